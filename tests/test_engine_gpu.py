"""GPU: the CUDA engine (through the C ABI) against (1) the fixtures recorded from the live reference and
(2) the C oracle on seeded random inputs at larger sizes.  fp64 bar: 1e-12 mixed error per step in one-step
re-sync mode; free-running tolerances are stated per env in helpers.ENGINE_TOL."""
import numpy as np
import pytest

from helpers import (ENGINE_TOL, FP32_QUANTILE, FP32_TOL, MIN_LIVE, MIN_LIVE_DEFAULT, EngineBackend, engine_vs_oracle, env_specs,
                     fp32_replay, load_golden, replay)

pytestmark = pytest.mark.gpu

GOLDEN_CASES = sorted(ENGINE_TOL)

# One-step bar is 1e-12 except for the fake laser: the reference's line/circle intersection
# (UGVForwardObstacleAvoidance.py:360-364) works with the slope m = tan(phi), |m| up to 1e16 on near-vertical rays, and
# loses ~eps * m^2 there, so a 1-ulp difference between numpy's and CUDA's tan() moves a range (the reference itself
# moves as much under a 1-ulp nudge: fixture twin_err).  Measured on B200: 1.3e-12 (ugvo), 5.1e-11 (ugvo_dppo2, 8 lanes x
# 400 steps x 37 rays); the bounds are 10x that.  Kinematic state, reward and flags stay exact.
ONE_STEP_TOL = {"ugvo": 2e-11, "ugvo_dppo2": 5e-10, "ugvo_edge": 2e-11}

# Fixtures whose actions were recorded from a closed loop around an open-loop-unstable plant: replaying them
# open loop amplifies a 1-ulp difference by e^(lambda*t) (inverted pendulum: lambda ~ 6/s, 5 s episodes -> 1e13),
# so only the one-step (re-sync) comparison is meaningful.  The C oracle still matches them bit-exactly.
OPEN_LOOP_UNSTABLE = {"cartpole_gentle": "open-loop replay of closed-loop actions on an unstable plant (e^30 gain)",
                      # the *_edge fixtures put every lane within one ulp of a zone edge / terminal threshold: their point is
                      # the bit-exact flag in one-step mode; free-running, the reference's own one-ulp twin crosses the
                      # edge at a different step in every lane (live fraction 0), so there is nothing to compare
                      "uav_att_edge": "every lane sits on a terminal threshold: one-step (re-sync) comparison only",
                      "uav_pos_edge": "every lane sits on a terminal threshold: one-step (re-sync) comparison only",
                      # collision_check's `distance <= r + r_vehicle` at equality and one ulp either side (UGVForward
                      # ObstacleAvoidance.py:261-272): the vehicle is parked there after every reset
                      "ugvo_edge": "every lane sits on the collision radius: one-step (re-sync) comparison only"}


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_engine_matches_reference_fixture_resync(name):
    """one-step mode: inject the reference state before every step; every output within 1e-12, flags exact."""
    g = load_golden(name)
    res = replay(g, EngineBackend(name, g["reward"].shape[1]), resync=True, name=name)
    assert res["flag_mismatch"] == 0 and res["done_mismatch"] == 0, res
    assert res["worst"]["time"] == 0.0, res
    for k, v in res["worst"].items():
        assert v <= ONE_STEP_TOL.get(name, 1e-12), (k, res)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_engine_matches_reference_fixture_free_running(name):
    """free-running over the whole fixture (1000 steps), state re-injected only after the reference's resets."""
    if name in OPEN_LOOP_UNSTABLE:
        pytest.skip(OPEN_LOOP_UNSTABLE[name])
    g = load_golden(name)
    res = replay(g, EngineBackend(name, g["reward"].shape[1]), resync=False, name=name)
    assert res["flag_mismatch"] == 0 and res["done_mismatch"] == 0, res
    assert res["worst"]["time"] == 0.0, res
    assert res["live_fraction"] >= MIN_LIVE.get(name, MIN_LIVE_DEFAULT), (name, res["live_fraction"])
    # within 1e4 x the reference's own drift under one-ulp nudges (floor 1e-12), and an absolute ceiling
    assert res["worst_ratio"] <= 1.0, res
    for k, v in res["worst"].items():
        assert v <= ENGINE_TOL[name], (k, res)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_engine_vs_oracle_random(name, oracle_lib):
    res = engine_vs_oracle(name, n=8192, steps=60, seed=7)
    assert res["flag_mismatch"] == 0, res
    assert res["worst"] <= res["tol"], res
    if name.startswith("cartpole") and name != "cartpole_gentle":
        assert res["terminals"] > 0, res  # episodes are short: the Philox auto-reset path is exercised


@pytest.mark.parametrize("name", ["cartpole", "uav_pos", "ugvo_dppo2"])
def test_engine_vs_oracle_sharded_offset(name, oracle_lib):
    """a shard that starts at a non-zero global instance index draws the resets of THOSE instances (multi-GPU contract)"""
    res = engine_vs_oracle(name, n=1000, steps=40, seed=11, offset=(1 << 33) + 12345)
    assert res["flag_mismatch"] == 0, res
    assert res["worst"] <= res["tol"], res


@pytest.mark.parametrize("name", ["cartpole", "soi", "fas_ppo2", "ugv_forward", "ugvo_dppo2", "uav_att_rand", "uav_pos",
                                  "uavr_hover", "twolink", "ballbalancer"])
def test_engine_f32_io_keeps_fp64_trajectory(name, oracle_lib):
    """b200env_io.io_dtype = F32 with dtype = F64: float32 actions in, float32 obs/reward out, state and arithmetic in
    fp64.  The state must meet the same fp64 tolerance as the all-fp64 engine; the float32 outputs are the oracle's fp64
    values rounded once (2^-24 relative + the fp64 tolerance)."""
    import torch
    res = engine_vs_oracle(name, n=4096, steps=40, seed=5, io_dtype=torch.float32)
    assert res["flag_mismatch"] == 0, res
    assert res["worst"] <= res["tol"], res
    assert 0.0 < res["worst_io"] <= 2.0 ** -24 + res["tol"], res


@pytest.mark.parametrize("name", sorted(FP32_TOL))
def test_engine_fp32_stated_tolerance(name):
    """fp32 mode (state, arithmetic and I/O in float32; `time` stays float64) against the fp64 reference fixtures:
    one-step (200 steps, state re-injected) and 100-step free-running errors of next_state and reward within the
    per-env tolerance stated in helpers.FP32_TOL; terminal flags identical on these fixtures."""
    import torch
    g = load_golden(name)
    L = g["reward"].shape[1]
    one_tol, free_tol = FP32_TOL[name]
    q = FP32_QUANTILE.get(name, 0.999)
    errs, fm = fp32_replay(g, EngineBackend(name, L, dtype=torch.float32), steps=200, resync=True)
    assert np.quantile(errs, q) <= one_tol, (name, float(np.quantile(errs, q)), float(errs.max()))
    assert fm == 0, (name, fm)
    assert np.median(errs) <= 5e-6, (name, float(np.median(errs)))  # typical error: a few fp32 roundings
    if free_tol is not None:
        errs, fm = fp32_replay(g, EngineBackend(name, L, dtype=torch.float32), steps=100, resync=False)
        assert np.quantile(errs, q) <= free_tol, (name, float(np.quantile(errs, q)), float(errs.max()))
        assert fm == 0, (name, fm)
