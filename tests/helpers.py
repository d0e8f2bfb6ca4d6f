"""Shared parity machinery: replay a golden fixture (recorded from the live reference by
oracle/gen_golden.py) through a backend -- the C oracle on the CPU or the CUDA engine -- and
report the worst deviation.  Metric: |a-b| <= tol * max(1, |b|)."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def mixed_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def env_specs():
    """golden fixture name -> (env class, constructor kwargs)."""
    import reinforcementlearningplatform_b200 as rlp
    return {
        "cartpole": (rlp.CartPole, {}),
        "cartpole_gentle": (rlp.CartPole, {}),
        "cartpole_wide": (rlp.CartPole, {}),
        "cartpole_angleonly_env": (rlp.CartPoleAngleOnly, {"variant": "env"}),
        "cartpole_angleonly_ppo2": (rlp.CartPoleAngleOnly, {"variant": "ppo2"}),
        "fas": (rlp.Flight_Attitude_Simulator, {}),
        "fas_ppo2": (rlp.Flight_Attitude_Simulator, {"variant": "ppo2"}),
        "fas_discrete": (rlp.FlightAttitudeSimulatorDiscrete, {}),
        "soi": (rlp.SecondOrderIntegration, {}),
        "soi_dppo2": (rlp.SecondOrderIntegration, {"variant": "dppo2"}),
        "ballbalancer": (rlp.BallBalancer1D, {}),
        "twolink": (rlp.TwoLinkManipulator, {}),
        "ugv_forward": (rlp.UGVForward, {}),
        "ugv_bidirectional": (rlp.UGVBidirectional, {}),
        "uavr_hover_outer": (rlp.uav_hover_outer_loop, {}),
        "uavr_hover": (rlp.uav_hover, {}),
        "uavr_inner": (rlp.uav_inner_loop, {}),
        "uavr_tracking": (rlp.uav_tracking_outer_loop, {}),
        "ugvo": (rlp.UGVForwardObstacleAvoidance, {}),
        "ugvo_dppo2": (rlp.UGVForwardObstacleAvoidance, {"variant": "dppo2"}),
        "ugvo_edge": (rlp.UGVForwardObstacleAvoidance, {}),
        "uav_pos": (rlp.UavPosCtrlRL, {"random_trajectory": True}),
        "uav_pos_dis": (rlp.UavPosCtrlRL, {"random_trajectory": True}),
        "uav_pos_wide": (rlp.UavPosCtrlRL, {"random_trajectory": True}),
        "uav_pos_rp0": (rlp.UavPosCtrlRL, {"random_trajectory": True, "random_pos0": True}),
        "uav_pos_crash": (rlp.UavPosCtrlRL, {"random_trajectory": True}),
        "uav_pos_edge": (rlp.UavPosCtrlRL, {"random_trajectory": True}),
        "uav_att": (rlp.UavAttCtrlRL, {"random_trajectory": False}),
        "uav_att_rand": (rlp.UavAttCtrlRL, {"random_trajectory": True}),
        "uav_att_edge": (rlp.UavAttCtrlRL, {"random_trajectory": True}),
    }


class OracleBackend:
    """C restatement (oracle/liboracle.so) behind the replay interface."""

    def __init__(self, name, lanes):
        from oracle import oracle
        from reinforcementlearningplatform_b200 import _lib
        cls, kw = env_specs()[name]
        host = cls(n_envs=lanes, host_only=True, **kw)
        sf, od, ad, dd = _lib.dims(cls.ENV_ID, host.VARIANT)
        self.env = oracle.OracleEnv(cls.ENV_ID, host._params, lanes, sf, od, ad, dd)

    def set_state(self, state, time, lanes=None):
        sel = slice(None) if lanes is None else lanes
        self.env.state[:, sel] = np.asarray(state).T
        self.env.time[sel] = time

    def step(self, actions, dis=None):
        e = self.env
        e.step(np.ascontiguousarray(actions.T), None if dis is None else np.ascontiguousarray(dis.T))
        return dict(obs=e.obs.T.copy(), next_obs=e.next_obs.T.copy(), reward=e.reward.copy(), done=e.done.copy(),
                    flag=e.flag.copy(), state=e.state.T.copy(), time=e.time.copy())


class EngineBackend:
    """CUDA engine (libb200env.so through the VecEnv mirror) behind the replay interface."""

    def __init__(self, name, lanes, dtype=None):
        import torch
        cls, kw = env_specs()[name]
        self.torch = torch
        self.env = cls(n_envs=lanes, device="cuda", dtype=dtype or torch.float64, **kw)

    def set_state(self, state, time, lanes=None):
        t = self.torch
        e = self.env
        st = e._state.cpu().numpy().astype(np.float64)
        tm = e._time.cpu().numpy()
        sel = slice(None) if lanes is None else lanes
        st[:, sel] = np.asarray(state).T
        tm[sel] = time
        e.set_state_buffers(t.from_numpy(st), t.from_numpy(tm))

    def step(self, actions, dis=None):
        t = self.torch
        e = self.env
        e.step_update(t.from_numpy(np.ascontiguousarray(actions)), None if dis is None else t.from_numpy(np.ascontiguousarray(dis)),
                      layout="envs_first")   # fixtures are [lanes, dim]; explicit because some have lanes == dim
        t.cuda.synchronize()
        f = lambda x: x.detach().cpu().numpy().astype(np.float64)
        return dict(obs=f(e.current_state), next_obs=f(e.next_state), reward=f(e.reward),
                    done=e._done.cpu().numpy().copy(), flag=e._flag.cpu().numpy().copy(),
                    state=f(e._state.t()), time=e._time.cpu().numpy().copy())


# leading state fields that are compared; the trailing ones are "lazy" (only written on terminal steps, see
# include/b200env.h: ref/dot_ref of the attitude env, pos_ref/dot_pos_ref of the position env)
STATE_CMP = {"uav_pos": 45, "uav_att": 30}
# chaotic plants (random torques on a double pendulum): compare one step at a time, re-syncing the engine from the oracle
RESYNC_EVERY_STEP = {"twolink"}


def _cmp_fields(name):
    for k, v in STATE_CMP.items():
        if name.startswith(k):
            return v
    return None


def replay(g, backend, resync=False, steps=None, name="", sens_k=None, floor=1e-12, chaos_cut=1e-11):
    """Run the fixture's actions through `backend`.  Free-running: state carried by the backend, re-injected
    only after the reference's resets.  resync=True: the fixture's state is injected before every step.

    Besides the raw worst mixed errors the result carries `worst_ratio`: the worst error divided by the per-lane
    tolerance max(floor, sens_k * running max of the fixture's twin_err), i.e. relative to how far the reference
    drifts from ITSELF when nudged by one ulp per step (its own sensitivity to rounding).  Once that self-drift of a
    lane exceeds `chaos_cut` the episode is no longer reproducible even by the reference (chaotic two-link arm,
    open-loop-unstable cart-pole): the lane is skipped until its next reset, where the state is re-injected."""
    T, L = g["reward"].shape
    if steps:
        T = min(T, steps)
    sens_k = sens_k or (1.0e6 if name in RESYNC_EVERY_STEP else 1.0e4)  # chaotic plants: errors grow like e^(lambda t)
    has_dis = "dis" in g
    nf = _cmp_fields(name)
    backend.set_state(g["state0"], g["time0"])
    worst = dict(obs=0.0, next_obs=0.0, reward=0.0, state=0.0, time=0.0)
    flag_mismatch = 0
    done_mismatch = 0
    first_bad = None
    run_sens = np.zeros(L)
    worst_ratio = 0.0
    lane_state = np.zeros(L)   # worst mixed state error per lane (drift statistics over seeds, tests/test_drift_gpu.py)
    live_steps, all_steps = 0, 0
    lane_err = lambda a, b: np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)), axis=1) if a.ndim == 2 else np.abs(a - b) / np.maximum(1.0, np.abs(b))
    for t in range(T):
        if resync and t > 0:
            st = np.where(np.isnan(g["reset_state"][t - 1]), g["state"][t - 1], g["reset_state"][t - 1])
            tm = np.where(np.isnan(g["reset_time"][t - 1]), g["time"][t - 1], g["reset_time"][t - 1])
            backend.set_state(st, tm)
        out = backend.step(g["actions"][t], g["dis"][t] if has_dis else None)
        if "twin_err" in g:
            run_sens = np.maximum(run_sens, g["twin_err"][t])
        tol = np.maximum(floor, sens_k * run_sens)
        live = (run_sens <= chaos_cut) | resync
        live_steps += int(np.sum(live))
        all_steps += L
        for k in ("obs", "next_obs", "reward", "state"):
            a, b = out[k], g[k][t]
            if k == "state" and nf:
                a, b = a[:, :nf], b[:, :nf]
            if k == "state":
                lane_state = np.maximum(lane_state, lane_err(np.asarray(a, np.float64), np.asarray(b, np.float64)))
            if live.any():
                worst[k] = max(worst[k], mixed_err(np.asarray(a)[live], np.asarray(b)[live]))
                if k != "obs":
                    worst_ratio = max(worst_ratio, float(np.max((lane_err(np.asarray(a, np.float64), b) / tol)[live])))
        worst["time"] = max(worst["time"], float(np.max(np.abs(out["time"] - g["time"][t]))))
        fm = int(np.sum((out["flag"] != g["flag"][t])[live]))
        dm = int(np.sum((out["done"] != g["done"][t])[live]))
        if (fm or dm) and first_bad is None:
            first_bad = (t, out["flag"].tolist(), g["flag"][t].tolist())
        flag_mismatch += fm
        done_mismatch += dm
        lanes = np.nonzero(g["done"][t])[0]
        if len(lanes):
            run_sens[lanes] = 0.0
            if not resync:
                backend.set_state(g["reset_state"][t][lanes], g["reset_time"][t][lanes], lanes)
    return dict(worst=worst, worst_ratio=worst_ratio, flag_mismatch=flag_mismatch, done_mismatch=done_mismatch,
                first_bad=first_bad, steps=T, lane_state=lane_state, live_fraction=live_steps / max(1, all_steps))


# ---------------------------------------------------------------------------------------------
# engine vs oracle on seeded random inputs, both with the in-kernel Philox auto-reset
# ---------------------------------------------------------------------------------------------
def action_bounds(name, env):
    ar = np.asarray(env.action_range, dtype=np.float64)
    return ar[:, 0], ar[:, 1]


def smoke_cases():
    return ["cartpole", "uav_att", "uav_pos"]


# fp64 tolerance (mixed metric |a - b| / max(1, |b|)) of (i) the free-running replay of the reference fixtures and (ii) the
# engine-vs-oracle comparison on seeded random inputs (8192 instances x 60 steps with Philox auto-resets), per fixture.
# About 10x the larger of the two values measured on B200 (tools/tolerance_report.py, tests/parity_report.py; round 2:
# cartpole 1.3e-13, fas / soi <= 7e-15, ballbalancer 1.6e-13, ugv 7e-14, uav_att 1.7e-14, uav_pos 3.8e-9 on the lane whose
# gains are redrawn every step -- profiles/r2/drift.md -- and 1.5e-11 on the wide / dis fixtures, uavr 7e-13, ugvo laser
# ranges 4.3e-10: the reference's line / circle intersection is ill-conditioned on steep rays, test_engine_gpu.py).
# twolink: chaotic arm, 1.7e-8 over the stretches the reference itself reproduces.
ENGINE_TOL = {
    "cartpole": 2e-12, "cartpole_gentle": 2e-11, "cartpole_wide": 2e-12, "cartpole_angleonly_env": 2e-12, "cartpole_angleonly_ppo2": 2e-12,
    "uav_pos": 5e-8, "uav_pos_dis": 2e-10, "uav_pos_wide": 2e-10, "uav_pos_rp0": 5e-10, "uav_pos_crash": 5e-11, "uav_pos_edge": 5e-11,
    "uav_att": 2e-13, "uav_att_rand": 2e-13, "uav_att_edge": 2e-13,
    "fas": 1e-13, "fas_ppo2": 1e-13, "fas_discrete": 1e-13, "soi": 1e-13, "soi_dppo2": 1e-13, "ballbalancer": 2e-12, "twolink": 2e-7,
    "ugv_forward": 1e-12, "ugv_bidirectional": 1e-12, "ugvo": 2e-9, "ugvo_dppo2": 2e-9, "ugvo_edge": 2e-9,
    "uavr_hover_outer": 2e-12, "uavr_hover": 1e-11, "uavr_inner": 2e-12, "uavr_tracking": 2e-12,
}
# Share of (step, lane) samples of a free-running replay that must stay comparable (lanes are dropped until their next
# reset once the reference's own one-ulp drift exceeds 1e-11).  A fixture that falls below its share fails instead of
# passing on an empty comparison.
MIN_LIVE = {"twolink": 0.5, "ugvo_dppo2": 0.4}
MIN_LIVE_DEFAULT = 0.95


def engine_vs_oracle(name, n, steps, seed, dtype=None, auto_reset=True, tol=None, offset=0, io_dtype=None):
    """Step the CUDA engine and the C oracle side by side from the same Philox reset with the same random actions.
    Auto-reset uses the same counter-based draws on both sides, so trajectories stay comparable across episodes.
    Returns the worst mixed error over obs/next_obs/reward/state and the number of flag mismatches."""
    import torch
    from oracle import oracle
    from reinforcementlearningplatform_b200 import _lib
    cls, kw = env_specs()[name]
    dtype = dtype or torch.float64
    env = cls(n_envs=n, device="cuda", dtype=dtype, seed=seed, auto_reset=auto_reset, env_index_offset=offset,
              io_dtype=io_dtype, **kw)
    io_dt = io_dtype or dtype
    # float32 I/O with fp64 arithmetic: the state must still meet the fp64 tolerance; the RL-facing outputs are the
    # fp64 values rounded once to float32 (half an ulp = 2^-24 relative)
    worst_io = 0.0
    sf, od, ad, dd = _lib.dims(cls.ENV_ID, env.VARIANT)
    orc = oracle.OracleEnv(cls.ENV_ID, env._params, n, sf, od, ad, dd, seed=seed, auto_reset=auto_reset, nthreads=8,
                           env_index_offset=offset)
    env.reset(True)
    orc.reset()
    torch.cuda.synchronize()
    f = lambda x: x.detach().cpu().numpy().astype(np.float64)
    nf = _cmp_fields(name)
    cs = (lambda a: a[:nf]) if nf else (lambda a: a)
    worst = mixed_err(cs(f(env._state)), cs(orc.state))
    lo, hi = action_bounds(name, env)
    rng = np.random.default_rng(seed)
    flag_mismatch = 0
    n_done = 0
    for t in range(steps):
        a = rng.uniform(lo[:, None], hi[:, None], size=(ad, n))
        d = rng.normal(0, 0.5, size=(dd, n)) if dd else None
        a_dev = torch.from_numpy(a).to("cuda", io_dt)
        d_dev = None if d is None else torch.from_numpy(d).to("cuda", io_dt)
        a_cpu = f(a_dev)  # the oracle sees exactly the values the engine sees (matters in fp32 mode)
        env.step_soa(a_dev, d_dev)
        orc.step(a_cpu, None if d is None else f(d_dev))
        torch.cuda.synchronize()
        sane0 = np.all(np.isfinite(orc.state) & (np.abs(orc.state) < 1e6), axis=0)
        fm = int(np.sum((env._flag.cpu().numpy() != orc.flag)[sane0]) + np.sum((env._done.cpu().numpy() != orc.done)[sane0]))
        flag_mismatch += fm
        n_done += int(orc.done.sum())
        # lanes whose reference state has blown up numerically (e.g. the two-link arm under sustained random torques
        # reaches |omega| ~ 1e18) carry no information: only sane lanes are compared
        sane = np.all(np.isfinite(orc.state) & (np.abs(orc.state) < 1e6), axis=0)
        for got, ref in ((env._obs, orc.obs), (env._next_obs, orc.next_obs), (env._reward, orc.reward),
                         (cs(env._state), cs(orc.state)), (env._reset_obs, orc.reset_obs)):
            gg, rr = f(got), np.asarray(ref)
            if io_dt != dtype and got.dtype == io_dt:
                worst_io = max(worst_io, mixed_err(gg[..., sane], rr[..., sane]))
            else:
                worst = max(worst, mixed_err(gg[..., sane], rr[..., sane]))
        worst = max(worst, float(np.max(np.abs(env._time.cpu().numpy() - orc.time))))
        if fm or name in RESYNC_EVERY_STEP:  # (after a flag mismatch trajectories diverge) re-sync the engine from the oracle
            env.set_state_buffers(torch.from_numpy(orc.state), torch.from_numpy(orc.time),
                                  torch.from_numpy(orc.episode.astype(np.int32)))
    return dict(worst=worst, worst_io=worst_io, flag_mismatch=flag_mismatch, terminals=n_done,
                tol=tol if tol is not None else ENGINE_TOL[name])


# ---------------------------------------------------------------------------------------------
# fp32 mode: stated per-env tolerances over short horizons (BASELINE north_star)
# ---------------------------------------------------------------------------------------------
# (one-step tolerance, 100-step free-running tolerance or None) on the mixed error |a-b| / max(1, |b|) of what the RL
# side sees (next_state and reward), against the fp64 reference fixtures; measured on B200 with
# tools/f32_tolerance_report.py and rounded up 2-4x.  The bound applies to the 99.9 % quantile over (step, lane)
# samples (on these fixtures that is within 10 % of the maximum).  None = no free-running statement: open-loop
# unstable (cart-pole replay) or chaotic (two-link arm) plants amplify any 1e-7 difference exponentially.
# One fp32-specific change was needed to get here: cos(arcsin(u)) in the thrust/attitude allocation is evaluated as
# sqrt(1 - u^2) in fp32 (uav_common.cuh cos_of_asin) because cosf(fl32(pi/2)) is negative where cos(fl64(pi/2)) is
# positive, which mirrored theta_d whenever the allocation saturated.
FP32_TOL = {
    "cartpole": (2e-6, 1e-5), "cartpole_gentle": (2e-6, None), "cartpole_angleonly_env": (2e-6, 1e-5),
    "cartpole_angleonly_ppo2": (2e-6, 1e-5),
    "fas": (2e-6, 1e-5), "fas_ppo2": (2e-6, 1e-5), "fas_discrete": (2e-6, 2e-5), "soi": (1e-6, 5e-6), "soi_dppo2": (1e-6, 5e-6),
    "ballbalancer": (4e-6, 3e-4), "twolink": (1e-5, None),
    "ugv_forward": (4e-5, 4e-5), "ugv_bidirectional": (2e-5, 2e-5),
    "uav_att": (3e-6, 5e-6), "uav_att_rand": (3e-6, 5e-6), "uav_att_edge": (3e-6, 5e-6),
    "uav_pos": (2e-5, 6e-5), "uav_pos_dis": (2e-5, 6e-5), "uav_pos_crash": (2e-6, 5e-6), "uav_pos_edge": (1e-4, 2e-4),
    "uavr_hover_outer": (2e-6, 5e-6), "uavr_hover": (3e-5, 6e-5), "uavr_inner": (6e-6, 1e-5), "uavr_tracking": (1e-5, 2e-5),
    # fake laser: the reference intersects in slope form (m = tan(phi), up to 1e7 in fp32); rays within ~1e-3 rad of
    # the vertical lose most of their digits, so the bound is on the 99 % quantile (max observed 7e-2 of 2.0 m range)
    "ugvo": (2e-3, 4e-3), "ugvo_dppo2": (2e-3, 4e-3),
}
FP32_QUANTILE = {"ugvo": 0.99, "ugvo_dppo2": 0.99}  # default 0.999


def fp32_replay(g, backend, steps, resync):
    """Replay `steps` steps of a fixture through the fp32 engine; returns the per-(step, lane) mixed error of
    (next_obs, reward) and the number of flag mismatches.  resync: inject the reference state before every step."""
    T, L = g["reward"].shape
    T = min(T, steps)
    has_dis = "dis" in g
    backend.set_state(g["state0"], g["time0"])
    errs = np.zeros((T, L))
    flag_mismatch = 0
    for t in range(T):
        if resync and t > 0:
            st = np.where(np.isnan(g["reset_state"][t - 1]), g["state"][t - 1], g["reset_state"][t - 1])
            tm = np.where(np.isnan(g["reset_time"][t - 1]), g["time"][t - 1], g["reset_time"][t - 1])
            backend.set_state(st, tm)
        out = backend.step(g["actions"][t], g["dis"][t] if has_dis else None)
        e_obs = np.max(np.abs(out["next_obs"] - g["next_obs"][t]) / np.maximum(1.0, np.abs(g["next_obs"][t])), axis=1)
        e_rew = np.abs(out["reward"] - g["reward"][t]) / np.maximum(1.0, np.abs(g["reward"][t]))
        errs[t] = np.maximum(e_obs, e_rew)
        flag_mismatch += int(np.sum(out["flag"] != g["flag"][t]))
        lanes = np.nonzero(g["done"][t])[0]
        if len(lanes) and not resync:
            backend.set_state(g["reset_state"][t][lanes], g["reset_time"][t][lanes], lanes)
    return errs, flag_mismatch
