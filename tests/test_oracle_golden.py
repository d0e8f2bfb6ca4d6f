"""CPU: the C restatement (oracle/) against the fixtures recorded from the live reference.
This is the pin of the oracle: the reference ships no golden vectors, so the
fixtures under tests/golden/ -- outputs of the unmodified reference classes -- are the authority."""
import numpy as np
import pytest

from helpers import OracleBackend, load_golden, replay

# name -> tolerance of the mixed metric.  0.0 = bit-exact (same glibc sin/cos as numpy, same operation order).
# The UAV envs cannot be bit-exact: numpy evaluates the vector pow/tanh of FNTSMC.py with AVX-512 SIMD kernels that
# differ from glibc by 1-3 ulp and np.dot/np.linalg.inv go through OpenBLAS.  One-step bar: 1e-12 (observed <= 1e-14).
CASES = {
    "cartpole": 0.0,
    "cartpole_gentle": 0.0,
    "cartpole_wide": 0.0,
    "cartpole_angleonly_env": 0.0,
    "cartpole_angleonly_ppo2": 0.0,
    "fas": 0.0, "fas_ppo2": 0.0, "fas_discrete": 0.0, "soi": 0.0, "soi_dppo2": 0.0, "ballbalancer": 1e-12, "twolink": 1e-12,
    "ugv_forward": 1e-12, "ugv_bidirectional": 1e-12, "ugvo": 1e-12, "ugvo_dppo2": 1e-12, "ugvo_edge": 1e-12,
    "uavr_hover_outer": 1e-12, "uavr_hover": 1e-12, "uavr_inner": 1e-12, "uavr_tracking": 1e-12,
    "uav_pos": 1e-12, "uav_pos_dis": 1e-12, "uav_pos_wide": 1e-12, "uav_pos_rp0": 1e-12, "uav_pos_crash": 1e-12, "uav_pos_edge": 1e-12,
    "uav_att": 1e-12, "uav_att_rand": 1e-12, "uav_att_edge": 1e-12,
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("resync", [False, True])
def test_oracle_matches_reference_fixture(name, resync, oracle_lib):
    g = load_golden(name)
    res = replay(g, OracleBackend(name, g["reward"].shape[1]), resync=resync, name=name)
    assert res["flag_mismatch"] == 0 and res["done_mismatch"] == 0, res
    assert res["worst"]["time"] == 0.0, res
    if resync or CASES[name] == 0.0:
        for k, v in res["worst"].items():
            assert v <= CASES[name], (k, res)
    else:
        # free-running: within 1e4 x the reference's own drift under one-ulp nudges (floor 1e-12), see helpers.replay
        assert res["worst_ratio"] <= 1.0, res


def test_random_pos0_reset_quirk_fixture_and_oracle(oracle_lib):
    """reset_uav_pos_ctrl(random_pos0=True), note N5: the reference loads p, q, r from the pos0 of the PREVIOUS reset
    (init_state = concat(pos0, vel0, angle0, pos0), uav.py:268).  Shown on the recorded fixture, then on the C
    restatement's own Philox resets: pos0 within 0.3 m of the trajectory start, p, q, r = previous pos0."""
    import numpy as np
    from oracle import oracle
    from reinforcementlearningplatform_b200 import _lib
    import reinforcementlearningplatform_b200 as rlp
    g = load_golden("uav_pos_rp0")
    T, L = g["reward"].shape
    seen = 0
    for l in range(L):
        prev_pos0 = g["state0"][l, 0:3]
        assert np.array_equal(g["state0"][l, 51:54], prev_pos0)
        for t in np.where(g["done"][:, l])[0]:
            rs = g["reset_state"][t, l]
            assert np.array_equal(rs[9:12], prev_pos0)          # p, q, r <- previous pos0
            assert np.array_equal(rs[51:54], rs[0:3])           # init_state[9:12] <- the new pos0
            traj0 = np.array([0.0, 0.0, 1.5]) + rs[33:36] * np.sin(rs[41:44])
            assert np.all(np.abs(rs[0:3] - traj0) <= 0.3 + 1e-12)
            prev_pos0 = rs[0:3].copy()
            seen += 1
    assert seen >= 3
    host = rlp.UavPosCtrlRL(n_envs=64, host_only=True, random_trajectory=True, random_pos0=True)
    sf, od, ad, dd = _lib.dims(_lib.UAV_POS, 1)
    assert sf == 54
    orc = oracle.OracleEnv(_lib.UAV_POS, host._params, 64, sf, od, ad, dd, seed=3)
    orc.reset()
    first = orc.state[0:3].copy()
    assert np.all(orc.state[9:12] == 0.0)
    traj0 = np.array([0.0, 0.0, 1.5])[:, None] + orc.state[33:36] * np.sin(orc.state[41:44])
    assert np.all(np.abs(first - traj0) <= 0.3 + 1e-12) and np.std(first) > 0.05
    orc.reset()
    assert np.array_equal(orc.state[9:12], first)
    assert np.array_equal(orc.state[51:54], orc.state[0:3]) and not np.array_equal(orc.state[0:3], first)


@pytest.mark.parametrize("name,h,both", [("cartpole", 0.002, False), ("cartpole_angleonly_env", 0.001, True),
                                         ("fas", 0.002, True), ("ballbalancer", 0.002, True)])
def test_time_loop_substep_trace(name, h, both, oracle_lib):
    """Note N1: `while self.time < tt` with h = dt / 10 takes 10 OR 11 RK4 sub-steps per control period,
    depending on how the float64 additions of h round.  The fixture's recorded `time` column gives the reference's count
    per step; the restatement must take exactly the same number on every step, and both counts occur (`both`: the
    CartPole fixture's episodes are too short for an 11)."""
    import numpy as np
    g = load_golden(name)
    T, L = g["reward"].shape
    be = OracleBackend(name, L)
    be.set_state(g["state0"], g["time0"])
    prev_time = g["time0"].copy()
    seen = set()
    for t in range(T):
        be.step(g["actions"][t])
        ref_count = np.rint((g["time"][t] - prev_time) / h).astype(int)
        assert np.array_equal(be.env.substeps, ref_count), (name, t, be.env.substeps, ref_count)
        seen.update(ref_count.tolist())
        prev_time = g["time"][t].copy()
        lanes = np.nonzero(g["done"][t])[0]
        if len(lanes):
            be.set_state(g["reset_state"][t][lanes], g["reset_time"][t][lanes], lanes)
            prev_time[lanes] = g["reset_time"][t][lanes]
    assert seen == ({10, 11} if both else {10}), seen


def test_oracle_ugvo_reset_draws_legal_maps(oracle_lib):
    """The engine-defined reset of UGVForwardObstacleAvoidance (Philox draws keyed by (seed, instance, episode); one block
    per obstacle candidate, the radius from the 22 bits the two 53-bit coordinates leave over -- csrc/ugvo.cu draw3,
    oracle/c/ugvo.c) against the rules of reset() :527-557 and Map.generate_circle_obs_training (map.py:120-174):
    clearances to start, target and between obstacles, radii in [r_min, r_max], radius resolution 2^-22 of the range;
    the draws depend on the global instance index only (not on how instances are split over calls)."""
    import reinforcementlearningplatform_b200 as rlp
    from oracle import oracle
    from reinforcementlearningplatform_b200 import _lib
    host = rlp.UGVForwardObstacleAvoidance(n_envs=1, variant='dppo2', host_only=True)
    p = host._params
    sf, od, ad, dd = _lib.dims(_lib.UGVO, host.VARIANT)
    n = 96
    orc = oracle.OracleEnv(_lib.UGVO, p, n, sf, od, ad, dd, seed=5)
    orc.reset()
    st = orc.state
    sx, sy, tx, ty, nobs = st[0], st[1], st[5], st[6], st[7].astype(int)
    assert np.all((sx >= p.st_margin) & (sx <= p.map_x - p.st_margin) & (sy >= p.st_margin) & (sy <= p.map_y - p.st_margin))
    assert np.all(np.hypot(tx - sx, ty - sy) >= p.safety_dis_st)
    assert nobs.min() >= 1 and nobs.max() <= p.obs_num and (nobs == p.obs_num).mean() > 0.2
    levels = set()
    for i in range(n):
        c = st[8:8 + 3 * nobs[i], i].reshape(-1, 3)
        cx, cy, r = c[:, 0], c[:, 1], c[:, 2]
        assert np.all((r >= p.r_min) & (r <= p.r_max)) and np.all((cx >= 0) & (cx <= p.map_x) & (cy >= 0) & (cy <= p.map_y))
        assert np.all(np.hypot(cx - sx[i], cy - sy[i]) > r + p.safety_dis_st)
        assert np.all(np.hypot(cx - tx[i], cy - ty[i]) > r + p.safety_dis_st)
        d = np.hypot(cx[:, None] - cx[None, :], cy[:, None] - cy[None, :]) + np.eye(len(r)) * 1e9
        assert np.all(d > r[:, None] + r[None, :] + p.safety_dis_obs)
        assert np.all(st[8 + 3 * nobs[i]:, i] == 0.0)                       # unused slots are zero
        u = (r - p.r_min) / (p.r_max - p.r_min) * 4194304.0                    # radius = lerp(r_min, r_max, k / 2^22)
        assert np.all(np.abs(u - np.round(u)) < 1e-6)
        levels.update(np.round(u).astype(int).tolist())
    assert len(levels) > 0.9 * nobs.sum()                                      # the 22 bits are actually used
    # instance i of an 96-instance call == instance 0 of a call offset by i
    for i in (0, 17, 95):
        one = oracle.OracleEnv(_lib.UGVO, p, 1, sf, od, ad, dd, seed=5, env_index_offset=i)
        one.reset()
        assert np.array_equal(one.state[:, 0], st[:, i])
