"""GPU, at the full sizes of BASELINE.json's configs (1,048,576 UAV / SOI instances, 65,536 CartPole, 262,144
UGV-FOA): size-independent properties of the step kernels plus an oracle check on a contiguous window of the batch.

 * sharding invariance: the whole batch in one launch == two half batches with `env_index_offset` (the multi-GPU
   contract: Philox resets are keyed by the global instance index), bit for bit, across auto-resets;
 * determinism: same seed, same actions -> same bits;
 * bookkeeping invariants: done == (flag != 0), flags from the env's code table, time == 0 and episode + 1 exactly where
   an auto-reset happened, time advanced by dt (or the RK4 time-loop's 10|11 h) elsewhere;
 * a window of 4096 consecutive instances taken from the middle of the batch, stepped by the C oracle from the same
   state with the same actions: within the engine-vs-oracle tolerance, flags exact.
"""
import numpy as np
import pytest

from helpers import ENGINE_TOL, mixed_err

pytestmark = pytest.mark.gpu

CASES = {  # name -> (constructor, kwargs, n, valid flags)
    "uav_pos": ("UavPosCtrlRL", {"random_trajectory": True}, 1 << 20, {0, 1, 2, 3}),
    "uav_att": ("UavAttCtrlRL", {"random_trajectory": True}, 1 << 20, {0, 1, 3}),
    "cartpole": ("CartPole", {}, 65536, {0, 1, 2, 3, 4}),
    "soi": ("SecondOrderIntegration", {}, 1 << 20, {0, 1, 2, 3}),
    "ugvo_dppo2": ("UGVForwardObstacleAvoidance", {"variant": "dppo2"}, 262144, {0, 1, 2, 3, 4}),
}
STEPS = 6


def _make(name, n, offset=0, seed=77):
    import torch
    import reinforcementlearningplatform_b200 as rlp
    cls, kw, _, _ = CASES[name]
    return getattr(rlp, cls)(n_envs=n, seed=seed, auto_reset=True, env_index_offset=offset, dtype=torch.float64, **kw)


def _actions(env, n, steps, seed):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    ar = torch.as_tensor(np.asarray(env.action_range, dtype=np.float64), device="cuda")
    lo, hi = ar[:, 0].view(1, -1, 1), ar[:, 1].view(1, -1, 1)
    return (lo + (hi - lo) * torch.rand((steps, ar.shape[0], n), generator=g, device="cuda", dtype=torch.float64)).contiguous()


def _short_episodes(env):
    """start most instances close to their time-out so that the few steps of the test cross many auto-resets"""
    import torch
    tmax = float(getattr(env, "timeMax", None) or env.time_max)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    env._time.copy_((tmax - 3 * env.dt) * (torch.rand(env.n_envs, generator=g, device="cuda") < 0.5).double())
    env._policy_obs_valid = False


@pytest.mark.parametrize("name", sorted(CASES))
def test_full_size_properties(name, oracle_lib):
    import torch
    from oracle import oracle
    from reinforcementlearningplatform_b200 import _lib
    _, _, n, valid = CASES[name]
    whole = _make(name, n)
    whole.reset(True)
    _short_episodes(whole)
    acts = _actions(whole, n, STEPS, 3)
    h = n // 2
    halves = [_make(name, h, 0), _make(name, n - h, h)]
    twin = _make(name, n)
    for e in halves + [twin]:
        e.reset(True)
        _short_episodes(e)
    # (the random time pattern above is drawn per env object: copy the whole batch's into the shards)
    halves[0]._time.copy_(whole._time[:h]); halves[1]._time.copy_(whole._time[h:]); twin._time.copy_(whole._time)
    assert torch.equal(torch.cat([halves[0]._state, halves[1]._state], 1), whole._state)  # Philox by global index
    sf, od, ad, dd = _lib.dims(whole.ENV_ID, whole.VARIANT)
    W, start = 4096, n // 2 - 1000   # the window straddles the shard boundary
    orc = oracle.OracleEnv(whole.ENV_ID, whole._params, W, sf, od, ad, dd, seed=77, env_index_offset=start,
                           auto_reset=True, nthreads=8)
    resets = 0
    worst = 0.0
    for t in range(STEPS):
        time_before = whole._time.clone()
        ep_before = whole._episode.clone()
        orc.state[:] = whole._state[:, start:start + W].cpu().numpy()
        orc.time[:] = whole._time[start:start + W].cpu().numpy()
        orc.episode[:] = whole._episode[start:start + W].cpu().numpy().astype(np.uint32)
        whole.step_soa(acts[t])
        twin.step_soa(acts[t])
        halves[0].step_soa(acts[t][:, :h].contiguous())
        halves[1].step_soa(acts[t][:, h:].contiguous())
        orc.step(acts[t][:, start:start + W].cpu().numpy())
        for f in ("_state", "_time", "_episode", "_next_obs", "_reward", "_done", "_flag", "_reset_obs"):
            a = getattr(whole, f)
            assert torch.equal(a, getattr(twin, f)), (name, t, f, "determinism")
            b = torch.cat([getattr(halves[0], f), getattr(halves[1], f)], dim=a.dim() - 1)
            assert torch.equal(a, b), (name, t, f, "sharding invariance")
        done, flag = whole._done.bool(), whole._flag
        assert torch.equal(done, flag != 0)
        assert set(torch.unique(flag).tolist()) <= valid, torch.unique(flag)
        assert torch.all(whole._time[done] == 0.0)
        assert torch.equal(whole._episode[done], ep_before[done] + 1) and torch.equal(whole._episode[~done], ep_before[~done])
        adv = (whole._time - time_before)[~done]
        assert float(adv.min()) > 0.98 * whole.dt and float(adv.max()) < 1.12 * whole.dt   # dt, or 10|11 sub-steps of dt/10
        assert torch.isfinite(whole._reward).all() and torch.isfinite(whole._next_obs).all()
        resets += int(done.sum())
        sl = slice(start, start + W)
        assert np.array_equal(whole._flag[sl].cpu().numpy(), orc.flag) and np.array_equal(whole._done[sl].cpu().numpy(), orc.done)
        for got, ref in ((whole._next_obs[:, sl], orc.next_obs), (whole._reward[sl], orc.reward), (whole._reset_obs[:, sl], orc.reset_obs),
                         (whole._time[sl], orc.time)):
            worst = max(worst, mixed_err(got.cpu().numpy(), ref))
    assert resets > n // 4, resets          # the auto-reset path ran at scale
    assert worst <= ENGINE_TOL[name], (name, worst)
