"""GPU: device-resident rollout buffer (rollout.py, vector form of utils/classes.py:250-311) filled directly by the
step kernels, and K-GAE over it (gae_flags) -- against the C oracle stepped side by side and the float-mask GAE."""
import numpy as np
import pytest

from helpers import env_specs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["soi", "fas_ppo2", "cartpole", "ugvo_dppo2", "uav_pos"])
def test_rollout_rows_equal_oracle_transitions(name, oracle_lib):
    """every row t of the buffer holds the transition (s, a, r, s', done, flag) the oracle produces at step t, rounded
    once to float32; GAE over the buffer equals the reference scan over the same float32 columns bit for bit."""
    import torch
    from oracle import oracle
    from reinforcementlearningplatform_b200 import RolloutBuffer, _lib
    from reinforcementlearningplatform_b200 import gae as G
    cls, kw = env_specs()[name]
    n, T, seed = 2048, 24, 9
    env = cls(n_envs=n, device="cuda", dtype=torch.float64, io_dtype=torch.float32, seed=seed, auto_reset=True, **kw)
    sf, od, ad, dd = _lib.dims(cls.ENV_ID, env.VARIANT)
    orc = oracle.OracleEnv(cls.ENV_ID, env._params, n, sf, od, ad, dd, seed=seed, auto_reset=True, nthreads=8)
    env.reset(True)
    orc.reset()
    buf = RolloutBuffer(T, env)
    ar = np.asarray(env.action_range, dtype=np.float64)
    rng = np.random.default_rng(seed)
    tol = 2.0 ** -24 + 1e-9
    mixed = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0
    for t in range(T):
        a32 = rng.uniform(ar[:, :1], ar[:, 1:], size=(ad, n)).astype(np.float32)
        buf.a[t].copy_(torch.from_numpy(a32))
        buf.step(env, t, buf.a[t])
        orc.step(a32.astype(np.float64))
        torch.cuda.synchronize()
        assert np.array_equal(buf.done[t].cpu().numpy(), orc.done), (name, t)
        assert np.array_equal(buf.flag[t].cpu().numpy(), orc.flag), (name, t)
        assert mixed(buf.s[t].cpu().numpy().astype(np.float64), orc.obs) <= tol, (name, t, "s")
        assert mixed(buf.s_[t].cpu().numpy().astype(np.float64), orc.next_obs) <= tol, (name, t, "s_")
        assert mixed(buf.r[t].cpu().numpy().astype(np.float64), orc.reward) <= tol, (name, t, "r")
        assert mixed(env.policy_state.t().cpu().numpy().astype(np.float64), orc.reset_obs) <= tol, (name, t, "policy")
    assert buf.index == T
    # GAE straight from the buffer == float-mask GAE == sequential restatement, bit for bit
    g = torch.Generator(device="cuda").manual_seed(1)
    vs = torch.randn((T, n), generator=g, device="cuda")
    vsn = torch.randn((T, n), generator=g, device="cuda")
    adv, vt = buf.gae(vs, vsn, 0.99, 0.95, normalize=False)
    done_f, succ_f = buf.done.float(), buf.success()
    adv2, vt2, _ = G.gae(buf.r, vs, vsn, done_f, succ_f, 0.99, 0.95)
    assert torch.equal(adv, adv2) and torch.equal(vt, vt2)
    adv_o, vt_o, _ = oracle.gae(buf.r.cpu().numpy(), vs.cpu().numpy(), vsn.cpu().numpy(), done_f.cpu().numpy(),
                                succ_f.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), adv_o) and np.array_equal(vt.cpu().numpy(), vt_o)
    # success rule of the train loops: terminal for a reason other than the time-out
    s_np = succ_f.cpu().numpy()
    assert np.array_equal(s_np, ((buf.done.cpu().numpy() != 0) & (buf.flag.cpu().numpy() != env.TIMEOUT_FLAG)).astype(np.float32))
    s, a, a_lp, r, s_, done, success = buf.to_tensor()
    assert s.shape == (T * n, env.state_dim) and a.shape == (T * n, env.action_dim) and r.shape == (T * n, 1)
    assert torch.equal(s[n + 3], buf.s[1, :, 3])


@pytest.mark.parametrize("name", ["soi", "fas_ppo2", "fas_discrete", "ugv_forward", "ballbalancer", "twolink", "cartpole",
                                  "cartpole_angleonly_env", "cartpole_angleonly_ppo2", "uav_att_rand"])
def test_rollout_collect_equals_step_by_step(name):
    """b200env_rollout (fused multi-step kernel for the generic families and CartPole, per-step launches for the others) fills the
    buffer with exactly the bits that T separate step calls produce, and leaves the env in the same state."""
    import torch
    from reinforcementlearningplatform_b200 import RolloutBuffer
    cls, kw = env_specs()[name]
    n, T, seed = 3000, 40, 4
    mk = lambda: cls(n_envs=n, device="cuda", dtype=torch.float64, io_dtype=torch.float32, seed=seed, auto_reset=True, **kw)
    e1, e2 = mk(), mk()
    e1.reset(True)
    e2.reset(True)
    b1, b2 = RolloutBuffer(T, e1), RolloutBuffer(T, e2)
    ar = torch.as_tensor(np.asarray(e1.action_range, dtype=np.float64), device="cuda", dtype=torch.float32)
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = ar[:, :1].view(1, -1, 1) + (ar[:, 1:] - ar[:, :1]).view(1, -1, 1) * torch.rand(b1.a.shape, generator=g, device="cuda")
    b1.a.copy_(a)
    b2.a.copy_(a)
    for t in range(T):
        b1.step(e1, t, b1.a[t])
    b2.collect(e2)
    torch.cuda.synchronize()
    for f in ("s", "s_", "r", "done", "flag"):
        assert torch.equal(getattr(b1, f), getattr(b2, f)), (name, f)
    assert torch.equal(e1._state, e2._state) and torch.equal(e1._time, e2._time) and torch.equal(e1._episode, e2._episode)
    assert torch.equal(e1.policy_state, e2.policy_state)
    assert int(b1.done.sum()) > 0 or name in ("uav_att_rand", "twolink", "ballbalancer", "ugv_forward", "soi", "fas_discrete")


@pytest.mark.parametrize("name", ["uav_pos", "uav_att_rand", "soi", "ugvo_dppo2", "cartpole"])
def test_chained_policy_observations_equal_the_copied_ones(name):
    """RolloutBuffer.step(chain_policy_obs=True) -- every step writes its policy-facing observation straight into the next
    row of buf.s, current_state is not stored -- fills the buffer with the same bits as the per-step row copy
    (store_policy_obs=True), and leaves env.policy_state and the persistent state identical."""
    import torch
    from reinforcementlearningplatform_b200 import RolloutBuffer
    cls, kw = env_specs()[name]
    n, T, seed = 3000, 24, 6
    mk = lambda: cls(n_envs=n, device="cuda", dtype=torch.float64, io_dtype=torch.float32, seed=seed, auto_reset=True, **kw)
    e1, e2 = mk(), mk()
    e1.reset(True)
    e2.reset(True)
    b1, b2 = RolloutBuffer(T, e1), RolloutBuffer(T, e2)
    ar = torch.as_tensor(np.asarray(e1.action_range, dtype=np.float64), device="cuda", dtype=torch.float32)
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = ar[:, :1].view(1, -1, 1) + (ar[:, 1:] - ar[:, :1]).view(1, -1, 1) * torch.rand(b1.a.shape, generator=g, device="cuda")
    b1.a.copy_(a)
    b2.a.copy_(a)
    b2.s[0].copy_(e2.policy_state.t())
    for t in range(T):
        b1.step(e1, t, b1.a[t], store_policy_obs=True)
        b2.step(e2, t, b2.a[t], chain_policy_obs=True)
    torch.cuda.synchronize()
    for f in ("s", "s_", "r", "done", "flag"):
        assert torch.equal(getattr(b1, f), getattr(b2, f)), (name, f)
    assert torch.equal(e1._state, e2._state) and torch.equal(e1._time, e2._time) and torch.equal(e1._episode, e2._episode)
    assert torch.equal(e1.policy_state, e2.policy_state)
