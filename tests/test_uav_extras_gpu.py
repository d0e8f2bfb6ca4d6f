"""GPU: the UAV env methods the reference scripts call besides step_update: get_param_from_actor on its
own, the fixed-gain control step, current/next_state_norm with save/load in the reference's CSV layout."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["pos", "att"])
def test_get_param_then_fixed_gain_step_equals_fused_step(kind):
    import torch
    import reinforcementlearningplatform_b200 as rlp
    cls = rlp.UavPosCtrlRL if kind == "pos" else rlp.UavAttCtrlRL
    n = 4096
    a_env, b_env = cls(n_envs=n, random_trajectory=True, seed=5), cls(n_envs=n, random_trajectory=True, seed=5)
    a_env.reset(True)
    b_env.reset(True)
    rng = np.random.default_rng(1)
    hi = 5.0 if kind == "pos" else 3.0
    for t in range(20):
        act = rng.uniform(0, hi, (8, n))
        act[rng.random((8, n)) < 0.25] = 0.0   # entries <= 0 keep the previous gain (note N6)
        act_d = torch.from_numpy(act).cuda()
        a_env.step_soa(act_d)                  # fused: get_param_from_actor + controller + step_update
        b_env.get_param_from_actor(act_d)      # the reference's two separate calls
        b_env.step_fixed_gains()
        assert torch.equal(a_env._state, b_env._state), t
        assert torch.equal(a_env._reward, b_env._reward) and torch.equal(a_env._flag, b_env._flag)
        assert torch.equal(a_env._next_obs, b_env._next_obs)


def test_state_norm_on_env_views_and_reference_csv_layout(tmp_path):
    import torch
    import reinforcementlearningplatform_b200 as rlp
    n = 8192
    env = rlp.UavPosCtrlRL(n_envs=n, random_trajectory=True, seed=2, auto_reset=True)
    env.reset(True)
    rng = np.random.default_rng(0)
    seen_cur, seen_next = [], []
    for t in range(5):
        env.step_soa(torch.from_numpy(rng.uniform(0, 5, (8, n))).cuda())
        s = env.current_state_norm(env.current_state, update=True)      # train.py:291
        s_ = env.next_state_norm(env.next_state, update=True)           # train.py:308
        assert s.shape == (n, 6) and s_.shape == (n, 6)
        seen_cur.append(env.current_state.cpu().numpy().copy())
        seen_next.append(env.next_state.cpu().numpy().copy())
    allc = np.concatenate(seen_cur)
    ms = env.current_state_norm.running_ms
    assert ms.n == 5 * n
    np.testing.assert_allclose(ms.mean, allc.mean(0), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(ms.std, allc.std(0), rtol=1e-10)
    np.testing.assert_allclose(s.cpu().numpy(), (seen_cur[-1] - allc.mean(0)) / (allc.std(0) + 1e-8), rtol=1e-9, atol=1e-10)
    # CSV in the reference's layout (uav_pos_ctrl_RL.py:208-233): header + one row per state dimension
    env.save_state_norm(str(tmp_path) + "/")
    lines = open(tmp_path / "state_norm.csv").read().strip().split("\n")
    assert lines[0] == "cur_n,cur_mean,cur_std,cur_S,next_n,next_mean,next_std,next_S" and len(lines) == 7
    env2 = rlp.UavPosCtrlRL(n_envs=16, random_trajectory=True)
    env2.load_norm_normalizer_from_file(str(tmp_path) + "/", "state_norm.csv")
    m2 = env2.current_state_norm.running_ms
    assert m2.n == ms.n and np.array_equal(m2.mean, ms.mean) and np.array_equal(m2.S, ms.S)
    x = torch.from_numpy(rng.normal(0, 1, (16, 6))).cuda()
    y = env2.current_state_norm(x, update=False)                       # evaluation.py / train.py:342
    np.testing.assert_allclose(y.cpu().numpy(), (x.cpu().numpy() - ms.mean) / (ms.std + 1e-8), rtol=1e-12, atol=1e-13)
    assert env2.current_state_norm.running_ms.n == ms.n


def test_random_pos0_engine_resets():
    """random_pos0=True through the engine: start within 0.3 m of the trajectory start; p, q, r of the second episode
    are the first episode's pos0 (note N5); identical to layout variant 0 until the first reset draws."""
    import torch
    import reinforcementlearningplatform_b200 as rlp
    n = 2048
    env = rlp.UavPosCtrlRL(n_envs=n, random_trajectory=True, random_pos0=True, seed=9)
    assert env._state.shape[0] == 54
    env.reset(True)
    st = env._state.cpu().numpy()
    first = st[0:3].copy()
    traj0 = np.array([0.0, 0.0, 1.5])[:, None] + st[33:36] * np.sin(st[41:44])
    assert np.all(np.abs(first - traj0) <= 0.3 + 1e-12) and np.all(st[9:12] == 0.0)
    np.testing.assert_array_equal(env.next_state.cpu().numpy().T[0:3], first - st[45:48])
    env.reset(True)
    st = env._state.cpu().numpy()
    assert np.array_equal(st[9:12], first) and np.array_equal(st[51:54], st[0:3])
