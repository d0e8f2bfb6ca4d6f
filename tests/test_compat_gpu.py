"""GPU: the single-instance rl_base view (compat.SingleEnv) driven with the call sequence of the reference train loops
(demonstration/PPO2/PPO2-4-CartPoleAngleOnly/train.py:184-216, PPO2-4-UavFntsmcParamPos/train.py:273-310,
PPO2-4-UavFntsmcParamAtt/train.py:254-290) on the recorded fixtures: numpy attributes, python scalars, same numbers."""
import numpy as np
import pytest

from helpers import env_specs, load_golden, mixed_err

pytestmark = pytest.mark.gpu


def _drive(name, lane, steps, uav=None):
    import reinforcementlearningplatform_b200 as rlp
    g = load_golden(name)
    cls, kw = env_specs()[name]
    env = rlp.single(cls(n_envs=1, **kw))
    assert isinstance(env.state_dim, int) and np.asarray(env.action_range).shape == (env.action_dim, 2)
    env.set_state(g["state0"][lane], g["time0"][lane])
    worst, episodes = 0.0, 0
    T = min(steps, g["reward"].shape[0])
    for t in range(T):
        env.current_state = env.next_state.copy()                      # train.py:193
        a = g["actions"][t, lane]
        if uav == "pos":
            env.get_param_from_actor(a)                                 # Pos/train.py:292
            action_4_uav = env.generate_action_4_uav()                  # :293
            env.step_update(action_4_uav)                               # :294
        elif uav == "att":
            env.get_param_from_actor(a)                                 # Att/train.py:265
            torque = env.att_control(None, None, None)                  # :268 (ref_inner is fused into the step)
            env.step_update([torque[0], torque[1], torque[2]])          # :269
        else:
            env.step_update(a)                                          # train.py:195
        assert isinstance(env.reward, float) and isinstance(env.is_terminal, bool) and isinstance(env.terminal_flag, int)
        assert env.next_state.dtype == np.float64 and env.next_state.shape == (env.state_dim,)
        worst = max(worst, mixed_err(env.next_state, g["next_obs"][t, lane]), mixed_err(env.current_state, g["obs"][t, lane]),
                    mixed_err(np.array([env.reward]), np.array([g["reward"][t, lane]])))
        assert env.is_terminal == bool(g["done"][t, lane]) and env.terminal_flag == int(g["flag"][t, lane]), (name, t)
        assert env.time == g["time"][t, lane]
        if env.is_terminal:                                             # train.py:187-191: the loop resets the env itself
            episodes += 1
            env.reset(True)                                             # Philox draw; the fixture's own reset is injected next
            assert not env.is_terminal and env.reward == 0.0 and env.time == 0.0
            env.set_state(g["reset_state"][t, lane], g["reset_time"][t, lane])
    return worst, episodes


@pytest.mark.parametrize("name,lane,tol", [("cartpole_angleonly_ppo2", 0, 1e-9), ("cartpole", 1, 1e-9), ("soi", 0, 1e-9),
                                           ("fas_discrete", 0, 1e-9), ("ugv_forward", 0, 1e-9), ("ugvo_dppo2", 0, 1e-8)])
def test_single_env_runs_the_reference_loop_shape(name, lane, tol):
    worst, episodes = _drive(name, lane, 400)
    assert worst <= tol, (name, worst)


def test_single_uav_pos_three_call_protocol():
    worst, _ = _drive("uav_pos", 1, 300, uav="pos")
    assert worst <= 1e-9, worst


def test_single_uav_att_three_call_protocol():
    worst, _ = _drive("uav_att_rand", 1, 300, uav="att")
    assert worst <= 1e-9, worst


def test_single_uav_reset_with_new_controller_params():
    """reset_uav_pos_ctrl_RL_tracking(new_pos_ctrl_parma=...) re-seeds the per-instance gains (Pos/train.py:275-286)."""
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200.envs import uav as U
    env = rlp.single(rlp.UavPosCtrlRL(n_envs=1, random_trajectory=True))
    par = U.train_pos_ctrl_param()
    env.reset_uav_pos_ctrl_RL_tracking(random_trajectroy=True, random_pos0=False, new_att_ctrl_param=None,
                                       new_pos_ctrl_parma=par)
    st = env._env._state[:, 0].cpu().numpy()
    i = env.STATE_FIELDS.index("k1_0")
    assert np.array_equal(st[i:i + 3], par.k1) and np.array_equal(st[i + 3:i + 6], par.k2)
    assert env.time == 0.0 and not env.is_terminal and env.next_state.shape == (6,)
    env.step_update(env.generate_action_4_uav())   # no get_param_from_actor: fixed gains (test_pos_tracking_ctrl.py loop)
    assert np.isfinite(env.reward) and env.time == pytest.approx(0.02)


@pytest.mark.parametrize("kind", ["pos", "att"])
def test_single_uav_view_has_every_attribute_the_train_scripts_touch(kind):
    """`grep -o 'env\\.[A-Za-z_0-9]*'` over PPO2-4-UavFntsmcParamPos/train.py and PPO2-4-UavFntsmcParamAtt/train.py: every name the
    scripts read or call on `env` / `env_test` resolves on the single-instance view (drawing calls are no-ops)."""
    import reinforcementlearningplatform_b200 as rlp
    common = ["action_dim", "action_range", "state_dim", "name", "dt", "time_max", "current_state", "next_state", "reward",
              "is_terminal", "terminal_flag", "time", "get_param_from_actor", "step_update", "current_state_norm",
              "next_state_norm", "save_state_norm", "visualization", "show_image"]
    if kind == "pos":
        env = rlp.single(rlp.UavPosCtrlRL(n_envs=1, random_trajectory=True))
        names = common + ["generate_action_4_uav", "reset_uav_pos_ctrl_RL_tracking"]
    else:
        env = rlp.single(rlp.UavAttCtrlRL(n_envs=1, random_trajectory=True))
        names = common + ["att_control", "reset_uav_att_ctrl_RL_tracking", "ref_att_amplitude", "ref_att_period",
                          "ref_att_bias_phase", "ref_att_bias_a"]
    for n in names:
        assert hasattr(env, n), n
    env.show_image(True)
    env.visualization()
    s = env.current_state_norm(np.zeros(env.state_dim), update=False)
    assert isinstance(s, np.ndarray) and s.dtype == np.float64 and s.shape == (env.state_dim,)


@pytest.mark.parametrize("name", ["cartpole", "cartpole_angleonly_ppo2", "fas_ppo2", "soi", "ballbalancer", "twolink",
                                  "ugv_forward", "ugvo_dppo2", "fas_discrete"])
def test_single_view_has_every_attribute_the_demo_scripts_touch(name, tmp_path):
    """Union of `env.<name>` over the reference's DDPG / DQN / PPO / PPO2 / SAC / TD3 train scripts for the non-UAV envs
    (demonstration/*/*/train.py) plus the descriptive lists of algorithm/rl_base.py:5-124."""
    import reinforcementlearningplatform_b200 as rlp
    cls, kw = env_specs()[name]
    env = rlp.single(cls(n_envs=1, **kw))
    names = ["action_dim", "action_range", "state_dim", "name", "dt", "current_action", "current_state", "next_state",
             "is_terminal", "terminal_flag", "reward", "reset", "step_update", "visualization", "save_state_norm",
             "current_state_norm", "next_state_norm", "action_num", "action_space", "action_step", "state_num", "state_step",
             "state_space", "state_range", "isStateContinuous", "isActionContinuous", "use_norm"]
    for n in names:
        assert hasattr(env, n), (name, n)
    assert hasattr(env, "timeMax") or hasattr(env, "time_max")
    assert len(env.action_num) == env.action_dim and len(env.state_range) == env.state_dim
    if name == "fas_discrete":
        assert env.action_num[0] == len(env.action_space[0])          # FlightAttitudeSimulatorDiscrete.py:58-59
    else:
        assert env.action_num == [np.inf] * env.action_dim and env.isActionContinuous == [True] * env.action_dim
    env.reset(True)
    s = env.current_state_norm(env.next_state, update=True)            # PPO-4-*/train.py, PPO2-4-SecondOrderIntegration/train.py
    assert isinstance(s, np.ndarray) and s.shape == (env.state_dim,)
    env.save_state_norm(str(tmp_path) + "/")                           # PPO2-4-*/train.py
    assert (tmp_path / "state_norm.csv").exists()
