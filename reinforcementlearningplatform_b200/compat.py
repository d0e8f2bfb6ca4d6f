"""Single-instance view of a vector env with the reference's scalar ``rl_base`` contract (algorithm/rl_base.py:4-162).

The reference's ``train.py`` loops and learners talk to ONE env object through numpy values:
``env.current_state = env.next_state.copy()``, ``a = agent.choose_action(env.current_state)``, ``env.step_update(a)``,
then ``env.reward`` (float), ``env.is_terminal`` (bool), ``env.terminal_flag`` (int), ``env.next_state`` (float64
array), and ``env.reset(True)`` when the episode ended (demonstration/PPO2/PPO2-4-CartPoleAngleOnly/train.py:184-216).
``SingleEnv(vec_env)`` gives exactly that on top of an ``n_envs == 1`` engine env, so those loops run unchanged while
every number still comes from the CUDA kernels (one launch + one small device->host copy per step; this is the
compatibility path, the batched classes are the fast one).

UAV envs: the reference loop calls ``env.get_param_from_actor(a)``, then ``env.generate_action_4_uav()`` (position) or
``ref_inner`` + ``env.att_control(...)`` (attitude), then ``env.step_update(action_4_uav)``
(PPO2-4-UavFntsmcParamPos/train.py:292-297, ...Att/train.py:265-276).  The engine fuses the three calls, so the view
defers: ``get_param_from_actor`` stores the gains, the control call returns a placeholder, ``step_update`` launches
the fused step with the stored gains.
"""
from __future__ import annotations

import numpy as np
import torch


class _HostNormalization:
    """``env.current_state_norm`` / ``env.next_state_norm`` as the scalar train loops use them (uav_pos_ctrl_RL.py:36-37,
    PPO2-4-UavFntsmcParamPos/train.py:291,308): numpy state in, numpy float64 state out -- the loop hands the result to
    ``agent.choose_action`` and stores it in the learner's numpy ``RolloutBuffer``.  The statistics live in the engine's
    device ``Normalization`` (one-sample updates follow the reference's Welford recurrence, csrc/norm.cu); ``running_ms`` and
    everything else is forwarded to it."""

    def __init__(self, norm):
        self._norm = norm

    def __call__(self, x, update: bool = True):
        return self._norm(np.asarray(x, dtype=np.float64), update).double().cpu().numpy()

    def __getattr__(self, name):
        return getattr(self.__dict__["_norm"], name)


class SingleEnv:
    _OWN = ("_env", "current_state", "next_state", "current_action", "reward", "is_terminal", "terminal_flag", "time",
            "_pending_gains", "_pack")

    def __init__(self, vec_env):
        if vec_env.n_envs != 1:
            raise ValueError("SingleEnv wraps an env built with n_envs=1")
        if vec_env.auto_reset:
            raise ValueError("SingleEnv: build the env with auto_reset=False (the train loop calls reset itself)")
        self._env = vec_env
        self._pending_gains = None
        od = vec_env.state_dim
        self._pack = torch.zeros(2 * od + 4, dtype=torch.float64, device=vec_env.device)
        self.current_state = np.zeros(od)
        self.next_state = np.zeros(od)
        self.current_action = np.zeros(vec_env.action_dim)
        self.reward, self.is_terminal, self.terminal_flag, self.time = 0.0, False, 0, 0.0

    def __getattr__(self, name):  # everything else (state_dim, action_dim, action_range, name, dt, timeMax, ...)
        return getattr(self.__dict__["_env"], name)

    # the env's running state normalisers (SecondOrderIntegration.py:76, uav_pos_ctrl_RL.py:36-37) with numpy in / numpy out
    @property
    def current_state_norm(self):
        if "_cur_norm_host" not in self.__dict__:
            self.__dict__["_cur_norm_host"] = _HostNormalization(self._env.current_state_norm)
        return self.__dict__["_cur_norm_host"]

    @property
    def next_state_norm(self):
        if "_next_norm_host" not in self.__dict__:
            self.__dict__["_next_norm_host"] = _HostNormalization(self._env.next_state_norm)
        return self.__dict__["_next_norm_host"]

    # ------------------------------------------------------------------ helpers
    def _pull(self, with_obs: bool):
        e, od = self._env, self._env.state_dim
        p = self._pack
        p[0:od] = e._obs[:, 0]
        p[od:2 * od] = e._next_obs[:, 0]
        p[2 * od] = e._reward[0]
        p[2 * od + 1] = e._done[0]
        p[2 * od + 2] = e._flag[0]
        p[2 * od + 3] = e._time[0]
        h = p.cpu().numpy()
        if with_obs:
            self.current_state = h[0:od].copy()
        self.next_state = h[od:2 * od].copy()
        self.reward = float(h[2 * od])
        self.is_terminal = bool(h[2 * od + 1])
        self.terminal_flag = int(h[2 * od + 2])
        self.time = float(h[2 * od + 3])

    # ------------------------------------------------------------------ rl_base
    def reset(self, random: bool = True):
        self._env.reset(random)
        self._pull(False)
        self.current_state = self.next_state.copy()
        self.current_action = np.zeros(self._env.action_dim)
        self.reward, self.is_terminal, self.terminal_flag = 0.0, False, 0

    def step_update(self, action):
        e = self._env
        a = np.asarray(action, dtype=np.float64).reshape(-1)
        self.current_action = a.copy()
        e._policy_obs_valid = False  # the caller may have changed the state; recompute current_state in-kernel
        e.step_soa(torch.as_tensor(a, dtype=e.io_dtype, device=e.device).view(-1, 1).contiguous())
        self._pull(True)

    def get_state(self):
        self._env.observe()
        return self._env._next_obs[:, 0].double().cpu().numpy()

    def get_reward(self, param=None):
        return self.reward

    def is_Terminal(self, param=None):
        return self.is_terminal

    def is_success(self):
        return self.is_terminal and self.terminal_flag != self._env.TIMEOUT_FLAG

    def visualization(self):  # OpenCV drawing is out of scope
        pass

    def show_image(self, iswait: bool = False):  # called once by the UAV train scripts (Pos/train.py:214); no window here
        pass

    # ------------------------------------------------------------------ parity / checkpoint helpers
    def set_state(self, state, time):
        self._env.set_state_buffers(np.asarray(state, dtype=np.float64).reshape(-1, 1), np.asarray([time]))
        self._env.observe()
        self._pull(False)
        self.time = float(time)


class SingleUavEnv(SingleEnv):
    """Deferred three-call protocol of the UavFntsmcParam train loops (see module docstring)."""
    _PLACEHOLDER_POS = [float("nan")] * 4   # what generate_action_4_uav() hands back; step_update ignores its values
    _PLACEHOLDER_ATT = np.full(3, np.nan)


    def get_param_from_actor(self, action_from_actor, update_k2: bool = True):
        a = np.array(action_from_actor, dtype=np.float64).reshape(-1)
        if not update_k2:
            a[3:6] = 0.0  # entries <= 0 leave the gain untouched (uav_pos_ctrl_RL.py:165-173)
        self._pending_gains = a

    def generate_action_4_uav(self, att_limit: bool = True):
        return list(self._PLACEHOLDER_POS)

    def att_control(self, ref=None, dot_ref=None, dot2_ref=None, att_only: bool = True):
        return self._PLACEHOLDER_ATT.copy()

    def step_update(self, action=None, dis=None):
        e = self._env
        g = self._pending_gains if self._pending_gains is not None else np.zeros(8)
        self._pending_gains = None
        self.current_action = np.asarray(g, dtype=np.float64).copy()
        e._policy_obs_valid = False
        d = None if dis is None else torch.as_tensor(np.asarray(dis, dtype=np.float64), dtype=e.io_dtype,
                                                      device=e.device).view(-1, 1).contiguous()
        e.step_soa(torch.as_tensor(g, dtype=e.io_dtype, device=e.device).view(-1, 1).contiguous(), d)
        self._pull(True)

    def _apply_new_params(self, new_att_ctrl_param, new_pos_ctrl_param):
        e, p = self._env, self._env._params
        for prefix, par in (("att_", new_att_ctrl_param), ("pos_", new_pos_ctrl_param)):
            if par is None:
                continue
            for name in ("k1", "k2", "alpha", "beta", "gamma", "lmd"):
                dst = getattr(p, prefix + name)
                for i, v in enumerate(np.asarray(getattr(par, name), dtype=np.float64).reshape(-1)[:3]):
                    dst[i] = float(v)
        e._hot_io = None

    def reset_uav_pos_ctrl_RL_tracking(self, random_trajectroy: bool = False, random_pos0: bool = False,
                                       yaw_fixed: bool = False, new_att_ctrl_param=None, new_pos_ctrl_parma=None,
                                       outer_param=None):
        """uav_pos_ctrl_RL.py:181-207 (argument names as spelled there)."""
        e = self._env
        if bool(random_pos0) != bool(getattr(e, "random_pos0", False)):
            raise ValueError("random_pos0 selects the state layout: pass it to the UavPosCtrlRL constructor")
        if outer_param is not None:
            raise NotImplementedError("outer_param: inject the trajectory with set_state_buffers()")
        self._apply_new_params(new_att_ctrl_param, new_pos_ctrl_parma)
        e._params.random_trajectory, e._params.yaw_fixed = int(random_trajectroy), int(yaw_fixed)
        self.reset(True)

    def reset_uav_att_ctrl_RL_tracking(self, random_trajectory: bool = False, yaw_fixed: bool = False,
                                       new_att_ctrl_param=None, outer_param=None):
        """uav_att_ctrl_RL.py:158-178."""
        e = self._env
        if outer_param is not None:
            raise NotImplementedError("outer_param: inject the trajectory with set_state_buffers()")
        self._apply_new_params(new_att_ctrl_param, None)
        e._params.random_trajectory, e._params.yaw_fixed = int(random_trajectory), int(yaw_fixed)
        self.reset(True)

    # trajectory parameters the attitude train loop reads to call ref_inner itself (Att/train.py:266-267)
    def _fields(self, first, count):
        i = self._env.STATE_FIELDS.index(first)
        return self._env._state[i:i + count, 0].double().cpu().numpy()

    @property
    def ref_att_amplitude(self):
        return self._fields("A_0", 3)

    @property
    def ref_att_period(self):
        return self._fields("T_0", 3)

    @property
    def ref_att_bias_phase(self):
        return self._fields("phase_0", 3)

    @property
    def ref_att_bias_a(self):
        return np.array([self._env._params.ref_bias_a[i] for i in range(3)])


def single(vec_env) -> SingleEnv:
    """The right single-instance view for an ``n_envs == 1`` engine env."""
    from .envs.uav import _UavBase
    return SingleUavEnv(vec_env) if isinstance(vec_env, _UavBase) else SingleEnv(vec_env)
