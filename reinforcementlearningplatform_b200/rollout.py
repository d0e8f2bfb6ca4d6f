"""Device-resident rollout buffer: the vector form of ``RolloutBuffer`` (utils/classes.py:250-311).

The reference keeps ``s, a, a_lp, r, s_, done, success`` as numpy ``[batch, dim]`` arrays, fills one row per env step
(``append`` :264-272) and converts everything to float32 tensors when the learner starts (``to_tensor`` :292-301).
Here one buffer holds ``T`` steps of all ``N`` instances of a vector env, time-major and field-major
(``s[T, S, N]``, ``r[T, N]`` ...), float32 from the start, and **the step kernel writes into it directly**:
``RolloutBuffer.step(env, t, action)`` points the env's output pointers (``b200env_io.obs / next_obs / reward / done /
flag``) at row ``t`` of the buffer, so a rollout costs no copy and no host round trip.  ``done`` is kept as the u8
``is_terminal`` column and ``flag`` as the i32 ``terminal_flag`` column; ``success`` is derived where it is consumed
(K-GAE, ``gae.gae_flags``) with the rule of the train loops: success = done and flag != env.TIMEOUT_FLAG
(e.g. demonstration/PPO2/PPO2-4-CartPoleAngleOnly/train.py:198-205).

Needs an env with ``io_dtype=torch.float32`` (state and arithmetic may stay float64).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from . import gae as _gae


class RolloutBuffer:
    def __init__(self, batch_size: int, env, store_log_prob: bool = True):
        """``batch_size`` = number of time steps T (the reference's ``batch_size``); instance count, state_dim and
        action_dim are taken from ``env`` (the reference passes state_dim / action_dim, utils/classes.py:251)."""
        if env.io_dtype != torch.float32:
            raise ValueError("RolloutBuffer needs an env created with io_dtype=torch.float32")
        self.batch_size = int(batch_size)
        self.n_envs = env.n_envs
        self.state_dim = env.state_dim
        self.action_dim = env.action_dim
        self.timeout_flag = int(env.TIMEOUT_FLAG)
        T, S, A, N, dev = self.batch_size, self.state_dim, self.action_dim, self.n_envs, env.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.s = torch.zeros((T, S, N), **f32)      # s        (utils/classes.py:255)
        self.a = torch.zeros((T, A, N), **f32)      # a
        self.a_lp = torch.zeros((T, A, N), **f32) if store_log_prob else None  # a_lp
        self.r = torch.zeros((T, N), **f32)         # r
        self.s_ = torch.zeros((T, S, N), **f32)     # s'
        self.done = torch.zeros((T, N), dtype=torch.uint8, device=dev)   # is_terminal
        self.flag = torch.zeros((T, N), dtype=torch.int32, device=dev)   # terminal_flag (success is derived from it)
        self.index = 0

    # ------------------------------------------------------------------ filling
    def step(self, env, t: int, action_soa: torch.Tensor, log_prob_soa: Optional[torch.Tensor] = None,
             dis_soa: Optional[torch.Tensor] = None, store_policy_obs: bool = False,
             chain_policy_obs: bool = False) -> None:
        """``env.step_update(a); buffer.append(s, a, a_lp, r, s_, done, success, t)`` of the train loops (e.g.
        PPO2-4-CartPoleAngleOnly/train.py:193-215) for every instance, in one kernel launch: the step kernel stores
        current_state, next_state, reward, is_terminal and terminal_flag straight into row ``t``.  ``action_soa`` is
        ``[action_dim, N]`` float32; pass ``self.a[t]`` itself to avoid the copy."""
        if action_soa.data_ptr() != self.a[t].data_ptr():
            self.a[t].copy_(action_soa)
        if log_prob_soa is not None and self.a_lp is not None:
            self.a_lp[t].copy_(log_prob_soa)
        if chain_policy_obs:
            # Row s[t] already IS the observation the policy acted on: the previous step wrote its policy-facing
            # observation there (the caller copies env.policy_state into s[0] once per rollout), and this step writes its
            # own into s[t + 1] -- or, for the last row, back into env.policy_state.  No per-step row copy, and the kernels
            # skip the current_state they would otherwise recompute.
            nxt = self.s[t + 1] if t + 1 < self.batch_size else env._reset_obs
            env.step_into(self.a[t], dis_soa, obs=None, next_obs=self.s_[t], reward=self.r[t], done=self.done[t],
                          flag=self.flag[t], policy_obs=nxt)
            self.index = t + 1
            return
        obs_row = self.s[t]
        if store_policy_obs and not (env.reuse_obs and env._policy_obs_valid):
            # Row s[t] = the observation the policy acted on (`s` of PPO2-4-UavFntsmcParamPos/train.py:290-303, the copy
            # of the previous next_state), not the current_state step_update recomputes against the new reference.  For
            # pure-observation envs the two are the same bits and step_into makes this copy itself.
            self.s[t].copy_(env._reset_obs)
            if getattr(self, "_obs_scratch", None) is None:
                self._obs_scratch = torch.empty_like(self.s[0])
            obs_row = self._obs_scratch
        env.step_into(self.a[t], dis_soa, obs=obs_row, next_obs=self.s_[t], reward=self.r[t], done=self.done[t],
                      flag=self.flag[t])
        self.index = t + 1

    def collect(self, env, t0: int = 0, steps: Optional[int] = None, dis: Optional[torch.Tensor] = None) -> None:
        """Rows ``t0 .. t0 + steps - 1`` in one call from the actions already stored in ``self.a`` (random-action
        benchmarks, open-loop replays, actions produced on the device ahead of time): ``b200env_rollout``."""
        steps = self.batch_size - t0 if steps is None else int(steps)
        sl = slice(t0, t0 + steps)
        env.rollout_into(steps, self.a[sl], obs=self.s[sl], next_obs=self.s_[sl], reward=self.r[sl],
                         done=self.done[sl], flag=self.flag[sl], dis=dis)
        self.index = t0 + steps

    # ------------------------------------------------------------------ reading
    def success(self) -> torch.Tensor:
        """float32 ``[T, N]`` success column as the reference's learner sees it (utils/classes.py:299)."""
        return ((self.done != 0) & (self.flag != self.timeout_flag)).float()

    def to_tensor(self):
        """``s, a, a_lp, r, s_, done, success`` like ``RolloutBuffer.to_tensor`` (utils/classes.py:292-301), as
        ``[T * N, dim]`` float32 views/copies in (t, instance) order for learners that want the flat batch."""
        T, N = self.batch_size, self.n_envs
        flat = lambda x: x.permute(0, 2, 1).reshape(T * N, -1)
        a_lp = None if self.a_lp is None else flat(self.a_lp)
        return (flat(self.s), flat(self.a), a_lp, self.r.reshape(T * N, 1), flat(self.s_),
                self.done.float().reshape(T * N, 1), self.success().reshape(T * N, 1))

    def gae(self, vs: torch.Tensor, vs_next: torch.Tensor, gamma: float, lmd: float, acc_mode: int = 0,
            normalize: bool = True, group=None):
        """adv, v_target of Proximal_Policy_Optimization2.learn (:88-100) over the whole buffer; ``vs`` / ``vs_next``
        are the critic's values of ``s`` / ``s_`` as ``[T, N]`` float32.  With a process group the normalisation
        statistics are all-reduced (global mean / std over all GPUs)."""
        adv, vt, stats = _gae.gae_flags(self.r, vs, vs_next, self.done, self.flag, self.timeout_flag, gamma, lmd,
                                        acc_mode)
        if normalize:
            _gae.normalize_advantage(adv, stats, group=group)
        return adv, vt
