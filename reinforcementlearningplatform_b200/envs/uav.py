"""Batched UavFntsmcParam attitude / position tracking envs (kernels K-UAVA / K-UAVP, csrc/uav.cu).

Mirrors ``environment/UavFntsmcParam``: ``uav.py`` (quadrotor + zones), ``FNTSMC.py`` (controllers),
``uav_att_ctrl_RL.py`` / ``uav_pos_ctrl_RL.py`` (RL wrappers) and ``ref_cmd.py``.  Constructor arguments are
duck-typed copies of the reference's ``uav_param`` / ``fntsmc_param`` objects (same attribute names), so a
``train.py`` can pass its own parameter objects unchanged.

One engine step = the fused loop body of the reference's train.py
(``get_param_from_actor(a)`` -> ``generate_action_4_uav()`` | ``ref_inner``+``att_control`` -> ``step_update``);
``step_update(a8)`` therefore takes the 8 controller gains the actor outputs.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib
from ..vec_env import VecEnvBase


def deg2rad(deg):  # utils/functions.py:4-5
    return deg * math.pi / 180.


class uav_param:
    """environment/UavFntsmcParam/uav.py:7-25 (same attribute names and defaults)."""

    def __init__(self):
        self.m = 0.8
        self.g = 9.8
        self.J = np.array([4.212e-3, 4.212e-3, 8.255e-3])
        self.d = 0.12
        self.CT = 2.168e-6
        self.CM = 2.136e-8
        self.J0 = 1.01e-5
        self.kr = 1e-3
        self.kt = 1e-3
        self.pos0 = np.array([0, 0, 0])
        self.vel0 = np.array([0, 0, 0])
        self.angle0 = np.array([0, 0, 0])
        self.pqr0 = np.array([0, 0, 0])
        self.dt = 0.01
        self.time_max = 20
        self.pos_zone = np.atleast_2d([[-5, 5], [-5, 5], [0, 3]])
        self.att_zone = np.atleast_2d([[deg2rad(-45), deg2rad(45)], [deg2rad(-45), deg2rad(45)],
                                       [deg2rad(-120), deg2rad(120)]])


class fntsmc_param:
    """environment/UavFntsmcParam/FNTSMC.py:4-14."""

    def __init__(self):
        self.k1 = np.array([1.2, 0.8, 1.5])
        self.k2 = np.array([0.2, 0.6, 1.5])
        self.alpha = np.array([1.2, 1.5, 1.2])
        self.beta = np.array([0.3, 0.3, 0.3])
        self.gamma = np.array([0.2, 0.2, 0.2])
        self.lmd = np.array([2.0, 2.0, 2.0])
        self.dim = 3
        self.dt = 0.01
        self.ctrl0 = np.array([0., 0., 0.])


def train_uav_param(kind: str) -> uav_param:
    """Quadrotor parameters of the PPO2 training scripts
    (PPO2-4-UavFntsmcParamAtt/train.py:29-48, PPO2-4-UavFntsmcParamPos/train.py:29-48)."""
    p = uav_param()
    p.dt = 0.02
    p.time_max = 10
    if kind == 'att':
        p.pos_zone = np.atleast_2d([[-np.inf, np.inf], [-np.inf, np.inf], [-np.inf, np.inf]])
        p.att_zone = np.atleast_2d([[deg2rad(-90), deg2rad(90)], [deg2rad(-90), deg2rad(90)],
                                    [deg2rad(-180), deg2rad(180)]])
    else:
        p.pos_zone = np.atleast_2d([[-3, 3], [-3, 3], [0, 3]])
        p.att_zone = np.atleast_2d([[deg2rad(-90), deg2rad(90)], [deg2rad(-90), deg2rad(90)],
                                    [deg2rad(-120), deg2rad(120)]])
    return p


def train_att_ctrl_param() -> fntsmc_param:
    """train.py:52-62 (both scripts)."""
    p = fntsmc_param()
    p.k1 = np.array([25., 25., 40.])
    p.k2 = np.array([0.1, 0.1, 0.2])
    p.alpha = np.array([2.5, 2.5, 2.5])
    p.beta = np.array([0.99, 0.99, 0.99])
    p.gamma = np.array([1.5, 1.5, 1.2])
    p.lmd = np.array([2.0, 2.0, 2.0])
    p.dt = 0.02
    return p


def train_pos_ctrl_param() -> fntsmc_param:
    """PPO2-4-UavFntsmcParamPos/train.py:66-76."""
    p = fntsmc_param()
    p.k1 = np.array([1.2, 0.8, 0.5])
    p.k2 = np.array([0.2, 0.6, 0.5])
    p.alpha = np.array([1.2, 1.5, 1.2])
    p.beta = np.array([0.3, 0.3, 0.5])
    p.gamma = np.array([0.2, 0.2, 0.2])
    p.lmd = np.array([2.0, 2.0, 2.0])
    p.dt = 0.02
    return p


def zero_gains(p: fntsmc_param) -> fntsmc_param:
    """reset_*_ctrl_param('zero') of the training scripts (Att/train.py:69-74, Pos/train.py:84-89)."""
    p.k1 = 0.01 * np.ones(3)
    p.k2 = 0.01 * np.ones(3)
    p.gamma = 0.01 * np.ones(3)
    p.lmd = 0.01 * np.ones(3)
    return p


def _set3(dst, src):
    for i, v in enumerate(np.asarray(src, dtype=np.float64).reshape(-1)):
        dst[i] = float(v)


class _UavBase(VecEnvBase):
    def _fill_common(self, p: _lib.UavParams, up: uav_param, att: fntsmc_param):
        p.m, p.g, p.kr, p.kt = up.m, up.g, up.kr, up.kt
        _set3(p.J, up.J)
        p.dt, p.time_max = up.dt, up.time_max
        mae = 3.0  # self.max_admissible_error, uav.py:91
        d1 = deg2rad(1)
        for i in range(3):
            p.pos_lo[i] = float(up.pos_zone[i][0]) - mae      # uav.py:184-190
            p.pos_hi[i] = float(up.pos_zone[i][1]) + mae
            p.att_lo[i] = float(up.att_zone[i][0]) + d1       # uav.py:197-203
            p.att_hi[i] = float(up.att_zone[i][1]) - d1
            p.att_zone_min[i] = float(up.att_zone[i][0])
            p.att_zone_max[i] = float(up.att_zone[i][1])
        p.t_term = up.time_max - up.dt / 2                    # uav.py:216
        # uav.py:64: init_state = concatenate((pos0, vel0, angle0, pos0))  -- pqr0 is NOT used (N5)
        _set3(p.init_state, np.concatenate((up.pos0, up.vel0, up.angle0, up.pos0)))
        for name in ("k1", "k2", "alpha", "beta", "gamma", "lmd"):
            _set3(getattr(p, "att_" + name), getattr(att, name))
        p.dot_att_ref_limit = 60. * np.pi / 180.              # uav_pos_ctrl.py:25
        p.att_limit = np.pi / 4                               # uav_pos_ctrl.py:314

    # ------------------------------------------------------------------ reference method names (host side, not hot)
    K1_SCALE, K2_DIV = 1.0, 1.0     # get_param_from_actor scaling of the first six actor outputs
    _GAIN_FIELDS = ("k1_0", "k2_0", "gamma_0", "lmd_0")

    def get_param_from_actor(self, action_from_actor) -> None:
        """``get_param_from_actor`` on its own (uav_pos_ctrl_RL.py:158-173, uav_att_ctrl_RL.py:141-156): overwrite the
        per-instance gains where the actor output is > 0.  ``step_update(a8)`` already fuses this; the separate call
        exists for scripts that set gains once and then run the controller with them (``step_fixed_gains``)."""
        a = self._as_soa(action_from_actor, 8).to(self.dtype)
        k1, k2, gm, ld = (self.STATE_FIELDS.index(f) for f in self._GAIN_FIELDS)
        st = self._state                # logical copy (the UAV state buffer is block-interleaved on the device)
        # a device-side divisor: torch turns division by a host scalar into a multiplication by its reciprocal,
        # which is not the IEEE quotient `a / 10` of the reference (and of the fused kernel)
        k2_div = torch.full((), self.K2_DIV, dtype=self.dtype, device=self.device)
        for i in range(3):
            st[k1 + i] = torch.where(a[i] > 0, a[i] * self.K1_SCALE, st[k1 + i])
            st[k2 + i] = torch.where(a[3 + i] > 0, a[3 + i] / k2_div, st[k2 + i])
            st[gm + i] = torch.where(a[6] > 0, a[6], st[gm + i])
            st[ld + i] = torch.where(a[7] > 0, a[7], st[ld + i])
        self._write_state(st)

    def step_fixed_gains(self, dis=None) -> None:
        """One control period with the gains currently stored per instance: ``generate_action_4_uav()`` /
        ``att_control()`` + ``step_update()`` of the reference's non-RL scripts (test_pos_tracking_ctrl.py:66-102).
        An all-zero actor output leaves every gain untouched (note N6), so this is the same fused kernel."""
        if getattr(self, "_zero_action", None) is None:
            self._zero_action = torch.zeros(8, self.n_envs, dtype=self.io_dtype, device=self.device)
        d = None if dis is None else self._as_soa(dis, self._dd)
        self.step_soa(self._zero_action, d)

    # reference attribute names
    @property
    def time_max(self):
        return self._uav_param.time_max

    @property
    def dt(self):
        return self._uav_param.dt


class UavAttCtrlRL(_UavBase):
    """``uav_att_ctrl_RL`` (uav_att_ctrl_RL.py:10-178): attitude tracking, obs = (att - ref, Euler rate - ref rate),
    action = 8 gains in [0, 3] (k1 x10, k2 /10 as in get_param_from_actor :141-156)."""
    ENV_ID = _lib.UAV_ATT
    K1_SCALE, K2_DIV = 10.0, 10.0   # k1 = 10 a, k2 = a / 10 (uav_att_ctrl_RL.py:149-152)
    TIMEOUT_FLAG = 1  # success = done and flag != 1 (PPO2-4-UavFntsmcParamPos/train.py:299-302, ...Att/train.py:278)
    STATE_FIELDS = tuple("phi theta psi p q r s1_0 s1_1 s1_2 k1_0 k1_1 k1_2 k2_0 k2_1 k2_2 gamma_0 gamma_1 gamma_2 "
                         "lmd_0 lmd_1 lmd_2 A_0 A_1 A_2 T_0 T_1 T_2 phase_0 phase_1 phase_2 "
                         "ref_0 ref_1 ref_2 dref_0 dref_1 dref_2".split())

    def __init__(self, n_envs: int = 1, _uav_param: uav_param = None, _uav_att_param: fntsmc_param = None,
                 random_trajectory: bool = False, yaw_fixed: bool = False, **kw):
        self._uav_param = _uav_param or train_uav_param('att')
        self._att_param = _uav_att_param or zero_gains(train_att_ctrl_param())
        self.random_trajectory, self.yaw_fixed = random_trajectory, yaw_fixed
        self.name = 'uav_pos_ctrl_RL'  # sic: uav_att_ctrl_RL.py:20
        self.staticGain = 2.0
        self.use_norm = False
        self.Q_att = np.array([1., 1., 1.])
        self.Q_pqr = np.array([0.01, 0.01, 0.01])
        self.R = np.array([0.01, 0.01, 0.01])
        super().__init__(n_envs, **kw)
        self.action_range = [[0, 3.0] for _ in range(8)]  # uav_att_ctrl_RL.py:36

    def make_params(self):
        up = self._uav_param
        p = _lib.UavParams()
        self._fill_common(p, up, self._att_param)
        _set3(p.Q_e, self.Q_att)
        _set3(p.Q_de, self.Q_pqr)
        _set3(p.R, self.R)
        # deterministic trajectory, uav_att_ctrl.py:163-166
        _set3(p.ref_amplitude, [np.pi / 3, np.pi / 3, np.pi / 2, 0.])
        _set3(p.ref_period, [5, 5, 5, 1.])
        _set3(p.ref_bias_a, [0., 0., 0., 0.])
        _set3(p.ref_bias_phase, [np.pi / 2, 0., 0., 0.])
        # random trajectory, uav_att_ctrl.py:156-161
        phi_max, theta_max, psi_max = (float(up.att_zone[i][1]) for i in range(3))
        _set3(p.traj_A_hi, [phi_max if phi_max < np.pi / 3 else np.pi / 3,
                            theta_max if theta_max < np.pi / 3 else np.pi / 3,
                            psi_max if psi_max < np.pi / 2 else np.pi / 2, 0.])
        p.traj_T_lo, p.traj_T_hi, p.traj_phase_hi = 3, 6, np.pi / 2
        p.random_trajectory, p.yaw_fixed = int(self.random_trajectory), int(self.yaw_fixed)
        return p


class UavPosCtrlRL(_UavBase):
    """``uav_pos_ctrl_RL`` (uav_pos_ctrl_RL.py:12-207): position tracking with the FNTSMC outer + inner loops,
    obs = (pos - ref, vel - ref vel), action = 8 outer-loop gains in [0, 5]; optional injected disturbance [3]."""
    ENV_ID = _lib.UAV_POS
    TIMEOUT_FLAG = 1  # success = done and flag != 1 (PPO2-4-UavFntsmcParamPos/train.py:299-302, ...Att/train.py:278)
    STATE_FIELDS = tuple("x y z vx vy vz phi theta psi p q r sig_0 sig_1 sig_2 s1_0 s1_1 s1_2 aref_0 aref_1 aref_2 "
                         "k1_0 k1_1 k1_2 k2_0 k2_1 k2_2 gamma_0 gamma_1 gamma_2 lmd_0 lmd_1 lmd_2 "
                         "A_0 A_1 A_2 A_3 T_0 T_1 T_2 T_3 phase_0 phase_1 phase_2 phase_3 "
                         "pref_0 pref_1 pref_2 dpref_0 dpref_1 dpref_2".split())

    def __init__(self, n_envs: int = 1, _uav_param: uav_param = None, _uav_att_param: fntsmc_param = None,
                 _uav_pos_param: fntsmc_param = None, random_trajectory: bool = True, yaw_fixed: bool = False,
                 random_pos0: bool = False, **kw):
        """``random_pos0``: ``reset_uav_pos_ctrl(random_pos0=True)`` (uav_pos_ctrl.py:510-513): every reset starts the
        quadrotor within +-0.3 m of the first trajectory point; selects state layout variant 1 (3 extra fields holding
        the reference's ``init_state[9:12]``, from which the NEXT reset loads p, q, r -- note N5, reproduced)."""
        self.random_pos0 = bool(random_pos0)
        if self.random_pos0:
            self.VARIANT = 1
            self.STATE_FIELDS = type(self).STATE_FIELDS + ("next_pqr0_0", "next_pqr0_1", "next_pqr0_2")
        self._uav_param = _uav_param or train_uav_param('pos')
        self._att_param = _uav_att_param or train_att_ctrl_param()
        self._pos_param = _uav_pos_param or zero_gains(train_pos_ctrl_param())
        self.random_trajectory, self.yaw_fixed = random_trajectory, yaw_fixed
        self.name = 'uav_pos_ctrl_RL'
        self.staticGain = 2.0
        self.use_norm = False
        self.Q_pos = np.array([1., 1., 1.])
        self.Q_vel = np.array([0.05, 0.05, 0.05])
        self.R = np.array([0.01, 0.01, 0.01])
        super().__init__(n_envs, **kw)
        if self.random_pos0 and not self.host_only:  # init_state[9:12] = pos0 of the constructor (uav.py:64)
            for k in range(3):
                self._state[51 + k].fill_(float(self._params.init_state[9 + k]))
        self.action_range = [[0, 5.0] for _ in range(8)]  # uav_pos_ctrl_RL.py:44

    def make_params(self):
        up = self._uav_param
        p = _lib.UavParams()
        self._fill_common(p, up, self._att_param)
        for name in ("k1", "k2", "alpha", "beta", "gamma", "lmd"):
            _set3(getattr(p, "pos_" + name), getattr(self._pos_param, name))
        _set3(p.Q_e, self.Q_pos)
        _set3(p.Q_de, self.Q_vel)
        _set3(p.R, self.R)
        # uav_pos_ctrl.py:399: center = concatenate((mean(pos_zone, axis=1), [mean(att_zone[2])]))
        center = np.concatenate((np.mean(np.asarray(up.pos_zone, dtype=float), axis=1),
                                 [np.mean(np.asarray(up.att_zone, dtype=float)[2])]))
        _set3(p.ref_bias_a, center)
        # deterministic trajectory, uav_pos_ctrl.py:419-422
        _set3(p.ref_amplitude, [1.5, 1.5, 0.3, 0.])
        _set3(p.ref_period, [6., 6., 10, 10])
        _set3(p.ref_bias_phase, [np.pi / 2, 0., 0., 0.])
        # random trajectory, uav_pos_ctrl.py:405-408
        _set3(p.traj_A_hi, [1.5, 1.5, 1.5, 0.])
        p.traj_T_lo, p.traj_T_hi, p.traj_phase_hi = 5, 10, 0.
        p.random_trajectory, p.yaw_fixed = int(self.random_trajectory), int(self.yaw_fixed)
        p.random_pos0 = int(self.random_pos0)
        _set3(p.init_pos_r, 0.3 * np.ones(3))                 # uav_pos_ctrl.py:512
        return p
