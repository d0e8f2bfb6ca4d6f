"""Batched UavRobust envs (kernel K-UAVR, csrc/uavrobust.cu): the quadrotor of ``environment/UavRobust/uav.py`` wrapped
as ``uav_hover_outer_loop``, ``uav_hover``, ``uav_inner_loop`` and ``uav_tracking_outer_loop``.  Constructor arguments
are duck-typed ``uav_param`` / ``fntsmc_param`` objects with the reference's attribute names."""
from __future__ import annotations

import numpy as np
from numpy import deg2rad

from .. import _lib
from ..vec_env import VecEnvBase
from .uav import _set3, fntsmc_param, uav_param


def robust_uav_param() -> uav_param:
    """Quadrotor parameters of the UavRobust training scripts (PPO-4-UavHoverOuterLoop/train.py:24-41)."""
    p = uav_param()
    p.dt, p.time_max = 0.01, 10
    p.pos_zone = np.atleast_2d([[-5, 5], [-5, 5], [0, 5]])
    p.att_zone = np.atleast_2d([[deg2rad(-45), deg2rad(45)], [deg2rad(-45), deg2rad(45)], [deg2rad(-120), deg2rad(120)]])
    return p


def robust_att_ctrl_param() -> fntsmc_param:
    """PPO-4-UavHoverOuterLoop/train.py:45-56."""
    p = fntsmc_param()
    p.k1 = np.array([25., 25., 40.])
    p.k2 = np.array([0.1, 0.1, 0.2])
    p.alpha = np.array([2.5, 2.5, 2.5])
    p.beta = np.array([0.99, 0.99, 0.99])
    p.gamma = np.array([1.5, 1.5, 1.2])
    p.lmd = np.array([2.0, 2.0, 2.0])
    p.dt = 0.01
    p.saturation = np.array([0.3, 0.3, 0.3])
    return p


class _UavRobustBase(VecEnvBase):
    ENV_ID = _lib.UAVROBUST
    TIMEOUT_FLAG = 1  # uav.py:543-560: flag 1 = time out
    STATE_FIELDS = tuple("x y z vx vy vz phi theta psi p q r s1_0 s1_1 s1_2 aref_0 aref_1 aref_2 daref_0 daref_1 daref_2 "
                         "pref_0 pref_1 pref_2 A_0 A_1 A_2 T_0 T_1 T_2 phase_0 phase_1 phase_2".split())
    Q = (1, 0.1, 0.02)
    msg_print_flag = True  # uav.py:109 (the train scripts switch the terminal prints off; the engine never prints)

    def __init__(self, n_envs: int = 1, UAV_param: uav_param = None, att_ctrl_param: fntsmc_param = None, **kw):
        self._up = UAV_param or robust_uav_param()
        self._att = att_ctrl_param or robust_att_ctrl_param()
        self.static_gain = 1.0
        self.e_pos_max, self.e_pos_min = np.array([5., 5., 5.]), -np.array([5., 5., 0.])
        self.vel_max, self.vel_min = np.array([3., 3., 3.]), -np.array([3., 3., 3.])
        self.dot_att_min = np.array([-deg2rad(60), -deg2rad(60), -deg2rad(1)])
        self.dot_att_max = np.array([deg2rad(60), deg2rad(60), deg2rad(1)])
        self.e_att_max = np.array([deg2rad(60), deg2rad(60), deg2rad(120)])
        self.e_att_min = -np.array([deg2rad(60), deg2rad(60), deg2rad(120)])
        self.e_dot_att_max = np.array([deg2rad(60), deg2rad(60), deg2rad(120)])
        self.e_dot_att_min = -np.array([deg2rad(60), deg2rad(60), deg2rad(120)])
        self.u_min, self.u_max, self.torque_min, self.torque_max = -8, 8, -0.3, 0.3
        super().__init__(n_envs, **kw)
        self.use_norm = True

    @property
    def dt(self):
        return self._up.dt

    @property
    def time_max(self):
        return self._up.time_max

    def make_params(self):
        up, att = self._up, self._att
        p = _lib.UavRobustParams()
        p.m, p.g, p.kr, p.kt, p.dt, p.time_max = up.m, up.g, up.kr, up.kt, up.dt, up.time_max
        p.t_term = up.time_max - up.dt / 2                      # uav.py:545
        _set3(p.J, up.J)
        pz, az = np.asarray(up.pos_zone, dtype=float), np.asarray(up.att_zone, dtype=float)
        _set3(p.pos_zone_min, pz[:, 0]); _set3(p.pos_zone_max, pz[:, 1])
        _set3(p.att_zone_min, az[:, 0]); _set3(p.att_zone_max, az[:, 1])
        _set3(p.pos0, up.pos0); _set3(p.vel0, up.vel0); _set3(p.angle0, up.angle0); _set3(p.pqr0, up.pqr0)
        for name in ("k1", "k2", "alpha", "beta", "gamma", "lmd"):
            _set3(getattr(p, "att_" + name), getattr(att, name))
        _set3(p.att_saturation, getattr(att, "saturation", np.zeros(3)))
        _set3(p.e_pos_span, self.e_pos_max - self.e_pos_min)
        _set3(p.vel_span, self.vel_max - self.vel_min)
        _set3(p.e_att_span, self.e_att_max - self.e_att_min)
        _set3(p.e_dot_att_span_neg, self.e_dot_att_min - self.e_dot_att_max)   # sic, UavHover.py:109 (N10)
        _set3(p.dot_att_min, self.dot_att_min); _set3(p.dot_att_max, self.dot_att_max)
        p.static_gain = self.static_gain
        p.Qx, p.Qv, p.R = self.Q
        _set3(p.target_lo, pz[:, 0] + 1.0); _set3(p.target_hi, pz[:, 1] - 1.0)   # generate_random_point(offset=1.0)
        p.variant = self.VARIANT
        return p


class uav_hover_outer_loop(_UavRobustBase):
    """environment/UavRobust/UavHoverOuterLoop.py:18-228: action = virtual acceleration (3, +-8), inner loop by FNTSMC."""
    VARIANT = 0

    def __init__(self, n_envs: int = 1, **kw):
        super().__init__(n_envs, **kw)
        self.name = 'uav_hover_outer_loop'
        self.action_range = [[self.u_min, self.u_max]] * 3


class uav_hover(_UavRobustBase):
    """environment/UavRobust/UavHover.py:18-254: obs 12, action = acceleration (3) + torque (3)."""
    VARIANT = 1

    def __init__(self, n_envs: int = 1, **kw):
        super().__init__(n_envs, **kw)
        self.name = 'uav_hover'
        self.action_range = [[self.u_min, self.u_max]] * 3 + [[self.torque_min, self.torque_max]] * 3


class uav_inner_loop(_UavRobustBase):
    """environment/UavRobust/UavInnerLoop.py:18-224: attitude tracking, action = torque (3, +-0.3)."""
    VARIANT = 2
    Q = (1, 0.1, 0.01)

    def __init__(self, n_envs: int = 1, **kw):
        super().__init__(n_envs, **kw)
        self.name = 'uav_inner_loop'
        self.action_range = [[self.torque_min, self.torque_max]] * 3

    def make_params(self):
        p = super().make_params()
        az = np.asarray(self._up.att_zone, dtype=float)
        # generate_random_signal, UavInnerLoop.py:197-210
        _set3(p.sig_A_hi, [az[0][1] if az[0][1] < np.pi / 3 else np.pi / 3, az[1][1] if az[1][1] < np.pi / 3 else np.pi / 3,
                           az[2][1] if az[2][1] < np.pi / 2 else np.pi / 2])
        p.sig_T_lo, p.sig_T_hi, p.sig_phase_hi = 3, 6, np.pi / 2
        return p


class uav_tracking_outer_loop(_UavRobustBase):
    """environment/UavRobust/UavTrackingOuterLoop.py:18-270: position tracking of a sinusoidal reference."""
    VARIANT = 3
    Q = (1, 0.1, 0.01)

    def __init__(self, n_envs: int = 1, **kw):
        super().__init__(n_envs, **kw)
        self.name = 'uav_tracking_outer_loop'
        self.action_range = [[self.u_min, self.u_max]] * 3

    def make_params(self):
        p = super().make_params()
        pz = np.asarray(self._up.pos_zone, dtype=float)
        center = np.mean(pz, axis=1)                             # UavTrackingOuterLoop.py:228
        _set3(p.ref_bias_a, center)
        _set3(p.sig_A_hi, [pz[k][1] - center[k] - 1 for k in range(3)])   # :236-238
        p.sig_T_lo, p.sig_T_hi, p.sig_phase_hi = 5, 10, np.pi / 2
        p.init_pos_r = 0.3                                       # :200
        return p
