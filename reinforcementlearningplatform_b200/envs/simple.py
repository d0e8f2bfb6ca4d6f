"""Batched mirrors of the small reference envs (kernels K-FAS, K-SOI, K-BB, K-TLM, K-UGV).

Each class keeps the reference's attribute names and constants (file:line cited per class) and fills the parameter
struct of include/b200env.h; thresholds the reference computes on the fly are evaluated here with the same Python
float expression so the device comparisons see bit-identical operands.
"""
from __future__ import annotations

import math

import numpy as np

from .. import _lib
from ..vec_env import VecEnvBase


def deg2rad(deg):  # utils/functions.py:4-5
    return deg * math.pi / 180.


class _Simple(VecEnvBase):
    def _reset_default(self, mask):
        raise NotImplementedError("reset(random=False): inject the initial state with set_state_buffers()")


class Flight_Attitude_Simulator(_Simple):
    """environment/FlightAttitudeSimulator/FlightAttitudeSimulator.py:9-287.  ``variant='ppo2'`` selects the
    PPO2/DPPO2 demo copy (timeMax = 10, reward Q = 1, R = 0.05; flight_attitude_simulator.py:42,211-224)."""
    ENV_ID = _lib.FAS
    TIMEOUT_FLAG = 3  # terminal_flag of a time-out: success = done and flag != 3 (PPO2-4-FlightAttitudeSimulator/train.py:190)
    OBS_IS_PURE = True
    STATE_FIELDS = ("theta", "dTheta")

    def __init__(self, n_envs: int = 1, variant: str = 'env', **kw):
        self.name = 'Flight_Attitude_Simulator'
        self.f_max, self.f_min = 4, -1.5                       # :21-22
        self.minTheta, self.maxTheta = deg2rad(-60.0), deg2rad(60.0)   # :24-25
        self.min_omega, self.max_omega = deg2rad(-90), deg2rad(90)     # :27-28
        self.dt = 0.02                                         # :38
        self.timeMax = 5 if variant == 'env' else 10           # :42
        self.L, self.J, self.k, self.m, self.dis, self.g = 0.362, 0.082, 0.09, 0.3, 0.3, 9.8   # :45-52
        self.staticGain = 2
        self.Q, self.R = (3., 0.0) if variant == 'env' else (1., 0.05)  # get_reward locals :218-219
        super().__init__(n_envs, **kw)
        self.action_range = np.array([[self.f_min, self.f_max]])
        self.use_norm = True

    def make_params(self):
        p = _lib.FasParams()
        p.L, p.k = self.L, self.k
        p.mgd = self.m * self.g * self.dis                     # self.m * self.g * self.dis   :234
        p.denom = self.J + self.m * self.dis ** 2              # :235
        p.dt, p.time_max = self.dt, self.timeMax
        p.min_theta, p.max_theta, p.min_omega, p.max_omega = self.minTheta, self.maxTheta, self.min_omega, self.max_omega
        p.static_gain = self.staticGain
        p.theta_term_hi = self.maxTheta + deg2rad(1)           # :199
        p.theta_term_lo = self.minTheta - deg2rad(1)           # :203
        p.Q, p.R = self.Q, self.R
        p.reset_lo, p.reset_hi = self.minTheta, self.maxTheta  # :270-271
        return p


class FlightAttitudeSimulatorDiscrete(_Simple):
    """environment/FlightAttitudeSimulator/FlightAttitudeSimulatorDiscrete.py:9-274 (the DQN-family demos).  The action
    handed to ``step_update`` is the force VALUE, an element of ``action_space[0]`` (:56-59), exactly as in the
    reference; ``action_index_to_value`` maps a batch of discrete indices to values on the device."""
    ENV_ID = _lib.FAS_DISCRETE
    TIMEOUT_FLAG = 2
    OBS_IS_PURE = True
    STATE_FIELDS = ("theta", "dTheta")

    def __init__(self, n_envs: int = 1, **kw):
        self.name = 'FlightAttitudeSimulatorDiscrete'
        self.f_max, self.f_min, self.f_step = 3.0, -1.6, 0.1            # :20-22
        self.theta_max, self.dtheta_max = deg2rad(60.0), deg2rad(90)    # :23-24
        self.static_gain = 2.0
        self.dt, self.timeMax = 0.02, 5.0                               # :28,30
        self.J, self.k, self.m, self.g = 0.082, 0.09, 0.3, 9.8          # :32-35
        self.dis, self.L = 0.3, 0.362                                   # :79,83
        super().__init__(n_envs, **kw)
        self.use_norm = True
        self.action_step = [self.f_step]
        self.action_range = [[self.f_min, self.f_max]]
        self.action_num = [int((self.f_max - self.f_min) / self.f_step + 1)]                           # :58
        self.action_space = [[self.f_min + i * self.f_step for i in range(self.action_num[0])]]        # :59
        self.isActionContinuous = False

    def make_params(self):
        p = _lib.FasDiscreteParams()
        p.a2 = -self.k / (self.J + self.m * self.dis ** 2)                                   # :202
        p.a1 = -self.m * self.g * self.dis / (self.dis + self.m * self.dis ** 2)             # :203 (sic)
        p.L, p.denom = self.L, self.J + self.m * self.dis ** 2                               # :204
        p.dt, p.time_max = self.dt, self.timeMax
        p.theta_max, p.dtheta_max, p.static_gain = self.theta_max, self.dtheta_max, self.static_gain
        p.theta_out = self.theta_max + deg2rad(1)                                            # :170,174
        p.Q, p.R = 3., 0.0                                                                   # :233-234
        p.bounce = -0.8
        return p

    def action_index_to_value(self, index):
        """[N] integer tensor of discrete action numbers -> [1, N] force values (``action_space[0][i]``)."""
        import torch
        table = torch.tensor(self.action_space[0], dtype=self.io_dtype, device=self.device)
        return table[index.to(self.device).long()].view(1, -1)


class SecondOrderIntegration(_Simple):
    """environment/SecondOrderIntegration/SecondOrderIntegration.py:13-352.  ``variant='dppo2'``: the DPPO2 demo copy
    (obs * static_gain :213, success terminal disabled :243-246, Q_vel = Q_acc = 0 :260-261)."""
    ENV_ID = _lib.SOI
    TIMEOUT_FLAG = 2  # terminal_flag of a time-out: success = done and flag != 2 (PPO2-4-SecondOrderIntegration/train.py:205)
    OBS_IS_PURE = True
    STATE_FIELDS = ("x", "y", "vx", "vy")

    def __init__(self, n_envs: int = 1, map_size=(5.0, 5.0), target=None, variant: str = 'env', **kw):
        self.name = 'SecondOrderIntegration'
        self.map_size = np.array(map_size, dtype=float)
        self.init_target = self.map_size / 2 if target is None else np.array(target, dtype=float)  # :333
        self.mass, self.fMax, self.fMin, self.admissible_error, self.vMax = 1.0, 3, -3, 0, 3      # :29-37
        self.k, self.dt, self.time_max = 0.15, 0.02, 5.0       # :39-42
        self.static_gain = 2
        self._variant = variant
        super().__init__(n_envs, **kw)
        self.action_range = np.array([[self.fMin, self.fMax], [self.fMin, self.fMax]])
        self.use_norm = True

    def make_params(self):
        p = _lib.SoiParams()
        p.map_x, p.map_y = self.map_size
        p.target_x, p.target_y = self.init_target
        p.mass, p.k, p.vmax, p.dt, p.time_max = self.mass, self.k, self.vMax, self.dt, self.time_max
        p.admissible_error = self.admissible_error
        dppo2 = self._variant == 'dppo2'
        p.obs_gain = self.static_gain if dppo2 else 1.0
        p.Q_pos, p.Q_vel, p.Q_acc = (1, 0.0, 0.0) if dppo2 else (1, 0.1, 0.05)   # :262-264
        p.reset_margin = 0.1                                   # :329-331
        p.success_terminal = 0 if dppo2 else 1
        return p


class BallBalancer1D(_Simple):
    """environment/BallBalancer/BallBalancer1D.py:14-322."""
    ENV_ID = _lib.BALLBALANCER
    TIMEOUT_FLAG = 2  # terminal_flag of a time-out: success = done and flag != 2 (PPO2-4-BallBalancer1D/train.py:201)
    OBS_IS_PURE = True
    STATE_FIELDS = ("pos", "vel", "theta", "error")

    def __init__(self, n_envs: int = 1, initVel: float = 0.0, target: float = 0.0, **kw):
        self.name = 'BallBalancer1D'
        self.initVel, self.target = initVel, target
        self.omegaMax, self.omegaMin = np.pi, -np.pi           # :34-35
        self.thetaMax, self.thetaMin = deg2rad(45.0), deg2rad(-45.0)
        self.vMin, self.vMax = -3, 3
        self.dt, self.timeMax = 0.02, 8                        # :47-51
        self.m, self.rBall, self.rMotor, self.L, self.g, self.J = 0.26, 0.02, 0.0245, 0.134, 9.81, 0.0000416  # :53-60
        self.K = (self.m * self.g * self.rBall ** 2 * self.rMotor) / ((self.m * self.rBall ** 2 + self.J) * self.L)  # :61-62
        self.staticGain = 2
        super().__init__(n_envs, **kw)
        self.action_range = np.array([[self.omegaMin, self.omegaMax]])
        self.use_norm = True

    def make_params(self):
        p = _lib.BallBalancerParams()
        p.K, p.L = self.K, self.L
        p.omega_min, p.omega_max, p.theta_min, p.theta_max = self.omegaMin, self.omegaMax, self.thetaMin, self.thetaMax
        p.v_min, p.v_max, p.dt, p.time_max = self.vMin, self.vMax, self.dt, self.timeMax
        p.static_gain, p.target, p.deg1 = self.staticGain, self.target, deg2rad(1)
        p.reset_theta_lo, p.reset_theta_hi = deg2rad(-40), deg2rad(40)   # :294
        p.reset_pos_lo, p.reset_pos_hi = -0.12, 0.12                     # :295
        p.init_vel = self.initVel
        return p


class TwoLinkManipulator(_Simple):
    """environment/RobotManipulator/TwoLinkManipulator.py:8-312."""
    ENV_ID = _lib.TWOLINK
    TIMEOUT_FLAG = 2  # terminal_flag of a time-out: success = done and flag != 2 (PPO2-4-TwoLinkManipulator/train.py:191)
    STATE_FIELDS = ("theta1", "theta2", "omega1", "omega2", "err_x", "err_y", "target_x", "target_y")

    def __init__(self, n_envs: int = 1, **kw):
        self.name = 'TwoLinkManipulator'
        self.basePos = np.array([1.0, 1.0])                    # :32
        self.l, self.m, self.g = 0.35, 0.5, 9.8                # :33-35
        self.J = self.m * (self.l ** 2) / 3                    # :36
        self.dt, self.time_max = 0.02, 8.0                     # :37-39
        self.thetaMax, self.omegaMax, self.torqueMax, self.miss = np.pi, np.pi, 5.0, 0.01   # :42-47
        self.init_endPos = np.array([1.0, 0.3])                # :19
        self.static_gain = 2
        super().__init__(n_envs, **kw)
        self.action_range = np.array([[-self.torqueMax, self.torqueMax], [-self.torqueMax, self.torqueMax]])
        self.use_norm = True

    def make_params(self):
        p = _lib.TwoLinkParams()
        p.l, p.m, p.g, p.J, p.dt, p.time_max = self.l, self.m, self.g, self.J, self.dt, self.time_max
        p.base_x, p.base_y = self.basePos
        p.theta_max, p.miss, p.omega_ok = self.thetaMax, self.miss, deg2rad(5)   # :194-197
        p.init_end_x, p.init_end_y = self.init_endPos
        p.r2_lo, p.r2_hi = 0.3 ** 2, (2 * self.l) ** 2         # :286
        p.Q_pos, p.Q_omega, p.Q_acc = 2.0, 0.1, 0.005          # :212-214
        return p


class UGVForward(_Simple):
    """environment/UGV/UGVForward.py:10-362."""
    ENV_ID = _lib.UGV
    TIMEOUT_FLAG = 2  # terminal_flag of a time-out: success = done and flag != 2 (PPO2-4-UGVForward/train.py:201)
    OBS_IS_PURE = True
    STATE_FIELDS = ("x", "y", "vel", "phi", "omega")
    BIDIRECTIONAL = 0

    def __init__(self, n_envs: int = 1, map_size=(5.0, 5.0), target=(2.5, 2.5), time_max: float = 10.0, **kw):
        self.name = 'UGVBidirectional' if self.BIDIRECTIONAL else 'UGVForward'
        self.map_size = np.array(map_size, dtype=float)
        self.init_target = np.array(target, dtype=float)
        self.dt, self.time_max, self.kf, self.kt = 0.02, time_max, 0.1, 0.1   # :43-48
        self.e_max = np.linalg.norm(self.map_size) / 2         # :53
        self.v_max, self.e_phi_max, self.omega_max = 3, np.pi, 2 * np.pi
        self.a_linear_max, self.a_angular_max = 3, 2 * np.pi
        self.static_gain = 1.
        super().__init__(n_envs, **kw)
        self.action_range = np.array([[-self.a_linear_max, self.a_linear_max], [-self.a_angular_max, self.a_angular_max]])
        self.use_norm = True

    def make_params(self):
        p = _lib.UgvParams()
        p.map_x, p.map_y = self.map_size
        p.target_x, p.target_y = self.init_target
        p.dt, p.time_max, p.kf, p.kt = self.dt, self.time_max, self.kf, self.kt
        p.e_max, p.v_max, p.e_phi_max, p.omega_max = float(self.e_max), self.v_max, self.e_phi_max, self.omega_max
        p.static_gain = self.static_gain
        p.Q_pos, p.Q_vel, p.Q_phi, p.Q_omega = 2., 0.0, 2., 1.0   # :264-267
        p.reset_d0 = 0.5                                       # :335
        p.bidirectional = self.BIDIRECTIONAL
        return p


class UGVBidirectional(UGVForward):
    """environment/UGV/UGVBidirectional.py:10-367 (signed position error, folded heading error, no v >= 0 clamp)."""
    BIDIRECTIONAL = 1


class UGVForwardObstacleAvoidance(_Simple):
    """environment/UGVForwardObstacleAvoidance/UGVForwardObstacleAvoidance.py:11-557 (``variant='env'``: dt = 0.1, 10
    circles) or the PPO2 / DPPO2 demo copies (``variant='ppo2'``: dt = 0.05, 10 circles; ``'dppo2'``: 15 circles;
    progress reward, success ignores omega, pose frozen when the previous velocity was negative).  Observation =
    4 kinematic terms + 37 fake-laser ranges (kernel K-UGVO: one warp per instance, lanes = rays)."""
    ENV_ID = _lib.UGVO
    TIMEOUT_FLAG = 2  # terminal_flag of a time-out: success = done and flag != 2 (PPO2-4-UGVForwardObstacleAvoidance/train.py:201)
    OBS_IS_PURE = True
    USES_WORK_LIST = True
    MAX_OBS = 16
    STATE_FIELDS = tuple(["x", "y", "vel", "phi", "omega", "target_x", "target_y", "n_obs"] +
                         [f"{c}_{k}" for k in range(16) for c in ("cx", "cy", "r")])

    def __init__(self, n_envs: int = 1, map_size=(5.0, 5.0), variant: str = 'env', obsNum: int = None, **kw):
        if variant not in ('env', 'ppo2', 'dppo2'):
            raise ValueError("variant must be 'env', 'ppo2' or 'dppo2'")
        self.name = 'UGVForwardObstacleAvoidance'
        self._variant = variant
        self.map_size = np.array(map_size, dtype=float)
        self.r_vehicle = 0.15                                  # :38
        self.dt = 0.1 if variant == 'env' else 0.05            # :42 / demo copy :40
        self.time_max, self.kf, self.kt = 15.0, 0.1, 0.1       # :44-47
        self.laserDis, self.laserBlind = 2.0, 0.0              # :49-50
        self.laserRange, self.laserStep = deg2rad(90), deg2rad(5)   # :51-52
        self.laserState = int(2 * self.laserRange / self.laserStep) + 1   # :53  (= 37)
        self.e_max = np.linalg.norm(self.map_size) / 2         # :58
        self.v_max, self.e_phi_max, self.omega_max = 3, np.pi, 2 * np.pi
        self.a_linear_max, self.a_angular_max = 3, 2 * np.pi
        self.static_gain = 1.
        self.obsNum = obsNum if obsNum is not None else (15 if variant == 'dppo2' else 10)   # reset() :535 / demo :543
        if not 0 < self.obsNum <= self.MAX_OBS:
            raise ValueError("obsNum must be in 1..16")
        super().__init__(n_envs, **kw)
        self.action_range = np.array([[-self.a_linear_max, self.a_linear_max], [-self.a_angular_max, self.a_angular_max]])
        self.use_norm = True

    def make_params(self):
        p = _lib.UgvoParams()
        p.map_x, p.map_y = self.map_size
        p.dt, p.time_max, p.kf, p.kt = self.dt, self.time_max, self.kf, self.kt
        p.e_max, p.v_max, p.e_phi_max, p.omega_max = float(self.e_max), self.v_max, self.e_phi_max, self.omega_max
        p.static_gain, p.r_vehicle = self.static_gain, self.r_vehicle
        p.laser_dis, p.laser_blind, p.laser_range = self.laserDis, self.laserBlind, self.laserRange
        p.Q_pos, p.Q_vel, p.Q_phi, p.Q_omega = 2., 0.0, 2., 1.0            # :452-455
        p.safety_dis_obs = p.safety_dis_st = 4 * self.r_vehicle             # :531-532
        p.r_min, p.r_max, p.st_margin = 0.2, 0.5, 0.3                       # :533-534, map.py:67
        p.n_rays, p.obs_num = self.laserState, self.obsNum
        p.variant = 0 if self._variant == 'env' else 1
        return p
