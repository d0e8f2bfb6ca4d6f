"""Batched CartPole family (kernel K-CP, csrc/cartpole.cu).

Mirrors ``environment/CartPole/CartPole.py``, ``CartPoleAngleOnly.py`` and the
PPO2 demo copy ``demonstration/PPO2/PPO2-4-CartPoleAngleOnly/cartpole_angleonly.py``.
Attribute names are the reference's (``theta_max``/``thetaMax``, ``fm``, ``timeMax`` ...).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib
from ..vec_env import VecEnvBase


def deg2rad(deg):  # utils/functions.py:4-5 (same expression, same rounding)
    return deg * math.pi / 180.


class CartPole(VecEnvBase):
    """CartPole.py:11-295.  obs = (theta, dtheta, x, dx) / max * 2; flags 1 angle, 2 position, 3 time, 4 success."""
    ENV_ID = _lib.CARTPOLE
    TIMEOUT_FLAG = 3  # success = done and flag != 3 (PPO2-4-CartPoleAngleOnly/train.py:198-205)
    OBS_IS_PURE = True
    VARIANT = 0
    STATE_FIELDS = ("theta", "dtheta", "x", "dx")

    def __init__(self, n_envs: int = 1, initTheta: float = 0., initX: float = 0., reset_theta_frac: float = 0.5, **kw):
        self.initTheta, self.initX = initTheta, initX
        self.theta_max = deg2rad(45)          # CartPole.py:26
        self.dtheta_max = deg2rad(90)         # :27
        self.x_max = 1.5                      # :28
        self.dx_max = 3                       # :29
        self.staticGain = 2.0                 # :31
        self.M, self.m, self.g, self.ell, self.kf, self.fm = 1.0, 0.1, 9.8, 0.2, 0.2, 8  # :33-38
        self.dt = 0.02                        # :40
        self.timeMax = 5                      # :41
        self.name = 'CartPole'
        # DPPO2 demo copy draws theta0 from +-theta_max/3 (DPPO2-4-CartPole/CartPole.py:273-274)
        self.reset_theta_frac = reset_theta_frac
        super().__init__(n_envs, **kw)
        self.action_range = np.array([[-self.fm, self.fm]])  # :63
        self.use_norm = True

    def make_params(self):
        p = _lib.CartPoleParams()
        p.M, p.m, p.g, p.ell, p.kf = self.M, self.m, self.g, self.ell, self.kf
        p.dt, p.time_max = self.dt, self.timeMax
        p.theta_max, p.dtheta_max, p.x_max, p.dx_max = self.theta_max, self.dtheta_max, self.x_max, self.dx_max
        p.static_gain, p.norm_boundless = self.staticGain, 4
        p.theta_term_hi = self.theta_max + deg2rad(1)        # CartPole.py:167
        p.theta_term_lo = -self.dtheta_max - deg2rad(1)      # CartPole.py:167 (sic: dtheta_max, note N2)
        p.reset_theta_lo = -self.theta_max * self.reset_theta_frac   # CartPole.py:272
        p.reset_theta_hi = self.theta_max * self.reset_theta_frac
        p.reset_x_lo, p.reset_x_hi = -self.x_max * 0.5, self.x_max * 0.5  # :273
        p.variant = self.VARIANT
        return p

    def _reset_default(self, mask):
        sel = slice(None) if mask is None else mask.to(self.device).bool()
        self._state[0, sel] = self.initTheta
        self._state[1, sel] = 0.
        self._state[2, sel] = self.initX
        self._state[3, sel] = 0.
        self._time[sel] = 0.


class CartPoleAngleOnly(CartPole):
    """CartPoleAngleOnly.py:11-299 (``variant='env'``: dt=0.01, timeMax=6, fm=8, time-loop RK4, sign-of-progress
    reward) or the PPO2/DPPO2 demo copy cartpole_angleonly.py:11-279 (``variant='ppo2'``: dt=0.02, timeMax=5, fm=5,
    one RK4 step of size dt, quadratic reward, success flag 4)."""
    STATE_FIELDS = ("theta", "dtheta", "x", "dx")

    def __init__(self, n_envs: int = 1, initTheta: float = 0., variant: str = 'env', **kw):
        if variant not in ('env', 'ppo2'):
            raise ValueError("variant must be 'env' or 'ppo2'")
        self.VARIANT = 1 if variant == 'env' else 2
        self._variant_name = variant
        super().__init__(n_envs, initTheta=initTheta, initX=0., **kw)
        self.name = 'CartPoleAngleOnly'
        self.action_range = np.array([[-self.fm, self.fm]])

    def make_params(self):
        self.thetaMax = deg2rad(45)
        self.norm_4_boundless_state = 4
        if self._variant_name == 'env':
            self.fm, self.dt, self.timeMax = 8, 0.01, 6       # CartPoleAngleOnly.py:36-39
        else:
            self.fm, self.dt, self.timeMax = 5, 0.02, 5       # cartpole_angleonly.py:37-40
        p = super().make_params()
        p.theta_max = self.thetaMax
        p.norm_boundless = self.norm_4_boundless_state
        p.theta_term_hi = self.thetaMax + deg2rad(1)         # CartPoleAngleOnly.py:150
        p.theta_term_lo = -self.thetaMax - deg2rad(1)
        p.reset_theta_lo, p.reset_theta_hi = -self.thetaMax / 2, self.thetaMax / 2  # :273
        p.reset_x_lo = p.reset_x_hi = 0.
        p.variant = self.VARIANT
        return p
