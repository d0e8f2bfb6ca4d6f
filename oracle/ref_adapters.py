"""Adapters that drive the UNMODIFIED reference env classes for fixture generation.

TEST INFRASTRUCTURE ONLY (container-side, needs /root/reference).  One adapter per
env variant: how to build it, reset it, read its internal state in the engine's SoA
field order, and run one ``step_update`` the way the reference's train.py does.
"""
from __future__ import annotations

import numpy as np

from . import ref_shim as R


class Adapter:
    name = ""
    cites = ""
    F = S = A = D = 0
    action_lo = action_hi = None

    def make(self):
        raise NotImplementedError

    def reset(self, env):
        env.reset(True)

    def internal(self, env):
        raise NotImplementedError

    def sample_action(self, rng, t, l, env=None):
        return rng.uniform(self.action_lo, self.action_hi)

    def sample_dis(self, rng, t, l):
        return np.zeros(self.D)

    def step(self, env, a, d):
        env.step_update(np.array(a, dtype=float))
        return (np.array(env.current_state, dtype=float), np.array(env.next_state, dtype=float),
                float(env.reward), bool(env.is_terminal), int(env.terminal_flag))

    def params_json(self):
        return {}


# --------------------------------------------------------------------- CartPole
class CartPoleA(Adapter):
    name = "cartpole"
    cites = "environment/CartPole/CartPole.py:145-295"
    F, S, A, D = 4, 4, 1, 0
    action_lo, action_hi = np.array([-8.]), np.array([8.])

    def make(self):
        return R.load("environment.CartPole.CartPole").CartPole(0., 0.)

    def internal(self, env):
        return np.array([env.theta, env.dtheta, env.x, env.dx], dtype=float), float(env.time)


class CartPoleGentleA(CartPoleA):
    """Stabilising feedback + noise: long episodes that reach the time-out flag (3) and the late-time
    10|11 sub-step pattern of the `while self.time < tt` loop (note N1)."""
    name = "cartpole_gentle"

    def sample_action(self, rng, t, l, env=None):
        k = 40.0 * env.theta + 6.0 * env.dtheta + 2.0 * env.x + 3.0 * env.dx  # only to keep the pole up
        return np.clip(np.array([k]) + rng.uniform(-0.5, 0.5, 1), -8, 8)


class CartPoleAngleOnlyEnvA(Adapter):
    name = "cartpole_angleonly_env"
    cites = "environment/CartPole/CartPoleAngleOnly.py:139-299"
    F, S, A, D = 4, 2, 1, 0
    action_lo, action_hi = np.array([-8.]), np.array([8.])

    def make(self):
        return R.load("environment.CartPole.CartPoleAngleOnly").CartPoleAngleOnly(0.)

    def internal(self, env):
        return np.array([env.theta, env.dtheta, env.x, env.dx], dtype=float), float(env.time)


class CartPoleAngleOnlyPPO2A(CartPoleAngleOnlyEnvA):
    name = "cartpole_angleonly_ppo2"
    cites = "demonstration/PPO2/PPO2-4-CartPoleAngleOnly/cartpole_angleonly.py:137-279"
    action_lo, action_hi = np.array([-5.]), np.array([5.])

    def make(self):
        m = R.load_file("demonstration/PPO2/PPO2-4-CartPoleAngleOnly/cartpole_angleonly.py", "ref_cartpole_angleonly_ppo2")
        return m.CartPoleAngleOnly(0.)


# name -> (factory, lanes, steps, seed)
REGISTRY = {
    "cartpole": (CartPoleA, 8, 1000, 1),
    "cartpole_gentle": (CartPoleGentleA, 2, 600, 11),
    "cartpole_angleonly_env": (CartPoleAngleOnlyEnvA, 4, 1000, 2),
    "cartpole_angleonly_ppo2": (CartPoleAngleOnlyPPO2A, 4, 1000, 3),
}
