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

    def sample_dis(self, rng, t, l, env=None):
        return np.zeros(self.D)

    def step(self, env, a, d):
        env.step_update(np.array(a, dtype=float))
        return (np.array(env.current_state, dtype=float), np.array(env.next_state, dtype=float),
                float(env.reward), bool(env.is_terminal), int(env.terminal_flag))

    def params_json(self):
        return {}

    #: attributes nudged by about one ulp (+-2.3e-16 * max(1, |v|)) in the twin instance (self-sensitivity measurement)
    perturb_attrs = ()
    EPS = 2.3e-16

    def perturb(self, env, k):
        for j, name in enumerate(self.perturb_attrs):
            v = float(getattr(env, name))
            setattr(env, name, v + (self.EPS if (k + j) % 2 == 0 else -self.EPS) * max(1.0, abs(v)))


# --------------------------------------------------------------------- CartPole
class CartPoleA(Adapter):
    name = "cartpole"
    cites = "environment/CartPole/CartPole.py:145-295"
    F, S, A, D = 4, 4, 1, 0
    action_lo, action_hi = np.array([-8.]), np.array([8.])
    perturb_attrs = ("theta", "dtheta", "dx")

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
    perturb_attrs = ("theta", "dtheta", "dx")

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


# ------------------------------------------------------------- UavFntsmcParam
def _uav_train_params(kind):
    """Parameter objects exactly as the PPO2 training scripts build them
    (PPO2-4-UavFntsmcParamAtt/train.py:29-62, PPO2-4-UavFntsmcParamPos/train.py:29-76)."""
    R.use_family("UavFntsmcParam")
    uav = R.load("uav")
    fn = R.load("FNTSMC")
    F = R.load("utils.functions")
    up = uav.uav_param()
    up.dt, up.time_max = 0.02, 10
    if kind == "att":
        up.pos_zone = np.atleast_2d([[-np.inf, np.inf], [-np.inf, np.inf], [-np.inf, np.inf]])
        up.att_zone = np.atleast_2d([[F.deg2rad(-90), F.deg2rad(90)], [F.deg2rad(-90), F.deg2rad(90)], [F.deg2rad(-180), F.deg2rad(180)]])
    else:
        up.pos_zone = np.atleast_2d([[-3, 3], [-3, 3], [0, 3]])
        up.att_zone = np.atleast_2d([[F.deg2rad(-90), F.deg2rad(90)], [F.deg2rad(-90), F.deg2rad(90)], [F.deg2rad(-120), F.deg2rad(120)]])
    att = fn.fntsmc_param()
    att.k1 = np.array([25., 25., 40.]); att.k2 = np.array([0.1, 0.1, 0.2]); att.alpha = np.array([2.5, 2.5, 2.5])
    att.beta = np.array([0.99, 0.99, 0.99]); att.gamma = np.array([1.5, 1.5, 1.2]); att.lmd = np.array([2.0, 2.0, 2.0])
    att.dim, att.dt, att.ctrl0 = 3, 0.02, np.array([0., 0., 0.])
    pos = fn.fntsmc_param()
    pos.k1 = np.array([1.2, 0.8, 0.5]); pos.k2 = np.array([0.2, 0.6, 0.5]); pos.alpha = np.array([1.2, 1.5, 1.2])
    pos.beta = np.array([0.3, 0.3, 0.5]); pos.gamma = np.array([0.2, 0.2, 0.2]); pos.lmd = np.array([2.0, 2.0, 2.0])
    pos.dim, pos.dt, pos.ctrl0 = 3, 0.02, np.array([0., 0., 0.])
    return up, att, pos


def _zero(p):
    p.k1 = 0.01 * np.ones(3); p.k2 = 0.01 * np.ones(3); p.gamma = 0.01 * np.ones(3); p.lmd = 0.01 * np.ones(3)


class UavPosA(Adapter):
    """Loop body of PPO2-4-UavFntsmcParamPos/train.py:273-297 around uav_pos_ctrl_RL."""
    name = "uav_pos"
    cites = ("environment/UavFntsmcParam/uav_pos_ctrl_RL.py:59-173, uav_pos_ctrl.py:302-376,467-533, uav.py:93-219, "
             "FNTSMC.py:47-69,112-137, ref_cmd.py:25-43")
    F, S, A, D = 51, 6, 8, 3
    perturb_attrs = ("vx", "vy", "vz", "p", "q", "r")
    random_trajectory = True
    with_dis = False

    def make(self):
        self.up, self.att, self.pos = _uav_train_params("pos")
        _zero(self.pos)
        m = R.load("uav_pos_ctrl_RL")
        self.ref_cmd = R.load("ref_cmd")
        return m.uav_pos_ctrl_RL(self.up, self.att, self.pos)

    def reset(self, env):
        _zero(self.pos)  # reset_pos_ctrl_param('zero'), train.py:84-89,275
        env.reset_uav_pos_ctrl_RL_tracking(random_trajectroy=self.random_trajectory, random_pos0=False,
                                           new_att_ctrl_param=None, new_pos_ctrl_parma=self.pos, outer_param=None)

    def internal(self, env):
        s = np.concatenate((env.uav_state_call_back(), env.pos_ctrl.sigma_o1, env.att_ctrl.s1, env.att_ref,
                            env.pos_ctrl.k1, env.pos_ctrl.k2, env.pos_ctrl.gamma, env.pos_ctrl.lmd,
                            env.ref_amplitude, env.ref_period, env.ref_bias_phase, env.pos_ref, env.dot_pos_ref))
        return np.array(s, dtype=float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        kind = l % 3
        if kind == 0:      # fresh gains over the whole action range every step
            return rng.uniform(0., 5., 8)
        if kind == 1:      # near the hand-tuned gains: long episodes that reach the time-out flag
            base = np.array([1.2, 0.8, 0.5, 0.2, 0.6, 0.5, 0.2, 2.0])
            return base * rng.uniform(0.8, 1.2, 8)
        a = rng.uniform(0., 3., 8)  # some entries <= 0: those gains keep their previous value (note N6)
        a[rng.random(8) < 0.3] = 0.
        return a

    def sample_dis(self, rng, t, l, env=None):
        if not self.with_dis:
            return np.zeros(3)
        tt = 0.02 * t  # UavRobust/ref_cmd.py:46-61 style sinusoids + noise
        return np.array([0.5 * np.sin(1.3 * tt) + 0.2, 0.4 * np.cos(0.9 * tt), 0.3 * np.sin(2.1 * tt + 1.0)]) + rng.normal(0, 0.05, 3)

    def step(self, env, a, d):
        a = np.array(a, dtype=float)
        env.dis = np.array(d, dtype=float)
        env.get_param_from_actor(a)
        action_4_uav = env.generate_action_4_uav()
        env.step_update(action_4_uav)
        return (np.array(env.current_state, dtype=float), np.array(env.next_state, dtype=float),
                float(env.reward), bool(env.is_terminal), int(env.terminal_flag))


class UavPosDisA(UavPosA):
    name = "uav_pos_dis"
    with_dis = True


class UavPosRandomPos0A(UavPosA):
    """reset_uav_pos_ctrl(random_pos0=True): uav_pos_ctrl.py:510-513, set_random_init_pos :457-465, uav.py:252-268.
    State layout variant 1: the trailing 3 fields are the reference's init_state[9:12] (note N5)."""
    name = "uav_pos_rp0"
    F = 54

    def reset(self, env):
        _zero(self.pos)
        env.reset_uav_pos_ctrl_RL_tracking(random_trajectroy=True, random_pos0=True, new_att_ctrl_param=None,
                                           new_pos_ctrl_parma=self.pos, outer_param=None)

    def internal(self, env):
        s, t = super().internal(env)
        return np.concatenate((s, np.array(env.init_state[9:12], dtype=float))), t

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:  # weak random gains: short episodes, many resets (each loads p, q, r from the previous pos0)
            return rng.uniform(0., 1.0, 8)
        return super().sample_action(rng, t, 1, env)


class UavAttA(Adapter):
    """Loop body of PPO2-4-UavFntsmcParamAtt/train.py:254-276 around uav_att_ctrl_RL."""
    name = "uav_att"
    cites = ("environment/UavFntsmcParam/uav_att_ctrl_RL.py:59-178, uav_att_ctrl.py:91-216, uav.py:93-219,285-360, "
             "FNTSMC.py:112-137, ref_cmd.py:4-22")
    F, S, A, D = 36, 6, 8, 0
    perturb_attrs = ("p", "q", "r")
    random_trajectory = False

    def make(self):
        self.up, self.att, _ = _uav_train_params("att")
        _zero(self.att)
        m = R.load("uav_att_ctrl_RL")
        self.ref_cmd = R.load("ref_cmd")
        return m.uav_att_ctrl_RL(self.up, self.att)

    def reset(self, env):
        _zero(self.att)  # reset_att_ctrl_param('zero'), train.py:69-74,256
        env.reset_uav_att_ctrl_RL_tracking(random_trajectory=self.random_trajectory, yaw_fixed=False,
                                           new_att_ctrl_param=self.att)

    def internal(self, env):
        s = np.concatenate((env.uav_att_pqr_call_back(), env.att_ctrl.s1, env.att_ctrl.k1, env.att_ctrl.k2,
                            env.att_ctrl.gamma, env.att_ctrl.lmd, env.ref_att_amplitude, env.ref_att_period,
                            env.ref_att_bias_phase, env.ref, env.dot_ref))
        return np.array(s, dtype=float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        kind = l % 3
        if kind == 0:
            return rng.uniform(0., 3., 8)
        if kind == 1:      # near the hand-tuned inner-loop gains (k1 = 10 a, k2 = a / 10)
            base = np.array([2.5, 2.5, 4.0, 1.0, 1.0, 2.0, 1.5, 2.0])
            return np.clip(base * rng.uniform(0.8, 1.2, 8), 0, 3)
        a = rng.uniform(0., 3., 8)
        a[rng.random(8) < 0.3] = 0.
        return a

    def step(self, env, a, d):
        a = np.array(a, dtype=float)
        env.get_param_from_actor(a)
        rhod, dot_rhod, _, _ = self.ref_cmd.ref_inner(env.time, env.ref_att_amplitude, env.ref_att_period,
                                                      env.ref_att_bias_a, env.ref_att_bias_phase)
        torque = env.att_control(rhod, dot_rhod, None)
        env.step_update([torque[0], torque[1], torque[2]])
        return (np.array(env.current_state, dtype=float), np.array(env.next_state, dtype=float),
                float(env.reward), bool(env.is_terminal), int(env.terminal_flag))


class UavAttRandA(UavAttA):
    name = "uav_att_rand"
    random_trajectory = True


REGISTRY.update({
    "uav_pos": (UavPosA, 6, 1000, 21),
    # wide fixtures: one full episode each for 24 / 32 independent seeds.  The 64 seeds x 1000 steps of SURVEY 8c-i are
    # not committed (28 MB per env): tests/test_live_reference_gpu.py records them from the staged reference at test time.
    "uav_pos_wide": (UavPosDisA, 24, 520, 41),
    "uav_pos_rp0": (UavPosRandomPos0A, 4, 600, 28),
    "uav_pos_dis": (UavPosDisA, 3, 1000, 22),
    "uav_att": (UavAttA, 6, 1000, 23),
    "uav_att_rand": (UavAttRandA, 3, 1000, 24),
})


class UavPosCrashA(UavPosDisA):
    """Weak gains + a strong constant disturbance: episodes end with the position-out flag (2)."""
    name = "uav_pos_crash"

    def sample_action(self, rng, t, l, env=None):
        return rng.uniform(0., 0.3, 8)

    def sample_dis(self, rng, t, l, env=None):
        return np.array([4.0, -3.0, 2.0]) * (1 + 0.5 * l) + rng.normal(0, 0.05, 3)


class UavPosEdgeA(UavPosA):
    """Reset followed by an injected near-edge attitude and a body-rate kick: attitude-out flag (3)."""
    name = "uav_pos_edge"

    def reset(self, env):
        super().reset(env)
        env.theta = 1.40 + 0.02 * (getattr(self, "_k", 0) % 3)
        env.q = 8.0
        env.phi = -0.3
        self._k = getattr(self, "_k", 0) + 1


class UavAttEdgeA(UavAttRandA):
    name = "uav_att_edge"

    def reset(self, env):
        super().reset(env)
        k = getattr(self, "_k", 0)
        if k % 2 == 0:
            env.phi, env.p = 1.45, 6.0
        else:
            env.psi, env.r, env.theta = 3.1235, (1.0 if k % 4 == 1 else 2.0), -0.4   # psi wraps past pi and leaves the zone; theta near its edge
        self._k = k + 1


REGISTRY.update({
    "uav_pos_crash": (UavPosCrashA, 2, 400, 25),
    "uav_pos_edge": (UavPosEdgeA, 2, 200, 26),
    "uav_att_edge": (UavAttEdgeA, 2, 200, 27),
})


# ------------------------------------------------------------- small envs
class FasA(Adapter):
    name = "fas"
    cites = "environment/FlightAttitudeSimulator/FlightAttitudeSimulator.py:173-287"
    F, S, A, D = 2, 2, 1, 0
    action_lo, action_hi = np.array([-1.5]), np.array([4.])
    perturb_attrs = ("theta", "dTheta")

    def make(self):
        return R.load("environment.FlightAttitudeSimulator.FlightAttitudeSimulator").Flight_Attitude_Simulator(0.)

    def internal(self, env):
        return np.array([env.theta, env.dTheta], dtype=float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        if l % 4 == 0:
            return rng.uniform(self.action_lo, self.action_hi)
        if l % 4 == 2:
            return rng.uniform(3.0, 4.0, 1)  # pushes up: upper-bound flag (1)
        hover = env.m * env.g * env.dis / env.L  # keeps the rod near level: episodes reach the time-out flag
        return np.clip(np.array([hover - 3.0 * env.theta - 0.8 * env.dTheta]) + rng.uniform(-0.3, 0.3, 1), -1.5, 4)


class FasPPO2A(FasA):
    name = "fas_ppo2"
    cites = "demonstration/PPO2/PPO2-4-FlightAttitudeSimulator/flight_attitude_simulator.py:170-280"

    def make(self):
        m = R.load_file("demonstration/PPO2/PPO2-4-FlightAttitudeSimulator/flight_attitude_simulator.py", "ref_fas_ppo2")
        return m.Flight_Attitude_Simulator(0.)


class FasDiscreteA(Adapter):
    name = "fas_discrete"
    cites = "environment/FlightAttitudeSimulator/FlightAttitudeSimulatorDiscrete.py:158-274"
    F, S, A, D = 2, 2, 1, 0
    action_lo, action_hi = np.array([-1.6]), np.array([3.0])
    perturb_attrs = ("theta", "dTheta")

    def make(self):
        m = R.load("environment.FlightAttitudeSimulator.FlightAttitudeSimulatorDiscrete")
        return m.FlightAttitudeSimulatorDiscrete(0.)

    def internal(self, env):
        return np.array([env.theta, env.dTheta], dtype=float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        space = env.action_space[0]
        if l % 3 == 0:  # uniformly random discrete action
            return np.array([space[int(rng.integers(len(space)))]])
        if l % 3 == 1:  # pushes up: bounce at +theta_max, stays inside theta_out -> time-out flag
            return np.array([space[int(rng.integers(len(space) - 8, len(space)))]])
        # nearest discrete force of a PD law around level: long episodes, small angles
        want = 0.81 - 3.0 * env.theta - 0.8 * env.dTheta
        return np.array([space[int(np.argmin(np.abs(np.array(space) - want)))]])


class SoiA(Adapter):
    name = "soi"
    cites = "environment/SecondOrderIntegration/SecondOrderIntegration.py:211-352"
    F, S, A, D = 4, 4, 2, 0
    action_lo, action_hi = np.array([-3., -3.]), np.array([3., 3.])
    perturb_attrs = ()

    def make(self):
        return R.load("environment.SecondOrderIntegration.SecondOrderIntegration").SecondOrderIntegration()

    def internal(self, env):
        return np.array([env.pos[0], env.pos[1], env.vel[0], env.vel[1]], dtype=float), float(env.time)

    def perturb(self, env, k):
        env.vel[0] += (self.EPS if k % 2 == 0 else -self.EPS) * max(1.0, abs(env.vel[0]))

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:
            return rng.uniform(self.action_lo, self.action_hi)
        e = env.target - env.pos  # PD towards the target: reaches the success flag
        return np.clip(4.0 * e - 2.5 * env.vel + rng.uniform(-0.05, 0.05, 2), -3, 3)


class SoiDPPO2A(SoiA):
    name = "soi_dppo2"
    cites = "demonstration/DPPO2/DPPO2-4-SecondOrderIntegration/SecondOrderIntegration.py:208-340"

    def make(self):
        m = R.load_file("demonstration/DPPO2/DPPO2-4-SecondOrderIntegration/SecondOrderIntegration.py", "ref_soi_dppo2")
        return m.SecondOrderIntegration()


class BallBalancerA(Adapter):
    name = "ballbalancer"
    cites = "environment/BallBalancer/BallBalancer1D.py:200-322"
    F, S, A, D = 4, 3, 1, 0
    action_lo, action_hi = np.array([-np.pi]), np.array([np.pi])
    perturb_attrs = ("vel", "theta")

    def make(self):
        return R.load("environment.BallBalancer.BallBalancer1D").BallBalancer1D()

    def internal(self, env):
        return np.array([env.pos, env.vel, env.theta, env.error], dtype=float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:
            return rng.uniform(-4.0, 4.0, 1)  # beyond +-pi: exercises the np.clip of the action
        goal = env.target + (0.04 if l % 4 == 3 else 0.0)  # lane 3 holds the ball off-target: time-out flag (2)
        u = -(25.0 * (env.pos - goal) + 12.0 * env.vel) - 6.0 * env.theta  # stabilising: success flag (3)
        return np.clip(np.array([u]) + rng.uniform(-0.02, 0.02, 1), -np.pi, np.pi)


class TwoLinkA(Adapter):
    name = "twolink"
    cites = "environment/RobotManipulator/TwoLinkManipulator.py:186-312"
    F, S, A, D = 8, 6, 2, 0
    action_lo, action_hi = np.array([-5., -5.]), np.array([5., 5.])

    def make(self):
        return R.load("environment.RobotManipulator.TwoLinkManipulator").TwoLinkManipulator()

    def perturb(self, env, k):
        env.omega[0] += (self.EPS if k % 2 == 0 else -self.EPS) * max(1.0, abs(env.omega[0]))

    def internal(self, env):
        return np.array([env.theta[0], env.theta[1], env.omega[0], env.omega[1], env.error[0], env.error[1],
                         env.target[0], env.target[1]], dtype=float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:
            return rng.uniform(self.action_lo, self.action_hi)
        return np.clip(-3.0 * env.omega + rng.uniform(-0.5, 0.5, 2), -5, 5)  # damped: long episodes (time-out flag)


class UgvForwardA(Adapter):
    name = "ugv_forward"
    cites = "environment/UGV/UGVForward.py:217-362"
    F, S, A, D = 5, 4, 2, 0
    action_lo, action_hi = np.array([-3., -2 * np.pi]), np.array([3., 2 * np.pi])
    perturb_attrs = ("vel", "phi", "omega")
    mod, cls = "environment.UGV.UGVForward", "UGVForward"

    def make(self):
        return getattr(R.load(self.mod), self.cls)()

    def internal(self, env):
        return np.array([env.pos[0], env.pos[1], env.vel, env.phi, env.omega], dtype=float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:
            return rng.uniform(self.action_lo, self.action_hi)
        # go-to-goal: heads for the target and brakes there (success flag), some noise
        e, ephi = env.get_e(), env.get_e_phi()
        a_lin = 2.0 * abs(e) * np.cos(ephi) - 2.5 * env.vel
        a_ang = 6.0 * ephi - 4.0 * env.omega
        return np.clip(np.array([a_lin, a_ang]) + rng.uniform(-0.02, 0.02, 2), self.action_lo, self.action_hi)


class UgvBidirectionalA(UgvForwardA):
    name = "ugv_bidirectional"
    cites = "environment/UGV/UGVBidirectional.py:217-367"
    mod, cls = "environment.UGV.UGVBidirectional", "UGVBidirectional"

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:
            return rng.uniform(self.action_lo, self.action_hi)
        e, ephi = env.get_e(), env.get_e_phi()
        a_lin = 2.0 * e - 2.5 * env.vel
        a_ang = 6.0 * ephi - 4.0 * env.omega
        return np.clip(np.array([a_lin, a_ang]) + rng.uniform(-0.02, 0.02, 2), self.action_lo, self.action_hi)


REGISTRY.update({
    "cartpole_wide": (CartPoleA, 32, 300, 42),
    "fas": (FasA, 4, 800, 31),
    "fas_ppo2": (FasPPO2A, 2, 1200, 32),
    "fas_discrete": (FasDiscreteA, 3, 700, 39),
    "soi": (SoiA, 4, 800, 33),
    "soi_dppo2": (SoiDPPO2A, 2, 600, 34),
    "ballbalancer": (BallBalancerA, 4, 1000, 35),
    "twolink": (TwoLinkA, 4, 900, 36),
    "ugv_forward": (UgvForwardA, 4, 1100, 37),
    "ugv_bidirectional": (UgvBidirectionalA, 4, 1100, 38),
})


# ------------------------------------------------------------- UGVForwardObstacleAvoidance
class UgvoA(Adapter):
    name = "ugvo"
    cites = "environment/UGVForwardObstacleAvoidance/UGVForwardObstacleAvoidance.py:261-557, map.py:65-174"
    F, S, A, D = 56, 41, 2, 0
    action_lo, action_hi = np.array([-3., -2 * np.pi]), np.array([3., 2 * np.pi])
    perturb_attrs = ("vel", "phi", "omega")

    def make(self):
        R.use_family("UGVForwardObstacleAvoidance")
        return R.load("UGVForwardObstacleAvoidance").UGVForwardObstacleAvoidance()

    def internal(self, env):
        s = [env.pos[0], env.pos[1], env.vel, env.phi, env.omega, env.target[0], env.target[1], float(len(env.obs))]
        for k in range(16):
            if k < len(env.obs):
                s += [env.obs[k][1][0], env.obs[k][1][1], env.obs[k][2][0]]
            else:
                s += [0., 0., 0.]
        return np.array(s, dtype=float), float(env.time)

    def reset(self, env):
        env.reset(True)

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:
            return np.array([rng.uniform(0.0, 3.0), rng.uniform(-2.0, 2.0)])  # drives around: collisions, out of map
        e, ephi = env.get_e(), env.get_e_phi()  # go-to-goal (may hit obstacles on the way)
        a_lin = 2.0 * abs(e) * np.cos(ephi) - 2.5 * env.vel
        a_ang = 6.0 * ephi - 4.0 * env.omega
        return np.clip(np.array([a_lin, a_ang]) + rng.uniform(-0.02, 0.02, 2), self.action_lo, self.action_hi)


class UgvoDPPO2A(UgvoA):
    name = "ugvo_dppo2"
    cites = "demonstration/DPPO2/DPPO2-4-UGVForwardObstacleAvoidance/UGVForwardObstacleAvoidance.py:259-555"

    def make(self):
        R.use_family("UGVForwardObstacleAvoidance")
        m = R.load_file("demonstration/DPPO2/DPPO2-4-UGVForwardObstacleAvoidance/UGVForwardObstacleAvoidance.py", "ref_ugvo_dppo2")
        return m.UGVForwardObstacleAvoidance()

    def sample_action(self, rng, t, l, env=None):
        if l % 2 == 0:  # includes braking below zero: exercises the pre-update `vel < 0` freeze (note N9)
            return np.array([rng.uniform(-3.0, 3.0), rng.uniform(-2.0, 2.0)])
        return super().sample_action(rng, t, 1, env)


class UgvoEdgeA(UgvoA):
    """Collision-radius equality (collision_check :261-272: `dis_two_points(pos, centre) <= r + r_vehicle`): after every
    reset the vehicle is parked (vel = omega = 0, zero actions) at a point whose distance to the first obstacle's centre --
    evaluated exactly like the reference does, np.linalg.norm of the difference -- is EQUAL to r + r_vehicle (lanes 0, 3),
    the closest representable distance above it (lanes 1, 4: no collision) or below it (lanes 2, 5).  The point is found
    by a search over +-12 ulp offsets of both coordinates around the nominal point."""
    name = "ugvo_edge"

    def make(self):
        env = super().make()
        n = getattr(self, "_made", 0)          # gen_golden makes (env, twin) per lane, in lane order
        self._made = n + 1
        env._edge_lane = n // 2
        return env

    def reset(self, env):
        env.reset(True)
        mode = env._edge_lane % 3
        c = np.array(env.obs[0][1], dtype=float)
        R = env.obs[0][2][0] + env.r_vehicle
        centre = np.array([env.x_size / 2.0, env.y_size / 2.0]) if hasattr(env, "x_size") else np.array([2.5, 2.5])
        u = centre - c
        u = u / max(np.linalg.norm(u), 1e-9) if np.linalg.norm(u) > 1e-9 else np.array([1.0, 0.0])
        nominal = c + R * u
        best = None
        for ix in range(-12, 13):
            px = nominal[0]
            for _ in range(abs(ix)):
                px = np.nextafter(px, np.inf if ix > 0 else -np.inf)
            for iy in range(-12, 13):
                py = nominal[1]
                for _ in range(abs(iy)):
                    py = np.nextafter(py, np.inf if iy > 0 else -np.inf)
                d = float(np.linalg.norm(np.array([px, py]) - c))
                if mode == 0:
                    key = (abs(d - R), 0)
                elif mode == 1:
                    key = (d - R if d > R else np.inf, 0)
                else:
                    key = (R - d if d < R else np.inf, 0)
                if best is None or key < best[0]:
                    best = (key, px, py, d)
        env.pos = np.array([best[1], best[2]], dtype=float)
        env.vel, env.omega = 0.0, 0.0
        self.last_edge = (mode, best[3], R)

    def sample_action(self, rng, t, l, env=None):
        return np.zeros(2)


REGISTRY.update({
    "ugvo": (UgvoA, 4, 300, 41),
    "ugvo_dppo2": (UgvoDPPO2A, 4, 400, 42),
    "ugvo_edge": (UgvoEdgeA, 6, 30, 43),
})


# ------------------------------------------------------------- UavRobust (own interpreter state: clashing module names)
def _robust_params():
    R.use_family("UavRobust")
    uav = R.load("environment.UavRobust.uav")
    fn = R.load("environment.UavRobust.FNTSMC")
    up = uav.uav_param()
    up.dt, up.time_max = 0.01, 10
    up.pos_zone = np.atleast_2d([[-5, 5], [-5, 5], [0, 5]])
    att = fn.fntsmc_param()
    att.k1 = np.array([25., 25., 40.]); att.k2 = np.array([0.1, 0.1, 0.2]); att.alpha = np.array([2.5, 2.5, 2.5])
    att.beta = np.array([0.99, 0.99, 0.99]); att.gamma = np.array([1.5, 1.5, 1.2]); att.lmd = np.array([2.0, 2.0, 2.0])
    att.dim, att.dt, att.ctrl0 = 3, 0.01, np.array([0., 0., 0.])
    att.saturation = np.array([0.3, 0.3, 0.3])
    return up, att, fn.fntsmc_param()


class _RobustBase(Adapter):
    F, D = 33, 3
    perturb_attrs = ("p", "q", "r")

    def reset(self, env):
        env.reset(random=True)

    def _common(self, env):
        s1 = env.att_ctrl.s1 if hasattr(env, "att_ctrl") else np.zeros(3)
        aref = getattr(env, "att_ref", np.zeros(3))
        daref = getattr(env, "dot_att_ref", np.zeros(3))
        return env.uav_state_call_back(), s1, aref, daref

    def step(self, env, a, d):
        if self.D and d is not None and hasattr(env, "dis"):
            env.dis = np.array(d, dtype=float)
        env.step_update(np.array(a, dtype=float))
        return (np.array(env.current_state, dtype=float), np.array(env.next_state, dtype=float),
                float(env.reward), bool(env.is_terminal), int(env.terminal_flag))

    def sample_dis(self, rng, t, l, env=None):
        if l % 2 == 0:
            return np.zeros(3)
        tt = 0.01 * t  # generate_uncertainty(time, is_ideal=False), UavRobust/ref_cmd.py:46-61
        w = 2 * np.pi / 4
        return np.array([0.8 * np.cos(w * tt) + 0.6 * np.sin(w * tt), 0.8 * np.sin(w * tt) + 0.6 * np.cos(w * tt),
                         0.8 * np.cos(w * tt) + 0.6 * np.sin(w * tt)])


class HoverOuterA(_RobustBase):
    name = "uavr_hover_outer"
    cites = "environment/UavRobust/UavHoverOuterLoop.py:81-196 (+ uav.py:429-560, FNTSMC.py:80-106, uav_pos_ctrl.py:41-76)"
    S, A = 6, 3

    def make(self):
        up, att, pos = _robust_params()
        m = R.load("environment.UavRobust.UavHoverOuterLoop")
        return m.uav_hover_outer_loop(up, pos, att, np.array([1.0, -1.0, 2.0]))

    def internal(self, env):
        x, s1, aref, daref = self._common(env)
        return np.concatenate((x, s1, aref, daref, env.pos_ref, np.zeros(9))).astype(float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        if l % 3 == 0:
            return rng.uniform(-8, 8, 3)  # wild accelerations: position / attitude out
        e = env.uav_pos() - env.pos_ref    # PD hover controller + noise: reaches the time-out flag
        return np.clip(-2.0 * e - 2.5 * env.uav_vel() + rng.uniform(-0.3, 0.3, 3), -8, 8)


class HoverA(HoverOuterA):
    name = "uavr_hover"
    cites = "environment/UavRobust/UavHover.py:101-226"
    S, A = 12, 6

    def make(self):
        up, att, pos = _robust_params()
        m = R.load("environment.UavRobust.UavHover")
        return m.uav_hover(up, pos, att, np.array([1.0, -1.0, 2.0]))

    def sample_action(self, rng, t, l, env=None):
        acc = super().sample_action(rng, t, l, env)
        if l % 3 == 0:
            tq = rng.uniform(-0.3, 0.3, 3)
        else:  # PD attitude torque towards the shaped reference
            e_att = env.uav_att() - env.att_ref
            tq = np.clip(-0.15 * e_att - 0.03 * env.uav_pqr() + rng.uniform(-0.005, 0.005, 3), -0.3, 0.3)
        return np.concatenate((acc, tq))


class InnerLoopA(_RobustBase):
    name = "uavr_inner"
    cites = "environment/UavRobust/UavInnerLoop.py:88-210 (+ uav_att_ctrl.py:33-46)"
    S, A, D = 6, 3, 0

    def make(self):
        up, att, _ = _robust_params()
        m = R.load("environment.UavRobust.UavInnerLoop")
        return m.uav_inner_loop(up, att, np.array([0.3, 0.3, 0.3]), np.array([5., 5., 5.]), np.zeros(3), np.array([np.pi / 2, 0., 0.]))

    def internal(self, env):
        x = env.uav_state_call_back()
        return np.concatenate((x, np.zeros(12), env.ref_amplitude, env.ref_period, env.ref_bias_phase)).astype(float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        if l % 3 == 0:
            return rng.uniform(-0.3, 0.3, 3)
        e = env.uav_att() - env.ref
        return np.clip(-0.2 * e - 0.03 * (env.dot_rho1() - env.dot_ref) + rng.uniform(-0.005, 0.005, 3), -0.3, 0.3)


class TrackingOuterA(_RobustBase):
    name = "uavr_tracking"
    cites = "environment/UavRobust/UavTrackingOuterLoop.py:90-255"
    S, A = 6, 3

    def make(self):
        up, att, pos = _robust_params()
        m = R.load("environment.UavRobust.UavTrackingOuterLoop")
        return m.uav_tracking_outer_loop(up, pos, att, np.array([1.5, 1.5, 0.3]), np.array([6., 6., 10.]),
                                         np.array([0., 0., 2.5]), np.array([np.pi / 2, 0., 0.]))

    def internal(self, env):
        x, s1, aref, daref = self._common(env)
        return np.concatenate((x, s1, aref, daref, np.zeros(3), env.ref_amplitude, env.ref_period, env.ref_bias_phase)).astype(float), float(env.time)

    def sample_action(self, rng, t, l, env=None):
        if l % 3 == 0:
            return rng.uniform(-8, 8, 3)
        e = env.uav_pos() - env.pos_ref
        de = env.uav_vel() - env.dot_pos_ref
        return np.clip(-2.0 * e - 2.5 * de + rng.uniform(-0.3, 0.3, 3), -8, 8)


REGISTRY.update({
    "uavr_hover_outer": (HoverOuterA, 3, 1500, 51),
    "uavr_hover": (HoverA, 3, 1500, 52),
    "uavr_inner": (InnerLoopA, 3, 1500, 53),
    "uavr_tracking": (TrackingOuterA, 3, 1500, 54),
})
