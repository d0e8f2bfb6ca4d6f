"""CPU: the C-ABI library loads, exports every symbol include/b200env.h declares, and the ctypes mirrors of the
parameter structs have the size the library was compiled with (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from reinforcementlearningplatform_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.load()


def declared_symbols():
    out = []
    for h in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", h)).read()
        out += re.findall(r"B200_API\s+[\w\s\*]+?\b(b200\w+)\s*\(", src)
    return sorted(set(out))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 7
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported by libb200env.so"


def test_param_struct_sizes(lib):
    from reinforcementlearningplatform_b200 import _lib
    for env_id, cls in _lib.PARAMS_OF.items():
        assert lib.b200env_params_bytes(env_id) == ctypes.sizeof(cls), (env_id, cls)


def test_bad_arguments_are_rejected(lib):
    from reinforcementlearningplatform_b200 import _lib
    io = _lib.IO()
    p = _lib.CartPoleParams()
    assert lib.b200env_step(99, 0, 4, ctypes.byref(p), ctypes.sizeof(p), ctypes.byref(io), 0, 0, 0, None) == -1
    assert lib.b200env_step(0, 7, 4, ctypes.byref(p), ctypes.sizeof(p), ctypes.byref(io), 0, 0, 0, None) == -2
    assert lib.b200env_step(0, 0, 0, ctypes.byref(p), ctypes.sizeof(p), ctypes.byref(io), 0, 0, 0, None) == -6
    assert lib.b200env_step(0, 0, 4, ctypes.byref(p), 8, ctypes.byref(io), 0, 0, 0, None) == -3
    assert lib.b200env_step(0, 0, 4, ctypes.byref(p), ctypes.sizeof(p), ctypes.byref(io), 0, 0, 0, None) == -4


def test_engine_refuses_cpu_device():
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200 import _lib
    with pytest.raises(_lib.B200EnvError):
        rlp.CartPole(n_envs=4, device="cpu")


def test_host_mirrors_carry_the_rl_base_description(lib):
    """Without a GPU: the host-only mirror of every env class has the dimensions the library reports (b200env_dims) and the
    descriptive lists of algorithm/rl_base.py:5-124 with one entry per dimension -- continuous everywhere except the
    discrete action set of FlightAttitudeSimulatorDiscrete (FlightAttitudeSimulatorDiscrete.py:58-59)."""
    import math
    import numpy as np
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200 import _lib
    makers = [lambda: rlp.CartPole(n_envs=1, host_only=True),
              lambda: rlp.CartPoleAngleOnly(n_envs=1, variant='env', host_only=True),
              lambda: rlp.CartPoleAngleOnly(n_envs=1, variant='ppo2', host_only=True),
              lambda: rlp.Flight_Attitude_Simulator(n_envs=1, host_only=True),
              lambda: rlp.FlightAttitudeSimulatorDiscrete(n_envs=1, host_only=True),
              lambda: rlp.SecondOrderIntegration(n_envs=1, host_only=True),
              lambda: rlp.BallBalancer1D(n_envs=1, host_only=True),
              lambda: rlp.TwoLinkManipulator(n_envs=1, host_only=True),
              lambda: rlp.UGVForward(n_envs=1, host_only=True),
              lambda: rlp.UGVBidirectional(n_envs=1, host_only=True),
              lambda: rlp.UGVForwardObstacleAvoidance(n_envs=1, variant='dppo2', host_only=True),
              lambda: rlp.UavAttCtrlRL(n_envs=1, host_only=True),
              lambda: rlp.UavPosCtrlRL(n_envs=1, host_only=True),
              lambda: rlp.uav_hover(n_envs=1, host_only=True)]
    for mk in makers:
        e = mk()
        _, od, ad, _ = _lib.dims(e.ENV_ID, e.VARIANT)
        assert (e.state_dim, e.action_dim) == (od, ad), type(e).__name__
        assert np.asarray(e.action_range, dtype=float).shape == (ad, 2), type(e).__name__
        assert len(e.state_num) == len(e.state_range) == len(e.isStateContinuous) == od
        assert len(e.action_num) == len(e.action_space) == ad
        if isinstance(e, rlp.FlightAttitudeSimulatorDiscrete):
            assert e.action_num[0] == len(e.action_space[0]) and e.action_num[0] > 1
        else:
            assert e.action_num == [math.inf] * ad and e.isActionContinuous == [True] * ad
        assert isinstance(e.name, str) and hasattr(e, "use_norm")
        with pytest.raises(AttributeError):
            e.no_such_attribute
