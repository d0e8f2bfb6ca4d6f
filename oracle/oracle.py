"""ctypes driver of oracle/liboracle.so -- the CPU checker.

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class OracleIO(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "state", "time", "episode", "action", "dis", "obs", "next_obs", "reward", "done", "flag", "reset_obs",
        "substeps")]


_lib = None


def build() -> None:
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        vp, i32, i64, u32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64
        lib.oracle_step.restype = i32
        lib.oracle_step.argtypes = [i32, i64, vp, C.POINTER(OracleIO), u32, u64, i64, i32]
        lib.oracle_reset.restype = i32
        lib.oracle_reset.argtypes = [i32, i64, vp, C.POINTER(OracleIO), vp, u64, i64]
        lib.oracle_observe.restype = i32
        lib.oracle_observe.argtypes = [i32, i64, vp, C.POINTER(OracleIO)]
        lib.oracle_gae.restype = None
        lib.oracle_gae.argtypes = [i64, i64, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp]
        lib.oracle_norm_seq.restype = None
        lib.oracle_norm_seq.argtypes = [i64, i32, vp, vp, vp, i32, C.c_double]
        lib.oracle_mc_returns.restype = None
        lib.oracle_mc_returns.argtypes = [i64, i64, vp, vp, C.c_double, vp]
        lib.oracle_philox4x32_10.restype = None
        lib.oracle_philox4x32_10.argtypes = [vp, vp, vp]
        _lib = lib
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    load().oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


class OracleEnv:
    """n instances of one env family stepped by the C restatement; numpy SoA buffers ([field, n])."""

    def __init__(self, env_id: int, params: C.Structure, n: int, F: int, S: int, A: int, D: int = 0,
                 seed: int = 0, env_index_offset: int = 0, auto_reset: bool = False, nthreads: int = 1):
        self.lib = load()
        self.env_id, self.params, self.n = env_id, params, int(n)
        self.F, self.S, self.A, self.D = F, S, A, D
        self.seed, self.off, self.auto_reset, self.nthreads = seed, env_index_offset, auto_reset, nthreads
        z = np.zeros
        self.state = z((F, n))
        self.time = z(n)
        self.episode = z(n, np.uint32)
        self.obs = z((S, n))
        self.next_obs = z((S, n))
        self.reset_obs = z((S, n))
        self.reward = z(n)
        self.done = z(n, np.uint8)
        self.flag = z(n, np.int32)
        self.substeps = z(n, np.int32)

    def _io(self, action=None, dis=None) -> OracleIO:
        io = OracleIO()
        p = lambda a: None if a is None else a.ctypes.data
        io.state, io.time, io.episode = p(self.state), p(self.time), p(self.episode)
        io.action, io.dis = p(action), p(dis)
        io.obs, io.next_obs, io.reset_obs = p(self.obs), p(self.next_obs), p(self.reset_obs)
        io.reward, io.done, io.flag, io.substeps = p(self.reward), p(self.done), p(self.flag), p(self.substeps)
        return io

    def step(self, action_soa: np.ndarray, dis_soa: np.ndarray | None = None) -> None:
        a = np.ascontiguousarray(action_soa, dtype=np.float64).reshape(self.A, self.n)
        d = None if dis_soa is None else np.ascontiguousarray(dis_soa, dtype=np.float64).reshape(self.D, self.n)
        self._keep = (a, d)
        io = self._io(a, d)
        rc = self.lib.oracle_step(self.env_id, self.n, C.byref(self.params), C.byref(io),
                                  1 if self.auto_reset else 0, self.seed, self.off, self.nthreads)
        assert rc == 0, rc

    def reset(self, mask: np.ndarray | None = None) -> None:
        io = self._io()
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = self.lib.oracle_reset(self.env_id, self.n, C.byref(self.params), C.byref(io),
                                   None if m is None else m.ctypes.data, self.seed, self.off)
        assert rc == 0, rc

    def observe(self) -> np.ndarray:
        io = self._io()
        rc = self.lib.oracle_observe(self.env_id, self.n, C.byref(self.params), C.byref(io))
        assert rc == 0, rc
        return self.next_obs


def gae(r, vs, vs_next, done, success, gamma, lmd):
    """C restatement of the reference GAE loop on time-major [T, N] float32 arrays -> (adv, v_target, stats)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (r, vs, vs_next, done, success)]
    T, N = arrs[0].shape
    adv = np.zeros((T, N), np.float32)
    vt = np.zeros((T, N), np.float32)
    stats = np.zeros(3, np.float64)
    load().oracle_gae(T, N, *[a.ctypes.data for a in arrs], np.float32(gamma), np.float32(gamma * lmd),
                      adv.ctypes.data, vt.ctypes.data, stats.ctypes.data)
    return adv, vt, stats


def norm_seq(x_soa, run=None, update=True, eps=1e-8):
    """C restatement of Normalization.__call__ fed sample by sample: x_soa [dim, rows] float64 ->
    (y [dim, rows], run [4, dim] = n, mean, S, std)."""
    x = np.ascontiguousarray(x_soa, dtype=np.float64)
    dim, rows = x.shape
    run = np.zeros((4, dim)) if run is None else np.ascontiguousarray(run, dtype=np.float64).copy()
    y = np.zeros_like(x)
    load().oracle_norm_seq(rows, dim, x.ctypes.data, y.ctypes.data, run.ctypes.data, 1 if update else 0, eps)
    return y, run


def mc_returns(r, done, gamma):
    """C restatement of the PPO (v1) return scan on time-major [T, N] arrays -> float32 [T, N]."""
    r = np.ascontiguousarray(r, dtype=np.float64)
    d = np.ascontiguousarray(done, dtype=np.uint8)
    T, N = r.shape
    out = np.zeros((T, N), np.float32)
    load().oracle_mc_returns(T, N, r.ctypes.data, d.ctypes.data, float(gamma), out.ctypes.data)
    return out
