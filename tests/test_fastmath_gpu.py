"""GPU: accuracy of the in-house fp64 sincos/exp/log/tanh/pow (csrc/fastmath64.cuh) against mpmath-grade references
(numpy float64 libm results, themselves < 1 ulp) on the argument ranges the env kernels use, plus IEEE edge cases."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eval(func, x, aux=None):
    import torch
    from reinforcementlearningplatform_b200 import _lib
    lib = _lib.load()
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    o0 = torch.empty_like(xd)
    o1 = torch.empty_like(xd) if aux is None else torch.from_numpy(np.ascontiguousarray(aux, dtype=np.float64)).cuda()
    _lib.check(lib.b200_fastmath_eval(func, xd.numel(), ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(o0.data_ptr()),
                                      ctypes.c_void_p(o1.data_ptr()), None), "b200_fastmath_eval")
    torch.cuda.synchronize()
    return o0.cpu().numpy(), o1.cpu().numpy()


def _ulps(got, ref):
    ref = np.asarray(ref, np.float64)
    return np.max(np.abs(got - ref) / np.spacing(np.abs(ref)))


def test_sincos_accuracy():
    rng = np.random.default_rng(1)
    x = np.concatenate((rng.uniform(-np.pi, np.pi, 200000), rng.uniform(-40, 40, 100000), rng.uniform(-1e3, 1e3, 50000),
                        [0.0, -0.0, np.pi / 2, np.pi, 1e-300, 1e-9]))
    s, c = _eval(0, x)
    # near the zeros of sin/cos the reference value is tiny: compare absolutely there (1 ulp of 1.0)
    es = np.abs(s - np.sin(x)) / np.maximum(np.spacing(np.abs(np.sin(x))), 2 ** -53 * 1e-3)
    ec = np.abs(c - np.cos(x)) / np.maximum(np.spacing(np.abs(np.cos(x))), 2 ** -53 * 1e-3)
    assert es.max() <= 2.5 and ec.max() <= 2.5, (es.max(), ec.max())


def test_sincos_special_values():
    s, c = _eval(0, np.array([np.inf, -np.inf, np.nan, 1e300]))
    assert np.isnan(s[:3]).all() and np.isnan(c[:3]).all()
    assert abs(s[3] - np.sin(1e300)) < 1e-15 and abs(c[3] - np.cos(1e300)) < 1e-15  # falls back to libdevice


def test_exp_accuracy_and_limits():
    rng = np.random.default_rng(2)
    x = np.concatenate((rng.uniform(-700, 700, 300000), rng.uniform(-40, 40, 200000), [0.0, -0.0, 1.0, -1.0]))
    e, _ = _eval(1, x)
    assert _ulps(e, np.exp(x)) <= 2.0
    e, _ = _eval(1, np.array([-np.inf, -1e4, -745.2, np.inf, 800.0, np.nan]))
    assert e[0] == 0 and e[1] == 0 and e[2] == 0 and np.isinf(e[3]) and np.isinf(e[4]) and np.isnan(e[5])


def test_log_accuracy_and_limits():
    rng = np.random.default_rng(3)
    x = np.concatenate((10.0 ** rng.uniform(-300, 300, 300000), rng.uniform(0.5, 2.0, 200000), rng.uniform(0, 1e-3, 1000),
                        [1.0, 2.2250738585072014e-308]))
    l, _ = _eval(2, x)
    ref = np.log(x)
    err = np.abs(l - ref) / np.maximum(np.spacing(np.abs(ref)), 2 ** -53 * 1e-3)
    assert err.max() <= 2.0, err.max()
    l, _ = _eval(2, np.array([0.0, -0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310]))
    assert l[0] == -np.inf and l[1] == -np.inf and np.isnan(l[2]) and l[3] == np.inf and np.isnan(l[4])
    assert l[5] == -np.inf and l[6] == -np.inf  # subnormal arguments are treated as 0 (documented in fastmath64.cuh)


def test_tanh_absolute_accuracy():
    rng = np.random.default_rng(4)
    x = np.concatenate((rng.uniform(-25, 25, 300000), rng.normal(0, 1e-3, 100000), [0.0, -0.0, 400.0, -400.0, 1e-200]))
    t, _ = _eval(3, x)
    assert np.max(np.abs(t - np.tanh(x))) <= 4e-16  # absolute, see fastmath64.cuh
    t, _ = _eval(3, np.array([np.inf, -np.inf, np.nan]))
    assert t[0] == 1 and t[1] == -1 and np.isnan(t[2])


def test_pow_from_log_matches_pow():
    rng = np.random.default_rng(5)
    x = np.concatenate((10.0 ** rng.uniform(-12, 2, 300000), [0.0, 0.0, 0.0, 1.0]))
    a = np.concatenate((rng.choice([0.2, 0.3, 0.5, 0.99, 1.2, 1.5, 2.5, 1.5 - 1, 2.5 - 1], 300000), [1.2, 0.0, 0.2, 0.0]))
    p, _ = _eval(4, x, aux=a)
    ref = np.power(x, a)
    rel = np.abs(p - ref) / np.maximum(np.abs(ref), 1e-300)
    assert rel.max() <= 2e-14, rel.max()   # (|a log x| + 1) ulp, see common.cuh:pow_from_log
    assert p[-4] == 0 and p[-3] == 1 and p[-2] == 0 and p[-1] == 1


def test_asin_accuracy_and_limits():
    rng = np.random.default_rng(6)
    x = np.concatenate((rng.uniform(-1, 1, 400000), rng.uniform(0.49, 0.51, 50000), 1 - 10.0 ** rng.uniform(-16, -1, 50000),
                        [0.0, -0.0, 0.5, -0.5, 1.0, -1.0]))
    a, _ = _eval(5, x)
    ref = np.arcsin(x)
    err = np.abs(a - ref) / np.maximum(np.spacing(np.abs(ref)), 2 ** -53 * 1e-3)
    assert err.max() <= 2.5, err.max()   # branch-free kernel + Goldschmidt sqrt on the folded half
    assert a[-2] == np.pi / 2 and a[-1] == -np.pi / 2 and a[-6] == 0
    a, _ = _eval(5, np.array([1.0000001, -2.0, np.nan]))
    assert np.all(np.isnan(a))
