"""CPU: host-side pieces of K-LEARN -- the keyed mini-batch permutation (evaluated on the host by the same function the
kernel inlines) and the flat parameter buffer that the nets' nn.Module parameters are re-pointed into."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("B", [1, 2, 3, 64, 1000, 4096, 100003, (1 << 20) + 7])
def test_keyed_permutation_is_a_bijection(B):
    from reinforcementlearningplatform_b200 import _lib
    lib = _lib.load()
    out = np.empty(B, dtype=np.int64)
    for key in (0, 1, 0x9E3779B97F4A7C15):
        _lib.check(lib.b200_ppo2_permutation(key, B, 0, B, out.ctypes.data_as(C.POINTER(C.c_int64))), "perm")
        assert out.min() == 0 and out.max() == B - 1 and np.unique(out).size == B
    if B >= 1000:
        a, b = np.empty(B, dtype=np.int64), np.empty(B, dtype=np.int64)
        lib.b200_ppo2_permutation(5, B, 0, B, a.ctypes.data_as(C.POINTER(C.c_int64)))
        lib.b200_ppo2_permutation(6, B, 0, B, b.ctypes.data_as(C.POINTER(C.c_int64)))
        assert np.mean(a == b) < 0.01                      # another key = another order
        assert abs(np.corrcoef(np.arange(B), a)[0, 1]) < 0.1  # no trace of the identity
        # a slice of the permutation equals the same positions of the whole
        part = np.empty(100, dtype=np.int64)
        lib.b200_ppo2_permutation(5, B, 300, 100, part.ctypes.data_as(C.POINTER(C.c_int64)))
        assert np.array_equal(part, a[300:400])
    assert lib.b200_ppo2_permutation(0, B, 0, B + 1, out.ctypes.data_as(C.POINTER(C.c_int64))) == -6   # B200ENV_ESIZE


def test_flat_params_keep_modules_working_and_follow_torch_parameter_order():
    import torch
    from reinforcementlearningplatform_b200.learn import FlatParams, fused_supported
    from reinforcementlearningplatform_b200.ppo2 import dppo2_nets, reference_nets
    torch.manual_seed(0)
    actor, critic = reference_nets(6, 8, "cpu")
    x = torch.randn(5, 6)
    ya, yc = actor(x).clone(), critic(x).clone()
    ref = torch.cat([p.detach().reshape(-1) for net in (actor, critic) for p in net.parameters()])
    fp = FlatParams([actor, critic])
    assert torch.equal(fp.flat, ref)                      # layer order, weight before bias = parameters() order
    assert fp.net_off == [0, sum(p.numel() for p in actor.parameters())]
    assert torch.equal(actor(x), ya) and torch.equal(critic(x), yc)
    with torch.no_grad():
        fp.flat.mul_(0.5)                                 # an in-place update of the flat buffer IS an update of the nets
    assert not torch.equal(actor(x), ya)
    assert all(p.data_ptr() >= fp.flat.data_ptr() and p.data_ptr() < fp.flat.data_ptr() + 4 * fp.flat.numel()
               for net in (actor, critic) for p in net.parameters())
    assert fused_supported(actor, critic)
    wide = dppo2_nets(41, 2, np.array([-3.0, -1.0]), np.array([3.0, 2.0]), "cpu")
    assert not fused_supported(*wide)
