"""GAE: the C restatement against the vectors captured from the reference learner (CPU), and the CUDA kernel against
both (GPU).  Bit-exact in float32 (acc_mode 0); normalised advantages within 1e-5 (different but equivalent
mean/std reduction order)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN


def cases():
    with np.load(os.path.join(GOLDEN, "gae.npz")) as z:
        g = {k: z[k] for k in z.files}
    out = []
    for k in range(int(g["n_cases"])):
        out.append({n: g[f"c{k}_{n}"] for n in ("r", "vs", "vs_", "done", "success", "adv", "v_target", "adv_norm")})
    return out, float(g["gamma"]), float(g["lmd"])


def test_oracle_gae_bit_exact_vs_reference(oracle_lib):
    from oracle import oracle
    cs, gamma, lmd = cases()
    for c in cs:
        col = lambda a: a.reshape(-1, 1)
        adv, vt, stats = oracle.gae(col(c["r"]), col(c["vs"]), col(c["vs_"]), col(c["done"]), col(c["success"]), gamma, lmd)
        assert np.array_equal(adv[:, 0], c["adv"])
        assert np.array_equal(vt[:, 0], c["v_target"])
        assert stats[2] == len(c["adv"])
        np.testing.assert_allclose(stats[0], c["adv"].astype(np.float64).sum(), rtol=1e-12)


@pytest.mark.gpu
def test_engine_gae_bit_exact_vs_reference():
    import torch
    from reinforcementlearningplatform_b200 import gae as G
    cs, gamma, lmd = cases()
    for c in cs:
        T = len(c["adv"])
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.reshape(T, 1))).cuda()
        adv, vt, stats = G.gae(dev(c["r"]), dev(c["vs"]), dev(c["vs_"]), dev(c["done"]), dev(c["success"]), gamma, lmd)
        assert np.array_equal(adv.cpu().numpy()[:, 0], c["adv"])
        assert np.array_equal(vt.cpu().numpy()[:, 0], c["v_target"])
        if T > 1:
            G.normalize_advantage(adv, stats)
            np.testing.assert_allclose(adv.cpu().numpy()[:, 0], c["adv_norm"], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("T,N", [(2048, 4096), (257, 1000), (3, 5), (1, 1)])
def test_engine_gae_vs_oracle_batched(T, N, oracle_lib):
    """many columns, ragged sizes: every column equals the sequential restatement bit for bit; stats match."""
    import torch
    from oracle import oracle
    from reinforcementlearningplatform_b200 import gae as G
    rng = np.random.default_rng(T * 31 + N)
    r = rng.normal(0, 1, (T, N)).astype(np.float32)
    vs = rng.normal(0, 2, (T, N)).astype(np.float32)
    vsn = rng.normal(0, 2, (T, N)).astype(np.float32)
    done = (rng.random((T, N)) < 0.01).astype(np.float32)
    succ = done * (rng.random((T, N)) < 0.5).astype(np.float32)
    adv_o, vt_o, st_o = oracle.gae(r, vs, vsn, done, succ, 0.99, 0.95)
    d = lambda a: torch.from_numpy(a).cuda()
    adv, vt, st = G.gae(d(r), d(vs), d(vsn), d(done), d(succ), 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), adv_o)
    assert np.array_equal(vt.cpu().numpy(), vt_o)
    np.testing.assert_allclose(st.cpu().numpy(), st_o, rtol=1e-10)
    # float64-carry mode (numpy 1.x behaviour): close to, not equal to, the float32 scan
    adv64, _, _ = G.gae(d(r), d(vs), d(vsn), d(done), d(succ), 0.99, 0.95, acc_mode=1)
    np.testing.assert_allclose(adv64.cpu().numpy(), adv_o, rtol=2e-4, atol=2e-4)
    if T * N > 1:
        ref = (adv_o.astype(np.float64) - adv_o.astype(np.float64).mean()) / (adv_o.astype(np.float64).std(ddof=1) + 1e-5)
        G.normalize_advantage(adv, st)
        np.testing.assert_allclose(adv.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_advantage_statistics_are_bit_reproducible_and_normalisation_divides():
    """K-GAE's (sum, sum of squares, count) come from a fixed-order reduction over per-block partial sums: two runs give
    the same bits (the atomicAdd path of round 1 did not), and agree with a float64 numpy sum to rounding.  The
    normalisation is a true float32 division by (std + eps), like `(adv - adv.mean()) / (adv.std() + 1e-5)`."""
    import torch
    from reinforcementlearningplatform_b200 import gae as G
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    T, N = 512, 40000
    mk = lambda: torch.randn((T, N), generator=g, device="cuda", dtype=torch.float32)
    r, vs, vsn = mk(), mk(), mk()
    done = (torch.rand((T, N), generator=g, device="cuda") < 0.01).float()
    succ = done * (torch.rand((T, N), generator=g, device="cuda") < 0.5).float()
    runs = [G.gae(r, vs, vsn, done, succ, 0.99, 0.95) for _ in range(4)]
    for adv, vt, st in runs[1:]:
        assert torch.equal(st, runs[0][2]) and torch.equal(adv, runs[0][0])
    adv, _, st = runs[0]
    a64 = adv.double().cpu().numpy()
    assert st[2].item() == T * N
    assert abs(st[0].item() - a64.sum()) <= 1e-9 * np.abs(a64).sum()
    assert abs(st[1].item() - (a64 ** 2).sum()) <= 1e-12 * (a64 ** 2).sum()
    mean, std = a64.mean(), a64.std(ddof=1)
    want = ((adv - np.float32(mean)) / np.float32(np.float32(std) + np.float32(1e-5))).clone()
    G.normalize_advantage(adv, st.clone())
    assert torch.equal(adv, want) or float((adv - want).abs().max()) <= 2.5e-7 * float(want.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("T,N", [(2048, 1), (1000, 7), (77, 130), (33, 4)])
def test_warp_scan_along_time_matches_float64_carry_scan(T, N):
    """acc_mode 2 (warp-level scan along time, for few columns) against acc_mode 1 (sequential float64 carry): outputs
    within one float32 ulp, statistics within 1e-12 relative; episode boundaries (done) and success flags exercised."""
    import torch
    from reinforcementlearningplatform_b200 import gae as G
    g = torch.Generator(device="cuda").manual_seed(T * 1000 + N)
    mk = lambda: torch.randn((T, N), generator=g, device="cuda", dtype=torch.float32)
    r, vs, vsn = mk(), mk(), mk()
    dn = torch.rand((T, N), generator=g, device="cuda") < 0.03
    succ = (dn & (torch.rand((T, N), generator=g, device="cuda") < 0.5)).float()
    done = dn.float()
    a1, v1, s1 = G.gae(r, vs, vsn, done, succ, 0.99, 0.95, acc_mode=1)
    a2, v2, s2 = G.gae(r, vs, vsn, done, succ, 0.99, 0.95, acc_mode=2)
    ulp = lambda x: torch.maximum(x.abs(), torch.tensor(1e-30, device="cuda")) * 2.0 ** -23
    assert bool(((a1 - a2).abs() <= ulp(a1)).all()) and bool(((v1 - v2).abs() <= ulp(v1)).all())
    assert float((a1 != a2).float().mean()) < 0.01            # all but a handful of roundings agree exactly
    assert torch.allclose(s1, s2, rtol=1e-9, atol=1e-9) and float(s2[2]) == T * N
    # and against the float32-sequential reference order: within the 1e-5 relative band of note N12
    a0, _, _ = G.gae(r, vs, vsn, done, succ, 0.99, 0.95, acc_mode=0)
    assert float(((a0 - a2).abs() / torch.clamp(a0.abs(), min=1.0)).max()) <= 1e-5
