"""Running normalisation (utils/classes.py:626-656) and the PPO v1 return scan (Proximal_Policy_Optimization.py:113-119):
the C restatement against vectors recorded from the reference's own classes (CPU), the CUDA kernels against both (GPU).
Bit-exact wherever the reference semantics apply (sample-by-sample feeding); the batched mode -- many instances per
step, a semantics the reference does not have -- is checked against float64 numpy statistics within 1e-12."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import GOLDEN


def golden():
    with np.load(os.path.join(GOLDEN, "norm.npz")) as z:
        return {k: z[k] for k in z.files}


def norm_cases(g):
    return [{n: g[f"n{k}_{n}"] for n in ("x", "y", "run", "x_eval", "y_eval")} for k in range(int(g["n_norm"]))]


def ret_cases(g):
    return [{n: g[f"r{k}_{n}"] for n in ("r", "done", "ret")} for k in range(int(g["n_ret"]))]


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


# ----------------------------------------------------------------------------------------------------------- CPU
def test_oracle_norm_bit_exact_vs_reference(oracle_lib):
    from oracle import oracle
    for c in norm_cases(golden()):
        y, run = oracle.norm_seq(c["x"].T)
        assert same_bits(y.T, c["y"])  # incl. the -0.0 of a negative first sample
        assert same_bits(run, c["run"])
        ye, run2 = oracle.norm_seq(c["x_eval"].T, run=run, update=False)
        assert same_bits(ye.T, c["y_eval"])
        assert same_bits(run2, run)


def test_oracle_mc_returns_bit_exact_vs_reference(oracle_lib):
    from oracle import oracle
    g = golden()
    for c in ret_cases(g):
        ret = oracle.mc_returns(c["r"].reshape(-1, 1), c["done"].reshape(-1, 1), float(g["gamma"]))
        assert np.array_equal(ret[:, 0], c["ret"])


def test_merge_rule_host_restatement():
    """The merge rule of csrc/norm.cu, restated in numpy: one-sample batches reproduce the reference recurrence bit
    for bit; merging shard statistics in rank order equals the statistics of the concatenated batch."""
    from reinforcementlearningplatform_b200.normalization import merge_stats_reference as merge
    c = norm_cases(golden())[0]
    dim = c["x"].shape[1]
    run = (np.zeros(dim), np.zeros(dim), np.zeros(dim))
    for row in c["x"]:
        run = merge(run, [(np.ones(dim), row, np.zeros(dim))])
    assert same_bits(run[0], c["run"][0]) and same_bits(run[1], c["run"][1]) and same_bits(run[2], c["run"][2])
    rng = np.random.default_rng(0)
    x = rng.normal(3.0, 2.0, (dim, 4096))
    shards = np.split(x, 4, axis=1)
    st = [(np.full(dim, s.shape[1]), s.mean(1), ((s - s.mean(1, keepdims=True)) ** 2).sum(1)) for s in shards]
    n, mean, S = merge((np.zeros(dim), np.zeros(dim), np.zeros(dim)), st)
    np.testing.assert_allclose(mean, x.mean(1), rtol=1e-13)
    np.testing.assert_allclose(S / n, x.var(1), rtol=1e-12)


# ----------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_engine_norm_seq_bit_exact_vs_reference():
    import torch
    import reinforcementlearningplatform_b200 as rlp
    for c in norm_cases(golden()):
        dim = c["x"].shape[1]
        nz = rlp.Normalization(dim)
        y = nz.seq(torch.from_numpy(np.ascontiguousarray(c["x"].T)).cuda())
        assert same_bits(y.cpu().numpy().T, c["y"])
        ms = nz.running_ms
        assert ms.n == c["run"][0, 0] and same_bits(ms.mean, c["run"][1]) and same_bits(ms.S, c["run"][2])
        assert same_bits(ms.std, c["run"][3])
        ye = nz.seq(torch.from_numpy(np.ascontiguousarray(c["x_eval"].T)).cuda(), update=False)
        assert same_bits(ye.cpu().numpy().T, c["y_eval"])


@pytest.mark.gpu
def test_engine_norm_batch_of_one_is_the_reference_update():
    """normalize_soa with N = 1, called once per sample like the train loop: bit-identical to the reference."""
    import torch
    import reinforcementlearningplatform_b200 as rlp
    for c in norm_cases(golden())[2:4] + [{k: v[:200] if k in ("x", "y") else v for k, v in norm_cases(golden())[0].items()}]:
        dim = c["x"].shape[1]
        nz = rlp.Normalization(dim, sync=False)
        xs = torch.from_numpy(c["x"]).cuda()
        ys = [nz(xs[t], update=True) for t in range(xs.shape[0])]
        y = torch.stack(ys).cpu().numpy()
        assert same_bits(y, c["y"])
        if c["x"].shape[0] == c["run"][0, 0]:
            assert same_bits(nz.running_ms.mean, c["run"][1]) and same_bits(nz.running_ms.S, c["run"][2])


@pytest.mark.gpu
@pytest.mark.parametrize("dim,N,dtype", [(6, 1 << 20, "f32"), (6, 100003, "f64"), (1, 65536, "f64"), (41, 5000, "f32"),
                                          (3, 7, "f64")])
def test_engine_norm_batched_vs_numpy(dim, N, dtype):
    import torch
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200.normalization import merge_stats_reference as merge
    rng = np.random.default_rng(dim * 1000 + N)
    tdt = torch.float32 if dtype == "f32" else torch.float64
    nz = rlp.Normalization(dim, sync=False)
    run = (np.zeros(dim), np.zeros(dim), np.zeros(dim))
    for step in range(3):
        x = (rng.normal(1.0 + step, 2.0, (dim, N)) * np.arange(1, dim + 1)[:, None]).astype(np.float32 if dtype == "f32" else np.float64)
        xd = torch.from_numpy(x).cuda()
        y = nz.normalize_soa(xd)
        x64 = x.astype(np.float64)
        run = merge(run, [(np.full(dim, N), x64.mean(1), ((x64 - x64.mean(1, keepdims=True)) ** 2).sum(1))])
        ms = nz.running_ms
        assert ms.n == run[0][0]
        np.testing.assert_allclose(ms.mean, run[1], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(ms.S, run[2], rtol=1e-11)
        ref = (x64 - run[1][:, None]) / (np.sqrt(run[2] / run[0])[:, None] + 1e-8)
        assert y.dtype == tdt
        np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=2e-6 if dtype == "f32" else 1e-11, atol=2e-6 if dtype == "f32" else 1e-11)
    # evaluation pass: statistics untouched
    before = nz.state_dict()["run"].clone()
    nz.normalize_soa(xd, update=False)
    assert torch.equal(before, nz.state_dict()["run"])
    # reference orientation: [N, dim] views of SoA buffers go through without a copy and come back as [N, dim]
    y2 = nz(xd.t(), update=False)
    assert y2.shape == (N, dim)
    assert torch.equal(y2.t(), nz.normalize_soa(xd, update=False))


@pytest.mark.gpu
def test_engine_norm_rank_ordered_merge_through_abi():
    """n_batches > 1 (what a rank sees after the all-gather): equals the host restatement of the merge rule."""
    import torch
    from reinforcementlearningplatform_b200 import _lib
    from reinforcementlearningplatform_b200.normalization import merge_stats_reference as merge
    lib = _lib.load()
    dim, G, N = 6, 4, 1000
    rng = np.random.default_rng(5)
    shards = [rng.normal(g, 1.0 + g, (dim, N)) for g in range(G)]
    batches = np.stack([np.stack([np.full(dim, N), s.mean(1), ((s - s.mean(1, keepdims=True)) ** 2).sum(1)]) for s in shards])
    run0 = np.stack([np.full(dim, 10.0), rng.normal(0, 1, dim), rng.uniform(5, 9, dim)])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    b_d, r_in, r_out = d(batches), d(run0), torch.zeros(3, dim, dtype=torch.float64, device="cuda")
    x = d(shards[1])
    y = torch.empty_like(x)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.b200_norm_merge_apply(_lib.F64, N, dim, p(x), p(y), p(b_d), G, p(r_in), p(r_out), 1, 1e-8, None), "merge")
    torch.cuda.synchronize()
    n, mean, S = merge(tuple(run0), [tuple(b) for b in batches])
    out = r_out.cpu().numpy()
    assert same_bits(out[0], n) and same_bits(out[1], mean) and same_bits(out[2], S)
    ref = (shards[1] - mean[:, None]) / (np.sqrt(S / n)[:, None] + 1e-8)
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-14, atol=1e-15)


@pytest.mark.gpu
def test_engine_mc_returns(oracle_lib):
    import torch
    from oracle import oracle
    from reinforcementlearningplatform_b200 import gae as G
    g = golden()
    gamma = float(g["gamma"])
    for c in ret_cases(g):
        T = len(c["ret"])
        ret = G.mc_returns(torch.from_numpy(c["r"].reshape(T, 1)).cuda(), torch.from_numpy(c["done"].reshape(T, 1)).cuda(), gamma)
        assert np.array_equal(ret.cpu().numpy()[:, 0], c["ret"])
    rng = np.random.default_rng(9)
    for T, N in [(2048, 4096), (257, 1000), (3, 5)]:
        r = rng.normal(0, 1, (T, N))
        done = (rng.random((T, N)) < 0.02).astype(np.uint8)
        ref = oracle.mc_returns(r, done, gamma)
        out = G.mc_returns(torch.from_numpy(r).cuda(), torch.from_numpy(done).cuda(), gamma)
        assert np.array_equal(out.cpu().numpy(), ref)
        out32 = G.mc_returns(torch.from_numpy(r.astype(np.float32)).cuda(), torch.from_numpy(done).cuda(), gamma)
        assert np.array_equal(out32.cpu().numpy(), oracle.mc_returns(r.astype(np.float32).astype(np.float64), done, gamma))


@pytest.mark.gpu
@pytest.mark.parametrize("T,N", [(16, 4096), (33, 1000), (4, 1)])
def test_engine_norm_rows_equals_row_by_row(T, N):
    """normalize_rows over a [T, N] reward column == T calls of normalize_soa(row) (row t enters the statistics after
    rows < t and is normalised with the statistics after row t), to float64 rounding of the batch sums."""
    import torch
    import reinforcementlearningplatform_b200 as rlp
    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn((T, N), generator=g, device="cuda") * 3.0 - 7.0).float()
    a, b = rlp.Normalization(1, device="cuda", sync=False), rlp.Normalization(1, device="cuda", sync=False)
    warm = torch.randn(64, generator=g, device="cuda").float()      # both start from the same non-empty statistics
    a.normalize_soa(warm.clone())
    b.normalize_soa(warm.clone())
    for rep in range(2):                                            # twice: the second rollout continues the first
        xa, xb = x.clone() + rep, x.clone() + rep
        for t in range(T):
            a.normalize_soa(xa[t], out=xa[t])
        b.normalize_rows(xb, out=xb)
        torch.cuda.synchronize()
        np.testing.assert_allclose(xb.cpu().numpy(), xa.cpu().numpy(), rtol=2e-6, atol=2e-6)
        ra, rb = a._run.cpu().numpy(), b._run.cpu().numpy()
        assert ra[0, 0] == rb[0, 0] == 64 + (rep + 1) * T * N
        np.testing.assert_allclose(rb, ra, rtol=1e-12)
    # an empty normaliser: the first row is the first batch (`self.mean = x` branch of the merge)
    c, d = rlp.Normalization(1, device="cuda", sync=False), rlp.Normalization(1, device="cuda", sync=False)
    xa, xb = x.clone(), x.clone()
    for t in range(T):
        c.normalize_soa(xa[t], out=xa[t])
    d.normalize_rows(xb, out=xb)
    np.testing.assert_allclose(xb.cpu().numpy(), xa.cpu().numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(d._run.cpu().numpy(), c._run.cpu().numpy(), rtol=1e-12)
