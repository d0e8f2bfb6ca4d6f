"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (sharding contract, flat gradient all-reduce, global
advantage statistics) and -- through the C oracle, which implements the same counter-based reset draws as the CUDA
kernels -- that trajectories do not depend on how instances are sharded over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from reinforcementlearningplatform_b200 import dist as D
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # 1. sharding contract
    n_total = 11
    n_local, off = D.shard(n_total, rank, world)
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local, off]))
    # 2. flat gradient average == mean of the per-rank gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    D.broadcast_parameters([net])
    x = torch.full((3, 4), float(rank + 1))
    net(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    red = D.FlatGradAllReducer(net.parameters())
    red()
    gathered = []
    for g in local:
        buf = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(buf, g)
        gathered.append(sum(buf) / world)
    ok_grad = all(torch.allclose(p.grad, m, atol=1e-6) for p, m in zip(net.parameters(), gathered))
    # 3. global advantage statistics: sum over ranks of (sum, sum sq, count)
    adv = torch.arange(5, dtype=torch.float64) + 10 * rank
    st = torch.tensor([adv.sum(), (adv ** 2).sum(), float(adv.numel())], dtype=torch.float64)
    D.allreduce_stats(st)
    # 4. oracle trajectories with the rank's offset (compared with the unsharded run in the parent)
    from oracle import oracle
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200 import _lib
    host = rlp.CartPole(n_envs=n_local, host_only=True)
    orc = oracle.OracleEnv(_lib.CARTPOLE, host._params, n_local, 4, 4, 1, 0, seed=99, env_index_offset=off, auto_reset=True)
    orc.reset()
    rng = np.random.default_rng(5)
    acts = rng.uniform(-8, 8, size=(40, 1, n_total))
    traj = []
    for t in range(40):
        orc.step(acts[t][:, off:off + n_local])
        traj.append(orc.state.copy())
    # 5. running normaliser: per-rank batch statistics gathered in rank order, merged by the rule of csrc/norm.cu
    from reinforcementlearningplatform_b200.normalization import gather_batch_stats, merge_stats_reference
    xs = np.random.default_rng(100 + rank).normal(rank, 1.0 + rank, (3, 500 + 100 * rank))
    mine = torch.from_numpy(np.stack([np.full(3, xs.shape[1], dtype=np.float64), xs.mean(1),
                                      ((xs - xs.mean(1, keepdims=True)) ** 2).sum(1)])[None])
    allb = gather_batch_stats(mine).numpy()
    run = merge_stats_reference((np.zeros(3), np.zeros(3), np.zeros(3)), [tuple(b) for b in allb])
    q.put((rank, [s.tolist() for s in sizes], ok_grad, st.tolist(), np.stack(traj), red.nbytes, [r.tolist() for r in run]))
    dist.destroy_process_group()


def test_world2_gloo(oracle_lib):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # sharding: contiguous cover of 11 instances
    assert res[0][1] == [[6, 0], [5, 6]]
    assert all(r[2] for r in res)
    # stats: both ranks hold the global sums
    adv = np.concatenate([np.arange(5) + 10 * r for r in range(world)]).astype(np.float64)
    for r in res:
        np.testing.assert_allclose(r[3], [adv.sum(), (adv ** 2).sum(), adv.size])
        assert r[5] == (4 * 8 + 8 + 8 * 2 + 2) * 4
    # normaliser: both ranks hold the same merged statistics, equal to those of the concatenated batch
    xs = np.concatenate([np.random.default_rng(100 + r).normal(r, 1.0 + r, (3, 500 + 100 * r)) for r in range(world)], axis=1)
    assert res[0][6] == res[1][6]
    np.testing.assert_allclose(res[0][6][0], np.full(3, xs.shape[1]))
    np.testing.assert_allclose(res[0][6][1], xs.mean(1), rtol=1e-13)
    np.testing.assert_allclose(np.array(res[0][6][2]) / xs.shape[1], xs.var(1), rtol=1e-12)
    # trajectories: the two shards side by side == one unsharded run (auto-reset draws keyed by the global index)
    from oracle import oracle
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200 import _lib
    host = rlp.CartPole(n_envs=11, host_only=True)
    orc = oracle.OracleEnv(_lib.CARTPOLE, host._params, 11, 4, 4, 1, 0, seed=99, env_index_offset=0, auto_reset=True)
    orc.reset()
    rng = np.random.default_rng(5)
    acts = rng.uniform(-8, 8, size=(40, 1, 11))
    for t in range(40):
        orc.step(acts[t])
        both = np.concatenate([res[0][4][t], res[1][4][t]], axis=1)
        assert np.array_equal(both, orc.state), t


def test_shard_covers_everything():
    from reinforcementlearningplatform_b200.dist import shard
    for n in (1, 7, 8, 1 << 20, (1 << 20) + 3):
        for w in (1, 2, 4, 8):
            parts = [shard(n, r, w) for r in range(w)]
            assert sum(p[0] for p in parts) == n
            off = 0
            for nl, o in parts:
                assert o == off
                off += nl
