"""GPU, two ranks on ONE device over gloo (NCCL refuses two ranks per GPU; gloo all-reduces CUDA tensors): a DPPO2-style
iteration of the device pipeline with the instances sharded over the ranks -- collect (K-POLICY -> step kernel -> K-NORM
with all-gathered statistics), K-GAE with the all-reduced advantage statistics, K-LEARN with ONE flat gradient all-reduce
per mini-batch (the synchronous form of the gradient push of demonstration/DPPO2/*/Distributed_PPO2.py:86-104) -- and the
invariants that make it data parallel: both ranks issue the same number of collectives although their shards differ by
one instance, and they hold bit-identical parameters after every learn()."""
import os
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200 import dist as D
    from reinforcementlearningplatform_b200.ppo2 import VecPPO2, reference_nets
    n_total = 2049                                       # shards of 1025 and 1024 instances
    n_local, off = D.shard(n_total, rank, world)
    env = rlp.CartPoleAngleOnly(n_envs=n_local, variant="ppo2", io_dtype=torch.float32, auto_reset=True, seed=3,
                                env_index_offset=off)
    torch.manual_seed(100 + rank)                        # different initial weights: rank 0's are broadcast
    actor, critic = reference_nets(env.state_dim, env.action_dim, "cuda", init_std=0.8, mean_act="identity")
    agent = VecPPO2(env, actor, critic, {"buffer_size": 16, "K_epochs": 2, "mini_batch_size": 4096}, std=0.8, seed=7)
    assert agent.fused is not None
    env.reset(True)
    sums = []
    for it in range(2):
        agent.collect()
        out = agent.learn()
        flat = agent.fused.params.flat
        mine = flat.detach().cpu()
        both = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        sums.append((bool(torch.equal(both[0], both[1])), float(flat.double().abs().sum()), out["actor_loss"], out["critic_loss"]))
    # B = 16 * n_local differs between the ranks (16400 vs 16384 samples): the mini-batch count was agreed (MIN over ranks)
    q.put((rank, n_local, off, agent.fused.step_count, sums))
    dist.destroy_process_group()


def test_two_rank_ppo2_iteration_keeps_parameters_identical():
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (r0, n0, o0, steps0, s0), (r1, n1, o1, steps1, s1) = res
    assert (n0, o0, n1, o1) == (1025, 0, 1024, 1025)
    assert steps0 == steps1 and steps0 == 2 * 2 * 4          # 2 iterations x K_epochs 2 x ceil(16384 / 4096) mini-batches
    for a, b in zip(s0, s1):
        assert a[0] and b[0]                                  # identical parameters on both ranks after every learn()
        assert a[1] == b[1]
        assert all(map(lambda v: v == v and abs(v) < 1e6, (a[2], a[3], b[2], b[3])))
    assert s0[0][1] != s0[1][1]                               # and they did change between the iterations
