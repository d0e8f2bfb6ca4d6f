#!/usr/bin/env python
"""PPO2 on a vector env, entirely on the device: the vector form of
demonstration/PPO2/PPO2-4-CartPoleAngleOnly/train.py (collection loop :186-216, learn :219-222) and of
PPO2-4-UavFntsmcParamPos/train.py.  Under torchrun every rank owns a shard of the instances and gradients are
all-reduced (DPPO2 as synchronous data parallelism).

    python examples/train_ppo2_vec.py --env cartpole_angleonly --envs 4096 --iters 30
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_ppo2_vec.py --env soi
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import reinforcementlearningplatform_b200 as rlp  # noqa: E402
from reinforcementlearningplatform_b200 import dist as D  # noqa: E402
from reinforcementlearningplatform_b200.ppo2 import VecPPO2, reference_nets  # noqa: E402

ENVS = {
    "cartpole_angleonly": (lambda **kw: rlp.CartPoleAngleOnly(variant="ppo2", **kw), dict(std=1.2, mean_act="identity")),
    "soi": (lambda **kw: rlp.SecondOrderIntegration(**kw), dict(std=0.8, mean_act="identity")),
    "fas": (lambda **kw: rlp.Flight_Attitude_Simulator(variant="ppo2", **kw), dict(std=0.8, mean_act="identity")),
    "uav_pos": (lambda **kw: rlp.UavPosCtrlRL(random_trajectory=True, **kw), dict(std=0.45, mean_act="relu")),
    # BASELINE config #5: the DPPO2 copy of UGVForwardObstacleAvoidance (41-dim observation with the 37-ray laser);
    # under torchrun this is DPPO2 as synchronous data parallelism (flat gradient all-reduce per mini-batch)
    "ugvo": (lambda **kw: rlp.UGVForwardObstacleAvoidance(variant="dppo2", **kw), dict(std=0.8, mean_act="identity")),
}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="cartpole_angleonly", choices=sorted(ENVS))
    ap.add_argument("--envs", type=int, default=4096, help="instances in total (sharded over ranks)")
    ap.add_argument("--steps", type=int, default=64, help="time steps per rollout (buffer_size)")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("--seed", type=int, default=3407)   # the reference script's seed (train.py:36)
    args = ap.parse_args(argv)
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(args.seed)
    n_local, off = D.shard(args.envs, rank, world)
    make, net_kw = ENVS[args.env]
    env = make(n_envs=n_local, device=dev, dtype=torch.float64, io_dtype=torch.float32, auto_reset=True, seed=args.seed,
               env_index_offset=off)
    actor, critic = reference_nets(env.state_dim, env.action_dim, dev, init_std=net_kw["std"], mean_act=net_kw["mean_act"])
    agent = VecPPO2(env, actor, critic, {"buffer_size": args.steps, "K_epochs": args.epochs,
                                          "mini_batch_size": 16384}, std=net_kw["std"],
                    seed=args.seed)
    env.reset(True)
    log = []
    for it in range(args.iters):
        t0 = time.perf_counter()
        mean_r = agent.collect()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        losses = agent.learn()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        done_rate = float(agent.buffer.done.float().mean())
        timeout_share = float(((agent.buffer.flag == env.TIMEOUT_FLAG) & (agent.buffer.done != 0)).float().sum() /
                              max(1.0, float(agent.buffer.done.sum())))
        row = {"iter": it, "mean_reward": mean_r, "done_rate": done_rate, "timeout_share": timeout_share,
               "collect_env_steps_per_s": args.steps * n_local * world / (t1 - t0), "learn_s": t2 - t1, **losses}
        log.append(row)
        if rank == 0:
            print(json.dumps(row))
    if world > 1:
        torch.distributed.destroy_process_group()
    return log


if __name__ == "__main__":
    main()
