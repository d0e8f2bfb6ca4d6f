"""One on-device collection rollout of VecPPO2 (1 M UAV-pos instances, 16 steps) for launch lists / ncu captures:
python tools/collect_once.py [rollouts]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from reinforcementlearningplatform_b200.ppo2 import VecPPO2, reference_nets  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
n = bench.WORKLOADS["uav_pos"]["n"]
env = bench.make_env("uav_pos", n, dev, 0, torch.float64, io_dtype=torch.float32)
env.reset(True)
actor, critic = reference_nets(env.state_dim, env.action_dim, dev, init_std=0.45)
agent = VecPPO2(env, actor, critic, {"buffer_size": 16, "K_epochs": 1}, std=0.45)
for _ in range(reps):
    agent.collect()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
agent.collect()
e1.record()
torch.cuda.synchronize()
print(f"collect: {e0.elapsed_time(e1) / 16:.4f} ms per step of {n} instances")
