"""Golden vectors for the 256-wide DPPO2 nets, produced by the UNMODIFIED classes of the reference demo.

TEST INFRASTRUCTURE ONLY (container-side).  Loads demonstration/DPPO2/DPPO2-4-UGVForwardObstacleAvoidance/train.py by
path (its training code sits under `if __name__ == '__main__'`, so only the class definitions run), builds its
PPOActor_Gaussian (:26-76, 41-256-256-2, mean = tanh(mean_layer) * gain + off) and PPOCritic (:79-107, 41-256-256-1) for
the demo's own env (state_dim 41, action_range, init_std = range / 2 / 3, train.py:128-144), and records mean, value and
the choose_action conventions of Distributed_PPO2.Worker.choose_action (:106-113) with a recorded N(0, 1) draw.

    python oracle/gen_golden_policy_wide.py   ->  tests/golden/policy_wide.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim as R  # noqa: E402

DEMO = "demonstration/DPPO2/DPPO2-4-UGVForwardObstacleAvoidance"


def main():
    R.install()
    sys.path.insert(0, os.path.join(R.REF, DEMO))
    tr = R.load_file(DEMO + "/train.py", "dppo2_ugvo_train")
    with R.quiet():
        env = tr.UGVForwardObstacleAvoidance()
    ar = np.array(env.action_range, dtype=np.float64)
    std0 = (ar[:, 1] - ar[:, 0]) / 2 / 3                                   # train.py:128
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    actor = tr.PPOActor_Gaussian(state_dim=env.state_dim, action_dim=env.action_dim, a_min=ar[:, 0], a_max=ar[:, 1],
                                 init_std=std0, use_orthogonal_init=True)
    critic = tr.PPOCritic(state_dim=env.state_dim, use_orthogonal_init=True)
    n = 400
    with torch.no_grad():
        for m in list(actor.modules()) + list(critic.modules()):
            if isinstance(m, torch.nn.Linear):
                m.bias.copy_(torch.from_numpy(rng.normal(0, 0.3, m.bias.shape).astype(np.float32)))
        actor.mean_layer.weight.mul_(150.0)   # gain-0.01 init gives means ~off: scale up so the tanh head is exercised
        s = torch.from_numpy(rng.uniform(-1.0, 1.0, (n, env.state_dim)).astype(np.float32) * float(env.static_gain))
        eps = torch.from_numpy(rng.normal(0, 1, (n, env.action_dim)).astype(np.float32))
        mean = actor(s)
        dist = actor.get_dist(s)
        a = mean + actor.std * eps                                          # Normal.sample() = loc + scale * N(0, 1)
        a = torch.maximum(torch.minimum(a, actor.a_max), actor.a_min)        # Distributed_PPO2.py:111
        lp = dist.log_prob(a)                                               # :112
        v = critic.net(s)
    out = dict(s=s.numpy(), eps=eps.numpy(), mean=mean.numpy(), action=a.numpy(), log_prob=lp.numpy(), value=v.numpy()[:, 0],
               a_min=actor.a_min.numpy(), a_max=actor.a_max.numpy(), std=actor.std.numpy().astype(np.float32),
               torch=np.array(torch.__version__))
    for lname in ("fc1", "fc2", "mean_layer"):
        out[f"actor_{lname}_w"] = getattr(actor, lname).weight.detach().numpy().copy()
        out[f"actor_{lname}_b"] = getattr(actor, lname).bias.detach().numpy().copy()
    for k, idx in enumerate((0, 2, 4)):
        out[f"critic_l{k}_w"] = critic.net[idx].weight.detach().numpy().copy()
        out[f"critic_l{k}_b"] = critic.net[idx].bias.detach().numpy().copy()
    print(f"S={env.state_dim} A={env.action_dim} range {ar.tolist()} std {std0.tolist()} mean range "
          f"[{out['mean'].min():.3f}, {out['mean'].max():.3f}] clamped "
          f"{float(np.mean((out['action'] == out['a_min']) | (out['action'] == out['a_max']))):.2f}")
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "policy_wide.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) // 1000, "kB")


if __name__ == "__main__":
    main()
