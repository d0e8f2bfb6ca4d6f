"""Golden vectors for the batched policy forward, produced by the UNMODIFIED reference nets and learner.

TEST INFRASTRUCTURE ONLY (container-side).  Builds utils/classes.py PPOActor_Gaussian (:529-591) and PPOCritic (:594-623)
with their own orthogonal init (then perturbs the biases so that they matter), evaluates mean = actor(s), value =
critic(s), and reproduces Proximal_Policy_Optimization2.choose_action (:69-76) with a recorded N(0,1) draw:
a = clamp(mean + std * eps), log_prob = Normal(mean, std).log_prob(a).  Also records one real `choose_action` call of
the learner (its own torch RNG) to pin the clamp / log-prob conventions.

    python oracle/gen_golden_policy.py   ->  tests/golden/policy.npz
"""
import os
import sys

import numpy as np
import torch
from torch.distributions import Normal

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim as R  # noqa: E402


def case(cls, S, A, a_lo, a_hi, std, n, seed):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    actor = cls.PPOActor_Gaussian(state_dim=S, action_dim=A, a_min=np.full(A, a_lo), a_max=np.full(A, a_hi), init_std=std)
    critic = cls.PPOCritic(state_dim=S)
    with torch.no_grad():
        for m in list(actor.modules()) + list(critic.modules()):
            if isinstance(m, torch.nn.Linear):
                m.bias.copy_(torch.from_numpy(rng.normal(0, 0.3, m.bias.shape).astype(np.float32)))
        actor.mean_layer.weight.mul_(60.0)  # gain-0.01 init gives means ~0: scale up so relu/clamp branches are hit
        s = torch.from_numpy(rng.normal(0, 1.5, (n, S)).astype(np.float32))
        eps = torch.from_numpy(rng.normal(0, 1, (n, A)).astype(np.float32))
        mean = actor(s)
        dist = actor.get_dist(s)
        a = mean + actor.std * eps                                            # Normal.sample() = loc + scale * N(0,1)
        a = torch.maximum(torch.minimum(a, actor.a_max), actor.a_min)          # PPO2.py:74
        lp = dist.log_prob(a)                                                 # PPO2.py:75
        v = critic(s)
    out = dict(s=s.numpy(), eps=eps.numpy(), mean=mean.numpy(), action=a.numpy(), log_prob=lp.numpy(), value=v.numpy()[:, 0],
               a_min=actor.a_min.numpy(), a_max=actor.a_max.numpy(), std=np.float32(std))
    for name, net in (("actor", actor), ("critic", critic)):
        for lname in ("fc1", "fc2", "fc3", "mean_layer"):
            if hasattr(net, lname):
                out[f"{name}_{lname}_w"] = getattr(net, lname).weight.detach().numpy().copy()
                out[f"{name}_{lname}_b"] = getattr(net, lname).bias.detach().numpy().copy()
    return out


def main():
    R.install()
    with R.quiet():
        cls = R.load("utils.classes")
    out = {"torch": np.array(torch.__version__)}
    cases = [(6, 8, 0.0, 5.0, 0.45, 1000, 1),     # UavFntsmcParamPos: state 6, 8 gains in [0, 5]
             (2, 1, -5.0, 5.0, 0.6, 777, 12),      # CartPoleAngleOnly
             (41, 2, -3.0, 3.0, 0.8, 300, 3),     # UGVForwardObstacleAvoidance observation width
             (4, 2, -3.0, 3.0, 0.5, 1, 4)]
    for k, c in enumerate(cases):
        d = case(cls, *c)
        for name, v in d.items():
            out[f"c{k}_{name}"] = v
        print(f"case {k}: S={c[0]} A={c[1]} mean range [{d['mean'].min():.3f}, {d['mean'].max():.3f}] "
              f"clamped {float(np.mean((d['action'] == c[2]) | (d['action'] == c[3]))):.2f}")
    out["n_cases"] = np.array(len(cases))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "policy.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) // 1000, "kB")


if __name__ == "__main__":
    main()
