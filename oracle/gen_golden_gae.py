"""Golden vectors for the GAE reverse scan, produced by the UNMODIFIED reference learner.

TEST INFRASTRUCTURE ONLY (container-side).  Builds algorithm/policy_base/Proximal_Policy_Optimization2 with the
reference's own critic (utils/classes.py:590-615), fills its RolloutBuffer with random transitions carrying the
done/success patterns of SURVEY.md section 8c(v), runs learn() with K_epochs = 0 (no parameter update), and captures
the locals vs, vs_, adv (before and after normalisation) and v_target with sys.settrace -- i.e. the numbers the
reference itself computes at Proximal_Policy_Optimization2.py:91-100.

    python oracle/gen_golden_gae.py   ->  tests/golden/gae.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim as R  # noqa: E402

GAMMA, LMD = 0.99, 0.95


def run_case(T, pattern, seed):
    R.install()
    with R.quiet():
        ppo_mod = R.load("algorithm.policy_base.Proximal_Policy_Optimization2")
        cls = R.load("utils.classes")
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    S, A = 4, 2
    env_msg = {'state_dim': S, 'action_dim': A, 'action_range': [[-1, 1]] * A, 'name': 'gae_fixture'}
    ppo_msg = {'gamma': GAMMA, 'K_epochs': 0, 'eps_clip': 0.2, 'buffer_size': T, 'a_lr': 1e-4, 'c_lr': 1e-3,
               'set_adam_eps': True, 'lmd': LMD, 'use_adv_norm': True, 'mini_batch_size': 64, 'entropy_coef': 0.01,
               'use_grad_clip': True, 'use_lr_decay': False, 'max_train_steps': int(5e6), 'using_mini_batch': False}
    actor = cls.PPOActor_Gaussian(state_dim=S, action_dim=A, a_min=-np.ones(A), a_max=np.ones(A))
    critic = cls.PPOCritic(state_dim=S)
    agent = ppo_mod.Proximal_Policy_Optimization2(env_msg, ppo_msg, actor, critic)
    b = agent.buffer
    b.s[:] = rng.normal(0, 1.5, (T, S))
    b.s_[:] = rng.normal(0, 1.5, (T, S))
    b.a[:] = rng.uniform(-1, 1, (T, A))
    b.a_lp[:] = rng.normal(-1, 0.3, (T, A))
    b.r[:] = rng.normal(-0.5, 2.0, (T, 1))
    done = np.zeros(T)
    if pattern == "none":
        pass
    elif pattern == "all":
        done[:] = 1
    elif pattern == "ends":
        done[0] = 1
        done[-1] = 1
    elif pattern == "episodic":
        done[np.arange(T) % 250 == 249] = 1
    else:  # random
        done[rng.random(T) < 0.02] = 1
    success = done * (rng.random(T) < 0.5)  # terminal-but-not-timeout transitions do not bootstrap
    b.done[:, 0] = done
    b.success[:, 0] = success
    cap = {}
    code = ppo_mod.Proximal_Policy_Optimization2.learn.__code__

    def tracer(frame, event, arg):
        if frame.f_code is not code:
            return None

        def local(frame, event, arg):
            if event == "line":
                loc = frame.f_locals
                if "v_target" in loc and "raw" not in cap and torch.is_tensor(loc.get("adv")):
                    cap["raw"] = loc["adv"].clone().numpy().ravel()
                    cap["v_target"] = loc["v_target"].clone().numpy().ravel()
                    cap["vs"] = loc["vs"].clone().numpy().ravel()
                    cap["vs_"] = loc["vs_"].clone().numpy().ravel()
                    cap["deltas"] = loc["deltas"].clone().numpy().ravel()
                elif "raw" in cap and torch.is_tensor(loc.get("adv")):
                    cap["norm"] = loc["adv"].clone().numpy().ravel()
            elif event == "return" and "raw" in cap and torch.is_tensor(frame.f_locals.get("adv")):
                cap["norm"] = frame.f_locals["adv"].clone().numpy().ravel()
            return local
        return local

    sys.settrace(tracer)
    try:
        with R.quiet():
            agent.learn(0, buf_num=1)
    finally:
        sys.settrace(None)
    return dict(r=b.r[:, 0].astype(np.float32), done=done.astype(np.float32), success=success.astype(np.float32),
                vs=cap["vs"], vs_=cap["vs_"], adv=cap["raw"], v_target=cap["v_target"], adv_norm=cap["norm"],
                deltas=cap["deltas"])


def main():
    cases = [(2048, "random", 1), (2048, "none", 2), (2048, "all", 3), (2048, "ends", 4), (1000, "episodic", 5),
             (37, "random", 6), (1, "none", 7)]
    out = {"gamma": np.array(GAMMA), "lmd": np.array(LMD), "numpy": np.array(np.__version__),
           "torch": np.array(torch.__version__)}
    for k, (T, pat, seed) in enumerate(cases):
        c = run_case(T, pat, seed)
        for name, v in c.items():
            out[f"c{k}_{name}"] = v
        out[f"c{k}_pattern"] = np.array(pat)
        print(f"case {k}: T={T} {pat}: adv[:3]={c['adv'][:3]} dtype={c['adv'].dtype} norm mean={c['adv_norm'].mean():.2e}")
    out["n_cases"] = np.array(len(cases))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "gae.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) // 1000, "kB")


if __name__ == "__main__":
    main()
