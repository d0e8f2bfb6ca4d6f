"""Times the UNMODIFIED reference (pure Python) in the build container: BASELINE.json config #1 and the per-env
step loops of BASELINE.md section 2.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (container-side; /root/reference does not exist on the GPU box, which is why
bench.py's CPU arm times the C restatement instead and quotes these numbers as `reference_python_loop`).

  * config #1: the collection loop and learner of demonstration/PPO2/PPO2-4-CartPoleAngleOnly/train.py (:138-222),
    driven headless -- reference env (the demo copy), reference PPOActor_Gaussian / PPOCritic / PPO2, reference
    Normalization -- seed 3407 (the script's own, :36), `epochs` buffers of int(timeMax / dt) * 4 = 1000 transitions
    each followed by agent.learn(); reports collection env-steps/s and learn() seconds per epoch.
  * per env: `seconds` of step_update with uniform random actions, reset(True) on terminal, one core.

    python oracle/bench_reference_python.py [--epochs 5] [--seconds 2]  ->  profiles/r1/reference_python_cpu.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_adapters as A  # noqa: E402
from oracle import ref_shim as R  # noqa: E402


def config1(epochs: int):
    R.install()
    with R.quiet():
        envmod = R.load_file("demonstration/PPO2/PPO2-4-CartPoleAngleOnly/cartpole_angleonly.py", "ref_cfg1_env")
        cls = R.load("utils.classes")
        ppo = R.load("algorithm.policy_base.Proximal_Policy_Optimization2")
    np.random.seed(3407)
    torch.manual_seed(3407)
    with R.quiet():
        env = envmod.CartPoleAngleOnly(0.)
    reward_norm = cls.Normalization(shape=1)
    env_msg = {'state_dim': env.state_dim, 'action_dim': env.action_dim, 'name': env.name, 'action_range': env.action_range}
    ppo_msg = {'gamma': 0.999, 'K_epochs': 30, 'eps_clip': 0.2, 'buffer_size': int(env.timeMax / env.dt) * 4,
               'state_dim': env.state_dim, 'action_dim': env.action_dim, 'a_lr': 3e-4, 'c_lr': 1e-3, 'set_adam_eps': True,
               'lmd': 0.95, 'use_adv_norm': True, 'mini_batch_size': 64, 'entropy_coef': 0.01, 'use_grad_clip': True,
               'use_lr_decay': True, 'max_train_steps': int(5e6), 'using_mini_batch': False}   # train.py:141-158
    ar = np.array(env.action_range)
    agent = ppo.Proximal_Policy_Optimization2(
        env_msg=env_msg, ppo_msg=ppo_msg,
        actor=cls.PPOActor_Gaussian(state_dim=env.state_dim, action_dim=env.action_dim, a_min=ar[:, 0], a_max=ar[:, 1],
                                    init_std=env.fm / 3, use_orthogonal_init=True),
        critic=cls.PPOCritic(state_dim=env.state_dim, use_orthogonal_init=True))
    rows, total = [], 0
    env.is_terminal = True
    sumr, sumr_list = 0., []
    for ep in range(epochs):
        idx, t0 = 0, time.perf_counter()
        with R.quiet():
            while idx < agent.buffer.batch_size:                      # train.py:186-216
                if env.is_terminal:
                    sumr_list.append(sumr)
                    sumr = 0.
                    env.reset(True)
                else:
                    env.current_state = env.next_state.copy()
                    a, a_lp = agent.choose_action(env.current_state)
                    env.step_update(a)
                    sumr += env.reward
                    success = 1 if (env.is_terminal and env.terminal_flag != 3) else 0
                    agent.buffer.append(s=env.current_state, a=a, log_prob=a_lp, r=reward_norm(env.reward),
                                        s_=env.next_state, done=1.0 if env.is_terminal else 0.0, success=success, index=idx)
                    idx += 1
            t1 = time.perf_counter()
            total += idx
            agent.learn(total, buf_num=1)                              # train.py:219-222
            t2 = time.perf_counter()
        rows.append({"epoch": ep, "transitions": idx, "collect_s": t1 - t0, "collect_env_steps_per_s": idx / (t1 - t0),
                     "learn_s": t2 - t1})
    return {"what": "PPO2-4-CartPoleAngleOnly/train.py loop, reference env + learner, 1 core", "epochs": rows,
            "mean_collect_env_steps_per_s": float(np.mean([r["collect_env_steps_per_s"] for r in rows])),
            "mean_learn_s": float(np.mean([r["learn_s"] for r in rows])),
            "episodes_finished": len(sumr_list) - 1}


def per_env(seconds: float):
    out = {}
    names = ["cartpole", "cartpole_angleonly_env", "fas", "fas_discrete", "soi", "ballbalancer", "twolink", "ugv_forward",
             "ugvo", "uav_att_rand", "uav_pos"]
    # (the DPPO2 copy of UGVForwardObstacleAvoidance is left out: its 15-obstacle map generation retries without bound,
    #  map.py:171-172, and does not terminate for some numpy seeds)
    for name in names:
        ad = A.REGISTRY[name][0]()
        rng = np.random.default_rng(1)
        np.random.seed(1)
        with R.quiet():
            env = ad.make()
            ad.reset(env)
            for t in range(50):
                ad.step(env, ad.sample_action(rng, t, 0, env), ad.sample_dis(rng, t, 0, env) if ad.D else None)
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < seconds:
                _, _, _, done, _ = ad.step(env, rng.uniform(ad.action_lo, ad.action_hi) if ad.action_lo is not None
                                           else ad.sample_action(rng, n, 0, env),
                                           ad.sample_dis(rng, n, 0, env) if ad.D else None)
                n += 1
                if done:
                    ad.reset(env)
            dt = time.perf_counter() - t0
        out[name] = {"env_steps_per_s": n / dt, "us_per_step": 1e6 * dt / n, "reference": ad.cites}
        print(f"{name:24s} {n / dt:10.0f} env-steps/s  ({1e6 * dt / n:8.1f} us/step)")
    return out


# bench.py workload -> adapter that reproduces the bench's action distribution with the reference's own loop body
WORKER_ADAPTER = {"uav_pos": "uav_pos", "uav_att": "uav_att_rand", "cartpole": "cartpole", "ugvo": "ugvo", "soi": "soi",
                  "fas": "fas_ppo2", "fas_discrete": "fas_discrete", "ballbalancer": "ballbalancer", "twolink": "twolink",
                  "ugv": "ugv_forward", "uavr_hover": "uavr_hover"}


def worker(workload: str, seconds: float, seed: int) -> dict:
    """One process = one host core: the reference's own step loop (e.g. PPO2-4-UavFntsmcParamPos/train.py:290-297:
    get_param_from_actor + generate_action_4_uav + step_update) with uniform random actions and reset(True) on terminal,
    200 warm-up steps, then `seconds` of stepping.  Prints one JSON line."""
    ad = A.REGISTRY[WORKER_ADAPTER[workload]][0]()
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    with R.quiet():
        env = ad.make()
        ad.reset(env)
        act = lambda k: (rng.uniform(ad.action_lo, ad.action_hi) if ad.action_lo is not None
                         else ad.sample_action(rng, k, 0, env))
        for t in range(200 if workload != "ugvo" else 20):
            if ad.step(env, act(t), ad.sample_dis(rng, t, 0, env) if ad.D else None)[3]:
                ad.reset(env)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            done = ad.step(env, act(n), ad.sample_dis(rng, n, 0, env) if ad.D else None)[3]
            n += 1
            if done:
                ad.reset(env)
        dt = time.perf_counter() - t0
    return {"steps": n, "seconds": dt, "reference_tree": R.REF, "cites": ad.cites}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--worker", default=None, help="bench.py workload name: run one timed loop and print JSON")
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    torch.set_num_threads(1)
    if args.worker:
        print(json.dumps(worker(args.worker, args.seconds, args.seed)))
        return
    res = {"python": sys.version.split()[0], "numpy": np.__version__, "torch": torch.__version__,
           "cores_used": 1, "host_cores": os.cpu_count()}
    res["config1"] = config1(args.epochs)
    print("config #1:", json.dumps({k: v for k, v in res["config1"].items() if k != "epochs"}))
    res["per_env"] = per_env(args.seconds)
    path = os.path.join(ROOT, "profiles", "r1", "reference_python_cpu.json")
    json.dump(res, open(path, "w"), indent=1)
    print("->", path)


if __name__ == "__main__":
    main()
