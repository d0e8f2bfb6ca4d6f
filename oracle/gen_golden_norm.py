"""Golden vectors for the running normalisation and the PPO (v1) return scan, from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (container-side).
  * `utils.classes.Normalization` (utils/classes.py:626-656) is fed sample by sample, exactly as the train loops do
    (`reward_norm(env.reward)`, `env.current_state_norm(env.current_state, update=True)`): inputs, outputs and the
    final (n, mean, S, std), including the first-sample quirk (std = x, output -0.0 for a negative first sample) and an
    `update=False` evaluation pass.
  * `Proximal_Policy_Optimization.learn` (algorithm/policy_base/Proximal_Policy_Optimization.py:107-119): the returns
    it computes, captured from the learner's locals with sys.settrace (K_epochs = 0, no parameter update).

    python oracle/gen_golden_norm.py   ->  tests/golden/norm.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim as R  # noqa: E402


def norm_case(cls, dim, rows, seed, scale, first=None):
    rng = np.random.default_rng(seed)
    x = rng.normal(0.3, 1.0, (rows, dim)) * scale
    if first is not None:
        x[0] = first
    nz = cls.Normalization(shape=dim)
    y = np.zeros_like(x)
    for t in range(rows):
        y[t] = nz(x[t].copy(), update=True) if dim > 1 else nz(float(x[t, 0]) if t % 2 else x[t].copy(), update=True)
    ms = nz.running_ms
    xe = rng.normal(0, 2.0, (16, dim)) * scale
    ye = np.stack([np.asarray(nz(xe[t].copy(), update=False)).reshape(dim) for t in range(16)])
    run = np.stack([np.full(dim, float(ms.n)), np.asarray(ms.mean, float).reshape(dim),
                    np.asarray(ms.S, float).reshape(dim), np.asarray(ms.std, float).reshape(dim)])
    return dict(x=x, y=y, run=run, x_eval=xe, y_eval=ye)


def main():
    R.install()
    with R.quiet():
        cls = R.load("utils.classes")
    out = {"numpy": np.array(np.__version__)}
    cases = [(6, 2000, 1, 1.0, None), (1, 1500, 2, 5.0, None), (6, 64, 3, 1.0, -np.abs(np.arange(6.0) + 0.5)),
             (3, 1, 4, 1.0, np.array([-2.0, 0.0, 3.0])), (41, 300, 5, 0.7, None)]
    for k, (dim, rows, seed, scale, first) in enumerate(cases):
        c = norm_case(cls, dim, rows, seed, scale, first)
        for name, v in c.items():
            out[f"n{k}_{name}"] = v
        print(f"norm case {k}: dim={dim} rows={rows} y[0]={c['y'][0][:3]} run n={c['run'][0, 0]}")
    out["n_norm"] = np.array(len(cases))

    # PPO v1 return scan: replicate the learner's loop through its own code object
    with R.quiet():
        mod = R.load("algorithm.policy_base.Proximal_Policy_Optimization")
    code = mod.Proximal_Policy_Optimization.learn.__code__
    rcases = [(1000, 11, 0.01), (2048, 12, 0.0), (500, 13, 1.0), (37, 14, 0.2), (1, 15, 0.0)]
    for k, (T, seed, p) in enumerate(rcases):
        rng = np.random.default_rng(seed)
        S, A = 4, 2

        class Stub:  # the attributes learn() touches before the scan finishes (Proximal_Policy_Optimization.py:107-119)
            pass
        agent = Stub()
        agent.gamma = 0.99
        agent.buffer = cls.RolloutBuffer(T, S, A)
        agent.buffer.r[:] = rng.normal(-0.5, 2.0, agent.buffer.r.shape)
        done = (rng.random(T) < p).astype(float)
        if T > 2 and p > 0:
            done[-1] = 1.0
        agent.buffer.done[:] = done.reshape(agent.buffer.done.shape)
        cap = {}

        def tracer(frame, event, arg):
            if frame.f_code is not code:
                return None

            def local(frame, event, arg):
                rw = frame.f_locals.get("rewards")
                if torch.is_tensor(rw) and "ret" not in cap:
                    cap["ret"] = rw.clone().numpy().reshape(-1)
                return local
            return local
        sys.settrace(tracer)
        try:
            with R.quiet():
                try:
                    mod.Proximal_Policy_Optimization.learn(agent)
                except Exception:
                    pass  # the stub has no nets: learn() stops after the scan, which is all that is captured
        finally:
            sys.settrace(None)
        assert "ret" in cap and cap["ret"].shape == (T,), (k, cap.keys())
        out[f"r{k}_r"] = agent.buffer.r.reshape(T).copy()
        out[f"r{k}_done"] = done.astype(np.uint8)
        out[f"r{k}_ret"] = cap["ret"]
        print(f"returns case {k}: T={T} p_done={p} ret[:3]={cap['ret'][:3]} dtype={cap['ret'].dtype}")
    out["n_ret"] = np.array(len(rcases))
    out["gamma"] = np.array(0.99)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "norm.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) // 1000, "kB")


if __name__ == "__main__":
    main()
