"""Generate the golden trajectory fixtures under tests/golden/ from the LIVE reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py [name ...]

The reference pins nothing itself (no tests, no golden vectors -- SURVEY.md section 4),
so these fixtures are the pin: outputs of the unmodified reference classes, imported
through oracle/ref_shim.py, driven with recorded random actions.  Each .npz holds, for
L independent lanes and T steps,

    state0 [L,F], time0 [L]          internal state after the initial reset
    actions [T,L,A] (, dis [T,L,D])  inputs
    next_obs [T,L,S], reward [T,L], done [T,L], flag [T,L]   outputs of step_update
    state [T,L,F], time [T,L]        internal state after the step (before any reset)
    reset_state [T,L,F]              state after the reference's own reset(True) that follows
                                     a terminal step (NaN rows elsewhere), reset_time likewise
    twin_err [T,L]                   self-sensitivity of the reference: mixed error of next_obs between the recorded run and
                                     a twin reference instance whose state is nudged by about one ulp after the reset and after
                                     every step (what any re-implementation with different rounding does).  Free-running
                                     tolerances are expressed as a multiple of its running maximum.
    meta                             numpy version, cpu flags, reference call sites

The checkers (C restatement in oracle/, CUDA engine) start from state0, apply the same
actions, compare every output each step and re-inject reset_state after terminal steps.
"""
from __future__ import annotations

import json
import os
import platform
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim as R  # noqa: E402
from oracle import ref_adapters as A  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def record(adapter: "A.Adapter", lanes: int, steps: int, seed: int, lane_offset: int = 0) -> dict:
    """lane_offset > 0: record lanes lane_offset .. lane_offset + lanes - 1 of a wider run with a generator of their own
    (oracle/live_record.py records one lane per process); 0 keeps the committed fixtures reproducible bit for bit."""
    rng = np.random.default_rng([seed, lane_offset]) if lane_offset else np.random.default_rng(seed)
    F, S, Adim, D = adapter.F, adapter.S, adapter.A, adapter.D
    out = {
        "state0": np.zeros((lanes, F)), "time0": np.zeros(lanes),
        "actions": np.zeros((steps, lanes, Adim)),
        "obs": np.zeros((steps, lanes, S)),
        "next_obs": np.zeros((steps, lanes, S)), "reward": np.zeros((steps, lanes)),
        "done": np.zeros((steps, lanes), np.uint8), "flag": np.zeros((steps, lanes), np.int32),
        "state": np.zeros((steps, lanes, F)), "time": np.zeros((steps, lanes)),
        "reset_state": np.full((steps, lanes, F), np.nan), "reset_time": np.full((steps, lanes), np.nan),
        "twin_err": np.zeros((steps, lanes)),
    }
    if D:
        out["dis"] = np.zeros((steps, lanes, D))
    for l in range(lanes):
        gl = l + lane_offset
        np.random.seed(seed * 1000 + gl)  # the reference's resets use the global numpy RNG
        with R.quiet():
            env = adapter.make()
            twin = adapter.make()
            rs = np.random.get_state()
            adapter.reset(env)
            np.random.set_state(rs)
            adapter.reset(twin)
            adapter.perturb(twin, 0)
        out["state0"][l], out["time0"][l] = adapter.internal(env)
        for t in range(steps):
            a = adapter.sample_action(rng, t, gl, env)
            d = adapter.sample_dis(rng, t, gl, env) if D else None
            with R.quiet():
                o, o2, r, done, flag = adapter.step(env, a, d)
                _, t2, _, tdone, _ = adapter.step(twin, a, d)
            out["twin_err"][t, l] = float(np.max(np.abs(t2 - o2) / np.maximum(1.0, np.abs(o2))))
            adapter.perturb(twin, t + 1)
            out["actions"][t, l] = a
            if D:
                out["dis"][t, l] = d
            out["obs"][t, l] = o
            out["next_obs"][t, l] = o2
            out["reward"][t, l] = r
            out["done"][t, l] = 1 if done else 0
            out["flag"][t, l] = flag
            out["state"][t, l], out["time"][t, l] = adapter.internal(env)
            if done:
                with R.quiet():
                    rs = np.random.get_state()
                    adapter.reset(env)
                    np.random.set_state(rs)
                    adapter.reset(twin)
                    adapter.perturb(twin, 0)
                out["reset_state"][t, l], out["reset_time"][t, l] = adapter.internal(env)
    meta = {
        "adapter": adapter.name, "reference": adapter.cites, "lanes": lanes, "steps": steps, "seed": seed,
        "numpy": np.__version__, "python": platform.python_version(), "machine": platform.machine(),
        "params": adapter.params_json(),
    }
    try:
        import cpuinfo
        meta["cpu"] = cpuinfo.get_cpu_info().get("brand_raw", "")
    except Exception:
        pass
    out["meta"] = np.array(json.dumps(meta))
    return out


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    names = argv or list(A.REGISTRY)
    for name in names:
        factory, lanes, steps, seed = A.REGISTRY[name]
        adapter = factory()
        data = record(adapter, lanes, steps, seed)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **data)
        nd = int(data["done"].sum())
        print(f"{name}: lanes={lanes} steps={steps} terminals={nd} -> {path} ({os.path.getsize(path)/1e3:.0f} kB)")


if __name__ == "__main__":
    main(sys.argv[1:])
