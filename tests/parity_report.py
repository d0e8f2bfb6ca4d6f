"""Prints (and writes to gpurun_out/parity_<backend>.json) the parity table of every golden fixture:
worst mixed error per output in one-step (re-sync) and free-running mode, flag mismatches, and the
free-running error relative to the reference's own sensitivity.  backend = engine (GPU) | oracle (CPU)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import ENGINE_TOL, EngineBackend, OracleBackend, load_golden, replay  # noqa: E402


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else "engine"  # engine | engine_f32 | oracle
    names = sys.argv[2:] or sorted(ENGINE_TOL)
    f32 = backend == "engine_f32"
    if f32:
        import torch
    rows = {}
    for name in names:
        g = load_golden(name)
        L = g["reward"].shape[1]
        for mode in ("resync", "free"):
            b = (EngineBackend(name, L, dtype=torch.float32) if f32 else EngineBackend(name, L)) \
                if backend.startswith("engine") else OracleBackend(name, L)
            # fp32: short horizon (100 steps), compared without the chaos cut-off so the raw drift is visible
            r = replay(g, b, resync=(mode == "resync"), name=name, steps=100 if f32 else None,
                       chaos_cut=1e300 if f32 else 1e-11)
            rows[f"{name}:{mode}"] = r
            w = r["worst"]
            print(f"{name:26s} {mode:6s} obs {w['obs']:.1e} next {w['next_obs']:.1e} rew {w['reward']:.1e} "
                  f"state {w['state']:.1e} time {w['time']:.0e} flags {r['flag_mismatch']}/{r['done_mismatch']} "
                  f"ratio {r['worst_ratio']:.2e} first_bad {r['first_bad']}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/parity_{backend}.json", "w"), indent=1, default=str)


if __name__ == "__main__":
    main()
