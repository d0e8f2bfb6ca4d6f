"""Per-field worst one-step error of the fp32 engine against the reference fixtures (diagnostic for the fp32
tolerance table in tests/test_engine_gpu.py).  Usage (GPU box): python tools/f32_field_report.py <fixture> ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import EngineBackend, load_golden  # noqa: E402

for name in sys.argv[1:]:
    g = load_golden(name)
    T, L = g["reward"].shape
    T = min(T, 200)
    b = EngineBackend(name, L, dtype=torch.float32)
    b.set_state(g["state0"], g["time0"])
    ws = None
    wn = None
    for t in range(T):
        if t > 0:
            st = np.where(np.isnan(g["reset_state"][t - 1]), g["state"][t - 1], g["reset_state"][t - 1])
            tm = np.where(np.isnan(g["reset_time"][t - 1]), g["time"][t - 1], g["reset_time"][t - 1])
            b.set_state(st, tm)
        out = b.step(g["actions"][t], g["dis"][t] if "dis" in g else None)
        nf = min(out["state"].shape[1], g["state"][t].shape[1])
        es = np.max(np.abs(out["state"][:, :nf] - g["state"][t][:, :nf]) / np.maximum(1, np.abs(g["state"][t][:, :nf])), axis=0)
        en = np.max(np.abs(out["next_obs"] - g["next_obs"][t]) / np.maximum(1, np.abs(g["next_obs"][t])), axis=0)
        ws = es if ws is None else np.maximum(ws, es)
        wn = en if wn is None else np.maximum(wn, en)
    print(name, "state per field:", " ".join(f"{v:.1e}" for v in ws))
    print(name, "next_obs per field:", " ".join(f"{v:.1e}" for v in wn))
