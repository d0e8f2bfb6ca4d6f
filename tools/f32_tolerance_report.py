"""Measures what tests/test_engine_gpu.py::test_engine_fp32_stated_tolerance asserts: quantiles of the fp32 engine's
mixed error on (next_obs, reward) against the fp64 reference fixtures, one-step and 100-step free-running.
Usage (GPU box): python tools/f32_tolerance_report.py [fixture ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import FP32_TOL, EngineBackend, fp32_replay, load_golden  # noqa: E402

for name in (sys.argv[1:] or sorted(FP32_TOL)):
    g = load_golden(name)
    L = g["reward"].shape[1]
    row = [name]
    for resync, steps in ((True, 200), (False, 100)):
        errs, fm = fp32_replay(g, EngineBackend(name, L, dtype=torch.float32), steps=steps, resync=resync)
        tol = FP32_TOL[name][0 if resync else 1]
        row.append(f"{'one' if resync else 'free'}: med {np.median(errs):.1e} q99 {np.quantile(errs, 0.99):.1e} "
                   f"q999 {np.quantile(errs, 0.999):.1e} max {errs.max():.1e} flags {fm}/{errs.size} tol {tol}")
    print(" | ".join(row))
