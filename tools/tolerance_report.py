"""Measured values behind the fp64 tolerances of tests/helpers.py / tests/test_engine_gpu.py (run on a B200):
engine vs C oracle on seeded random inputs (the three configurations the tests use), worst mixed error per fixture name.
The fixture replays (one-step / free-running) are printed by tests/parity_report.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import ENGINE_TOL, engine_vs_oracle  # noqa: E402

rows = {}
for name in sorted(ENGINE_TOL):
    a = engine_vs_oracle(name, n=8192, steps=60, seed=7)
    b = engine_vs_oracle(name, n=1000, steps=40, seed=11, offset=(1 << 33) + 12345)
    c = engine_vs_oracle(name, n=4096, steps=40, seed=5, io_dtype=torch.float32)
    rows[name] = {"random": a["worst"], "offset": b["worst"], "io32_state": c["worst"], "terminals": a["terminals"],
                  "flag_mismatch": a["flag_mismatch"] + b["flag_mismatch"] + c["flag_mismatch"]}
    print(f"{name:26s} random {a['worst']:.1e}  offset {b['worst']:.1e}  io32 {c['worst']:.1e}  terminals {a['terminals']}  "
          f"flags {rows[name]['flag_mismatch']}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tolerance_report.json"), "w"), indent=1)
