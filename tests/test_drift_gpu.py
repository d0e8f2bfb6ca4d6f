"""How far does the shipped engine drift from the reference over whole episodes, and why?

One-step (re-sync) parity is 1e-12 everywhere (test_engine_gpu.py).  Free-running, every difference in the last bit is
amplified by the closed loop (controller gains redrawn every step): the reference drifts from ITSELF by `twin` when nudged
one ulp per step, the C oracle -- reference operation order, but glibc instead of numpy's SIMD pow / tanh -- by more.  This
test puts four numbers side by side, per lane, on the committed UAV fixtures and on 64 seeds x 1000 steps recorded from the
reference at test time: twin drift, C oracle, the strict build of the engine (`make strict`: pow(), tan(), IEEE divisions,
LU-pivot inverse, libdevice, no FMA contraction) and the shipped engine (DESIGN.md section 3 lists its shortcuts).

Measured on B200 (profiles/r2/drift.md): no single shortcut explains the drift -- reverting any ONE of them moves the
worst lane of `uav_pos` between 8e-9 and 4e-8 in either direction (shipped 4e-8, oracle 5e-9, strict 7e-10), i.e. the
number is the amplification (~1e7 on lanes whose gains are redrawn every step) of which last bits happened to flip.  The
bound is therefore statistical: over 64 seeds the shipped engine's typical (median) lane drifts no more than 3x the
oracle's, its worst lane no more than 10x the oracle's worst, and absolute ceilings hold per fixture."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
STRICT = os.path.join(ROOT, "reinforcementlearningplatform_b200", "libb200env_strict.so")
FIXTURES = ["uav_pos", "uav_pos_wide", "uav_pos_dis", "uav_att", "uav_att_rand", "uavr_hover"]
# absolute ceilings on the free-running state error (mixed metric), about 10x the values measured on B200
CEILING = {"uav_pos": 5e-7, "uav_pos_wide": 5e-10, "uav_pos_dis": 2e-9, "uav_att": 1e-12, "uav_att_rand": 1e-12,
           "uavr_hover": 1e-11, "live_uav_pos_dis": 5e-7}


def report(backend, specs, lib=None):
    env = dict(os.environ)
    if lib:
        env["B200ENV_LIB"] = lib
    out = subprocess.run([sys.executable, os.path.join(HERE, "drift_report.py"), backend] + specs, env=env,
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_free_running_drift_is_bounded_and_decomposed():
    assert os.path.exists(STRICT), "libb200env_strict.so missing: run `make -C reinforcementlearningplatform_b200/csrc strict`"
    from oracle import ref_shim
    specs = list(FIXTURES)
    with tempfile.TemporaryDirectory() as tmp:
        if ref_shim.available():       # 64 seeds x 1000 steps from the reference itself (not committed: 28 MB)
            from oracle.live_record import record_parallel
            g = record_parallel("uav_pos_dis", lanes=64, steps=1000, seed=123)
            path = os.path.join(tmp, "live_uav_pos_dis.npz")
            np.savez(path, **g)
            specs.append(path + ":uav_pos_dis")
        oracle, strict, shipped = report("oracle", specs), report("engine", specs, STRICT), report("engine", specs)
    rows = {}
    for name in shipped:
        med = lambda d: float(np.median(d[name]["lane_state"]))
        rows[name] = {"lanes": shipped[name]["lanes"], "steps": shipped[name]["steps"],
                      "twin_max": float(np.max(shipped[name]["lane_twin"])), "oracle_max": oracle[name]["state"],
                      "strict_max": strict[name]["state"], "shipped_max": shipped[name]["state"],
                      "oracle_median": med(oracle), "strict_median": med(strict), "shipped_median": med(shipped)}
        r = rows[name]
        print(f"{name:18s} {r['lanes']:3d} lanes  twin {r['twin_max']:.1e} | max: oracle {r['oracle_max']:.1e} strict "
              f"{r['strict_max']:.1e} shipped {r['shipped_max']:.1e} | median lane: oracle {r['oracle_median']:.1e} strict "
              f"{r['strict_median']:.1e} shipped {r['shipped_median']:.1e}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "drift.json"), "w"), indent=1)
    for name, r in rows.items():
        assert shipped[name]["flag_mismatch"] == 0 and strict[name]["flag_mismatch"] == 0, name
        assert r["shipped_max"] <= CEILING[name], (name, r)
        if r["lanes"] >= 16:           # enough seeds for a distribution
            assert r["shipped_median"] <= max(3.0 * r["oracle_median"], 1e-13), (name, r)
            assert r["shipped_max"] <= max(10.0 * r["oracle_max"], 1e-13), (name, r)
