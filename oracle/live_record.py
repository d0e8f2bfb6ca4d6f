"""Record a WIDE fixture from the reference at test time: many lanes, one process per lane.

TEST INFRASTRUCTURE ONLY.  SURVEY.md 8c-i asks for 64 seeds x 1000 steps per environment; as committed .npz files those
are ~28 MB each, so tests/test_live_reference_gpu.py produces them on the spot from whichever reference tree is present
(/root/reference in the build container, the byte-compiled staging oracle/_ref on the GPU box) with the recorder of
oracle/gen_golden.py, lanes spread over the host cores as plain subprocesses (no CUDA, OMP_NUM_THREADS=1).

    python oracle/live_record.py <adapter> <lane> <steps> <seed> <out.npz>     (worker mode)"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def record_parallel(name: str, lanes: int, steps: int, seed: int, procs: int | None = None, timeout: float = 900) -> dict:
    """Same dict as a committed fixture (oracle/gen_golden.py docstring) for `lanes` independent seeds."""
    procs = procs or min(lanes, os.cpu_count() or 1)
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    parts = [None] * lanes
    with tempfile.TemporaryDirectory() as tmp:
        pending, running = list(range(lanes)), []
        while pending or running:
            while pending and len(running) < procs:
                l = pending.pop(0)
                out = os.path.join(tmp, f"lane{l}.npz")
                p = subprocess.Popen([sys.executable, os.path.abspath(__file__), name, str(l), str(steps), str(seed), out],
                                     env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                running.append((l, out, p))
            l, out, p = running.pop(0)
            _, err = p.communicate(timeout=timeout)
            if p.returncode != 0:
                for _, _, q in running:
                    q.kill()
                raise RuntimeError(f"live_record worker for lane {l} failed:\n{err[-2000:]}")
            with np.load(out) as z:
                parts[l] = {k: z[k] for k in z.files}
    res = {}
    for k in parts[0]:
        if k == "meta":
            res[k] = parts[0][k]
        elif parts[0][k].ndim >= 2 and parts[0][k].shape[0] == steps and parts[0][k].shape[1] == 1:
            res[k] = np.concatenate([p[k] for p in parts], axis=1)      # [T, L, ...]
        else:
            res[k] = np.concatenate([p[k] for p in parts], axis=0)      # [L, ...]
    return res


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from oracle import gen_golden, ref_adapters as A
    name, lane, steps, seed, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    data = gen_golden.record(A.REGISTRY[name][0](), 1, steps, seed, lane_offset=lane + 1)
    np.savez(out, **data)
