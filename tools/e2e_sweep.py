"""A/B of the host-buffer (e2e) path of bench.py: shard counts.  python tools/e2e_sweep.py [steps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n = bench.WORKLOADS["uav_pos"]["n"]
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for rep in range(2):
    for chunks in (1, 2, 3, 4, 8):
        nn = (n // chunks) * chunks
        ms, en, h2d, d2h = bench.timed_e2e("uav_pos", nn, steps, 3, False, 11, dev, 0, torch.float64, chunks=chunks,
                                           io_dtype=torch.float32)
        print(f"chunks={chunks}: {en * steps / (ms * 1e-3):.4e} env-steps/s, {ms / steps:.4f} ms/step", flush=True)
    p = bench.copy_probe(n, 8, 6, steps, dev, False)
    print(f"copy probe: {n * steps / (p * 1e-3):.4e}, {p / steps:.4f} ms/step", flush=True)
