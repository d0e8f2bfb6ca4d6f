"""One b200env_rollout launch (profiling helper): python tools/rollout_once.py <workload> <envs> <steps>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

w, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
r, g = bench.rollout_pipeline(w, n, T, "cuda", False)
print(w, n, T, "ms_rollout", r, "env-steps/s %.3e" % (T * n / (r * 1e-3)))
