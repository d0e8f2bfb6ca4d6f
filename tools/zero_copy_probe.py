"""Experiment: the host-buffer step with MAPPED pinned memory instead of copies -- the step kernel reads the float32
actions straight from pinned host memory and writes policy_state / reward / is_terminal straight into pinned host memory
(unified virtual addressing: a cudaHostAlloc'ed buffer is addressable from the device by its host pointer), so that one
launch moves the bytes of a step over PCIe in both directions while it computes.  python tools/zero_copy_probe.py [steps]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
n = bench.WORKLOADS["uav_pos"]["n"]


def run(chunks, zc_in, zc_out):
    nc = n // chunks
    envs = [bench.make_env("uav_pos", nc, dev, c * nc, torch.float64, io_dtype=torch.float32) for c in range(chunks)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(chunks)]
    A, S = envs[0].action_dim, envs[0].state_dim
    rng = np.random.default_rng(5)
    ar = np.asarray(envs[0].action_range, dtype=np.float64)
    host_a = [[torch.from_numpy(rng.uniform(ar[:, :1], ar[:, 1:], size=(A, nc))).float().pin_memory() for _ in range(2)]
              for _ in range(chunks)]
    dev_a = [torch.empty((A, nc), dtype=torch.float32, device=dev) for _ in range(chunks)]
    h_obs = [torch.zeros((S, nc), dtype=torch.float32).pin_memory() for _ in range(chunks)]
    h_rew = [torch.zeros((nc,), dtype=torch.float32).pin_memory() for _ in range(chunks)]
    h_done = [torch.zeros((nc,), dtype=torch.uint8).pin_memory() for _ in range(chunks)]
    for c, e in enumerate(envs):
        e.reset(True)
        if zc_out:  # the kernel's policy-facing outputs land in pinned host memory
            e._reset_obs, e._reward, e._done = h_obs[c], h_rew[c], h_done[c]
            e._hot_io = None
    ready = [torch.cuda.Event() for _ in range(chunks)]
    torch.cuda.synchronize()

    def one(k):
        for c in range(chunks):
            if k > 0:
                ready[c].synchronize()
            with torch.cuda.stream(streams[c]):
                if zc_in:
                    envs[c].step_soa(host_a[c][k & 1])
                else:
                    dev_a[c].copy_(host_a[c][k & 1], non_blocking=True)
                    envs[c].step_soa(dev_a[c])
                if not zc_out:
                    h_obs[c].copy_(envs[c]._reset_obs, non_blocking=True)
                    h_rew[c].copy_(envs[c]._reward, non_blocking=True)
                    h_done[c].copy_(envs[c]._done, non_blocking=True)
                ready[c].record(streams[c])
    for k in range(5):
        one(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        one(5 + k)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    chk = float(h_obs[0].double().abs().sum()) + float(h_rew[0].double().sum())
    return chunks * nc * steps / (ms * 1e-3), ms / steps, chk


for rep in range(2):
    for chunks, zi, zo in ((2, False, False), (1, True, True), (2, True, True), (4, True, True), (2, True, False), (2, False, True)):
        v, ms, chk = run(chunks, zi, zo)
        print(f"chunks={chunks} zero-copy in={zi} out={zo}: {v:.4e} env-steps/s, {ms:.4f} ms/step (checksum {chk:.6e})", flush=True)
