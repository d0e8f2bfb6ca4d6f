import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from helpers import EngineBackend, load_golden
name = "uavr_hover"
g = load_golden(name)
T, L = g["reward"].shape
b = EngineBackend(name, L, dtype=torch.float32)
b.set_state(g["state0"], g["time0"])
cnt = 0
for t in range(200):
    if t > 0:
        st = np.where(np.isnan(g["reset_state"][t - 1]), g["state"][t - 1], g["reset_state"][t - 1])
        tm = np.where(np.isnan(g["reset_time"][t - 1]), g["time"][t - 1], g["reset_time"][t - 1])
        b.set_state(st, tm)
        prev = st
    else:
        prev = g["state0"]
    out = b.step(g["actions"][t], g["dis"][t] if "dis" in g else None)
    e = np.abs(out["next_obs"] - g["next_obs"][t])
    bad = np.argwhere(e > 0.1)
    for (l, f) in bad[:2]:
        if cnt < 6:
            cnt += 1
            print("t", t, "lane", l, "field", f, "got", out["next_obs"][l, f], "ref", g["next_obs"][t][l, f])
            print("  prev state 15..20", prev[l, 15:21], "action", g["actions"][t][l])
            print("  new state got 15..20", out["state"][l, 15:21], "ref", g["state"][t][l, 15:21])
            print("  x got 6..11", out["state"][l, 6:12], "ref", g["state"][t][l, 6:12])
