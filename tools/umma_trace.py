"""Event trace of the tcgen05 policy kernel (variant built with -DB200_UMMA_TRACE, see tools/build_variant.sh):
    SRC=policy_umma tools/build_variant.sh trace -DB200_UMMA_TRACE
    B200ENV_LIB=$PWD/tools/variants/libb200env_trace.so python tools/umma_trace.py [out.json]
Prints, for CTA 0, the cycle stamps of the first tiles: epilogue groups (wait D / got D / chunk stored / arrived /
sampled) and the MMA issue lane (wait A / got A / layer committed)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench  # noqa: E402
from reinforcementlearningplatform_b200 import _lib  # noqa: E402

lib = _lib.load()
kind = os.environ.get("KIND", "policy")
microbench.run(kind, 1)
buf = (C.c_longlong * (4 * 2048))()
cnt = (C.c_int * 4)()
lib.b200_umma_trace_read(buf, cnt)          # discard the warm-up launches
microbench_res = microbench.run(kind, 1)
lib.b200_umma_trace_read(buf, cnt)
ev = []
for role in range(4):
    for k in range(cnt[role]):
        ev.append((buf[role * 2048 + 2 * k + 1], role, buf[role * 2048 + 2 * k]))
ev.sort()
t0 = ev[0][0]
names = {1: "EG wait D", 2: "EG got D", 3: "EG chunk stored", 4: "EG arrived", 5: "EG sampled", 6: "  chunk: ld issued",
         7: "  chunk: dst free", 8: "  chunk: ld done", 9: "  chunk: computed"}
out = []
for t, role, eid in ev[:400]:
    if role < 2:
        what = f"{names[eid // 100]} L{eid % 100}"
    else:
        what = {1: "MMA wait A", 2: "MMA got A", 3: "MMA committed"}[eid // 1000] + f" L{eid % 100}"
    out.append((t - t0, ["EG0", "EG1", "MMA0", "MMA1"][role], what))
    print(f"{t - t0:8d}  {['EG0', 'EG1', 'MMA0', 'MMA1'][role]:4s} {what}")
if len(sys.argv) > 1:
    json.dump({"events": out, "bench": microbench_res}, open(sys.argv[1], "w"))
