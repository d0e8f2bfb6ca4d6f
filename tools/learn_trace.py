"""Phase breakdown of K-LEARN's gradient kernel (variant built with -DLEARN_TRACE):
    SRC=learn tools/build_variant.sh ltrace -DLEARN_TRACE
    B200ENV_LIB=$PWD/tools/variants/libb200env_ltrace.so MB=262144 python tools/learn_trace.py
Cycles thread 0 of block 0 (an actor block) spent between the phase barriers, summed over its tiles."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench  # noqa: E402
from reinforcementlearningplatform_b200 import _lib  # noqa: E402

lib = _lib.load()
buf = (C.c_longlong * 16)()
microbench.run("learn", 2)
lib.b200_learn_trace_read(buf)                      # discard warm-up
res = microbench.run("learn", 1)
lib.b200_learn_trace_read(buf)
names = {0: "sample indices", 1: "gather", 2: "fwd L0", 3: "fwd L1", 4: "fwd L2", 5: "fwd L3", 6: "loss", 7: "dW L0 (+ tail)",
         9: "dW L1", 10: "dH L1 + dZ", 11: "dW L2", 12: "dH L2 + dZ", 13: "dW L3", 14: "dH L3 + dZ", 15: "top-of-tile barrier"}
tot = sum(buf)
launches = 4                                        # 3 warm-up + 1 timed call of microbench.run(kind, 1)
for k in sorted(names):
    print(f"{names[k]:22s} {buf[k] / launches:12.0f} cycles  {100.0 * buf[k] / max(tot, 1):5.1f} %")
print("total", tot / launches, res["ms_per_step"], "ms per update")
