"""Issue-rate probe of tcgen05.mma (b200_umma_rate): cycles per 128 x N x 8 TF32 MMA for chains accumulating into 1..4
TMEM tiles in rotation, SS (A from shared memory) and TS (A from TMEM) forms."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from reinforcementlearningplatform_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.b200_umma_rate.restype = C.c_int
lib.b200_umma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
count = 2400
print("cycles per MMA (128 x N x 8, kind::tf32), chain of %d" % count)
for ts in (0, 1):
    for N in (16, 32, 64):
        row = []
        for n_acc in (1, 2, 8):      # 8: two issuing warps, one accumulator each
            for _ in range(2):
                out.zero_()
                _lib.check(lib.b200_umma_rate(N, count, n_acc, ts, C.c_void_p(out.data_ptr()), None), "rate")
                torch.cuda.synchronize()
            row.append(out.max().item() / count)
        print(f"{'TS' if ts else 'SS'} N={N:3d}: " + "  ".join(f"{k}: {v:6.1f}" for k, v in zip(("1 acc", "2 acc", "2 issuers (per issuer)"), row)))
