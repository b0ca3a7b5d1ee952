"""SYRK timing at C3 size for a few split-K settings (VEL_SYRK_SK)."""
import os, sys
import torch
sys.path.insert(0, ".")
from velocity_b200 import _lib
from velocity_b200.device import ptr, stream_ptr
L = _lib.lib()
m, k = 1794, 12288
ld = k
E = torch.randn((m, ld), dtype=torch.float64, device="cuda")
S = torch.zeros((m, m), dtype=torch.float64, device="cuda")
work = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
for sk in sys.argv[1:]:
    os.environ["VEL_SYRK_SK"] = sk
    best = 1e9
    for _ in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.vel_syrk_lower_sub(ptr(E), ld, m, k, ptr(S), m, ptr(work), work.numel(), stream_ptr()), "syrk")
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("SK=%s: %.3f ms  %.1f TFLOP/s" % (sk, best, m * (m + 1.0) * k / best / 1e9))
