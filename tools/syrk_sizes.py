import os, sys
import torch
sys.path.insert(0, ".")
from velocity_b200 import _lib
from velocity_b200.device import ptr, stream_ptr
L = _lib.lib()
for m, k in [(1794, 12288), (1800, 12288), (1920, 12288), (2048, 12288), (1024, 12288), (4096, 4096)]:
    E = torch.randn((m, k), dtype=torch.float64, device="cuda")
    S = torch.zeros((m, m), dtype=torch.float64, device="cuda")
    work = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    best = 1e9
    for _ in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.vel_syrk_lower_sub(ptr(E), k, m, k, ptr(S), m, ptr(work), work.numel(), stream_ptr()), "syrk")
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    nb = (m + 127) // 128; bme = ((m + nb - 1) // nb + 7) // 8 * 8
    print("m=%d (nb=%d, bme=%d): %.3f ms  %.1f TFLOP/s useful, %.1f TFLOP/s executed tiles" % (m, nb, bme, best, m * (m + 1.0) * k / best / 1e9,
          nb * (nb + 1) / 2 * bme * bme * 2.0 * k / best / 1e9))
