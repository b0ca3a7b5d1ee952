"""One vel_syrk_lower_sub call at C3 size (for ncu)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from velocity_b200 import _lib
from velocity_b200.device import ptr, stream_ptr
L = _lib.lib()
m, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1794, 12288)
ld = (k + 31) // 32 * 32
E = torch.randn((m, ld), dtype=torch.float64, device="cuda")
S = torch.zeros((m, m), dtype=torch.float64, device="cuda")
work = torch.empty(max(L.vel_syrk_lower_sub_workspace(m, k), 16), dtype=torch.uint8, device="cuda")
for _ in range(3):
    _lib.check(L.vel_syrk_lower_sub(ptr(E), ld, m, k, ptr(S), m, ptr(work), work.numel(), stream_ptr()), "syrk")
torch.cuda.synchronize()
