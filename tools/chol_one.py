"""One vel_spd_solve call at C3 size (for ncu)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from velocity_b200 import _lib
from velocity_b200.device import ptr, stream_ptr
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1794
A = torch.randn((n, n + 8), dtype=torch.float64, device="cuda")
S0 = A @ A.T + torch.eye(n, dtype=torch.float64, device="cuda")
b = torch.randn(n, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(3):
    S = S0.clone()
    _lib.check(L.vel_spd_solve(ptr(S), n, n, ptr(b), ptr(info), stream_ptr()), "spd")
torch.cuda.synchronize()
