"""Phase breakdown of chol_solve_kernel (debug build with -DVEL_CHOL_TIMING, see csrc/dense_f64.cu)."""
import ctypes as C, sys
import torch
L = C.CDLL("tools/micro/libchol_timing.so")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1794
A = torch.randn((n, n + 8), dtype=torch.float64, device="cuda")
S0 = A @ A.T + torch.eye(n, dtype=torch.float64, device="cuda")
b = torch.randn(n, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
L.vel_spd_solve.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
out = (C.c_ulonglong * 28)()
for rep in range(3):
    S = S0.clone()
    L.vel_chol_timing(out, 1)
    rc = L.vel_spd_solve(S.data_ptr(), n, n, b.data_ptr(), info.data_ptr(), None)
    L.vel_chol_timing(out, 0)
names = ["A load diag", "A chol_block", "B row solves", "sync1", "C trailing", "sync2", "back step", "back sync"]
tot = max(sum(out[:8]), 1)
for k, nm in enumerate(names):
    print("%-14s %8.1f us  %5.1f%%" % (nm, out[k] / 1e3, 100.0 * out[k] / tot))
print("total %.1f us (CTA 0), rc %d" % (tot / 1e3, rc))
print("chol_block clocks (both fwd+back phases call it once per panel): (i) 8x8 warp %d  (ii) sub-panel %d  (iii) trailing %d  loop %d  -> per 64x64 block (i)=%.0f (ii)=%.0f (iii)=%.0f clk" % (out[8], out[9], out[10], out[11], out[8]/29, out[9]/29, out[10]/29))

names2 = ["D load", "D chol_block", "D tri_inverse", "D store+signal", "flag waits", "T tasks", "U tasks", "other"]
for role, lab in ((0, "panel CTAs (sum over %d CTAs)"), (1, "worker CTAs (sum)")):
    print(lab % 29 if role == 0 else lab)
    for k, nm in enumerate(names2):
        print("   %-16s %9.1f us" % (nm, out[12 + 8 * role + k] / 1e3))
