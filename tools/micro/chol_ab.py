"""Same-box A/B of two builds of csrc/dense_f64.cu (tools/micro/libdense_old.so / libdense_new.so): vel_spd_solve at n = 1794."""
import ctypes as C, sys
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1794
A = torch.randn((n, n + 8), dtype=torch.float64, device="cuda")
S0 = A @ A.T + torch.eye(n, dtype=torch.float64, device="cuda")
b0 = torch.randn(n, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
libs = {v: C.CDLL("tools/micro/libdense_%s.so" % v) for v in ("old", "new", "C", "D", "E")}
for L in libs.values():
    L.vel_spd_solve.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
for rnd in range(3):
    for v, L in libs.items():
        best = 1e9
        for rep in range(8):
            S, b = S0.clone(), b0.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.vel_spd_solve(S.data_ptr(), n, n, b.data_ptr(), info.data_ptr(), None)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("round %d %s: %.4f ms (rc %d, info %d)" % (rnd, v, best, rc, info.item()))
