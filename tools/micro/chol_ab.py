"""Same-box A/B of several builds of csrc/dense_f64.cu: vel_spd_solve at n = 1794, best of 8, three rounds.
Build each variant as tools/micro/libdense_<name>.so  (nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
--expt-relaxed-constexpr -shared -o tools/micro/libdense_<name>.so <variant of dense_f64.cu> velocity_b200/csrc/api.cu) and list the names below;
boxes differ by 1-3 %, so only numbers taken in one call compare (profiles/README.md has the results of round 2)."""
import ctypes as C, sys
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1794
A = torch.randn((n, n + 8), dtype=torch.float64, device="cuda")
S0 = A @ A.T + torch.eye(n, dtype=torch.float64, device="cuda")
b0 = torch.randn(n, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
libs = {v: C.CDLL("tools/micro/libdense_%s.so" % v) for v in (sys.argv[2:] or ["base", "G"])}
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
