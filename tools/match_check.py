"""Tensor-core matcher (tcgen05) vs the CUDA-core popcount form vs the oracle; timings.  GPU box only."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cv_oracle as O  # noqa: E402 (checker)
from velocity_b200 import match  # noqa: E402

rng = np.random.default_rng(4)
ok = True
for nq, nt in [(128, 256), (200, 300), (1000, 777), (4096, 4096), (8192, 8192)]:
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = t[rng.integers(0, nt, nq)].copy()
    flip = rng.integers(0, 256, (nq, 3))
    for k in range(3):
        q[np.arange(nq), flip[:, k] // 8] ^= (1 << (flip[:, k] % 8)).astype(np.uint8)
    if nt > 20:
        t[11] = t[5]  # exact duplicate rows: lowest index must win
    res = {}
    for mode in ("popc", "tc"):
        os.environ["VEL_MATCH_FORCE"] = mode
        res[mode] = match.knn2_hamming256(q, t)
    oi, od = O.knn2_hamming(q, t)
    e_tc = np.array_equal(res["tc"][0], oi) and np.array_equal(res["tc"][1], od)
    e_pc = np.array_equal(res["popc"][0], oi) and np.array_equal(res["popc"][1], od)
    ok &= e_tc and e_pc
    print("nq=%5d nt=%5d  tcgen05 == oracle: %s   popcount == oracle: %s" % (nq, nt, e_tc, e_pc), flush=True)
    if not e_tc:
        bad = np.nonzero((res["tc"][0] != oi).any(1) | (res["tc"][1] != od).any(1))[0]
        print("   first mismatches:", bad[:5], res["tc"][0][bad[:3]], oi[bad[:3]], res["tc"][1][bad[:3]], od[bad[:3]])

dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
for mode in ("popc", "tc"):
    os.environ["VEL_MATCH_FORCE"] = mode
    for _ in range(3):
        match.knn2_hamming256(dq, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        match.knn2_hamming256(dq, dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-5s 8192x8192x256bit: %.1f us per frame pair -> %.1f T int8-op/s (34.36 Gop)" % (mode, ms * 1e3, 34.36e9 / (ms * 1e-3) / 1e12))
sys.exit(0 if ok else 1)
