"""Does the H2D rate of a pinned buffer depend on where the allocation happened to land?  Allocates the C3 frame buffer (622 MB)
several times and times a full upload of each (CUDA events, best of 3):  python tools/pinned_probe.py"""
import torch

n = 300 * 1080 * 1920
dev = torch.device("cuda", 0)
dst = torch.empty(n, dtype=torch.uint8, device=dev)
bufs = []
for i in range(6):
    b = torch.empty(n, dtype=torch.uint8).pin_memory()
    b[::4096] = 1
    bufs.append(b)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(b, non_blocking=True); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("pinned buffer %d: %.2f ms per 622 MB upload = %.1f GB/s" % (i, best, n / best / 1e6))
# and chunked, as the sequence uploads (25 frames per copy)
b = bufs[0]
ch = 25 * 1080 * 1920
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for lo in range(0, n, ch):
        dst[lo:lo + ch].copy_(b[lo:lo + ch], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("chunked (25 frames per copy): %.2f ms = %.1f GB/s" % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6))
