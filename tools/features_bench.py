"""Timing of K9 (goodFeaturesToTrack on the GPU) against cv2 on the host: python tools/features_bench.py"""
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, ".")
from velocity_b200 import features, synth  # noqa: E402

for h, w in ((1080, 1920), (2160, 3840)):
    im = synth.texture(h, w, 5)
    d = torch.from_numpy(im).cuda()
    for n, q in ((1000, 0.01), (4096, 0.001)):
        for _ in range(3):
            out, cnt, _ = features.harris_corners_device(d, n, q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out, cnt, _ = features.harris_corners_device(d, n, q)
        e1.record()
        torch.cuda.synchronize()
        gpu_ms = e0.elapsed_time(e1) / 10
        t0 = time.perf_counter()
        for _ in range(3):
            ref = cv2.goodFeaturesToTrack(im, n, q, 0, blockSize=5, useHarrisDetector=True)
        cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
        got = out[: int(cnt.item())].cpu().numpy()
        print("%dx%d n=%d q=%g: GPU %.3f ms (device-resident frame, incl. workspace alloc)  cv2 (%d threads) %.1f ms  equal=%s" % (
            w, h, n, q, gpu_ms, cv2.getNumThreads(), cpu_ms, np.array_equal(got, ref.reshape(-1, 2))))
