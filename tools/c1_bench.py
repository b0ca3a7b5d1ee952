"""BASELINE configs[0] (C1): the reference's own clips (20 frames each, decoded frames in tests/_refdata) through
velocity_b200.pipeline.run_speed_estimation -- frames per second of the per-frame loop (tracker + pose + bookkeeping; decode
excluded: the GPU box has no encoded stream), against the same loop driven by the CPU oracle with cv2 as its arithmetic provider,
and the speed both recover.  python tools/c1_bench.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from velocity_b200 import pipeline  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ("IMG_4134", "IMG_4119"):
    d = np.load(os.path.join(ROOT, "tests", "_refdata", name + ".npz"))
    frames, q, K, times = d["frames"], d["q"], d["K"], d["times"]
    for rep in range(2):                      # second pass: warm (library loaded, scratch allocated)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = pipeline.run_speed_estimation(list(frames), q, K, times, verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    per = out["S"][1:, 1]
    print("%s: %d frames, %d..%d tracks: %.1f ms total, per-frame loop %.2f ms median (%.0f fps), speed %.2f +/- %.2f km/h" % (
        name, len(frames), out["tracks"].max(), out["tracks"].min(), dt * 1e3, float(np.median(per)) * 1e3, 1.0 / float(np.median(per)),
        out["speed_mean"], out["speed_std"]))
