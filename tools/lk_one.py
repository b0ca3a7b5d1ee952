"""One K1+K2 launch of the C2 workload (development aid for ncu): python tools/lk_one.py [pairs] [reps]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from velocity_b200 import synth  # noqa: E402
from velocity_b200.lk import FrameBatch, lk_params, track_pairs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames, _ = synth.plane_sequence(min(B + 1, 24), seed=1234, Z0=40.0)
idx = [i % len(frames) if (i // len(frames)) % 2 == 0 else len(frames) - 1 - i % len(frames) for i in range(B + 1)]
pts = torch.from_numpy(synth.harris_tracks(frames[0], 4096)).cuda()
dev = torch.from_numpy(np.stack([frames[i] for i in idx])).cuda()
lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
params = lk_params(fbt=1.0, **lk)
fb = FrameBatch(dev, lk["winSize"], lk["maxLevel"])
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(reps):
    fb.build()
    torch.cuda.synchronize()
    e[0].record()
    out, st, err, _ = track_pairs(fb, fb, pts, params, 0, 1, B)
    e[1].record()
    torch.cuda.synchronize()
print("B=%d track %.3f ms (%.2f us/pair) valid %.4f" % (B, e[0].elapsed_time(e[1]), e[0].elapsed_time(e[1]) * 1e3 / B, st.float().mean().item()))
