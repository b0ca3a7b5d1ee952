"""One launch of the 51x51 lk_fine configuration (development aid for ncu): python tools/lk_fine_one.py [pairs]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from velocity_b200 import synth  # noqa: E402
from velocity_b200.lk import FrameBatch, lk_params, track_pairs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames, _ = synth.plane_sequence(B + 1, seed=1234, Z0=40.0)
pts = torch.from_numpy(synth.harris_tracks(frames[0], 4096)).cuda()
dev = torch.from_numpy(np.stack(frames)).cuda()
lk = dict(winSize=(51, 51), maxLevel=0, criteria=(3, 30, 0.001))
params = lk_params(fbt=0.3, **lk)
fb = FrameBatch(dev, lk["winSize"], lk["maxLevel"]).build()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(3):
    torch.cuda.synchronize()
    e[0].record()
    out, st, err, _ = track_pairs(fb, fb, pts, params, 0, 1, B)
    e[1].record()
    torch.cuda.synchronize()
print("B=%d track %.3f ms (%.1f us/pair) valid %.4f" % (B, e[0].elapsed_time(e[1]), e[0].elapsed_time(e[1]) * 1e3 / B, st.float().mean().item()))
