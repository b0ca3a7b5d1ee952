"""Ad-hoc timing of K1+K2 on synthetic 1080p pairs (development aid; bench.py is the contract)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from velocity_b200 import synth  # noqa: E402
from velocity_b200.lk import FrameBatch, lk_params, track_pairs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Z0 = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
frames, _ = synth.plane_sequence(B + 1, seed=1234, Z0=Z0)
pts = torch.from_numpy(synth.harris_tracks(frames[0], 4096)).cuda()
dev = torch.from_numpy(np.stack(frames)).cuda()
for name, lk, fbt in [("c2_fwd", dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1)), None),
                      ("c2_fb", dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1)), 1.0),
                      ("coarse5_fb", dict(winSize=(15, 15), maxLevel=4, criteria=(3, 10, 0.1)), 1.0),
                      ("fine_fb", dict(winSize=(51, 51), maxLevel=0, criteria=(3, 30, 0.001)), 0.3)]:
    params = lk_params(fbt=fbt, **lk)
    fb = FrameBatch(dev, lk["winSize"], lk["maxLevel"])
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(3):
        torch.cuda.synchronize()
        e[0].record()
        fb.build()
        e[1].record()
        out, st, err, _ = track_pairs(fb, fb, pts, params, 0, 1, B)
        e[2].record()
        torch.cuda.synchronize()
    print("%-11s B=%d pyramid %.3f ms (%.1f us/frame)  track %.3f ms (%.1f us/pair)  valid %.3f" % (
        name, B, e[0].elapsed_time(e[1]), e[0].elapsed_time(e[1]) * 1e3 / (B + 1), e[1].elapsed_time(e[2]),
        e[1].elapsed_time(e[2]) * 1e3 / B, st.float().mean().item()))
