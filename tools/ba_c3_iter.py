"""Three LM iterations of the bundle adjustment at C3 size (nt=4096, nc=299) -- the launch list of one iteration is read
from `ncu --metrics gpu__time_duration.sum` over this script (tools: see profiles/README.md)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from velocity_b200 import NLS, synth

K = synth.K_1080P
NT, F = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 300)
pw = synth.scene_points(NT, seed=7)
P, cw = synth.scene_observations(pw, F, step=0.02, noise=0.1, seed=11)
z = np.concatenate((P[0].T.ravel(), P[1].T.ravel())).astype(np.float64)
x = np.concatenate((pw + 0.01, cw[1:], np.zeros((F - 1, 3)))).ravel()
ba = NLS.BundleAdjuster(K, z, x, NT, F - 1)
for it in range(4):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); ba.accumulate(); e[1].record(); ba.solve(); e[2].record()
    torch.cuda.synchronize()
    print("it %d: accumulate %.3f ms, solve %.3f ms, rms_delta %.3e" % (it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), ba.rms_delta.item()))

# the device-resident loop (vel_ba_iterate): 10 iterations enqueued at once, no host synchronisation in between (tol = 0: all run)
lb = NLS.BundleAdjuster(K, z, x, NT, F - 1)
lb.iterate(2, 0.0)
for rep in range(3):
    lb.reset(z, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); h = lb.iterate(10, 0.0); e1.record()
    torch.cuda.synchronize()
    print("device loop: %d iterations, %.3f ms per iteration (events), %.3f ms wall; rms_delta %s" % (
        len(h), e0.elapsed_time(e1) / len(h), (time.perf_counter() - t0) * 1e3 / len(h), " ".join("%.2e" % r for _, r in h[:5])))
print("max |x_loop - x_stepwise| after 4 iterations:", end=" ")
lb.reset(z, x); lb.iterate(4, 0.0)
print(float((lb.x - ba.x).abs().max()))
