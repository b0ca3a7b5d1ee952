import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from velocity_b200 import NLS, synth
K = synth.K_1080P
for NT, F in [(512, 20), (4096, 100), (4096, 300)]:
    pw = synth.scene_points(NT, seed=7)
    P, cw = synth.scene_observations(pw, F, step=0.02, noise=0.1, seed=11)
    z = np.concatenate((P[0].T.ravel(), P[1].T.ravel())).astype(np.float64)
    x = np.concatenate((pw + 0.01, cw[1:], np.zeros((F - 1, 3)))).ravel()
    t = time.perf_counter(); ba = NLS.BundleAdjuster(K, z, x, NT, F - 1); torch.cuda.synchronize(); print("init", time.perf_counter() - t)
    for it in range(3):
        t0 = time.perf_counter(); ba.accumulate(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ba.solve(); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
        print(NT, F, "accumulate host %.4f  sync %.4f | solve host %.4f sync %.4f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3))
