"""Extract the inputs of the reference's end-to-end run (vidExample.py:19-39,75-91) into
tests/_refdata/<clip>.npz: gray frames, plate corners q, intrinsics K, frame times.

    python tools/make_refdata.py      # authoring container only (needs /root/reference)

tests/_refdata is git-ignored (video-derived data, ~25 MB per clip) but travels to the GPU box with
the repo snapshot; tests/test_e2e_gpu.py skips when it is absent.  The expected outputs are the
committed pins in tests/golden/e2e_vidExample.npz (made by running the reference itself)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import scipy.io  # noqa: E402

from oracle import ref_shim  # noqa: E402

ns = ref_shim.load()
out_dir = os.path.join(ROOT, "tests", "_refdata")
os.makedirs(out_dir, exist_ok=True)
cwd = os.getcwd()
os.chdir(ref_shim.REF_ROOT)
try:
    for clip, start in [("IMG_4134.MOV", 19), ("IMG_4119.MOV", 41)]:
        n = 20
        cam, cap = ns.images.getCameraParams("./data/" + clip, platform="iPhone 6s")
        mat = scipy.io.loadmat("./matlab/" + cam["filename"] + ".mat")
        q = mat["q"].astype(np.float32)
        q /= 2                                   # vidExample.py:36-39 (4K -> 1080p)
        cam["IntrinsicMatrix"][:2, :2] /= 2
        K = cam["IntrinsicMatrix"]
        frames, times = [], []
        cap.set(1, start)
        for i in range(n):
            times.append(cap.get(cv2.CAP_PROP_POS_MSEC) / 1000)
            ok, bgr = cap.read()
            assert ok
            frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
        cap.release()
        path = os.path.join(out_dir, clip.split(".")[0] + ".npz")
        np.savez_compressed(path, frames=np.stack(frames), q=q, K=np.asarray(K, float), times=np.array(times, np.float32))
        print(path, os.path.getsize(path) / 1e6, "MB", "K=", K.tolist())
finally:
    os.chdir(cwd)
