"""BASELINE config 5 shape: frames shard across ranks (300 frames per GPU), every rank runs KLT on its
own 4K shard (no collective), then ONE global bundle adjustment over all cameras with the per-shard
JtJ/Jtr blocks all-gathered over NCCL (velocity_b200.ba_exchange) and solved redundantly on every rank.

    torchrun --nproc-per-node N tools/c5_sharded.py [--frames-per-gpu 300] [--width 3840 --height 2160]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from velocity_b200 import NLS, synth  # noqa: E402
from velocity_b200.lk import FrameBatch, lk_params, track_pairs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames-per-gpu", type=int, default=300)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--tracks", type=int, default=4096)
ap.add_argument("--ba-iters", type=int, default=10)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
FPG, NT, W, H = args.frames_per_gpu, args.tracks, args.width, args.height
scale = W / 1920.0
K = synth.K_1080P.copy()
K[:2, :2] *= scale
K[2, :2] = [W / 2 + 0.5, H / 2 + 0.5]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


# ---- stage 1: KLT on this rank's shard (one halo frame: pair (FPG-1, FPG) needs frame FPG) --------------
uniq, _ = synth.plane_sequence(12, h=H, w=W, seed=2025 + rank, Z0=40.0, K_row=K)
idx = [i % 12 if (i // 12) % 2 == 0 else 11 - (i % 12) for i in range(FPG + 1)]
stack = np.stack(uniq)
frames = torch.from_numpy(stack).to(dev)[torch.tensor(idx, device=dev)]
pts = torch.from_numpy(synth.harris_tracks(uniq[0], NT)).to(dev)
LK = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
params = lk_params(fbt=1.0, **LK)
fb = FrameBatch(frames, LK["winSize"], LK["maxLevel"])
for _ in range(2):
    fb.build()
    out = track_pairs(fb, fb, pts, params, 0, 1, FPG)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fb.build()
out = track_pairs(fb, fb, pts, params, 0, 1, FPG)
e1.record()
barrier()
t_klt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t_klt, op=dist.ReduceOp.MAX)
valid = out[1].float().mean().item()
del frames, fb, out
torch.cuda.empty_cache()

# ---- stage 2: global BA, cameras sharded across ranks --------------------------------------------------
F = FPG * world
pw = synth.scene_points(NT, seed=7)
P, cw = synth.scene_observations(pw, F, K_row=K, step=0.02 * 300 / max(F, 300) * 1.0, noise=0.1, seed=11)
rng = np.random.default_rng(3)
z = np.concatenate((P[0].T.ravel(), P[1].T.ravel())).astype(np.float64)
x0 = np.concatenate((pw + rng.normal(0, 0.05, pw.shape), (cw + rng.normal(0, 0.01, cw.shape))[1:], np.zeros((F - 1, 3)))).ravel()
ba = NLS.BundleAdjuster(K, z, x0, NT, F - 1, shard=world > 1)
ba.step()                      # warm-up: cuBLAS / cuSOLVER lazy loading, NCCL channel setup
ba.x.copy_(torch.from_numpy(x0).to(dev))
barrier()
t0 = time.perf_counter()
hist = []
for it in range(args.ba_iters):
    hist.append(ba.step())
barrier()
t_ba = (time.perf_counter() - t0) * 1e3
gathered = [torch.empty_like(ba.x) for _ in range(world)]
if world > 1:
    dist.all_gather(gathered, ba.x)
    identical = all(torch.equal(gathered[0], g) for g in gathered)
else:
    identical = True
if rank == 0:
    xs = ba.x.cpu().numpy()
    print("C5-shape run on %d GPU(s): %dx%d, %d frames per GPU, %d tracks" % (world, W, H, FPG, NT))
    print("  KLT  (per-rank shard, max over ranks): %.2f ms for %d pairs -> %.0f frames/s whole job, valid %.3f" % (
        t_klt.item(), FPG, world * FPG / t_klt.item() * 1e3, valid))
    print("  BA   nt=%d nc=%d nx=%d: %d iterations in %.1f ms (%.1f ms/iteration), bit-identical across ranks: %s" % (
        NT, F - 1, x0.size, args.ba_iters, t_ba, t_ba / args.ba_iters, identical))
    print("       rms residual: " + " ".join("%.4f" % h[0] for h in hist))
    print("       point error vs truth: %.3e -> %.3e m" % (np.abs(x0[:3 * NT].reshape(-1, 3) - pw).max(), np.abs(xs[:3 * NT].reshape(-1, 3) - pw).max()))
if world > 1:
    dist.destroy_process_group()
