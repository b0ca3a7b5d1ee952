"""Per-step timeline of the bench's e2e loop: host time spent in prefetch() / run(), GPU time between successive step starts and the
GPU's idle gap before each step (start mark of step s minus the end of step s-1's last kernel):  python tools/e2e_steps.py [steps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from velocity_b200.sfm import SfmSequence

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
K, frames_np, p0_np, p3_np, times_np, Z = bench.make_sequence(bench.SEED, bench.SEED)
fh = torch.from_numpy(frames_np).pin_memory()
p0, p3, tm = torch.from_numpy(p0_np).pin_memory(), torch.from_numpy(p3_np).pin_memory(), torch.from_numpy(times_np).pin_memory()
seq = SfmSequence(K, bench.H, bench.W, bench.NFRAMES, bench.NPTS, fbt=bench.FBT, ba_iters=bench.BA_ITERS, chunk=25, **bench.LK)
out = dict(S=torch.empty((bench.NFRAMES, 9)).pin_memory(), S_ba=torch.empty((bench.NFRAMES, 9)).pin_memory(),
           B=torch.empty((bench.NFRAMES, 14)).pin_memory(), P=torch.empty((5, bench.NPTS, bench.NFRAMES)).pin_memory())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    seq.run(fh, p0, p3, tm, out=out)
for trial in range(3):
    torch.cuda.synchronize()
    seq.marks = []
    rows = []
    ends = []
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    t0 = time.perf_counter()
    seq.prefetch(fh, p0, p3, tm)
    for s in range(steps):
        ta = time.perf_counter()
        flush.zero_()
        if s + 1 < steps:
            seq.prefetch(fh, p0, p3, tm)
        tb = time.perf_counter()
        seq.run(fh, p0, p3, tm, out=out, sync=False)
        tc = time.perf_counter()
        ev = torch.cuda.Event(enable_timing=True); ev.record(); ends.append(ev)
        rows.append(((tb - ta) * 1e3, (tc - tb) * 1e3))
    seq.wait_results()
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    torch.cuda.synchronize()
    marks, seq.marks = seq.marks, None
    starts = [ev for nm, ev in marks if nm == "start"]
    tracks = [ev for nm, ev in marks if nm == "track"]
    print("trial %d: %.2f ms/step (events), %.2f wall" % (trial, e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps))
    for s in range(steps):
        gap = ends[s - 1].elapsed_time(starts[s]) if s else e0.elapsed_time(starts[0])
        print("  step %2d: host prefetch %.2f ms, host run %.2f ms | gpu: idle before start %.2f, track %.2f, rest %.2f" % (
            s, rows[s][0], rows[s][1], gap, starts[s].elapsed_time(tracks[s]), tracks[s].elapsed_time(ends[s])))
