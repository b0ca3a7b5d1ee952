"""e2e step breakdown with and without cross-sequence upload prefetch (SfmSequence.prefetch)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from velocity_b200.sfm import SfmSequence

K, frames_np, p0_np, p3_np, times_np, Z = bench.make_sequence(bench.SEED, bench.SEED)
dev = torch.device("cuda", 0)
fh = torch.from_numpy(frames_np).pin_memory()
p0, p3, tm = torch.from_numpy(p0_np).pin_memory(), torch.from_numpy(p3_np).pin_memory(), torch.from_numpy(times_np).pin_memory()
seq = SfmSequence(K, bench.H, bench.W, bench.NFRAMES, bench.NPTS, fbt=bench.FBT, ba_iters=bench.BA_ITERS, chunk=25, **bench.LK)
out = dict(S=torch.empty((bench.NFRAMES, 9)).pin_memory(), S_ba=torch.empty((bench.NFRAMES, 9)).pin_memory(), B=torch.empty((bench.NFRAMES, 14)).pin_memory())
for _ in range(3):
    seq.run(fh, p0, p3, tm, out=out)
for mode in ("plain", "prefetch", "plain", "prefetch"):
    torch.cuda.synchronize()
    seq.marks = []
    t0 = time.perf_counter()
    steps = 6
    if mode == "prefetch":
        seq.prefetch(fh, p0, p3, tm)
    for s in range(steps):
        if mode == "prefetch" and s + 1 < steps:
            seq.prefetch(fh, p0, p3, tm)
        seq.run(fh, p0, p3, tm, out=out)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    marks, seq.marks = seq.marks, None
    st = {}
    for (na, ea), (nb, eb) in zip(marks[:-1], marks[1:]):
        if nb != "start":
            st.setdefault(nb, []).append(ea.elapsed_time(eb))
    print(mode, "wall %.2f ms/step" % wall, {k: [round(x, 1) for x in v] for k, v in st.items()})

# timeline: when do the prefetched uploads actually execute relative to the compute stream?
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record()
tl = []
def cmark(name, stream):
    ev = torch.cuda.Event(enable_timing=True); ev.record(stream); tl.append((name, ev))
comp = torch.cuda.current_stream()
cmark("pf0 begin", seq.copy_stream); seq.prefetch(fh, p0, p3, tm); cmark("pf0 end", seq.copy_stream)
for s in range(3):
    cmark("pf%d begin" % (s + 1), seq.copy_stream); seq.prefetch(fh, p0, p3, tm); cmark("pf%d end" % (s + 1), seq.copy_stream)
    cmark("run%d begin" % s, comp)
    seq.marks = []
    seq.run(fh, p0, p3, tm, out=out)
    for nm, ev in seq.marks:
        tl.append(("run%d %s" % (s, nm), ev))
    seq.marks = None
torch.cuda.synchronize()
for nm, ev in tl:
    print("%-16s %8.2f ms" % (nm, base.elapsed_time(ev)))
