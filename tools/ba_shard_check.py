"""torchrun --nproc-per-node N tools/ba_shard_check.py : camera-sharded bundle adjustment over NCCL
must reproduce the single-GPU result and the reference golden (run on the GPU box)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import sfm_oracle as S  # noqa: E402  (checker only)
from util import golden  # noqa: E402
from velocity_b200 import NLS  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = golden("ba_medium")
z, x, nt, nc = S._ba_pack(g["P"], g["pw0"], g["cw0"])
ba = NLS.BundleAdjuster(g["K"], z, x, nt, nc, shard=True)
single = NLS.BundleAdjuster(g["K"], z, x, nt, nc, shard=False)
for it in range(10):
    f, xr = ba.step()
    f1, xr1 = single.step()
    if xr < 1e-7:
        break
xs, x1 = ba.x.cpu().numpy(), single.x.cpu().numpy()
cw, pw = S._ba_unpack(xs, nt, nc)
ok_ref = np.allclose(cw, g["cw"], rtol=1e-6, atol=1e-7) and np.allclose(pw, g["pw"], rtol=1e-6, atol=1e-7)
ok_single = np.allclose(xs, x1, rtol=1e-7, atol=1e-8)  # all-reduce order differs from the sequential sum
gathered = [torch.empty_like(ba.x) for _ in range(world)]
dist.all_gather(gathered, ba.x)
identical = all(torch.equal(gathered[0], t) for t in gathered)
print("rank %d/%d: iterations %d  matches reference %s  matches single-GPU %s  bit-identical across ranks %s" % (
    rank, world, it + 1, ok_ref, ok_single, identical), flush=True)
dist.destroy_process_group()
sys.exit(0 if (ok_ref and ok_single and identical) else 1)
