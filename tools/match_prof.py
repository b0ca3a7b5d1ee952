import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from velocity_b200 import match
os.environ["VEL_MATCH_FORCE"] = sys.argv[1] if len(sys.argv) > 1 else "tc"
rng = np.random.default_rng(1)
t = torch.from_numpy(rng.integers(0, 256, (8192, 32), dtype=np.uint8)).cuda()
q = torch.from_numpy(rng.integers(0, 256, (8192, 32), dtype=np.uint8)).cuda()
for _ in range(4):
    match.knn2_hamming256(q, t)
torch.cuda.synchronize()
