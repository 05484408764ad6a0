"""One n=2048 LAP on structured -cdist costs (for the ncu capture of lap_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pleas_merging_b200 import ops
rng = np.random.default_rng(0); n = 2048
X = rng.standard_normal((n, 512)).astype(np.float32)
Y = X[rng.permutation(n)] + 0.5 * rng.standard_normal((n, 512)).astype(np.float32)
A = -np.sqrt(np.maximum((X * X).sum(1)[:, None] + (Y * Y).sum(1)[None] - 2 * X @ Y.T, 0)).astype(np.float32)
Ad = torch.from_numpy(A).cuda()
for _ in range(2):
    outs, obj, st = ops.lap_solve_batched([Ad], True)
torch.cuda.synchronize(); print("ok", float(obj[0]), int(st[0]))
