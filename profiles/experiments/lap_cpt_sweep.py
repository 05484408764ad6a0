"""LAP v2: columns-per-thread sweep on N(0,1) costs and on a hard structured instance (row/column
offsets dominate: c_ij = a_i + b_j + noise, like the activation-matching costs of random-init nets).
Run once per PLB_LAP_COLS_PER_THREAD value (the library reads it at first use)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pleas_merging_b200 import ops
rng = np.random.default_rng(0)
out = []
for n in (256, 512, 1024, 2048, 4096):
    for kind in ("randn", "offsets"):
        A = rng.standard_normal((n, n)).astype(np.float32)
        if kind == "offsets":
            A = (A + 30 * rng.standard_normal((n, 1)) + 30 * rng.standard_normal((1, n))).astype(np.float32)
        Ad = torch.from_numpy(A).cuda()
        ops.lap_solve_batched([Ad], True); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.lap_solve_batched([Ad], True); e1.record(); torch.cuda.synchronize()
        out.append("%s n=%d: %.2f ms" % (kind, n, e0.elapsed_time(e1)))
print("cpt=%s impl=%s | " % (os.environ.get("PLB_LAP_COLS_PER_THREAD", "auto"), os.environ.get("PLB_LAP_IMPL", "v3")) + " | ".join(out))
