"""One fp64 Cholesky solve of the largest ResNet-50 PLeaS layer shape (K=4608, 512 right-hand sides):
timed with CUDA events; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pleas_merging_b200 import ops
n, nrhs = int(sys.argv[1]) if len(sys.argv) > 1 else 4608, int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(0)
U = torch.randn(2 * n, n, device="cuda", dtype=torch.float64, generator=g)
G0 = U.T @ U
B0 = torch.randn(n, nrhs, device="cuda", dtype=torch.float64, generator=g)
for r in range(reps):
    G, B = G0.clone(), B0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    info = ops.chol_solve_(G, B, 1e-6)
    e1.record()
    torch.cuda.synchronize()
    print("n=%d nrhs=%d: %.2f ms (info %d)" % (n, nrhs, e0.elapsed_time(e1), int(info.item())))
res = (G0 @ B - B0).abs().max() / B0.abs().max()
print("relative residual %.2e" % float(res))
t0 = time.perf_counter(); X = torch.linalg.solve(G0, B0); torch.cuda.synchronize()
t0 = time.perf_counter(); X = torch.cholesky_solve(B0, torch.linalg.cholesky(G0)); torch.cuda.synchronize()
print("torch (cuSOLVER) cholesky+solve: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
