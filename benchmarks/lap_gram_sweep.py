"""BASELINE.json config 5: standalone LAP + Gram microbench sweep.

LAP: n in {64..4096}, i.i.d. N(0,1) and structured -cdist costs, batch sizes 1 and 37 (the
ResNet-50 group mix), GPU kernel (through the C ABI) vs SciPy linear_sum_assignment on the host,
with assignment equality checked.  Gram: C in {64..4096}, K in {1e4..1e6}, vs torch fp32 matmul
(allow_tf32=False) on the same GPU, with relative error against fp64 reported.
Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pleas_merging_b200 import ops  # noqa: E402


def cuda_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def lap_sweep():
    from scipy.optimize import linear_sum_assignment as lsa

    rng = np.random.default_rng(0)
    for n in (64, 128, 256, 512, 1024, 2048, 4096):
        for dist in ("randn", "cdist"):
            if dist == "randn":
                A = rng.standard_normal((n, n)).astype(np.float32)
            else:
                X = rng.standard_normal((n, 512)).astype(np.float32)
                Y = X[rng.permutation(n)] + 0.5 * rng.standard_normal((n, 512)).astype(np.float32)
                A = -np.sqrt(np.maximum((X * X).sum(1)[:, None] + (Y * Y).sum(1)[None] - 2 * X @ Y.T, 0)).astype(np.float32)
            t0 = time.perf_counter()
            _, col = lsa(A, maximize=True)
            t_scipy = (time.perf_counter() - t0) * 1e3
            Ad = torch.from_numpy(A).cuda()
            outs, _, st = ops.lap_solve_batched([Ad], True)
            same = bool((outs[0].cpu().numpy() == col).all())
            t_gpu = cuda_time(lambda: ops.lap_solve_batched([Ad], True), iters=3, warm=1)
            print(json.dumps({"bench": "lap", "n": n, "dist": dist, "scipy_ms": round(t_scipy, 3),
                              "gpu_ms": round(t_gpu, 3), "identical_to_scipy": same,
                              "algorithmic_GBps": round(n * n * 4 / (t_gpu * 1e-3) / 1e9, 2)}), flush=True)
    # ResNet-50 group mix in one launch (37 problems)
    sizes = [64] * 7 + [128] * 8 + [256] * 13 + [512] * 7 + [1024, 2048]
    mats = [rng.standard_normal((n, n)).astype(np.float32) for n in sizes]
    t0 = time.perf_counter()
    cols = [lsa(m, maximize=True)[1] for m in mats]
    t_scipy = (time.perf_counter() - t0) * 1e3
    dm = [torch.from_numpy(m).cuda() for m in mats]
    outs, _, _ = ops.lap_solve_batched(dm, True)
    same = all(bool((o.cpu().numpy() == c).all()) for o, c in zip(outs, cols))
    t_gpu = cuda_time(lambda: ops.lap_solve_batched(dm, True), iters=3, warm=1)
    print(json.dumps({"bench": "lap_batch37_rn50_mix", "scipy_ms_total": round(t_scipy, 2), "gpu_ms_one_launch": round(t_gpu, 2),
                      "identical_to_scipy": same}), flush=True)


def gram_sweep():
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    for C in (64, 256, 1024, 2048, 4096):
        for K in (10_000, 100_000, 1_000_000):
            if C * K > 1.1e9:
                continue
            x = torch.relu(torch.randn(C, K, generator=g, device="cuda"))
            y = torch.relu(torch.randn(C, K, generator=g, device="cuda"))
            kb = (K + 15) // 16
            pa, pb = ops.Planes(C, kb, "cuda"), ops.Planes(C, kb, "cuda")
            plan = ops.GemmPlan(pa, pb, C, C, kb)
            out = torch.empty(C, C, device="cuda")

            def ours():
                ops.pack_split(x, 0, pa)
                ops.pack_split(y, 0, pb)
                plan.run()
                plan.finalize(out)

            def gemm_only():
                plan.run()

            t_all = cuda_time(ours)
            t_gemm = cuda_time(gemm_only)
            t_torch = cuda_time(lambda: torch.matmul(x, y.T))
            ref = (x[:64].double() @ y.double().T)
            ours()
            err = float((out[:64].double() - ref).abs().max() / ref.abs().max())
            err_t = float((torch.matmul(x[:64], y.T).double() - ref).abs().max() / ref.abs().max())
            fl = 2.0 * C * C * K
            print(json.dumps({"bench": "gram", "C": C, "K": K, "ours_ms_pack+gemm+finalize": round(t_all, 3),
                              "ours_gemm_ms": round(t_gemm, 3), "torch_fp32_matmul_ms": round(t_torch, 3),
                              "ours_gemm_alg_TFLOPs": round(fl / (t_gemm * 1e-3) / 1e12, 1),
                              "ours_total_alg_TFLOPs": round(fl / (t_all * 1e-3) / 1e12, 1),
                              "torch_TFLOPs": round(fl / (t_torch * 1e-3) / 1e12, 1),
                              "relerr_ours": float(f"{err:.2e}"), "relerr_torch_fp32": float(f"{err_t:.2e}")}), flush=True)
            del x, y, pa, pb, plan, out


if __name__ == "__main__":
    which = sys.argv[1:] or ["lap", "gram"]
    if "lap" in which:
        lap_sweep()
    if "gram" in which:
        gram_sweep()
