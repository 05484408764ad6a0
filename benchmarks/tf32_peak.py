"""Measures the dense TF32 tensor-core rate of the box the way MEASURED_PEAKS.json measured bf16:
torch.matmul fp32 8192^3 with allow_tf32=True (cuBLAS), best of 10 (burst) and back to back for ~3 s
(sustained).  Prints one JSON line; bench.py imports `measure` for its roofline denominators."""
import json
import time

import torch


def measure(n=8192, sustain_s=3.0, device="cuda"):
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=device)
        b = torch.randn(n, n, device=device)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            time.sleep(0.05)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 0
        t0 = time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < sustain_s:
            for _ in range(20):
                a @ b
            reps += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        sustained = reps * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return {"tf32_tflops": best, "tf32_tflops_sustained": sustained,
            "how": f"torch.matmul fp32 {n}^3 allow_tf32=True: best of 10 (burst), back to back {sustain_s:.0f} s (sustained)"}


if __name__ == "__main__":
    print(json.dumps(measure()))
