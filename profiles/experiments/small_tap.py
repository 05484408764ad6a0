"""The HBM-bound small-C tap (ResNet-50 stem: C=64, K=32*112*112) through cross_statistic, a few times,
with CUDA-event timing of pack / GEMM / finalize — the target of the ncu capture gemm_c64_r01."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pleas_merging_b200 import ops
C, HW = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), (int(sys.argv[2]) if len(sys.argv) > 2 else 112)
x = torch.randn(32, C, HW, HW, device="cuda"); y = torch.randn(32, C, HW, HW, device="cuda")
ops.GEMM_TIMER, ops.DIRECT_TIMER = [], []
for i in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    G = ops.cross_statistic(x, y, 1, ops.MODE_NEG_CDIST)
    e1.record()
    torch.cuda.synchronize()
    K = 32 * HW * HW
    if ops.DIRECT_TIMER:  # fused narrow-tap kernel: reads the fp32 activations once
        a, b, nbytes, f = ops.DIRECT_TIMER[-1]
        print("tap C=%d K=%d: total %.3f ms, fused kernel %.3f ms = %.2f TB/s of activation reads (%.1f algorithmic TFLOP/s)"
              % (C, K, e0.elapsed_time(e1), a.elapsed_time(b), nbytes / a.elapsed_time(b) / 1e9, f / a.elapsed_time(b) / 1e9))
    else:
        a, b, f, bn, _ = ops.GEMM_TIMER[-1]
        print("tap C=%d K=%d: total %.3f ms, GEMM %.3f ms = %.2f TB/s of plane reads (%.1f algorithmic TFLOP/s)"
              % (C, K, e0.elapsed_time(e1), a.elapsed_time(b), 2 * C * K * 8 / a.elapsed_time(b) / 1e9, f / a.elapsed_time(b) / 1e9))
