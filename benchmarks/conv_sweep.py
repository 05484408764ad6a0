"""ResNet-50 convolution classes (batch 32, both source models per launch): the library's 3xTF32 tcgen05
implicit GEMM (plb_conv2d_forward) against cuDNN exact fp32 (what ATen runs with allow_tf32 = False).
Prints one JSON line per class; CUDA-event timed, operands rotating through more than the L2 size."""
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from pleas_merging_b200 import conv  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
# (Cin, H, Cout, k, stride, pad, count in ResNet-50)
CLASSES = [
    (3, 224, 64, 7, 2, 3, 1),
    (64, 56, 64, 1, 1, 0, 1), (64, 56, 64, 3, 1, 1, 3), (64, 56, 256, 1, 1, 0, 4), (256, 56, 64, 1, 1, 0, 2),
    (256, 56, 128, 1, 1, 0, 1), (128, 56, 128, 3, 2, 1, 1), (256, 56, 512, 1, 2, 0, 1),
    (128, 28, 512, 1, 1, 0, 4), (512, 28, 128, 1, 1, 0, 3), (128, 28, 128, 3, 1, 1, 3),
    (512, 28, 256, 1, 1, 0, 1), (256, 28, 256, 3, 2, 1, 1), (512, 28, 1024, 1, 2, 0, 1),
    (256, 14, 1024, 1, 1, 0, 6), (1024, 14, 256, 1, 1, 0, 5), (256, 14, 256, 3, 1, 1, 5),
    (1024, 14, 512, 1, 1, 0, 1), (512, 14, 512, 3, 2, 1, 1), (1024, 14, 2048, 1, 2, 0, 1),
    (512, 7, 2048, 1, 1, 0, 3), (2048, 7, 512, 1, 1, 0, 2), (512, 7, 512, 3, 1, 1, 2),
]
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 32


def timed(fn, reps):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        fn(r)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tot_lib = tot_dnn = 0.0
with torch.no_grad():
    for cin, h, cout, k, s, p, count in CLASSES:
        ma = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).cuda()
        mb = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).cuda()
        nbuf = max(2, int(300e6 // (NB * cin * h * h * 4 * 2)) + 1)
        xs = [(torch.randn(NB, cin, h, h, device="cuda"), torch.randn(NB, cin, h, h, device="cuda")) for _ in range(nbuf)]
        pair = conv.ConvPair(ma, mb)
        ya, yb = pair(*xs[0])
        ra, rb = ma(xs[0][0]), mb(xs[0][1])
        err = max(((ya - ra).abs().max() / ra.abs().max()).item(), ((yb - rb).abs().max() / rb.abs().max()).item())
        t_lib = timed(lambda r: pair(*xs[r % nbuf]), 20)
        t_dnn = timed(lambda r: (ma(xs[r % nbuf][0]), mb(xs[r % nbuf][1])), 20)
        oh = (h + 2 * p - k) // s + 1
        flops = 2.0 * 2 * NB * oh * oh * cout * cin * k * k
        nbytes = 4.0 * 2 * NB * (cin * h * h + cout * oh * oh)
        tot_lib += t_lib * count
        tot_dnn += t_dnn * count
        print(json.dumps({"cin": cin, "hw": h, "cout": cout, "k": k, "stride": s, "count": count,
                          "lib_us": round(t_lib, 1), "cudnn_us": round(t_dnn, 1),
                          "lib_tflops": round(flops / t_lib / 1e6, 1), "cudnn_tflops": round(flops / t_dnn / 1e6, 1),
                          "lib_gbs": round(nbytes / t_lib / 1e3), "max_rel_diff_vs_cudnn": float(f"{err:.2e}")}), flush=True)
print(json.dumps({"resnet50_pair_convs_ms": {"lib": round(tot_lib / 1e3, 3), "cudnn": round(tot_dnn / 1e3, 3)}, "batch": NB}))
