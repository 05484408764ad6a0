"""Per-tap-class timing of the fused cross-Gram kernels on the ResNet-50 tap mix (B = 32, 224x224):

    python benchmarks/gram_tap_sweep.py [--impl tma|packed|direct] [--json out.jsonl]

For every (C, HW) class of SURVEY.md 8a the tap's kernels (Gram + epilogue) are timed with CUDA events
over rotating operand pairs whose total size exceeds the 126 MB L2 (cold operands), and reported as
algorithmic TFLOP/s (2 C^2 K), tensor-pipe fraction (3 x algorithmic / measured TF32 rate) and
algorithmic GB/s (2 C K 4 bytes) against the measured HBM copy rate."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pleas_merging_b200 import ops  # noqa: E402

CLASSES = [(64, 112, 3), (64, 56, 19), (128, 56, 3), (128, 28, 21), (256, 56, 14), (256, 28, 3), (256, 14, 33),
           (512, 28, 18), (512, 14, 3), (512, 7, 15), (1024, 14, 26), (2048, 7, 14)]  # (C, H = W, taps per batch)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="tma", choices=["tma", "packed", "direct"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--tf32-peak", type=float, default=757.0)
    ap.add_argument("--hbm-peak", type=float, default=6550.0)
    ap.add_argument("--only", default=None, help="C,H filter, e.g. 256,56")
    ap.add_argument("--nbuf", type=int, default=None, help="operand pairs to rotate through (default: enough to exceed L2; "
                    "1 keeps a small tap L2-resident)")
    args = ap.parse_args()
    if args.impl != "tma":
        ops.TMA_GRAM = False
    if args.impl == "packed":
        ops.DIRECT_MAX_ROWS = 0
    dev = torch.device("cuda")
    rows, total_ms = [], 0.0
    for C, H, count in CLASSES:
        if args.only and args.only != f"{C},{H}":
            continue
        B = args.batch
        nbytes = 2 * B * C * H * H * 4
        nbuf = args.nbuf or max(2, min(16, int(300e6 // nbytes) + 1))
        g = torch.Generator(device=dev).manual_seed(C * 1000 + H)
        xs = [torch.relu(torch.randn(B, C, H, H, generator=g, device=dev)) for _ in range(nbuf)]
        ys = [torch.relu(torch.randn(B, C, H, H, generator=g, device=dev)) for _ in range(nbuf)]
        out = torch.zeros(C, C, device=dev)
        q = torch.zeros(2, C, dtype=torch.float64, device=dev)
        K = B * H * H
        if args.impl == "tma" and ops.tma_gram_eligible(xs[0], ys[0], 1):
            plan = ops.TmaGramPlan(C, B, H * H, dev)
            kind = f"tma cg{plan.cta_group} splits{plan.splits}"

            def run(i):
                plan.run(xs[i % nbuf], ys[i % nbuf], 1, q[0], q[1])
                plan.finalize(out, ops.MODE_NEG_CDIST, q[0], q[1], accumulate=True)
        elif args.impl != "packed" and ops.direct_gram_eligible(xs[0], ys[0], 1):
            plan = ops.DirectGramPlan(C, K, dev)
            kind = f"direct splits{plan.splits}"

            def run(i):
                plan.run(xs[i % nbuf], ys[i % nbuf], 1, q[0], q[1])
                plan.finalize(out, ops.MODE_NEG_CDIST, q[0], q[1], accumulate=True)
        else:
            kb = (K + 15) // 16
            pa, pb = ops.Planes(C, kb, dev), ops.Planes(C, kb, dev)
            plan = ops.GemmPlan(pa, pb, C, C, kb)
            kind = f"pack+gemm splits{plan.splits}"

            def run(i):
                ops.pack_split_pair(xs[i % nbuf], ys[i % nbuf], 1, pa, pb, q[0], q[1])
                plan.run()
                plan.finalize(out, ops.MODE_NEG_CDIST, q[0], q[1], accumulate=True)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        # the reps are captured into ONE CUDA graph (like the calibration step): device time without host launch gaps
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            graph.capture_begin()
            for i in range(args.reps):
                run(i)
            graph.capture_end()
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        flops, byts = 2.0 * C * C * K, 2.0 * C * K * 4
        row = {"C": C, "HW": H * H, "K": K, "taps": count, "impl": kind, "us": ms * 1e3,
               "alg_tflops": flops / ms / 1e9, "tensor_frac": 3 * flops / ms / 1e9 / args.tf32_peak,
               "alg_gbs": byts / ms / 1e6, "hbm_frac": byts / ms / 1e6 / args.hbm_peak,
               "ms_per_step": ms * count}
        total_ms += ms * count
        rows.append(row)
        print(f"C={C:5d} HW={H * H:6d} x{count:2d} {kind:28s} {ms * 1e3:8.1f} us  {row['alg_tflops']:6.1f} TF/s "
              f"(tensor {row['tensor_frac']:.2f})  {row['alg_gbs']:7.0f} GB/s (hbm {row['hbm_frac']:.2f})  "
              f"{row['ms_per_step']:.3f} ms/step", flush=True)
    print(f"total {total_ms:.3f} ms per ResNet-50 batch for these classes")
    if args.json:
        with open(args.json, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
