"""Benchmark of the merge hot path on BASELINE.json's config 2: ResNet-50 pair, activation
matching over synthetic 224x224 batches of 32 (+ the rest of the merge: LAP, partial_merge,
PLeaS closed form over MAX_STEPS+1 batches).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one calibration batch through activation matching's accumulation loop (two model
forwards + one fused pack/GEMM/epilogue per tap — 174 taps for ResNet-50).  ``value`` is
calibration samples/s with the batches already resident in HBM; ``e2e`` is the same metric
through the public ``activation_matching`` call with pinned HOST batches (H2D copy of every
batch and the D2H read of the permutations inside the timed region).  The whole merge
(activation matching -> LAP -> partial_merge -> PLeaS) is timed once as ``merge_wall_s``.

``--impl reference`` times the reference's CPU path (the oracle restatement, torch CPU + numpy
+ the C LAP port, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32
HW = 224
METRIC = "rn50_pair_merge_calib_samples_per_s"
UNIT = "samples/s"


def num_classes_of(model_name):
    return 345 if model_name == "resnet101_domainnet" else 1000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="resnet50", help="torchvision arch; 'resnet101_domainnet' = config 4 "
                    "(ResNet-101 with a 345-class head, run_domainnet.py:182-186)")
    ap.add_argument("--batch", type=int, default=None, help="samples per batch (default 32; 16 for config 4)")
    ap.add_argument("--pleas-steps", type=int, default=None, help="MAX_STEPS of the PLeaS pass (default 400, "
                    "or 4*steps when steps < 100)")
    ap.add_argument("--no-merge", action="store_true", help="skip the one-off whole-merge timing")
    ap.add_argument("--tf32-convs", action="store_true", help="let cuDNN use TF32 for the source models' convolutions "
                    "(PyTorch's default; NOT the headline: activations then differ from the fp32 CPU reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=8, help="samples per CPU-baseline batch")
    args = ap.parse_args()
    global BATCH, METRIC
    if args.batch is None:
        args.batch = 16 if args.model == "resnet101_domainnet" else 32
    BATCH = args.batch
    if args.model != "resnet50":
        METRIC = f"{args.model}_pair_merge_calib_samples_per_s"
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 7:
                reasons |= {n for n, v in zip(names, r[3:7]) if v.lower().startswith("active")}
        mx = max((float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def make_models(name, device=None):
    import torch
    import torchvision

    def build(seed):
        torch.manual_seed(seed)
        if name == "resnet101_domainnet":
            m = torchvision.models.resnet101()
            m.fc = torch.nn.Linear(2048, 345)
            return m.eval()
        return getattr(torchvision.models, name)().eval()

    m1, m2 = build(0), build(1)
    if device is not None:
        m1, m2 = m1.to(device), m2.to(device)
    return m1, m2


# ------------------------------------------------------------------------------ reference arm

def run_reference(args):
    """The reference's CPU path (oracle port) on a bounded sample: K steps of `cpu_sample`
    samples each through matching_costs (two forwards + 174 -cdist taps), all host threads."""
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host thread it can
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count())
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_oracle as O

    torch.set_num_threads(os.cpu_count())
    m1, m2 = make_models(args.model)
    import pleas_merging_b200 as P

    spec = P.get_permutation_spec(m1, ((1, 3, 64, 64),))
    jspec = [{"key": (k.key, k.axis), "size": pg.size, "state": sorted((a.key, a.axis) for a in pg.state),
              "node": sorted((a.key, a.axis) for a in pg.node)} for k, pg in spec.items()]
    b = args.cpu_sample
    g = torch.Generator().manual_seed(123)
    steps = min(args.steps, 3)  # bounded: ~10 s of CPU work per step on 8 cores
    warm = min(args.warmup, 1)
    loader = [(torch.randn(b, 3, HW, HW, generator=g), 0) for _ in range(steps + warm)]
    if warm:
        O.matching_costs(jspec, m1, m2, loader[:warm], warm, "cdist", "sum")
    t0 = time.perf_counter()
    O.matching_costs(jspec, m1, m2, loader[warm:], steps, "cdist", "sum")
    dt = time.perf_counter() - t0
    v = steps * b / dt
    sample = f"{steps} steps x {b} samples of {HW}x{HW} through the oracle port of activation_matching's cost loop"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} pair, activation_matching cost loop (-cdist), {b} samples/step "
                               f"(bounded sample of the 32-sample batch), CPU reference path"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------ B200 arm

def cpu_baseline(args, spec):
    import torch

    from oracle import ref_oracle as O

    torch.set_num_threads(os.cpu_count())
    m1, m2 = make_models(args.model)
    jspec = [{"key": (k.key, k.axis), "size": pg.size, "state": sorted((a.key, a.axis) for a in pg.state),
              "node": sorted((a.key, a.axis) for a in pg.node)} for k, pg in spec.items()]
    b = args.cpu_sample
    g = torch.Generator().manual_seed(123)
    loader = [(torch.randn(b, 3, HW, HW, generator=g), 0) for _ in range(3)]
    O.matching_costs(jspec, m1, m2, loader[:1], 1, "cdist", "sum")  # warm-up
    t0 = time.perf_counter()
    O.matching_costs(jspec, m1, m2, loader[1:], 2, "cdist", "sum")
    dt = time.perf_counter() - t0
    return {"value": 2 * b / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"2 batches x {b} samples of {HW}x{HW} through the oracle port of the activation_matching "
                      f"cost loop (2 forwards + all -cdist taps), after 1 warm-up batch"}


def run_b200(args):
    # rank 0 prints ONE JSON line on stdout: anything else that writes to fd 1 (NCCL's version banner,
    # library chatter) is sent to stderr for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(line, flush=True)


def _run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    import pleas_merging_b200 as P
    from pleas_merging_b200 import _native, ops
    import importlib

    AM = importlib.import_module("pleas_merging_b200.methods.activation_matching")

    # exact-fp32 library forwards: the permutations must reproduce the reference's
    torch.backends.cudnn.allow_tf32 = bool(args.tf32_convs)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    K, W = args.steps, max(args.warmup, 0)
    m1, m2 = make_models(args.model, device)
    spec = P.get_permutation_spec(m1, ((1, 3, HW, HW),))
    taps = sum(len(pg.node) for pg in spec.values())

    # synthetic calibration batches, resident in HBM (weak scaling: K batches per GPU)
    gen = torch.Generator(device=device).manual_seed(123 + rank)
    n_dev = min(K + W, 16)  # distinct batches; each step's activations (~10 GB) dwarf the 126 MB L2
    dev_batches = [torch.randn(BATCH, 3, HW, HW, generator=gen, device=device) for _ in range(n_dev)]

    runner = AM.CalibrationRunner(spec, m1, m2, ops.MODE_NEG_CDIST, accumulate="sum", use_cuda_graph=True)
    acc = runner.acc

    def step(i):
        runner.run(dev_batches[i % n_dev])

    with torch.inference_mode():
        _native.LAUNCH_COUNTS.clear()
        runner._eager(dev_batches[0])  # un-captured batch: counts this library's launches per step
        launches_per_step = sum(_native.LAUNCH_COUNTS.values())
        for i in range(max(W, 2)):  # >= 2 so the CUDA graph of the step exists before timing
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        torch.cuda.nvtx.range_push("plb_timed")
        for i in range(K):
            step(W + i)
        torch.cuda.nvtx.range_pop()
        if world > 1:  # the path's one exchange: all-reduce of the cost accumulators over NVLink
            dist.all_reduce(acc.flat)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        # per-launch CUDA-event timing of the dominant kernel: the timed region replays a CUDA graph
        # (events cannot bracket nodes of a replay), so the same steps are re-run un-captured with
        # events around every GEMM launch on the launching stream
        ops.GEMM_TIMER, ops.PACK_TIMER, ops.DIRECT_TIMER = [], [], []
        overlap_was, acc.overlap = acc.overlap, False  # time the kernel alone, not time-sliced with cuDNN
        torch.cuda.nvtx.range_push("plb_eager")
        for i in range(min(K, 5)):
            runner._eager(dev_batches[(W + i) % n_dev])
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        acc.overlap = overlap_was
        timer, ops.GEMM_TIMER = ops.GEMM_TIMER, None
        pack_timer, ops.PACK_TIMER = ops.PACK_TIMER, None
        direct_timer, ops.DIRECT_TIMER = ops.DIRECT_TIMER, None
        launches = launches_per_step * K
    gemm_ms = sum(t[0].elapsed_time(t[1]) for t in timer)
    gemm_flops = sum(t[2] for t in timer)
    by_class = {}  # tensor-bound launches (one large tap each) vs the grouped small-tap launches, per tile width
    for a, b, f, bn, nprob in timer:
        c = by_class.setdefault(f"bn{bn}_" + ("single" if nprob == 1 else "grouped"), [0.0, 0.0, 0])
        c[0] += a.elapsed_time(b)
        c[1] += f
        c[2] += 1
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * BATCH / (ms_max / 1e3)

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.model} pair (random init, eval), activation_matching accumulation over "
                                  f"{K} batches of {BATCH}x3x{HW}x{HW} per GPU, -cdist statistic on {taps} taps, "
                                  f"accumulate=sum, 3xTF32 tcgen05 GEMM, "
                                  + ("TF32 cuDNN forwards (secondary number)" if args.tf32_convs else "exact-fp32 cuDNN forwards"),
                      "parallelism": f"batch-sharded x{world}, one NCCL all-reduce of the cost matrices",
                      "l2": "inputs larger than L2: every step streams ~10 GB of activations"},
           "gpu_launches": launches}
    pk = peaks()
    tf32_peak = pk["bf16_tflops_sustained"] / 2.0
    achieved = 3.0 * gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    out["roofline"] = {
        "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the ncu --set full capture
        # profiles/gemm_v2_c2048_r01_raw.csv (C=2048, K=1568 tap: the 13.15-GFLOP shape that makes up
        # 72 of the 174 launches and 88 % of the GEMM FLOPs of a step); its algorithmic operand bytes are
        # (2048+2048)*1568*8 = 51.4 MB of packed planes, i.e. no re-reads
        "traffic": 53.36e6, "traffic_algorithmic_bytes": 51.4e6, "kernel": "gemm3xtf32_v2_kernel",
        "note": f"achieved = 3 x algorithmic FLOPs (3xTF32 issues three tensor-pipe passes; algorithmic = "
                f"2*C^2*K per tap, {gemm_flops / min(K, 5) / 1e12:.3f} TFLOP per step) / summed CUDA-event time of "
                f"{len(timer)} GEMM launches ({min(K, 5)} steps re-run un-captured right after the timed "
                f"CUDA-graph region); peak = {pk['source']} sustained bf16 "
                f"{pk['bf16_tflops_sustained']} TFLOP/s / 2 (TF32 dense rate)",
        "algorithmic_tflops": gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0,
        "kernel_ms_per_step": gemm_ms / min(K, 5), "kernel_share_of_step": (gemm_ms / min(K, 5)) / (ms / K),
        # launches split by class: "single" = one tap (>= 3 GFLOP, tensor-bound) per launch, "grouped" = the
        # small taps of a batch in one persistent launch per tile width (HBM/latency-bound: C/4 flop/B)
        "by_class": {k: {"launches": v[2], "ms_per_step": v[0] / min(K, 5),
                         "algorithmic_tflops": v[1] / (v[0] / 1e3) / 1e12 if v[0] > 0 else 0.0,
                         "tensor_frac": 3.0 * v[1] / (v[0] / 1e3) / 1e12 / tf32_peak if v[0] > 0 else 0.0}
                     for k, v in sorted(by_class.items())}}
    # second-largest kernel of this library: the operand pack (HBM-bound by construction: reads every
    # activation once, writes its tf32 hi/lo planes, accumulates the row sums of squares)
    pack_ms = sum(a.elapsed_time(b) for a, b, _ in pack_timer)
    pack_bytes = sum(n for _, _, n in pack_timer)
    if pack_ms > 0:
        gbs = pack_bytes / (pack_ms / 1e3) / 1e9
        out["roofline_pack"] = {
            "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
            "traffic": 255.9e6, "traffic_algorithmic_bytes": 308.3e6, "kernel": "pack_split_kernel",
            "kernel_ms_per_step": pack_ms / min(K, 5), "launches_per_step": len(pack_timer) // min(K, 5),
            "note": "achieved = algorithmic bytes (12 B per activation element: 4 read + 8 written as tf32 hi/lo "
                    "planes) / summed CUDA-event time of the pack launches of the same un-captured steps; traffic = "
                    "dram read (102.8 MB = the operand) + write (153.1 MB; the rest of the 205.5 MB of planes is still in L2 "
                    "when the launch ends) of the C=64, K=401408 launch in profiles/pack_r01_raw.csv"}
    # fused narrow-tap kernel (C <= 128): HBM-bound, reads each fp32 activation once
    d_ms = sum(t[0].elapsed_time(t[1]) for t in direct_timer)
    if d_ms > 0:
        d_bytes = sum(t[2] for t in direct_timer)
        gbs = d_bytes / (d_ms / 1e3) / 1e9
        out["roofline_narrow_taps"] = {
            "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
            "traffic": None, "kernel": "gram_direct_kernel", "kernel_ms_per_step": d_ms / min(K, 5),
            "launches_per_step": len(direct_timer) // min(K, 5),
            "algorithmic_tflops": sum(t[3] for t in direct_timer) / (d_ms / 1e3) / 1e12,
            "note": "achieved = algorithmic bytes (both fp32 activations of the tap, read once: 2*C*K*4) / summed "
                    "CUDA-event time of the launches of the same un-captured steps"}
    out["clocks"] = clocks

    # ---- e2e through the public API with pinned host batches (H2D + LAP + D2H of the perms inside)
    with torch.inference_mode():
        host = [(b.cpu().pin_memory(), 0) for b in dev_batches]
    Ke = K
    loader = [host[i % n_dev] for i in range(Ke * world)]  # weak scaling: K batches per GPU
    P.activation_matching(spec, m1, m2, loader[:2 * world], 2 * world, accumulate="sum",
                          distributed=world > 1)  # warm the public path
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    perm, costs = P.activation_matching(spec, m1, m2, loader, Ke * world, output_costs=True, accumulate="sum",
                                        distributed=world > 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    d2h = sum(p.numel() * 8 for p in perm.values())
    out["e2e"] = {"value": Ke * world * BATCH / dt, "unit": UNIT, "h2d_bytes_per_step": BATCH * 3 * HW * HW * 4,
                  "d2h_bytes_per_step": d2h / Ke, "wall_s": dt,
                  "note": "activation_matching(spec, m1, m2, pinned host loader, K, accumulate='sum'): H2D of "
                          "every batch, all taps, batched GPU LAP, D2H of the permutations"}

    # ---- the whole merge once: activation matching -> partial_merge -> PLeaS closed form.
    # Multi-GPU runs time the SAME fixed job (strong scaling): the 100 + 401 batches are dealt to the
    # ranks, cost matrices are all-reduced, normal equations reduced onto the layers' owner ranks,
    # solves run layer-parallel and the fitted weights are all-reduced.
    if not args.no_merge:
        ps = args.pleas_steps if args.pleas_steps is not None else (400 if K >= 100 else 4 * K)
        dist_on = world > 1
        t_am = dt
        if dist_on:
            aloader = [host[i % n_dev] for i in range(Ke)]
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            perm, costs = P.activation_matching(spec, m1, m2, aloader, Ke, output_costs=True, accumulate="sum",
                                                distributed=True)
            torch.cuda.synchronize()
            t_am = time.perf_counter() - t0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model3 = P.partial_merge(spec, m1, m2, perm, costs, 0.0)
        torch.cuda.synchronize()
        t_merge = time.perf_counter() - t0
        ploader = [host[i % n_dev] for i in range(ps + 1)]
        t0 = time.perf_counter()
        tstats = {}
        P.train(ploader, m1, m2, model3, spec, perm, costs, 0.0, False, ps, None,
                num_classes=num_classes_of(args.model), model_type="rn50", stats=tstats, distributed=dist_on)
        torch.cuda.synchronize()
        t_train = time.perf_counter() - t0
        walls = torch.tensor([t_am, t_merge, t_train], dtype=torch.float64, device=device)
        if dist_on:
            dist.all_reduce(walls, op=dist.ReduceOp.MAX)
        t_am, t_merge, t_train = walls.tolist()
        out["merge"] = {"activation_matching_s": t_am, "am_batches": Ke, "partial_merge_s": t_merge,
                        "pleas_train_s": t_train, "pleas_batches": ps + 1,
                        "merge_wall_s": t_am + t_merge + t_train, "scaling": "strong" if dist_on else None,
                        "pleas_samples_per_s": (ps + 1) * BATCH / t_train, "pleas_timing": tstats.get("_timing")}

    if rank != 0:  # rank 0 reports
        dist.destroy_process_group()
        return None

    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
        out["cpu_baseline"] = cpu_baseline(args, spec)
    if world > 1:
        dist.destroy_process_group()
    return json.dumps(out)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
