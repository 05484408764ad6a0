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
+ the C LAP port, all host threads) on a bounded sample of the same workload (full 32-sample batches).

Beside the headline the line carries: ``roofline`` (dominant kernel: the TMA-fed 3xTF32 Gram of the wide
taps, against the TF32 dense rate MEASURED in this run), ``roofline_narrow_taps`` / ``roofline_packed_gemm`` /
``roofline_pack`` (the other statistics kernels), ``roofline_normal_eq`` (PLeaS normal-equation GEMMs),
``roofline_lap``, ``roofline_chol``, ``cpu_baseline`` (activation-matching cost loop on the host cores),
``cpu_baseline_merge`` (the reference's Adam step on the host cores, extrapolated to the 401-step merge) and
``gpu_library_baseline`` (the reference's own GPU path — ATen cdist per tap + SciPy behind a D2H — on this B200).
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


def workload_of(model_name, batch=None):
    """The workload both arms run (identical `config.workload`): what differs is who executes it."""
    return (f"{model_name} pair (random init, eval), activation_matching cost accumulation over batches of "
            f"{BATCH if batch is None else batch}x3x{HW}x{HW}, -cdist statistic on all taps, accumulate=sum, "
            f"fp32-accurate forwards (no TF32 / bf16 rounding of the products)")


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
    ap.add_argument("--cpu-sample", type=int, default=32, help="samples per CPU-baseline batch (32 = the config's batch)")
    ap.add_argument("--cpu-merge-steps", type=int, default=2, help="timed Adam steps of the CPU merge baseline")
    ap.add_argument("--no-extra-rooflines", action="store_true", help="skip the K4 / LAP / Cholesky / GPU-library legs")
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
    steps = min(args.steps, 20)  # ~6 s of CPU work per 32-sample step: the driver's 20 steps are honoured
    warm = min(args.warmup, 2)
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
        "config": {"workload": workload_of(args.model, b)},
        "implementation": f"CPU reference path: oracle port of activation_matching's cost loop (torch CPU forwards + numpy "
                          f"-cdist per tap), {torch.get_num_threads()} host threads, {steps} steps of {b} samples after "
                          f"{warm} warm-up",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------ B200 arm

def _jspec(spec):
    return [{"key": (k.key, k.axis), "size": pg.size, "state": sorted((a.key, a.axis) for a in pg.state),
             "node": sorted((a.key, a.axis) for a in pg.node)} for k, pg in spec.items()]


def cpu_baseline(args, spec):
    """Activation-matching cost loop of the reference on the host cores (oracle port): full 32-sample batches."""
    import torch

    from oracle import ref_oracle as O

    torch.set_num_threads(os.cpu_count())
    m1, m2 = make_models(args.model)
    jspec = _jspec(spec)
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


def cpu_baseline_merge(args, spec, perm, costs, am_batches, pleas_batches, cpu_am_samples_per_s):
    """The PLeaS half of the merge on the host cores: the oracle port of the reference's Adam step
    (pleas_merging.py:234-302) timed for a few steps on full batches and extrapolated linearly to the
    MAX_STEPS + 1 steps of the job, plus the activation-matching loop extrapolated from cpu_baseline."""
    import numpy as np
    import torch

    from oracle import ref_oracle as O

    torch.set_num_threads(os.cpu_count())
    m1, m2 = make_models(args.model)
    jspec = _jspec(spec)
    p_np = {(k.key, k.axis): v.numpy() for k, v in perm.items()}
    c_np = {(k.key, k.axis): v.float().cpu().numpy() for k, v in costs.items()}
    blocks = O.get_blocks(jspec, p_np, c_np, 0.0)
    init = O.merged_state(jspec, {k: v.numpy() for k, v in m1.state_dict().items()},
                          {k: v.numpy() for k, v in m2.state_dict().items()}, blocks)
    n = 1 + args.cpu_merge_steps
    g = torch.Generator().manual_seed(321)
    loader = [(torch.randn(BATCH, 3, HW, HW, generator=g), 0) for _ in range(n)]
    marks = [time.perf_counter()]
    O.pleas_adam_train(jspec, m1, m2, blocks, init, loader, n - 1, num_classes=num_classes_of(args.model),
                       on_step=lambda idx: marks.append(time.perf_counter()))
    steps = np.diff(marks)[1:]  # first step = warm-up
    s_per_step = float(np.mean(steps))
    am_s = am_batches * BATCH / cpu_am_samples_per_s
    return {"adam_s_per_step": s_per_step, "timed_steps": int(len(steps)), "cores": torch.get_num_threads(),
            "kind": "port", "extrapolated": True,
            "activation_matching_s": am_s, "pleas_train_s": s_per_step * pleas_batches,
            "merge_wall_s": am_s + s_per_step * pleas_batches,
            "sample": f"{len(steps)} Adam steps of the reference's PLeaS loop (oracle port of pleas_merging.py:234-302, "
                      f"batches of {BATCH}x3x{HW}x{HW}) after 1 warm-up step, scaled linearly to {pleas_batches} steps; "
                      f"activation matching scaled from cpu_baseline to {am_batches} batches"}


def gpu_library_baseline(args, spec, m1, m2, host_batches, nb=8):
    """BASELINE.md section 3, second baseline: the reference's own GPU path on this B200 — per tap a
    movedim/reshape copy + ATen ``torch.cdist`` (cuBLAS SGEMM + ~12 elementwise kernels,
    activation_matching.py:31-46), cost matrices copied to the host and SciPy's linear_sum_assignment
    (solvers.py:29-31) — driven through the same dual-model graph via the generic plug-in path."""
    import torch
    from scipy.optimize import linear_sum_assignment

    import pleas_merging_b200 as P

    def torch_cdist(x, y, a):
        x = torch.movedim(x, a, 0).reshape(x.shape[a], -1)
        y = torch.movedim(y, a, 0).reshape(y.shape[a], -1)
        return -torch.cdist(x[None], y[None])[0]

    def scipy_lsa(A, maximize=True):
        ri, ci = linear_sum_assignment(A.detach().cpu().numpy(), maximize=maximize)
        return torch.tensor(ci)

    loader = [host_batches[i % len(host_batches)] for i in range(nb)]
    P.activation_matching(spec, m1, m2, loader[:2], 2, cross_features=torch_cdist, lsa_solver=scipy_lsa,
                          accumulate="sum")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    P.activation_matching(spec, m1, m2, loader, nb, cross_features=torch_cdist, lsa_solver=scipy_lsa, accumulate="sum")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": nb * BATCH / dt, "unit": UNIT, "batches": nb, "wall_s": dt,
            "what": "reference GPU path on this B200: ATen torch.cdist per tap (fp32 cuBLAS, allow_tf32=False) through "
                    "the dual-model graph, pinned host batches, SciPy linear_sum_assignment behind a D2H copy"}


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r02_ncu_summary.json)."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        for row in json.load(f):
            if row["kernel"] == kernel:
                return row["dram_bytes"], row
    return None, None


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


def tiny_parity(P, world, device):
    """Multi-GPU runs: batch-sharded activation matching + PLeaS on the tiny pair must equal the
    single-GPU result before anything is timed (the driver's GPU test box has one GPU)."""
    import torch

    from oracle import tinynet

    m1, m2 = tinynet.make_pair(12, 10)
    spec = P.get_permutation_spec(m1, ((1, 3, 16, 16),))
    m1, m2 = m1.to(device), m2.to(device)
    loader = tinynet.make_loader(2 * world + 1, 4, 16)
    p1, c1 = P.activation_matching(spec, m1, m2, loader, len(loader), output_costs=True, accumulate="sum")
    p2, c2 = P.activation_matching(spec, m1, m2, loader, len(loader), output_costs=True, accumulate="sum",
                                   distributed=True)
    ok = all(torch.equal(p1[k], p2[k]) for k in spec)
    ok = ok and all(float((c1[k] - c2[k]).abs().max()) <= 2e-6 * float(c1[k].abs().max()) for k in spec)
    outs = []
    for dist_on in (False, True):
        m3 = P.partial_merge(spec, m1, m2, p1, c1, 0.0)
        P.train(loader, m1, m2, m3, spec, p1, c1, 0.0, False, len(loader) - 1, None, num_classes=10,
                model_type="rn18", distributed=dist_on)
        outs.append({k: v.clone() for k, v in m3.state_dict().items()})
    ok = ok and all(torch.allclose(outs[0][k], outs[1][k], rtol=1e-4, atol=1e-6) for k in outs[0])
    flag = torch.tensor([int(ok)], device=device)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    return bool(flag.item())


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

    import importlib

    import pleas_merging_b200 as P
    from pleas_merging_b200 import _native, ops
    from pleas_merging_b200 import conv as conv_mod

    AM = importlib.import_module("pleas_merging_b200.methods.activation_matching")
    PM = importlib.import_module("pleas_merging_b200.methods.pleas_merging")

    # exact-fp32 library forwards: the permutations must reproduce the reference's
    torch.backends.cudnn.allow_tf32 = bool(args.tf32_convs)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    parity_ok = None
    if world > 1:
        parity_ok = tiny_parity(P, world, device)
        if not parity_ok:
            raise RuntimeError("sharded activation matching / PLeaS differ from the single-GPU result on the tiny pair")

    # roofline denominators: HBM copy rate from MEASURED_PEAKS.json; the TF32 dense rate is measured here
    pk = peaks()
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import tf32_peak

    tf32 = tf32_peak.measure(sustain_s=2.0, device=device)
    tf32_burst = tf32["tf32_tflops"]

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
        launch_mix = dict(_native.LAUNCH_COUNTS)
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
        # per-launch CUDA-event timing of the statistics kernels: the timed region replays a CUDA graph
        # (events cannot bracket nodes of a replay), so the same steps are re-run un-captured with
        # events around every launch on the launching stream
        ops.GEMM_TIMER, ops.PACK_TIMER, ops.DIRECT_TIMER = [], [], []
        conv_mod.CONV_TIMER = []
        overlap_was, acc.overlap = acc.overlap, False  # time the kernel alone, not time-sliced with cuDNN
        n_eager = min(K, 5)
        torch.cuda.nvtx.range_push("plb_eager")
        for i in range(n_eager):
            runner._eager(dev_batches[(W + i) % n_dev])
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        acc.overlap = overlap_was
        timer, ops.GEMM_TIMER = ops.GEMM_TIMER, None
        pack_timer, ops.PACK_TIMER = ops.PACK_TIMER, None
        direct_timer, ops.DIRECT_TIMER = ops.DIRECT_TIMER, None
        conv_timer, conv_mod.CONV_TIMER = conv_mod.CONV_TIMER, None
        launches = launches_per_step * K
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * BATCH / (ms_max / 1e3)
    step_ms = ms / K

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_of(args.model) if not args.tf32_convs else
                      workload_of(args.model).replace("fp32-accurate forwards (no TF32 / bf16 rounding of the products)",
                                                     "TF32 cuDNN forwards (secondary number)")},
           "implementation": {"kernels": f"3xTF32 tcgen05 Gram kernels on {taps} taps, the source models' convolutions on the "
                                         f"3xTF32 tcgen05 implicit-GEMM kernel (PLB_CONV=0: cuDNN fp32), {K} batches per GPU "
                                         f"replayed as one CUDA graph per batch",
                              "parallelism": f"batch-sharded x{world}, one NCCL all-reduce of the cost matrices",
                              "l2": "inputs larger than L2: every step streams ~10 GB of activations"},
           "gpu_launches": launches, "launches_per_step": launch_mix}
    if parity_ok is not None:
        out["parity_ok"] = parity_ok
    out["peaks"] = {"hbm_gbs": pk["hbm_gbs"], "hbm_source": pk["source"], "tf32_tflops_burst": tf32_burst,
                    "tf32_tflops_sustained": tf32["tf32_tflops_sustained"], "tf32_how": tf32["how"],
                    "bf16_tflops_burst": pk["bf16_tflops"]}

    def tensor_roofline(kernel, entries, note, traffic_kernel=None):
        """entries: (event0, event1, algorithmic flops).  3xTF32 issues three tensor-pipe passes per flop."""
        t_ms = sum(a.elapsed_time(b) for a, b, _ in entries)
        fl = sum(f for _, _, f in entries)
        if t_ms <= 0:
            return None
        ach = 3.0 * fl / (t_ms / 1e3) / 1e12
        traffic, src = ncu_traffic(traffic_kernel or kernel)
        r = {"bound": "tensor", "achieved": ach, "peak": tf32_burst, "unit": "TFLOP/s", "frac": ach / tf32_burst,
             "traffic": traffic, "kernel": kernel, "algorithmic_tflops": fl / (t_ms / 1e3) / 1e12,
             "kernel_ms_per_step": t_ms / n_eager, "launches_per_step": len(entries) // n_eager,
             "kernel_share_of_step": (t_ms / n_eager) / step_ms,
             "frac_of_sustained": ach / tf32["tf32_tflops_sustained"], "note": note}
        if src:
            r["traffic_source"] = src
        return r

    def hbm_roofline(kernel, entries, note, traffic_kernel=None):
        """entries: (event0, event1, algorithmic bytes)."""
        t_ms = sum(a.elapsed_time(b) for a, b, _ in entries)
        by = sum(n for _, _, n in entries)
        if t_ms <= 0:
            return None
        gbs = by / (t_ms / 1e3) / 1e9
        traffic, src = ncu_traffic(traffic_kernel or kernel)
        r = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
             "traffic": traffic, "kernel": kernel, "kernel_ms_per_step": t_ms / n_eager,
             "launches_per_step": len(entries) // n_eager, "kernel_share_of_step": (t_ms / n_eager) / step_ms,
             "note": note}
        if src:
            r["traffic_source"] = src
        return r

    common = (f"CUDA events on the launching stream around every launch of {n_eager} steps re-run un-captured right "
              f"after the timed CUDA-graph region; tensor peak = TF32 dense rate measured in this run "
              f"({tf32['how']}: burst {tf32_burst:.0f}, sustained {tf32['tf32_tflops_sustained']:.0f} TFLOP/s); "
              f"isolated launches are judged against the burst figure")
    # the wide-tap launches split by size: a tap of >= 3 GFLOP keeps the machine busy for tens of microseconds;
    # the small ones (C = 256..512 at 14x14) are a few microseconds of work behind the same launch and pipeline ramp,
    # and an un-captured launch additionally waits for the host to submit it
    BIG = 3e9
    wide = [(a, b, f) for a, b, _, f, rows in direct_timer if rows > 128 and f >= BIG]
    wide_small = [(a, b, f) for a, b, _, f, rows in direct_timer if rows > 128 and f < BIG]
    narrow = [(a, b, n) for a, b, n, _, rows in direct_timer if rows <= 128]
    out["roofline"] = tensor_roofline(
        "gram_tma_kernel<2,128,128>", wide,
        "wide taps of >= 3 GFLOP (C >= 256, TMA-eligible; 92 % of the step's Gram FLOPs that take this kernel): achieved = "
        "3 x algorithmic FLOPs (2*C^2*K per tap) / summed launch time; " + common)
    if out["roofline"] is not None and wide_small:
        r = tensor_roofline("gram_tma_kernel<2,128,128>", wide_small,
                            "the same kernel on the wide taps below 3 GFLOP (launch / ramp / submission bound); " + common)
        out["roofline"]["small_taps"] = {k: r[k] for k in ("achieved", "frac", "algorithmic_tflops", "kernel_ms_per_step",
                                                           "launches_per_step")}
    packed = [(a, b, f) for a, b, f, _, _ in timer]
    if out["roofline"] is None:  # PLB_TMA_GRAM=0: the packed-plane GEMM is the dominant kernel again
        out["roofline"] = tensor_roofline("gemm3xtf32_v2_kernel", packed, "packed-plane 3xTF32 GEMM; " + common)
    else:
        r = tensor_roofline("gemm3xtf32_v2_kernel", packed,
                            "taps the tensor map cannot describe (7x7 planes: 196-byte channel stride; flattened [B, C] "
                            "taps) keep pack + packed-plane GEMM; " + common)
        if r:
            out["roofline_packed_gemm"] = r
    r = hbm_roofline("gram_tma_kernel<1,64,64>", narrow,
                     "narrow taps (C <= 128, arithmetic intensity C/4 flop/B): achieved = algorithmic bytes (both fp32 "
                     "activations of the tap, read once: 2*C*K*4) / summed launch time; " + common)
    if r:
        n_ms = sum(a.elapsed_time(b) for a, b, _, _, rows in direct_timer if rows <= 128)
        r["algorithmic_tflops"] = sum(f for _, _, _, f, rows in direct_timer if rows <= 128) / (n_ms / 1e3) / 1e12
        out["roofline_narrow_taps"] = r
    r = hbm_roofline("pack_split_pair_kernel", pack_timer,
                     "operand pack of the taps that stay on the packed path: 12 B per element (4 read + 8 written as tf32 "
                     "hi/lo planes); " + common)
    if r:
        out["roofline_pack"] = r
    if conv_timer:
        # every launch is judged against the roof that bounds it: tensor when 3 x flops / TF32 rate exceeds
        # bytes / HBM copy rate (the 3x3 layers, the 1x1 layers with >= 512 input channels), HBM otherwise
        def conv_bound(f, n):
            return "tensor" if 3.0 * f / (tf32_burst * 1e12) >= n / (pk["hbm_gbs"] * 1e9) else "hbm"

        conv_note = ("forward convolutions of BOTH source models, one launch per layer pair with the eval-mode "
                     "BatchNorm (+ ReLU) behind it as a second output (plb_conv2d_affine_forward): 3xTF32 implicit "
                     "GEMM, A operand split in registers and fed from TMEM, weights packed once + TMA; ")
        tens = [(a, b, f) for a, b, f, n, _ in conv_timer if conv_bound(f, n) == "tensor"]
        membound = [(a, b, n) for a, b, f, n, _ in conv_timer if conv_bound(f, n) == "hbm"]
        r = tensor_roofline(
            "conv3xtf32_kernel", tens,
            conv_note + "the launches whose tensor time exceeds their HBM time; achieved = 3 x algorithmic FLOPs "
            "(2*N*OH*OW*Cout*Cin*KH*KW per model) / summed launch time; " + common)
        if r:
            out["roofline_conv"] = r
        r = hbm_roofline(
            "conv3xtf32_kernel", membound,
            conv_note + "the launches whose HBM time exceeds their tensor time (1x1 layers with few input channels, the "
            "3-channel stem): achieved = algorithmic bytes (x read once, every output written once) / summed launch "
            "time; " + common, traffic_kernel="conv3xtf32_kernel (HBM-bound class)")
        if r:
            out["roofline_conv_hbm"] = r
        c_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in conv_timer)
        out["conv_ms_per_step"] = c_ms / n_eager
        # the line's `roofline` is the dominant kernel's: since the forwards moved onto this library's kernel, that is
        # whichever of the Gram kernel's wide taps and the convolution's tensor-bound launches takes more of the step
        if out.get("roofline_conv") and out.get("roofline") and \
                out["roofline_conv"]["kernel_ms_per_step"] > out["roofline"]["kernel_ms_per_step"]:
            out["roofline_gram"] = out["roofline"]
            out["roofline"] = out["roofline_conv"]
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
    ps = args.pleas_steps if args.pleas_steps is not None else (400 if K >= 100 else 4 * K)
    if not args.no_merge:
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

    if world == 1 and not args.no_extra_rooflines:
        # ---- K4: normal-equation GEMMs of one PLeaS batch (symmetric U^T U and U^T Y per layer), event-timed
        with torch.no_grad():
            blocks = P.get_blocks(spec, perm, costs, 0.0)
            pb = dict(blocks)
            for axis, pg in spec.items():
                for ax in pg.state:
                    pb[ax] = pb[axis]
            model3n = P.partial_merge(spec, m1, m2, perm, costs, 0.0)
            lr = PM.LstsqRunner(m1, m2, model3n, pb, num_classes_of(args.model), False, "rn50", use_cuda_graph=False)
            m1.eval()
            m2.eval()
            lr._eager(dev_batches[0])
            torch.cuda.synchronize()
            ops.GEMM_TIMER = []
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for i in range(3):
                lr._eager(dev_batches[(i + 1) % n_dev])
            ev1.record()
            torch.cuda.synchronize()
            k4, ops.GEMM_TIMER = ops.GEMM_TIMER, None
            lr.close()
        k4_ms = sum(e[0].elapsed_time(e[1]) for e in k4)
        k4_fl = sum(e[2] for e in k4)
        ach = 3.0 * k4_fl / (k4_ms / 1e3) / 1e12
        out["roofline_normal_eq"] = {
            "bound": "tensor", "achieved": ach, "peak": tf32_burst, "unit": "TFLOP/s", "frac": ach / tf32_burst,
            "traffic": None, "kernel": "gemm3xtf32_v2_kernel (U^T U symmetric + U^T Y per layer)",
            "algorithmic_tflops": k4_fl / (k4_ms / 1e3) / 1e12, "kernel_ms_per_step": k4_ms / 3,
            "launches_per_step": len(k4) // 3, "pleas_step_ms_eager": ev0.elapsed_time(ev1) / 3,
            "note": "PLeaS closed form, one batch: per trained layer the im2col pack of X-bar (fused gather-average), "
                    "G += U^T U and R += U^T Y-bar; achieved = 3 x algorithmic FLOPs (2*K_l^2*L_l for the full Gram although "
                    "only the tiles touching the lower triangle are computed, + 2*Co*K_l*L_l) / summed CUDA-event time of "
                    "the GEMM launches of 3 un-captured steps — the symmetric skip therefore shows as a fraction above what "
                    "the tensor pipe really sustains"}
        # ---- K2: batched LAP on the real cost matrices of this run
        mats = list(costs.values())
        ops.lap_solve_batched(mats, True)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            ops.lap_solve_batched(mats, True)
        ev1.record()
        torch.cuda.synchronize()
        lap_ms = ev0.elapsed_time(ev1) / 3
        lap_bytes = sum(m.numel() * 4 for m in mats)
        gbs = lap_bytes / (lap_ms / 1e3) / 1e9
        out["roofline_lap"] = {
            "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
            "traffic": None, "kernel": "lap_kernel_v3", "ms": lap_ms, "problems": len(mats),
            "sizes": sorted({m.shape[0] for m in mats}),
            "note": "all permutation groups of the pair in ONE launch (one CTA per problem), real -cdist cost matrices of "
                    "this run; algorithmic bytes = sum n^2 * 4 (each matrix read once).  The solve is a chain of dependent "
                    "augmenting-path steps (latency-bound by construction; ~125 k steps of ~0.75 us for the n = 2048 group, "
                    "instruction-issue bound at 32 warps: profiles/r02_notes.md): the fraction of the HBM rate is reported "
                    "because the contract asks for it, the time is what matters"}
        # ---- K5: blocked fp64 Cholesky + triangular solves, the largest layer shape of the pair
        n_ch, nrhs = 4608, 512
        A = torch.randn(n_ch, n_ch + 64, dtype=torch.float64, device=device)
        G0 = A @ A.T
        B0 = torch.randn(n_ch, nrhs, dtype=torch.float64, device=device)
        ops.chol_solve_(G0.clone(), B0.clone(), 1e-6)
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            Gc, Bc = G0.clone(), B0.clone()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            ops.chol_solve_(Gc, Bc, 1e-6)
            ev1.record()
            torch.cuda.synchronize()
            times.append(ev0.elapsed_time(ev1))
        ch_ms = min(times)
        ch_fl = n_ch ** 3 / 3.0 + 2.0 * n_ch * n_ch * nrhs
        out["roofline_chol"] = {
            "bound": "fp64 SIMT + launch latency", "achieved": ch_fl / (ch_ms / 1e3) / 1e12, "peak": None,
            "unit": "TFLOP/s", "frac": None, "traffic": None, "kernel": "potrf/trsm/rank_update/trsv (chol.cu)",
            "ms": ch_ms, "n": n_ch, "nrhs": nrhs,
            "note": "K = 4608 (layer4 3x3 convolutions), 512 right-hand sides: n^3/3 + 2 n^2 nrhs fp64 FLOPs / best of 3; "
                    "no measured fp64 peak exists for this pool (MEASURED_PEAKS.json has HBM and bf16 only), so no fraction "
                    "is claimed; all 54 solves of the pair are 0.3-0.4 s of the merge"}
        del A, G0, B0
        # ---- the reference's own GPU path on this B200 (second baseline of BASELINE.md section 3)
        try:
            out["gpu_library_baseline"] = gpu_library_baseline(args, spec, m1, m2, host)
        except Exception as e:  # never lose the headline to the baseline leg
            out["gpu_library_baseline"] = {"error": repr(e)}

    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
        out["cpu_baseline"] = cpu_baseline(args, spec)
        if not args.no_merge and args.cpu_merge_steps > 0:
            out["cpu_baseline_merge"] = cpu_baseline_merge(args, spec, perm, costs, Ke, ps + 1,
                                                           out["cpu_baseline"]["value"])
    if world > 1:
        dist.destroy_process_group()
    return json.dumps(out)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
