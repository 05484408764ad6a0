"""Generate tests/golden/* from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run as

    PYTHONHASHSEED=0 python -m oracle.make_golden

PYTHONHASHSEED pins the iteration order of the reference's ``set[Axis]`` members
(SURVEY.md §3.2 step 4), which leaks into fp32 summation order.  The script imports
the reference through oracle/import_reference.py, runs its public hot-path functions
on seeded synthetic inputs and stores inputs-by-seed + outputs as small fixtures:

  lap_golden.npz      SciPy linear_sum_assignment (the routine behind
                      pleas/core/solvers.py:29-31) on known-answer / random / tie-heavy
                      / structured instances
  spec_*.json         get_permutation_spec (pleas/core/compiler.py:786-797) for the tiny
                      net, ResNet-18/50 (+ fc=Identity variant) and ResNet-101
  tiny_golden.pt      TinyResNet pair: per-tap cross features, costs and permutations
                      of activation_matching (verbatim "last batch" semantics and the
                      accumulate-fixed sum), weight_matching, get_blocks/partial_merge at
                      several ratios, reference Adam ``train`` losses and the fp64
                      least-squares optimum built from the reference's own
                      get_model_orig_activations
  rn18_golden.pt      ResNet-18 pair at 64x64: activation_matching permutations +
                      objectives (both cross-feature functions), weight_matching perms
  budget_golden.json  count_linear_flops terms and partial_merge_flops values (tiny, RN18, RN50)
"""
import json
import os
import sys
import time

import numpy as np
import torch

from oracle import tinynet
from oracle.import_reference import accumulate_fixed_costs, load_reference

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def spec_to_json(spec):
    return [
        {
            "key": [k.key, k.axis],
            "size": pg.size,
            "state": sorted([a.key, a.axis] for a in pg.state),
            "node": sorted([a.key, a.axis] for a in pg.node),
        }
        for k, pg in spec.items()
    ]


def gen_lap():
    from scipy.optimize import linear_sum_assignment as lsa

    rng = np.random.default_rng(0)
    inst = {}

    def add(name, A, maximize):
        A = np.ascontiguousarray(A, dtype=np.float32)
        r, c = lsa(A, maximize=maximize)
        assert (r == np.arange(len(r))).all()
        inst[f"{name}/A"] = A
        inst[f"{name}/maximize"] = np.array(maximize)
        inst[f"{name}/col"] = c.astype(np.int64)
        inst[f"{name}/obj"] = np.array(A.astype(np.float64)[r, c].sum())

    add("scipy_doc", [[4, 1, 3], [2, 0, 5], [3, 2, 2]], False)  # -> [1,0,2], cost 5
    add("n1", [[3.5]], True)
    add("n2", [[1, 2], [2, 1]], True)
    add("const5", np.ones((5, 5)), True)  # scipy: identity on constant matrices
    add("zeros7", np.zeros((7, 7)), True)
    for n in (3, 5, 17, 33, 64, 100, 128):
        add(f"randn{n}", rng.standard_normal((n, n)), True)
        add(f"randn{n}_min", rng.standard_normal((n, n)), False)
    for n in (6, 31, 64):  # integer costs with many ties (exercises the tie-break rule)
        add(f"ties{n}", rng.integers(0, 4, (n, n)), True)
        add(f"ties{n}_min", rng.integers(0, 3, (n, n)), False)
    for n in (48, 96):  # structured -cdist costs like activation matching produces
        X = rng.standard_normal((n, 256)).astype(np.float32)
        pi = rng.permutation(n)
        Y = X[pi] + 0.5 * rng.standard_normal((n, 256)).astype(np.float32)
        D = -np.sqrt(np.maximum((X * X).sum(1)[:, None] + (Y * Y).sum(1)[None] - 2 * X @ Y.T, 0))
        add(f"cdist{n}", D, True)
    # duplicate rows / dead channels: identical rows and an all-zero row+column
    A = rng.standard_normal((12, 12)).astype(np.float32)
    A[3] = A[7]
    A[5] = 0
    A[:, 2] = 0
    add("dup12", A, True)
    np.savez_compressed(os.path.join(GOLD, "lap_golden.npz"), **inst)
    print("lap_golden:", len(inst) // 4, "instances")


def gen_specs(ref):
    import torchvision

    out = {}
    torch.manual_seed(0)
    out["tiny"] = tinynet.TinyResNet(12, 10).eval()
    out["resnet18"] = torchvision.models.resnet18().eval()
    out["resnet50"] = torchvision.models.resnet50().eval()
    m = torchvision.models.resnet50().eval()
    m.fc = torch.nn.Identity()  # experiments/different_label_space/run_torchvision.py:189-191
    out["resnet50_nofc"] = m
    out["resnet101"] = torchvision.models.resnet101().eval()
    for name, model in out.items():
        shape = (1, 3, 16, 16) if name == "tiny" else (1, 3, 64, 64)
        spec = ref.compiler.get_permutation_spec(model, (shape,))
        with open(os.path.join(GOLD, f"spec_{name}.json"), "w") as f:
            json.dump(spec_to_json(spec), f, separators=(",", ":"))
        print(f"spec_{name}: {len(spec)} groups, {sum(len(pg.node) for pg in spec.values())} taps")


def axis_key(a):
    return f"{a.key}:{a.axis}"


def run_am(ref, spec, m1, m2, loader, cross_features, fixed):
    """activation_matching (pleas/methods/activation_matching.py:139-177); ``fixed`` swaps in
    the accumulate-fixed cost loop."""
    axes = [ax for pg in spec.values() for ax in pg.node]
    gm = ref.am.build_cross_module(m1, m2, axes, cross_features)
    if fixed:
        costs = accumulate_fixed_costs(ref, spec, gm, loader, len(loader))
    else:
        costs = ref.am.compute_matching_costs(spec, gm, loader, len(loader))
    perm = {k: ref.solvers.scipy_solve_lsa(v) for k, v in costs.items()}
    return perm, costs, gm


def gen_tiny(ref):
    m1, m2 = tinynet.make_pair(12, 10)
    loader = tinynet.make_loader(3, 4, 16)
    spec = ref.compiler.get_permutation_spec(m1, ((1, 3, 16, 16),))
    G = {"width": 12, "num_classes": 10, "loader": [3, 4, 16, 123]}

    for cf_name, cf in (("cdist", ref.am.cross_features_cdist), ("inner", ref.am.cross_features_inner_product)):
        for fixed in (False, True):
            perm, costs, gm = run_am(ref, spec, m1, m2, loader, cf, fixed)
            tag = f"am/{cf_name}/{'sum' if fixed else 'reference'}"
            G[f"{tag}/perm"] = {axis_key(k): v.clone() for k, v in perm.items()}
            G[f"{tag}/costs"] = {axis_key(k): v.clone() for k, v in costs.items()}
        # per-tap cross features of the LAST batch (activation -> cost boundary)
        with torch.inference_mode():
            _, cross = gm(loader[-1][0])
        G[f"am/{cf_name}/taps_last"] = {f"{k[0]}:{k[1]}": v.clone() for k, v in cross.items()}

    # weight matching (pleas/methods/weight_matching.py:22-95)
    calls = [0]

    def counting_solver(A, maximize=True):
        calls[0] += 1
        return ref.solvers.scipy_solve_lsa(A, maximize)

    perm_wm, costs_wm = ref.wm.weight_matching(
        spec, m1.state_dict(), m2.state_dict(), max_iter=100, seed=0, verbose=False,
        return_costs=True, lsa_solver=counting_solver,
    )
    G["wm/perm"] = {axis_key(k): v.clone() for k, v in perm_wm.items()}
    G["wm/costs"] = {axis_key(k): v.clone() for k, v in costs_wm.items()}
    G["wm/lap_calls"] = calls[0]

    # get_blocks / partial_merge (pleas/methods/partial_matching.py:47-202)
    perm = {k: G["am/cdist/sum/perm"][axis_key(k)] for k in spec}
    costs = {k: G["am/cdist/sum/costs"][axis_key(k)] for k in spec}
    keys = list(spec.keys())
    ratio_cases = {
        "r0": 0.0,
        "r05": 0.5,
        "r1": 1.0,
        "mixed": {k: [0.0, 0.3, 1.0, 0.7][i % 4] for i, k in enumerate(keys)},
    }
    for name, ratios in ratio_cases.items():
        model3, blocks = ref.pm.partial_merge(spec, m1, m2, perm, costs, ratios, return_blocks=True)
        G[f"pm/{name}/ratios"] = ratios if not isinstance(ratios, dict) else {axis_key(k): v for k, v in ratios.items()}
        G[f"pm/{name}/blocks"] = {axis_key(k): [t.clone() for t in v] for k, v in blocks.items()}
        G[f"pm/{name}/state"] = {k: v.detach().clone() for k, v in model3.state_dict().items()}

    # PLeaS train (pleas/methods/pleas_merging.py:305-405): reference Adam vs fp64 optimum
    tloader = tinynet.make_loader(24, 4, 16, seed=321)
    for name in ("r0", "r05"):
        ratios = ratio_cases[name]
        model3, blocks = ref.pm.partial_merge(spec, m1, m2, perm, costs, ratios, return_blocks=True)
        init_state = {k: v.detach().clone() for k, v in model3.state_dict().items()}
        steps = 23  # reference consumes MAX_STEPS+1 batches (idx > MAX_STEPS breaks)
        t0 = time.time()
        model3 = ref.pl.train(tloader, m1, m2, model3, spec, perm, costs, ratios, False, steps, None,
                              num_classes=10, model_type="rn18")
        print(f"reference train {name}: {time.time() - t0:.1f}s")
        G[f"train/{name}/adam_state"] = {k: v.detach().clone() for k, v in model3.state_dict().items()}
        G[f"train/{name}/layer_stats"] = layer_losses(ref, spec, m1, m2, blocks, init_state,
                                                      model3.state_dict(), tloader, 10)
    G["train/loader"] = [24, 4, 16, 321]
    G["train/max_steps"] = 23
    torch.save(G, os.path.join(GOLD, "tiny_golden.pt"))
    print("tiny_golden.pt written:", os.path.getsize(os.path.join(GOLD, "tiny_golden.pt")) // 1024, "KiB")


def layer_losses(ref, spec, m1, m2, blocks, init_state, adam_state, loader, num_classes):
    """Per trained layer: mean squared error of the merged layer on the reference's own
    (X-bar, Y-bar) pairs (pleas_merging.py:63-149, 281-283) with the partial_merge init, the
    reference's Adam result, and the fp64 least-squares optimum (normal equations built with
    F.unfold from the same pairs; masked entries per get_gradient_mask stay at their init)."""
    import torch.nn.functional as F

    perm_blocks = dict(blocks)
    for axis, pg in spec.items():
        for ax in pg.state:
            perm_blocks[ax] = perm_blocks[axis]
    a1, a2 = {}, {}
    h1 = ref.pl.capture_inputs(m1, a1)
    h2 = ref.pl.capture_inputs(m2, a2)
    layers = {n: mod for n, mod in m1.named_modules() if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear))}
    acc = {}
    with torch.no_grad():
        for x, _ in loader:
            m1(x)
            m2(x)
            snap1, snap2 = dict(a1), dict(a2)
            for name, mod in layers.items():
                X, Y = ref.pl.get_model_orig_activations(m1, m2, perm_blocks, name, snap1, snap2,
                                                         num_classes=num_classes, model_type="rn18")
                X, Y = X.double(), Y.double()
                if isinstance(mod, torch.nn.Conv2d):
                    U = F.unfold(X, mod.kernel_size, mod.dilation, mod.padding, mod.stride)  # [B, K, L]
                    U = U.transpose(1, 2).reshape(-1, U.shape[1])
                    T = Y.flatten(2).transpose(1, 2).reshape(-1, Y.shape[1])
                else:
                    U, T = X, Y
                if mod.bias is not None:
                    U = torch.cat([U, torch.ones(U.shape[0], 1, dtype=U.dtype)], 1)
                d = acc.setdefault(name, {"G": 0, "R": 0, "yy": 0, "n": 0})
                d["G"] = d["G"] + U.T @ U
                d["R"] = d["R"] + U.T @ T
                d["yy"] = d["yy"] + (T * T).sum(0)
                d["n"] += T.shape[0]
    for h in h1 + h2:
        h.remove()

    def flat_w(state, name, mod):
        W = state[f"{name}.weight"].double().flatten(1)
        if mod.bias is not None:
            W = torch.cat([W, state[f"{name}.bias"].double()[:, None]], 1)
        return W  # [Co', K']

    def loss(W, d):
        # sum_o (w_o^T G w_o - 2 w_o^T r_o + yy_o) / (n * Co)
        q = ((W @ d["G"]) * W).sum(1) - 2 * (W * d["R"].T).sum(1) + d["yy"]
        return float(q.sum() / (d["n"] * W.shape[0]))

    model3_dict = {n: mod for n, mod in layers.items()}
    stats = {}
    # gradient masks exactly as the reference builds them (pleas_merging.py:11-60)
    class _L:  # minimal object exposing .parameters() with merged shapes
        def __init__(self, ps):
            self.ps = ps

        def parameters(self):
            return self.ps

    mw = {}
    for name, mod in layers.items():
        ps = [init_state[f"{name}.weight"]]
        if mod.bias is not None:
            ps.append(init_state[f"{name}.bias"])
        mw[name] = _L(ps)
    masks = ref.pl.get_gradient_mask(perm_blocks, mw)
    mi = 0
    for name, mod in model3_dict.items():
        d = acc[name]
        W0 = flat_w(init_state, name, mod)
        Wa = flat_w(adam_state, name, mod)
        mask_w = masks[mi].double().flatten(1)
        mi += 1
        if mod.bias is not None:
            mask_w = torch.cat([mask_w, masks[mi].double()[:, None]], 1)
            mi += 1
        Wls = W0.clone()
        for o in range(W0.shape[0]):
            free = mask_w[o] > 0
            if free.sum() == 0:
                continue
            # minimum-norm UPDATE from the init (directions U never excites keep their init
            # value, exactly like a gradient method started at W0): dW = pinv(Gff) (r - G w0)
            Gff = d["G"][free][:, free]
            grad = d["R"][free, o] - d["G"][free] @ W0[o]
            ev, V = torch.linalg.eigh(Gff)
            inv = torch.where(ev > 1e-11 * ev.max(), 1.0 / ev, torch.zeros_like(ev))
            Wls[o, free] = W0[o, free] + V @ (inv * (V.T @ grad))
        stats[name] = {
            "loss_init": loss(W0, d), "loss_adam": loss(Wa, d), "loss_lstsq": loss(Wls, d),
            "masked": int((mask_w == 0).sum()),
        }
        print(f"  {name}: init {stats[name]['loss_init']:.5f} adam {stats[name]['loss_adam']:.5f} "
              f"lstsq {stats[name]['loss_lstsq']:.5f} masked {stats[name]['masked']}")
    assert mi == len(masks)
    return stats


def gen_rn18(ref):
    import torchvision

    torch.manual_seed(0)
    m1 = torchvision.models.resnet18().eval()
    torch.manual_seed(1)
    m2 = torchvision.models.resnet18().eval()
    g = torch.Generator().manual_seed(123)
    loader = [(torch.randn(4, 3, 64, 64, generator=g), 0) for _ in range(2)]
    spec = ref.compiler.get_permutation_spec(m1, ((1, 3, 64, 64),))
    G = {"loader": [2, 4, 64, 123]}
    for cf_name, cf in (("cdist", ref.am.cross_features_cdist), ("inner", ref.am.cross_features_inner_product)):
        for fixed in (False, True):
            perm, costs, _ = run_am(ref, spec, m1, m2, loader, cf, fixed)
            tag = f"am/{cf_name}/{'sum' if fixed else 'reference'}"
            G[f"{tag}/perm"] = {axis_key(k): v.to(torch.int16) for k, v in perm.items()}
            G[f"{tag}/obj"] = {
                axis_key(k): float(costs[k].double()[torch.arange(len(perm[k])), perm[k]].sum()) for k in perm
            }
            G[f"{tag}/cost_absmax"] = {axis_key(k): float(v.abs().max()) for k, v in costs.items()}
    calls = [0]

    def counting_solver(A, maximize=True):
        calls[0] += 1
        return ref.solvers.scipy_solve_lsa(A, maximize)

    perm_wm = ref.wm.weight_matching(spec, m1.state_dict(), m2.state_dict(), max_iter=100, seed=0,
                                     verbose=False, lsa_solver=counting_solver)
    G["wm/perm"] = {axis_key(k): v.to(torch.int16) for k, v in perm_wm.items()}
    G["wm/lap_calls"] = calls[0]
    torch.save(G, os.path.join(GOLD, "rn18_golden.pt"))
    print("rn18_golden.pt written:", os.path.getsize(os.path.join(GOLD, "rn18_golden.pt")) // 1024, "KiB",
          "wm lap calls", calls[0])


def gen_budget(ref):
    """count_linear_flops (pleas/core/utils.py:558-617) and partial_merge_flops
    (pleas/methods/partial_matching.py:205-226) on three models -> budget_golden.json."""
    import torchvision

    out = {}
    for name in ("tiny", "resnet18", "resnet50"):
        torch.manual_seed(0)
        if name == "tiny":
            m, shape = tinynet.TinyResNet(12, 10).eval(), (1, 3, 16, 16)
        else:
            m, shape = getattr(torchvision.models, name)().eval(), (1, 3, 64, 64)
        spec = ref.compiler.get_permutation_spec(m, (shape,))
        flops, terms = ref.utils.count_linear_flops(spec, m, (shape,))
        keys = list(spec.keys())
        cases = {"r0": 0.0, "r1": 1.0, "r03": 0.3,
                 "mixed": {k: [0.0, 0.25, 1.0, 0.6][i % 4] for i, k in enumerate(keys)}}
        out[name] = {
            "flops": int(flops),
            "terms": [[int(c)] + [[a.key, a.axis] for a in axes] for c, *axes in terms],
            "merge_flops": {cn: float(ref.pm.partial_merge_flops(spec, terms, r)) for cn, r in cases.items()},
            "mixed_ratios": [[k.key, k.axis, cases["mixed"][k]] for k in keys],
        }
    with open(os.path.join(GOLD, "budget_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("budget_golden:", {k: v["flops"] for k, v in out.items()})


def gen_eval(ref):
    """get_fc_perm / permute_final_features (pleas/methods/pleas_merging.py:408-465) on the tiny
    pair with the activation-matching result stored in tiny_golden.pt -> eval_golden.pt."""
    m1, _ = tinynet.make_pair(12, 10)
    spec = ref.compiler.get_permutation_spec(m1, ((1, 3, 16, 16),))
    T = torch.load(os.path.join(GOLD, "tiny_golden.pt"))
    perm = {k: T["am/cdist/sum/perm"][axis_key(k)] for k in spec}
    costs = {k: T["am/cdist/sum/costs"][axis_key(k)] for k in spec}
    keys = list(spec.keys())
    cases = {"r0": 0.0, "r05": 0.5, "r1": 1.0, "mixed": {k: [0.3, 0.0, 0.7, 1.0][i % 4] for i, k in enumerate(keys)}}
    G = {}
    gen = torch.Generator().manual_seed(77)
    for name, ratios in cases.items():
        fc_perm = ref.pl.get_fc_perm(perm, spec, costs, ratios)
        n, m = len(fc_perm[0]), len(fc_perm[2])
        feats = torch.randn(5, n + 2 * m, generator=gen)
        G[f"{name}/ratios"] = ratios if not isinstance(ratios, dict) else {axis_key(k): v for k, v in ratios.items()}
        G[f"{name}/fc_perm"] = [t.clone() for t in fc_perm]
        G[f"{name}/features"] = feats
        G[f"{name}/out0"] = ref.pl.permute_final_features(feats, fc_perm, 0).clone()
        G[f"{name}/out1"] = ref.pl.permute_final_features(feats, fc_perm, 1).clone()
    torch.save(G, os.path.join(GOLD, "eval_golden.pt"))
    print("eval_golden.pt written:", {k: [len(t) for t in v] for k, v in G.items() if k.endswith("fc_perm")})


def gen_wmp(ref):
    """weight_matching_partial (pleas/methods/partial_matching.py:337-463) on the tiny pair:
    returned permutation and both rewritten state dicts (inplace=True) -> wmp_golden.pt."""
    G = {}
    m1, m2 = tinynet.make_pair(12, 10)
    spec = ref.compiler.get_permutation_spec(m1, ((1, 3, 16, 16),))
    keys = list(spec.keys())
    cases = {"r05": {k: 0.5 for k in keys}, "r025": {k: 0.25 for k in keys},
             "mixed": {k: [0.5, 1.0, 0.25, 0.75][i % 4] for i, k in enumerate(keys)}}
    for name, ratios in cases.items():
        sa = {k: v.clone() for k, v in m1.state_dict().items()}
        sb = {k: v.clone() for k, v in m2.state_dict().items()}
        perm = ref.pm.weight_matching_partial(spec, sa, sb, ratios, max_iter=20, inplace=True, verbose=False, seed=0)
        G[f"{name}/ratios"] = {axis_key(k): v for k, v in ratios.items()}
        G[f"{name}/perm"] = {axis_key(k): v.clone() for k, v in perm.items()}
        G[f"{name}/state_a"] = {k: v.clone() for k, v in sa.items()}
        G[f"{name}/state_b"] = {k: v.clone() for k, v in sb.items()}
        print("wmp", name, {k: tuple(v.shape) for k, v in list(sa.items())[:2]})
    torch.save(G, os.path.join(GOLD, "wmp_golden.pt"))


def gen_api(ref):
    """Positional parameter names and plain default values of the reference's public functions on the
    path (the drop-in contract) -> api_signatures.json."""
    import inspect

    funcs = {
        "get_permutation_spec": ref.compiler.get_permutation_spec,
        "activation_matching": ref.am.activation_matching, "cross_features_cdist": ref.am.cross_features_cdist,
        "cross_features_inner_product": ref.am.cross_features_inner_product,
        "build_cross_module": ref.am.build_cross_module, "compute_matching_costs": ref.am.compute_matching_costs,
        "weight_matching": ref.wm.weight_matching, "partial_merge": ref.pm.partial_merge,
        "get_blocks": ref.pm.get_blocks, "expand_ratios": ref.pm.expand_ratios,
        "build_partial_merge_model": ref.pm.build_partial_merge_model,
        "partial_merge_flops": ref.pm.partial_merge_flops, "qp_ratios": ref.pm.qp_ratios,
        "weight_matching_partial": ref.pm.weight_matching_partial,
        "apply_perm_with_padding": ref.pm.apply_perm_with_padding, "remove_zero_block": ref.pm.remove_zero_block,
        "train": ref.pl.train, "get_fc_perm": ref.pl.get_fc_perm,
        "permute_final_features": ref.pl.permute_final_features, "eval_perm_model": ref.pl.eval_perm_model,
        "eval_whole_model": ref.pl.eval_whole_model, "train_eval_linear_probe": ref.pl.train_eval_linear_probe,
        "scipy_solve_lsa": ref.solvers.scipy_solve_lsa, "apply_perm": ref.utils.apply_perm,
        "count_linear_flops": ref.utils.count_linear_flops,
    }
    out = {}
    for name, fn in funcs.items():
        params = []
        for prm in inspect.signature(fn).parameters.values():
            d = prm.default
            if d is inspect.Parameter.empty:
                params.append([prm.name, "<required>"])
            elif isinstance(d, (int, float, str, bool, type(None))):
                params.append([prm.name, d])
            elif isinstance(d, tuple):
                params.append([prm.name, list(d)])
            else:
                params.append([prm.name, "<callable>"])
        out[name] = params
    with open(os.path.join(GOLD, "api_signatures.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("api_signatures:", len(out), "functions")


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":
        sys.exit("run with PYTHONHASHSEED=0 (pins the reference's set iteration order)")
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = load_reference()
    which = sys.argv[1:] or ["lap", "specs", "tiny", "rn18", "budget", "eval", "wmp", "api"]
    if "lap" in which:
        gen_lap()
    if "specs" in which:
        gen_specs(ref)
    if "tiny" in which:
        gen_tiny(ref)
    if "rn18" in which:
        gen_rn18(ref)
    if "budget" in which:
        gen_budget(ref)
    if "eval" in which:
        gen_eval(ref)
    if "wmp" in which:
        gen_wmp(ref)
    if "api" in which:
        gen_api(ref)


if __name__ == "__main__":
    main()
