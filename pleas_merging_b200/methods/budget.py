"""FLOP accounting and the layer-wise ("zip") budget rule for partial merging.

Host-side bookkeeping only.  ``count_linear_flops`` / ``partial_merge_flops`` follow
pleas/core/utils.py:558-617 and pleas/methods/partial_matching.py:205-226; ``get_zip_ratios``
is the drivers' rule (experiments/different_label_space/run_torchvision.py:32-54) restated on
``Axis.key`` — as written there it calls ``str.startswith`` on ``Axis`` keys and cannot run.
``qp_ratios`` keeps the reference's signature and objective (partial_matching.py:229-257) but
solves the small bilinear program on the host without Gurobi (which is licence-gated and absent
here, so this function has NO reference output to pin against: parity unpinned; its tests check
feasibility, KKT structure and agreement with brute force / SciPy's SLSQP on small instances).
"""
import math
from typing import Dict, List, Tuple

import torch
import torch.fx
from torch import nn
from torch.fx.passes.shape_prop import ShapeProp

from ..core.utils import Axis, InputsOrShapes, PermutationSpec
from .partial_matching import Ratios, expand_ratios

Term = Tuple  # (coefficient, Axis) or (coefficient, Axis, Axis): FLOPs = coeff * prod(group sizes)


def count_linear_flops(spec: PermutationSpec, model: nn.Module, inputs_or_shapes: InputsOrShapes):
    """(MACs of all Conv2d/Linear layers, terms) where each term expresses a layer's MACs as
    coefficient x sizes of the permutation groups its weight axes belong to."""
    device = next(iter(model.parameters())).device
    inputs = [torch.randn(*ios).to(device) if isinstance(ios, tuple) else ios.to(device) for ios in inputs_or_shapes]
    gm = torch.fx.symbolic_trace(model)
    sp = ShapeProp(gm)
    sp.propagate(*inputs)
    group_of = {ax: k for k, pg in spec.items() for ax in pg.state}
    flops, terms = 0, []
    for node in gm.graph.nodes:
        if node.op != "call_module":
            continue
        mod = gm.get_submodule(node.target)
        if not isinstance(mod, (nn.Conv2d, nn.Linear)):
            continue
        shape = node.meta["tensor_meta"].shape
        coeff = shape[0]
        if isinstance(mod, nn.Conv2d):
            coeff *= math.prod(shape[2:]) * math.prod(mod.kernel_size)
        sout, sin = mod.weight.shape[:2]
        flops += coeff * sin * sout
        axes = []
        for ax, size in ((Axis(f"{node.target}.weight", 1), sin), (Axis(f"{node.target}.weight", 0), sout)):
            if ax in group_of:
                axes.append(group_of[ax])
            else:
                coeff *= size
        terms.append((coeff, *axes))
    return flops, terms


def partial_merge_flops(spec: PermutationSpec, terms: List[Term], ratios: Ratios) -> float:
    """MACs of the partially merged model: a group with ratio r has (1 + r) x its units, and the
    model-1-only x model-2-only blocks of a two-axis layer are empty."""
    ratios = expand_ratios(spec, ratios)
    total = 0
    for coeff, *axes in terms:
        if len(axes) == 0:
            total += coeff
        elif len(axes) == 1:
            total += coeff * spec[axes[0]].size * (1 + ratios[axes[0]])
        else:
            a, b = axes
            r1, r2 = ratios[a], ratios[b]
            total += coeff * spec[a].size * spec[b].size * ((1 + r1) * (1 + r2) - 2 * r1 * r2)
    return total


def get_zip_ratios(spec: PermutationSpec, budget_ratio: float, base_budget_ratios) -> Dict[Axis, float]:
    """Layer-wise rule: with i the position of ``budget_ratio`` in ``base_budget_ratios`` (five
    entries, e.g. merge_configs.BUDGET_RATIOS['rn50']), groups of ``layer1..layer{4-i}`` are merged
    (ratio 0), deeper ``layer*`` groups are kept separate (ratio 1) and all other groups merged."""
    depth = {b: 4 - i for i, b in enumerate(base_budget_ratios)}[budget_ratio]
    out = {}
    for k in spec:
        if k.key.startswith("layer"):
            layer = int(k.key.split(".")[0][len("layer"):])
            out[k] = 0.0 if layer <= depth else 1.0
        else:
            out[k] = 0.0
    return out


def _flops_poly(spec, terms, keys):
    """f(r) = const + lin . r + sum_{a<=b} quad[a, b] r_a r_b over the spec's groups (index order
    ``keys``), from ``partial_merge_flops``'s per-term formulas."""
    idx = {k: i for i, k in enumerate(keys)}
    m = len(keys)
    const, lin, quad = 0.0, [0.0] * m, {}
    for coeff, *axes in terms:
        if len(axes) == 0:
            const += coeff
        elif len(axes) == 1:
            a = idx[axes[0]]
            base = coeff * spec[axes[0]].size
            const += base
            lin[a] += base
        else:
            a, b = idx[axes[0]], idx[axes[1]]
            base = coeff * spec[axes[0]].size * spec[axes[1]].size
            const += base  # (1 + ra)(1 + rb) - 2 ra rb = 1 + ra + rb - ra rb
            lin[a] += base
            lin[b] += base
            key = (min(a, b), max(a, b))
            quad[key] = quad.get(key, 0.0) - base
    return const, lin, quad


def qp_ratios(spec: PermutationSpec, terms: List[Term], flops_budget: float, obj_weights: Dict[Axis, float],
              iters: int = 200) -> Dict[str, float]:
    """Per-group ratios maximising ``sum_k r_k max(w_k, 1e-5)`` subject to
    ``partial_merge_flops(r) / partial_merge_flops(0) <= flops_budget`` and ``0 <= r <= 1``
    (reference partial_matching.py:229-257, same return type: ``{key string: ratio}``).

    The program is linear in every single ratio, so at a KKT point each group is at 0, at 1, or
    (at most a few) in between with ``w_k = lambda * df/dr_k``.  Solved by successive linear
    programming: linearise the FLOP constraint at the current point, solve the resulting
    fractional knapsack exactly (greedy by w / gradient), move there with a damped step and
    restore the exact constraint by shrinking toward the origin (bisection); best of three starts."""
    keys = list(spec.keys())
    m = len(keys)
    w = [max(float(obj_weights[k]), 1e-5) for k in keys]
    const, lin, quad = _flops_poly(spec, terms, keys)
    budget = flops_budget * const  # f(0) = const

    def f(r):
        v = const + sum(l * x for l, x in zip(lin, r))
        for (a, b), q in quad.items():
            v += q * r[a] * r[b]
        return v

    def grad(r):
        g = list(lin)
        for (a, b), q in quad.items():
            if a == b:
                g[a] += 2 * q * r[a]
            else:
                g[a] += q * r[b]
                g[b] += q * r[a]
        return g

    if f([1.0] * m) <= budget:
        return {k.key: 1.0 for k in keys}
    if budget <= const:
        return {k.key: 0.0 for k in keys}

    def knapsack(r, g):
        """max w.x  s.t.  f(r) + g.(x - r) <= budget, 0 <= x <= 1 (g > 0: more units cost more)."""
        room = budget - f(r) + sum(gi * ri for gi, ri in zip(g, r))
        x = [0.0] * m
        for i in sorted(range(m), key=lambda i: -w[i] / max(g[i], 1e-30)):
            if g[i] <= 0:
                x[i] = 1.0
                continue
            take = min(1.0, max(0.0, room / g[i]))
            x[i] = take
            room -= take * g[i]
            if room <= 0:
                break
        return x

    def shrink(y):
        """Largest s in [0, 1] with f(s y) <= budget, applied (f is non-decreasing in every ratio:
        df/dr_k >= 0 on the box, so it is monotone along the ray from the origin)."""
        if f(y) <= budget:
            return y
        a, b = 0.0, 1.0
        for _ in range(60):
            t = 0.5 * (a + b)
            if f([t * v for v in y]) <= budget:
                a = t
            else:
                b = t
        return [a * v for v in y]

    best, best_val = [0.0] * m, 0.0
    for start in ([0.0] * m, [0.5] * m, [1.0] * m):
        r = shrink(start)
        step = 1.0
        for _ in range(iters):
            x = knapsack(r, grad(r))
            # the knapsack vertex sits on the LINEARISED constraint, so the move is tangent to the true
            # one and curvature can push it outside: restore feasibility by shrinking toward the origin
            # (a second-order loss in the objective against the first-order gain of the move)
            cand = shrink([ri + step * (xi - ri) for ri, xi in zip(r, x)])
            if sum(wi * ci for wi, ci in zip(w, cand)) > sum(wi * ri for wi, ri in zip(w, r)) * (1 + 1e-9) + 1e-12:
                r = cand
            else:
                step *= 0.5
                if step < 1e-5:
                    break
        val = sum(wi * ri for wi, ri in zip(w, r))
        if val > best_val:
            best, best_val = r, val
    return {k.key: max(min(float(v), 1.0), 0.0) for k, v in zip(keys, best)}
