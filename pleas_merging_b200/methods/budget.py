"""FLOP accounting and the layer-wise ("zip") budget rule for partial merging.

Host-side bookkeeping only.  ``count_linear_flops`` / ``partial_merge_flops`` follow
pleas/core/utils.py:558-617 and pleas/methods/partial_matching.py:205-226; ``get_zip_ratios``
is the drivers' rule (experiments/different_label_space/run_torchvision.py:32-54) restated on
``Axis.key`` — as written there it calls ``str.startswith`` on ``Axis`` keys and cannot run.
The reference's ``qp_ratios`` (a non-convex QP solved by Gurobi on the authors' private
sensitivity files) is out of scope.
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
