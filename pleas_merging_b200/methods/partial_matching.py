"""Budgeted partial merging on B200 (drop-in for pleas/methods/partial_matching.py:30-202).

``get_blocks`` picks, per permutation group, which matched unit pairs are merged and which
are kept separate (quantile of the matched costs); ``build_partial_merge_model`` assembles the
block-structured weights ``[merged | model-1 only | model-2 only]``.  Both run as kernels
(csrc/blocks.cu): sort + torch.quantile-compatible threshold + order-preserving compaction,
and a single gather/average/scatter pass per tensor — the reference does the assembly with
fancy indexing into CPU tensors.  Index results are bit-identical to the reference's.

Note: as in the reference, no assignment problem is solved here; the
``lsa_solver`` / ``zero_augmented`` arguments are accepted and ignored.
"""
from copy import copy, deepcopy
from typing import Dict, Union

import torch
from torch.nn import Module

from .. import ops
from ..core.solvers import b200_solve_lsa
from ..core.utils import Axis, Permutation, PermutationSpec, set_attr

Ratios = Union[float, Dict[Axis, float]]


def expand_ratios(spec: PermutationSpec, ratios: Ratios) -> Dict[Axis, float]:
    """reference :30-44"""
    if not isinstance(ratios, dict):
        return {ax: ratios for ax in spec}
    return ratios


def _device_of(costs):
    for c in costs.values():
        if torch.is_tensor(c) and c.is_cuda:
            return c.device
    return torch.device("cuda", torch.cuda.current_device())


def get_blocks(spec: PermutationSpec, perm: Permutation, costs, ratios: Ratios, lsa_solver=b200_solve_lsa):
    """{group key: (Q[mask], P[mask], Q[~mask], P[~mask])} as int64 CUDA tensors (reference
    :47-89): ``mask = c >= torch.quantile(c, ratio)`` over the matched costs ``c_i = C[i, P_i]``;
    a ratio within 1e-3 of 1.0 forces the identity permutation."""
    ratios = expand_ratios(spec, ratios)
    device = _device_of(costs)
    # every group's kernel is enqueued first; ONE readback of all merged-unit counts then sizes the views
    axes = list(perm.keys())
    sizes = [int(perm[a].numel()) for a in axes]
    bufs = torch.empty(4, max(sum(sizes), 1), dtype=torch.int64, device=device)
    counts = torch.zeros(max(len(axes), 1), dtype=torch.int32, device=device)
    perms_dev = torch.cat([perm[a].reshape(-1).to(torch.int64) for a in axes]).to(device) if axes else None
    off, spans = 0, []
    for gi, (axis, n) in enumerate(zip(axes, sizes)):
        r = float(ratios[axis])
        C = costs[axis].to(device=device, dtype=torch.float32)
        ops.get_blocks_launch(C, perms_dev[off:off + n], r, abs(r - 1.0) < 1e-3, bufs[:, off:off + n],
                              counts[gi:gi + 1])
        spans.append((off, n))
        off += n
    merged = counts.cpu().tolist()
    blocks = {}
    for axis, (o, n), m in zip(axes, spans, merged):
        blocks[axis] = (bufs[0, o:o + m], bufs[1, o:o + m], bufs[2, o:o + n - m], bufs[3, o:o + n - m])
    return blocks


def build_partial_merge_model(spec: PermutationSpec, model1: Module, model2: Module, blocks) -> Module:
    """reference :91-185.  The merged module is a deep copy of ``model1`` in eval mode whose
    blocked tensors are replaced by frozen Parameters (BatchNorm running statistics included,
    exactly like the reference).  It stays on the models' CUDA device (the reference moves it to
    the CPU and every caller moves it back)."""
    blocks = copy(blocks)
    for axis, pg in spec.items():
        for ax in pg.state:
            blocks[ax] = blocks[axis]
    axes_by_tensor = {}
    for pg in spec.values():
        for ax in pg.state:
            axes_by_tensor.setdefault(ax.key, set()).add(ax.axis)

    D1, D2, D3 = model1.state_dict(), model2.state_dict(), {}
    device = next(iter(model1.parameters())).device
    if device.type != "cuda":
        device = torch.device("cuda", torch.cuda.current_device())
    for name, axes in axes_by_tensor.items():
        if name not in D1 or name not in D2:
            print(f"Could not find - {name}")
            continue
        assert len(axes) in {1, 2}
        if len(axes) == 2:
            assert axes == {0, 1}  # axis 0 = output units, axis 1 = input units
        W1 = D1[name].to(device=device, dtype=torch.float32)
        W2 = D2[name].to(device=device, dtype=torch.float32)
        D3[name] = ops.block_merge(W1, W2, {ax: blocks[Axis(name, ax)] for ax in axes})

    model3 = deepcopy(model1).eval().to(device)
    for name, tensor in D3.items():
        set_attr(model3, name.split("."), torch.nn.Parameter(tensor, requires_grad=False))
    return model3


def partial_merge(spec: PermutationSpec, model1: Module, model2: Module, perm: Permutation, costs,
                  ratios: Ratios, zero_augmented: bool = False, return_blocks=False):
    """reference :188-202"""
    blocks = get_blocks(spec, perm, costs, ratios, zero_augmented)
    model3 = build_partial_merge_model(spec, model1, model2, blocks)
    if return_blocks:
        return model3, blocks
    return model3
