"""Partial weight matching — drop-in for the reference's experimental
``weight_matching_partial`` / ``apply_perm_with_padding`` / ``remove_zero_block``
(pleas/methods/partial_matching.py:260-463; no driver calls them, they are part of the published
module surface).

Like weight matching it is a coordinate ascent over the permutation groups in a seeded random
order, but a visit also decides which units of the group are MERGED (the best-matched
``merge = n - int(n (1 - ratio))`` pairs of the assignment) and which stay SEPARATE, and rewrites
both state dicts into the expanded layout

    model A:  [ separate (rep) | zeros (rep) | merged (merge) ]      rep = int(n (1 - ratio))
    model B:  [ zeros (rep) | separate (rep) | merged (merge) ]

along every state axis of the group, so that the two models can later be averaged unit by unit.
On the next visit the zero block is stripped again before the cost matrix is built.

Kernels: the cost build ``A = sum_ax W_a W_b^T`` goes through ``cross_weights`` (default: this
library's 3xTF32 tcgen05 Gram, ``cross_features_inner_product``) and the assignment through
``lsa_solver`` (default: the GPU LAP kernel, SciPy-identical); selection and rewriting are a
handful of gathers per tensor.  Both callables are the reference's plug points and may be replaced.

Reference behaviours kept on purpose: the merged set is ``matched cost > quantile(1 - ratio)``
patched to exactly ``merge`` units (the reference's fix-up order), ``state_as`` is rewritten IN
PLACE even when ``inplace=False`` (only ``state_bs`` is shallow-copied, :360-361), and the returned
permutation is the composition of the raw LAP results.
"""
from collections.abc import Sequence
from copy import copy, deepcopy
from typing import Union

import torch

from ..core.solvers import b200_solve_lsa
from ..core.utils import Permutation, PermutationSpec, StateDict, apply_perm, make_identity_perm
from .activation_matching import cross_features_inner_product
from .partial_matching import Ratios


def remove_zero_block(tensor, axis, block_size, final_size, beginning=False):
    """Strips the all-zero block of an expanded tensor (reference :260-275): at the beginning of
    ``axis`` for model B, second position for model A.  Tensors that are not expanded yet
    (extent < ``final_size``) pass through."""
    if tensor.shape[axis] < final_size:
        return tensor
    if axis not in (0, 1):
        raise NotImplementedError("remove_zero_block: only axes 0 and 1 carry permutation groups")
    if beginning:
        if float(tensor.narrow(axis, 0, block_size).sum()) != 0.0:
            raise AssertionError((final_size, beginning, block_size))
        return tensor.narrow(axis, block_size, tensor.shape[axis] - block_size)
    head = tensor.narrow(axis, 0, block_size)
    tail = tensor.narrow(axis, 2 * block_size, tensor.shape[axis] - 2 * block_size)
    return torch.cat([head, tail], dim=axis)


def apply_perm_with_padding(perm: Permutation, padding: Permutation, size: int, pad_ahead: bool,
                            spec: PermutationSpec, state: Union[torch.nn.Module, StateDict], inplace=False,
                            skip_missing=True):
    """Rewrites every state axis of the groups in ``perm`` as ``[zeros | separate | merged]``
    (``pad_ahead``) or ``[separate | zeros | merged]`` where ``merged = index_select(w, P)`` and
    ``separate = index_select(w, padding)`` (reference :278-334)."""
    if isinstance(state, torch.nn.Module):
        assert inplace
        state.load_state_dict(apply_perm(perm, spec, state.state_dict(), inplace=inplace))
        return state
    if not inplace:
        state = copy(state)
    for key, P in perm.items():
        if P is None:
            continue
        sep = padding[key]
        for ax in spec[key].state:
            if skip_missing and ax.key not in state:
                continue
            w = remove_zero_block(state[ax.key], ax.axis, len(sep), size, pad_ahead)
            merged = torch.index_select(w, ax.axis, P.to(w.device))
            separate = torch.index_select(w, ax.axis, sep.to(w.device))
            zeros = torch.zeros_like(separate)
            parts = (zeros, separate, merged) if pad_ahead else (separate, zeros, merged)
            out = torch.cat(parts, dim=ax.axis)
            if not float(out.abs().sum()) > 0.0:
                raise ValueError("Zero norm")
            state[ax.key] = out
    return state


def _merge_mask(matched, ratio, merge_size):
    """Boolean mask of the units to merge: matched cost above the (1 - ratio) quantile, then the
    reference's fix-ups (:421-431) until exactly ``merge_size`` are selected."""
    mask = matched > torch.quantile(matched, 1.0 - ratio)
    count = int(mask.sum())
    if count > merge_size:
        mask[matched.argmin()] = False
    elif count < merge_size:
        for i in range(len(mask)):
            if not bool(mask[i]):
                mask[i] = True
                if int(mask.sum()) == merge_size:
                    break
    return mask


def weight_matching_partial(spec: PermutationSpec, state_as: Union[StateDict, Sequence[StateDict]],
                            state_bs: Union[StateDict, Sequence[StateDict]], ratios: Ratios, max_iter=100,
                            init_perm=None, inplace=False, skip_suffixes=("running_mean", "running_var"),
                            skip_missing=True, lsa_solver=b200_solve_lsa,
                            cross_weights=cross_features_inner_product, verbose=True, seed=0) -> Permutation:
    """Same signature and result as the reference (:337-463): the composed permutation per group;
    the state dicts are rewritten into the expanded layout as a side effect."""
    if isinstance(state_as, dict):
        state_as = [state_as]
    if isinstance(state_bs, dict):
        state_bs = [state_bs]
    assert len(state_as) == len(state_bs)
    if not inplace:
        state_bs = [copy(sb) for sb in state_bs]
    perm = make_identity_perm(spec) if init_perm is None else deepcopy(init_perm)
    if init_perm is not None:
        for sb in state_bs:
            apply_perm(init_perm, spec, sb, inplace=True)
    names = list(perm.keys())
    device = next(iter(state_as[0].values())).device
    rng = torch.Generator()
    rng.manual_seed(seed)
    skip_suffixes = tuple(skip_suffixes)

    with torch.no_grad():
        for iteration in range(max_iter):
            progress = False
            for ix in torch.randperm(len(names), generator=rng):
                p = names[ix]
                pg, ratio = spec[p], ratios[p]
                n = pg.size
                rep = int(n * (1 - ratio))
                merge = int(n * ratio)
                if merge + rep < n:
                    merge += 1
                final = 2 * rep + merge
                A = torch.zeros(n, n, device=device)
                for ax in pg.state:
                    if ax.key.endswith(skip_suffixes):
                        continue
                    for sa, sb in zip(state_as, state_bs):
                        if skip_missing and not (ax.key in sa and ax.key in sb):
                            continue
                        w_a = remove_zero_block(sa[ax.key], ax.axis, rep, final, False)
                        w_b = remove_zero_block(sb[ax.key], ax.axis, rep, final, True)
                        A.add_(cross_weights(w_a.contiguous(), w_b.contiguous(), ax.axis))
                assert float(A.norm()) > 0
                newP = lsa_solver(A).to(A.device)
                rows = torch.arange(n, device=A.device)
                matched = A[rows, newP]
                old_l, new_l = A.diag().sum(), matched.sum()
                mask = _merge_mask(matched, ratio, merge)
                a_merged, b_merged = rows[mask], newP[mask]
                a_sep, b_sep = rows[~mask], newP[~mask]
                assert len(a_merged) == len(b_merged) == merge, (len(a_merged), len(b_merged), merge)
                assert len(a_sep) == len(b_sep) == rep, (len(a_sep), len(b_sep), rep)
                progress = progress or bool(new_l > old_l + 1e-12)
                if verbose:
                    print(f"{iteration}/{p.key}:{p.axis}: {float(new_l - old_l)}")
                perm[p] = perm[p][newP.to(perm[p].device)]
                for sb in state_bs:
                    apply_perm_with_padding({p: b_merged}, {p: b_sep}, final, True, spec, sb, inplace=True)
                for sa in state_as:
                    apply_perm_with_padding({p: a_merged}, {p: a_sep}, final, False, spec, sa, inplace=True)
            if not progress:
                break
    return perm
