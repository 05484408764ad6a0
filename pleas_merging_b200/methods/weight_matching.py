"""Weight matching (Git Re-Basin coordinate ascent) on B200 — drop-in for
pleas/methods/weight_matching.py:22-95.

Per visited group the reference builds ``A = sum_ax W_a^(ax) W_b^(ax)^T`` with one
transposed copy + SGEMM + add per state axis, syncs on ``A.norm()``, copies A to the host for
SciPy, and rewrites every tensor of the group with ``index_select``.  Here a visit is: pack
the group's model-B operands K-concatenated into one plane pair (the model-A planes are
packed once per group), ONE 3xTF32 tcgen05 GEMM with a fused split-K reduction, the GPU LAP
kernel on the device-resident cost matrix, a device-side progress test, and gather kernels
for the permutation — the only host round trip is one 4-byte progress flag per sweep.  The
visit order replays ``torch.randperm`` on a seeded CPU generator exactly like the reference;
visits that do not share a tensor commute, so each sweep is scheduled by the levels of its
conflict DAG and the visits of a level are solved in ONE batched LAP launch (the LAPs are 96 %
of the time on a ResNet-50 pair) with results identical to the sequential order.
"""
from collections.abc import Sequence
from copy import copy, deepcopy
from typing import Union

import torch

from .. import ops
from ..core.solvers import b200_solve_lsa
from ..core.utils import Permutation, PermutationSpec, StateDict, apply_perm, make_identity_perm
from .activation_matching import cross_features_inner_product


# warm-started assignment solves between sweeps (PLB_WM_WARM=0: every solve cold, i.e. SciPy's tie-breaking too)
WARM_START = __import__("os").environ.get("PLB_WM_WARM", "1") == "1"


class _GroupPlan:
    """Packed operands and GEMM plan of one permutation group."""

    def __init__(self, pg, key, state_as, state_bs, skip_suffixes, skip_missing, device):
        self.n = pg.size
        self.operands = []  # (pair index, tensor name, axis, kb offset)
        kb_total = 0
        for ax in sorted(pg.state, key=lambda a: (a.key, a.axis)):
            if ax.key.endswith(tuple(skip_suffixes)):
                continue
            for si, (sa, sb) in enumerate(zip(state_as, state_bs)):
                if skip_missing and not (ax.key in sa and ax.key in sb):
                    continue
                outer, rows, inner = ops.as_rows_view(sa[ax.key], ax.axis)
                assert rows == self.n
                self.operands.append((si, ax.key, ax.axis, kb_total))
                kb_total += (outer * inner + 15) // 16
        assert kb_total > 0, f"group {key} has no weights to match"
        self.kb = kb_total
        self.pa = ops.Planes(self.n, kb_total, device)
        self.pb = ops.Planes(self.n, kb_total, device)
        for si, name, axis, off in self.operands:
            ops.pack_split(state_as[si][name], axis, self.pa, kb_offset=off)
        self.plan = ops.GemmPlan(self.pa, self.pb, self.n, self.n, kb_total)
        self.cost = torch.empty(self.n, self.n, dtype=torch.float32, device=device)

    def build_cost(self, state_bs):
        for si, name, axis, off in self.operands:
            ops.pack_split(state_bs[si][name], axis, self.pb, kb_offset=off)
        self.plan.run()
        self.plan.finalize(self.cost, ops.MODE_INNER, accumulate=False)
        return self.cost


def _to_device_state(state, keys, device):
    """Replaces the group tensors of ``state`` by contiguous fp32 CUDA copies; returns
    ``{key: (device, dtype)}`` of the entries it replaced so an in-place caller can be handed its
    own device / dtype back."""
    moved = {}
    for k in keys:
        if k in state:
            t = state[k]
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                if t.device != device or t.dtype != torch.float32:
                    moved[k] = (t.device, t.dtype)
                state[k] = t.detach().to(device=device, dtype=torch.float32).contiguous()
    return moved


def weight_matching(
    spec: PermutationSpec,
    state_as: Union[StateDict, Sequence[StateDict]],
    state_bs: Union[StateDict, Sequence[StateDict]],
    max_iter=100,
    init_perm=None,
    inplace=False,
    skip_suffixes=("running_mean", "running_var"),
    skip_missing=True,
    lsa_solver=b200_solve_lsa,
    cross_weights=cross_features_inner_product,
    verbose=True,
    seed=0,
    return_costs=False,
) -> Permutation:
    if isinstance(state_as, dict):
        state_as = [state_as]
    if isinstance(state_bs, dict):
        state_bs = [state_bs]
    assert len(state_as) == len(state_bs)
    if not inplace:
        state_bs = [copy(sb) for sb in state_bs]
    first = next(iter(state_as[0].values()))
    device = first.device if first.is_cuda else torch.device("cuda", torch.cuda.current_device())

    group_keys = {ax.key for pg in spec.values() for ax in pg.state}
    state_as = [copy(sa) for sa in state_as]
    moved_bs = []
    for sa, sb in zip(state_as, state_bs):
        _to_device_state(sa, group_keys, device)
        moved_bs.append(_to_device_state(sb, group_keys, device))

    perm = make_identity_perm(spec) if init_perm is None else deepcopy(init_perm)
    if init_perm is not None:
        for sb in state_bs:
            apply_perm(init_perm, spec, sb, inplace=True)
    perm = {k: p.to(device=device, dtype=torch.int64).contiguous() for k, p in perm.items()}
    perm_names = list(perm.keys())
    all_costs = {}
    rng = torch.Generator()
    rng.manual_seed(seed)
    fused = lsa_solver is b200_solve_lsa and cross_weights is cross_features_inner_product
    plans = {}
    duals_prev = {}
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    gains = torch.zeros(max(len(perm_names), 1), dtype=torch.float64, device=device)  # one slot per visit of a sweep
    visited = []

    # Two visits of a sweep commute unless their groups touch a common tensor (the cost of one then
    # depends on the permutation the other applies).  Scheduling a sweep by levels of that conflict
    # DAG — a visit runs once every EARLIER conflicting visit has been applied — gives exactly the
    # sequential result, and the visits of a level share one batched LAP launch.
    touched = {p: {ax.key for ax in spec[p].state} for p in perm_names}

    def levels(order):
        lvl, out = [], []
        for j, p in enumerate(order):
            l = 0
            for i in range(j):
                if touched[order[i]] & touched[p]:
                    l = max(l, lvl[i] + 1)
            lvl.append(l)
            while len(out) <= l:
                out.append([])
            out[l].append(p)
        return out

    def finish_visit(iteration, p, A, newP):
        ops.wm_progress(A, newP, flag, gains[len(visited):len(visited) + 1])
        visited.append(p)
        perm[p] = ops.compose_perm(perm[p], newP)
        all_costs[p] = A
        for sb in state_bs:
            apply_perm({p: newP}, spec, sb, inplace=True)

    with torch.no_grad():
        for iteration in range(max_iter):
            statuses = []
            order = [perm_names[i] for i in torch.randperm(len(perm_names), generator=rng)]
            if fused:
                for wave in levels(order):
                    mats = []
                    for p in wave:
                        if p not in plans:
                            plans[p] = _GroupPlan(spec[p], p, state_as, state_bs, skip_suffixes, skip_missing, device)
                        mats.append(plans[p].build_cost(state_bs))
                    # every group is re-solved once per sweep on costs that change little: the solver starts from
                    # the column duals of the group's previous visit (same optimum, far fewer augmenting steps)
                    newPs, _, status, duals = ops.lap_solve_batched(
                        mats, True, v_init=[duals_prev.get(p) for p in wave] if WARM_START else None,
                        return_duals=True)
                    statuses.append(status)
                    for p, A, newP, v in zip(wave, mats, newPs, duals):
                        duals_prev[p] = v[newP]  # B's units are about to be permuted by newP: so are its columns
                        finish_visit(iteration, p, A, newP)
            else:
                for p in order:
                    pg = spec[p]
                    n = pg.size
                    A = torch.zeros(n, n, device=device)
                    for ax in pg.state:
                        if ax.key.endswith(tuple(skip_suffixes)):
                            continue
                        for sa, sb in zip(state_as, state_bs):
                            if skip_missing and not (ax.key in sa and ax.key in sb):
                                continue
                            A.add_(cross_weights(sa[ax.key], sb[ax.key], ax.axis))
                    assert A.norm() > 0
                    newP = lsa_solver(A).to(device=device, dtype=torch.int64)
                    finish_visit(iteration, p, A, newP)
            if statuses:
                ops.raise_on_lap_status(torch.cat(statuses))
            progress = bool(flag.item())  # the sweep's one synchronising readback
            flag.zero_()
            if verbose:  # the reference prints newL - oldL per visit (:82-83); same lines, after the sweep
                for p, g in zip(visited, gains[:len(visited)].tolist()):
                    print(f"{iteration}/{p.key}:{p.axis}: {g}")
            visited.clear()
            if not progress:
                break
        assert all(bool(c.any()) for c in all_costs.values()), "a group's weight cost matrix is all zero"
        perm = {k: v.cpu() for k, v in perm.items()}
        if inplace:  # the caller's dicts get their tensors back on the device / dtype they came with
            for sb, moved in zip(state_bs, moved_bs):
                for k, (dev0, dt0) in moved.items():
                    sb[k] = sb[k].to(device=dev0, dtype=dt0)
        if return_costs:
            return perm, all_costs
        return perm
