"""Data model of the merge path: permutable axes, permutation groups, permutations.

Mirrors the reference's public types (pleas/core/utils.py:16-56) and permutation helpers
(:143-246) so specs and permutations are interchangeable between the two packages:
``Axis`` hashes like the reference's frozen dataclass (``hash((key, axis))``) and compares
by value with any object exposing ``key``/``axis``, so a dict keyed by one package's Axis can
be indexed with the other's.
"""
from copy import copy
from dataclasses import dataclass
from typing import Dict, Sequence, Set, Tuple, Union

import torch
from torch import nn


@dataclass(frozen=True, eq=False)
class Axis:
    """Axis ``axis`` of the state-dict tensor / fx node named ``key``."""
    key: str
    axis: int

    def __eq__(self, other):
        try:
            return self.key == other.key and self.axis == other.axis
        except AttributeError:
            return NotImplemented

    def __hash__(self):
        return hash((self.key, self.axis))

    def __iter__(self):  # allows Axis(*ax) / tuple(ax)
        yield self.key
        yield self.axis

    def __str__(self):
        return f"{self.key}:{self.axis}"

    __repr__ = __str__


@dataclass
class PermutationGroup:
    """Axes that must be permuted together: ``state`` = state-dict tensor axes, ``node`` =
    fx node output axes living in the same permutation space."""
    size: int
    state: Set[Axis]
    node: Set[Axis]


PermutationKey = Axis
PermutationSpec = Dict[PermutationKey, PermutationGroup]
Permutation = Dict[PermutationKey, torch.Tensor]
StateDict = Dict[str, torch.Tensor]
InputsOrShapes = Union[Tuple[tuple, ...], Tuple[torch.Tensor, ...]]


def get_attr(obj, names: Sequence[str]):
    for n in names:
        obj = getattr(obj, n)
    return obj


def set_attr(obj, names: Sequence[str], val):
    setattr(get_attr(obj, names[:-1]), names[-1], val)


def make_identity_perm(spec: PermutationSpec) -> Permutation:
    return {k: torch.arange(pg.size) for k, pg in spec.items()}


def make_random_perm(spec: PermutationSpec, generator=None) -> Permutation:
    return {k: torch.randperm(pg.size, generator=generator) for k, pg in spec.items()}


def invert_perm(perm):
    if isinstance(perm, dict):
        return {k: invert_perm(p) for k, p in perm.items()}
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(len(perm), device=perm.device, dtype=perm.dtype)
    return inv


def perm_eq(perm1: Permutation, perm2: Permutation) -> bool:
    return len(perm1) == len(perm2) and all(bool((perm2[k].cpu() == p.cpu()).all()) for k, p in perm1.items())


def index_select_axis(weight: torch.Tensor, axis: int, P: torch.Tensor) -> torch.Tensor:
    """index_select along one axis; float32 CUDA tensors go through the library's gather
    kernel, anything else (CPU state dicts, integer buffers) through torch."""
    if weight.is_cuda and weight.dtype == torch.float32:
        from .. import ops

        return ops.gather_axis(weight, axis, P)
    return torch.index_select(weight, axis, P.to(weight.device))


def apply_perm(perm: Permutation, spec: PermutationSpec, state, inplace=False, skip_missing=True):
    """Applies per-group permutations to every state axis of each group
    (reference: pleas/core/utils.py:203-246)."""
    if isinstance(state, nn.Module):
        assert inplace
        state.load_state_dict(apply_perm(perm, spec, state.state_dict(), inplace=True))
        return state
    if not inplace:
        state = copy(state)
    for key, P in perm.items():
        if P is None:
            continue
        pg = spec[key]
        assert P.shape == (pg.size,)
        for ax in pg.state:
            if skip_missing and ax.key not in state:
                continue
            state[ax.key] = index_select_axis(state[ax.key], ax.axis, P)
    return state


def reset_running_stats(net):
    """reference: pleas/core/utils.py:383-392"""
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.reset_running_stats()
