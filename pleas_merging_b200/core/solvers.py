"""Linear-sum-assignment solver plug point.

``lsa_solver(A, maximize=True) -> LongTensor[n]`` with ``result[i]`` = column assigned to row
``i`` is the reference's solver plug-in interface (pleas/core/solvers.py:18-33, used at
pleas/methods/activation_matching.py:173 and pleas/methods/weight_matching.py:78).  Here it is
the library's GPU shortest-augmenting-path kernel (csrc/lap.cu): same algorithm, float64
arithmetic and tie-breaking as SciPy's ``linear_sum_assignment``, so the assignment is
identical — without the D2H copy of the cost matrix.  No SciPy, no CPU fallback.
"""
import torch

from .. import ops


def b200_solve_lsa(A, maximize=True):
    """Drop-in for ``scipy_solve_lsa``: float32 cost matrix (any device; moved to the current
    CUDA device if needed) -> CPU int64 tensor of column indices, like the reference returns."""
    if not torch.is_tensor(A):
        A = torch.as_tensor(A)
    if A.dim() != 2 or A.shape[0] != A.shape[1]:
        raise ValueError(f"expected a square cost matrix, got {tuple(A.shape)}")
    A = A.detach().to(device="cuda" if not A.is_cuda else A.device, dtype=torch.float32)
    if A.shape[0] == 0:
        return torch.empty(0, dtype=torch.int64)
    outs, _, status = ops.lap_solve_batched([A], maximize)
    ops.raise_on_lap_status(status)
    return outs[0].cpu()


def solve_lsa_batched(costs, maximize=True):
    """All groups of one matching call in a single launch; returns CPU int64 tensors."""
    outs, _, status = ops.lap_solve_batched(list(costs), maximize)
    ops.raise_on_lap_status(status)
    if not outs:
        return []
    flat = torch.cat(outs).cpu()  # one D2H copy for all groups
    return [t.clone() for t in flat.split([o.numel() for o in outs])]


# Name-compatible alias for code written against the reference's module.
scipy_solve_lsa = b200_solve_lsa
