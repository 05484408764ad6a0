"""pleas_merging_b200 — B200-native implementation of the PLeaS-Merging merge hot path.

Drop-in for the reference's Python API (``get_permutation_spec``, ``activation_matching``,
``weight_matching``, ``partial_merge``, ``pleas_merging.train``) on hand-written sm_100a CUDA
kernels behind a C ABI (include/pleas_b200.h).  There is no CPU fallback: the kernels' shared
library must be built (``python -m pleas_merging_b200.build``) and a CUDA device present.
"""
from .core.compiler import check_permutation_spec, get_permutation_spec
from .core.solvers import b200_solve_lsa
from .core.utils import (Axis, Permutation, PermutationGroup, PermutationSpec, apply_perm, invert_perm,
                         make_identity_perm, make_random_perm, perm_eq)
from .methods import (activation_matching, clear_caches, cross_features_cdist, cross_features_correlation,
                      cross_features_inner_product, get_blocks,
                      partial_merge, reset_bn_stats, train, weight_matching)

__version__ = "0.1.0"
