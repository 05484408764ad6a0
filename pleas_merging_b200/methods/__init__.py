"""Merge methods — same public names as pleas/methods/__init__.py:12-35 (hot-path subset)."""
from .activation_matching import (activation_matching, build_cross_module, clear_caches, compute_matching_costs,
                                  cross_features_cdist, cross_features_correlation,
                                  cross_features_inner_product)
from .partial_matching import build_partial_merge_model, expand_ratios, get_blocks, partial_merge
from .bn_stats import reset_bn_stats
from .evaluation import (eval_perm_model, eval_whole_model, get_fc_perm, permute_final_features,
                         train_eval_linear_probe)
from .budget import count_linear_flops, get_zip_ratios, partial_merge_flops, qp_ratios
from .pleas_merging import train
from .weight_matching import weight_matching
from .weight_matching_partial import apply_perm_with_padding, remove_zero_block, weight_matching_partial

__all__ = ["activation_matching", "build_cross_module", "compute_matching_costs", "cross_features_cdist",
           "cross_features_inner_product", "cross_features_correlation", "clear_caches", "weight_matching", "partial_merge", "get_blocks", "expand_ratios",
           "build_partial_merge_model", "train", "reset_bn_stats", "count_linear_flops", "partial_merge_flops", "get_zip_ratios", "qp_ratios",
           "weight_matching_partial", "apply_perm_with_padding", "remove_zero_block", "get_fc_perm", "permute_final_features", "eval_perm_model", "eval_whole_model", "train_eval_linear_probe"]
