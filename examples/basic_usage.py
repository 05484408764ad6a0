"""The reference README's "Basic Usage" (README.md:50-106) on the B200 library, with synthetic
data in place of a dataset: spec -> activation matching -> partial merge -> PLeaS -> BN reset.

    python examples/basic_usage.py [resnet18|resnet50] [num_batches]
"""
import os
import sys
import time

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pleas_merging_b200.core.compiler import get_permutation_spec  # noqa: E402
from pleas_merging_b200.methods.activation_matching import activation_matching  # noqa: E402
from pleas_merging_b200.methods.bn_stats import reset_bn_stats  # noqa: E402
from pleas_merging_b200.methods.partial_matching import partial_merge  # noqa: E402
from pleas_merging_b200.methods.pleas_merging import train as pleas_train  # noqa: E402


def main(arch="resnet18", num_batches=8, batch=16, hw=224, max_steps=8, verbose=True):
    torch.manual_seed(0)
    model1 = getattr(torchvision.models, arch)().cuda().eval()
    torch.manual_seed(1)
    model2 = getattr(torchvision.models, arch)().cuda().eval()
    g = torch.Generator().manual_seed(123)
    dataloader = [(torch.randn(batch, 3, hw, hw, generator=g).pin_memory(), 0) for _ in range(max(num_batches, max_steps + 1))]

    t0 = time.perf_counter()
    spec = get_permutation_spec(model1, ((1, 3, hw, hw),))
    perm, costs = activation_matching(spec, model1, model2, dataloader, num_batches=num_batches, output_costs=True)
    # a ratio of 0.0 merges a group completely, 1.0 keeps both models' units
    budget_ratios = {key: 0.0 for key in spec.keys()}
    merged_model = partial_merge(spec, model1, model2, perm, costs, budget_ratios)
    stats = {}
    optimized_model = pleas_train(dataloader, model1, model2, merged_model, spec, perm, costs, budget_ratios,
                                  WANDB=False, MAX_STEPS=max_steps, wandb_run=None, stats=stats)
    optimized_model = reset_bn_stats(optimized_model, dataloader, num_batches=num_batches)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        y = optimized_model(dataloader[0][0].cuda())
    layers = [k for k in stats if not k.startswith("_")]
    gain = sum(stats[k]["objective_init"] - stats[k]["objective_fit"] for k in layers)
    if verbose:
        print(f"{arch}: {len(spec)} permutation groups, {len(layers)} layers fitted, merge pipeline {dt:.2f} s, "
              f"least-squares objective improved by {gain:.4g}, merged logits finite: {bool(torch.isfinite(y).all())}")
    return optimized_model, stats


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["resnet18"]), num_batches=int(sys.argv[2]) if len(sys.argv) > 2 else 8)
