"""BatchNorm statistics re-estimation after merging.

Every reference driver follows ``train`` with the same loop
(experiments/shared_label_space/run_domainnet.py:327-341,
experiments/different_label_space/run_torchvision.py:268-274): put the merged model in train
mode, optionally ``reset_running_stats`` every BatchNorm2d, and stream ~100 batches through it
without gradients so the running statistics match the merged weights.  The arithmetic is the
user's own modules (cuDNN batch norm in training mode — the normalised activations feed the
next layer, so it cannot be replaced by a side reduction); what this helper adds is the B200
plumbing of the other loops: CUDA-graph replay of the forward and H2D prefetch of the next batch.
"""
import torch

from ..core.utils import reset_running_stats
from ..graphs import GraphedStep
from ..parallel import BatchSharder, device_prefetch


def reset_bn_stats(model, dataloader, num_batches=101, reset=True, use_cuda_graph=True):
    """Re-estimates BatchNorm running statistics of ``model`` in place and returns it in eval mode.

    ``num_batches=101`` reproduces the drivers' ``if idx > 100: break`` loops; ``reset=True`` is the
    DomainNet driver (reset, then exponential moving average with the modules' own momentum),
    ``reset=False`` the torchvision driver."""
    device = next(iter(model.parameters())).device
    if device.type != "cuda":
        raise RuntimeError("pleas_merging_b200 runs on a CUDA device: move the model to cuda first")
    model.train()
    if reset:
        reset_running_stats(model)
    step = GraphedStep(lambda x: model(x.float()), None, use_cuda_graph)
    with torch.no_grad():
        for _, x in device_prefetch(BatchSharder(((b[0], 0) for b in dataloader), num_batches, 0, 1), device):
            step(x)
    torch.cuda.synchronize(device)
    step.clear()
    return model.eval()
