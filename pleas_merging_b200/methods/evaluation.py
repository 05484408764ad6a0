"""Evaluation of merged models with the source models' classifiers (reference
pleas/methods/pleas_merging.py:408-585).

A partially merged backbone emits features in the merged layout ``[merged | separate-1 |
separate-2]`` of the classifier's input group; ``permute_final_features`` brings them back to
the channel order source model ``idx`` was trained with, so that model's own ``fc`` can be put
on top.  These are plain inference loops over the user's modules (library forwards); accuracy
is top-1 computed on device — the reference uses ``torchmetrics.Accuracy(task="multiclass")``,
which is the same number.
"""
import torch

from ..core.utils import Axis
from .partial_matching import get_blocks


def get_fc_perm(perm, spec, costs, budget_ratios):
    """Blocks ``(b1, b2, b1c, b2c)`` of the group that feeds the classifier — the one whose state
    holds ``Axis('fc.weight', 1)`` (reference :408-433)."""
    perm_blocks = get_blocks(spec, perm, costs, budget_ratios, False)
    for k, pg in spec.items():
        if Axis("fc.weight", 1) in pg.state:
            return perm_blocks[k]
    raise ValueError("No fc perm found")


def permute_final_features(features, fc_perm, idx):
    """Features ``[N, n + 2m]`` of the merged backbone -> ``[N, n + m]`` in source model ``idx``'s
    channel order (reference :436-465): take the merged block and that model's separate block,
    then undo the block ordering with the inverse of ``cat[b, bc]``."""
    bi1, bi2, bi1c, bi2c = fc_perm
    ni, mi = len(bi1), len(bi1c)
    own = slice(ni, ni + mi) if idx == 0 else slice(ni + mi, ni + 2 * mi)
    sliced = torch.cat([features[:, :ni], features[:, own]], dim=1)
    order = torch.cat([bi1, bi1c] if idx == 0 else [bi2, bi2c], dim=0)
    return sliced[:, torch.argsort(order).to(sliced.device)]


class _Top1:
    """Running top-1 accuracy on device (one host read at the end)."""

    def __init__(self, device):
        self.hit = torch.zeros((), dtype=torch.int64, device=device)
        self.n = 0

    def update(self, logits, y):
        self.hit += (logits.argmax(dim=1) == y.to(logits.device)).sum()
        self.n += int(y.shape[0])

    def compute(self):
        return self.hit.float() / max(self.n, 1)


def eval_perm_model(model, fc, dataloader, num_classes, fc_perm, idx):
    """Top-1 accuracy of the merged ``model`` under source model ``idx``'s classifier ``fc``
    (reference :468-496).  ``num_classes`` is kept for signature compatibility."""
    model.eval()
    device = next(model.parameters()).device
    acc = _Top1(device)
    with torch.no_grad():
        for x, y in dataloader:
            feats = permute_final_features(model(x.to(device, non_blocking=True)), fc_perm, idx)
            acc.update(fc(feats), y)
    return acc.compute()


def eval_whole_model(model, dataloader, num_classes):
    """Top-1 accuracy of ``model`` as it is (reference :574-585)."""
    model.eval()
    device = next(model.parameters()).device
    acc = _Top1(device)
    with torch.no_grad():
        for x, y in dataloader:
            acc.update(model(x.to(device, non_blocking=True)), y)
    out = acc.compute()
    print(out)
    return out


def train_eval_linear_probe(model, train_dataloader, test_dataloader, num_classes, wandb_run, dataset_name, lr=1e-3,
                            epochs=10, input_shape=(1, 3, 224, 224)):
    """Adam-trained linear classifier on the frozen merged backbone, then its test accuracy
    (reference :499-570; same optimiser, cosine schedule to lr/10 and logging keys).
    ``wandb_run`` may be None.  Returns the trained ``nn.Linear``."""
    device = next(model.parameters()).device
    model.eval()
    with torch.no_grad():
        out_feats = model(torch.randn(*input_shape, device=device)).shape[-1]
    fc = torch.nn.Linear(out_feats, int(num_classes)).to(device)
    optimizer = torch.optim.Adam(fc.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, epochs * len(train_dataloader), eta_min=lr / 10)
    loss_fn = torch.nn.CrossEntropyLoss()
    log = wandb_run.log if wandb_run is not None else (lambda metrics: None)
    for epoch in range(epochs):
        fc.train()
        acc, total = _Top1(device), torch.zeros((), device=device)
        loss = None
        for x, y in train_dataloader:
            x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
            with torch.no_grad():
                feats = model(x)
            y_hat = fc(feats)
            loss = loss_fn(y_hat, y)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            sched.step()
            acc.update(y_hat.detach(), y)
            total += loss.detach()
        log({f"{dataset_name}_linear_probe_train_acc": acc.compute().item(),
             f"{dataset_name}_linear_probe_train_loss": float(loss) if loss is not None else 0.0,
             "epoch": epoch, f"{dataset_name}_total_loss": float(total) / max(len(train_dataloader), 1)})
    fc.eval()
    acc = _Top1(device)
    with torch.no_grad():
        for x, y in test_dataloader:
            acc.update(fc(model(x.to(device, non_blocking=True))), y)
    log({f"{dataset_name}_linear_probe_acc": acc.compute().item()})
    return fc
