"""Small fx-traceable residual CNNs used by the golden-vector generator and the tests.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The nets exercise every op
class the ResNet hot path has (conv3x3/1x1, BN, in-place ReLU shared between call
sites, residual add, strided downsample branch, max-pool, global average pool,
flatten, linear with bias) at channel counts that are NOT multiples of 8/32 so the
kernels' ragged edges are covered.
"""
import torch
from torch import nn


class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)  # one module, two call sites (like torchvision)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(
                nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout)
            )

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class TinyResNet(nn.Module):
    """conv-bn-relu-maxpool, two residual stages, avgpool, flatten, fc."""

    def __init__(self, width=12, num_classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(2)
        self.layer1 = nn.Sequential(_Block(width, width, 1))
        self.layer2 = nn.Sequential(_Block(width, 2 * width, 2))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(2 * width, num_classes)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer2(self.layer1(x))
        x = torch.flatten(self.avgpool(x), 1)
        return self.fc(x)


def make_pair(width=12, num_classes=10, seeds=(0, 1)):
    """Two random-init nets with perturbed BN statistics so BN is not a no-op in eval."""
    nets = []
    for s in seeds:
        torch.manual_seed(s)
        m = TinyResNet(width, num_classes)
        g = torch.Generator().manual_seed(100 + s)
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=g)
                mod.bias.data = 0.2 * torch.randn(mod.bias.shape, generator=g)
                mod.running_mean.data = 0.2 * torch.randn(mod.running_mean.shape, generator=g)
                mod.running_var.data = 0.5 + torch.rand(mod.running_var.shape, generator=g)
        nets.append(m.eval())
    return nets


def make_loader(num_batches, batch, hw=16, seed=123):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 3, hw, hw, generator=g), 0) for _ in range(num_batches)]
