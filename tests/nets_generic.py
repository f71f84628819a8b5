"""Small generic networks for the lowering / converter tests (NOT reference models: the converter accepts any
nn.Module, nn2bnn.py:32-45).  `drop` builds the dropout module so that the same architecture can be instantiated
with the product's MCDropout (GPU) or with an oracle-mask-injecting stand-in (CPU reference)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import philox


class Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        idt = x if self.down is None else self.down(x)     # the shortcut is traced AFTER the main branch
        out = out + idt
        return F.relu(out)


class SmallResNet(nn.Module):
    """stem -> block(64) -> dropout -> block(128, /2) -> dropout -> global pool -> dropout -> fc; two outputs."""

    def __init__(self, drop, classes=10):
        super().__init__()
        self.stem = nn.Conv2d(3, 64, 3, 1, 1)
        self.b1 = Block(64, 64, 1)
        self.d1 = drop(0.25)
        self.b2 = Block(64, 128, 2)
        self.d2 = drop(0.5)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.d3 = drop(0.125)
        self.fc = nn.Linear(128, classes)
        self.aux_pool = nn.AvgPool2d(16)
        self.aux_fc = nn.Linear(64, classes)

    def forward(self, x):
        x = self.d1(self.b1(F.relu(self.stem(x))))
        aux = self.aux_fc(torch.flatten(self.aux_pool(x), 1))
        x = self.d2(self.b2(x))
        x = self.d3(self.pool(x).flatten(1))
        return [aux, self.fc(x)]


def plain_cnn():
    """LeNet-like stack in the order relu(maxpool(conv)) - what the converter wraps leaf by leaf."""
    return nn.Sequential(nn.Conv2d(1, 8, 5, padding=2), nn.MaxPool2d(2), nn.ReLU(), nn.Conv2d(8, 16, 3, padding=1),
                         nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(), nn.Linear(16 * 7 * 7, 32), nn.ReLU(),
                         nn.Linear(32, 10))


class InjectedDropout(nn.Module):
    """CPU stand-in: x * keep / (1 - p) with the oracle's Philox mask of (seed, stream, sample)."""

    def __init__(self, p, stream, seed, mode="element"):
        super().__init__()
        self.p, self.stream, self.seed, self.mode, self.sample = p, stream, seed, mode, 0

    def forward(self, x):
        keep = torch.from_numpy(philox.keep_mask(self.seed, self.stream, self.sample, tuple(x.shape), self.p,
                                                 mode=self.mode))
        return x * keep.to(x.dtype) * (0.0 if self.p >= 1 else float(np.float32(1.0) / (np.float32(1.0) - np.float32(self.p))))


def randomize_bn(model, seed=3):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.normal_(0, 0.1, generator=g)
    return model


def reference_mean(model, x, sites, S, sample0=0):
    """float64 CPU mean over S passes of `model` whose stochastic modules are InjectedDropout (`sites`)."""
    model = model.double().eval()
    acc = None
    with torch.no_grad():
        for s in range(S):
            for m in sites:
                m.sample = sample0 + s
            out = model(x.double())
            out = out if isinstance(out, (list, tuple)) else [out]
            acc = [o.clone() for o in out] if acc is None else [a + o for a, o in zip(acc, out)]
    return [a / S for a in acc]


class MCDropout(nn.Dropout):
    """Same shape as the REFERENCE's own class (resnet18.py:207-210): an nn.Dropout subclass, not this package's."""

    def forward(self, x):
        return F.dropout(x, self.p, True, self.inplace)


class RefStyleBlock(nn.Module):
    """BasicBlock written like the reference's (resnet18.py:32-48): `out += residual`, module ReLU, shortcut last."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out += residual
        return self.relu(out)


class RefStyleNet(nn.Module):
    """Two-exit network in the style of ResNet18MCEarlyExit.forward (resnet18.py:302-346): stem without ReLU, stage =
    Sequential(blocks, dropout), exit branch on F.relu of an already non-negative tensor, F.avg_pool2d + view heads,
    a side-effect attribute, list output."""

    def __init__(self, drop, classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = nn.Sequential(nn.Sequential(RefStyleBlock(64, 64, 1)), drop(0.5))
        self.layer2 = nn.Sequential(RefStyleBlock(64, 128, 2), RefStyleBlock(128, 128, 1))
        self.ex1conv1 = nn.Conv2d(64, 128, 3, 2, 1, bias=False)
        self.ex1bn1 = nn.BatchNorm2d(128)
        self.exit1_dropout = drop(0.25)
        self.ex1linear = nn.Linear(128, classes)
        self.exit_dropout = drop(0.25)
        self.linear = nn.Linear(128, classes)

    def forward(self, x):
        out = self.bn1(self.conv1(x))
        out = self.layer1(out)
        out1 = self.ex1bn1(self.ex1conv1(F.relu(out)))
        out1 = F.avg_pool2d(F.relu(out1), 8)
        middle1_fea = out1
        out1 = out1.view(out1.size(0), -1)
        out1 = self.ex1linear(self.exit1_dropout(out1))
        out = self.layer2(out)
        out = F.avg_pool2d(F.relu(out), 8)
        out = out.view(out.size(0), -1)
        out = self.linear(self.exit_dropout(out))
        self.intermediary_output_list = (out, [out1], None, [middle1_fea])
        return [out1, out]
