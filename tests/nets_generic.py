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
