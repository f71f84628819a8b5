"""Plain torch networks the converter fixtures (tests/golden/make_golden_converter.py) are frozen on.
name -> (factory, input shape without batch, p, nSamples, seed)."""
import torch.nn as nn


def _mlp():
    return nn.Sequential(nn.Linear(12, 32), nn.ReLU(), nn.Linear(32, 16), nn.ReLU(), nn.Linear(16, 4))


def _cnn():
    return nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                         nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                         nn.Flatten(), nn.Linear(32 * 4 * 4, 64), nn.ReLU(), nn.Linear(64, 10))


class _Nested(nn.Module):
    """nested containers + BatchNorm (not a convertible leaf) + a strided conv"""

    def __init__(self):
        super().__init__()
        self.features = nn.Sequential(
            nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU()),
            nn.Sequential(nn.Conv2d(8, 16, 3, stride=2, padding=1), nn.BatchNorm2d(16), nn.ReLU()))
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.flat = nn.Flatten()
        self.head = nn.Linear(16, 6)

    def forward(self, x):
        return self.head(self.flat(self.pool(self.features(x))))


NETS = {
    "mlp": (_mlp, (12,), 0.5, 4, 1234),
    "cnn": (_cnn, (3, 16, 16), 0.25, 5, 99),
    "nested": (lambda: _Nested().eval(), (3, 8, 8), 0.125, 3, 7),
}
