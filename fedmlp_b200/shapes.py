"""state_dict layouts of the backbones named by BASELINE.json, without importing torchvision.

Only the *shapes* matter to the hot path (FedAvg streams the parameters; the CNN itself stays on
cuDNN and is out of scope).  DenseNet121 = torchvision.models.densenet121 with the classifier
swapped to n_classes (reference model/all_models.py:29-130): 727 entries, 606 of them <= 1024
elements, 121 int64 `num_batches_tracked` scalars.
"""
from __future__ import annotations

from collections import OrderedDict

import torch


def _bn(prefix, n, out):
    out[prefix + ".weight"] = ((n,), torch.float32)
    out[prefix + ".bias"] = ((n,), torch.float32)
    out[prefix + ".running_mean"] = ((n,), torch.float32)
    out[prefix + ".running_var"] = ((n,), torch.float32)
    out[prefix + ".num_batches_tracked"] = ((), torch.int64)


def densenet121_state_shapes(n_classes: int, growth=32, blocks=(6, 12, 24, 16), init=64, bn_size=4):
    """OrderedDict key -> (shape, dtype) in torchvision's state_dict order."""
    out = OrderedDict()
    out["features.conv0.weight"] = ((init, 3, 7, 7), torch.float32)
    _bn("features.norm0", init, out)
    nf = init
    for b, layers in enumerate(blocks, start=1):
        for i in range(layers):
            cin = nf + i * growth
            p = f"features.denseblock{b}.denselayer{i + 1}"
            _bn(p + ".norm1", cin, out)
            out[p + ".conv1.weight"] = ((bn_size * growth, cin, 1, 1), torch.float32)
            _bn(p + ".norm2", bn_size * growth, out)
            out[p + ".conv2.weight"] = ((growth, bn_size * growth, 3, 3), torch.float32)
        nf += layers * growth
        if b != len(blocks):
            _bn(f"features.transition{b}.norm", nf, out)
            out[f"features.transition{b}.conv.weight"] = ((nf // 2, nf, 1, 1), torch.float32)
            nf //= 2
    _bn("features.norm5", nf, out)
    out["classifier.weight"] = ((n_classes, nf), torch.float32)
    out["classifier.bias"] = ((n_classes,), torch.float32)
    return out


def efficientnet_b0_state_shapes(n_classes: int):
    """torchvision.models.efficientnet_b0 with `n_classes` outputs: 360 entries, 4,067,498 fp32 values and 49
    int64 counters at 14 classes (the reference's EfficientNet comes from efficientnet_pytorch, which differs
    slightly; only the sizes matter to FedAvg).  Shapes are read from torchvision's meta-device model."""
    import torchvision

    with torch.device("meta"):
        model = torchvision.models.efficientnet_b0(num_classes=n_classes)
    return OrderedDict((k, (tuple(v.shape), v.dtype)) for k, v in model.state_dict().items())


def synth_state_dict(shapes, seed, device="cpu", base=None, counter=100):
    """Seeded synthetic state_dict of the given layout: N(0, 0.02) around `base`."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, (shp, dt) in shapes.items():
        if dt == torch.int64:
            sd[k] = torch.tensor(counter, dtype=torch.int64).reshape(shp).to(device)
        else:
            v = torch.randn(shp, generator=g) * 0.02
            if base is not None:
                v = v + base[k].cpu()
            sd[k] = v.to(device)
    return sd


def count_params(shapes):
    f = sum(int(torch.Size(s).numel()) for s, d in shapes.values() if d == torch.float32)
    i = sum(int(torch.Size(s).numel()) for s, d in shapes.values() if d == torch.int64)
    return f, i
