"""GPU parity of the on-device evaluation metrics (csrc/eval.cu) against the reference's own
globaltest / classtest outputs (tests/golden/eval.npz) and the oracle on synthetic test sets."""
import types

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import fedmlp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
KEYS = ("mAP", "BACC", "R", "F1", "auc", "P", "hamming_loss")


@pytest.fixture(scope="module")
def E(lib):
    from fedmlp_b200 import evaluation
    return evaluation


def test_globaltest_golden(E):
    z = gu.load("eval.npz")
    probs, labels = torch.from_numpy(z["probs"]).to(DEV), torch.from_numpy(z["labels"]).to(DEV)
    r = E.multilabel_metrics(probs, labels, scores_are_probs=True)
    for k in KEYS:
        assert abs(float(r[k]) - float(z[f"globaltest/{k}"])) <= 1e-12, k
    assert isinstance(r["mAP"], torch.Tensor) and r["mAP"].dtype == torch.float32      # torch.tensor(APs).mean()
    for i in range(probs.shape[1]):
        d = E.multilabel_metrics(probs, labels, scores_are_probs=True, classid=i)
        for k in ("BACC", "R", "F1", "P"):
            assert abs(float(d[k]) - float(z[f"classtest/{i}/{k}"])) <= 1e-12, (i, k)
    # from logits: the GPU sigmoid may differ from the CPU one in the last ulp, which can move a sample across a tie
    r2 = E.multilabel_metrics(torch.from_numpy(z["logits"]).to(DEV), labels)
    for k in KEYS:
        assert abs(float(r2[k]) - float(z[f"globaltest/{k}"])) <= 1e-4, k


def test_ties_golden(E):
    z = gu.load("eval.npz")
    counts, aa = E.class_statistics(torch.from_numpy(z["ties/p"]).to(DEV)[:, None].contiguous(),
                                    torch.from_numpy(z["ties/y"]).to(DEV)[:, None].contiguous(), scores_are_probs=True)
    assert abs(float(aa[0, 0]) - float(z["ties/ap"])) <= 1e-15
    assert abs(float(aa[0, 1]) - float(z["ties/auc"])) <= 1e-15
    assert counts[0, :2].tolist() == [6, 6]


@pytest.mark.parametrize("N,C,levels", [(1, 1, 0), (300, 5, 0), (4097, 14, 0), (25596, 14, 0), (5000, 3, 7), (2500, 32, 50)])
def test_metrics_vs_oracle(E, N, C, levels):
    g = torch.Generator().manual_seed(N + C)
    labels = (torch.rand(N, C, generator=g) < torch.linspace(0.03, 0.5, C)).float()
    labels[0] = 1.0
    if N > 1:
        labels[1] = 0.0
    logits = 1.5 * torch.randn(N, C, generator=g) + 1.2 * (labels - 0.5)
    probs = torch.sigmoid(logits)
    if levels:                                   # quantised scores: massive ties
        probs = torch.round(probs * levels) / levels
    ref = O.eval_metrics(probs.numpy(), labels.numpy())
    got = E.multilabel_metrics(probs.to(DEV), labels.to(DEV), scores_are_probs=True)
    for k in KEYS:
        a, b = float(got[k]), float(ref[k])
        assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-11 * max(1.0, abs(b)), (k, a, b)
    _, aa = E.class_statistics(probs.to(DEV), labels.to(DEV), scores_are_probs=True)
    if N > 1:
        np.testing.assert_allclose(aa[:, 0].cpu().numpy(), np.array(ref["APs"]), rtol=1e-12, atol=0)
    # deterministic
    _, aa2 = E.class_statistics(probs.to(DEV), labels.to(DEV), scores_are_probs=True)
    assert torch.equal(aa, aa2) or (torch.isnan(aa) == torch.isnan(aa2)).all()


def test_globaltest_drop_in(E):
    """Same signature as the reference's globaltest(net, test_dataset, args): dict samples, .targets, net -> (feature, logits)."""
    class DS(torch.utils.data.Dataset):
        def __init__(self):
            g = torch.Generator().manual_seed(1)
            self.targets = (torch.rand(200, 4, generator=g) < 0.3).float().numpy()
            self.targets[0] = 1; self.targets[1] = 0
            self.x = torch.randn(200, 6, generator=g) + torch.from_numpy(self.targets) @ torch.randn(4, 6, generator=g)
        def __getitem__(self, i):
            return {"image": self.x[i], "target": self.targets[i]}
        def __len__(self):
            return 200

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(6, 4)
        def forward(self, x):
            return x, self.fc(x)

    torch.manual_seed(2)
    ds, net = DS(), Net().to(DEV)
    args = types.SimpleNamespace(batch_size=16, device=DEV, n_classes=4, num_workers=0)
    got = E.globaltest(net, ds, args)
    with torch.no_grad():
        probs = torch.sigmoid(net(ds.x.to(DEV))[1]).cpu().numpy()
    ref = O.eval_metrics(probs, ds.targets)
    for k in KEYS:
        assert abs(float(got[k]) - float(ref[k])) <= 1e-9, k
    one = E.classtest(net, ds, args, 2)
    assert set(one) == {"BACC", "R", "F1", "P"}


def test_no_cpu_fallback(E):
    with pytest.raises(Exception):
        E.multilabel_metrics(torch.rand(10, 3), torch.zeros(10, 3))
