"""Drop-in LocalUpdate.train_FedMLP / round loop on the GPU: the orchestration is
checked against the oracle on the tensors each call actually used (per-step losses, selections,
masks, prototypes, t, aggregated weights)."""
import types
from collections import OrderedDict
from copy import deepcopy

import numpy as np
import pytest
import torch
from torch import nn

from oracle import fedmlp_oracle as O

pytestmark = pytest.mark.gpu


class SynthDataset(torch.utils.data.Dataset):
    """Sample contract of dataset/all_dataset.py:23-41 (two-view dict, numpy target row)."""

    def __init__(self, n, c, dim, seed):
        g = torch.Generator().manual_seed(seed)
        self.targets = (torch.rand(n, c, generator=g) < 0.3).float().numpy().astype(np.float32)
        for col in range(c):
            self.targets[col::c, col][:3] = 1.0
        shift = torch.randn(c, dim, generator=g)
        base = torch.randn(n, dim, generator=g)
        self.view1 = (base + torch.from_numpy(self.targets) @ shift).float()
        self.view2 = (self.view1 + 0.05 * torch.randn(n, dim, generator=g)).float()

    def __getitem__(self, index):
        return {"image_aug_1": self.view1[index], "image_aug_2": self.view2[index],
                "target": self.targets[index], "index": index, "image_id": str(index)}

    def __len__(self):
        return len(self.targets)


class TinyNet(nn.Module):
    def __init__(self, dim, feat_dim, c):
        super().__init__()
        self.fc1, self.bn, self.fc2 = nn.Linear(dim, feat_dim), nn.BatchNorm1d(feat_dim), nn.Linear(feat_dim, c)

    def forward(self, x):
        f = torch.relu(self.bn(self.fc1(x)))
        return f, self.fc2(f)


def _args(**over):
    a = types.SimpleNamespace(batch_size=16, annotation_num=1, n_classes=5, n_clients=3, local_ep=1, base_lr=1e-3,
                              device="cuda", U=0.7, L=0.3, rounds_FedMLP_stage1=2, clean_threshold=0.1,
                              noise_threshold=0.2)
    a.__dict__.update(over)
    return a


def _setup(n_clients=3, per=104, C=5, dim=12, feat=64):
    ds = SynthDataset(n_clients * per, C, dim, seed=5)
    rows, cols = np.where(ds.targets == 1)
    class_neg_idx = [rows[np.where(cols == i)[0]] for i in range(C)]
    dict_users = {k: list(range(k * per, (k + 1) * per)) for k in range(n_clients)}
    torch.manual_seed(3)
    net = TinyNet(dim, feat, C).cuda()
    return ds, class_neg_idx, dict_users, net


def test_local_update_calls_match_oracle(lib):
    from fedmlp_b200.local_training import LocalUpdate
    args = _args()
    ds, class_neg_idx, dict_users, netglob = _setup()
    k = 1
    local = LocalUpdate(args, k, deepcopy(ds), dict_users[k], class_neg_idx, class_neg_idx, active_class_list=[k])
    act, neg = [k], [c for c in range(5) if c != k]
    # hidden positives: labels of the non-annotated classes are all zero (p_pos_1 = 0, main.py:62-66)
    assert float(local.labels[:, neg].sum()) == 0 and torch.equal(local.labels[:, k], local.targets_true[:, k])
    # ---- stage 1, last round: losses + first prototypes / t
    net = deepcopy(netglob)
    ret = local.train_FedMLP(1, [0] * 5, [], None, neg, act, net)
    assert len(ret) == 8 and ret[4] == neg and ret[5] == act
    steps = local.last["steps"]
    assert len(steps) == 7 and steps[-1]["z1"].shape[0] == 8                     # 104 = 6*16 + 8
    for st in steps:
        ref = O.stage1_loss(st["z1"].cpu(), st["z2"].cpu(), st["z3"].cpu(), st["z4"].cpu(), st["y"].cpu(), act, neg, 16)
        assert abs(float(st["loss"]) - float(ref)) <= 1e-5 * abs(float(ref))
    assert abs(ret[1] - np.mean([float(s["loss"]) for s in steps])) < 1e-6
    ref_p, ref_n, ref_t = O.prototype_build(local.last["proto_feat"].cpu(), local.labels.cpu(), local.last["proto_logits"].cpu(),
                                            act, neg, 0.3, 0.7, guard_empty=False)
    np.testing.assert_allclose(ret[7].numpy(), ref_p.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(ret[6], ref_t)
    assert not ret[7].is_cuda and ret[7].shape == (10, 64)
    # ---- two stage-2 rounds against an oracle TaggingState fed with the same features
    proto_glob = torch.relu(torch.randn(10, 64, generator=torch.Generator().manual_seed(1))) + 0.05
    proto_glob[2 * k:2 * k + 2] = ret[7][2 * k:2 * k + 2]
    state = O.TaggingState(dict_users[k], neg)
    for rnd in (2, 3):
        ret = local.train_FedMLP(rnd, [0] * 5, proto_glob, None, neg, act, net)
        state.step(local.last["tag_feat"].cpu(), proto_glob, args.clean_threshold, args.noise_threshold, python_sort=True)
        assert local.traindata_idx == [[float(v) for v in l] for l in state.traindata_idx]
        assert local.idxss == O.remaining_indices(dict_users[k], state.traindata_idx)
        tgt, dis = O.mask_fill(local.targets_true.cpu().numpy(), dict_users[k], act, neg, state.traindata_idx)
        pos = {d: p for p, d in enumerate(dict_users[k])}
        for st in local.last["steps"]:
            ref = O.stage2_loss(st["z"].cpu(), st["zg"].cpu(), st["y"].cpu(), st["distill"].cpu())
            assert abs(float(st["loss"]) - float(ref)) <= 1e-5 * abs(float(ref))
        y_all, distill_all, _ = local.tagger.fill(local.targets_true)
        np.testing.assert_array_equal(y_all.cpu().numpy(), tgt)
        np.testing.assert_array_equal(distill_all.cpu().numpy(), dis)
        ref_p, ref_n, ref_t = O.prototype_build(local.last["proto_feat"].cpu(), local.labels.cpu(), local.last["proto_logits"].cpu(),
                                                act, neg, 0.3, 0.7, guard_empty=True)
        np.testing.assert_allclose(ret[7].numpy(), ref_p.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(ret[6], ref_t)
        assert local.class_num_list[neg[0]] == len(state.traindata_idx[1])


def test_round_loop_runs_and_aggregates_like_the_reference(lib):
    import fedmlp_b200 as F
    from fedmlp_b200.local_training import LocalUpdate, run_fedmlp_rounds
    args = _args()
    ds, class_neg_idx, dict_users, netglob = _setup()
    trainers = [LocalUpdate(args, i, deepcopy(ds), dict_users[i], class_neg_idx, class_neg_idx, active_class_list=[i])
                for i in range(3)]
    dict_len = [len(dict_users[i]) for i in range(3)]
    before = deepcopy(netglob.state_dict())
    seen = []
    tao, Prototype, hist = run_fedmlp_rounds(args, netglob, trainers, dict_len, rounds=4,
                                             on_round_end=lambda r, n, t, p: seen.append(r))
    assert seen == [0, 1, 2, 3] and len(hist) == 4 and all(np.isfinite(hist))
    assert not torch.equal(before["fc1.weight"], netglob.state_dict()["fc1.weight"])
    assert netglob.state_dict()["bn.num_batches_tracked"].dtype == torch.int64      # float32 average cast back on load
    P = Prototype.cpu()
    assert P.shape == (10, 64)
    assert torch.isfinite(P[:6]).all()            # classes 0..2 are annotated by clients 0..2
    assert torch.isnan(P[6:]).all()               # classes 3, 4: nobody annotates -> 0/0 like FedAvg.py:85-86
    assert tao.shape == (5,) and np.all((tao >= 0) & (tao <= 1))
    assert all(len(tr.traindata_idx) == 8 for tr in trainers)
    # aggregation of the final round == reference arithmetic on the clients' weights
    w_locals = [deepcopy(netglob.state_dict()) for _ in range(3)]
    for i, w in enumerate(w_locals):
        for kname in w:
            if w[kname].is_floating_point():
                w[kname] += 0.01 * (i + 1)
    ref = O.fedavg([OrderedDict((k, v.cpu()) for k, v in w.items()) for w in w_locals], dict_len)
    out = F.FedAvg(w_locals, dict_len)
    for kname in ref:
        assert torch.equal(out[kname].cpu(), ref[kname])


def test_flat_adam_matches_torch_adam(lib):
    """Fused flat Adam vs torch.optim.Adam with the reference's hyper-parameters
    (lr, betas=(0.9, 0.999), weight_decay=5e-4): identical gradients are fed to both for several
    steps (so the comparison isolates the optimizer arithmetic); BatchNorm buffers untouched."""
    from fedmlp_b200.optim import FlatAdam
    torch.manual_seed(0)
    ref_net = TinyNet(12, 64, 5)
    net = deepcopy(ref_net).cuda()
    opt_ref = torch.optim.Adam(ref_net.parameters(), lr=3e-3, betas=(0.9, 0.999), weight_decay=5e-4)
    opt = FlatAdam(net, lr=3e-3, betas=(0.9, 0.999), weight_decay=5e-4)
    g = torch.Generator().manual_seed(1)
    buffers_before = {k: v.clone() for k, v in net.named_buffers()}
    for it in range(8):
        scale = 10.0 ** (it % 4 - 2)                       # gradients from 1e-2 to 1e+1
        for (name, p_ref), (_, p) in zip(ref_net.named_parameters(), net.named_parameters()):
            grad = torch.randn(p_ref.shape, generator=g) * scale
            p_ref.grad = grad.clone()
            p.grad.copy_(grad)                             # views of the flat gradient buffer
        opt_ref.step()
        opt.step(zero_grad=True)
        assert float(opt.grad.abs().max()) == 0.0
    for (name, p_ref), (_, p) in zip(ref_net.named_parameters(), net.named_parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), p_ref.detach().numpy(), rtol=1e-5, atol=1e-7)
    for k, v in net.named_buffers():
        assert torch.equal(v, buffers_before[k])           # running statistics are not parameters
    # a real training step works through the flat gradient views, and the flat state_dict feeds
    # FedAvg's one-launch path directly
    x = torch.randn(16, 12, generator=g).cuda()
    y = (torch.rand(16, 5, generator=g) < 0.3).float().cuda()
    before = net.fc2.weight.detach().clone()
    _, z = net(x)
    torch.nn.functional.binary_cross_entropy_with_logits(z, y).backward()
    assert float(opt.grad.abs().max()) > 0.0
    opt.step(zero_grad=True)
    assert not torch.equal(before, net.fc2.weight.detach())
    import fedmlp_b200 as F
    out = F.FedAvg([opt.flat, opt.flat.clone()], [3, 5])
    np.testing.assert_allclose(out["fc1.weight"].cpu().numpy(), net.state_dict()["fc1.weight"].cpu().numpy(), rtol=1e-6, atol=1e-7)
