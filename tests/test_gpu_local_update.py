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


# ---------------------------------------------------------------------------------------------------------
# Whole-call equivalence with the RECORDED reference run (tests/golden/flow.npz = the unmodified
# LocalUpdate.train_FedMLP driven on CPU by oracle/make_golden.py): same dataset, model, seeds, call sequence.
def test_local_update_whole_call_matches_recorded_reference_run(lib):
    import copy
    import random

    import golden_util as gu
    import oracle.make_golden as G
    from fedmlp_b200.local_training import LocalUpdate

    z = gu.load("flow.npz")
    N, C, FEAT, CLIENT = int(z["meta/N"]), int(z["meta/C"]), int(z["meta/FEAT"]), int(z["meta/client"])
    DIM = 12

    class GpuTinyNet(G.TinyNet):
        """The golden run's model on the GPU.  __deepcopy__ draws from the global RNG exactly like the
        reference run's (a fresh TinyNet is initialised, then overwritten), so the loaders shuffle alike."""

        def __deepcopy__(self, memo):
            new = GpuTinyNet(self.fc1.in_features, self.fc1.out_features, self.fc2.out_features)
            new.load_state_dict(copy.deepcopy(self.state_dict()))
            new.role = "global"
            new.train(self.training)
            return new.to(self.fc1.weight.device)

    # ---- the exact preamble of oracle/make_golden.py:make_flow
    torch.manual_seed(7); np.random.seed(7); random.seed(7)
    ds = G.SynthDataset(N, C, DIM, seed=99)
    assert np.array_equal(ds.targets, z["meta/targets_true"])
    idxs = [int(v) for v in z["meta/idxs"]]
    rows, cols = np.where(ds.targets == 1)
    class_neg_idx = [rows[np.where(cols == i)[0]] for i in range(C)]
    args = G.make_flow_args(device="cuda")
    net = GpuTinyNet(DIM, FEAT, C)                       # initialised on the CPU from the global RNG, like the golden run
    G.TRACE.clear()
    local = LocalUpdate(args, CLIENT, copy.deepcopy(ds), idxs, class_neg_idx, class_neg_idx, active_class_list=[CLIENT])
    tao = [0] * C
    neg_in = [c for c in range(C) if c != CLIENT]

    def step_losses(prefix):
        kinds = [str(k) for k in z[f"{prefix}/kinds"]]
        return [float(z[f"{prefix}/{i}/loss"]) for i, k in enumerate(kinds) if k == "loss"]

    # ---- last stage-1 round
    rnd = args.rounds_FedMLP_stage1 - 1
    work = copy.deepcopy(net).cuda()
    ret = local.train_FedMLP(rnd, tao, [], None, neg_in, [CLIENT], work)
    got = [float(s["loss"]) for s in local.last["steps"]]
    ref = step_losses("s1")
    assert len(got) == len(ref)
    np.testing.assert_allclose(got, ref, rtol=2e-4)               # same batches in the same order, CPU vs GPU fp32
    assert abs(ret[1] - float(z["s1/loss_mean"])) <= 2e-4 * abs(float(z["s1/loss_mean"]))
    assert ret[4] == [int(v) for v in z["s1/neg_list"]] and ret[5] == [int(v) for v in z["s1/act_list"]]
    np.testing.assert_allclose(ret[6], z["s1/t"], rtol=0, atol=1.5 / len(idxs))
    np.testing.assert_allclose(ret[7].numpy(), z["s1/proto"], rtol=2e-3, atol=2e-4)
    proto_glob = torch.from_numpy(z["proto_glob"].copy())
    # ---- two stage-2 rounds: tagging state, losses, prototypes, t
    for r in range(2):
        rnd = args.rounds_FedMLP_stage1 + r
        work2 = copy.deepcopy(work)
        ret = local.train_FedMLP(rnd, tao, proto_glob, None, list(neg_in), [CLIENT], work2)
        for j in range(2 * len(neg_in)):
            assert local.traindata_idx[j] == [float(v) for v in z[f"s2_{r}/traindata_idx/{j}"]], (r, j)
        for j in range(len(neg_in)):
            assert sorted(local.idxss[j]) == [int(v) for v in z[f"s2_{r}/idxss/{j}"]]
        got = [float(s["loss"]) for s in local.last["steps"]]
        np.testing.assert_allclose(got, step_losses(f"s2_{r}"), rtol=5e-4)
        assert abs(ret[1] - float(z[f"s2_{r}/loss_mean"])) <= 5e-4 * abs(float(z[f"s2_{r}/loss_mean"]))
        np.testing.assert_allclose(ret[6], z[f"s2_{r}/t"], rtol=0, atol=1.5 / len(idxs))
        np.testing.assert_allclose(ret[7].numpy(), z[f"s2_{r}/proto"], rtol=5e-3, atol=5e-4)
        work = work2
    G.TRACE.clear()


def test_round_loop_flat_parameters_match_the_reference_copies(lib):
    """run_fedmlp_rounds(flat=True) — one model per client living in a flat buffer, FlatAdam, FedAvg on the flat
    buffers, one flat copy per load — against flat=False (deepcopy / torch.optim.Adam / load_state_dict, the
    reference's plumbing) from identical seeds."""
    from fedmlp_b200.local_training import LocalUpdate, run_fedmlp_rounds
    outs = []
    for flat in (False, True):
        args = _args()
        ds, class_neg_idx, dict_users, netglob = _setup()
        torch.manual_seed(11)
        trainers = [LocalUpdate(args, i, deepcopy(ds), dict_users[i], class_neg_idx, class_neg_idx, active_class_list=[i])
                    for i in range(3)]
        dict_len = [len(dict_users[i]) for i in range(3)]
        tao, Prototype, hist = run_fedmlp_rounds(args, netglob, trainers, dict_len, rounds=4, flat=flat)
        outs.append(dict(sd={k: v.detach().cpu().clone() for k, v in netglob.state_dict().items()}, tao=tao,
                         proto=Prototype.cpu(), hist=hist, idx=[tr.traindata_idx for tr in trainers]))
    a, b = outs
    assert b["sd"]["bn.num_batches_tracked"].dtype == torch.int64
    assert torch.equal(a["sd"]["bn.num_batches_tracked"], b["sd"]["bn.num_batches_tracked"])
    for k in a["sd"]:
        if a["sd"][k].is_floating_point():
            np.testing.assert_allclose(b["sd"][k].numpy(), a["sd"][k].numpy(), rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(b["hist"], a["hist"], rtol=1e-3)
    np.testing.assert_allclose(b["tao"], a["tao"], atol=2.0 / 104)
    m = ~torch.isnan(a["proto"])
    assert torch.equal(torch.isnan(a["proto"]), torch.isnan(b["proto"]))
    np.testing.assert_allclose(b["proto"][m].numpy(), a["proto"][m].numpy(), rtol=5e-3, atol=5e-4)


class _ConvBackbone(nn.Module):
    def __init__(self, dim, feat_dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, feat_dim, 1)

    def forward(self, x):                       # x [B, dim] -> a [B, dim, 3, 3] "image"
        img = x[:, :, None, None] * torch.linspace(0.5, 1.5, 9, device=x.device).view(1, 1, 3, 3)
        return self.conv(img)


def test_local_update_fused_tail_tagging_matches_the_materialised_path(lib):
    """SURVEY 8f.1 in the drop-in: a FusedTail model tags through the pool+score kernel (no [N, D] matrix);
    same selections / masks as the same model run through the materialised feature path."""
    import fedmlp_b200 as F
    from fedmlp_b200.local_training import LocalUpdate

    class Plain(nn.Module):                     # same arithmetic, but without .tagging -> materialised path
        def __init__(self, fused):
            super().__init__()
            self.fused = fused

        def forward(self, x):
            return self.fused(x)

    args = _args()
    ds, class_neg_idx, dict_users, _ = _setup()
    torch.manual_seed(21)
    fused = F.FusedTail(_ConvBackbone(12, 64), nn.Linear(64, 5)).cuda()
    k, act, neg = 1, [1], [0, 2, 3, 4]
    proto_glob = torch.relu(torch.randn(10, 64, generator=torch.Generator().manual_seed(1))) + 0.05
    res = []
    for model in (fused, Plain(fused)):
        torch.manual_seed(5)
        local = LocalUpdate(args, k, deepcopy(ds), dict_users[k], class_neg_idx, class_neg_idx, active_class_list=[k])
        local.lr = 0.0                               # keep the model fixed: both passes see the same weights
        net = deepcopy(model)
        ret = local.train_FedMLP(2, [0] * 5, proto_glob, None, neg, act, net)
        res.append(dict(idx=local.traindata_idx, idxss=local.idxss, feat=local.last["tag_feat"], sim=local.tagger.sim.cpu().clone()))
    assert res[0]["feat"] is None and res[1]["feat"] is not None          # the fused pass never built [N, D]
    np.testing.assert_allclose(res[0]["sim"][neg].numpy(), res[1]["sim"][neg].numpy(), rtol=0, atol=1e-6)
    assert res[0]["idx"] == res[1]["idx"] and res[0]["idxss"] == res[1]["idxss"]
