"""Edge cases of the hot path on the GPU: empty and ragged segments, more clients than one launch
takes (grouping path), more classes than one similarity launch takes (class-split path), maximum
client counts, degenerate inputs."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import fedmlp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def F(lib):
    import fedmlp_b200
    return fedmlp_b200


def test_many_small_clients_grouping_path(F):
    """70 clients (> 64 per launch) incl. empty ones: similarity, selection, fill, prototypes."""
    C, D = 5, 128
    rng = np.random.default_rng(0)
    sizes = [int(v) for v in rng.integers(0, 40, size=70)]
    sizes[3] = 0; sizes[64] = 0; sizes[69] = 57
    seg_rows = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    S, N = len(sizes), seg_rows[-1]
    feat, labels, logits = O.synth_client(N, D, C, seed=1)
    proto = O.synth_prototypes(feat, labels)
    active = [[s % C] for s in range(S)]
    missing = [[c for c in range(C) if c != s % C] for s in range(S)]
    tb = F.TagBatch(seg_rows, C, active, missing, device=DEV)
    counts, sel, cap = tb.step(feat.to(DEV), proto.to(DEV), 0.2, 0.3)
    sim = tb.sim.cpu().numpy()
    y, distill, sup = tb.fill(labels.to(DEV))
    res = F.build_prototypes(feat.to(DEV), labels.to(DEV), logits.to(DEV), active, missing, 0.3, 0.7, True, seg_rows=seg_rows)
    for s in range(S):
        r0, r1 = seg_rows[s], seg_rows[s + 1]
        got = tb.traindata_idx(s)
        if r1 == r0:
            assert all(len(l) == 0 for l in got)
            assert counts[s].sum() == 0
            continue
        ref = O.tag_similarity(feat[r0:r1], proto, missing[s])
        st = O.TaggingState(list(range(r1 - r0)), missing[s])
        st.step(feat[r0:r1], proto, 0.2, 0.3)
        for i, c in enumerate(missing[s]):
            np.testing.assert_allclose(sim[c, r0:r1], ref[c].numpy(), rtol=0, atol=1e-6)
            assert got[2 * i] == [float(v) for v in st.traindata_idx[2 * i]]
            assert got[2 * i + 1] == [float(v) for v in st.traindata_idx[2 * i + 1]]
        tgt, dis = O.mask_fill(labels[r0:r1].numpy(), list(range(r1 - r0)), active[s], missing[s], st.traindata_idx)
        np.testing.assert_array_equal(y[r0:r1].cpu().numpy(), tgt)
        np.testing.assert_array_equal(distill[r0:r1].cpu().numpy(), dis)
        ref_p, ref_n, ref_t = O.prototype_build(feat[r0:r1], labels[r0:r1], logits[r0:r1], active[s], missing[s], 0.3, 0.7, True)
        np.testing.assert_allclose(res.proto[s].cpu().numpy(), ref_p.numpy(), rtol=1e-5, atol=1e-5)
        assert res.cnt[s].cpu().tolist() == ref_n


@pytest.mark.parametrize("mode", ["pair", "folded"])
def test_twenty_classes_split_path(F, mode):
    """C = 20 > 16 classes per similarity launch: the class set is split across launches."""
    N, D, C = 300, 96, 20
    feat, labels, _ = O.synth_client(N, D, C, seed=2)
    proto = O.synth_prototypes(feat, labels) + 0.01
    missing = [c for c in range(C) if c != 7]
    ref = O.tag_similarity(feat, proto, missing)
    sim = F.tag_similarity(feat.to(DEV), proto.to(DEV), missing, mode=mode).cpu().numpy()
    for c in missing:
        np.testing.assert_allclose(sim[c], ref[c].numpy(), rtol=0, atol=1e-6)
    assert np.isnan(sim[7]).all()


def test_select_degenerate_inputs(F):
    C = 3
    # all-NaN similarities, all-equal similarities, a single row, everything already tagged
    cases = {
        "nan": np.full((C, 50), np.nan, dtype=np.float32),
        "equal": np.full((C, 50), 0.25, dtype=np.float32),
        "single": np.array([[0.5], [-0.5], [0.0]], dtype=np.float32),
    }
    for name, sims in cases.items():
        n = sims.shape[1]
        tb = F.TagBatch([0, n], C, [[]], [[0, 1, 2]], device=DEV)
        tb.sim.copy_(torch.from_numpy(sims).to(DEV))
        counts, sel, cap = tb.select(0.5, 0.5)
        counts, sel = counts.cpu().numpy(), sel.cpu().numpy()
        for c in range(C):
            s = sims[c]
            ok = ~np.isnan(s)
            ref = O.split_and_select(np.where(ok, s, 0), 0.5, 0.5, valid=ok)
            assert counts[0, c].tolist() == [ref["n_clean"], ref["n_noise"], ref["m"], ref["k"]], name
            assert sel[0, c, 0, :ref["m"]].tolist() == ref["clean"], name
            assert sel[0, c, 1, :ref["k"]].tolist() == ref["noise"], name
    tb = F.TagBatch([0, 20], C, [[]], [[0, 1, 2]], device=DEV)
    tb.sim.copy_(torch.randn(C, 20).to(DEV))
    tb.tag.fill_(1)
    counts, _, _ = tb.select(0.5, 0.5)
    assert int(counts.sum()) == 0 and int(tb.remaining_count.sum()) == 0


def test_fedavg_max_clients_and_empty(F):
    g = torch.Generator().manual_seed(3)
    K, P = 64, 4099
    bufs = [torch.randn(P, generator=g) for _ in range(K)]
    w = list(range(1, K + 1))
    acc = bufs[0] * w[0]
    for b, n in zip(bufs[1:], w[1:]):
        acc += b * n
    ref = acc / sum(w)
    out = F.fedavg_flat_buffers([b.to(DEV) for b in bufs], w)
    assert torch.equal(out.cpu(), ref)
    # a state_dict with an empty tensor and a 0-dim fp32 scalar
    clients = [OrderedDict(a=torch.randn(0), b=torch.randn((), generator=g), c=torch.randn(5, generator=g)) for _ in range(3)]
    ref = O.fedavg(clients, [1, 2, 3])
    out = F.FedAvg([OrderedDict((k, v.to(DEV)) for k, v in c.items()) for c in clients], [1, 2, 3])
    for k in ref:
        assert out[k].shape == ref[k].shape and torch.equal(out[k].cpu(), ref[k])
    with pytest.raises(KeyError):
        F.FedAvg([OrderedDict(a=torch.zeros(1, device=DEV)), OrderedDict(b=torch.zeros(1, device=DEV))], [1, 1])
    with pytest.raises(TypeError):
        F.FedAvg([OrderedDict(a=torch.zeros(2, device=DEV, dtype=torch.float16))], [1])


@pytest.mark.parametrize("mode", ["pair", "folded"])
@pytest.mark.parametrize("N,D,C,signed,act", [(55000, 1024, 5, False, 0), (85000, 1280, 14, True, 3)])
def test_full_size_configs(F, mode, N, D, C, signed, act):
    """BASELINE.json configs at full size: ICH 55,000 x 1024 (C=5, post-ReLU features) and ChestXray14
    85,000 x 1280 (C=14, 13 missing classes, EfficientNet-style SIGNED features), in both similarity modes
    (folded is what the batched round / bench runs): similarity and selection against the oracle."""
    feat, labels, _ = O.synth_client(N, D, C, seed=N, signed=signed)
    proto = O.synth_prototypes(feat, labels)
    missing = [c for c in range(C) if c != act]
    ref = O.tag_similarity(feat, proto, missing)
    tb = F.TagBatch([0, N], C, [[act]], [missing], device=DEV)
    tb.step(feat.to(DEV), proto.to(DEV), 0.005, 0.01, mode=mode)
    sim = tb.sim.cpu().numpy()
    got = tb.traindata_idx(0)
    for i, c in enumerate(missing):
        r = ref[c].numpy()
        np.testing.assert_allclose(sim[c], r, rtol=0, atol=1e-6)
        sel_ref = O.split_and_select(r, 0.005, 0.01)
        for side, key in ((0, "clean"), (1, "noise")):
            mine, theirs = {int(v) for v in got[2 * i + side]}, set(sel_ref[key])
            # north_star waiver: differences only among rows within 1e-6 of the k-th similarity (or of the sign
            # threshold 0, which moves a row between the two sides and with it the counts by at most the
            # number of such rows)
            near0 = int((np.abs(r) < 1e-6).sum())
            if mine != theirs:
                kth = r[sel_ref[key][-1]] if sel_ref[key] else 0.0
                assert all(abs(r[d] - kth) < 1e-6 or abs(r[d]) < 1e-6 for d in mine ^ theirs)
            assert abs(len(mine) - len(theirs)) <= near0


# ---------------------------------------------------------------------------------------------------------
# Round-1 advisor findings: client-count limits, model_dist operands, flat-view safety, FlatAdam
def test_fedavg_more_than_64_clients_with_int64_counters(F):
    """FedAvg over 70 clients with integer dict_len and BatchNorm counters (the normal main.py case), scattered
    and flat: bit-exact incl. the int64 -> float32 quirk (utils/FedAvg.py:9-13)."""
    g = torch.Generator().manual_seed(3)
    K = 70
    clients = [OrderedDict(w=torch.randn(33, 7, generator=g), b=torch.randn(5, generator=g),
                           n=torch.tensor(1000 + 37 * k, dtype=torch.int64)) for k in range(K)]
    dict_len = [4000 + 13 * k for k in range(K)]
    ref = O.fedavg(clients, dict_len)
    cu = [OrderedDict((k, v.to(DEV)) for k, v in c.items()) for c in clients]
    for inputs in (cu, [F.FlatStateDict.from_state_dict(c) for c in cu]):
        out = F.FedAvg(inputs, dict_len)
        for k in ref:
            assert out[k].dtype == ref[k].dtype
            assert torch.equal(out[k].cpu(), ref[k]), k


def test_fedavg_proto_and_tao_more_than_64_clients(F):
    K, C, D = 150, 5, 64
    g = torch.Generator().manual_seed(5)
    protos = [torch.randn(2 * C, D, generator=g) for _ in range(K)]
    weights = [3000 + 7 * k for k in range(K)]
    active = [[k for k in range(K) if k % C == c] for c in range(C)]
    active[2] = []                                            # nobody annotates class 2 -> NaN rows
    active[4] = [k for k in range(K) if k % C == 4 and k >= 64]   # only clients of the later groups
    ref = O.fedavg_proto(protos, weights, active)
    out = F.FedAvg_proto([p.to(DEV) for p in protos], weights, active).cpu()
    ok = ~torch.isnan(ref)
    assert torch.isnan(out[~ok]).all() and not torch.isnan(out[ok]).any()
    np.testing.assert_allclose(out[ok].numpy(), ref[ok].numpy(), rtol=1e-5, atol=1e-6)
    taos = [torch.rand(C, generator=g).double().numpy() for _ in range(K)]
    missing = [[k for k in range(K) if k % C != c] for c in range(C)]
    missing[1] = []
    np.testing.assert_allclose(F.FedAvg_tao(taos, weights, missing), O.fedavg_tao(taos, weights, missing), rtol=1e-12)
    np.testing.assert_allclose(F.FedAvg_tao(taos, weights), O.fedavg_tao(taos, weights), rtol=1e-12)


def test_model_dist_noncontiguous_and_rscfed_with_counters(F):
    """model_dist on channels_last (non-contiguous) conv weights, and RSCFed on state_dicts WITH int64 BatchNorm
    counters: model_dist(w_i, Fed_w output) pairs an int64 entry with its float32 average (FedNoRo.py:106-115)."""
    g = torch.Generator().manual_seed(9)

    def client(k):
        return OrderedDict([("conv.weight", torch.randn(8, 4, 3, 3, generator=g)), ("bn.weight", torch.randn(8, generator=g)),
                            ("bn.num_batches_tracked", torch.tensor(10 + k, dtype=torch.int64)),
                            ("fc.weight", torch.randn(5, 8, generator=g))])
    clients = [client(k) for k in range(6)]
    cl_last = [OrderedDict((k, (v.to(DEV).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v.to(DEV)))
                           for k, v in c.items()) for c in clients]
    assert not cl_last[0]["conv.weight"].is_contiguous()
    ref = O.model_dist(clients[0], clients[1])
    for _ in range(20):          # repeated calls recycle allocator blocks: a dangling operand would show up here
        assert abs(F.model_dist(cl_last[0], cl_last[1]) - ref) <= 1e-5 * ref
    cu = [OrderedDict((k, v.to(DEV)) for k, v in c.items()) for c in clients]
    dma, dict_len = [[0, 1, 2], [3, 4, 5], [0, 2, 4], [1, 3, 5]], [500, 400, 300, 200, 100, 50]
    ref_r = O.rscfed(dma, clients, 3, dict_len, 4)
    out = F.RSCFed(dma, cu, 3, dict_len, 4)
    for k in ref_r:
        np.testing.assert_allclose(out[k].cpu().numpy(), ref_r[k].numpy(), rtol=1e-5, atol=1e-6)


def test_flat_state_dict_setitem_and_flat_view_checks(F):
    from fedmlp_b200.flat import flat_view_of
    sd = OrderedDict(a=torch.randn(6, device=DEV), b=torch.randn(3, 3, device=DEV))
    f1, f2 = F.FlatStateDict.from_state_dict(sd), F.FlatStateDict.from_state_dict(sd)
    new_a = torch.full((6,), 2.0, device=DEV)
    f1["a"] = new_a                                  # copies into the flat view: FedAvg's flat path sees it
    assert f1["a"].data_ptr() == f1.flat_f32.data_ptr() and torch.equal(f1["a"], new_a)
    out = F.FedAvg([f1, f2], [1, 1])
    assert torch.equal(out["a"].cpu(), ((new_a * 1 + sd["a"] * 1) / 2).cpu())
    with pytest.raises(KeyError):
        f1["zzz"] = new_a
    # a single-tensor dict whose storage ends with the tensor: the padded flat read would run past it
    lone = OrderedDict(w=torch.randn(6, device=DEV))
    assert flat_view_of(lone) is None or lone["w"].untyped_storage().nbytes() >= 32
    res = F.FedAvg([lone, OrderedDict(w=torch.ones(6, device=DEV))], [1, 3])
    assert torch.equal(res["w"].cpu(), ((lone["w"] * 1 + torch.ones(6, device=DEV) * 3) / 4).cpu())


def test_flat_adam_rebinds_after_zero_grad_set_to_none(F):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.ReLU(), torch.nn.Linear(8, 4)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.ReLU(), torch.nn.Linear(8, 4)).to(DEV)
    ref.load_state_dict(net.state_dict())
    net[2].bias.requires_grad_(False); ref[2].bias.requires_grad_(False)      # frozen: untouched by both
    opt = F.FlatAdam(net, lr=1e-2, weight_decay=5e-4)
    topt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=1e-2, weight_decay=5e-4)
    x = torch.randn(32, 16, device=DEV)
    for it in range(4):
        net.zero_grad()                              # set_to_none=True: detaches the flat gradient views
        topt.zero_grad()
        net(x).square().mean().backward()
        ref(x).square().mean().backward()
        opt.step()
        topt.step()
    for a, b in zip(net.parameters(), ref.parameters()):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=1e-5, atol=1e-7)
