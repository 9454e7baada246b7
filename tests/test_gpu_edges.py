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


def test_full_size_configs(F):
    """BASELINE.json configs at full size: ICH 55,000 x 1024 (C=5) and ChestXray14 85,000 x 1280
    (C=14, 13 missing classes): similarity and selection against the oracle."""
    for (N, D, C, signed, act) in [(55000, 1024, 5, False, 0), (85000, 1280, 14, True, 3)]:
        feat, labels, _ = O.synth_client(N, D, C, seed=N)
        proto = O.synth_prototypes(feat, labels)
        missing = [c for c in range(C) if c != act]
        ref = O.tag_similarity(feat, proto, missing)
        tb = F.TagBatch([0, N], C, [[act]], [missing], device=DEV)
        tb.step(feat.to(DEV), proto.to(DEV), 0.005, 0.01)
        sim = tb.sim.cpu().numpy()
        got = tb.traindata_idx(0)
        for i, c in enumerate(missing):
            r = ref[c].numpy()
            np.testing.assert_allclose(sim[c], r, rtol=0, atol=1e-6)
            sel_ref = O.split_and_select(r, 0.005, 0.01)
            for side, key in ((0, "clean"), (1, "noise")):
                mine, theirs = {int(v) for v in got[2 * i + side]}, set(sel_ref[key])
                # north_star waiver: differences only among rows within 1e-6 of the k-th similarity
                if mine != theirs:
                    kth = r[sel_ref[key][-1]]
                    assert all(abs(r[d] - kth) < 1e-6 for d in mine ^ theirs)
                assert len(mine) == len(theirs)
