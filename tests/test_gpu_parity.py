"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star): pseudo-label masks and selected indices bit-exact except
where |similarity - threshold| < 1e-6; prototypes, aggregated weights and losses within 1e-5
relative (fp32).  FedAvg on one GPU is additionally bit-exact against the CPU oracle because the
kernel folds in the reference's order with IEEE mul/add/div.
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import fedmlp_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def F(lib):
    import fedmlp_b200
    return fedmlp_b200


def cuda(x):
    return torch.as_tensor(np.array(x)).to(DEV) if not isinstance(x, torch.Tensor) else x.to(DEV)


# =============================================================================== K1 FedAvg
@pytest.mark.parametrize("tag", ["int", "float", "float_odd"])
def test_fedavg_golden(F, tag):
    z = gu.load("fedavg.npz")
    names = [str(n) for n in z["names"]]
    clients = [OrderedDict((n, cuda(z[f"in/{k}/{n}"].copy())) for n in names) for k in range(4)]
    w = z[f"weights/{tag}"].tolist()
    if tag == "int":
        w = [int(v) for v in w]
    out = F.FedAvg(clients, w)                       # scattered tensors -> multi-tensor kernel
    flat = F.FedAvg([F.FlatStateDict.from_state_dict(c) for c in clients], w)   # flat kernel
    assert list(out.keys()) == names == list(flat.keys())
    for n in names:
        ref = z[f"out/{tag}/{n}"]
        for o in (out, flat):
            assert o[n].dtype == torch.float32 and o[n].is_cuda and tuple(o[n].shape) == ref.shape
            np.testing.assert_array_equal(o[n].cpu().numpy(), ref)


def _densenet_like_state(seed, base=None):
    g = torch.Generator().manual_seed(seed)
    shapes = [("conv0.weight", (64, 3, 7, 7)), ("norm0.weight", (64,)), ("norm0.bias", (64,)),
              ("norm0.running_mean", (64,)), ("norm0.running_var", (64,)), ("norm0.num_batches_tracked", ()),
              ("block.conv1.weight", (128, 64, 1, 1)), ("block.norm1.weight", (128,)),
              ("block.norm1.num_batches_tracked", ()), ("block.conv2.weight", (32, 128, 3, 3)),
              ("odd.weight", (1023,)), ("odd2.weight", (7, 13)), ("classifier.weight", (5, 1024)),
              ("classifier.bias", (5,))]
    sd = OrderedDict()
    for n, shp in shapes:
        if "num_batches" in n:
            sd[n] = torch.tensor(100 + seed, dtype=torch.int64)
        else:
            b = base[n] if base is not None else torch.zeros(shp)
            sd[n] = b + 0.02 * torch.randn(shp, generator=g)
    return sd


@pytest.mark.parametrize("K", [1, 5, 8, 11])
def test_fedavg_state_dicts_bit_exact(F, K):
    base = _densenet_like_state(0)
    clients = [_densenet_like_state(10 + k, base) for k in range(K)]
    w = [5000 + 3 * k for k in range(K)]
    ref = O.fedavg(clients, w)
    gpu_clients = [OrderedDict((n, v.to(DEV)) for n, v in c.items()) for c in clients]
    out = F.FedAvg(gpu_clients, w)
    flat = F.FedAvg([F.FlatStateDict.from_state_dict(c) for c in gpu_clients], w)
    for n in ref:
        np.testing.assert_array_equal(out[n].cpu().numpy(), ref[n].numpy())
        np.testing.assert_array_equal(flat[n].cpu().numpy(), ref[n].numpy())
    # CPU inputs come back as CPU tensors with the same values
    cpu_out = F.FedAvg(clients, w)
    for n in ref:
        assert not cpu_out[n].is_cuda
        np.testing.assert_array_equal(cpu_out[n].numpy(), ref[n].numpy())
    # int64 counters come out float32 (true-division quirk)
    assert out["norm0.num_batches_tracked"].dtype == torch.float32


@pytest.mark.parametrize("K,P,off", [(3, 1, 0), (8, 4096, 0), (8, 100003, 0), (8, 100003, 1), (70, 50001, 0), (130, 7777, 3)])
def test_fedavg_flat_buffers(F, K, P, off):
    """Flat kernel incl. K > 64 chaining (reference order kept), odd lengths and misaligned bases."""
    g = torch.Generator().manual_seed(K * 1000 + P)
    bufs = [torch.randn(P + 8, generator=g) for _ in range(K)]
    w = [float(v) for v in (torch.rand(K, generator=g) * 3 + 0.5)]
    acc = bufs[0][off:off + P] * w[0]
    for i in range(1, K):
        acc += bufs[i][off:off + P] * w[i]
    ref = acc / sum(w)
    dbufs = [b.to(DEV)[off:off + P] for b in bufs]
    out = F.fedavg_flat_buffers(dbufs, w)
    np.testing.assert_array_equal(out.cpu().numpy(), ref.numpy())


def test_fedavg_proto_golden(F):
    z = gu.load("fedavg.npz")
    protos = [cuda(p.copy()) for p in z["proto/in"]]
    out = F.FedAvg_proto(protos, z["proto/weight"].tolist(), gu.parse_lists(z["proto/lists"]))
    np.testing.assert_array_equal(out.cpu().numpy(), z["proto/out"])
    out_cpu = F.FedAvg_proto([p.cpu() for p in protos], z["proto/weight"].tolist(), gu.parse_lists(z["proto/lists"]))
    assert not out_cpu.is_cuda
    np.testing.assert_array_equal(out_cpu.numpy(), z["proto/out"])
    taos = [t.copy() for t in z["tao/in"]]
    np.testing.assert_array_equal(F.FedAvg_tao(taos, z["proto/weight"].tolist(), gu.parse_lists(z["tao/lists"])),
                                  z["tao/out_lists"])


# =============================================================================== K3 similarity
@pytest.mark.parametrize("mode", ["pair", "folded"])
def test_tag_sim_golden(F, mode):
    z = gu.load("tagging.npz")
    feat, proto = cuda(z["feat"].copy()), cuda(z["proto"].copy())
    C = proto.shape[0] // 2
    sim = F.tag_similarity(feat, proto, list(range(C)), mode=mode).cpu().numpy()
    for c in range(C):
        np.testing.assert_allclose(sim[c], z[f"sim/{c}"], rtol=0, atol=1e-6)
    # only the requested classes are written
    part = F.tag_similarity(feat, proto, [1, 3]).cpu().numpy()
    assert np.isnan(part[[0, 2, 4]]).all()
    np.testing.assert_allclose(part[[1, 3]], np.stack([z["sim/1"], z["sim/3"]]), rtol=0, atol=1e-6)


@pytest.mark.parametrize("N,D,C,signed", [(6875, 1024, 5, False), (3001, 1280, 14, True), (517, 100, 3, False), (5, 1024, 5, False)])
@pytest.mark.parametrize("mode", ["pair", "folded"])
def test_tag_sim_vs_oracle(F, N, D, C, signed, mode):
    feat, labels, _ = O.synth_client(N, D, C, seed=N + D, signed=signed)
    proto = O.synth_prototypes(feat, labels)
    missing = [c for c in range(C) if c != 1]
    ref = O.tag_similarity(feat, proto, missing)
    sim = F.tag_similarity(feat.to(DEV), proto.to(DEV), missing, mode=mode).cpu().numpy()
    for c in missing:
        np.testing.assert_allclose(sim[c], ref[c].numpy(), rtol=0, atol=1e-6)
    assert np.isnan(sim[1]).all()


def test_tag_sim_segments(F):
    """Batched clients: each segment scores its own missing classes against the shared prototypes."""
    C, D = 5, 1024
    sizes = [700, 1, 1333, 64]
    feats, labs = [], []
    for s, n in enumerate(sizes):
        f, l, _ = O.synth_client(n, D, C, seed=50 + s)
        feats.append(f); labs.append(l)
    feat, lab = torch.cat(feats), torch.cat(labs)
    proto = O.synth_prototypes(feat, lab)
    seg_rows = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    missing = [[c for c in range(C) if c != s % C] for s in range(len(sizes))]
    sim = F.tag_similarity(feat.to(DEV), proto.to(DEV), missing, seg_rows=seg_rows).cpu().numpy()
    for s in range(len(sizes)):
        r0, r1 = seg_rows[s], seg_rows[s + 1]
        ref = O.tag_similarity(feat[r0:r1], proto, missing[s])
        for c in range(C):
            if c in missing[s]:
                np.testing.assert_allclose(sim[c, r0:r1], ref[c].numpy(), rtol=0, atol=1e-6)
            else:
                assert np.isnan(sim[c, r0:r1]).all()


# =============================================================================== K3b selection
def _select_on(F, sims, missing, C, clean_frac, noise_frac, seg_rows=None, tag_init=None):
    n = sims.shape[1]
    seg_rows = seg_rows or [0, n]
    S = len(seg_rows) - 1
    act = [[c for c in range(C) if c not in missing[s]] for s in range(S)]
    tb = F.TagBatch(seg_rows, C, act, missing, device=DEV)
    tb.sim.copy_(cuda(sims))
    if tag_init is not None:
        tb.tag.copy_(cuda(tag_init))
    counts, sel, cap = tb.select(clean_frac, noise_frac)
    return tb, counts.cpu().numpy(), sel.cpu().numpy()


def test_select_ties_golden(F):
    z = gu.load("tagging.npz")
    vals = z["ties/vals"]
    n_clean, n_noise = int(np.sum(vals >= 0)), int(np.sum(vals < 0))
    for m in range(n_clean + 1):
        for k in range(n_noise + 1):
            cf, nf = min(1.0, (m + 0.5) / n_clean), min(1.0, (k + 0.5) / n_noise)
            tb, counts, sel = _select_on(F, vals[None, :], [[0]], 1, cf, nf)
            assert counts[0, 0].tolist() == [n_clean, n_noise, m, k]
            assert sel[0, 0, 0, :m].tolist() == z[f"ties/max/{m}"].tolist()
            assert sel[0, 0, 1, :k].tolist() == z[f"ties/min/{k}"].tolist()
            tag = tb.tag.cpu().numpy()[0]
            assert sorted(np.nonzero(tag == 1)[0].tolist()) == sorted(z[f"ties/max/{m}"].tolist())
            assert sorted(np.nonzero(tag == 2)[0].tolist()) == sorted(z[f"ties/min/{k}"].tolist())


def test_select_golden_sims(F):
    z = gu.load("tagging.npz")
    C = 5
    sims = np.stack([z[f"sim/{c}"] for c in range(C)])
    for n in (0, 1, 3, 10, 40):
        for c in range(C):
            nc, nn = int(np.sum(sims[c] >= 0)), int(np.sum(sims[c] < 0))
            if n > nc or n > nn:
                continue
            tb, counts, sel = _select_on(F, sims, [[c]], C, (n + 0.5) / nc, (n + 0.5) / nn)
            assert counts[0, c].tolist() == [nc, nn, n, n]
            assert sel[0, c, 0, :n].tolist() == z[f"max/{c}/{n}"].tolist()
            assert sel[0, c, 1, :n].tolist() == z[f"min/{c}/{n}"].tolist()


@pytest.fixture
def select_cluster(lib, request):
    """Forces the number of CTAs (one thread-block cluster) per (client, class) item of the selection kernel."""
    from fedmlp_b200 import _cabi as cabi
    lib.fmlp_set_tuning(cabi.TUNE_SELECT_CLUSTER, request.param)
    yield request.param
    lib.fmlp_set_tuning(cabi.TUNE_SELECT_CLUSTER, -1)


@pytest.mark.parametrize("select_cluster", [0, 2, 8], indirect=True)
def test_select_ties_golden_clustered(F, select_cluster):
    """The reference's tie order (utils/utils.py:24-35) when an item is split over the CTAs of a cluster."""
    z = gu.load("tagging.npz")
    vals = z["ties/vals"]
    n_clean, n_noise = int(np.sum(vals >= 0)), int(np.sum(vals < 0))
    for m in (0, 1, n_clean // 2, n_clean):
        for k in (0, 1, n_noise // 2, n_noise):
            cf, nf = min(1.0, (m + 0.5) / n_clean), min(1.0, (k + 0.5) / n_noise)
            tb, counts, sel = _select_on(F, vals[None, :], [[0]], 1, cf, nf)
            assert counts[0, 0].tolist() == [n_clean, n_noise, m, k]
            assert sel[0, 0, 0, :m].tolist() == z[f"ties/max/{m}"].tolist()
            assert sel[0, 0, 1, :k].tolist() == z[f"ties/min/{k}"].tolist()


@pytest.mark.parametrize("select_cluster", [0, 2, 8], indirect=True)
@pytest.mark.parametrize("N,cf,nf", [(55000, 0.005, 0.01), (6875, 0.005, 0.01), (85000, 0.005, 0.01), (2000, 0.3, 0.6), (1000, 1.0, 1.0), (37, 0.0, 0.0),
                                     (150000, 0.005, 0.01)])
def test_select_vs_oracle_random(F, N, cf, nf, select_cluster):
    """select_cluster 0 = automatic (1 CTA up to 16,384 rows, then a cluster of 2 / 4 / 8; 150,000 rows: 8 CTAs that
    re-read their keys), 2 / 8 = forced split."""
    rng = np.random.default_rng(N)
    C = 4
    sims = rng.normal(0, 0.02, size=(C, N)).astype(np.float32)
    sims[:, rng.integers(0, N, size=N // 10)] = sims[:, rng.integers(0, N, size=N // 10)]   # exact duplicates
    sims[0, :5] = [0.0, -0.0, np.nan, 0.0, -0.0]
    valid = rng.random((C, N)) < 0.8
    tag_init = np.where(valid, 0, rng.integers(1, 3, size=(C, N))).astype(np.uint8)
    tb, counts, sel = _select_on(F, sims, [[0, 1, 2, 3]], C, cf, nf, tag_init=tag_init)
    tag = tb.tag.cpu().numpy()
    for c in range(C):
        s = np.where(np.isnan(sims[c]), np.float32(0), sims[c])
        v = valid[c] & ~np.isnan(sims[c])
        ref = O.split_and_select(s, cf, nf, valid=v)
        assert counts[0, c].tolist() == [ref["n_clean"], ref["n_noise"], ref["m"], ref["k"]]
        assert sel[0, c, 0, :ref["m"]].tolist() == ref["clean"]
        assert sel[0, c, 1, :ref["k"]].tolist() == ref["noise"]
        exp = tag_init[c].copy()
        exp[ref["clean"]] = 1
        exp[ref["noise"]] = 2
        np.testing.assert_array_equal(tag[c], exp)
        assert int(tb.remaining_count[0, c]) == int((exp == 0).sum())      # = number of distill entries


def _assert_boundary_waiver(got, ref, ids, sims, stats):
    diff = set(int(v) for v in got) ^ set(int(v) for v in ref)
    pos = {d: p for p, d in enumerate(ids)}
    picked = [sims[p] for p in stats["clean"] + stats["noise"]]
    for d in diff:
        s = sims[pos[d]]
        assert abs(s) < 1e-6 or min(abs(s - q) for q in picked) < 1e-6, f"index {d} differs away from a threshold"


def test_select_segments_and_rounds(F):
    """Three tagging rounds over 3 batched clients == three independent oracle TaggingStates."""
    C, D = 5, 256
    sizes = [900, 333, 1201]
    seg_rows = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    missing = [[c for c in range(C) if c != s] for s in range(3)]
    active = [[s] for s in range(3)]
    ids = torch.cat([torch.randperm(n, generator=torch.Generator().manual_seed(n)) + 10000 * s for s, n in enumerate(sizes)])
    tb = F.TagBatch(seg_rows, C, active, missing, dataset_idx=ids, device=DEV)
    states = [O.TaggingState(ids[seg_rows[s]:seg_rows[s + 1]].tolist(), missing[s]) for s in range(3)]
    labs = None
    for rnd in range(3):
        fs, ls = [], []
        for s, n in enumerate(sizes):
            f, l, _ = O.synth_client(n, D, C, seed=77 * rnd + s)
            fs.append(f); ls.append(l)
        feat, labs = torch.cat(fs), torch.cat(ls)
        proto = O.synth_prototypes(feat, labs)
        tb.step(feat.to(DEV), proto.to(DEV), 0.02, 0.05)
        for s in range(3):
            seg_ids = ids[seg_rows[s]:seg_rows[s + 1]].tolist()
            sims_ref, stats = states[s].step(feat[seg_rows[s]:seg_rows[s + 1]], proto, 0.02, 0.05)
            got = tb.traindata_idx(s)
            for i, c in enumerate(missing[s]):
                exact = (got[2 * i] == [float(v) for v in states[s].traindata_idx[2 * i]]
                         and got[2 * i + 1] == [float(v) for v in states[s].traindata_idx[2 * i + 1]])
                if not exact:   # waiver: only rows whose similarity is within 1e-6 of a threshold may differ
                    _assert_boundary_waiver(got[2 * i] + got[2 * i + 1],
                                            states[s].traindata_idx[2 * i] + states[s].traindata_idx[2 * i + 1],
                                            seg_ids, sims_ref[c].numpy(), stats[i])
                    # keep the two states in lock-step for the next round
                    states[s].traindata_idx[2 * i] = [int(v) for v in got[2 * i]]
                    states[s].traindata_idx[2 * i + 1] = [int(v) for v in got[2 * i + 1]]
            rem = tb.remaining(s)
            for i, c in enumerate(missing[s]):
                tagged = {int(v) for v in got[2 * i]} | {int(v) for v in got[2 * i + 1]}
                assert rem[i] == sorted(set(seg_ids) - tagged)
    # label / mask fill from the final state
    y, distill, sup = tb.fill(labs.to(DEV))
    for s in range(3):
        r0, r1 = seg_rows[s], seg_rows[s + 1]
        tgt, dis = O.mask_fill(labs[r0:r1].numpy(), ids[r0:r1].tolist(), active[s], missing[s],
                               [[int(v) for v in l] for l in tb.traindata_idx(s)])
        np.testing.assert_array_equal(y[r0:r1].cpu().numpy(), tgt)
        np.testing.assert_array_equal(distill[r0:r1].cpu().numpy(), dis)
        np.testing.assert_array_equal(sup[r0:r1].cpu().numpy(), 1 - dis)


def test_mask_fill_golden(F):
    z = gu.load("maskfill.npz")
    idxs = z["idxs"].tolist()
    tdi = [[float(v) for v in s.split(",") if v] for s in z["traindata_idx"].tolist()]
    active, negative = z["active"].tolist(), z["negative"].tolist()
    C = 5
    tb = F.TagBatch([0, len(idxs)], C, [active], [negative], dataset_idx=torch.tensor(idxs), device=DEV)
    tag = np.zeros((C, len(idxs)), dtype=np.uint8)
    pos = {d: p for p, d in enumerate(idxs)}
    for i, c in enumerate(negative):
        for d in tdi[2 * i]:
            if int(d) in pos:
                tag[c, pos[int(d)]] = 1
        for d in tdi[2 * i + 1]:
            if int(d) in pos:
                tag[c, pos[int(d)]] = 2     # the noise list wins when an index is in both (:1463-1466)
    tb.tag.copy_(cuda(tag))
    y, distill, sup = tb.fill(cuda(z["targets_true"][idxs]))
    np.testing.assert_array_equal(y.cpu().numpy(), z["out_target"])
    np.testing.assert_array_equal(distill.cpu().numpy(), z["out_distill"])


# =============================================================================== K2 prototypes
@pytest.mark.parametrize("N,D,C,signed", [(5000, 1024, 5, False), (3001, 1280, 14, True), (40000, 1024, 5, False), (130, 64, 3, False)])
def test_prototypes_vs_oracle(F, N, D, C, signed):
    feat, labels, logits = O.synth_client(N, D, C, seed=3 * N + C, signed=signed)
    active, missing = [2], [c for c in range(C) if c != 2]
    ref_p, ref_n, ref_t = O.prototype_build(feat, labels, logits, active, missing, 0.3, 0.7, guard_empty=True)
    res = F.build_prototypes(feat.to(DEV), labels.to(DEV), logits.to(DEV), active, missing, 0.3, 0.7, guard_empty=True)
    p = res.proto[0].cpu().numpy()
    scale = float(ref_p.abs().max())
    np.testing.assert_allclose(p, ref_p.numpy(), rtol=1e-5, atol=1e-5 * scale)
    assert res.cnt[0].cpu().tolist() == ref_n
    # counts may differ only for probabilities within 1e-6 of L / U (sigmoid ulp differences)
    probs = torch.sigmoid(logits)
    near = ((probs - 0.3).abs() < 1e-6) | ((probs - 0.7).abs() < 1e-6)
    got_t = res.tcnt[0].cpu().numpy()
    exp_t = np.round(ref_t * N).astype(np.int64)
    assert np.all(np.abs(got_t - exp_t) <= near.sum(0).numpy())
    assert np.all(p[[r for r in range(2 * C) if r // 2 not in active]] == 0)


def test_prototypes_empty_group_and_multi_active(F):
    N, D, C = 700, 128, 6
    feat, labels, logits = O.synth_client(N, D, C, seed=9)
    labels[:, 4] = 0.0                                  # class 4 has no positives
    active = [0, 1, 3, 4, 5]                            # 5 active classes -> two passes of the kernel
    for guard in (True, False):
        ref_p, ref_n, ref_t = O.prototype_build(feat, labels, logits, active, [2], 0.3, 0.7, guard_empty=guard)
        res = F.build_prototypes(feat.to(DEV), labels.to(DEV), logits.to(DEV), active, [2], 0.3, 0.7, guard_empty=guard)
        p = res.proto[0].cpu().numpy()
        assert res.cnt[0].cpu().tolist() == ref_n
        if guard:
            assert np.all(p[9] == 0)
        else:
            assert np.isnan(p[9]).all() and np.isnan(ref_p[9].numpy()).all()
        ok = ~np.isnan(ref_p.numpy())
        np.testing.assert_allclose(p[ok], ref_p.numpy()[ok], rtol=1e-5,
                                   atol=1e-5 * float(ref_p[~torch.isnan(ref_p)].abs().max()))


def test_prototypes_segments(F):
    C, D = 5, 1024
    sizes = [1500, 7, 2222, 129]
    fs, ls, zs = [], [], []
    for s, n in enumerate(sizes):
        f, l, z = O.synth_client(n, D, C, seed=200 + s)
        fs.append(f); ls.append(l); zs.append(z)
    seg_rows = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    active = [[s % C] for s in range(4)]
    tcls = [[c for c in range(C) if c != s % C] for s in range(4)]
    res = F.build_prototypes(torch.cat(fs).to(DEV), torch.cat(ls).to(DEV), torch.cat(zs).to(DEV), active, tcls,
                             0.3, 0.7, guard_empty=True, seg_rows=seg_rows)
    t = res.t()
    for s in range(4):
        ref_p, ref_n, ref_t = O.prototype_build(fs[s], ls[s], zs[s], active[s], tcls[s], 0.3, 0.7, guard_empty=True)
        np.testing.assert_allclose(res.proto[s].cpu().numpy(), ref_p.numpy(), rtol=1e-5, atol=1e-5 * float(ref_p.abs().max()))
        assert res.cnt[s].cpu().tolist() == ref_n
        np.testing.assert_allclose(t[s], ref_t, rtol=0, atol=2.0 / sizes[s])


# =============================================================================== K4 losses
def _rand_logits(B, C, seed, extreme=False):
    g = torch.Generator().manual_seed(seed)
    zs = [torch.randn(B, C, generator=g) * 2 for _ in range(4)]
    if extreme:
        for z in zs:
            z.view(-1)[::7] *= 15.0          # |z| up to ~100: exercises the log clamp at -100 and the 1e-12 eps
    y = (torch.rand(B, C, generator=g) < 0.3).float()
    return zs, y


@pytest.mark.parametrize("B,C,extreme", [(32, 5, False), (8, 5, False), (32, 14, True), (1, 1, False), (85000, 14, False), (1000, 8, True)])
def test_loss_stage1_vs_oracle(F, B, C, extreme):
    zs, y = _rand_logits(B, C, B + C, extreme)
    a = min(2, C - 1)
    active, missing = [a], [c for c in range(C) if c != a]
    bs = 32
    z1 = zs[0].to(DEV).requires_grad_(True)
    z2 = zs[1].to(DEV).requires_grad_(True)
    loss = F.fedmlp_stage1_loss(z1, z2, zs[2].to(DEV), zs[3].to(DEV), y.to(DEV), active, missing, bs)
    (loss * 1.0).backward()
    if C == 1:
        assert torch.isnan(loss)          # 0/0 of the empty missing set, like the reference (:958-959)
        return
    ref_loss, r1, r2 = O.loss_and_grads(lambda p, q, c, d, t: O.stage1_loss(p, q, c, d, t, active, missing, bs),
                                        *zs, y, n_grad=2)
    assert abs(float(loss.detach()) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    s1, s2 = float(r1.abs().max()), float(r2.abs().max())
    np.testing.assert_allclose(z1.grad.cpu().numpy(), r1.numpy(), rtol=1e-5, atol=1e-6 * s1)
    np.testing.assert_allclose(z2.grad.cpu().numpy(), r2.numpy(), rtol=1e-5, atol=1e-6 * s2)


@pytest.mark.parametrize("variant", ["sup", "sup_dis"])
@pytest.mark.parametrize("B,C,extreme", [(32, 5, False), (8, 14, True), (85000, 14, False), (3, 2, False)])
def test_loss_stage2_vs_oracle(F, B, C, extreme, variant):
    zs, y = _rand_logits(B, C, 7 * B + C, extreme)
    g = torch.Generator().manual_seed(B)
    distill = (torch.rand(B, C, generator=g) < 0.6).float()
    distill[:, 0] = 0
    ref_loss, rdz = O.loss_and_grads(lambda z, zg, t, d: O.stage2_loss(z, zg, t, d, variant), zs[0], zs[1], y, distill, n_grad=1)
    z = zs[0].to(DEV).requires_grad_(True)
    loss = F.fedmlp_stage2_loss(z, zs[1].to(DEV), y.to(DEV), distill.to(DEV), variant)
    (loss * 3.0).backward()                      # upstream gradient is applied on the device
    assert abs(float(loss.detach()) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    np.testing.assert_allclose(z.grad.cpu().numpy(), 3.0 * rdz.numpy(), rtol=1e-5, atol=3e-6 * float(rdz.abs().max()))


@pytest.mark.parametrize("variant", ["sup", "sup_dis"])
def test_loss_stage2_segmented(F, variant):
    """One launch over several clients == per-client oracle losses (own denominators)."""
    from fedmlp_b200.losses import launch_stage2_seg, LOSS2_VARIANTS
    C = 5
    sizes = [700, 0, 33, 6875, 1, 2049]
    seg_rows = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    N = seg_rows[-1]
    zs, y = _rand_logits(N, C, 99, extreme=True)
    g = torch.Generator().manual_seed(4)
    distill = (torch.rand(N, C, generator=g) < 0.6).float()
    distill[:, 1] = 0
    loss = torch.empty(len(sizes), device=DEV)
    dz = torch.empty(N, C, device=DEV)
    launch_stage2_seg(zs[0].to(DEV), zs[1].to(DEV), y.to(DEV), distill.to(DEV), seg_rows, LOSS2_VARIANTS[variant], loss, dz)
    # same launch with the per-(client, class) distill counts supplied (what fmlp_tag_select reports):
    # the kernel skips its counting phase and must give identical results
    counts = torch.stack([distill[seg_rows[s]:seg_rows[s + 1]].sum(0) for s in range(len(sizes))]).to(torch.int32)
    loss_b = torch.empty(len(sizes), device=DEV)
    dz_b = torch.empty(N, C, device=DEV)
    launch_stage2_seg(zs[0].to(DEV), zs[1].to(DEV), y.to(DEV), distill.to(DEV), seg_rows, LOSS2_VARIANTS[variant], loss_b, dz_b,
                      seg_class_distill=counts.to(DEV))
    assert torch.equal(dz_b, dz) and torch.equal(loss_b[~torch.isnan(loss_b)], loss[~torch.isnan(loss)])
    loss, dz = loss.cpu(), dz.cpu()
    for s, n in enumerate(sizes):
        r0, r1 = seg_rows[s], seg_rows[s + 1]
        if n == 0:
            assert torch.isnan(loss[s])
            continue
        ref_loss, rdz = O.loss_and_grads(lambda z, zg, t, d: O.stage2_loss(z, zg, t, d, variant),
                                         zs[0][r0:r1], zs[1][r0:r1], y[r0:r1], distill[r0:r1], n_grad=1)
        assert abs(float(loss[s]) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
        np.testing.assert_allclose(dz[r0:r1].numpy(), rdz.numpy(), rtol=1e-5, atol=1e-6 * float(rdz.abs().max()))


def test_losses_golden_flow(F):
    """Per-step losses and logit gradients recorded from the reference's own training loop."""
    flow = gu.load("flow.npz")
    act, neg, bs = flow["s1/act_list"].tolist(), flow["s1/neg_list"].tolist(), int(flow["meta/batch_size"])
    for st in gu.stage1_steps(gu.Trace(flow, "s1")):
        loss, dz1, dz2 = F.fused_loss_and_grad_stage1(cuda(st["z1"]), cuda(st["z2"]), cuda(st["z3"]), cuda(st["z4"]),
                                                      cuda(st["y"]), act, neg, bs)
        assert abs(float(loss) - st["loss"]) <= 1e-5 * abs(st["loss"])
        np.testing.assert_allclose(dz1.cpu().numpy(), st["dz1"], rtol=1e-5, atol=1e-6 * np.abs(st["dz1"]).max())
        np.testing.assert_allclose(dz2.cpu().numpy(), st["dz2"], rtol=1e-5, atol=1e-6 * np.abs(st["dz2"]).max())
    for r in range(2):
        for st in gu.stage2_round(gu.Trace(flow, f"s2_{r}"))["steps"]:
            loss, dz = F.fused_loss_and_grad_stage2(cuda(st["z"]), cuda(st["zg"]), cuda(st["y"]), cuda(st["distill"]))
            assert abs(float(loss) - st["loss"]) <= 1e-5 * abs(st["loss"])
            np.testing.assert_allclose(dz.cpu().numpy(), st["dz"], rtol=1e-5, atol=1e-6 * np.abs(st["dz"]).max())


# =============================================================================== recorded flow, end to end
@pytest.mark.parametrize("mode", ["pair", "folded"])
def test_golden_flow_tagging_and_prototypes(F, mode):
    """Replay the reference's recorded stage-2 rounds through the CUDA path: same selected
    indices (traindata_idx, in pick order), same remaining sets, same masks, same prototypes/t —
    with either similarity formulation."""
    flow = gu.load("flow.npz")
    neg, act = flow["s1/neg_list"].tolist(), flow["s1/act_list"].tolist()
    C = int(flow["meta/C"])
    proto_glob = cuda(flow["proto_glob"].copy())
    cf, nf = float(flow["meta/clean_threshold"]), float(flow["meta/noise_threshold"])
    L, U = float(flow["meta/L"]), float(flow["meta/U"])
    # stage-1 prototypes (unguarded divide)
    idx, lab, feat, logit = gu.stage1_proto_pass(gu.Trace(flow, "s1"))
    res = F.build_prototypes(cuda(feat), cuda(lab), cuda(logit), act, neg, L, U, guard_empty=False)
    np.testing.assert_allclose(res.proto[0].cpu().numpy(), flow["s1/proto"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(res.t()[0], flow["s1/t"])
    rd0 = gu.stage2_round(gu.Trace(flow, "s2_0"))
    idx0, _, feat0, _ = rd0["extract"]
    tb = F.TagBatch([0, len(idx0)], C, [act], [neg], dataset_idx=torch.from_numpy(idx0), device=DEV)
    true = flow["meta/targets_true"]
    for r, rd in enumerate([rd0, gu.stage2_round(gu.Trace(flow, "s2_1"))]):
        idx_r, _, feat_r, _ = rd["extract"]
        order = {int(d): p for p, d in enumerate(idx_r.tolist())}
        perm = [order[int(d)] for d in idx0.tolist()]
        tb.step(cuda(feat_r[perm]), proto_glob, cf, nf, mode=mode)
        got = tb.traindata_idx(0)
        for j in range(2 * len(neg)):
            assert got[j] == flow[f"s2_{r}/traindata_idx/{j}"].tolist()
        rem = tb.remaining(0)
        for j in range(len(neg)):
            assert rem[j] == flow[f"s2_{r}/idxss/{j}"].tolist()
        y, distill, sup = tb.fill(cuda(true[idx0]))
        pos = {int(d): p for p, d in enumerate(idx0.tolist())}
        y, distill = y.cpu().numpy(), distill.cpu().numpy()
        for st in rd["steps"]:
            rows = [pos[int(d)] for d in st["idx"]]
            np.testing.assert_array_equal(y[rows], st["target"])
            np.testing.assert_array_equal(distill[rows], st["distill"])
        pidx, plab, pfeat, plogit = rd["proto_pass"]
        res = F.build_prototypes(cuda(pfeat), cuda(plab), cuda(plogit), act, neg, L, U, guard_empty=True)
        np.testing.assert_allclose(res.proto[0].cpu().numpy(), flow[f"s2_{r}/proto"], rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(res.t()[0], flow[f"s2_{r}/t"])


# =============================================================================== other aggregators (§8f.3)
def test_other_aggregators_golden(F):
    """model_dist / DaAgg / RSCFed / FedAvg_rela through the kernels vs the reference's own outputs."""
    z = gu.load("aggregators.npz")
    names = [str(n) for n in z["names"]]
    clients = [OrderedDict((n, cuda(z[f"in/{k}/{n}"].copy())) for n in names) for k in range(6)]
    dict_len = z["dict_len"].tolist()
    for a, b in ((0, 1), (2, 5)):
        ref = float(z[f"model_dist/{a}_{b}"])
        assert abs(F.model_dist(clients[a], clients[b]) - ref) <= 1e-5 * ref
        flat_a, flat_b = F.FlatStateDict.from_state_dict(clients[a]), F.FlatStateDict.from_state_dict(clients[b])
        assert abs(F.model_dist(flat_a, flat_b) - ref) <= 1e-5 * ref
    res = F.DaAgg(clients, dict_len, z["daagg/clean"].tolist(), z["daagg/noisy"].tolist())
    for n in names:
        assert res[n].dtype == torch.float32
        np.testing.assert_allclose(res[n].cpu().numpy(), z[f"daagg/out/{n}"], rtol=1e-5, atol=1e-6)
    fclients = [OrderedDict((n, v) for n, v in sd.items() if v.dtype == torch.float32) for sd in clients]
    res = F.RSCFed(z["rscfed/dma"].tolist(), fclients, 3, dict_len, 4)
    for n in res:
        np.testing.assert_allclose(res[n].cpu().numpy(), z[f"rscfed/out/{n}"], rtol=1e-5, atol=1e-6)
    protos = [cuda(p.copy()) for p in z["rela/in"]]
    out = F.FedAvg_rela(protos, dict_len, gu.parse_lists(z["rela/lists"]))
    np.testing.assert_array_equal(out.cpu().numpy(), z["rela/out"])


def test_model_dist_densenet_sized(F):
    from fedmlp_b200.shapes import densenet121_state_shapes, synth_state_dict
    shapes = densenet121_state_shapes(5)
    a = synth_state_dict(shapes, 1)
    b = synth_state_dict(shapes, 2)
    ref = O.model_dist(a, b)
    got = F.model_dist(OrderedDict((k, v.to(DEV)) for k, v in a.items()), OrderedDict((k, v.to(DEV)) for k, v in b.items()))
    assert abs(got - ref) <= 1e-5 * ref
    assert abs(F.model_dist(a, b) - ref) <= 1e-5 * ref          # CPU dicts are staged to the GPU
