"""Pin the oracle (oracle/fedmlp_oracle.py) to outputs of the reference itself (tests/golden).

The fixtures were produced by oracle/make_golden.py from /root/reference: direct calls of
utils/FedAvg.py, CosineSimilarityFast, max_m_indices/min_n_indices, DatasetSplit_pseudo and a
recorded run of the unmodified LocalUpdate.train_FedMLP.  Exact comparisons are used for
integer / index / element-wise results; sums that go through BLAS / vectorised reductions are
compared to 1e-6 because their summation order depends on the CPU and thread count.
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import fedmlp_oracle as O


# ------------------------------------------------------------------------------------ FedAvg
def _fedavg_inputs(z):
    names = [str(n) for n in z["names"]]
    K = 4
    clients = [OrderedDict((n, torch.from_numpy(z[f"in/{k}/{n}"].copy())) for n in names) for k in range(K)]
    return names, clients


@pytest.mark.parametrize("tag", ["int", "float", "float_odd"])
def test_fedavg_matches_reference(tag):
    z = gu.load("fedavg.npz")
    names, clients = _fedavg_inputs(z)
    w = z[f"weights/{tag}"].tolist()
    if tag == "int":
        w = [int(v) for v in w]
    out = O.fedavg(clients, w)
    assert list(out.keys()) == names
    for n in names:
        ref = z[f"out/{tag}/{n}"]
        assert str(out[n].dtype) == str(z[f"outdtype/{tag}/{n}"])
        assert out[n].dtype == torch.float32          # int64 counters come out float32
        np.testing.assert_array_equal(out[n].numpy(), ref)


def test_fedavg_proto_and_tao_match_reference():
    z = gu.load("fedavg.npz")
    protos = [torch.from_numpy(p.copy()) for p in z["proto/in"]]
    weight = z["proto/weight"].tolist()
    lists = gu.parse_lists(z["proto/lists"])
    out = O.fedavg_proto(protos, weight, lists)
    np.testing.assert_array_equal(out.numpy(), z["proto/out"])      # NaN rows compare equal here
    assert np.isnan(z["proto/out"][6]).all() and np.isnan(z["proto/out"][7]).all()
    taos = [t.copy() for t in z["tao/in"]]
    neg = gu.parse_lists(z["tao/lists"])
    np.testing.assert_array_equal(O.fedavg_tao(taos, weight, neg), z["tao/out_lists"])
    np.testing.assert_array_equal(O.fedavg_tao(taos, weight), z["tao/out_plain"])


# ------------------------------------------------------------------------------------ tagging
def test_cosine_and_selection_match_reference():
    z = gu.load("tagging.npz")
    feat, proto = torch.from_numpy(z["feat"].copy()), torch.from_numpy(z["proto"].copy())
    C = proto.shape[0] // 2
    sims = O.tag_similarity(feat, proto, list(range(C)))
    for c in range(C):
        np.testing.assert_allclose(O.cosine_similarity_fast(feat, proto[2 * c:2 * c + 1]).numpy(), z[f"cos0/{c}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(sims[c].numpy(), z[f"sim/{c}"], rtol=0, atol=1e-6)
        ref_sim = z[f"sim/{c}"]
        for n in (0, 1, 3, 10, 40):
            for fn in (O.top_positions_python, O.top_positions):
                assert fn(ref_sim.tolist() if fn is O.top_positions_python else ref_sim, n, True) == z[f"max/{c}/{n}"].tolist()
                assert fn(ref_sim.tolist() if fn is O.top_positions_python else ref_sim, n, False) == z[f"min/{c}/{n}"].tolist()
    vals = z["ties/vals"]
    for n in range(len(vals) + 1):
        assert O.top_positions_python(vals.tolist(), n, True) == z[f"ties/max/{n}"].tolist()
        assert O.top_positions(vals, n, True) == z[f"ties/max/{n}"].tolist()
        assert O.top_positions_python(vals.tolist(), n, False) == z[f"ties/min/{n}"].tolist()
        assert O.top_positions(vals, n, False) == z[f"ties/min/{n}"].tolist()


def test_mask_fill_matches_reference():
    z = gu.load("maskfill.npz")
    idxs = z["idxs"].tolist()
    tdi = [[float(v) for v in s.split(",") if v] for s in z["traindata_idx"].tolist()]
    labels = z["targets_true"][idxs]
    tgt, dis = O.mask_fill(labels, idxs, z["active"].tolist(), z["negative"].tolist(), tdi)
    np.testing.assert_array_equal(tgt, z["out_target"])
    np.testing.assert_array_equal(dis, z["out_distill"])
    assert z["out_idx"].tolist() == idxs


# ------------------------------------------------------------------------------------ recorded flow
@pytest.fixture(scope="module")
def flow():
    return gu.load("flow.npz")


def test_flow_stage1_losses_and_grads(flow):
    tr = gu.Trace(flow, "s1")
    steps = gu.stage1_steps(tr)
    assert len(steps) == 7 and steps[-1]["z1"].shape[0] == 8          # 104 = 6*16 + 8: partial last batch
    act, neg, bs = flow["s1/act_list"].tolist(), flow["s1/neg_list"].tolist(), int(flow["meta/batch_size"])
    losses = []
    for st in steps:
        t = {k: torch.from_numpy(np.array(v)) for k, v in st.items() if k != "loss"}
        np.testing.assert_array_equal(O.bce_on_probs(t["p1"], t["y"]).numpy(), st["bce1"])
        loss, dz1, dz2 = O.loss_and_grads(lambda a, b, c, d, y: O.stage1_loss(a, b, c, d, y, act, neg, bs),
                                          t["z1"], t["z2"], t["z3"], t["z4"], t["y"], n_grad=2)
        assert abs(float(loss) - st["loss"]) <= 1e-6 * abs(st["loss"])
        np.testing.assert_allclose(dz1.numpy(), st["dz1"], rtol=1e-5, atol=1e-10)
        np.testing.assert_allclose(dz2.numpy(), st["dz2"], rtol=1e-5, atol=1e-10)
        losses.append(st["loss"])
    assert abs(np.mean(losses) - float(flow["s1/loss_mean"])) < 1e-9


def test_flow_stage1_prototypes(flow):
    tr = gu.Trace(flow, "s1")
    idx, lab, feat, logit = gu.stage1_proto_pass(tr)
    assert idx.tolist() == flow["meta/idxs"].tolist()                 # shuffle=False pass (:975)
    act, neg = flow["s1/act_list"].tolist(), flow["s1/neg_list"].tolist()
    proto, num, t = O.prototype_build(torch.from_numpy(feat), torch.from_numpy(lab), torch.from_numpy(logit),
                                      act, neg, float(flow["meta/L"]), float(flow["meta/U"]), guard_empty=False,
                                      batch=4 * int(flow["meta/batch_size"]))
    np.testing.assert_allclose(proto.numpy(), flow["s1/proto"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(t, flow["s1/t"])


@pytest.mark.parametrize("r", [0, 1])
def test_flow_stage2_round(flow, r):
    rd = gu.stage2_round(gu.Trace(flow, f"s2_{r}"))
    idx, lab, feat, _ = rd["extract"]
    neg, act = flow["s1/neg_list"].tolist(), flow["s1/act_list"].tolist()
    proto = torch.from_numpy(flow["proto_glob"].copy())
    cf, nf = float(flow["meta/clean_threshold"]), float(flow["meta/noise_threshold"])
    # similarity: every recorded CosineSimilarityFast call
    for call in rd["cos"]:
        out = O.cosine_similarity_fast(torch.from_numpy(call["x1"].copy()), torch.from_numpy(call["x2"].copy()))
        np.testing.assert_allclose(out.numpy(), call["out"], rtol=0, atol=1e-6)
    # selection on the recorded similarity lists (exact)
    for s in rd["sel"]:
        got = O.top_positions_python(s["values"].tolist(), s["n"], s["kind"] == "max")
        assert got == s["out"].tolist()
        assert O.top_positions(s["values"], s["n"], s["kind"] == "max") == s["out"].tolist()
    # counts int(frac*len(side)) as recorded in the n argument of the calls
    for i in range(len(neg)):
        v = rd["sel"][2 * i]["values"]
        assert rd["sel"][2 * i]["n"] == int(cf * int(np.sum(v >= 0)))
        assert rd["sel"][2 * i + 1]["n"] == int(nf * int(np.sum(v < 0)))
    # label / mask fill as seen by the training loop
    tdi = [flow[f"s2_{r}/traindata_idx/{j}"].tolist() for j in range(2 * len(neg))]
    true = flow["meta/targets_true"]
    for st in rd["steps"]:
        tgt, dis = O.mask_fill(true[st["idx"]], st["idx"].tolist(), act, neg, tdi)
        np.testing.assert_array_equal(tgt, st["target"])
        np.testing.assert_array_equal(dis, st["distill"])
        t = {k: torch.from_numpy(np.array(v)) for k, v in st.items() if k not in ("loss", "idx")}
        loss, dz = O.loss_and_grads(lambda z, zg, y, d: O.stage2_loss(z, zg, y, d, "sup"),
                                    t["z"], t["zg"], t["y"], t["distill"], n_grad=1)
        assert abs(float(loss) - st["loss"]) <= 1e-6 * abs(st["loss"])
        np.testing.assert_allclose(dz.numpy(), st["dz"], rtol=1e-5, atol=1e-10)
    assert abs(np.mean([s["loss"] for s in rd["steps"]]) - float(flow[f"s2_{r}/loss_mean"])) < 1e-9
    # remaining candidates
    rem = O.remaining_indices(flow["meta/idxs"].tolist(), tdi)
    for j in range(len(neg)):
        assert rem[j] == flow[f"s2_{r}/idxss/{j}"].tolist()
    # prototypes / t of the round (guarded divide)
    pidx, plab, pfeat, plogit = rd["proto_pass"]
    p, num, t = O.prototype_build(torch.from_numpy(pfeat), torch.from_numpy(plab), torch.from_numpy(plogit), act, neg,
                                  float(flow["meta/L"]), float(flow["meta/U"]), guard_empty=True,
                                  batch=4 * int(flow["meta/batch_size"]))
    np.testing.assert_allclose(p.numpy(), flow[f"s2_{r}/proto"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(t, flow[f"s2_{r}/t"])


def test_flow_tagging_state_across_rounds(flow):
    """TaggingState (candidate bookkeeping across rounds) reproduces self.traindata_idx of the
    reference after round 1 and after round 2 (append, then extend)."""
    neg = flow["s1/neg_list"].tolist()
    proto = torch.from_numpy(flow["proto_glob"].copy())
    cf, nf = float(flow["meta/clean_threshold"]), float(flow["meta/noise_threshold"])
    rd0 = gu.stage2_round(gu.Trace(flow, "s2_0"))
    idx0, _, feat0, _ = rd0["extract"]
    state = O.TaggingState(idx0.tolist(), neg)
    state.step(torch.from_numpy(feat0), proto, cf, nf, python_sort=True)
    for j in range(2 * len(neg)):
        assert state.traindata_idx[j] == [int(v) for v in flow[f"s2_0/traindata_idx/{j}"]]
    rd1 = gu.stage2_round(gu.Trace(flow, "s2_1"))
    idx1, _, feat1, _ = rd1["extract"]
    order = {int(d): p for p, d in enumerate(idx1.tolist())}
    perm = [order[d] for d in idx0.tolist()]                     # round-2 features in round-1 row order
    state.step(torch.from_numpy(feat1[perm]), proto, cf, nf, python_sort=True)
    for j in range(2 * len(neg)):
        assert state.traindata_idx[j] == [int(v) for v in flow[f"s2_1/traindata_idx/{j}"]]


# ------------------------------------------------------------------------------------ other aggregators (§8f.3)
def _agg_inputs(z):
    names = [str(n) for n in z["names"]]
    return names, [OrderedDict((n, torch.from_numpy(z[f"in/{k}/{n}"].copy())) for n in names) for k in range(6)]


def test_other_aggregators_match_reference():
    z = gu.load("aggregators.npz")
    names, clients = _agg_inputs(z)
    dict_len = z["dict_len"].tolist()
    assert abs(O.model_dist(clients[0], clients[1]) - float(z["model_dist/0_1"])) <= 1e-6 * float(z["model_dist/0_1"])
    assert abs(O.model_dist(clients[2], clients[5]) - float(z["model_dist/2_5"])) <= 1e-6 * float(z["model_dist/2_5"])
    res = O.daagg(clients, dict_len, z["daagg/clean"].tolist(), z["daagg/noisy"].tolist())
    for n in names:
        assert str(res[n].dtype) == str(z[f"daagg/dtype/{n}"])
        np.testing.assert_allclose(res[n].numpy(), z[f"daagg/out/{n}"], rtol=1e-6, atol=1e-7)
    fclients = [OrderedDict((n, v) for n, v in sd.items() if v.dtype == torch.float32) for sd in clients]
    res = O.rscfed(z["rscfed/dma"].tolist(), fclients, 3, dict_len, 4)
    for n in res:
        np.testing.assert_allclose(res[n].numpy(), z[f"rscfed/out/{n}"], rtol=1e-6, atol=1e-7)
    protos = [torch.from_numpy(p.copy()) for p in z["rela/in"]]
    with np.errstate(all="ignore"):
        out = O.fedavg_rela(protos, dict_len, gu.parse_lists(z["rela/lists"]))
    np.testing.assert_array_equal(out.numpy(), z["rela/out"])


def test_pooled_features_and_sims_match_reference():
    """the oracle's model-tail restatement against torchvision's DenseNet.forward tail and
    the reference's CosineSimilarityFast (tests/golden/pool.npz, oracle/make_golden.py:make_pool)."""
    z = gu.load("pool.npz")
    fmap, proto = torch.from_numpy(z["fmap"]), torch.from_numpy(z["proto"])
    for tag, relu in (("", True), ("_norelu", False)):
        feat, sims = O.pool_tag(fmap, proto, list(range(5)), relu=relu)
        np.testing.assert_allclose(feat.numpy(), z["feat" + tag], rtol=1e-6, atol=1e-8)
        for c in range(5):
            np.testing.assert_allclose(sims[c].numpy(), z[f"sim{tag}/{c}"], rtol=0, atol=1e-6)
    assert O.pooled_features(fmap)[3, 7] == 0.0


def test_eval_metrics_match_reference_globaltest():
    """the oracle's restatement of sklearn AP / ROC-AUC and utils/multilabel_metrixs.py against
    the reference's own globaltest() output (tests/golden/eval.npz, oracle/make_golden.py:make_eval)."""
    z = gu.load("eval.npz")
    r = O.eval_metrics(z["probs"], z["labels"])
    for k in ("mAP", "BACC", "R", "F1", "auc", "P", "hamming_loss"):
        assert abs(float(r[k]) - float(z[f"globaltest/{k}"])) <= 1e-12, k
    assert abs(O.average_precision(z["ties/y"], z["ties/p"]) - float(z["ties/ap"])) <= 1e-15
    assert abs(O.roc_auc(z["ties/y"], z["ties/p"]) - float(z["ties/auc"])) <= 1e-15
