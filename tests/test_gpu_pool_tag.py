"""GPU parity of the fused pooling + tagging kernel (csrc/pool_tag.cu) against the oracle,
the golden vectors of the reference (tests/golden/pool.npz, tagging.npz) and the unfused K3 kernel."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import fedmlp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def P(lib):
    from fedmlp_b200 import pooling
    return pooling


def _case(B, D, H, W, C, seed, signed=True):
    g = torch.Generator().manual_seed(seed)
    fmap = torch.randn(B, D, H, W, generator=g)
    if not signed:
        fmap = fmap.abs()
    proto = torch.relu(torch.randn(2 * C, D, generator=g)) + 0.05
    return fmap, proto


@pytest.mark.parametrize("mode", ["folded", "pair"])
@pytest.mark.parametrize("relu", [True, False])
def test_pool_tag_golden(P, mode, relu):
    z = gu.load("pool.npz")
    tag = "" if relu else "_norelu"
    fmap, proto = torch.from_numpy(z["fmap"]).to(DEV), torch.from_numpy(z["proto"]).to(DEV)
    B, C = fmap.shape[0], proto.shape[0] // 2
    for layout in ("nchw", "nhwc"):
        x = fmap if layout == "nchw" else fmap.contiguous(memory_format=torch.channels_last)
        table = P.build_sim_table(proto, [0, 1, 2, 3, 4], mode)
        sim = torch.full((C, B + 3), float("nan"), device=DEV)
        feat = P.pool_tag(x, table, sim_out=sim, col0=2, relu=relu)
        np.testing.assert_allclose(feat.cpu().numpy(), z["feat" + tag], rtol=1e-5, atol=1e-7)
        got = sim.cpu().numpy()
        assert np.isnan(got[:, :2]).all() and np.isnan(got[:, B + 2:]).all()      # untouched columns
        for c in range(C):
            np.testing.assert_allclose(got[c, 2:B + 2], z[f"sim{tag}/{c}"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("B,D,H,W,C,missing,layout,mode", [
    (128, 1024, 7, 7, 5, [1, 2, 3, 4], "nchw", "folded"),      # DenseNet121 / ICH batch
    (128, 1024, 7, 7, 5, [0, 1, 2, 4], "nhwc", "pair"),
    (37, 1280, 7, 7, 14, [c for c in range(14) if c != 5], "nchw", "folded"),   # EfficientNet-B0 / CXR14
    (37, 1280, 7, 7, 14, [c for c in range(14) if c != 5], "nhwc", "pair"),
    (1, 1024, 7, 7, 5, [0], "nchw", "pair"),
    (700, 1024, 7, 7, 5, [1, 2, 3, 4], "nchw", "folded"),      # more samples than resident clusters
    (300, 2208, 7, 7, 14, [0, 3, 13], "nhwc", "folded"),       # DenseNet161 width: ragged channel ranges
    (5, 1000, 7, 7, 3, [1], "nchw", "folded"),                 # D not a multiple of the stage width
    (9, 64, 8, 8, 4, [0, 1, 3], "nchw", "pair"),               # even HW (rotated smem reads)
    (9, 64, 8, 8, 4, [0, 1, 3], "nhwc", "folded"),
    (6, 256, 14, 14, 5, [2, 4], "nchw", "folded"),             # HW = 196: narrow stages
    (6, 256, 14, 14, 5, [2, 4], "nhwc", "pair"),
    (11, 40, 1, 1, 5, [0, 1, 2, 3], "nchw", "pair"),           # HW = 1: pooling is the identity
    (4, 512, 16, 16, 32, list(range(31)), "nchw", "folded"),   # 31 class vectors, HW = 256
])
def test_pool_tag_vs_oracle(P, B, D, H, W, C, missing, layout, mode):
    fmap, proto = _case(B, D, H, W, C, seed=B * 131 + D)
    ref_feat, ref_sim = O.pool_tag(fmap, proto, missing, relu=True)
    x = fmap.to(DEV)
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    table = P.build_sim_table(proto.to(DEV), missing, mode)
    sim = torch.full((C, B), float("nan"), device=DEV)
    feat = P.pool_tag(x, table, sim_out=sim, relu=True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(feat.cpu().numpy(), ref_feat.numpy(), rtol=1e-5, atol=1e-7)
    got = sim.cpu().numpy()
    for c in range(C):
        if c in missing:
            np.testing.assert_allclose(got[c], ref_sim[c].numpy(), rtol=0, atol=2e-6)
        else:
            assert np.isnan(got[c]).all()


def test_pool_only_and_empty(P):
    fmap, _ = _case(17, 128, 7, 7, 2, seed=5)
    feat = P.pool_tag(fmap.to(DEV), relu=False)
    np.testing.assert_allclose(feat.cpu().numpy(), O.pooled_features(fmap, relu=False).numpy(), rtol=1e-5, atol=1e-7)
    out = P.pool_tag(torch.zeros(0, 128, 7, 7, device=DEV))
    assert tuple(out.shape) == (0, 128)
    with pytest.raises(Exception):
        P.pool_tag(torch.zeros(2, 6, 7, 7, device=DEV))           # D % 4 != 0 -> loud error, no fallback


def test_pool_tag_nan_propagates_like_torch(P):
    fmap, proto = _case(3, 64, 7, 7, 2, seed=9)
    fmap[1, 5, 3, 3] = float("nan")
    feat = P.pool_tag(fmap.to(DEV)).cpu()
    ref = O.pooled_features(fmap)
    assert torch.isnan(feat[1, 5]) and torch.isnan(ref[1, 5])
    assert not torch.isnan(feat[0]).any() and not torch.isnan(feat[2]).any()


def test_streamed_tagging_equals_batched(P, lib):
    """Batch-by-batch fused pooling + scoring into TagBatch.sim, then selection: the same picks as the
    unfused flow on the concatenated pooled features (reference :1026-1112)."""
    from fedmlp_b200 import TagBatch
    N, D, C, missing = 1000, 1024, 5, [0, 1, 3, 4]
    g = torch.Generator().manual_seed(31)
    fmap = torch.randn(N, D, 7, 7, generator=g)
    proto = torch.relu(torch.randn(2 * C, D, generator=g)) + 0.05
    feat_ref = O.pooled_features(fmap)
    a = TagBatch([0, N], C, [[2]], [missing])
    b = TagBatch([0, N], C, [[2]], [missing])
    table = P.build_sim_table(proto.to(DEV), missing, "folded")
    feats = []
    for r0 in range(0, N, 128):
        feats.append(P.pool_tag(fmap[r0:r0 + 128].to(DEV), table, sim_out=a.sim, col0=r0))
    ca, sa, _ = a.select(0.05, 0.1)
    feat = torch.cat(feats)
    np.testing.assert_allclose(feat.cpu().numpy(), feat_ref.numpy(), rtol=1e-5, atol=1e-7)
    cb, sb, _ = b.step(feat, proto.to(DEV), 0.05, 0.1, mode="folded")
    assert torch.equal(ca, cb)
    # identical picks unless two sims differ by < 1e-6 around a cut (north_star waiver): compare as sets per side
    sa, sb = sa.cpu().numpy(), sb.cpu().numpy()
    sim_a, sim_b = a.sim.cpu().numpy(), b.sim.cpu().numpy()
    assert np.nanmax(np.abs(sim_a - sim_b)) < 1e-6
    for c in missing:
        for side in range(2):
            n = int(ca[0, c, 2 + side])
            da = set(sa[0, c, side, :n].tolist()) ^ set(sb[0, c, side, :n].tolist())
            for row in da:       # any disagreement must sit within the waiver of the cut value
                cut = np.sort(sim_b[c][sb[0, c, side, :n]])[0 if side == 0 else -1]
                assert abs(sim_b[c][row] - cut) < 1e-6
    # oracle selection on the oracle's similarities agrees as well
    ref_sim = O.tag_similarity(feat_ref, proto, missing)
    for c in missing:
        sel = O.split_and_select(ref_sim[c].numpy(), 0.05, 0.1)
        ref_vals = ref_sim[c].numpy()
        for side, key, n in ((0, "clean", sel["m"]), (1, "noise", sel["k"])):
            if abs(int(ca[0, c, side]) - (sel["n_clean"], sel["n_noise"])[side]) == 0:
                assert int(ca[0, c, 2 + side]) == n
                for row in set(sa[0, c, side, :n].tolist()) ^ set(sel[key]):
                    cut = ref_vals[sel[key][-1]]
                    assert abs(ref_vals[row] - cut) < 2e-6


def test_fused_tail_module(P):
    """FusedTail keeps the reference's `net(x) -> (feature, logits)` contract: kernel tail under no_grad,
    torch tail with autograd, identical values; tagging() fills the similarity columns."""
    torch.manual_seed(3)
    feats = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 3, stride=2, padding=1), torch.nn.BatchNorm2d(64)).to(DEV)
    net = P.FusedTail(feats, torch.nn.Linear(64, 5).to(DEV)).eval()
    x = torch.randn(10, 3, 14, 14, device=DEV)
    with torch.no_grad():
        f0, z0 = net(x)
    f2, z2 = net(x)          # grad enabled: torch tail with autograd
    assert f2.requires_grad
    np.testing.assert_allclose(f0.cpu().numpy(), f2.detach().cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(z0.cpu().numpy(), z2.detach().cpu().numpy(), rtol=1e-4, atol=1e-6)
    proto = torch.rand(10, 64, device=DEV) + 0.05
    table = P.build_sim_table(proto, [0, 2, 3], "pair")
    sim = torch.full((5, 16), float("nan"), device=DEV)
    f3, _ = net.tagging(x, table, sim, 4)
    ref = O.tag_similarity(f0.cpu(), proto.cpu(), [0, 2, 3])
    for c in (0, 2, 3):
        np.testing.assert_allclose(sim[c, 4:14].cpu().numpy(), ref[c].numpy(), rtol=0, atol=2e-6)
    assert torch.isnan(sim[1]).all() and torch.isnan(sim[:, :4]).all()
