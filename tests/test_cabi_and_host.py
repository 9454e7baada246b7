"""CPU-side checks: the C-ABI library loads and exports every symbol include/fedmlp_b200.h
declares (no compute calls without a GPU), argument validation returns the documented status
codes, and the host-side plumbing (flat layouts, DenseNet121 shapes, no-fallback errors)."""
import ctypes
import re
from collections import OrderedDict
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(lib):
    from fedmlp_b200 import _cabi
    header = (ROOT / "include" / "fedmlp_b200.h").read_text()
    declared = set(re.findall(r"\b(fmlp_[a-z0-9_]+)\s*\(", header))
    declared -= {"fmlp_status", "fmlp_stream_t"}
    assert len(declared) >= 20
    raw = ctypes.CDLL(str(_cabi.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_cabi.SIGNATURES) == declared
    assert lib.fmlp_abi_version() == _cabi.ABI_VERSION
    assert b"workspace" in lib.fmlp_status_string(-3)


def test_argument_validation_without_gpu(lib):
    from fedmlp_b200 import _cabi as c
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.fmlp_fedavg_flat_f32(None, None, 1, 10, 1.0, 1, None, None) == -1
    assert lib.fmlp_fedavg_flat_f32(c.ptr_array([16]), c.f32_array([1.0]), 65, 10, 1.0, 1, 16, None) == -1
    assert lib.fmlp_tag_sim_f32(None, 4, 4, None, 5, 1, None, None, None, 0, 0, None, 0, None) == -1
    assert lib.fmlp_loss_stage2_f32(None, None, None, None, 1, 5, 0, None, None, None, 0, None) == -1
    # misaligned / unsupported shapes
    rows = c.i64_array([0, 8])
    m = c.u32_array([1])
    assert lib.fmlp_tag_sim_f32(16, 6, 6, 16, 5, 1, rows, m, 16, 8, 0, 16, 1 << 20, None) == -2        # D % 4 != 0
    assert lib.fmlp_tag_sim_f32(16, 8, 8, 16, 5, 1, rows, m, 16, 8, 0, 16, 8, None) == -3              # workspace too small
    assert lib.fmlp_tag_sim_ws_bytes(5, 1024) >= (2 * 5 * 1024 + 10) * 4
    assert lib.fmlp_tag_select_ws_bytes(8, 5, 100) == 8 * 5 * 2 * 100 * 8
    assert lib.fmlp_loss_ws_bytes(32, 5) >= 4 * 2048 * 4


def test_no_cpu_fallback():
    import fedmlp_b200 as F
    from fedmlp_b200._cabi import FedMLPNativeError
    x = torch.zeros(8, 4)
    with pytest.raises(FedMLPNativeError):
        F.tag_similarity(x, torch.zeros(10, 4), [0])
    with pytest.raises(FedMLPNativeError):
        F.build_prototypes(x, torch.zeros(8, 5), None, [0], [1])
    with pytest.raises(FedMLPNativeError):
        F.fedmlp_stage2_loss(torch.zeros(4, 5), None, torch.zeros(4, 5), torch.zeros(4, 5))
    with pytest.raises(FedMLPNativeError):
        F.pool_tag(torch.zeros(2, 8, 7, 7))                       # fused model tail: CUDA tensors only
    with pytest.raises(FedMLPNativeError):
        from fedmlp_b200 import evaluation
        evaluation.multilabel_metrics(torch.zeros(6, 3), torch.zeros(6, 3))
    if not torch.cuda.is_available():
        with pytest.raises(FedMLPNativeError):
            F.FedAvg([OrderedDict(w=torch.zeros(3))], [1])
        with pytest.raises(Exception):
            F.build_sim_table(torch.zeros(10, 8), [0])


def test_new_entry_points_validate_arguments(lib):
    """pool_tag / sim_table / eval / adam reject bad arguments before touching the device."""
    import ctypes as C
    null = C.c_void_p(0)
    assert lib.fmlp_sim_table_bytes(5, 1024) >= (2 * 5 * 1024 + 10) * 4
    assert lib.fmlp_sim_table_bytes(0, 1024) == 0
    assert lib.fmlp_eval_ws_bytes(100, 5) > 0 and lib.fmlp_eval_ws_bytes(100, 0) == 0
    assert lib.fmlp_pool_tag_f32(null, 0, 1, 8, 49, 1, null, 0, 0, 1, null, 0, null, 0, null) == -1
    assert lib.fmlp_sim_table_build_f32(null, 5, 8, 1, 1, null, null) == -1
    assert lib.fmlp_eval_multilabel_f32(null, null, 10, 3, 0, 0.5, null, null, null, 0, null) == -1


def test_flat_layout_and_densenet_shapes():
    from fedmlp_b200.flat import FlatStateDict, flat_view_of, layout_of
    from fedmlp_b200.shapes import count_params, densenet121_state_shapes, synth_state_dict
    shapes = densenet121_state_shapes(5)
    assert len(shapes) == 727 and count_params(shapes) == (7042629, 121)
    assert count_params(densenet121_state_shapes(14)) == (7051854, 121)
    sd = synth_state_dict(shapes, seed=1)
    lay = layout_of(sd)
    assert lay.n_i64 == 121 and lay.n_f32 >= 7042629 and lay.n_f32 % 4 == 0
    flat = FlatStateDict.from_state_dict(sd)
    assert list(flat.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(flat[k], sd[k]) and flat[k].dtype == sd[k].dtype
    for i, k in enumerate(lay.keys):
        if not lay.is_int[i]:
            assert lay.offsets[i] % 4 == 0                                      # 16-byte aligned tensors
    assert flat_view_of(flat) is not None
    assert flat_view_of(sd) is None                                             # separately allocated tensors
    plain = OrderedDict(flat.items())                                           # same views, plain dict
    assert flat_view_of(plain) == flat_view_of(flat)
    clone = flat.clone()
    clone["classifier.bias"].add_(1)
    assert not torch.equal(clone["classifier.bias"], flat["classifier.bias"])


@pytest.mark.gpu
def test_fedavg_tao_matches_oracle_bitwise():
    import numpy as np
    import fedmlp_b200 as F
    from oracle import fedmlp_oracle as O
    rng = np.random.default_rng(0)
    t = [rng.random(5) for _ in range(4)]
    w = [5000, 4999, 5001, 1234]
    lists = [[1, 2], [0, 3], [], [0, 1, 2, 3], [2]]
    np.testing.assert_array_equal(F.FedAvg_tao(t, w, lists), O.fedavg_tao(t, w, lists))
    np.testing.assert_array_equal(F.FedAvg_tao(t, w), O.fedavg_tao(t, w))


def test_eval_host_arithmetic_matches_reference_globaltest():
    """Host half of the on-device evaluation (fedmlp_b200.evaluation.metrics_from_class_statistics): given the
    per-class sums the kernel produces (computed here with numpy), the result dicts equal the reference's own
    globaltest / classtest outputs (tests/golden/eval.npz)."""
    import golden_util as gu
    from fedmlp_b200.evaluation import metrics_from_class_statistics
    from oracle import fedmlp_oracle as O
    z = gu.load("eval.npz")
    probs, y = z["probs"], z["labels"] != 0
    pred = probs > 0.5
    N, C = probs.shape
    cnt = np.stack([y.sum(0), (~y).sum(0), pred.sum(0), (y & pred).sum(0), (~y & ~pred).sum(0), (y != pred).sum(0)], axis=1)
    aa = np.array([[O.average_precision(y[:, c], probs[:, c]), O.roc_auc(y[:, c], probs[:, c])] for c in range(C)])
    res = metrics_from_class_statistics(cnt, aa, N)
    for k in ("mAP", "BACC", "R", "F1", "auc", "P", "hamming_loss"):
        assert abs(float(res[k]) - float(z[f"globaltest/{k}"])) <= 1e-12, k
    assert res["mAP"].dtype == torch.float32
    for i in range(C):
        one = metrics_from_class_statistics(cnt, aa, N, classid=i)
        for k in ("BACC", "R", "F1", "P"):
            assert abs(float(one[k]) - float(z[f"classtest/{i}/{k}"])) <= 1e-12, (i, k)
    # a class nobody predicts: Precision skips it but still divides by C; F1 stays finite
    cnt2 = cnt.copy(); cnt2[0, 2] = 0; cnt2[0, 3] = 0
    r2 = metrics_from_class_statistics(cnt2, aa, N)
    assert np.isfinite(r2["P"]) and np.isfinite(r2["F1"])


def test_pool_layout_detection():
    """pooling._layout_of: NCHW vs channels_last vs strided inputs (pure host logic)."""
    from fedmlp_b200 import pooling
    x = torch.zeros(2, 8, 7, 7)
    assert pooling._layout_of(x)[1:] == (pooling.FMAP_NCHW, 49)
    xl = x.contiguous(memory_format=torch.channels_last)
    t, layout, hw = pooling._layout_of(xl)
    assert layout == pooling.FMAP_NHWC and hw == 49 and t.data_ptr() == xl.data_ptr()
    xs = torch.zeros(2, 16, 7, 7)[:, ::2]                 # neither layout: densified to NCHW
    t, layout, hw = pooling._layout_of(xs)
    assert layout == pooling.FMAP_NCHW and t.is_contiguous()
    assert pooling._layout_of(torch.zeros(3, 8, 49))[1:] == (pooling.FMAP_NCHW, 49)
    with pytest.raises(ValueError):
        pooling._layout_of(torch.zeros(3, 8))


def test_tuning_knobs_and_new_entry_points_without_gpu(lib):
    """fmlp_set_tuning / fmlp_get_tuning (scheduling knobs, include/fedmlp_b200.h) and the argument checks of the
    round-2 aggregation-tails entry point: host-only behaviour, no CUDA call."""
    from fedmlp_b200 import _cabi as c
    for knob in (c.TUNE_PROTO_PAD_SMEM_KB, c.TUNE_SIM_REQUEST_SMEM_KB, c.TUNE_SIM_SMEM_BUDGET_KB, c.TUNE_SELECT_CLUSTER):
        prev = lib.fmlp_get_tuning(knob)
        assert lib.fmlp_set_tuning(knob, 8) == 0 and lib.fmlp_get_tuning(knob) == 8
        assert lib.fmlp_set_tuning(knob, -1) == 0 and lib.fmlp_get_tuning(knob) == -1     # back to "unset"
        assert lib.fmlp_set_tuning(knob, -2) == -1
        lib.fmlp_set_tuning(knob, prev)
    assert lib.fmlp_set_tuning(99, 1) == -1 and lib.fmlp_get_tuning(99) == -1
    w = c.f64_array([1.0, 2.0])
    masks = c.u64_array([1, 2, 3])
    # null prototypes / zero total weight / counters without an output buffer
    assert lib.fmlp_agg_tails_local_f32(None, 2, 3, 8, w, masks, 16, None, None, None, None, 0, 3.0, None, None, None) == -1
    assert lib.fmlp_agg_tails_local_f32(16, 2, 3, 8, w, masks, 16, None, None, None, None, 0, 0.0, None, None, None) == -1
    assert lib.fmlp_agg_tails_local_f32(16, 2, 3, 8, w, masks, 16, None, None, None, c.ptr_array([16, 16]), 4, 3.0, None, None, None) == -1
    assert lib.fmlp_agg_tails_local_f32(16, 2, 3, 8, w, masks, 16, 16, None, None, None, 0, 3.0, 16, None, None) == -1   # tao needs rows / masks


def test_fast_layout_match_is_strict_about_sizes_and_dtypes():
    """flat.layout_of(sd, like=...) (the per-client layout check of FedAvg on scattered state_dicts) accepts only dicts
    whose keys, element counts and dtypes equal client 0's."""
    from collections import OrderedDict

    import torch

    from fedmlp_b200.flat import layout_of
    a = OrderedDict(w=torch.zeros(3, 4), b=torch.zeros(4), n=torch.zeros((), dtype=torch.int64))
    lay = layout_of(a)
    same = OrderedDict(w=torch.ones(3, 4), b=torch.ones(4), n=torch.ones((), dtype=torch.int64))
    assert layout_of(same, like=lay) is lay
    assert layout_of(OrderedDict(w=torch.ones(3, 5), b=torch.ones(4), n=torch.ones((), dtype=torch.int64)), like=lay) is not lay
    assert layout_of(OrderedDict(w=torch.ones(3, 4), c=torch.ones(4), n=torch.ones((), dtype=torch.int64)), like=lay) is not lay
    assert layout_of(OrderedDict(w=torch.ones(3, 4), b=torch.ones(4, dtype=torch.int64), n=torch.ones((), dtype=torch.int64)), like=lay) is not lay


def test_host_copy_many_packs_exactly(lib):
    """fmlp_host_copy_many (host threads): byte-exact copies incl. tensors split between two threads, empty entries."""
    import numpy as np
    import torch
    g = torch.Generator().manual_seed(5)
    sizes = [0, 1, 7, 300_000, 5, 900_001, 64, 0, 123_457]
    srcs = [torch.randn(n, generator=g) for n in sizes]
    dst = torch.full((sum(sizes) + 16,), -7.0)
    offs = np.cumsum([0] + sizes[:-1]).astype(np.int64)
    sp = np.array([s.data_ptr() if s.numel() else 0 for s in srcs], dtype=np.int64)
    dp = np.ascontiguousarray(dst.data_ptr() + 4 * offs)
    nb = np.array([4 * n for n in sizes], dtype=np.int64)
    for threads in (1, 3, 8):
        dst.fill_(-7.0)
        assert lib.fmlp_host_copy_many(sp.ctypes.data, dp.ctypes.data, nb.ctypes.data, len(sizes), threads) == 0
        assert torch.equal(dst[:sum(sizes)], torch.cat(srcs))
        assert bool((dst[sum(sizes):] == -7.0).all())
    assert lib.fmlp_host_copy_many(None, None, None, 2, 1) == -1


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver launches beside the GPU arm): one JSON line on stdout with the
    GPU arm's metric / unit / config keys, impl = reference, an e2e block of its own value and zero copy bytes."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--rows-per-client", "200",
                          "--clients-per-gpu", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=300, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "client_samples_per_sec_per_fedmlp_round" and d["unit"] == "client-samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    import bench
    a = bench.parse_args.__globals__["argparse"].Namespace(sim_mode="folded", clients_per_gpu=2, rows_per_client=200)
    # the GPU arm builds its `config` with the same call: equal objects for the same workload, N and flags
    assert d["config"] == bench.Workload("ich55k", a).full_config(a, 1)
    assert "l2" in d["config"] and "workload" in d["config"] and "collective" not in d["config"]
