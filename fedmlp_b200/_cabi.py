"""ctypes binding of libfedmlp_b200.so (the C ABI declared in include/fedmlp_b200.h).

There is deliberately NO fallback: if the shared library is missing or a CUDA device is not
available, the product path raises.  The CPU restatement of the reference lives under oracle/
and is test infrastructure only.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os as _os

# FEDMLP_B200_LIB: alternative build of the same library (tuning experiments under tools/); the default is in-tree
LIB_PATH = Path(_os.environ.get("FEDMLP_B200_LIB") or (Path(__file__).resolve().parent / "lib" / "libfedmlp_b200.so"))

ABI_VERSION = 4
TUNE_PROTO_PAD_SMEM_KB, TUNE_SIM_REQUEST_SMEM_KB, TUNE_SIM_SMEM_BUDGET_KB, TUNE_SELECT_CLUSTER = 0, 1, 2, 3
MAX_CLASSES = 32
MAX_SEGMENTS = 64
MAX_CLIENTS = 64
FEDAVG_CHUNK = 2048

FEDAVG_DIVIDE = 1
FEDAVG_ACCUMULATE = 2
SIM_PAIR = 0
SIM_FOLDED = 1
LOSS2_SUP = 0
LOSS2_SUP_DIS = 1

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u32 = C.c_uint32
_f = C.c_float
_d = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/fedmlp_b200.h one to one
SIGNATURES = {
    "fmlp_abi_version": (_i, []),
    "fmlp_host_copy_many": (_i, [_p, _p, _p, C.c_int64, _i]),
    "fmlp_set_tuning": (_i, [_i, _i]),
    "fmlp_get_tuning": (_i, [_i]),
    "fmlp_status_string": (C.c_char_p, [_i]),
    "fmlp_sm_count": (_i, []),
    "fmlp_launch_count": (C.c_ulonglong, []),
    "fmlp_fedavg_flat_f32": (_i, [_p, _p, _i, _i64, _f, _i, _p, _p]),
    "fmlp_fedavg_flat_i64": (_i, [_p, _p, _i, _i64, _d, _i, _i, _p, _p]),
    "fmlp_fedavg_multi_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i, _p, _i, _f, _i, _p]),
    "fmlp_fedavg_multi_i64": (_i, [_p, _p, _p, _p, _i64, _i, _p, _i, _d, _i, _i, _p]),
    "fmlp_fedavg_allreduce_f32": (_i, [_p, _p, _i, _i64, _p, _p, _p, _i64, _i, _i, _i, _p, _p]),
    "fmlp_fedavg_allreduce_q_buffer_floats": (_sz, [_i64, _i64, _i]),
    "fmlp_fedavg_allreduce_q_f32": (_i, [_p, _p, _p, _i, _i64, _i64, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _i, _i, _p]),
    "fmlp_agg_tail_pack_f64": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p]),
    "fmlp_agg_finalize_f32": (_i, [_p, _p, _i, _i, _i, _d, _p, _p, _p, _p]),
    "fmlp_proto_avg_f32": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "fmlp_agg_tails_local_f32": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _d, _p, _p, _p]),
    "fmlp_tao_avg_f64": (_i, [_p, _i, _i, _p, _p, _d, _p, _p]),
    "fmlp_model_dist_ws_bytes": (_sz, [_i64, _i]),
    "fmlp_model_dist_f32": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _p, _p, _sz, _p]),
    "fmlp_proto_ws_bytes": (_sz, [_i64, _i, _i, _i]),
    "fmlp_proto_build_f32": (_i, [_p, _i64, _i, _p, _p, _i, _i, _i, _p, _p, _p, _f, _f, _i, _p, _p, _p, _p, _sz, _p]),
    "fmlp_tag_sim_ws_bytes": (_sz, [_i, _i]),
    "fmlp_tag_sim_f32": (_i, [_p, _i64, _i, _p, _i, _i, _p, _p, _p, _i64, _i, _p, _sz, _p]),
    "fmlp_sim_table_bytes": (_sz, [_i, _i]),
    "fmlp_sim_table_build_f32": (_i, [_p, _i, _i, _u32, _i, _p, _p]),
    "fmlp_pool_tag_f32": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _u32, _i, _p, _i64, _p, _i64, _p]),
    "fmlp_tag_select_ws_bytes": (_sz, [_i, _i, _i64]),
    "fmlp_tag_select": (_i, [_p, _i64, _p, _i64, _i, _i, _p, _p, _d, _d, _p, _p, _p, _i64, _p, _sz, _p]),
    "fmlp_mask_fill": (_i, [_p, _p, _i64, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "fmlp_eval_ws_bytes": (_sz, [_i64, _i]),
    "fmlp_eval_multilabel_f32": (_i, [_p, _p, _i64, _i, _i, _f, _p, _p, _p, _sz, _p]),
    "fmlp_loss_ws_bytes": (_sz, [_i64, _i]),
    "fmlp_loss_stage1_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i, _u32, _u32, _i, _p, _p, _p, _p, _sz, _p]),
    "fmlp_loss_stage2_f32": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _sz, _p]),
    "fmlp_loss_stage2_seg_f32": (_i, [_p, _p, _p, _p, _i, _i, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "fmlp_fill_loss_stage2_f32": (_i, [_p, _p, _i64, _p, _p, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "fmlp_adam_step_f32": (_i, [_p, _p, _p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i64, _i, _p]),
    "fmlp_scale_f32": (_i, [_p, _i64, _p, _p]),
}


class FedMLPNativeError(RuntimeError):
    pass


_lib = None


def load(path: Path | None = None):
    """dlopen the library and bind every symbol of the header; raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise FedMLPNativeError(
            f"{p} not found: build it with `python -m fedmlp_b200._build` "
            "(fedmlp_b200 has no CPU or PyTorch fallback for its kernels)"
        )
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    v = lib.fmlp_abi_version()
    if v != ABI_VERSION:
        raise FedMLPNativeError(f"ABI version mismatch: library {v}, bindings {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def lib():
    return load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().fmlp_status_string(rc).decode()
        raise FedMLPNativeError(f"{what or 'fedmlp_b200 call'} failed with status {rc}: {msg}")


# ----------------------------------------------------------------------------- helpers
def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream on `device` (kernels are launched on it)."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FedMLPNativeError(
                "fedmlp_b200 kernels need CUDA tensors; got a tensor on "
                f"{t.device} (there is no CPU fallback)"
            )


def i64_array(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def u32_array(vals):
    return (C.c_uint32 * len(vals))(*[int(v) & 0xFFFFFFFF for v in vals])


def u64_array(vals):
    return (C.c_uint64 * len(vals))(*[int(v) for v in vals])


def f32_array(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def f64_array(vals):
    return (C.c_double * len(vals))(*[float(v) for v in vals])


def ptr_array(vals):
    return (C.c_void_p * len(vals))(*[int(v) for v in vals])


def class_mask(classes) -> int:
    m = 0
    for c in classes:
        c = int(c)
        if not 0 <= c < MAX_CLASSES:
            raise ValueError(f"class index {c} outside [0, {MAX_CLASSES})")
        m |= 1 << c
    return m
