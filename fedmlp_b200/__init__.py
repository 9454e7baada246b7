"""fedmlp_b200 — B200-native (sm_100a) implementation of FedMLP's per-round hot path.

Reference call surface kept (szbonaldo/FedMLP):
    FedAvg, Fed_w, FedAvg_proto, FedAvg_tao            utils/FedAvg.py
    fedmlp_stage1_loss, fedmlp_stage2_loss             loss blocks of LocalUpdate.train_FedMLP
    build_prototypes, tag_similarity, TagBatch         prototype / pseudo-label tagging blocks
    pool_tag, build_sim_table, FusedTail               model tail (relu + avg-pool) fused with the tagging scores
    LocalUpdate (train_FedMLP)                         utils/local_training.py  (fedmlp_b200.local_training)
All arithmetic runs in hand-written CUDA kernels (fedmlp_b200/csrc, C ABI in include/fedmlp_b200.h)
loaded through ctypes; there is no CPU / PyTorch fallback.
"""
from . import _cabi
from .fedavg import (DaAgg, FedAvg, FedAvg_proto, FedAvg_rela, FedAvg_tao, Fed_w, RSCFed, fedavg_flat_buffers,
                     model_dist)
from .flat import FlatLayout, FlatStateDict, flatten_module_, layout_of
from .losses import (fedmlp_stage1_loss, fedmlp_stage2_loss, fused_loss_and_grad_stage1,
                     fused_loss_and_grad_stage2)
from .optim import FlatAdam
from .pooling import FusedTail, SimTable, build_sim_table, pool_tag
from .prototypes import PrototypeResult, build_prototypes
from .tagging import TagBatch, tag_similarity

__all__ = [
    "FedAvg", "Fed_w", "FedAvg_proto", "FedAvg_tao", "FedAvg_rela", "RSCFed", "DaAgg", "model_dist", "fedavg_flat_buffers",
    "FlatLayout", "FlatStateDict", "flatten_module_", "layout_of", "FlatAdam",
    "fedmlp_stage1_loss", "fedmlp_stage2_loss", "fused_loss_and_grad_stage1", "fused_loss_and_grad_stage2",
    "PrototypeResult", "build_prototypes", "TagBatch", "tag_similarity",
    "FusedTail", "SimTable", "build_sim_table", "pool_tag",
]
__version__ = "0.1.0"
