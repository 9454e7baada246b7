"""The distributed aggregation entry points with the CUDA kernel as the rank-local reduce
(world size 1 on the test box; the 2-rank algebra is covered by tests/test_dist_gloo.py and the
N-GPU path by bench.py --gpus N)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import fedmlp_oracle as O

pytestmark = pytest.mark.gpu


def test_fedavg_distributed_single_rank_cuda(lib):
    from fedmlp_b200 import dist as fd
    from fedmlp_b200.shapes import densenet121_state_shapes, synth_state_dict
    shapes = OrderedDict(list(densenet121_state_shapes(5).items())[:60])
    base = synth_state_dict(shapes, 0)
    clients = [synth_state_dict(shapes, 1 + k, base=base, counter=100 + k) for k in range(6)]
    weights = [5000, 4999, 5001, 1234, 777, 6875]
    ref = O.fedavg(clients, weights)
    out = fd.FedAvg_distributed([OrderedDict((k, v.cuda()) for k, v in c.items()) for c in clients], weights)
    assert list(out.keys()) == list(ref.keys())
    for k in ref:
        assert out[k].dtype == torch.float32
        if ref[k].numel() == 1 and "num_batches" in k:
            assert torch.equal(out[k].cpu(), ref[k])
        else:
            scale = float(ref[k].abs().max()) + 1e-30
            np.testing.assert_allclose(out[k].cpu().numpy(), ref[k].numpy(), rtol=1e-5, atol=1e-6 * scale)


def test_proto_tao_distributed_single_rank_cuda(lib):
    from fedmlp_b200 import dist as fd
    K, C, D = 6, 5, 1024
    g = torch.Generator().manual_seed(2)
    protos = [torch.randn(2 * C, D, generator=g) for _ in range(K)]
    weights = [5000 + k for k in range(K)]
    active = [[k for k in range(K) if k % C == c] for c in range(C)]
    active[3] = []
    ref = O.fedavg_proto(protos, weights, active)
    out = fd.FedAvg_proto_distributed([p.cuda() for p in protos], weights, active, C).cpu()
    ok = ~torch.isnan(ref)
    assert torch.isnan(out[~ok]).all()
    np.testing.assert_allclose(out[ok].numpy(), ref[ok].numpy(), rtol=1e-5, atol=1e-6)
    taos = [torch.rand(C, generator=g).double().numpy() for _ in range(K)]
    missing = [[k for k in range(K) if k % C != c] for c in range(C)]
    np.testing.assert_allclose(fd.FedAvg_tao_distributed(taos, weights, missing, C), O.fedavg_tao(taos, weights, missing), rtol=1e-12)
