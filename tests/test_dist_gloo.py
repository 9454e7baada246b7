"""N>1 path on CPU: world_size-2 gloo processes exercise the distributed aggregation algebra
(client sharding, weight pre-normalisation, all-reduce, int64 counters, prototype / tao tails).
The rank-local reduce is injected from the oracle here because the CUDA kernel needs a GPU; on
the GPU box the same functions run with the kernel (tests/test_gpu_dist.py, bench.py --gpus N)."""
import os
import socket
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fedmlp_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_clients(K, seed=0):
    g = torch.Generator().manual_seed(seed)
    base = OrderedDict(w=torch.randn(33, 7, generator=g), b=torch.randn(7, generator=g))
    clients = []
    for k in range(K):
        clients.append(OrderedDict(w=base["w"] + 0.02 * torch.randn(33, 7, generator=g),
                                   b=base["b"] + 0.02 * torch.randn(7, generator=g),
                                   n=torch.tensor(100 + 3 * k, dtype=torch.int64)))
    return clients


def _oracle_local_reduce(bufs, weights, out):
    acc = bufs[0] * np.float32(weights[0])
    for b, w in zip(bufs[1:], weights[1:]):
        acc = acc + b * np.float32(w)
    out.copy_(acc)
    return out


def _worker(rank, world, port, K, C, D, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fedmlp_b200 import dist as fd
        clients = _make_clients(K)
        weights = [5000 + 7 * k for k in range(K)]
        mine = fd.shard_clients(K, world, rank)
        out = fd.FedAvg_distributed([clients[i] for i in mine], [weights[i] for i in mine],
                                    local_reduce=_oracle_local_reduce)
        # prototypes / tao
        g = torch.Generator().manual_seed(5)
        protos = [torch.randn(2 * C, D, generator=g) for _ in range(K)]
        taos = [torch.rand(C, generator=g).double().numpy() for _ in range(K)]
        active = [[k for k in range(K) if k % C == c] for c in range(C)]
        active[C - 1] = []                                   # a class nobody annotates -> NaN rows
        missing = [[k for k in range(K) if k % C != c] for c in range(C)]
        pos = {gid: p for p, gid in enumerate(mine)}
        act_local = [[pos[k] for k in lst if k in pos] for lst in active]
        mis_local = [[pos[k] for k in lst if k in pos] for lst in missing]
        proto = fd.FedAvg_proto_distributed([protos[i] for i in mine], [weights[i] for i in mine], act_local, C,
                                            local_proto_avg=O.fedavg_proto)
        tao = fd.FedAvg_tao_distributed([taos[i] for i in mine], [weights[i] for i in mine], mis_local, C)
        ret[rank] = dict(out={k: v.clone() for k, v in out.items()}, proto=proto.clone(), tao=tao.copy(),
                         mine=list(mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [5, 8])
def test_distributed_aggregation_matches_single_process(K):
    world, C, D = 2, 4, 16
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), K, C, D, ret), nprocs=world, join=True)
    clients = _make_clients(K)
    weights = [5000 + 7 * k for k in range(K)]
    ref = O.fedavg(clients, weights)
    assert sorted(ret[0]["mine"] + ret[1]["mine"]) == list(range(K))
    for rank in range(world):
        out = ret[rank]["out"]
        assert list(out.keys()) == list(ref.keys())
        for k in ref:
            assert out[k].dtype == torch.float32
            if k == "n":
                assert torch.equal(out[k], ref[k])                     # int64 path is exact
            else:
                np.testing.assert_allclose(out[k].numpy(), ref[k].numpy(), rtol=1e-5, atol=1e-6)
    # both ranks hold the identical global model
    for k in ref:
        assert torch.equal(ret[0]["out"][k], ret[1]["out"][k])
    g = torch.Generator().manual_seed(5)
    protos = [torch.randn(2 * C, D, generator=g) for _ in range(K)]
    taos = [torch.rand(C, generator=g).double().numpy() for _ in range(K)]
    active = [[k for k in range(K) if k % C == c] for c in range(C)]
    active[C - 1] = []
    missing = [[k for k in range(K) if k % C != c] for c in range(C)]
    ref_proto = O.fedavg_proto(protos, weights, active)
    ref_tao = O.fedavg_tao(taos, weights, missing)
    for rank in range(world):
        p = ret[rank]["proto"]
        assert torch.isnan(p[2 * (C - 1):]).all() and torch.isnan(ref_proto[2 * (C - 1):]).all()
        np.testing.assert_allclose(p[:2 * (C - 1)].numpy(), ref_proto[:2 * (C - 1)].numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ret[rank]["tao"], ref_tao, rtol=1e-12)


def test_shard_clients_partitions():
    from fedmlp_b200.dist import shard_clients
    for n in (1, 5, 8, 64, 65):
        for world in (1, 2, 4, 8):
            parts = [list(shard_clients(n, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_fused_allreduce_chunk_layout():
    """Host-side slicing of the fused fold + all-reduce kernel (fmlp_fedavg_allreduce_f32 contract:
    slice_len % (4 * n_chunks) == 0 and world * slice_len >= P)."""
    from fedmlp_b200.dist import FusedFedAvgAllReduce as A
    for P in (4, 1000, 7042752, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            for nc in (0, 1, 2, 4, 7, 16, 99):
                n, L = A.chunk_layout(P, world, nc)
                assert 1 <= n <= A.MAX_CHUNKS
                assert L % (4 * n) == 0 and world * L >= P
                assert world * (L - 4 * n) < P or L == 4 * n          # no more than one vector of padding per slice
