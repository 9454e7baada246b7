"""Two-rank NCCL tests (need >= 2 GPUs; skipped on a single-GPU box): the fused
fold + all-reduce kernel over symmetric memory against the single-process oracle, and against the
NCCL path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, P, n_chunks, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fedmlp_b200 import dist as fd
        g = torch.Generator().manual_seed(7)
        all_bufs = [torch.randn(P, generator=g) for _ in range(K * world)]
        weights = [5000 + 13 * i for i in range(K * world)]
        total = float(sum(weights))
        mine = list(range(rank * K, (rank + 1) * K))
        bufs = [all_bufs[i].cuda() for i in mine]
        wn = [weights[i] / total for i in mine]
        fused = fd.FusedFedAvgAllReduce(P, n_chunks=n_chunks)
        outs = []
        for it in range(3):                                   # epochs advance, buffers are reused
            out = fused(bufs, wn).clone()
            torch.cuda.synchronize()
            outs.append(out.cpu())
        nccl = fd.fedavg_flat_distributed(bufs, [weights[i] for i in mine], total_weight=total).cpu()
        ret[rank] = dict(fused=outs, nccl=nccl)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("K,P,n_chunks", [(3, 1000, 4), (8, 1 << 20, 4), (1, 4, 4), (8, 1 << 20, 1), (5, 4 * 7771, 16),
                                          (8, 7042752, 4)])
def test_fused_fedavg_allreduce_two_ranks(lib, K, P, n_chunks):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), K, P, n_chunks, ret), nprocs=world, join=True)
    g = torch.Generator().manual_seed(7)
    all_bufs = [torch.randn(P, generator=g) for _ in range(K * world)]
    weights = [5000 + 13 * i for i in range(K * world)]
    acc = all_bufs[0].double() * weights[0]
    for b, w in zip(all_bufs[1:], weights[1:]):
        acc += b.double() * w
    ref = (acc / sum(weights)).float()
    for rank in range(world):
        for out in ret[rank]["fused"]:
            np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
            assert torch.equal(out, ret[0]["fused"][0])           # bit-identical on every rank and every call
        np.testing.assert_allclose(ret[rank]["nccl"].numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------
# Round-2 aggregation exchange (csrc/fedavg_allreduce_q.cu): work-queue kernel, NVLS or peer-to-peer path,
# parameters + prototype tail + fp64 tail (class weights, tao, int64 counters) in one exchange.
# world = 1 runs on a single GPU (the driver's test box), world = 2 needs two.
def _agg_inputs(world, K, P, C, D, J):
    g = torch.Generator().manual_seed(11)
    n = K * world
    flats = [torch.randn(P, generator=g) * 0.05 for _ in range(n)]
    active = [[k % C] for k in range(n)]
    missing = [[c for c in range(C) if c != k % C] for k in range(n)]
    protos = []
    for k in range(n):
        p = torch.zeros(2 * C, D)
        for c in active[k]:
            p[2 * c:2 * c + 2] = torch.randn(2, D, generator=g)
        protos.append(p)
    rows = [500 + 7 * k for k in range(n)]
    weights = [float(r) for r in rows]
    tcnt = [torch.randint(0, rows[k], (C,), generator=g, dtype=torch.int32) for k in range(n)]
    counters = [torch.randint(0, 2000, (J,), generator=g, dtype=torch.int64) for k in range(n)]
    return flats, protos, tcnt, counters, rows, weights, active, missing


def _agg_worker(rank, world, port, K, P, C, D, J, n_chunks, multicast, split, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fedmlp_b200 import dist as fd
        flats, protos, tcnt, counters, rows, weights, active, missing = _agg_inputs(world, K, P, C, D, J)
        mine = list(range(rank * K, (rank + 1) * K))
        agg = fd.FedMLPAggregation(P, C, D, J, n_chunks=n_chunks, use_multicast=bool(multicast), split=bool(split))
        outs = []
        for it in range(3):                                   # epochs advance, buffers are reused
            params, proto, tao, cnt = agg([flats[i].cuda() for i in mine], [protos[i].cuda() for i in mine],
                                          torch.stack([tcnt[i] for i in mine]).cuda(), [weights[i] for i in mine],
                                          [rows[i] for i in mine], [active[i] for i in mine], [missing[i] for i in mine],
                                          float(sum(weights)), [counters[i].cuda() for i in mine])
            torch.cuda.synchronize()
            outs.append(dict(params=params.cpu().clone(), proto=proto.cpu().clone(), tao=tao.cpu().clone(),
                             cnt=None if cnt is None else cnt.cpu().clone()))
        ret[rank] = dict(outs=outs, path=agg.exchange.path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,K,P,C,D,J,n_chunks,multicast,split", [
    (1, 3, 4 * 1000, 5, 64, 7, 4, 0, 0), (1, 8, 1 << 20, 5, 1024, 121, 4, 0, 0), (1, 2, 4, 14, 32, 0, 16, 0, 0),
    (1, 8, 1 << 20, 5, 1024, 121, 4, 0, 1), (1, 3, 4 * 1000, 14, 128, 0, 2, 0, 1),
    (2, 3, 4 * 1000, 5, 64, 7, 4, 0, 0), (2, 3, 4 * 1000, 5, 64, 7, 4, 1, 0), (2, 8, 7042752, 5, 1024, 121, 4, 1, 0),
    (2, 8, 7042752, 5, 1024, 121, 4, 0, 0), (2, 5, 4 * 7771, 14, 1280, 49, 16, 1, 0), (2, 1, 4, 5, 4, 0, 1, 1, 0),
    (2, 8, 7042752, 5, 1024, 121, 4, 1, 1), (2, 5, 4 * 7771, 14, 1280, 49, 8, 0, 1)])
def test_queued_aggregation(lib, world, K, P, C, D, J, n_chunks, multicast, split):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import fedmlp_oracle as O
    ret = mp.Manager().dict()
    mp.spawn(_agg_worker, args=(world, _free_port(), K, P, C, D, J, n_chunks, multicast, split, ret), nprocs=world, join=True)
    flats, protos, tcnt, counters, rows, weights, active, missing = _agg_inputs(world, K, P, C, D, J)
    n = K * world
    acc = flats[0].double() * weights[0]
    for b, w in zip(flats[1:], weights[1:]):
        acc += b.double() * w
    ref_params = (acc / sum(weights)).float()
    class_active = [[k for k in range(n) if c in active[k]] for c in range(C)]
    class_missing = [[k for k in range(n) if c in missing[k]] for c in range(C)]
    ref_proto = O.fedavg_proto(protos, weights, class_active)                        # utils/FedAvg.py:72-93
    taos = [(tcnt[k].double() / rows[k]).numpy() for k in range(n)]
    ref_tao = O.fedavg_tao(taos, weights, class_missing)                             # utils/FedAvg.py:51-70
    ref_cnt = torch.stack([counters[k] * int(weights[k]) for k in range(n)]).sum(0).float() / float(sum(weights)) if J else None
    for rank in range(world):
        for out in ret[rank]["outs"]:
            np.testing.assert_allclose(out["params"].numpy(), ref_params.numpy(), rtol=1e-5, atol=1e-6)
            ok = ~torch.isnan(ref_proto)
            assert torch.isnan(out["proto"][~ok]).all()
            np.testing.assert_allclose(out["proto"][ok].numpy(), ref_proto[ok].numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(out["tao"].numpy(), ref_tao, rtol=1e-12)
            if J:
                assert torch.equal(out["cnt"], ref_cnt)                              # exact int64 sums, FedAvg.py:9-13
            for key in ("params", "proto", "tao"):                                   # bit-identical on every rank / call
                a, b = out[key], ret[0]["outs"][0][key]
                assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
