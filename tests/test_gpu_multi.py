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
