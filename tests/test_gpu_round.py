"""Round-level parity: the batched multi-client hot path (ClientShard.round_hot_path) against the
oracle run client by client, the way the reference's sequential loop (main.py:135) would."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import fedmlp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("two_streams", [False, True, "sim_first", "proto_first"])
@pytest.mark.parametrize("mode", ["pair", "folded"])
def test_round_hot_path_matches_per_client_oracle(lib, mode, two_streams):
    """two_streams: the DAG form bench.py replays ({sim, select, fill+loss} || {prototypes, FedAvg} with the small
    aggregation tails forked off behind the prototype pass) must give the same results as the serial order."""
    from fedmlp_b200.round import ClientShard
    side = torch.cuda.Stream(device=DEV) if two_streams else None
    schedule = two_streams if isinstance(two_streams, str) else "concurrent"   # ClientShard.schedule of the DAG form

    C, D, P = 5, 256, 10007 + 1          # P multiple of 4
    sizes = [700, 333, 1201, 64]
    S = len(sizes)
    active = [[s % C] for s in range(S)]
    cf, nf = 0.03, 0.06
    shard = ClientShard(sizes, C, active, device=DEV, clean_frac=cf, noise_frac=nf, sim_mode=mode)
    shard.schedule = schedule
    states = [O.TaggingState(list(range(n)), shard.missing[s]) for s, n in enumerate(sizes)]
    g = torch.Generator().manual_seed(11)
    flats = [torch.randn(P, generator=g) for _ in range(S)]
    for rnd in range(2):
        data = [O.synth_client(n, D, C, seed=100 * rnd + s) for s, n in enumerate(sizes)]
        data2 = [O.synth_client(n, D, C, seed=100 * rnd + 50 + s) for s, n in enumerate(sizes)]
        feat = torch.cat([d[0] for d in data]); labels = torch.cat([d[1] for d in data])
        logits = torch.cat([d[2] for d in data])
        feat2 = torch.cat([d[0] for d in data2]); logits2 = torch.cat([d[2] for d in data2])
        zg = torch.randn(sum(sizes), C, generator=g) * 2
        proto = O.synth_prototypes(feat, labels)
        counters = [torch.arange(7, dtype=torch.int64) * (s + 1) + 100 * rnd for s in range(S)]
        res = shard.round_hot_path(feat.to(DEV), proto.to(DEV), logits.to(DEV), zg.to(DEV), labels.to(DEV),
                                   feat2.to(DEV), logits2.to(DEV), [f.to(DEV) for f in flats], sizes,
                                   aggregate_tails=True, counters=[c.to(DEV) for c in counters], side_stream=side)
        torch.cuda.synchronize()
        t = res.protos.t()
        for s, n in enumerate(sizes):
            r0, r1 = shard.seg_rows[s], shard.seg_rows[s + 1]
            sims_ref, stats = states[s].step(feat[r0:r1], proto, cf, nf)
            got = shard.tagger.traindata_idx(s)
            for j in range(2 * len(shard.missing[s])):
                if got[j] != [float(v) for v in states[s].traindata_idx[j]]:
                    # north_star waiver: only rows whose similarity is within 1e-6 of a picked one may differ
                    c = shard.missing[s][j // 2]
                    sims = sims_ref[c].numpy()
                    diff = {int(v) for v in got[j]} ^ set(states[s].traindata_idx[j])
                    picked = [sims[p] for p in stats[j // 2]["clean"] + stats[j // 2]["noise"]]
                    assert all(min(abs(sims[d] - q) for q in picked) < 1e-6 for d in diff)
                    states[s].traindata_idx[j] = [int(v) for v in got[j]]
            tgt, dis = O.mask_fill(labels[r0:r1].numpy(), list(range(n)), active[s], shard.missing[s], states[s].traindata_idx)
            ref_loss, rdz = O.loss_and_grads(lambda z, g_, y, d: O.stage2_loss(z, g_, y, d), logits[r0:r1], zg[r0:r1],
                                             torch.from_numpy(tgt), torch.from_numpy(dis), n_grad=1)
            # by-products of the fused fill + loss kernel (DatasetSplit_pseudo :1456-1477, sup_cls :1173)
            np.testing.assert_array_equal(shard._plan.y[r0:r1].cpu().numpy(), tgt)
            np.testing.assert_array_equal(shard._plan.distill[r0:r1].cpu().numpy(), dis)
            np.testing.assert_array_equal(shard._plan.sup[r0:r1].cpu().numpy(), 1.0 - dis)
            assert abs(float(res.losses[s]) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
            np.testing.assert_allclose(res.dz[r0:r1].cpu().numpy(), rdz.numpy(), rtol=1e-5, atol=1e-6 * float(rdz.abs().max()))
            ref_p, ref_n, ref_t = O.prototype_build(feat2[r0:r1], labels[r0:r1], logits2[r0:r1], active[s], shard.missing[s],
                                                    0.3, 0.7, guard_empty=True)
            np.testing.assert_allclose(res.protos.proto[s].cpu().numpy(), ref_p.numpy(), rtol=1e-5, atol=1e-5 * float(ref_p.abs().max()))
            assert res.protos.cnt[s].cpu().tolist() == ref_n
            np.testing.assert_allclose(t[s], ref_t, rtol=0, atol=1.5 / n)
        ref_glob = O.fedavg([OrderedDict(w=f) for f in flats], sizes)["w"]
        assert torch.equal(res.global_flat.cpu(), ref_glob)
        # the small tails of the server aggregation (main.py:218-234) on the clients' own outputs
        class_active = [[s for s in range(S) if c in active[s]] for c in range(C)]
        class_missing = [[s for s in range(S) if c in shard.missing[s]] for c in range(C)]
        ref_pg = O.fedavg_proto([res.protos.proto[s].cpu() for s in range(S)], sizes, class_active)
        got_pg = res.proto_glob.cpu()
        assert torch.equal(torch.isnan(got_pg), torch.isnan(ref_pg))
        assert torch.equal(torch.nan_to_num(got_pg), torch.nan_to_num(ref_pg))           # bit-exact kernel (FedAvg.py:72-93)
        np.testing.assert_allclose(res.tao.cpu().numpy(), O.fedavg_tao([t[s] for s in range(S)], sizes, class_missing), rtol=1e-12)
        ref_cnt = O.fedavg([OrderedDict(n=c) for c in counters], sizes)["n"]
        assert torch.equal(res.counters.cpu(), ref_cnt)                                   # int64 sums -> float32 (FedAvg.py:9-13)
