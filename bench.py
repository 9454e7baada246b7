#!/usr/bin/env python
"""bench.py — FedMLP round hot path (tag + loss + prototypes + aggregation) on B200.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--config NAME]

A "step" is one pass of the per-round hot path over the clients resident on one GPU:
  tag (similarity + selection + mask fill) over N_total x D features of the incoming global model,
  stage-2 loss forward+backward over every client's rows, prototype/t construction over the
  N_total x D features of the locally trained model, and the server aggregation of main.py:218-234
  (FedAvg over the K client parameter buffers, FedAvg_proto, FedAvg_tao, the int64 BatchNorm counters).
Headline workload = BASELINE.json configs[1] ("ich55k": ICH 5-class DenseNet121, 8 clients on one B200,
55,000 synthetic feature rows, D=1024, P=7,042,629).  The C=14 workloads of configs[2] ("cxr14_64c") and
configs[3] ("effb0_85k") are run behind it and reported in the `configs` array of the same JSON line.
With --gpus N>1 every rank holds its own clients (weak scaling) and only the aggregation crosses GPUs: one
fused fold + all-reduce kernel per rank over NVLink / NVSwitch (csrc/fedavg_allreduce_q.cu); the result is
checked against an fp64 fold and the NCCL path inside the run (`parity_ok`, exit code 3 on mismatch).

Prints ONE JSON line (contract in the task statement): value = client-samples/s with inputs resident in
HBM (every step one CUDA-graph replay at any N); e2e = the same through host buffers (pinned H2D of every
input and D2H of every result inside the timed region); roofline = dominant kernel vs the measured HBM
peak; cpu_baseline = the reference's own CPU code (oracle/ref_round.py) timed on this box's cores.
`--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ICH_PREVALENCE = [0.015, 0.176, 0.128, 0.174, 0.230]
METRIC = "client_samples_per_sec_per_fedmlp_round"
UNIT = "client-samples/s"

# BASELINE.json configs[1..3] as synthetic shapes
CONFIGS = {
    "ich55k": dict(baseline_config=1, classes=5, dim=1024, clients_per_gpu=8, rows_per_client=6875, backbone="DenseNet121",
                   signed=False, loss="sup",
                   title="ICH 5-class DenseNet121, 8 clients/GPU x 6,875 rows (55k features per GPU), D=1024"),
    "cxr14_64c": dict(baseline_config=2, classes=14, dim=1024, clients_per_gpu=8, rows_per_client=5889, backbone="DenseNet121",
                      signed=False, loss="sup",
                      title="ChestXray14 14-class DenseNet121, 8 clients/GPU x 5,889 rows (64 clients on 8 GPUs), D=1024"),
    "effb0_85k": dict(baseline_config=3, classes=14, dim=1280, clients_per_gpu=1, rows_per_client=85000, backbone="EfficientNet-B0",
                      signed=True, loss="sup_dis",
                      title="ChestXray14 14-class EfficientNet-B0, one client x 85,000 rows/GPU, D=1280, 13 missing classes, "
                            "BCE + teacher-consistency loss"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="ich55k", choices=sorted(CONFIGS), help="headline workload (BASELINE configs[1] by default)")
    ap.add_argument("--extra-configs", default="auto", help="comma list of further configs for the `configs` array; "
                                                            "auto = the other BASELINE configs, none = skip")
    ap.add_argument("--clients-per-gpu", type=int, default=0, help="override the config")
    ap.add_argument("--rows-per-client", type=int, default=0, help="override the config")
    ap.add_argument("--sim-mode", default="folded", choices=["pair", "folded"])
    ap.add_argument("--cpu-clients", type=int, default=0, help="clients in the CPU arm (0 = all of the workload)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--collective", default="auto", choices=["auto", "queue_split", "push_split", "queue", "fused_r01", "nccl"],
                    help="N>1: auto = push_split up to 4 GPUs, queue_split (NVLS) beyond — the measured best of each "
                         "(profiles/r02_multi_gpu_variants.txt); push_split = the parameters through the fold + two-shot peer-store kernel "
                         "on its own stream, the packed tails through the work-queue kernel; queue_split = work-queue fold + NVLS/P2P all-reduce kernel for the parameters on its own stream "
                         "from the start of the round + a second single-chunk launch for the packed tails after the prototype pass "
                         "(default); queue = one exchange for parameters + tails; fused_r01 = the round-1 cooperative peer-store "
                         "kernel (parameters only); nccl = local fold + NCCL all-reduce (parameters only)")
    ap.add_argument("--streams", type=int, default=0, choices=[0, 2, 3],
                    help="3: tagging/loss chain || prototypes + tails || parameter aggregation; 2: the last two share a stream; "
                         "0 = auto (2 on one GPU, where all three chains are HBM-bound; 3 with a collective)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    return ap.parse_args()


class Workload:
    def __init__(self, name, a=None):
        c = dict(CONFIGS[name])
        self.name = name
        self.C, self.D = c["classes"], c["dim"]
        self.S = (a.clients_per_gpu if a is not None and a.clients_per_gpu else c["clients_per_gpu"])
        self.n = (a.rows_per_client if a is not None and a.rows_per_client else c["rows_per_client"])
        self.backbone, self.signed, self.loss = c["backbone"], c["signed"], c["loss"]
        self.baseline_config = c["baseline_config"]
        self.title = c["title"] if (self.S, self.n) == (c["clients_per_gpu"], c["rows_per_client"]) else \
            f"{c['title'].split(',')[0]}, {self.S} clients/GPU x {self.n} rows, D={self.D}"

    def description(self):
        return f"{self.title}: tag + loss + prototypes + aggregation per round (BASELINE.json configs[{self.baseline_config}])"

    def config_keys(self, P=None):
        return {"workload": self.description(), "name": self.name, "clients_per_gpu": self.S, "rows_per_client": self.n,
                "feature_dim": self.D, "classes": self.C, "backbone": self.backbone, "params_per_client": P}

    def full_config(self, a, world):
        """The `config` object of the JSON line — the SAME for the GPU arm and the reference arm of one (workload, N,
        flags), computed from shapes alone (no device): workload keys + how L2 is kept out of the measurement +
        how the clients are spread over the GPUs."""
        import torch
        from fedmlp_b200.shapes import count_params
        shapes = self.state_shapes()
        P = count_params(shapes)[0]
        Ppad = sum((int(torch.Size(sh).numel()) + 3) // 4 * 4 for sh, dt in shapes.values() if dt == torch.float32)
        ab = alg_bytes(self, {"N": self.S * self.n, "Ppad": Ppad})
        total = ab["sim"] + ab["proto"] + ab["fedavg"] + ab["loss"] + ab["select_fill"]
        cfg = self.config_keys(P)
        cfg.update({"sim_mode": a.sim_mode,
                    "l2": f"inputs larger than L2: {round(total / 1e6)} MB streamed per step vs 126 MB L2, no flush needed",
                    "parallelism": (f"clients sharded over {world} GPU(s), one process per GPU; only the aggregation crosses GPUs, on "
                                    "its own streams, concurrent with the tagging/loss chain") if world > 1 else "single GPU",
                    "not_in_timed_step": "host-list bookkeeping of traindata_idx (two small device clones per round, keep_history)"})
        return cfg

    def state_shapes(self):
        from fedmlp_b200.shapes import densenet121_state_shapes, efficientnet_b0_state_shapes
        return densenet121_state_shapes(self.C) if self.backbone == "DenseNet121" else efficientnet_b0_state_shapes(self.C)


# ----------------------------------------------------------------------------------- workload
def make_device_inputs(w: Workload, rank, dev):
    """Seeded synthetic inputs of the named shapes, created on the device (setup, untimed)."""
    import torch
    from fedmlp_b200.shapes import count_params

    S, n, C, D = w.S, w.n, w.C, w.D
    N = S * n
    g = torch.Generator(device=dev).manual_seed(1037 + rank)
    prev = torch.tensor(ICH_PREVALENCE if C == 5 else torch.linspace(0.02, 0.20, C).tolist(), device=dev)
    labels = (torch.rand(N, C, generator=g, device=dev) < prev).float()
    shift = torch.randn(C, D, generator=g, device=dev) * 0.25

    def feats():
        x = torch.randn(N, D, generator=g, device=dev)
        if w.signed:      # EfficientNet-style signed features (swish tail), SURVEY §8d
            x = x * torch.sigmoid(torch.randn(N, D, generator=g, device=dev))
            return (x + labels @ shift).contiguous()
        return torch.relu(x) + torch.relu(labels @ shift)

    feat_tag, feat_proto = feats(), feats()
    logits = torch.randn(N, C, generator=g, device=dev) * 2
    logits_glob = torch.randn(N, C, generator=g, device=dev) * 2
    logits_proto = torch.randn(N, C, generator=g, device=dev) * 2
    # server prototypes = masked means of the features (small, sign-mixed similarities as in practice)
    rows = []
    for c in range(C):
        for v in (0.0, 1.0):
            m = labels[:, c] == v
            rows.append(feat_tag[m].mean(0) if bool(m.any()) else torch.zeros(D, device=dev))
    proto = torch.stack(rows).contiguous()
    shapes = w.state_shapes()
    P, J = count_params(shapes)
    Ppad = sum((int(torch.Size(s).numel()) + 3) // 4 * 4 for s, d in shapes.values() if d == torch.float32)
    base = torch.randn(Ppad, generator=g, device=dev) * 0.05
    flats = [(base + 0.02 * torch.randn(Ppad, generator=g, device=dev)).contiguous() for _ in range(S)]
    counters = [torch.full((J,), 100 + rank * S + k, dtype=torch.int64, device=dev) for k in range(S)]
    weights = [n] * S
    return dict(N=N, P=P, J=J, Ppad=Ppad, labels=labels, feat_tag=feat_tag, feat_proto=feat_proto, logits=logits,
                logits_glob=logits_glob, logits_proto=logits_proto, proto=proto, flats=flats, weights=weights,
                counters=counters, n_keys=len(shapes))


def alg_bytes(w: Workload, inp):
    """Algorithmic bytes per launch."""
    N, D, C, K, P = inp["N"], w.D, w.C, w.S, inp["Ppad"]
    M = C - 1
    return {
        "sim": 4 * N * D + 4 * 2 * C * D + 4 * M * N,
        "proto": 4 * N * D + 4 * N * C * 2 + 4 * 2 * C * D * K,
        "fedavg": (K + 1) * 4 * P,
        "loss": 5 * 4 * N * C,
        "select_fill": 4 * M * N * 4 + 2 * 4 * N * C,
    }


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.path = f"/tmp/fedmlp_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------- CPU arm
def run_cpu_arm(w: Workload, steps, warmup, n_clients=0, budget_s=150.0):
    """The reference's CPU path over the workload's clients (all of them unless n_clients says otherwise):
    oracle/ref_round.py calls the reference's own FedAvg / FedAvg_proto / FedAvg_tao / CosineSimilarityFast /
    max_m_indices / min_n_indices / DatasetSplit_pseudo / LogitAdjust_Multilabel (unmodified files, mounted
    or vendored under oracle/_ref) and restates the inline blocks of train_FedMLP around them."""
    import torch
    from fedmlp_b200.shapes import count_params, synth_state_dict
    from oracle import ref_loader, ref_round

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    have_ref = ref_loader.available()
    K = w.S
    nc = n_clients or K
    clients, proto = ref_round.make_clients(nc, w.n, w.D, w.C, signed=w.signed)
    shapes = w.state_shapes()
    P, J = count_params(shapes)
    base = synth_state_dict(shapes, 1037)
    sds = [synth_state_dict(shapes, 1038 + k, base=base, counter=100 + k) for k in range(K)]
    weights = [w.n] * K
    if have_ref:
        ref = ref_round.load_reference()
        kind = "reference"
        step = lambda: ref_round.round_step(ref, clients, proto, sds, weights, w.C)
    else:      # neither /root/reference nor oracle/_ref: the oracle's restatement of the same path
        from oracle import fedmlp_oracle as O
        kind = "port"
        for cl in clients:
            cl["state"] = O.TaggingState(cl["idxs"], cl["missing"])

        def step():
            t0 = time.perf_counter()
            for cl in clients:
                cl["state"].traindata_idx = [[] for _ in cl["state"].traindata_idx]
                cl["state"].step(cl["feat"], proto, 0.005, 0.01, python_sort=True)
                tgt, dis = O.mask_fill(cl["labels"].numpy(), cl["state"].dataset_idx, cl["active"], cl["missing"],
                                       cl["state"].traindata_idx)
                O.loss_and_grads(lambda z, zg, y, d: O.stage2_loss(z, zg, y, d), cl["logits"], cl["zg"],
                                 torch.from_numpy(tgt), torch.from_numpy(dis), n_grad=1)
                O.prototype_build(cl["feat2"], cl["labels"], cl["logits2"], cl["active"], cl["missing"], 0.3, 0.7, True)
            t1 = time.perf_counter()
            O.fedavg(sds, weights)
            return dict(t_clients=t1 - t0, t_fedavg=time.perf_counter() - t1, t_tail=0.0)

    t_begin = time.perf_counter()
    for _ in range(warmup):
        step()
        if time.perf_counter() - t_begin > 0.4 * budget_s:
            break
    tc, tf, tt = [], [], []
    for _ in range(steps):
        r = step()
        tc.append(r["t_clients"]); tf.append(r["t_fedavg"]); tt.append(r["t_tail"])
        if time.perf_counter() - t_begin > budget_s:      # bounded: a few minutes whatever K/W the driver passes
            break
    scale = K / nc
    t_step = statistics.mean(tc) * scale + statistics.mean(tf) + statistics.mean(tt)
    sample = (f"{nc} of {K} clients ({nc * w.n} rows) through tag + mask fill + stage-2 loss fwd/bwd + prototypes"
              + (f" (time x{scale:g})" if scale != 1 else "") + f" + FedAvg over {K} {w.backbone} state_dicts "
              f"({len(shapes)} tensors) + FedAvg_proto / FedAvg_tao, {len(tc)} steps; "
              + ("the reference's own utils/FedAvg.py, CosineSimilarityFast, max_m/min_n_indices, DatasetSplit_pseudo, "
                 f"LogitAdjust_Multilabel ({ref_loader.source()} files) with the inline blocks of train_FedMLP restated around them"
                 if have_ref else "oracle port (reference files not available)"))
    return dict(value=K * w.n / t_step, t_step=t_step, t_clients=statistics.mean(tc) * scale, t_fedavg=statistics.mean(tf),
                cores=cores, P=P, kind=kind, sample=sample, steps_executed=len(tc))


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = Workload(a.config, a)
    r = run_cpu_arm(w, max(1, a.steps), max(0, min(a.warmup, 2)), n_clients=a.cpu_clients)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["t_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": w.full_config(a, a.gpus),
        "steps_executed": r["steps_executed"],
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "fedavg_gbs": (w.S + 1) * 4 * r["P"] / r["t_fedavg"] / 1e9,
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_result(json.dumps(line))


# ----------------------------------------------------------------------------------- GPU arm
def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel):
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


def bind_numa(dev_index):
    """Run this rank on the CPUs of its GPU's NUMA node, so that pinned buffers allocated afterwards sit
    next to the GPU's PCIe root (8 ranks pulling through one node halved the e2e rate in round 1)."""
    info = {"node": None, "cpus": None}
    try:
        import torch
        bus = torch.cuda.get_device_properties(dev_index).pci_bus_id
        dom = torch.cuda.get_device_properties(dev_index).pci_domain_id
        devid = torch.cuda.get_device_properties(dev_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        info["node"] = node
        if node >= 0:
            cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            cpus = set()
            for part in cpulist.split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus"] = len(allowed)
    except Exception as exc:
        info["error"] = f"{type(exc).__name__}: {exc}"[:120]
    return info


class Runner:
    """One workload on this rank: inputs, shard, aggregation, the step function."""

    def __init__(self, a, w: Workload, world, rank, dev, lib):
        import torch
        import torch.distributed as dist
        from fedmlp_b200 import _cabi as cabi
        from fedmlp_b200.round import ClientShard

        self.a, self.w, self.world, self.rank, self.dev, self.lib = a, w, world, rank, dev, lib
        self.inp = inp = make_device_inputs(w, rank, dev)
        S, C = w.S, w.C
        self.shard = ClientShard([w.n] * S, C, [[(rank * S + k) % C] for k in range(S)], device=dev, sim_mode=a.sim_mode)
        self.shard.keep_history = False      # host-list bookkeeping (two small device clones per round) is off the timed path
        self.shard.loss_variant = cabi.LOSS2_SUP_DIS if w.loss == "sup_dis" else cabi.LOSS2_SUP
        self.fed_out = torch.empty(inp["Ppad"], dtype=torch.float32, device=dev)
        # FMLP_MAIN_PRIORITY / FMLP_SIDE_PRIORITY: CUDA stream priorities of the tagging / loss chain (the stream the round's
        # graph is captured on) and of the prototype / aggregation chain (0 = default, -1 .. = higher); measured on one
        # GPU: a higher side priority costs 8 % (profiles/r02_exp_coresidency.txt)
        mp = int(os.environ.get("FMLP_MAIN_PRIORITY", "0"))
        self.capture_stream = torch.cuda.Stream(device=dev, priority=mp) if mp != 0 else None
        self.side_stream = torch.cuda.Stream(device=dev, priority=int(os.environ.get("FMLP_SIDE_PRIORITY", "0")))
        n_streams = a.streams or (2 if world == 1 else 3)
        # the parameter exchange gets a HIGH-priority stream: its (persistent, small-footprint) CTAs must be placed
        # before the prototype / similarity kernels of the other streams fill every SM's register file
        prio = int(os.environ.get("FMLP_AGG_PRIORITY", "-1"))
        self.agg_stream = torch.cuda.Stream(device=dev, priority=prio) if n_streams == 3 else None
        self.total_w = float(sum(inp["weights"]) * world)
        self.w_norm = [x / self.total_w for x in inp["weights"]]          # pre-normalised: the all-reduce yields the mean
        self.agg = self.fused = None
        self.collective = "none (single GPU: FedAvg + FedAvg_proto + FedAvg_tao + counter kernels)"
        if world > 1:
            self.collective = "nccl all_reduce after the local fold (parameters only)"
            if a.collective == "auto":
                a.collective = "push_split" if world <= 4 else "queue_split"
            try:
                if a.collective == "push_split":
                    from fedmlp_b200.dist import FedMLPAggregation
                    self.agg = FedMLPAggregation(inp["Ppad"], C, w.D, inp["J"], device=dev, split=True, params_impl="push")
                    self.collective = (f"parameters: fold + two-shot all-reduce by posted peer stores in one kernel ({self.agg.exchange.n_chunks} "
                                       "chunks) on its own high-priority stream from the start of the round; prototype sums + fp64 tail: "
                                       f"work-queue kernel ({self.agg.tails_exchange.path}), one chunk, after the prototype pass")
                elif a.collective in ("queue", "queue_split"):
                    from fedmlp_b200.dist import FedMLPAggregation
                    self.agg = FedMLPAggregation(inp["Ppad"], C, w.D, inp["J"], device=dev, split=(a.collective == "queue_split"))
                    ex = self.agg.exchange
                    self.collective = (f"work-queue fold + all-reduce kernel ({ex.path}), {ex.n_chunks} chunks: " +
                                       ("parameters in their own exchange from the start of the round; prototype sums + fp64 tail (class "
                                        f"weights, tao, {inp['J']} int64 counters) in a second single-chunk launch after the prototype pass"
                                        if self.agg.split else
                                        f"parameters + prototype sums + fp64 tail (class weights, tao, {inp['J']} int64 counters) in one exchange"))
                elif a.collective == "fused_r01":
                    from fedmlp_b200.dist import FusedFedAvgAllReduce
                    self.fused = FusedFedAvgAllReduce(inp["Ppad"], device=dev)
                    self.collective = (f"round-1 cooperative fold + two-shot all-reduce over peer stores ({self.fused.n_chunks} chunks), "
                                       "parameters only")
            except Exception as exc:
                self.agg = self.fused = None
                self.collective += f" (requested '{a.collective}' unavailable: {type(exc).__name__}: {exc})"[:240]
        self.dist = dist

    def step(self, timers=None, overlap=True, data=None, only=None):
        """One round hot path.  overlap: {prototypes -> aggregation} on a side stream, concurrent with
        {sim -> select -> fill -> loss}; the per-stage event timing (timers) runs the stages back to back on one
        stream so every kernel is timed alone."""
        side = self.side_stream if (overlap and timers is None and only is None) else None
        aggs = self.agg_stream if side is not None else None
        inp = data if data is not None else self.inp
        sh, w = self.shard, self.w
        args = (inp["feat_tag"], inp["proto"], inp["logits"], inp["logits_glob"], inp["labels"], inp["feat_proto"],
                inp["logits_proto"], inp["flats"])
        if self.world == 1:
            return sh.round_hot_path(*args, inp["weights"], timers=timers, fedavg_out=self.fed_out, side_stream=side,
                                     agg_stream=aggs, aggregate_tails=True, counters=inp["counters"], only=only)
        if self.agg is not None and self.agg.split:
            def params_fn(bufs, wts):
                return self.agg.aggregate_params(bufs, inp["weights"], self.total_w)

            def tails_fn(protos):
                return self.agg.aggregate_tails([protos.proto[k] for k in range(w.S)], protos.tcnt, inp["weights"], [w.n] * w.S,
                                                sh.active, sh.missing, self.total_w, inp["counters"])
            return sh.round_hot_path(*args, self.w_norm, timers=timers, fedavg_out=self.fed_out, divide=False,
                                     side_stream=side, agg_stream=aggs, params_fn=params_fn, tails_fn=tails_fn, only=only)
        if self.agg is not None:
            def agg_fn(bufs, wts, protos):
                return self.agg(bufs, [protos.proto[k] for k in range(w.S)], protos.tcnt, inp["weights"], [w.n] * w.S,
                                sh.active, sh.missing, self.total_w, inp["counters"])
            return sh.round_hot_path(*args, self.w_norm, timers=timers, fedavg_out=self.fed_out, divide=False,
                                     side_stream=side, aggregate_fn=agg_fn, only=only)
        if self.fused is not None:
            return sh.round_hot_path(*args, self.w_norm, timers=timers, fedavg_out=self.fed_out, divide=False,
                                     side_stream=side, aggregate_fn=lambda bufs, wts, protos: self.fused(bufs, wts), only=only)
        return sh.round_hot_path(*args, self.w_norm, timers=timers, fedavg_out=self.fed_out, divide=False,
                                 side_stream=side, after_aggregate=lambda g: self.dist.all_reduce(g), only=only)

    # ---- untimed parity gate for the multi-GPU exchange -------------------------------------------------
    def parity(self):
        """Global parameters of one step against (a) an fp64 weighted fold all-reduced in fp64 and (b) the
        local-fold + NCCL path; aggregated prototypes / tao / counters against fp64 sums.  All ranks must also
        hold bit-identical results."""
        import torch
        dist, inp, w = self.dist, self.inp, self.w
        res = self.step(overlap=False)
        torch.cuda.synchronize()
        got = res.global_flat[:inp["Ppad"]].clone()
        acc = torch.zeros(inp["Ppad"], dtype=torch.float64, device=self.dev)
        for b, wt in zip(inp["flats"], inp["weights"]):
            acc += b.double() * float(wt)
        dist.all_reduce(acc)
        ref = acc / self.total_w
        scale = float(ref.abs().max())
        out = {"params_max_rel_err_vs_fp64": float((got.double() - ref).abs().max()) / scale}
        from fedmlp_b200.dist import fedavg_flat_distributed
        nccl = fedavg_flat_distributed(inp["flats"], inp["weights"], total_weight=self.total_w)
        out["params_max_rel_err_vs_nccl_path"] = float((got - nccl).abs().max()) / scale
        chk = torch.tensor([float(got.double().sum())], dtype=torch.float64, device=self.dev)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["identical_on_all_ranks"] = bool(float(lo.item()) == float(hi.item()))
        ok = out["params_max_rel_err_vs_fp64"] <= 1e-5 and out["params_max_rel_err_vs_nccl_path"] <= 1e-5 and out["identical_on_all_ranks"]
        if res.proto_glob is not None:
            C, D, S = w.C, w.D, w.S
            pl = self.shard._plan
            psum = torch.zeros(2 * C, D, dtype=torch.float64, device=self.dev)
            wsum = torch.zeros(C, dtype=torch.float64, device=self.dev)
            tnum = torch.zeros(C, dtype=torch.float64, device=self.dev)
            tden = torch.zeros(C, dtype=torch.float64, device=self.dev)
            csum = torch.zeros(max(inp["J"], 1), dtype=torch.float64, device=self.dev)
            for k in range(S):
                wt = float(inp["weights"][k])
                psum += pl.proto[k].double() * wt
                for c in self.shard.active[k]:
                    wsum[c] += wt
                for c in self.shard.missing[k]:
                    tnum[c] += float(pl.tcnt[k, c]) / w.n * wt
                    tden[c] += wt
                if inp["J"]:
                    csum += inp["counters"][k].double() * wt
            for t in (psum, wsum, tnum, tden, csum):
                dist.all_reduce(t)
            ref_p = psum / wsum.repeat_interleave(2).unsqueeze(1)
            m = ~torch.isnan(ref_p)
            gp = res.proto_glob.double()
            out["proto_max_rel_err"] = float((gp[m] - ref_p[m]).abs().max()) / max(float(ref_p[m].abs().max()), 1e-30) if bool(m.any()) else 0.0
            out["proto_nan_rows_match"] = bool(torch.equal(torch.isnan(gp), torch.isnan(ref_p)))
            ref_t = torch.where(tden > 0, tnum / tden.clamp_min(1e-300), torch.ones_like(tden))
            out["tao_max_abs_err"] = float((res.tao - ref_t).abs().max())
            if inp["J"]:
                ref_c = (csum.to(torch.int64).to(torch.float32) / torch.tensor(self.total_w, dtype=torch.float32, device=self.dev))
                out["counters_exact"] = bool(torch.equal(res.counters, ref_c[:inp["J"]]))
            ok = ok and out["proto_max_rel_err"] <= 1e-5 and out["proto_nan_rows_match"] and out["tao_max_abs_err"] <= 1e-12 \
                and out.get("counters_exact", True)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["parity_ok"] = bool(int(flag.item()) == 1)
        return out

    # ---- timing -----------------------------------------------------------------------------------------
    def fence(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def measure(self, steps, warmup, want_graph=True):
        import torch
        a, w, world, dev, lib, inp = self.a, self.w, self.world, self.dev, self.lib, self.inp
        for _ in range(max(warmup, 3)):
            self.step()
        self.fence()
        # ---- timed region A (value): K steps, each ONE CUDA-graph replay of the round (any N: the fused
        #      aggregation kernel is an ordinary launch; the NCCL path stays eager)
        graph, graph_note, launches_per_step = None, "eager launches", None
        self.schedule_note = None
        if want_graph and not a.no_graph and (world == 1 or self.agg is not None or self.fused is not None):
            try:
                # Single GPU: the round's DAG can start the similarity and prototype kernels together or give the
                # similarity kernel the machine first (ClientShard.schedule; same work, same results).  Which is
                # faster depends on the shape, so both graphs are timed here, UNTIMED warm-up, and the better one is
                # the graph the timed region replays.
                candidates = ["concurrent", "sim_first"] if (world == 1 and not os.environ.get("FMLP_ROUND_SCHEDULE")) else [None]
                tried = {}
                best = None
                for sched in candidates:
                    if sched is not None:
                        self.shard.schedule = sched
                    l0 = lib.fmlp_launch_count()
                    g_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_, stream=self.capture_stream):
                        self.step()
                    n_launch = lib.fmlp_launch_count() - l0
                    t_ms = None
                    if len(candidates) > 1:
                        for _ in range(3):
                            g_.replay()
                        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        t0.record()
                        for _ in range(40):
                            g_.replay()
                        t1.record()
                        torch.cuda.synchronize()
                        t_ms = t0.elapsed_time(t1) / 40
                        tried[sched] = round(t_ms, 5)
                    if best is None or (t_ms is not None and t_ms < best[0]):
                        best = (t_ms, sched, g_, n_launch)
                _, sched, g_, launches_per_step = best
                if sched is not None:
                    self.shard.schedule = sched
                    self.schedule_note = {"chosen": sched, "warmup_ms_per_step": tried}
                graph = g_
                graph_note = (f"CUDA-graph replay of the {launches_per_step}-launch round, " +
                              ("three-stream DAG {sim,select,fill,loss} || {proto,tails} || {parameter aggregation}"
                               if self.agg_stream is not None and (world == 1 or (self.agg is not None and self.agg.split))
                               else "two-stream DAG {sim,select,fill,loss} || {proto,aggregation}"))
                for _ in range(3):
                    graph.replay()
            except Exception as exc:          # fall back to eager timing, say so
                graph, graph_note = None, f"eager launches (graph capture failed: {type(exc).__name__}: {exc})"[:200]
                torch.cuda.synchronize()
        self.fence()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if graph is not None:
            for _ in range(steps):
                graph.replay()
        else:
            for _ in range(steps):
                self.step()
        ev1.record()
        self.fence()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        ms_step = float(t.item()) / steps

        # ---- timed region B (kernels / roofline): the same K steps on ONE stream with a CUDA event after every
        #      stage, in the serial order sim, prototypes, aggregation, select+fill, loss.  With a graph: the
        #      serial round is captured with external event-record nodes and replayed back to back in groups of
        #      three — the events of the last replay of a group are read, so every stage is timed in steady state
        #      behind its real predecessor, without eager launch gaps.  Otherwise eager launches + events.
        launches0 = lib.fmlp_launch_count()
        stage_note = "eager launches, CUDA event after every stage"
        stage_graph = None
        if graph is not None:
            try:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2):
                    evs = self.step(timers="external").events
                stage_graph = g2
                stage_note = ("single-stream CUDA graph (sim, proto, aggregation, select+fill, loss) with an external CUDA event after "
                              "every stage; replays in back-to-back groups of 3, events of the last replay read")
            except Exception as exc:
                stage_graph = None
                stage_note += f" (graph with external events unavailable: {type(exc).__name__})"
                torch.cuda.synchronize()
        order = ["start", "sim", "proto", "fedavg", "select_fill", "loss"]
        results = []
        if stage_graph is not None:
            stage_graph.replay(); torch.cuda.synchronize()
            t_b = 0.0
            n_groups = max(1, min(steps, 200))
            for _ in range(n_groups):
                if world > 1:
                    self.dist.barrier()
                for _ in range(3):
                    stage_graph.replay()
                torch.cuda.synchronize()
                results.append({q: evs[p].elapsed_time(evs[q]) for p, q in zip(order[:-1], order[1:])})
                t_b += evs["start"].elapsed_time(evs["loss"])
            serial_ms_step = t_b / n_groups
            launches = (launches_per_step or 0) * steps
        else:
            torch.cuda._sleep(int(30e6))      # let the host run ahead of the GPU
            eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eb0.record()
            raw = [self.step(timers=True).events for _ in range(steps)]
            eb1.record()
            self.fence()
            launches = lib.fmlp_launch_count() - launches0
            serial_ms_step = eb0.elapsed_time(eb1) / steps
            results = [{q: e[p].elapsed_time(e[q]) for p, q in zip(order[:-1], order[1:])} for e in raw]

        in_round = {k: statistics.median(r[k] for r in results) for k in order[1:]}
        # ---- timed region C (roofline): every streaming stage's launches back to back on the main stream — a
        #      CUDA graph holding that stage alone is replayed K times between two events (eager launches when
        #      graphs are off), so the launch latency of one replay hides behind the previous one exactly as
        #      between the kernels of the real round.  Region B brackets every stage with event nodes, which puts
        #      each stage's launch latency (~6-9 us) inside its own interval; both are reported.  Inputs of every
        #      stage (>= 225 MB) exceed the 126 MB L2, so no flush is needed between the launches.
        kms = dict(in_round)
        loop_note = "not measured"
        reps = max(20, min(steps, 200))
        for k in ("sim", "proto", "fedavg"):
            try:
                g3 = None
                if graph is not None:
                    g3 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g3):
                        self.step(only=k)
                run_once = (g3.replay if g3 is not None else (lambda k=k: self.step(only=k)))
                for _ in range(3):
                    run_once()
                self.fence()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                for _ in range(reps):
                    run_once()
                c1.record()
                self.fence()
                tk = torch.tensor([c0.elapsed_time(c1) / reps], dtype=torch.float64, device=dev)
                if world > 1:
                    self.dist.all_reduce(tk, op=self.dist.ReduceOp.MAX)
                kms[k] = float(tk.item())
                loop_note = (f"{reps} back-to-back " + ("replays of a CUDA graph holding the stage alone" if g3 is not None else "eager launches")
                             + " between two CUDA events, per stage")
            except Exception as exc:
                loop_note = f"stage loop failed ({type(exc).__name__}: {exc})"[:160] + "; in-round event brackets used"
                torch.cuda.synchronize()
        ab = alg_bytes(w, inp)
        peak, peak_src = peak_hbm()
        kernels = {}
        for k in ("sim", "proto", "fedavg", "loss", "select_fill"):
            gbs = ab[k] / (kms[k] * 1e-3) / 1e9
            kernels[k] = {"ms": round(kms[k], 5), "alg_bytes": ab[k], "gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4),
                          "ms_in_round_between_event_nodes": round(in_round[k], 5)}
        kernels["fedavg"]["includes"] = "FedAvg + FedAvg_proto + tail pack + finalize launches" if world == 1 else self.collective
        if world > 1:
            link_bytes = 2 * (world - 1) / world * 4 * inp["Ppad"]
            kernels["fedavg"].pop("gbs", None); kernels["fedavg"].pop("frac_of_hbm_peak", None)
            kernels["fedavg"]["nvlink_floor_ms_two_shot_770gbs"] = round(link_bytes / 770e9 * 1e3, 5)
            kernels["fedavg"]["nvlink_floor_ms_nvls_770gbs"] = round(4 * inp["Ppad"] / 770e9 * 1e3, 5)
        dom = max(("sim", "proto", "fedavg") if world == 1 else ("sim", "proto"), key=lambda k: kms[k])
        dom_names = {"sim": "tag_sim_kernel", "proto": "proto_accum_kernel", "fedavg": "fedavg_flat_kernel"}
        roofline = {"kernel": dom_names[dom], "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(kernels[dom]["gbs"] / peak, 4), "traffic": traffic_from_profiles(dom_names[dom]) if w.name == "ich55k" else None,
                    "peak_source": peak_src, "alg_bytes_per_launch": ab[dom], "avg_launch_ms": round(kms[dom], 5)}
        stream_bytes = ab["sim"] + ab["proto"] + ab["fedavg"] + ab["loss"] + ab["select_fill"]
        return dict(ms_per_step=ms_step, value=inp["N"] * world / (ms_step * 1e-3), kernels=kernels, roofline=roofline,
                    schedule=self.schedule_note,
                    launches=int(launches), launches_per_step=launches_per_step, graph_note=graph_note,
                    stage_note=f"ms: {loop_note}; ms_in_round_between_event_nodes: {stage_note}",
                    serial_ms_per_step=serial_ms_step, step_alg_bytes=stream_bytes,
                    step_frac_of_hbm_peak=round(stream_bytes / (ms_step * 1e-3) / 1e9 / peak, 4), graph=graph)

    def loss_sweep(self, rows_list=(32, 1024, 8192, 85000)):
        """configs[3]: stage-2 (BCE + teacher consistency) and stage-1 loss fwd+bwd over `rows` logits rows."""
        import torch
        import fedmlp_b200 as F
        C = self.w.C
        out = {}
        g = torch.Generator(device=self.dev).manual_seed(5)
        for rows in rows_list:
            z = [torch.randn(rows, C, generator=g, device=self.dev) * 2 for _ in range(4)]
            y = (torch.rand(rows, C, generator=g, device=self.dev) < 0.1).float()
            d = (torch.rand(rows, C, generator=g, device=self.dev) < 0.5).float()
            active, missing = [0], list(range(1, C))

            def s2():
                zz = z[0].detach().requires_grad_(True)
                F.fedmlp_stage2_loss(zz, z[1], y, d, variant="sup_dis").backward()

            def s1():
                z1, z2 = z[0].detach().requires_grad_(True), z[1].detach().requires_grad_(True)
                F.fedmlp_stage1_loss(z1, z2, z[2], z[3], y, active, missing, 32).backward()

            res = {}
            for name, fn in (("stage2_sup_dis_us", s2), ("stage1_us", s1)):
                try:
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        fn()
                    e1.record(); torch.cuda.synchronize()
                    res[name] = round(e0.elapsed_time(e1) / 20 * 1e3, 2)
                except Exception as exc:
                    res[name] = f"{type(exc).__name__}"[:60]
            out[str(rows)] = res
        out["note"] = "autograd.Function wrappers (fwd + bwd through the C ABI, incl. torch dispatch), eager"
        return out


def gpu_arm(a):
    import torch
    import torch.distributed as dist

    from fedmlp_b200 import _cabi as cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = bind_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.load()
    w = Workload(a.config, a)
    run = Runner(a, w, world, rank, dev, lib)
    inp = run.inp

    parity = None
    rc = 0
    if world > 1:
        parity = run.parity()
        if not parity["parity_ok"]:
            rc = 3
    sampler = ClockSampler(local) if rank == 0 else None
    m = run.measure(a.steps, a.warmup)

    # ---- end to end: pinned host inputs -> H2D -> round -> D2H of every result, per step
    e2e = None
    if not a.skip_e2e:
        try:
            e2e = run_e2e(a, run, m.get("graph"), numa)
        except Exception as exc:      # report instead of losing the whole line (same exception on every rank)
            e2e = {"value": None, "unit": UNIT, "error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    api = None
    if world == 1 and not a.skip_e2e:
        try:
            api = measure_api(w, dev)
        except Exception as exc:
            api = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.synchronize()

    # ---- the other BASELINE configs at this N (C = 14 shapes), fewer steps, no e2e / CPU leg
    extra = []
    names = [] if a.extra_configs == "none" else ([n for n in ("cxr14_64c", "effb0_85k") if n != a.config]
                                                  if a.extra_configs == "auto" else [n for n in a.extra_configs.split(",") if n])
    global headline_collective
    headline_collective = run.collective
    del inp
    run.inp = None
    m.pop("graph", None)
    headline = m
    for name in names:
        try:
            del run
            torch.cuda.empty_cache()
            w2 = Workload(name)
            run = Runner(a, w2, world, rank, dev, lib)
            entry = {"name": name, **w2.config_keys(run.inp["P"])}
            if world > 1:
                entry["parity"] = run.parity()
                if not entry["parity"]["parity_ok"]:
                    rc = 3
            m2 = run.measure(min(a.steps, 200), min(a.warmup, 5))
            m2.pop("graph", None)
            entry.update(ms_per_step=m2["ms_per_step"], value=m2["value"], unit=UNIT, kernels=m2["kernels"], roofline=m2["roofline"],
                         serial_ms_per_step=m2["serial_ms_per_step"], step_frac_of_hbm_peak=m2["step_frac_of_hbm_peak"],
                         gpu_launches_per_step=m2["launches_per_step"], timing=m2["graph_note"], collective=run.collective,
                         schedule=m2.get("schedule"))
            if name == "effb0_85k" and world == 1:
                entry["loss_sweep"] = run.loss_sweep()
            extra.append(entry)
        except Exception as exc:
            extra.append({"name": name, "error": f"{type(exc).__name__}: {exc}"[:300]})
            torch.cuda.synchronize()

    if rank == 0:
        cpu = None
        if world == 1 and not a.skip_cpu_baseline:
            r = run_cpu_arm(w, steps=5, warmup=1, n_clients=a.cpu_clients, budget_s=30.0)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                   "ms_per_step": r["t_step"] * 1e3, "fedavg_gbs": (w.S + 1) * 4 * r["P"] / r["t_fedavg"] / 1e9}
        cfg = w.full_config(a, world)      # identical to the reference arm's `config` for the same workload, N and flags
        line = {
            "metric": METRIC, "value": headline["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": headline["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "collective": headline_collective,
            "fedavg_gbs": headline["kernels"]["fedavg"].get("gbs"),
            "roofline": headline["roofline"], "kernels": headline["kernels"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": headline["launches"], "gpu_launches_per_step": headline["launches_per_step"], "clocks": clocks,
            "step_frac_of_hbm_peak": headline["step_frac_of_hbm_peak"],
            "timing": {"value": headline["graph_note"], "kernels": headline["stage_note"],
                       "serial_ms_per_step": headline["serial_ms_per_step"], "schedule": headline.get("schedule")},
            "api": api,
            "parity": parity, "parity_ok": (parity or {}).get("parity_ok") if world > 1 else None,
            "numa": numa, "configs": extra,
        }
        print_result(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


headline_collective = None


def measure_api(w: Workload, dev, reps=10):
    """The reference-shaped entry points as a caller uses them (wall clock around the call + a device sync, median
    of `reps`): FedAvg(w_locals, dict_len) (utils/FedAvg.py:7-14) on K real-layout state_dicts — 727 separately
    allocated CUDA tensors each (what main.py:196 collects), FlatStateDicts (fedmlp_b200.flat), and CPU tensors
    (the reference's stage 2 returns net.cpu() weights, local_training.py:1251) — and one tagging pass of a client
    through TagBatch.step (similarity + selection, :1052-1112)."""
    import torch
    import fedmlp_b200 as F
    from fedmlp_b200.shapes import synth_state_dict

    shapes = w.state_shapes()
    K = 8
    base = synth_state_dict(shapes, 1037)
    cpu_sds = [synth_state_dict(shapes, 1038 + k, base=base, counter=100 + k) for k in range(K)]
    gpu_sds = [type(sd)((k, v.to(dev)) for k, v in sd.items()) for sd in cpu_sds]
    flat_sds = [F.FlatStateDict.from_state_dict(sd) for sd in gpu_sds]
    dict_len = [w.n] * K

    def wall(fn):
        ts = []
        for _ in range(reps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        return round(statistics.median(ts[2:]), 4)

    out = {"fedavg_state_dicts_ms": wall(lambda: F.FedAvg(gpu_sds, dict_len)),
           "fedavg_flat_state_dicts_ms": wall(lambda: F.FedAvg(flat_sds, dict_len)),
           "fedavg_cpu_state_dicts_ms": wall(lambda: F.FedAvg(cpu_sds, dict_len)),
           "fedavg_flat_buffers_ms": wall(lambda: F.fedavg_flat_buffers([f.flat_f32 for f in flat_sds], dict_len)),
           "tensors_per_state_dict": len(shapes), "clients": K}
    n, C, D = w.n, w.C, w.D
    g = torch.Generator(device=dev).manual_seed(3)
    feat = torch.relu(torch.randn(n, D, generator=g, device=dev))
    proto = torch.relu(torch.randn(2 * C, D, generator=g, device=dev)) + 0.05
    tb = F.TagBatch([0, n], C, [[0]], [list(range(1, C))], device=dev)
    out["train_FedMLP_tag_ms"] = wall(lambda: tb.step(feat, proto, 0.005, 0.01, mode="folded"))
    out["note"] = ("wall clock incl. Python + a device sync; scattered dicts: one pointer-table pass over 727 x K tensors per call "
                   "(table re-uploaded only when a pointer changed) + a 727-entry result dict; flat state_dicts: K pointers + the result dict; flat buffers: K pointers, one [P] tensor out; CPU dicts: pinned staging (pooled) + H2D + D2H of the result")
    return out


def run_e2e(a, run, graph, numa):
    """Same round through host buffers: every step copies ALL its inputs from pinned host memory and reads every
    result back to pinned host memory.  The three legs run on three streams and are double-buffered across steps
    (H2D of step i+1 overlaps compute + D2H of step i); the host consumes the results of step i-1 while step i
    is in flight.  PCIe-bound.  The pinned buffers are allocated after the rank was bound to its GPU's NUMA node."""
    import torch
    dist = run.dist
    inp, dev, world = run.inp, run.dev, run.world

    names = ["feat_tag", "logits", "logits_glob", "labels", "feat_proto", "logits_proto", "proto"]
    host = {k: inp[k].cpu().pin_memory() for k in names}
    host_flats = [f.cpu().pin_memory() for f in inp["flats"]]
    host_cnt = [c.cpu().pin_memory() for c in inp["counters"]]
    h2d = (sum(v.numel() * v.element_size() for v in host.values()) + sum(f.numel() * 4 for f in host_flats)
           + sum(c.numel() * 8 for c in host_cnt))
    sets = [inp, dict(inp)]
    for k in names:
        sets[1][k] = torch.empty_like(inp[k])
    sets[1]["flats"] = [torch.empty_like(f) for f in inp["flats"]]
    sets[1]["counters"] = [torch.empty_like(c) for c in inp["counters"]]
    main_s = torch.cuda.current_stream(dev)
    copy_s, d2h_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]      # inputs of set b landed
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]    # compute finished reading set b
    ev_out = [torch.cuda.Event(), torch.cuda.Event()]     # results of the step using set b are on the host
    for e in ev_free + ev_out:
        e.record(main_s)
    out_host = [{}, {}]
    d2h_bytes = [0]

    def e2e_step(i):
        b = i & 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_free[b])
            for k in names:
                sets[b][k].copy_(host[k], non_blocking=True)
            for d, h in zip(sets[b]["flats"], host_flats):
                d.copy_(h, non_blocking=True)
            for d, h in zip(sets[b]["counters"], host_cnt):
                d.copy_(h, non_blocking=True)
            ev_in[b].record(copy_s)
        main_s.wait_event(ev_in[b])
        main_s.wait_event(ev_out[b ^ 1])          # the previous step's results have left the shared output buffers
        r = run.step(data=sets[b])
        ev_free[b].record(main_s)
        outs = {"counts": r.counts, "sel": r.sel, "losses": r.losses, "dz": r.dz, "proto": r.protos.proto,
                "cnt": r.protos.cnt, "tcnt": r.protos.tcnt, "global": r.global_flat}
        for k in ("proto_glob", "tao", "counters"):
            if getattr(r, k) is not None:
                outs[k] = getattr(r, k)
        with torch.cuda.stream(d2h_s):
            d2h_s.wait_event(ev_free[b])
            n = 0
            for k, v in outs.items():
                if k not in out_host[b] or out_host[b][k].shape != v.shape:
                    out_host[b][k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                out_host[b][k].copy_(v, non_blocking=True)
                n += v.numel() * v.element_size()
            ev_out[b].record(d2h_s)
        d2h_bytes[0] = n
        ev_out[b ^ 1].synchronize()               # the host consumes the results of step i-1 here

    steps = a.e2e_steps or min(a.steps, 20)
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_s)
    for i in range(steps):
        e2e_step(i)
    torch.cuda.synchronize()                      # the last step's results are on the host too
    e1.record(main_s)
    torch.cuda.synchronize()
    ms_local = e0.elapsed_time(e1)
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    return {"value": inp["N"] * world / (ms_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h_bytes[0]), "ms_per_step": ms_step, "steps": steps,
            "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / steps,
            "h2d_gbs_this_rank": round(h2d / (ms_local / steps * 1e-3) / 1e9, 1), "numa_node": numa.get("node"),
            "note": "H2D / compute / D2H on three streams, double-buffered across steps; every input (features, logits, labels, "
                    "prototypes, client parameters, counters) is re-uploaded every step; PCIe-bound"}


def main():
    a = parse_args()
    # Keep stdout clean for the ONE JSON line: libraries (NCCL banner, torchrun notices) write to
    # fd 1, so fd 1 is pointed at stderr for the duration of the run and restored for the result.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    captured = []
    global print_result
    print_result = captured.append
    rc = 0
    try:
        if a.impl == "reference":
            reference_arm(a)
        else:
            rc = gpu_arm(a) or 0
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in captured:
        print(line, flush=True)
    sys.exit(rc)


print_result = print


if __name__ == "__main__":
    main()
