#!/usr/bin/env python
"""bench.py — FedMLP round hot path (tag + loss + prototypes + FedAvg) on B200.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

A "step" is one pass of the per-round hot path over the clients resident on one GPU:
  tag (similarity + selection + mask fill) over N_total x D features of the incoming global model,
  stage-2 loss forward+backward over every client's rows, prototype/t construction over the
  N_total x D features of the locally trained model, FedAvg over the K client parameter buffers.
Workload at N=1 = BASELINE.json configs[1]: ICH 5-class DenseNet121, 8 clients on one B200,
55,000 synthetic feature rows (6,875 per client), D=1024, P=7,042,629 fp32 parameters per client.
With --gpus N>1 every rank holds its own 8 clients (weak scaling) and only FedAvg crosses GPUs
(weighted partial sums + one NCCL all-reduce of the flat parameter buffer).

Prints ONE JSON line (contract in the task statement): value = client-samples/s with inputs
resident in HBM; e2e = the same through host buffers (pinned H2D of every input and D2H of every
result inside the timed region); roofline = dominant kernel vs the measured HBM peak;
cpu_baseline = the oracle port of the reference's CPU path timed on this box's cores.
`--impl reference` times that CPU path alone (the reference is pure Python and cannot travel
to the GPU box, so the arm is the oracle port — kind "port").
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clients-per-gpu", type=int, default=8)
    ap.add_argument("--rows-per-client", type=int, default=6875)
    ap.add_argument("--classes", type=int, default=5)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--sim-mode", default="folded", choices=["pair", "folded"])
    ap.add_argument("--cpu-clients", type=int, default=2, help="clients in the bounded CPU-baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--graph-multi", type=int, default=0,
                    help="N>1 with the fused collective: 1 = replay the round (incl. the fold+all-reduce kernel) as a CUDA graph")
    ap.add_argument("--proto-on-side", type=int, default=1,
                    help="N>1: 1 = prototypes share the side stream with the aggregation, 0 = stay on the main chain")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused fold+all-reduce kernel over peer memory, or local fold + NCCL all-reduce")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------- workload
def workload_name(a):
    return (f"ICH {a.classes}-class DenseNet121, {a.clients_per_gpu} clients/GPU x {a.rows_per_client} rows, "
            f"D={a.dim}: tag+loss+prototypes+FedAvg per round")


def make_device_inputs(a, rank, dev):
    """Seeded synthetic inputs of the named shapes, created on the device (setup, untimed)."""
    import torch
    from fedmlp_b200.shapes import count_params, densenet121_state_shapes

    S, n, C, D = a.clients_per_gpu, a.rows_per_client, a.classes, a.dim
    N = S * n
    g = torch.Generator(device=dev).manual_seed(1037 + rank)
    prev = torch.tensor(ICH_PREVALENCE if C == 5 else torch.linspace(0.02, 0.20, C).tolist(), device=dev)
    labels = (torch.rand(N, C, generator=g, device=dev) < prev).float()
    shift = torch.randn(C, D, generator=g, device=dev) * 0.25
    feat_tag = torch.relu(torch.randn(N, D, generator=g, device=dev)) + torch.relu(labels @ shift)
    feat_proto = torch.relu(torch.randn(N, D, generator=g, device=dev)) + torch.relu(labels @ shift)
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
    shapes = densenet121_state_shapes(C)
    P, J = count_params(shapes)
    Ppad = sum((int(torch.Size(s).numel()) + 3) // 4 * 4 for s, d in shapes.values() if d == torch.float32)
    base = torch.randn(Ppad, generator=g, device=dev) * 0.05
    flats = [(base + 0.02 * torch.randn(Ppad, generator=g, device=dev)).contiguous() for _ in range(S)]
    weights = [n] * S
    return dict(N=N, P=P, J=J, Ppad=Ppad, labels=labels, feat_tag=feat_tag, feat_proto=feat_proto, logits=logits,
                logits_glob=logits_glob, logits_proto=logits_proto, proto=proto, flats=flats, weights=weights)


def alg_bytes(a, inp):
    """Algorithmic bytes per launch."""
    N, D, C, K, P = inp["N"], a.dim, a.classes, a.clients_per_gpu, inp["Ppad"]
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


# ----------------------------------------------------------------------------------- CPU path (oracle port)
def cpu_round_sample(a, n_clients, seed=1037):
    """Build the bounded CPU sample: `n_clients` clients of the workload + K state_dicts."""
    import torch
    from oracle import fedmlp_oracle as O
    from fedmlp_b200.shapes import densenet121_state_shapes, synth_state_dict

    C, D, n = a.classes, a.dim, a.rows_per_client
    clients = []
    for k in range(n_clients):
        feat, labels, logits = O.synth_client(n, D, C, seed=seed + k)
        feat2, _, logits2 = O.synth_client(n, D, C, seed=seed + 100 + k)
        clients.append(dict(feat=feat, labels=labels, logits=logits, feat2=feat2, logits2=logits2,
                            zg=torch.randn(n, C, generator=torch.Generator().manual_seed(seed + 200 + k)) * 2,
                            active=[k % C], missing=[c for c in range(C) if c != k % C],
                            state=O.TaggingState(list(range(n)), [c for c in range(C) if c != k % C])))
    proto = O.synth_prototypes(clients[0]["feat"], clients[0]["labels"])
    shapes = densenet121_state_shapes(C)
    base = synth_state_dict(shapes, seed)
    sds = [synth_state_dict(shapes, seed + 1 + k, base=base, counter=100 + k) for k in range(a.clients_per_gpu)]
    return dict(clients=clients, proto=proto, sds=sds, weights=[n] * a.clients_per_gpu)


def cpu_round_step(a, sample):
    """One pass of the reference's CPU path (oracle port) over the sample; returns (t_clients, t_fedavg)."""
    import torch
    from oracle import fedmlp_oracle as O

    t0 = time.perf_counter()
    for cl in sample["clients"]:
        # tag: similarity + python-sorted selection (utils/utils.py:24-35) + label/mask fill
        cl["state"].traindata_idx = [[] for _ in cl["state"].traindata_idx]
        cl["state"].step(cl["feat"], sample["proto"], 0.005, 0.01, python_sort=True)
        tgt, dis = O.mask_fill(cl["labels"].numpy(), cl["state"].dataset_idx, cl["active"], cl["missing"],
                               cl["state"].traindata_idx)
        # stage-2 loss forward + backward over the client's rows
        O.loss_and_grads(lambda z, zg, y, d: O.stage2_loss(z, zg, y, d), cl["logits"], cl["zg"],
                         torch.from_numpy(tgt), torch.from_numpy(dis), n_grad=1)
        # prototypes + t
        O.prototype_build(cl["feat2"], cl["labels"], cl["logits2"], cl["active"], cl["missing"], 0.3, 0.7, True)
    t1 = time.perf_counter()
    O.fedavg(sample["sds"], sample["weights"])
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def run_cpu_arm(a, steps, warmup):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = cpu_round_sample(a, a.cpu_clients)
    for _ in range(warmup):
        cpu_round_step(a, sample)
    tc, tf = [], []
    for _ in range(steps):
        c, f = cpu_round_step(a, sample)
        tc.append(c); tf.append(f)
    scale = a.clients_per_gpu / a.cpu_clients
    t_step = statistics.mean(tc) * scale + statistics.mean(tf)
    n_total = a.clients_per_gpu * a.rows_per_client
    P = 7042629 if a.classes == 5 else None
    return dict(value=n_total / t_step, t_step=t_step, t_clients=statistics.mean(tc), t_fedavg=statistics.mean(tf),
                cores=cores, P=P,
                sample=(f"{a.cpu_clients} of {a.clients_per_gpu} clients ({a.cpu_clients * a.rows_per_client} rows) through "
                        f"tag+mask-fill+loss+prototypes (time x{scale:g}) + FedAvg over all {a.clients_per_gpu} "
                        f"DenseNet121 state_dicts (727 tensors), {steps} steps after {warmup} warm-up"))


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    # keep the whole run within a few minutes whatever K/W the driver passes
    budget_steps = 40
    if steps + warmup > budget_steps:
        warmup = min(warmup, 5)
        steps_run = budget_steps - warmup
    else:
        steps_run = steps
    r = run_cpu_arm(a, steps_run, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["t_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "steps_executed": steps_run},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "fedavg_gbs": ((a.clients_per_gpu + 1) * 4 * r["P"] / r["t_fedavg"] / 1e9) if r["P"] else None,
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


def gpu_arm(a):
    import torch
    import torch.distributed as dist

    from fedmlp_b200 import _cabi as cabi
    from fedmlp_b200.round import ClientShard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.load()

    inp = inp0 = make_device_inputs(a, rank, dev)
    S, C = a.clients_per_gpu, a.classes
    shard = ClientShard([a.rows_per_client] * S, C, [[(rank * S + k) % C] for k in range(S)], device=dev,
                        sim_mode=a.sim_mode)
    shard.keep_history = False      # host-list bookkeeping (two small device copies per round) is off the hot path
    fed_out = torch.empty(inp["Ppad"], dtype=torch.float32, device=dev)
    if world > 1:
        total_w = float(sum(inp["weights"]) * world)
        w_norm = [w / total_w for w in inp["weights"]]          # pre-normalised: all-reduce yields the mean

    side_stream = torch.cuda.Stream(device=dev)
    fused, collective = None, "none"
    if world > 1:
        collective = "nccl all_reduce after the local fold"
        if a.collective == "fused":
            try:
                from fedmlp_b200.dist import FusedFedAvgAllReduce
                fused = FusedFedAvgAllReduce(inp["Ppad"], device=dev)
                collective = (f"fused fold + chunk-pipelined two-shot all-reduce over NVLink peer memory "
                              f"(one kernel per rank, {fused.n_chunks} chunks)")
            except Exception as exc:
                fused = None
                collective += f" (fused path unavailable: {type(exc).__name__}: {exc})"[:200]

    def step(timers=None, overlap=True, data=None):
        """One round hot path.  overlap: {prototypes -> FedAvg (-> all-reduce)} on a side stream,
        concurrent with {sim -> select -> fill -> loss}; the per-stage event timing (timers) runs
        the stages back to back on one stream so every kernel is timed alone."""
        side = side_stream if (overlap and timers is None) else None
        inp = data if data is not None else inp0
        if world == 1:
            return shard.round_hot_path(inp["feat_tag"], inp["proto"], inp["logits"], inp["logits_glob"], inp["labels"],
                                        inp["feat_proto"], inp["logits_proto"], inp["flats"], inp["weights"],
                                        timers=timers, fedavg_out=fed_out, side_stream=side)
        if fused is not None:
            return shard.round_hot_path(inp["feat_tag"], inp["proto"], inp["logits"], inp["logits_glob"], inp["labels"],
                                        inp["feat_proto"], inp["logits_proto"], inp["flats"], w_norm, timers=timers,
                                        fedavg_out=fed_out, divide=False, side_stream=side, proto_on_side=a.proto_on_side,
                                        aggregate_fn=lambda bufs, w: fused(bufs, w))
        return shard.round_hot_path(inp["feat_tag"], inp["proto"], inp["logits"], inp["logits_glob"], inp["labels"],
                                    inp["feat_proto"], inp["logits_proto"], inp["flats"], w_norm, timers=timers,
                                    fedavg_out=fed_out, divide=False, side_stream=side, proto_on_side=a.proto_on_side,
                                    after_aggregate=lambda g: dist.all_reduce(g))

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        step()
    fence()
    # ---- timed region A (value): K steps, each ONE CUDA-graph replay of the round (single GPU;
    #      the launch-bound inner loop is captured once, as the task's design rules ask)
    graph, graph_note = None, "eager launches"
    launches_per_step = None
    # N>1 stays eager: a graph serialises the cooperative fold+all-reduce kernel against the other
    # stream (measured slower) and NCCL capture is not used
    if (world == 1 or (a.graph_multi and fused is not None)) and not a.no_graph:
        try:
            l0 = lib.fmlp_launch_count()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                step()
            launches_per_step = lib.fmlp_launch_count() - l0
            graph, graph_note = g_, ("CUDA-graph replay of the 7-launch round, two-stream DAG {sim,select,fill,loss} || {proto,FedAvg"
                                     + ("+all-reduce}" if world > 1 else "}"))
            for _ in range(3):
                graph.replay()
        except Exception as exc:          # fall back to eager timing, say so
            graph, graph_note = None, f"eager launches (graph capture failed: {type(exc).__name__})"
            torch.cuda.synchronize()
    fence()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if graph is not None:
        for _ in range(a.steps):
            graph.replay()
    else:
        for _ in range(a.steps):
            step()
    ev1.record()
    fence()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps

    # ---- timed region B (kernels / roofline): the same K steps on ONE stream with a CUDA event
    #      after every stage.  N=1: the serial round is captured in a second graph with external
    #      event-record nodes, so the events bracket the kernels without eager launch gaps; each
    #      replay is followed by a sync to read its events.  Otherwise: eager launches + events.
    launches0 = lib.fmlp_launch_count()
    results = []          # only the per-stage events / times are kept
    stage_note = "eager launches, CUDA event after every stage"
    stage_graph = None
    if graph is not None:
        try:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                evs = step(timers="external").events
            stage_graph, stage_note = g2, "single-stream CUDA graph with an external CUDA event after every stage"
        except Exception as exc:
            stage_graph = None
            stage_note += f" (graph with external events unavailable: {type(exc).__name__})"
            torch.cuda.synchronize()
    order = ["start", "sim", "select_fill", "loss", "proto", "fedavg"]
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if stage_graph is not None:
        stage_graph.replay(); torch.cuda.synchronize()
        t_b = 0.0
        for _ in range(a.steps):
            stage_graph.replay()
            torch.cuda.synchronize()
            results.append({q: evs[p].elapsed_time(evs[q]) for p, q in zip(order[:-1], order[1:])})
            t_b += evs["start"].elapsed_time(evs["fedavg"])
        eager_ms_step = t_b / a.steps
        launches = (launches_per_step or 7) * a.steps
    else:
        torch.cuda._sleep(int(30e6))      # let the host run ahead of the GPU
        eb0.record()
        raw = []
        for _ in range(a.steps):
            raw.append(step(timers=True).events)
        eb1.record()
        fence()
        launches = lib.fmlp_launch_count() - launches0
        eager_ms_step = eb0.elapsed_time(eb1) / a.steps
        results = [{q: e[p].elapsed_time(e[q]) for p, q in zip(order[:-1], order[1:])} for e in raw]

    # per-kernel device times from the events recorded inside the timed steps
    kms = {k: statistics.mean(r[k] for r in results) for k in order[1:]}
    ab = alg_bytes(a, inp)
    peak, peak_src = peak_hbm()
    kernels = {}
    for k in ("sim", "proto", "fedavg", "loss", "select_fill"):
        gbs = ab[k] / (kms[k] * 1e-3) / 1e9
        kernels[k] = {"ms": round(kms[k], 5), "alg_bytes": ab[k], "gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)}
    if world > 1:
        kernels["fedavg"]["note"] = collective
        kernels["fedavg"].pop("gbs", None); kernels["fedavg"].pop("frac_of_hbm_peak", None)
        kernels["nccl_allreduce_alone"] = measure_allreduce(fed_out, world)
    dom = max(("sim", "proto", "fedavg") if world == 1 else ("sim", "proto"), key=lambda k: kms[k])
    dom_names = {"sim": "tag_sim_kernel", "proto": "proto_accum_kernel", "fedavg": "fedavg_flat_kernel"}
    roofline = {"kernel": dom_names[dom], "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": round(kernels[dom]["gbs"] / peak, 4), "traffic": traffic_from_profiles(dom_names[dom]),
                "peak_source": peak_src, "alg_bytes_per_launch": ab[dom], "avg_launch_ms": round(kms[dom], 5)}
    n_total_all = inp["N"] * world
    value = n_total_all / (ms_step * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> round -> D2H of every result, per step
    e2e = None
    if not a.skip_e2e:
        try:
            e2e = run_e2e(a, inp, shard, step, fed_out, world, dev)
        except Exception as exc:      # report instead of losing the whole line (same exception on every rank)
            e2e = {"value": None, "unit": UNIT, "error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not a.skip_cpu_baseline:
            r = run_cpu_arm(a, steps=3, warmup=1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                   "ms_per_step": r["t_step"] * 1e3, "fedavg_gbs": (S + 1) * 4 * inp["P"] / r["t_fedavg"] / 1e9}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "clients_per_gpu": S, "rows_per_client": a.rows_per_client,
                       "feature_dim": a.dim, "classes": C, "params_per_client": inp["P"], "sim_mode": a.sim_mode,
                       "l2": f"inputs larger than L2: {round((ab['sim'] + ab['proto'] + ab['fedavg']) / 1e6)} MB streamed per step vs 126 MB L2, no flush needed",
                       "parallelism": (f"clients sharded over {world} GPU(s); FedAvg = local weighted partial + NCCL all-reduce, "
                                       "on a side stream with the prototype pass, concurrent with the tagging/loss chain")
                       if world > 1 else "single GPU", "collective": collective},
            "fedavg_gbs": kernels["fedavg"].get("gbs"),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / a.steps, "clocks": clocks,
            "timing": {"value": graph_note, "kernels": stage_note, "serial_ms_per_step": eager_ms_step},
        }
        print_result(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_allreduce(buf, world, iters=20):
    """Stand-alone NCCL all-reduce of the flat parameter buffer (bus bandwidth vs NVLink)."""
    import torch
    import torch.distributed as dist

    for _ in range(3):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = buf.numel() * 4
    return {"ms": round(ms, 5), "bytes": nbytes, "bus_gbs": round(2 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9, 1),
            "nvlink_peak_gbs_per_direction": 900}


def run_e2e(a, inp, shard, step_fn, fed_out, world, dev):
    """Same round through host buffers: every step copies ALL its inputs from pinned host memory
    (680 MB) and reads every result back to pinned host memory (30 MB).  The three legs run on three
    streams and are double-buffered across steps (H2D of step i+1 overlaps compute + D2H of step i);
    the host consumes the results of step i-1 while step i is in flight.  PCIe-bound."""
    import torch
    import torch.distributed as dist

    names = ["feat_tag", "logits", "logits_glob", "labels", "feat_proto", "logits_proto", "proto"]
    host = {k: inp[k].cpu().pin_memory() for k in names}
    host_flats = [f.cpu().pin_memory() for f in inp["flats"]]
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + sum(f.numel() * 4 for f in host_flats)
    sets = [inp, dict(inp)]
    for k in names:
        sets[1][k] = torch.empty_like(inp[k])
    sets[1]["flats"] = [torch.empty_like(f) for f in inp["flats"]]
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
            ev_in[b].record(copy_s)
        main_s.wait_event(ev_in[b])
        main_s.wait_event(ev_out[b ^ 1])          # the previous step's results have left the shared output buffers
        r = step_fn(data=sets[b])
        ev_free[b].record(main_s)
        outs = {"counts": r.counts, "sel": r.sel, "losses": r.losses, "dz": r.dz, "proto": r.protos.proto,
                "cnt": r.protos.cnt, "tcnt": r.protos.tcnt, "global": r.global_flat}
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
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    return {"value": inp["N"] * world / (ms_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h_bytes[0]), "ms_per_step": ms_step, "steps": steps,
            "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / steps,
            "note": "H2D / compute / D2H on three streams, double-buffered across steps; PCIe-bound"}


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
    try:
        if a.impl == "reference":
            reference_arm(a)
        else:
            gpu_arm(a)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in captured:
        print(line, flush=True)


print_result = print


if __name__ == "__main__":
    main()
