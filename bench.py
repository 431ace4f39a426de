#!/usr/bin/env python
"""bench.py - samples/sec of one DeepFM training step (fwd + loss + bwd + update) on synthetic
Criteo-shaped batches (BASELINE.json configs[1]: 26 tables x 1M rows, emb_dim 32, batch 65536).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c2small]

Prints ONE JSON line (see the task contract): `value` = device-timed whole-job samples/s with the
batch already resident in HBM; `e2e` = the same step through the public API
(`Model.train_on_batch(host arrays)`) including pinned-H2D of the inputs and the D2H read of the loss;
`roofline` = the dominant kernel against the measured peaks in MEASURED_PEAKS.json; `cpu_baseline` =
the CPU oracle (torch-CPU restatement of the reference math) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1]
    "c2": dict(workload="DeepFM synthetic Criteo: 26 tables x 1M rows, emb_dim=32, batch=65536, 13 dense",
               n_sparse=26, n_dense=13, vocab=1000000, dim=32, batch=65536, hidden=(256, 128, 64)),
    # tiny variant for CPU smoke runs of this script
    "c2small": dict(workload="DeepFM synthetic Criteo (small): 26 tables x 10k rows, emb_dim=32, batch=4096",
                    n_sparse=26, n_dense=13, vocab=10000, dim=32, batch=4096, hidden=(256, 128, 64)),
}
LR = 0.01
N_BATCHES = 4      # distinct pre-generated batches cycled through the timed steps


def feature_columns(cfg):
    from deepctr_b200.feature_column import SparseFeat, DenseFeat
    cols = [SparseFeat("C%d" % (i + 1), cfg["vocab"], cfg["dim"]) for i in range(cfg["n_sparse"])]
    cols += [DenseFeat("I%d" % (i + 1), 1) for i in range(cfg["n_dense"])]
    return cols


def synth_batches(cfg, n, rank=0):
    """uniform ids (worst case for the gather: no reuse), U(0,1) dense, Bernoulli(0.25) labels; seed 2020."""
    rng = np.random.RandomState(2020 + rank)
    out = []
    for _ in range(n):
        ids = rng.randint(0, cfg["vocab"], size=(cfg["batch"], cfg["n_sparse"])).astype(np.int32)
        dense = rng.rand(cfg["batch"], cfg["n_dense"]).astype(np.float32)
        y = (rng.rand(cfg["batch"]) < 0.25).astype(np.float32)
        out.append((ids, dense, y))
    return out


def as_inputs(cfg, ids, dense):
    """What a user holds: one contiguous host array per feature (e.g. DataFrame columns)."""
    x = {"C%d" % (i + 1): np.ascontiguousarray(ids[:, i]) for i in range(cfg["n_sparse"])}
    x.update({"I%d" % (i + 1): np.ascontiguousarray(dense[:, i]) for i in range(cfg["n_dense"])})
    return x


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def gemm_launch_times(cfg, precision, dev, reps=5):
    """CUDA-event duration of every DNN GEMM launch of one step (forward / dgrad / wgrad per layer, operands
    as ops.dense passes them: pre-split planes in bf16x3 mode), each timed alone after an L2 flush.
    Returns [(label, m, n, k, us)]."""
    import torch
    from deepctr_b200 import _lib as L, kernels as K, ops
    B = cfg["batch"]
    dims = [cfg["n_sparse"] * cfg["dim"] + cfg["n_dense"]] + list(cfg["hidden"])
    prec = L.GEMM_BF16X3 if precision == "bf16x3" else L.GEMM_FP32
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
    out = []
    for li in range(len(dims) - 1):
        kin, nout = dims[li], dims[li + 1]
        ld = (kin + 3) // 4 * 4
        xw = torch.randn((B, ld), device=dev)
        x, w, dz = xw[:, :kin], torch.randn((kin, nout), device=dev), torch.randn((B, nout), device=dev)
        pl = (lambda t: K.split_planes(t)) if prec == L.GEMM_BF16X3 else (lambda t: None)
        xp, wp, dzp = pl(x), pl(w), pl(dz)
        dxw = torch.empty((B, ld), device=dev)
        calls = [
            ("fwd", B, nout, kin, lambda: K.gemm(x, w, precision=prec, m=B, n=nout, k=kin, a_planes=xp, b_planes=wp)),
            ("dgrad", B, kin, nout, lambda: K.gemm(dz, w, c=dxw[:, :kin], trans_b=True, precision=prec, m=B, n=kin,
                                                   k=nout, a_planes=dzp, b_planes=wp)),
            ("wgrad", kin, nout, B, lambda: K.gemm(x, dz, trans_a=True, precision=prec,
                                                   split_k=ops._split_k(kin, nout, B), m=kin, n=nout, k=B,
                                                   a_planes=xp, b_planes=dzp)),
        ]
        for name, m, n, k, fn in calls:
            fn()
            ts = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            out.append(("%s %d->%d" % (name, kin, nout), m, n, k, ts[len(ts) // 2]))
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "which": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "which": "fallback"}


# ------------------------------------------------------------------------------------------------
def cpu_step_factory(cfg, threads):
    """The reference math (oracle/) as one DeepFM SGD step on host cores: fwd + BCE + autograd bwd +
    row-wise SGD on the gathered rows (the reference's dense-Adam-over-tables semantics, SURVEY.md
    App. C, cannot run at this table size on any hardware)."""
    import torch
    from oracle import ops as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1024)
    F, E, nd, V = cfg["n_sparse"], cfg["dim"], cfg["n_dense"], cfg["vocab"]
    tables = [torch.randn(V, E, generator=g) * 1e-2 for _ in range(F)]
    lin = [torch.zeros(V, 1) for _ in range(F)]
    dims = [F * E + nd] + list(cfg["hidden"])
    ks = [(torch.randn(dims[i], dims[i + 1], generator=g) * (2.0 / (dims[i] + dims[i + 1])) ** 0.5).requires_grad_()
          for i in range(len(dims) - 1)]
    bs = [torch.zeros(d, requires_grad=True) for d in dims[1:]]
    wd = (torch.randn(dims[-1], 1, generator=g) * 0.1).requires_grad_()
    wl = (torch.randn(nd, 1, generator=g) * 0.1).requires_grad_()
    gb = torch.zeros(1, requires_grad=True)
    dense_params = ks + bs + [wd, wl, gb]

    def step(ids, dense, y):
        idx = torch.from_numpy(ids.astype(np.int64))
        rows = [O.embedding_lookup(tables[f], idx[:, f]).detach().requires_grad_() for f in range(F)]
        lrows = [O.embedding_lookup(lin[f], idx[:, f]).detach().requires_grad_() for f in range(F)]
        x = torch.cat(rows, dim=1)
        d = torch.from_numpy(dense)
        logit = O.linear(torch.cat(lrows, dim=-1), d, wl) + O.fm(x)
        h = O.dnn(torch.cat([x.flatten(1), d], dim=-1), ks, bs, "relu")
        logit = logit + h @ wd
        loss = O.binary_crossentropy(y, O.prediction(logit, gb, "binary"))
        loss.backward()
        with torch.no_grad():
            for p in dense_params:
                p -= LR * p.grad
                p.grad = None
            for f in range(F):
                tables[f].index_add_(0, idx[:, f], rows[f].grad[:, 0, :], alpha=-LR)
                lin[f].index_add_(0, idx[:, f], lrows[f].grad[:, 0, :], alpha=-LR)
        return float(loss)

    return step


def run_cpu(cfg, steps, warmup, sample_batch):
    import torch
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    step = cpu_step_factory(cfg, avail)
    small = dict(cfg, batch=sample_batch)
    data = synth_batches(small, 2)
    # "all the host threads it can use": torch's intra-op pool stops scaling (and can collapse) far below
    # the core count of a big host for these op sizes, so probe a few pool sizes and keep the fastest
    best, threads = None, avail
    for cand in sorted(set(min(avail, c) for c in (8, 16, 32, 64, avail))):
        torch.set_num_threads(cand)
        step(*data[0])
        t0 = time.perf_counter()
        step(*data[1])
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, threads = dt, cand
    torch.set_num_threads(threads)
    for i in range(warmup):
        step(*data[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        step(*data[i % 2])
    dt = time.perf_counter() - t0
    return steps * sample_batch / dt, dt / steps * 1e3, threads


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2ctr", choices=["b2ctr", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("B2CTR_GEMM", "auto"))
    ap.add_argument("--cpu-sample-batch", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        # the reference's own CPU path cannot run here (TensorFlow absent): the oracle port is timed
        if rank != 0:
            return
        sb = min(args.cpu_sample_batch, cfg["batch"])
        v, ms, threads = run_cpu(cfg, max(1, min(args.steps, 10)), min(warmup, 3), sb)
        line = {"impl": "reference", "metric": "samples/sec fwd+bwd DeepFM Criteo-synth", "value": v,
                "unit": "samples/s", "n_gpus": args.gpus, "steps": min(args.steps, 10), "warmup": min(warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "global_batch": sb, "optimizer": "sgd",
                           "note": "CPU arm (no GPU is used; n_gpus echoes the launch): oracle port = torch-CPU "
                                   "restatement of the reference math; TensorFlow is not installable here; each "
                                   "step is a %d-sample slice of the batch" % sb},
                "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                 "sample": "%d steps x %d samples of the c2 workload" % (min(args.steps, 10), sb)},
                "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the b2ctr path has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from deepctr_b200 import _lib as L, kernels as K, ops
    from deepctr_b200.engine import SGD
    from deepctr_b200.models import DeepFM

    precision = args.precision
    if precision == "auto":
        precision = "bf16x3"
    ops.set_gemm_precision(precision)
    cols = feature_columns(cfg)
    model = DeepFM(cols, cols, dnn_hidden_units=cfg["hidden"], l2_reg_linear=0, l2_reg_embedding=0, l2_reg_dnn=0)
    model.compile(SGD(LR), "binary_crossentropy", embedding_update="sparse")
    host = synth_batches(cfg, N_BATCHES, rank)
    dev = torch.device("cuda", local_rank)
    dev_batches = []
    for ids, dense, y in host:
        ids_d, dense_d, y_d = torch.from_numpy(ids).to(dev), torch.from_numpy(dense).to(dev), torch.from_numpy(y).to(dev)
        x = {"C%d" % (i + 1): ids_d[:, i:i + 1] for i in range(cfg["n_sparse"])}
        x["__dense__"] = dense_d
        dev_batches.append((x, ids_d, dense_d, y_d))

    def dev_inputs(b):
        _, ids_d, dense_d, y_d = dev_batches[b]
        x = {"C%d" % (i + 1): ids_d[:, i:i + 1] for i in range(cfg["n_sparse"])}
        x.update({"I%d" % (i + 1): dense_d[:, i:i + 1] for i in range(cfg["n_dense"])})
        return x, y_d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    # warm-up: >= `warmup` steps, extended until every distinct batch has its step graph captured (the
    # model replays the whole training step as one CUDA graph per input-buffer set once warm)
    i = 0
    while i < warmup or (i < warmup + N_BATCHES + 4 and model._graph_eligible()
                         and len(model._step_graphs) < N_BATCHES):
        x, y = dev_inputs(i % N_BATCHES)
        model.train_step(x, y)
        i += 1
    warmup_done = i
    barrier()
    L.reset_launch_count()
    model.replayed_launches = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        x, y = dev_inputs((warmup_done + i) % N_BATCHES)
        model.train_step(x, y)
    e1.record()
    barrier()
    launches = L.launch_count() + model.replayed_launches
    graph_replays = args.steps if model._step_graphs else 0
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    # per-kernel durations: the same K steps once more, launched eagerly with a CUDA-event pair around every
    # kernel group (the timed region above replays graphs, which cannot carry per-kernel events)
    K.PROFILE = {}
    for i in range(args.steps):
        x, y = dev_inputs((warmup_done + i) % N_BATCHES)
        model.train_step(x, y)
    torch.cuda.synchronize()
    prof = K.profile_summary()
    K.PROFILE = None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * cfg["batch"] * args.steps / (ms / 1e3)

    # ---- end-to-end through the public API: host arrays in, loss out -------------------------------
    # model.fit(x, y, batch_size=B) over `steps` batches of host arrays: every step packs its inputs into
    # pinned staging buffers, copies them H2D and reads its loss back D2H (asynchronously; the host waits
    # once per epoch, as Keras' fit does between epochs).
    host_x = {}
    reps = (args.steps + N_BATCHES - 1) // N_BATCHES
    for i in range(cfg["n_sparse"]):
        host_x["C%d" % (i + 1)] = np.concatenate([np.ascontiguousarray(h[0][:, i]) for h in host] * reps)[:args.steps * cfg["batch"]]
    for i in range(cfg["n_dense"]):
        host_x["I%d" % (i + 1)] = np.concatenate([np.ascontiguousarray(h[1][:, i]) for h in host] * reps)[:args.steps * cfg["batch"]]
    host_y = np.concatenate([h[2] for h in host] * reps)[:args.steps * cfg["batch"]]
    warm = {k: v[:3 * cfg["batch"]] for k, v in host_x.items()}
    model.fit(warm, host_y[:3 * cfg["batch"]], batch_size=cfg["batch"], epochs=1, shuffle=False, verbose=0)
    model._feeder.h2d_bytes = 0
    model.d2h_bytes = 0
    barrier()
    e0.record()
    model.fit(host_x, host_y, batch_size=cfg["batch"], epochs=1, shuffle=False, verbose=0)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * cfg["batch"] * args.steps / (e2e_ms / 1e3)
    h2d = model._feeder.h2d_bytes // args.steps
    d2h = model.d2h_bytes // args.steps

    gemm_times = gemm_launch_times(cfg, precision, dev) if rank == 0 else []
    if world > 1:
        model.close()          # step graphs hold NCCL kernels, the planner holds IPC mappings of peer shards
        dist.barrier()
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)

    peaks = measured_peaks()
    F, E, nd, B = cfg["n_sparse"], cfg["dim"], cfg["n_dense"], cfg["batch"]
    # algorithmic bytes per sample
    gather_fwd_bytes = F * 4 + F * E * 4 + F * E * 4 + F * 4 + 2 * nd * 4
    scatter_bwd_bytes = F * 4 + F * E * 4 + F * E * 4 + 2 * F * E * 4 + 2 * F * 4
    dims = [F * E + nd] + list(cfg["hidden"]) + [1]
    kernels = {}
    for name, (count, total_ms) in prof.items():
        kernels[name] = {"launches": count, "ms_per_step": total_ms / args.steps}
    def frac_hbm(name, bytes_per_sample):
        if name not in prof or prof[name][0] == 0:
            return None
        avg_ms = prof[name][1] / prof[name][0]
        a = bytes_per_sample * B / (avg_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": a / peaks["hbm_gbs"], "traffic": None, "kernel": name,
                "avg_launch_ms": avg_ms, "peak_source": peaks["which"],
                "algorithmic_bytes_per_launch": bytes_per_sample * B}
    roof_gather = frac_hbm("embed_gather_uniform_fwd", gather_fwd_bytes)
    roof_scatter = frac_hbm("embed_scatter_uniform_bwd", scatter_bwd_bytes)
    # DRAM bytes per launch from the committed ncu --set full captures of this workload (profiles/README.md)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and args.config == "c2" and world == 1:
        traffic = json.load(open(tpath))
    for r in (roof_gather, roof_scatter):
        if r is not None and r["kernel"] in traffic:
            r["traffic"] = traffic[r["kernel"]]
    gemm_ms = sum(v[1] for k, v in prof.items() if k.startswith("gemm"))
    # the DNN GEMM launches, each timed alone on the device (the eager per-group events above include the host's
    # launch gaps): algorithmic flops = 2*M*N*K per launch; in bf16x3 mode the tensor pipe executes 3x that
    roof_gemm = None
    if gemm_times:
        flops = sum(2.0 * m * n * k for _, m, n, k, _ in gemm_times)
        us = sum(t for *_, t in gemm_times)
        a = flops / (us * 1e-6) / 1e12
        roof_gemm = {"bound": "tensor", "achieved": a, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": a / peaks["bf16_tflops_sustained"], "traffic": None,
                     "kernel": ("gemm_planes_ws_kernel (tcgen05 cta_group::2, split-bf16)" if precision == "bf16x3"
                                else "sgemm_kernel (fp32 FFMA)"),
                     "launches_per_step": len(gemm_times), "avg_launch_ms": us / len(gemm_times) / 1e3,
                     "algorithmic_flops_per_step": flops,
                     "tensor_pipe_frac": (3.0 if precision == "bf16x3" else 1.0) * a / peaks["bf16_tflops_sustained"],
                     "note": "achieved = 2*M*N*K algorithmic flops / CUDA-event launch time, each launch timed alone "
                             "after an L2 flush; bf16x3 issues 3 bf16 MMAs per fp32 product, so frac <= 1/3",
                     "per_launch_us": {lab: round(t, 1) for lab, _, _, _, t in gemm_times},
                     "peak_source": peaks["which"] + " (dense bf16, sustained)"}
    cands = [r for r in (roof_gather, roof_scatter) if r is not None]
    shares = {"gather+scatter_ms": sum(prof.get(k, (0, 0))[1] for k in ("embed_gather_uniform_fwd",
                                                                       "embed_scatter_uniform_bwd")) / args.steps,
              "gemm_ms": (sum(t for *_, t in gemm_times) / 1e3) if gemm_times else gemm_ms / args.steps,
              "gemm_ms_eager_with_launch_gaps": gemm_ms / args.steps, "step_ms": ms_per_step,
              "measured": "eager pass of the same %d steps with a CUDA-event pair per kernel group" % args.steps}
    dominant = roof_gemm if (roof_gemm is not None and shares["gemm_ms"] > shares["gather+scatter_ms"]) else \
        (max(cands, key=lambda r: r["avg_launch_ms"]) if cands else None)

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (the N > 1 lines carry null)
        sb = min(args.cpu_sample_batch, cfg["batch"])
        v, cms, threads = run_cpu(cfg, 4, 1, sb)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "4 steps x %d samples of the same workload (oracle port, torch-CPU)" % sb}

    line = {"metric": "samples/sec fwd+bwd DeepFM Criteo-synth", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "global_batch": world * B, "optimizer": "sgd (fused row-wise)",
                       "gemm_precision": precision, "parallelism": ("tables row-sharded over %d GPUs (%s), dense part data-parallel" %
                                       (world, "NVLink peer loads / red.add" if getattr(model.planner, "peer_mode", False)
                                        else "NCCL all-to-all")) if world > 1 else "1 gpu",
                       "l2_flush": "none: %d distinct batches cycle; per step the path touches %.1f GB of "
                                   "randomly addressed table rows + activations, >> 126 MB L2"
                                   % (N_BATCHES, (gather_fwd_bytes + scatter_bwd_bytes) * B / 1e9)},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                    "api": "Model.fit(host arrays, batch_size=%d)" % B},
            "gpu_launches": int(launches), "graph_replays": int(graph_replays), "clocks": sampler.summary(),
            "roofline": dominant, "roofline_gather_fwd": roof_gather, "roofline_scatter_bwd": roof_scatter,
            "roofline_gemm": roof_gemm, "kernel_ms_per_step": kernels, "shares": shares,
            "cpu_baseline": cpu}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        os._exit(0)            # all ranks passed the barrier above; skip collective teardown at exit


if __name__ == "__main__":
    main()
