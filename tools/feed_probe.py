#!/usr/bin/env python
"""Where does Model.fit's end-to-end time go on this box?  C2-shaped DeepFM, host arrays in.
Separates the per-step pipeline rate from the fixed per-epoch cost (thread start, first batch staged with nothing to
overlap, end-of-epoch sync) by fitting epochs of 20 / 40 / 80 steps, and times the producer's stages per batch."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    cfg = bench.resolve(bench.CONFIGS[os.environ.get("PROBE_CONFIG", "c2")], 1)
    from deepctr_b200 import ops, inputs as I
    from deepctr_b200.engine import SGD
    ops.set_gemm_precision("bf16x3")
    model = bench.build_model(cfg)
    model.compile(SGD(bench.LR), "binary_crossentropy", embedding_update="sparse")
    B = cfg["batch"]
    host = bench.synth_batches(cfg, 4, 0, "uniform")
    out = {"cores": len(os.sched_getaffinity(0)), "pack_threads": I._pack_threads()}

    def arrays(steps):
        reps = (steps + 3) // 4
        x = {k: np.concatenate([bench.user_inputs(h[0])[k] for h in host] * reps)[:steps * B] for k in bench.user_inputs(host[0][0])}
        y = np.concatenate([h[1] for h in host] * reps)[:steps * B]
        return x, y

    x, y = arrays(8)
    model.fit(x, y, batch_size=B, epochs=1, shuffle=False, verbose=0)
    for steps in (20, 40, 80):
        x, y = arrays(steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.fit(x, y, batch_size=B, epochs=1, shuffle=False, verbose=0)
        torch.cuda.synchronize()
        out["fit_%d_steps_ms" % steps] = (time.perf_counter() - t0) * 1e3
    out["per_step_ms"] = (out["fit_80_steps_ms"] - out["fit_20_steps_ms"]) / 60.0
    out["fixed_ms"] = out["fit_20_steps_ms"] - 20 * out["per_step_ms"]
    # timeline of one 20-step fit: host time of every step launch, device time at which every step finished
    x, y = arrays(20)
    orig = model._loss_step
    host_t, evs = [], []

    def traced(*a, **k):
        host_t.append(time.perf_counter())
        out_ = orig(*a, **k)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        evs.append(ev)
        return out_

    model._loss_step = traced
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    model.fit(x, y, batch_size=B, epochs=1, shuffle=False, verbose=0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    model._loss_step = orig
    out["trace_fit_ms"] = (t1 - t0) * 1e3
    out["trace_host_launch_ms"] = [round((t - t0) * 1e3, 3) for t in host_t]
    out["trace_dev_done_ms"] = [round(e0.elapsed_time(e), 3) for e in evs]
    # producer stages in isolation (no training step running)
    x, y = arrays(40)
    from deepctr_b200.inputs import slice_inputs
    st = model._stage_stream
    t_slice = t_stage = 0.0
    for i in range(40):
        t0 = time.perf_counter()
        bx, by = slice_inputs(x, slice(i * B, (i + 1) * B)), y[i * B:(i + 1) * B]
        t1 = time.perf_counter()
        staged = model._stage_batch(bx, by, st)
        t2 = time.perf_counter()
        model._feeder.consumed(staged[3])
        torch.cuda.synchronize()
        if i >= 8:
            t_slice += t1 - t0
            t_stage += t2 - t1
    out["producer_slice_ms"] = t_slice / 32 * 1e3
    out["producer_stage_ms"] = t_stage / 32 * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
