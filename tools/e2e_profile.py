"""Host-side breakdown of Model.fit at the C2 bench shape: staging (pack + H2D) vs step launch vs fit."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from deepctr_b200 import ops  # noqa: E402
from deepctr_b200.engine import SGD  # noqa: E402
from deepctr_b200.models import DeepFM  # noqa: E402
from deepctr_b200.inputs import slice_inputs  # noqa: E402


def main():
    cfg = bench.CONFIGS["c2"]
    ops.set_gemm_precision("bf16x3")
    cols = bench.feature_columns(cfg)
    model = DeepFM(cols, cols, dnn_hidden_units=cfg["hidden"], l2_reg_linear=0, l2_reg_embedding=0, l2_reg_dnn=0)
    model.compile(SGD(bench.LR), "binary_crossentropy", embedding_update="sparse")
    B, steps = cfg["batch"], 24
    host = bench.synth_batches(cfg, 4)
    x = {}
    for i in range(cfg["n_sparse"]):
        x["C%d" % (i + 1)] = np.concatenate([np.ascontiguousarray(h[0][:, i]) for h in host] * 6)
    for i in range(cfg["n_dense"]):
        x["I%d" % (i + 1)] = np.concatenate([np.ascontiguousarray(h[1][:, i]) for h in host] * 6)
    y = np.concatenate([h[2] for h in host] * 6)
    model.fit({k: v[:8 * B] for k, v in x.items()}, y[:8 * B], batch_size=B, epochs=1, shuffle=False, verbose=0)
    torch.cuda.synchronize()
    # (1) staging alone
    t0 = time.perf_counter()
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        t_a = time.perf_counter()
        bx, by = slice_inputs(x, sl), y[sl]
        t_b = time.perf_counter()
        staged = model._stage_batch(bx, by, model._stage_stream)
        if s == steps - 1:
            print("last batch: slice %.3f ms, stage %.3f ms" % ((t_b - t_a) * 1e3, (time.perf_counter() - t_b) * 1e3))
    torch.cuda.synchronize()
    print("staging only: %.3f ms/batch" % ((time.perf_counter() - t0) / steps * 1e3))
    # (2) step launch alone (graph replay), inputs already staged
    staged = model._stage_batch(slice_inputs(x, slice(0, B)), y[:B])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        model._loss_step(None, None, True, staged=staged)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("step launch: host %.3f ms/step, device-complete %.3f ms/step, graphs=%d" %
          ((t1 - t0) / steps * 1e3, (time.perf_counter() - t0) / steps * 1e3, len(model._step_graphs)))
    # (3) fit: slope (per step) and intercept (per call)
    res = {}
    for nst in (6, 12, 24, 24):
        xs = {k: v[:nst * B] for k, v in x.items()}
        t0 = time.perf_counter()
        model.fit(xs, y[:nst * B], batch_size=B, epochs=1, shuffle=False, verbose=0)
        torch.cuda.synchronize()
        res[nst] = (time.perf_counter() - t0) * 1e3
        print("fit %d steps: %.3f ms total, %.3f ms/step" % (nst, res[nst], res[nst] / nst))
    print("slope %.3f ms/step, intercept %.3f ms" % ((res[24] - res[12]) / 12, res[12] - 12 * (res[24] - res[12]) / 12))
    # host-side: how long does the staging call take INSIDE fit (worker thread), and where
    import deepctr_b200.inputs as I
    rec = {"stage": [], "feed": [], "labels": [], "fill": [], "upload": [], "slot": []}

    def wrap(obj, name, key):
        orig_f = getattr(obj, name)

        def w(*a, **kw):
            t = time.perf_counter()
            try:
                return orig_f(*a, **kw)
            finally:
                rec[key].append((time.perf_counter() - t) * 1e3)
        setattr(obj, name, w)
        return orig_f

    fd = model._feeder
    o1 = wrap(model, "_stage_batch", "stage")
    o2 = wrap(fd, "feed", "feed")
    o3 = wrap(fd, "labels", "labels")
    o4 = wrap(fd, "_fill", "fill")
    o5 = wrap(fd, "_upload", "upload")
    o6 = wrap(fd, "_next_slot", "slot")
    model.fit(x, y, batch_size=B, epochs=1, shuffle=False, verbose=0)
    torch.cuda.synchronize()
    for k, v in rec.items():
        print("in-fit %-7s n=%3d mean %.3f ms  max %.3f" % (k, len(v), sum(v) / max(len(v), 1), max(v) if v else 0))
    model._stage_batch, fd.feed, fd.labels, fd._fill, fd._upload, fd._next_slot = o1, o2, o3, o4, o5, o6
    # GPU-side timeline of one fit: events around every step on the compute stream
    evs = []
    orig = model._loss_step

    def traced(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        staged = kw.get("staged")
        if staged is not None and staged[2] is not None:
            torch.cuda.current_stream().wait_event(staged[2])     # so that e0 is after the H2D wait
        e0.record()
        out = orig(*a, **kw)
        e1.record()
        evs.append((e0, e1, time.perf_counter()))
        return out

    model._loss_step = traced
    t0 = time.perf_counter()
    model.fit(x, y, batch_size=B, epochs=1, shuffle=False, verbose=0)
    torch.cuda.synchronize()
    model._loss_step = orig
    durs = [e0.elapsed_time(e1) for e0, e1, _ in evs]
    gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(len(evs) - 1)]
    host = [(evs[i + 1][2] - evs[i][2]) * 1e3 for i in range(len(evs) - 1)]
    print("step GPU ms:", " ".join("%.2f" % d for d in durs))
    print("gap  GPU ms:", " ".join("%.2f" % d for d in gaps))
    print("host period:", " ".join("%.2f" % d for d in host))
    from deepctr_b200.inputs import Feeder
    print("copy calibration:", getattr(Feeder, "_COPY_TIMES", None))


if __name__ == "__main__":
    main()
