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
    # (3) fit
    t0 = time.perf_counter()
    model.fit(x, y, batch_size=B, epochs=1, shuffle=False, verbose=0)
    torch.cuda.synchronize()
    print("fit: %.3f ms/step" % ((time.perf_counter() - t0) / steps * 1e3))
    from deepctr_b200.inputs import Feeder
    print("copy calibration:", getattr(Feeder, "_COPY_TIMES", None))


if __name__ == "__main__":
    main()
