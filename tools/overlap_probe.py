"""Does concurrent H2D staging slow the replayed training step?  (a) steps alone over the 3 ring slots,
(b) with a side stream doing only the H2D copies, (c) side stream also running pack_rows."""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from deepctr_b200 import ops, kernels as K  # noqa: E402
from deepctr_b200.engine import SGD  # noqa: E402
from deepctr_b200.models import DeepFM  # noqa: E402
from deepctr_b200.inputs import slice_inputs  # noqa: E402


def main():
    cfg = bench.CONFIGS["c2"]
    ops.set_gemm_precision("bf16x3")
    cols = bench.feature_columns(cfg)
    model = DeepFM(cols, cols, dnn_hidden_units=cfg["hidden"], l2_reg_linear=0, l2_reg_embedding=0, l2_reg_dnn=0)
    model.compile(SGD(bench.LR), "binary_crossentropy", embedding_update="sparse")
    B = cfg["batch"]
    host = bench.synth_batches(cfg, 4)
    x = {}
    for i in range(cfg["n_sparse"]):
        x["C%d" % (i + 1)] = np.concatenate([np.ascontiguousarray(h[0][:, i]) for h in host] * 3)
    for i in range(cfg["n_dense"]):
        x["I%d" % (i + 1)] = np.concatenate([np.ascontiguousarray(h[1][:, i]) for h in host] * 3)
    y = np.concatenate([h[2] for h in host] * 3)
    model.fit(x, y, batch_size=B, epochs=1, shuffle=False, verbose=0)
    staged = [model._stage_batch(slice_inputs(x, slice(s * B, (s + 1) * B)), y[s * B:(s + 1) * B]) for s in range(3)]
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    pin_i = torch.empty(B * 26, dtype=torch.int32).pin_memory()
    pin_f = torch.empty(B * 13, dtype=torch.float32).pin_memory()
    dev_i = torch.empty(B * 26, dtype=torch.int32, device="cuda")
    dev_f = torch.empty(B * 13, dtype=torch.float32, device="cuda")
    dev_p = torch.empty((B, 13), dtype=torch.float32, device="cuda")

    def steps(n=60):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(n):
            model._loss_step(None, None, True, staged=staged[s % 3])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    print("(a) alone: %.3f ms/step" % steps())
    for mode in ("h2d", "h2d+pack", "pack"):
        stop = [False]

        def bg():
            torch.cuda.set_device(0)
            with torch.cuda.stream(side):
                while not stop[0]:
                    if "h2d" in mode:
                        dev_i.copy_(pin_i, non_blocking=True)
                        dev_f.copy_(pin_f, non_blocking=True)
                    if "pack" in mode:
                        K.pack_rows(dev_f, [1] * 13, B, out=dev_p)
                    side.synchronize()
                    time.sleep(0.0008)

        th = threading.Thread(target=bg)
        th.start()
        time.sleep(0.05)
        r = steps()
        stop[0] = True
        th.join()
        print("(b) with %-9s in the background: %.3f ms/step" % (mode, r))


if __name__ == "__main__":
    main()
