#!/usr/bin/env python
"""Debug aid: which call invalidates the step-graph capture of the dense-Adam DeepFM test?
Wraps deepctr_b200._lib.check so that the capture status of the current stream is read after every library call
and reports the first call after which the capture is invalidated (with the Python stack)."""
import ctypes as C
import os
import sys
import traceback
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import b2_helpers as H  # noqa: E402
from deepctr_b200 import _lib as L, ops  # noqa: E402
from deepctr_b200.models import DeepFM  # noqa: E402

rt = C.CDLL("libcudart.so.12")
rt.cudaStreamIsCapturing.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
state = {"last_ok": None, "reported": False, "n": 0}
_check = L.check


def status():
    st = C.c_int(-1)
    rc = rt.cudaStreamIsCapturing(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(st))
    return rc, st.value


def check(status_code, what=""):
    _check(status_code, what)
    rc, st = status()
    state["n"] += 1
    if (st == 2 or rc != 0) and not state["reported"]:
        state["reported"] = True
        print("capture INVALIDATED after call #%d %r (rc=%d, status=%d); last good call: %r"
              % (state["n"], what, rc, st, state["last_ok"]))
        print("".join(traceback.format_stack(limit=12)))
    elif st == 1:
        state["last_ok"] = what


L.check = check
import deepctr_b200.kernels as K  # noqa: E402
for mod in (K,):
    if hasattr(mod, "L"):
        mod.L.check = check

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
ops.set_gemm_precision(prec)
rng = np.random.RandomState(13)
cols, x, y = H.criteo_like(rng, 96)
model = DeepFM(cols, cols, dnn_hidden_units=(32, 16), l2_reg_linear=0, l2_reg_embedding=0)
H.randomize_weights(model, rng)
model.compile("adam", "binary_crossentropy", embedding_update="dense")
with warnings.catch_warnings(record=True) as wl:
    warnings.simplefilter("always")
    for i in range(5):
        print("train_on_batch", i, model.train_on_batch(x, y), "graphs:", len(model._step_graphs), "mode:", model.step_graph)
    print("fit...")
    model.fit(x, y, batch_size=32, epochs=2, verbose=0, validation_split=0.25)
    print("graphs:", len(model._step_graphs), "mode:", model.step_graph)
    for w_ in wl:
        print("WARNING:", str(w_.message)[:300])
