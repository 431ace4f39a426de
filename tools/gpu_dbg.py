"""debug: xDeepFM C3-shaped step-graph capture + per-weight update errors (eager vs graph)."""
import sys, os, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import bench as BN, b2_helpers as H
from oracle import models as OM, ops as O
from deepctr_b200 import engine as E, ops
from deepctr_b200.engine import SGD
warnings.simplefilter("always")
for mode in ("auto", "off"):
    cfg = dict(BN.CONFIGS["c3"], vocab=10000, batch=4096)
    E.clear_session()
    rng = np.random.RandomState(7)
    model = BN.build_model(cfg)
    H.randomize_weights(model, rng, std=0.05)
    cols = BN.feature_columns(cfg)
    data = [(BN.user_inputs(x), y) for x, y in BN.synth_batches(cfg, 3, 0, "uniform")]
    lr = 0.05
    model.compile(SGD(lr), "binary_crossentropy", embedding_update="sparse", step_graph=mode)
    for step in range(4):
        x, y = data[step % 3]
        W = H.oracle_weights(model, requires_grad=True)
        logit, pred = OM.xdeepfm(x, cols, cols, W, cin_layer_size=cfg["cin"])
        want = O.binary_crossentropy(y, pred); want.backward()
        got = model.train_on_batch(x, y)
        new, old = H.flat_params(H.oracle_weights(model)), H.flat_params(W)
        worst = []
        for name, p in old.items():
            if p.grad is None: continue
            upd = lr * p.grad.numpy(); w0 = p.detach().numpy()
            err = np.abs(new[name].numpy() - (w0 - upd)).max() / (np.abs(upd).max() + 1e-12)
            worst.append((float(err), name))
        worst.sort(reverse=True)
        print(mode, step, "loss", got, float(want.detach()), "graphs", len(model._step_graphs), "worst", worst[:4], flush=True)
