"""GPU parity of the DeepFM path end to end (feature columns -> fused gather -> FM -> DNN -> logit ->
loss -> backward -> update) against the CPU oracle.  Tolerance from BASELINE.json's north_star:
fp32 logits within 1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import models as OM
from oracle import ops as O
import b2_helpers as H

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gemm_precision")]


def _build(rng, **kw):
    from deepctr_b200.models import DeepFM
    cols, x, y = H.criteo_like(rng, kw.pop("n", 96), **kw)
    model = DeepFM(cols, cols, dnn_hidden_units=(32, 16), l2_reg_linear=0, l2_reg_embedding=0)
    H.randomize_weights(model, rng)
    return model, cols, x, y


@pytest.mark.parametrize("dim", [8, 6])     # 8: Criteo fast path; 6: generic gather + standalone FM
def test_deepfm_forward_matches_oracle(cuda, dim):
    rng = np.random.RandomState(11)
    model, cols, x, y = _build(rng, dim=dim)
    model.compile("sgd", "binary_crossentropy", embedding_update="dense")
    W = H.oracle_weights(model)
    logit, want = OM.deepfm(x, cols, cols, W)
    for rep in range(3):      # step 0 unfused, later steps run the fused FM / linear / dense tail
        got = model.predict(x, batch_size=64)
        assert got.shape == (len(y), 1)
        assert H.rel_err(got, want.numpy()) < 1e-4, "rep %d" % rep
    if dim == 8:
        p = model.planner
        assert p.fast and p.fm_hint is not None and p.lin_hint and p.tail_hint is not None


@pytest.mark.parametrize("mode", ["dense", "sparse"])
def test_deepfm_train_step_matches_oracle_sgd(cuda, mode):
    rng = np.random.RandomState(12)
    model, cols, x, y = _build(rng)
    from deepctr_b200.engine import SGD
    lr = 0.05
    model.compile(SGD(lr), "binary_crossentropy", embedding_update=mode)
    for step in range(3):     # covers the unfused first step and the fused later ones
        W = H.oracle_weights(model, requires_grad=True)
        logit, pred = OM.deepfm(x, cols, cols, W)
        loss = O.binary_crossentropy(y, pred)
        loss.backward()
        got_loss = model.train_on_batch(x, y)
        assert abs(got_loss - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
        W2 = H.oracle_weights(model)
        new, old = H.flat_params(W2), H.flat_params(W)
        for name, p in old.items():
            if p.grad is None:
                continue
            want = (p.detach() - lr * p.grad).numpy()
            got = new[name].numpy()
            scale = np.abs(lr * p.grad.numpy()).max() + 1e-12
            err = np.abs(got - want).max() / scale
            assert err < 2e-3, "step %d weight %s: update mismatch %.3e (relative to max update)" % (step, name, err)


def test_deepfm_adam_dense_matches_keras_semantics(cuda):
    """Keras Adam is dense over the tables: every row moves once m, v are non-zero."""
    rng = np.random.RandomState(13)
    model, cols, x, y = _build(rng)
    model.compile("adam", "binary_crossentropy", embedding_update="dense")
    tabs_before = {w.name: w.value() for w in model.planner.tables()}
    l0 = model.train_on_batch(x, y)
    l1 = model.train_on_batch(x, y)
    for _ in range(20):
        l2 = model.train_on_batch(x, y)
    assert l2 < l0
    hist = model.fit(x, y, batch_size=32, epochs=2, verbose=0, validation_split=0.25)
    assert len(hist.history["loss"]) == 2 and len(hist.history["val_loss"]) == 2
    ev = model.evaluate(x, y, batch_size=50)
    assert np.isfinite(ev)


def test_inputs_as_list_and_2d_columns(cuda):
    """Appendix F.1: inputs may be a list ordered like get_feature_names, columns [N] or [N,1]."""
    from deepctr_b200.feature_column import get_feature_names
    rng = np.random.RandomState(14)
    model, cols, x, y = _build(rng)
    names = get_feature_names(cols)
    a = model.predict(x, batch_size=256)
    b = model.predict([x[n].reshape(-1, 1) for n in names], batch_size=17)
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)


def test_save_and_load_weights_roundtrip(cuda, tmp_path):
    rng = np.random.RandomState(15)
    model, cols, x, y = _build(rng)
    a = model.predict(x, batch_size=256)
    p = str(tmp_path / "w")
    model.save_weights(p)
    for w in model.weights:
        w.set_value(np.zeros(w.shape, np.float32))
    model.load_weights(p)
    # the second predict runs the fused FM / linear epilogues (different fp32 summation order)
    np.testing.assert_allclose(a, model.predict(x, batch_size=256), rtol=1e-5, atol=1e-6)


def test_training_step_graph_replay_matches_eager(cuda):
    """After two eager steps the whole training step is captured per staging-ring slot and replayed as a CUDA
    graph (engine.Model._loss_step).  Same kernels, same order: the losses must agree with a model that keeps
    launching eagerly, up to the order of the fp32 atomics in the embedding scatter."""
    import numpy as np
    from deepctr_b200.models import DeepFM
    from deepctr_b200.engine import SGD
    from deepctr_b200.feature_column import SparseFeat, DenseFeat, VarLenSparseFeat
    cols = [SparseFeat("C%d" % i, 50 + i, 8) for i in range(5)] + [DenseFeat("I%d" % i, 1) for i in range(3)] + \
           [VarLenSparseFeat(SparseFeat("V", 30, 8), maxlen=4, combiner="mean")]
    rng = np.random.RandomState(3)
    n, bs = 96 * 9, 96
    x = {"C%d" % i: rng.randint(0, 50 + i, n).astype(np.int32) for i in range(5)}
    x.update({"I%d" % i: rng.rand(n).astype(np.float32) for i in range(3)})
    x["V"] = rng.randint(0, 30, (n, 4)).astype(np.int32)
    y = rng.randint(0, 2, n).astype(np.float32)
    losses = {}
    w0 = None
    for mode in ("off", "auto"):
        m = DeepFM(cols, cols, dnn_hidden_units=(16, 8), seed=7)
        m.compile(SGD(0.05), "binary_crossentropy", step_graph=mode)
        if w0 is None:
            w0 = [w.copy() for w in m.get_weights()]
        else:
            m.set_weights(w0)            # layer names (hence initialiser streams) differ between instances
        h = m.fit(x, y, batch_size=bs, epochs=3, shuffle=False, verbose=0)
        per_batch = [m.test_on_batch({k: v[:bs] for k, v in x.items()}, y[:bs])]
        losses[mode] = (h.history["loss"], per_batch, m)
    m = losses["auto"][2]
    assert len(m._step_graphs) == 3 and m.replayed_launches > 0      # one graph per staging-ring slot
    assert not losses["off"][2]._step_graphs
    np.testing.assert_allclose(losses["auto"][0], losses["off"][0], rtol=2e-5)
    np.testing.assert_allclose(losses["auto"][1], losses["off"][1], rtol=2e-5)
    for wa, wo in zip(m.weights, losses["off"][2].weights):
        np.testing.assert_allclose(wa.value(), wo.value(), rtol=1e-4, atol=1e-6)
    # a model with dropout keeps launching eagerly (the Philox offset is a by-value kernel argument)
    md = DeepFM(cols, cols, dnn_hidden_units=(16, 8), dnn_dropout=0.5, seed=7)
    md.compile(SGD(0.05), "binary_crossentropy")
    md.fit(x, y, batch_size=bs, epochs=1, shuffle=False, verbose=0)
    assert not md._step_graphs


def test_adam_training_is_graph_replayed_and_matches_eager(cuda):
    """Adam keeps its step count on the device (b2ctr_adam_step_dev), so the default optimizer of the reference's
    examples is replayed as a step graph too; dense (Keras) embedding updates."""
    import numpy as np
    from deepctr_b200.models import DeepFM
    from deepctr_b200.feature_column import SparseFeat, DenseFeat
    cols = [SparseFeat("C%d" % i, 40 + i, 8) for i in range(4)] + [DenseFeat("I%d" % i, 1) for i in range(2)]
    rng = np.random.RandomState(5)
    n, bs = 64 * 8, 64
    x = {"C%d" % i: rng.randint(0, 40 + i, n).astype(np.int32) for i in range(4)}
    x.update({"I%d" % i: rng.rand(n).astype(np.float32) for i in range(2)})
    y = rng.randint(0, 2, n).astype(np.float32)
    hist, w0 = {}, None
    for mode in ("off", "auto"):
        m = DeepFM(cols, cols, dnn_hidden_units=(16, 8), seed=3)
        m.compile("adam", "binary_crossentropy", step_graph=mode)
        if w0 is None:
            w0 = [w.copy() for w in m.get_weights()]
        else:
            m.set_weights(w0)
        h = m.fit(x, y, batch_size=bs, epochs=2, shuffle=False, verbose=0)
        hist[mode] = (h.history["loss"], m)
    assert len(hist["auto"][1]._step_graphs) == 3 and not hist["off"][1]._step_graphs
    np.testing.assert_allclose(hist["auto"][0], hist["off"][0], rtol=2e-5)
    for wa, wo in zip(hist["auto"][1].weights, hist["off"][1].weights):
        np.testing.assert_allclose(wa.value(), wo.value(), rtol=1e-4, atol=1e-6)
