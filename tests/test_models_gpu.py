"""GPU parity of the xDeepFM / DCN / AutoInt / DIN builders (forward logits and one SGD step) against
the CPU oracle, plus varlen / hash / shared-embedding feature handling.  fp32 logits within 1e-4 rel."""
import numpy as np
import pytest
import torch

from oracle import models as OM
from oracle import ops as O
import b2_helpers as H

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gemm_precision")]


def _check_step(model, oracle_fn, x, y, lr=0.05, steps=2, tol=3e-3):
    from deepctr_b200.engine import SGD
    model.compile(SGD(lr), "binary_crossentropy", embedding_update="dense")
    for step in range(steps):
        W = H.oracle_weights(model, requires_grad=True)
        _, pred = oracle_fn(W)
        got_pred = model.predict(x, batch_size=len(y))
        assert H.rel_err(got_pred, pred.detach().numpy()) < 1e-4
        loss = O.binary_crossentropy(y, pred)
        loss.backward()
        got = model.train_on_batch(x, y)
        assert abs(got - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
        new, old = H.flat_params(H.oracle_weights(model)), H.flat_params(W)
        for name, p in old.items():
            if p.grad is None or "moving" in name:
                continue
            want = (p.detach() - lr * p.grad).numpy()
            scale = np.abs(lr * p.grad.numpy()).max() + 1e-9
            err = np.abs(new[name].numpy() - want).max() / scale
            assert err < tol, "step %d %s: %.3e" % (step, name, err)


def test_xdeepfm(cuda):
    from deepctr_b200.models import xDeepFM
    rng = np.random.RandomState(21)
    cols, x, y = H.criteo_like(rng, 64, n_sparse=4, n_dense=2, dim=4)
    model = xDeepFM(cols, cols, dnn_hidden_units=(16, 8), cin_layer_size=(8, 6), l2_reg_linear=0,
                    l2_reg_embedding=0)
    H.randomize_weights(model, rng, 0.2)
    _check_step(model, lambda W: OM.xdeepfm(x, cols, cols, W, (8, 6), True, "relu"), x, y)
    # empty CIN is legal (tests/models/xDeepFM_test.py)
    m2 = xDeepFM(cols, cols, dnn_hidden_units=(8,), cin_layer_size=())
    assert m2.predict(x, batch_size=64).shape == (64, 1)


@pytest.mark.parametrize("param,cross_num,hidden", [("vector", 2, (16, 8)), ("matrix", 1, (8,)), ("vector", 1, ())])
def test_dcn(cuda, param, cross_num, hidden):
    from deepctr_b200.models import DCN
    rng = np.random.RandomState(22)
    cols, x, y = H.criteo_like(rng, 48, n_sparse=3, n_dense=2, dim=4)
    model = DCN(cols, cols, cross_num=cross_num, cross_parameterization=param, dnn_hidden_units=hidden,
                l2_reg_linear=0, l2_reg_embedding=0, l2_reg_cross=0)
    H.randomize_weights(model, rng, 0.2)
    _check_step(model, lambda W: OM.dcn(x, cols, cols, W, cross_num, param, use_dnn=len(hidden) > 0), x, y)
    with pytest.raises(ValueError):
        DCN(cols, cols, cross_num=0, dnn_hidden_units=())


def test_dcn_empty_linear_columns(cuda):
    """DCN([], cols) is legal and yields a constant-zero linear logit (tests/models/DCN_test.py:25-32)."""
    from deepctr_b200.models import DCN
    rng = np.random.RandomState(23)
    cols, x, y = H.criteo_like(rng, 16, n_sparse=2, n_dense=1, dim=4)
    model = DCN([], cols, cross_num=1, dnn_hidden_units=(4,))
    assert model.predict(x, batch_size=16).shape == (16, 1)


def test_autoint(cuda):
    from deepctr_b200.models import AutoInt
    rng = np.random.RandomState(24)
    cols, x, y = H.criteo_like(rng, 40, n_sparse=4, n_dense=2, dim=4)
    model = AutoInt(cols, cols, att_layer_num=2, att_embedding_size=3, att_head_num=2, dnn_hidden_units=(8,),
                    l2_reg_linear=0, l2_reg_embedding=0)
    H.randomize_weights(model, rng, 0.3)
    _check_step(model, lambda W: OM.autoint(x, cols, cols, W, 2, 3, 2, True), x, y)
    with pytest.raises(ValueError):
        AutoInt(cols, cols, att_layer_num=0, dnn_hidden_units=())


def _din_data(rng, n=24, T=6, vocab=30):
    from deepctr_b200.feature_column import SparseFeat, VarLenSparseFeat, DenseFeat
    cols = [SparseFeat('user', 8, embedding_dim=4), SparseFeat('gender', 3, embedding_dim=4),
            SparseFeat('item_id', vocab, embedding_dim=8), SparseFeat('cate_id', 7, embedding_dim=4),
            DenseFeat('pay_score', 1)]
    cols += [VarLenSparseFeat(SparseFeat('hist_item_id', vocab, embedding_dim=8, embedding_name='item_id'),
                              maxlen=T, length_name="seq_length"),
             VarLenSparseFeat(SparseFeat('hist_cate_id', 7, embedding_dim=4, embedding_name='cate_id'),
                              maxlen=T, length_name="seq_length")]
    lens = rng.randint(1, T + 1, size=n)
    hi = rng.randint(1, vocab, size=(n, T))
    hc = rng.randint(1, 7, size=(n, T))
    for b in range(n):
        hi[b, lens[b]:] = 0
        hc[b, lens[b]:] = 0
    hc[0, 1] = 0     # a position valid for item but padded for cate: the AND of the masks matters
    x = {'user': rng.randint(0, 8, n).astype(np.int32), 'gender': rng.randint(0, 3, n).astype(np.int32),
         'item_id': rng.randint(1, vocab, n).astype(np.int32), 'cate_id': rng.randint(1, 7, n).astype(np.int32),
         'pay_score': rng.rand(n).astype(np.float32), 'hist_item_id': hi.astype(np.int32),
         'hist_cate_id': hc.astype(np.int32), 'seq_length': lens.astype(np.int32)}
    y = (rng.rand(n) < 0.4).astype(np.float32)
    return cols, x, y


@pytest.mark.parametrize("act,norm", [("sigmoid", False), ("dice", False), ("sigmoid", True)])
def test_din(cuda, act, norm):
    from deepctr_b200.models import DIN
    rng = np.random.RandomState(25)
    cols, x, y = _din_data(rng)
    hist = ["item_id", "cate_id"]
    model = DIN(cols, hist, dnn_hidden_units=(8, 4), att_hidden_size=(6, 5), att_activation=act,
                att_weight_normalization=norm, l2_reg_embedding=0)
    H.randomize_weights(model, rng, 0.3)
    # shared tables: hist_item_id uses item_id's table, created with mask_zero (tests/feature_test.py:35-50)
    emb = model.get_layer("sparse_emb_item_id")
    assert emb.mask_zero
    assert not any(l.name == "sparse_seq_emb_hist_item_id" for l in model.layers)
    W = H.oracle_weights(model)
    _, want = OM.din(x, cols, hist, W, act, norm, training=False)
    got = model.predict(x, batch_size=len(y))
    assert H.rel_err(got, want.numpy()) < 1e-4
    if act != "dice":    # dice in training uses batch statistics: covered by the layer test
        _check_step(model, lambda W: OM.din(x, cols, hist, W, act, norm), x, y)
    else:
        from deepctr_b200.engine import SGD
        model.compile(SGD(0.05), "binary_crossentropy", embedding_update="dense")
        l0 = model.train_on_batch(x, y)
        for _ in range(10):
            l1 = model.train_on_batch(x, y)
        assert np.isfinite(l1) and l1 < l0


def test_varlen_pooled_weighted_hashed_features(cuda):
    """The feature mix of tests/utils.py::get_test_data: sum/mean/max varlen bags with and without
    weights / lengths, hashed sparse features, two groups - all served by the fused generic gather."""
    from deepctr_b200.feature_column import SparseFeat, VarLenSparseFeat, DenseFeat
    from deepctr_b200.models import DeepFM
    rng = np.random.RandomState(26)
    n, T = 40, 5
    cols = [SparseFeat("s0", 20, 4), SparseFeat("s1", 16, 4, use_hash=True, dtype="int64"),
            DenseFeat("d0", 2),
            VarLenSparseFeat(SparseFeat("v_sum", 12, 4), maxlen=T, combiner="sum", length_name="len_a"),
            VarLenSparseFeat(SparseFeat("v_mean", 12, 4), maxlen=T, combiner="mean"),
            VarLenSparseFeat(SparseFeat("v_max", 12, 4, use_hash=True, dtype="int64"), maxlen=T, combiner="max",
                             length_name="len_a", weight_name="w_a", weight_norm=True),
            VarLenSparseFeat(SparseFeat("v_w", 12, 4), maxlen=T, combiner="sum", weight_name="w_b",
                             weight_norm=False)]
    lens = rng.randint(1, T + 1, size=n)

    def seq(vocab):
        a = rng.randint(1, vocab, size=(n, T))
        for b in range(n):
            a[b, lens[b]:] = 0
        return a

    x = {"s0": rng.randint(0, 20, n).astype(np.int32), "s1": rng.randint(0, 10 ** 6, n).astype(np.int64),
         "d0": rng.rand(n, 2).astype(np.float32), "v_sum": seq(12).astype(np.int32),
         "v_mean": seq(12).astype(np.int32), "v_max": seq(10 ** 5).astype(np.int64), "v_w": seq(12).astype(np.int32),
         "len_a": lens.astype(np.int32), "w_a": rng.rand(n, T, 1).astype(np.float32),
         "w_b": rng.rand(n, T, 1).astype(np.float32)}
    y = (rng.rand(n) < 0.5).astype(np.float32)
    model = DeepFM(cols, cols, dnn_hidden_units=(8,), l2_reg_linear=0, l2_reg_embedding=0)
    H.randomize_weights(model, rng, 0.3)
    _check_step(model, lambda W: OM.deepfm(x, cols, cols, W), x, y)
    # every lookup + pooling chain was planned into the fused launch
    p = model.planner
    assert all(s.pool != 0 for s in p.slots if s.maxlen > 1)


def test_feature_column_api_errors():
    """tests/feature_test.py:53-60 and feature_column.py:24-31 behaviours (no GPU needed to raise)."""
    from deepctr_b200.feature_column import SparseFeat, VarLenSparseFeat, build_input_features
    from deepctr_b200.models import DeepFM
    with pytest.raises(ValueError, match="same embedding_name"):
        cols = [SparseFeat("item", 10, 4), VarLenSparseFeat(SparseFeat("hist", 11, 4, embedding_name="item"), 3)]
        DeepFM(cols, cols)
    with pytest.raises(ValueError, match="requires use_hash=True"):
        build_input_features([SparseFeat("s", 10, 4, dtype="string")])
