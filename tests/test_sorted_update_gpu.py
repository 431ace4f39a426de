"""GPU: the deterministic fused embedding update (sort by (feature, id) + ordered segmented reduce;
b2ctr_embed_update_sorted) - north_star's "bit-exact segment sums" for the backward pass.

* bit-exact against a numpy restatement that sums the gradient rows of duplicate ids in ascending sample
  order in fp32 and applies ONE update per row;
* bit-identical from run to run (the atomic path is not);
* Keras' (lazy, sparse) Adagrad on the same path, against torch.optim.Adagrad on the CPU oracle.
"""
import numpy as np
import pytest
import torch

import b2_helpers as H
from oracle import models as OM
from oracle import ops as O

pytestmark = pytest.mark.gpu


def _setup(rng, B, F, V, dim, cuda, with_fm=True):
    from deepctr_b200 import kernels as K
    tabs = [rng.normal(0, 0.3, size=(V, dim)).astype(np.float32) for _ in range(F)]
    lins = [rng.normal(0, 0.3, size=(V,)).astype(np.float32) for _ in range(F)]
    idx = rng.randint(0, V, size=(B, F)).astype(np.int32)
    idx[:, 0] = rng.randint(0, 3, size=B)                    # heavy duplicates in feature 0
    idx[5, 1] = -1                                           # an out-of-range id: skipped
    ldx = F * dim + 4
    x = np.zeros((B, ldx), np.float32)
    for f in range(F):
        ok = (idx[:, f] >= 0) & (idx[:, f] < V)
        x[ok, f * dim:(f + 1) * dim] = tabs[f][idx[ok, f]]
    dx = rng.normal(size=(B, ldx)).astype(np.float32)
    dfm = rng.normal(size=(B,)).astype(np.float32) if with_fm else None
    dlin = rng.normal(size=(B,)).astype(np.float32)
    return tabs, lins, idx, x, dx, dfm, dlin, ldx


def _want(tabs, lins, idx, x, dx, dfm, dlin, dim, lr, fm_mask, adagrad=False, acc0=0.1, eps=1e-7):
    """numpy restatement: per (feature, id) ascending-sample fp32 sums, one update per row."""
    B, F = idx.shape
    f32 = np.float32
    S = np.zeros((B, dim), f32)
    for f in range(F):
        if (fm_mask >> f) & 1:
            S = (S + x[:, f * dim:(f + 1) * dim]).astype(f32)
    out_t = [t.copy() for t in tabs]
    out_l = [l.copy() for l in lins]
    acc_t = [np.full_like(t, acc0) for t in tabs]
    acc_l = [np.full_like(l, acc0) for l in lins]
    for f in range(F):
        V = tabs[f].shape[0]
        for i in np.unique(idx[:, f]):
            if i < 0 or i >= V:
                continue
            g = np.zeros((dim,), f32)
            gl = f32(0)
            for b in np.nonzero(idx[:, f] == i)[0]:
                r = dx[b, f * dim:(f + 1) * dim].astype(f32)
                if dfm is not None and (fm_mask >> f) & 1:
                    r = (r + (f32(dfm[b]) * (S[b] - x[b, f * dim:(f + 1) * dim]).astype(f32)).astype(f32)).astype(f32)
                g = (g + r).astype(f32)
                gl = f32(gl + f32(dlin[b]))
            if adagrad:
                acc_t[f][i] = (acc_t[f][i] + (g * g).astype(f32)).astype(f32)
                out_t[f][i] = out_t[f][i] - f32(lr) * g / (np.sqrt(acc_t[f][i]) + f32(eps))
                acc_l[f][i] = f32(acc_l[f][i] + f32(gl * gl))
                out_l[f][i] = out_l[f][i] - f32(lr) * gl / (np.sqrt(acc_l[f][i]) + f32(eps))
            else:
                out_t[f][i] = (out_t[f][i] - (f32(lr) * g).astype(f32)).astype(f32)
                out_l[f][i] = f32(out_l[f][i] - f32(f32(lr) * gl))
    return out_t, out_l


@pytest.mark.parametrize("dim,F", [(8, 5), (32, 26), (128, 3)])
@pytest.mark.parametrize("with_fm", [True, False])
def test_sorted_sgd_update_is_bit_exact_and_deterministic(cuda, dim, F, with_fm):
    from deepctr_b200 import kernels as K
    rng = np.random.RandomState(31)
    B, V, lr = 257, 23, 0.05
    tabs, lins, idx, x, dx, dfm, dlin, ldx = _setup(rng, B, F, V, dim, cuda, with_fm)
    fm_mask = (1 << F) - 1 if with_fm else 0
    want_t, want_l = _want(tabs, lins, idx, x, dx, dfm, dlin, dim, lr, fm_mask)
    results = []
    for rep in range(2):
        td = [torch.tensor(t).to(cuda) for t in tabs]
        ld_ = [torch.tensor(l).to(cuda) for l in lins]
        xd, idd = torch.tensor(x).to(cuda), torch.tensor(idx).to(cuda)
        feats = [K.make_feature(td[f], idd[:, f], xd) for f in range(F)]
        plan = K.UniformPlan(feats, ld_, None, xd, None, None, fm_mask)
        K.embed_update_sorted(plan, torch.tensor(dx).to(cuda), torch.tensor(dfm).to(cuda) if with_fm else None,
                              torch.tensor(dlin).to(cuda), 0, lr, lr, 1e-7, None, None, B)
        results.append(([t.cpu().numpy() for t in td], [l.cpu().numpy() for l in ld_]))
    for f in range(F):
        assert np.array_equal(results[0][0][f], results[1][0][f])          # run-to-run bit-identical
        assert np.array_equal(results[0][1][f], results[1][1][f])
        assert np.array_equal(results[0][0][f], want_t[f]), "table %d" % f  # ordered fp32 segment sums: bit-exact
        assert np.array_equal(results[0][1][f], want_l[f]), "linear %d" % f


def test_sorted_adagrad_update(cuda):
    from deepctr_b200 import kernels as K
    rng = np.random.RandomState(32)
    B, F, V, dim, lr = 300, 6, 19, 16, 0.1
    tabs, lins, idx, x, dx, dfm, dlin, ldx = _setup(rng, B, F, V, dim, cuda, True)
    fm_mask = (1 << F) - 1
    want_t, want_l = _want(tabs, lins, idx, x, dx, dfm, dlin, dim, lr, fm_mask, adagrad=True)
    td = [torch.tensor(t).to(cuda) for t in tabs]
    ld_ = [torch.tensor(l).to(cuda) for l in lins]
    acc = [torch.full_like(t, 0.1) for t in td]
    lacc = [torch.full_like(l, 0.1) for l in ld_]
    xd, idd = torch.tensor(x).to(cuda), torch.tensor(idx).to(cuda)
    feats = [K.make_feature(td[f], idd[:, f], xd) for f in range(F)]
    plan = K.UniformPlan(feats, ld_, None, xd, None, None, fm_mask)
    K.embed_update_sorted(plan, torch.tensor(dx).to(cuda), torch.tensor(dfm).to(cuda), torch.tensor(dlin).to(cuda),
                          1, lr, lr, 1e-7, acc, lacc, B)
    for f in range(F):
        np.testing.assert_allclose(td[f].cpu().numpy(), want_t[f], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(ld_[f].cpu().numpy(), want_l[f], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("opt_name", ["sgd", "adagrad"])
def test_deepfm_deterministic_sparse_training(cuda, opt_name):
    """model level: embedding_update='sparse_deterministic' (SGD) / row-wise Adagrad vs the CPU oracle trained with
    torch.optim on dense gradients (Keras' sparse Adagrad is lazy: rows outside the batch do not move)."""
    from deepctr_b200 import engine as E, ops
    from deepctr_b200.engine import SGD, Adagrad
    from deepctr_b200.models import DeepFM
    ops.set_gemm_precision("fp32")
    rng = np.random.RandomState(33)
    cols, x, y = H.criteo_like(rng, 96, dim=8)
    E.clear_session()
    model = DeepFM(cols, cols, dnn_hidden_units=(16, 8), l2_reg_linear=0, l2_reg_embedding=0)
    H.randomize_weights(model, rng)
    lr = 0.05
    model.compile(SGD(lr) if opt_name == "sgd" else Adagrad(lr), "binary_crossentropy",
                  embedding_update="sparse_deterministic" if opt_name == "sgd" else "sparse")
    assert model.planner.sorted_update
    W = H.oracle_weights(model, requires_grad=True)
    params = list(H.flat_params(W).values())
    topt = (torch.optim.SGD(params, lr=lr) if opt_name == "sgd" else
            torch.optim.Adagrad(params, lr=lr, initial_accumulator_value=0.1, eps=1e-7))
    for step in range(3):
        topt.zero_grad()
        logit, pred = OM.deepfm(x, cols, cols, W)
        loss = O.binary_crossentropy(y, pred)
        loss.backward()
        got = model.train_on_batch(x, y)
        assert abs(got - float(loss.detach())) < 2e-5 * max(1.0, abs(float(loss.detach()))), (step, got, float(loss.detach()))
        topt.step()
    new = H.flat_params(H.oracle_weights(model))
    for name, p in H.flat_params(W).items():
        np.testing.assert_allclose(new[name].numpy(), p.detach().numpy(), rtol=2e-4, atol=2e-6, err_msg=name)
    ops.set_gemm_precision("bf16x3")
