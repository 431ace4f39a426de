"""GPU: BASELINE.json configs[0] - the reference's own example (examples/run_classification_criteo.py)
on its own sample data (tests/golden/criteo_sample.txt, a verbatim copy of the reference's 200-row data
fixture): pandas preprocessing -> DeepFM(26 sparse + 13 dense, embedding_dim=8) -> compile("adam",
"binary_crossentropy") -> fit(batch_size=256) -> predict, against the CPU oracle trained with the same
Keras semantics (dense Adam over every table, L2 1e-5 on embeddings and the linear part)."""
import os

import numpy as np
import pytest
import torch

from oracle import models as OM
from oracle import ops as O
import b2_helpers as H

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gemm_precision")]
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "criteo_sample.txt")


def _prepare():
    import pandas as pd
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import LabelEncoder, MinMaxScaler
    from deepctr_b200.feature_column import SparseFeat, DenseFeat, get_feature_names
    data = pd.read_csv(DATA)
    sparse = ['C' + str(i) for i in range(1, 27)]
    dense = ['I' + str(i) for i in range(1, 14)]
    data[sparse] = data[sparse].fillna('-1')
    data[dense] = data[dense].fillna(0)
    for feat in sparse:
        data[feat] = LabelEncoder().fit_transform(data[feat])
    data[dense] = MinMaxScaler(feature_range=(0, 1)).fit_transform(data[dense])
    cols = [SparseFeat(f, vocabulary_size=int(data[f].max()) + 1, embedding_dim=8) for f in sparse] + \
           [DenseFeat(f, 1) for f in dense]
    names = get_feature_names(cols + cols)
    train, test = train_test_split(data, test_size=0.2, random_state=2020)
    return cols, names, train, test


def test_preprocessing_matches_example_script():
    cols, names, train, test = _prepare()
    card = [c.vocabulary_size for c in cols[:26]]
    # LabelEncoder ids are dense: vocabulary_size = max + 1 = number of distinct values incl. the '-1' fill
    assert card == [27, 92, 172, 157, 12, 7, 183, 19, 2, 142, 173, 170, 166, 14, 170, 168, 9, 127, 44, 4, 169, 6,
                    10, 125, 20, 90]
    assert len(train) == 160 and len(test) == 40 and len(names) == 39
    assert abs(float(train['label'].mean()) - 0.245) < 0.05


def test_deepfm_criteo_sample_training_matches_oracle(cuda):
    from deepctr_b200.models import DeepFM
    cols, names, train, test = _prepare()
    x_tr = {n: train[n] for n in names}                  # pandas Series, as in the example script
    y_tr = train[['label']].values
    x_te = {n: test[n] for n in names}
    model = DeepFM(cols, cols, task='binary')            # reference defaults: l2 1e-5, dnn (256,128,64)
    rng = np.random.RandomState(0)
    H.randomize_weights(model, rng, 0.05)                # the default 1e-4 / zero inits carry no signal
    model.compile("adam", "binary_crossentropy", metrics=['binary_crossentropy'], embedding_update="dense")

    # ---- oracle: same weights, same batches, Keras Adam with L2 regularisers ----
    W = H.oracle_weights(model, requires_grad=True)
    leaves = H.flat_params(W)
    l2 = {n: (1e-5 if (n.startswith("tables/") or n == "linear_kernel") else 0.0) for n in leaves}
    m = {n: torch.zeros_like(t) for n, t in leaves.items()}
    v = {n: torch.zeros_like(t) for n, t in leaves.items()}
    # before any training: logits-level parity at the north_star tolerance (1e-4) on the held-out rows
    x_te_np = {n: np.asarray(x_te[n]) for n in names}
    with torch.no_grad():
        _, want0 = OM.deepfm(x_te_np, cols, cols, W)
    assert H.rel_err(model.predict(x_te, batch_size=256), want0.numpy()) < 1e-4
    split = int(len(y_tr) * 0.8)                         # validation_split=0.2 takes the LAST 20 %
    xo = {n: np.asarray(x_tr[n])[:split] for n in names}
    yo = y_tr[:split].reshape(-1)
    epochs, want_losses = 6, []
    for step in range(1, epochs + 1):
        _, pred = OM.deepfm(xo, cols, cols, W)
        data_loss = O.binary_crossentropy(yo, pred)
        data_loss.backward()
        reg = sum(l2[n] * float((t.detach().double() ** 2).sum()) for n, t in leaves.items())
        want_losses.append(float(data_loss) + reg)
        with torch.no_grad():
            lr_t = 1e-3 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
            for n, t in leaves.items():
                if t.grad is None:
                    continue
                g = t.grad + 2 * l2[n] * t
                m[n] = 0.9 * m[n] + 0.1 * g
                v[n] = 0.999 * v[n] + 0.001 * g * g
                t -= lr_t * m[n] / (v[n].sqrt() + 1e-7)
                t.grad = None
    hist = model.fit(x_tr, y_tr, batch_size=256, epochs=epochs, verbose=0, validation_split=0.2, shuffle=False)
    got = hist.history["loss"]
    # Keras reports the loss BEFORE the update of that step, regularisation included; fit() adds the
    # regulariser evaluated after the epoch's update: compare the data term trajectory step by step
    assert len(got) == epochs and len(hist.history["val_loss"]) == epochs
    for a, b in zip(got, want_losses):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (got, want_losses)
    assert got[-1] < got[0]
    # predictions on the held-out 40 rows after 6 dense-Adam steps (fp32 Adam amplifies last-bit differences of the
    # gradients through 1/sqrt(v): the trajectories are compared at 2e-3, the untrained model above at 1e-4)
    pred = model.predict(x_te, batch_size=256)
    _, want = OM.deepfm(x_te_np, cols, cols, W)
    assert pred.shape == (40, 1)
    assert H.rel_err(pred, want.detach().numpy()) < 2e-3
    from sklearn.metrics import log_loss, roc_auc_score
    y_te = test[['label']].values
    assert np.isfinite(log_loss(y_te, pred.astype(np.float64)))
    assert 0.0 <= roc_auc_score(y_te, pred) <= 1.0
