"""GPU: the reference's example script examples/run_din.py as an acceptance test: the script's
own data, DIN with its defaults (Dice attention units), compile('adam'), fit(epochs=10, validation_split=0.5); then
the trained model's predictions against the CPU oracle on the trained weights."""
import numpy as np
import pytest

import b2_helpers as H
from oracle import models as OM

pytestmark = pytest.mark.gpu


def _get_xy_fd():
    """the data of examples/run_din.py:8-34"""
    from deepctr_b200.feature_column import SparseFeat, VarLenSparseFeat, DenseFeat, get_feature_names
    feature_columns = [SparseFeat('user', 3, embedding_dim=10), SparseFeat('gender', 2, embedding_dim=4),
                       SparseFeat('item_id', 3 + 1, embedding_dim=8), SparseFeat('cate_id', 2 + 1, embedding_dim=4),
                       DenseFeat('pay_score', 1)]
    feature_columns += [
        VarLenSparseFeat(SparseFeat('hist_item_id', vocabulary_size=3 + 1, embedding_dim=8, embedding_name='item_id'),
                         maxlen=4, length_name="seq_length"),
        VarLenSparseFeat(SparseFeat('hist_cate_id', 2 + 1, embedding_dim=4, embedding_name='cate_id'), maxlen=4,
                         length_name="seq_length")]
    behavior_feature_list = ["item_id", "cate_id"]
    feature_dict = {'user': np.array([0, 1, 2]), 'gender': np.array([0, 1, 0]), 'item_id': np.array([1, 2, 3]),
                    'cate_id': np.array([1, 2, 2]), 'pay_score': np.array([0.1, 0.2, 0.3]),
                    'hist_item_id': np.array([[1, 2, 3, 0], [3, 2, 1, 0], [1, 2, 0, 0]]),
                    'hist_cate_id': np.array([[1, 2, 2, 0], [2, 2, 1, 0], [1, 2, 0, 0]]),
                    'seq_length': np.array([3, 3, 2])}
    x = {name: feature_dict[name] for name in get_feature_names(feature_columns)}
    return x, np.array([1, 0, 1]), feature_columns, behavior_feature_list


def test_run_din_example(cuda):
    import warnings
    from deepctr_b200.models import DIN
    x, y, feature_columns, behavior_feature_list = _get_xy_fd()
    model = DIN(feature_columns, behavior_feature_list)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")            # metrics=[...] is accepted and ignored with a warning
        model.compile('adam', 'binary_crossentropy', metrics=['binary_crossentropy'])
        history = model.fit(x, y, verbose=0, epochs=10, validation_split=0.5)
    assert len(history.history['loss']) == 10 and len(history.history['val_loss']) == 10
    assert np.all(np.isfinite(history.history['loss'])) and np.all(np.isfinite(history.history['val_loss']))
    pred = model.predict(x, batch_size=256)
    assert pred.shape == (3, 1) and np.all((pred > 0) & (pred < 1))
    # the trained model against the oracle on the trained weights (inference: Dice uses its moving statistics)
    W = H.oracle_weights(model)
    _, want = OM.din(x, feature_columns, behavior_feature_list, W, "dice", False, training=False)
    assert H.rel_err(pred, want.numpy()) < 1e-4
