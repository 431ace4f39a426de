"""xDeepFM (Lian et al. 2018) - drop-in for the reference builder deepctr/models/xdeepfm.py:18-70.
logit = first-order term + DNN tower, plus a projection of the Compressed Interaction Network over the
[B, fields, E] embedding matrix when `cin_layer_size` is not empty."""
from ..layers.interaction import CIN
from ._tower import Tower, total


def xDeepFM(linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 128, 64),
            cin_layer_size=(128, 128,), cin_split_half=True, cin_activation='relu', l2_reg_linear=0.00001,
            l2_reg_embedding=0.00001, l2_reg_dnn=0, l2_reg_cin=0, seed=1024, dnn_dropout=0,
            dnn_activation='relu', dnn_use_bn=False, task='binary'):
    t = Tower(linear_feature_columns + dnn_feature_columns, linear_feature_columns, dnn_feature_columns, seed,
              l2_reg_linear, l2_reg_embedding)
    fields = t.field_matrix()
    deep = t.project(t.mlp(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn))
    logit = total([t.linear_logit, deep])
    if len(cin_layer_size) > 0:
        explicit = CIN(cin_layer_size, cin_activation, cin_split_half, l2_reg_cin, seed)(fields)
        logit = total([logit, t.project(explicit)])
    return t.finish(logit, task)
