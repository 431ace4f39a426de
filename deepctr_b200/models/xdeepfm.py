"""xDeepFM builder - drop-in for deepctr/models/xdeepfm.py:18-70.
logit = linear + DNN tower (+ Dense1(CIN(field embeddings)) when cin_layer_size is non-empty)."""
from ..engine import Model, Dense
from ..feature_column import build_input_features, get_linear_logit, input_from_feature_columns
from ..layers.core import PredictionLayer, DNN
from ..layers.interaction import CIN
from ..layers.utils import concat_func, add_func, combined_dnn_input


def xDeepFM(linear_feature_columns, dnn_feature_columns, dnn_hidden_units=(256, 128, 64),
            cin_layer_size=(128, 128,), cin_split_half=True, cin_activation='relu', l2_reg_linear=0.00001,
            l2_reg_embedding=0.00001, l2_reg_dnn=0, l2_reg_cin=0, seed=1024, dnn_dropout=0,
            dnn_activation='relu', dnn_use_bn=False, task='binary'):
    features = build_input_features(linear_feature_columns + dnn_feature_columns)
    inputs_list = list(features.values())

    linear_logit = get_linear_logit(features, linear_feature_columns, seed=seed, prefix='linear',
                                    l2_reg=l2_reg_linear)
    emb_list, dense_value_list = input_from_feature_columns(features, dnn_feature_columns,
                                                            l2_reg_embedding, seed)
    field_matrix = concat_func(emb_list, axis=1)           # [B, m, D]

    dnn_out = DNN(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn,
                  seed=seed)(combined_dnn_input(emb_list, dense_value_list))
    final_logit = add_func([linear_logit, Dense(1, use_bias=False)(dnn_out)])

    if len(cin_layer_size) > 0:
        cin_out = CIN(cin_layer_size, cin_activation, cin_split_half, l2_reg_cin, seed)(field_matrix)
        final_logit = add_func([final_logit, Dense(1, use_bias=False)(cin_out)])

    return Model(inputs=inputs_list, outputs=PredictionLayer(task)(final_logit))
