"""Deep & Cross Network builder - drop-in for deepctr/models/dcn.py:22-78.
Model inputs come from dnn_feature_columns only (dcn.py:48), as in the reference."""
from ..engine import Model, Dense, Concatenate
from ..feature_column import build_input_features, get_linear_logit, input_from_feature_columns
from ..layers.core import PredictionLayer, DNN
from ..layers.interaction import CrossNet
from ..layers.utils import add_func, combined_dnn_input


def DCN(linear_feature_columns, dnn_feature_columns, cross_num=2, cross_parameterization='vector',
        dnn_hidden_units=(256, 128, 64), l2_reg_linear=1e-5, l2_reg_embedding=1e-5, l2_reg_cross=1e-5,
        l2_reg_dnn=0, seed=1024, dnn_dropout=0, dnn_use_bn=False, dnn_activation='relu', task='binary'):
    if len(dnn_hidden_units) == 0 and cross_num == 0:
        raise ValueError("Either hidden_layer or cross layer must > 0")

    features = build_input_features(dnn_feature_columns)
    inputs_list = list(features.values())
    linear_logit = get_linear_logit(features, linear_feature_columns, seed=seed, prefix='linear',
                                    l2_reg=l2_reg_linear)
    emb_list, dense_value_list = input_from_feature_columns(features, dnn_feature_columns,
                                                            l2_reg_embedding, seed)
    dnn_input = combined_dnn_input(emb_list, dense_value_list)

    def deep():
        return DNN(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, seed=seed)(dnn_input)

    def cross():
        return CrossNet(cross_num, parameterization=cross_parameterization, l2_reg=l2_reg_cross)(dnn_input)

    if len(dnn_hidden_units) > 0 and cross_num > 0:
        deep_out = deep()
        stack = Concatenate()([cross(), deep_out])
    elif len(dnn_hidden_units) > 0:
        stack = deep()
    elif cross_num > 0:
        stack = cross()
    else:
        raise NotImplementedError
    final_logit = add_func([Dense(1, use_bias=False)(stack), linear_logit])
    return Model(inputs=inputs_list, outputs=PredictionLayer(task)(final_logit))
