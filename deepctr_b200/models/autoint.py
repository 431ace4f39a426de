"""AutoInt builder - drop-in for deepctr/models/autoint.py:21-84."""
from ..engine import Model, Dense, Concatenate, Flatten
from ..feature_column import build_input_features, get_linear_logit, input_from_feature_columns
from ..layers.core import PredictionLayer, DNN
from ..layers.interaction import InteractingLayer
from ..layers.utils import concat_func, add_func, combined_dnn_input


def AutoInt(linear_feature_columns, dnn_feature_columns, att_layer_num=3, att_embedding_size=8, att_head_num=2,
            att_res=True, dnn_hidden_units=(256, 128, 64), dnn_activation='relu', l2_reg_linear=1e-5,
            l2_reg_embedding=1e-5, l2_reg_dnn=0, dnn_use_bn=False, dnn_dropout=0, seed=1024, task='binary'):
    if len(dnn_hidden_units) <= 0 and att_layer_num <= 0:
        raise ValueError("Either hidden_layer or att_layer_num must > 0")

    features = build_input_features(dnn_feature_columns)
    inputs_list = list(features.values())
    linear_logit = get_linear_logit(features, linear_feature_columns, seed=seed, prefix='linear',
                                    l2_reg=l2_reg_linear)
    emb_list, dense_value_list = input_from_feature_columns(features, dnn_feature_columns,
                                                            l2_reg_embedding, seed)
    att = concat_func(emb_list, axis=1)
    for _ in range(att_layer_num):
        att = InteractingLayer(att_embedding_size, att_head_num, att_res)(att)
    att_output = Flatten()(att)
    dnn_input = combined_dnn_input(emb_list, dense_value_list)

    if len(dnn_hidden_units) > 0 and att_layer_num > 0:
        deep_out = DNN(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, seed=seed)(dnn_input)
        final_logit = Dense(1, use_bias=False)(Concatenate()([att_output, deep_out]))
    elif len(dnn_hidden_units) > 0:
        deep_out = DNN(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, seed=seed)(dnn_input)
        final_logit = Dense(1, use_bias=False)(deep_out)
    elif att_layer_num > 0:
        final_logit = Dense(1, use_bias=False)(att_output)
    else:
        raise NotImplementedError
    final_logit = add_func([final_logit, linear_logit])
    return Model(inputs=inputs_list, outputs=PredictionLayer(task)(final_logit))
