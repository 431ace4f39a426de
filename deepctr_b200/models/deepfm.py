"""DeepFM builder - drop-in for deepctr/models/deepfm.py:22-65 (same signature and defaults).
logit = linear + DNN tower + sum over fm_group of FM(group embeddings);  out = PredictionLayer(task)."""
from itertools import chain

from ..engine import Model, Dense
from ..feature_column import build_input_features, get_linear_logit, DEFAULT_GROUP_NAME, input_from_feature_columns
from ..layers.core import PredictionLayer, DNN
from ..layers.interaction import FM
from ..layers.utils import concat_func, add_func, combined_dnn_input


def DeepFM(linear_feature_columns, dnn_feature_columns, fm_group=(DEFAULT_GROUP_NAME,),
           dnn_hidden_units=(256, 128, 64), l2_reg_linear=0.00001, l2_reg_embedding=0.00001, l2_reg_dnn=0,
           seed=1024, dnn_dropout=0, dnn_activation='relu', dnn_use_bn=False, task='binary'):
    features = build_input_features(linear_feature_columns + dnn_feature_columns)
    inputs_list = list(features.values())

    linear_logit = get_linear_logit(features, linear_feature_columns, seed=seed, prefix='linear',
                                    l2_reg=l2_reg_linear)
    group_embedding_dict, dense_value_list = input_from_feature_columns(
        features, dnn_feature_columns, l2_reg_embedding, seed, support_group=True)

    fm_logits = [FM()(concat_func(embs, axis=1))
                 for group, embs in group_embedding_dict.items() if group in fm_group]

    all_embs = list(chain.from_iterable(group_embedding_dict.values()))
    dnn_out = DNN(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn,
                  seed=seed)(combined_dnn_input(all_embs, dense_value_list))
    dnn_logit = Dense(1, use_bias=False)(dnn_out)

    output = PredictionLayer(task)(add_func([linear_logit, dnn_logit] + fm_logits))
    return Model(inputs=inputs_list, outputs=output)
