"""DeepFM (Guo et al. 2017) - drop-in for the reference builder deepctr/models/deepfm.py:22-65.
logit = first-order term + DNN tower + sum over `fm_group` of FM(embeddings of that group)."""
from ..feature_column import DEFAULT_GROUP_NAME
from ..layers.interaction import FM
from ._tower import Tower, total


def DeepFM(linear_feature_columns, dnn_feature_columns, fm_group=(DEFAULT_GROUP_NAME,),
           dnn_hidden_units=(256, 128, 64), l2_reg_linear=0.00001, l2_reg_embedding=0.00001, l2_reg_dnn=0,
           seed=1024, dnn_dropout=0, dnn_activation='relu', dnn_use_bn=False, task='binary'):
    t = Tower(linear_feature_columns + dnn_feature_columns, linear_feature_columns, dnn_feature_columns, seed,
              l2_reg_linear, l2_reg_embedding, grouped=True)
    second_order = [FM()(t.field_matrix(embs)) for name, embs in t.groups.items() if name in fm_group]
    deep = t.project(t.mlp(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn))
    return t.finish(total([t.linear_logit, deep] + second_order), task)
