"""AutoInt (Song et al. 2019) - drop-in for the reference builder deepctr/models/autoint.py:21-84: stacked
multi-head self-attention over the field embeddings next to (or instead of) the DNN tower.  Inputs come
from `dnn_feature_columns` only (:49)."""
from ..engine import Concatenate, Flatten
from ..layers.interaction import InteractingLayer
from ._tower import Tower, total


def AutoInt(linear_feature_columns, dnn_feature_columns, att_layer_num=3, att_embedding_size=8, att_head_num=2,
            att_res=True, dnn_hidden_units=(256, 128, 64), dnn_activation='relu', l2_reg_linear=1e-5,
            l2_reg_embedding=1e-5, l2_reg_dnn=0, dnn_use_bn=False, dnn_dropout=0, seed=1024, task='binary'):
    has_deep, has_att = len(dnn_hidden_units) > 0, att_layer_num > 0
    if not has_deep and not has_att:
        raise ValueError("Either hidden_layer or att_layer_num must > 0")
    t = Tower(dnn_feature_columns, linear_feature_columns, dnn_feature_columns, seed, l2_reg_linear,
              l2_reg_embedding)
    fields = t.field_matrix()
    for _ in range(att_layer_num):
        fields = InteractingLayer(att_embedding_size, att_head_num, att_res)(fields)
    attended = Flatten()(fields)
    x = t.flat_input()
    branches = [attended] if has_att else []
    if has_deep:
        branches.append(t.mlp(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, inputs=x))
    stack = Concatenate()(branches) if len(branches) == 2 else branches[0]
    return t.finish(total([t.project(stack), t.linear_logit]), task)
