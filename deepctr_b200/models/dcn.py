"""Deep & Cross Network (Wang et al. 2017 / 2020) - drop-in for the reference builder
deepctr/models/dcn.py:22-78.  As there, the model's inputs are those of `dnn_feature_columns` only (:48)."""
from ..engine import Concatenate
from ..layers.interaction import CrossNet
from ._tower import Tower, total


def DCN(linear_feature_columns, dnn_feature_columns, cross_num=2, cross_parameterization='vector',
        dnn_hidden_units=(256, 128, 64), l2_reg_linear=1e-5, l2_reg_embedding=1e-5, l2_reg_cross=1e-5,
        l2_reg_dnn=0, seed=1024, dnn_dropout=0, dnn_use_bn=False, dnn_activation='relu', task='binary'):
    has_deep, has_cross = len(dnn_hidden_units) > 0, cross_num > 0
    if not has_deep and not has_cross:
        raise ValueError("Either hidden_layer or cross layer must > 0")
    t = Tower(dnn_feature_columns, linear_feature_columns, dnn_feature_columns, seed, l2_reg_linear,
              l2_reg_embedding)
    x0 = t.flat_input()
    branches = []
    if has_deep:
        branches.append(t.mlp(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn))
    if has_cross:
        branches.insert(0, CrossNet(cross_num, parameterization=cross_parameterization, l2_reg=l2_reg_cross)(x0))
    stack = Concatenate()(branches) if len(branches) == 2 else branches[0]
    return t.finish(total([t.project(stack), t.linear_logit]), task)
