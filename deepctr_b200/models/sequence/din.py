"""Deep Interest Network builder - drop-in for deepctr/models/sequence/din.py:20-96.
History columns are the VarLen columns named "hist_" + f for f in history_feature_list and must share
tables with their query feature through embedding_name (examples/run_din.py:12-15)."""
from ...engine import Model, Dense, Flatten
from ...feature_column import SparseFeat, VarLenSparseFeat, DenseFeat, build_input_features
from ...inputs import (create_embedding_matrix, embedding_lookup, get_dense_input, varlen_embedding_lookup,
                       get_varlen_pooling_list)
from ...layers.core import DNN, PredictionLayer
from ...layers.sequence import AttentionSequencePoolingLayer
from ...layers.utils import concat_func, combined_dnn_input


def DIN(dnn_feature_columns, history_feature_list, dnn_use_bn=False, dnn_hidden_units=(256, 128, 64),
        dnn_activation='relu', att_hidden_size=(80, 40), att_activation="dice", att_weight_normalization=False,
        l2_reg_dnn=0, l2_reg_embedding=1e-6, dnn_dropout=0, seed=1024, task='binary'):
    features = build_input_features(dnn_feature_columns)
    cols = dnn_feature_columns or []
    sparse_cols = [c for c in cols if isinstance(c, SparseFeat)]
    dense_cols = [c for c in cols if isinstance(c, DenseFeat)]
    varlen_cols = [c for c in cols if isinstance(c, VarLenSparseFeat)]

    hist_names = ["hist_" + f for f in history_feature_list]
    history_cols = [c for c in varlen_cols if c.name in hist_names]
    other_varlen_cols = [c for c in varlen_cols if c.name not in hist_names]
    inputs_list = list(features.values())

    embedding_dict = create_embedding_matrix(dnn_feature_columns, l2_reg_embedding, seed, prefix="")
    query_embs = embedding_lookup(embedding_dict, features, sparse_cols, history_feature_list,
                                  history_feature_list, to_list=True)
    key_embs = embedding_lookup(embedding_dict, features, history_cols, hist_names, hist_names, to_list=True)
    deep_embs = embedding_lookup(embedding_dict, features, sparse_cols, mask_feat_list=history_feature_list,
                                 to_list=True)
    dense_value_list = get_dense_input(features, dense_cols)
    seq_embed_dict = varlen_embedding_lookup(embedding_dict, features, other_varlen_cols)
    deep_embs += list(get_varlen_pooling_list(seq_embed_dict, features, other_varlen_cols, to_list=True))

    keys_emb = concat_func(key_embs, mask=True)
    deep_input_emb = concat_func(deep_embs)
    query_emb = concat_func(query_embs, mask=True)
    hist = AttentionSequencePoolingLayer(att_hidden_size, att_activation,
                                         weight_normalization=att_weight_normalization,
                                         supports_masking=True)([query_emb, keys_emb])

    deep_input_emb = Flatten()(concat_func([deep_input_emb, hist]))
    dnn_input = combined_dnn_input([deep_input_emb], dense_value_list)
    tower = DNN(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, seed=seed)(dnn_input)
    output = PredictionLayer(task)(Dense(1, use_bias=False)(tower))
    return Model(inputs=inputs_list, outputs=output)
