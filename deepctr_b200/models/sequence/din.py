"""Deep Interest Network (Zhou et al. 2018) - drop-in for the reference builder
deepctr/models/sequence/din.py:20-96.

`history_feature_list` names the query features (e.g. ["item_id", "cate_id"]); the behaviour sequences are the
VarLen columns called "hist_" + name and must share their table with the query column through `embedding_name`
(examples/run_din.py:12-15).  The candidate attends over its own history (local activation unit), the attended
history joins the other embeddings and the dense features in front of the DNN tower."""
from ...engine import Dense, Flatten, Model
from ...feature_column import DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features
from ... import inputs as I
from ...layers.core import DNN, PredictionLayer
from ...layers.sequence import AttentionSequencePoolingLayer
from ...layers.utils import combined_dnn_input, concat_func


def _of_type(columns, kind):
    return [c for c in (columns or []) if isinstance(c, kind)]


def DIN(dnn_feature_columns, history_feature_list, dnn_use_bn=False, dnn_hidden_units=(256, 128, 64),
        dnn_activation='relu', att_hidden_size=(80, 40), att_activation="dice", att_weight_normalization=False,
        l2_reg_dnn=0, l2_reg_embedding=1e-6, dnn_dropout=0, seed=1024, task='binary'):
    features = build_input_features(dnn_feature_columns)
    singles = _of_type(dnn_feature_columns, SparseFeat)
    sequences = _of_type(dnn_feature_columns, VarLenSparseFeat)
    behaviour_names = ["hist_" + name for name in history_feature_list]
    behaviours = [c for c in sequences if c.name in behaviour_names]
    pooled_sequences = [c for c in sequences if c.name not in behaviour_names]

    tables = I.create_embedding_matrix(dnn_feature_columns, l2_reg_embedding, seed, prefix="")
    # candidate (query) and behaviour (keys) embeddings keep their Keras masks: the attention needs them
    query = I.embedding_lookup(tables, features, singles, history_feature_list, history_feature_list, to_list=True)
    keys = I.embedding_lookup(tables, features, behaviours, behaviour_names, behaviour_names, to_list=True)
    context = I.embedding_lookup(tables, features, singles, mask_feat_list=history_feature_list, to_list=True)
    dense_values = I.get_dense_input(features, _of_type(dnn_feature_columns, DenseFeat))
    other_seq = I.varlen_embedding_lookup(tables, features, pooled_sequences)
    context = context + list(I.get_varlen_pooling_list(other_seq, features, pooled_sequences, to_list=True))

    keys_matrix = concat_func(keys, mask=True)               # [B, T, sum E]
    context_matrix = concat_func(context)                    # [B, 1, sum E]
    query_matrix = concat_func(query, mask=True)             # [B, 1, sum E]
    interest = AttentionSequencePoolingLayer(att_hidden_size, att_activation,
                                             weight_normalization=att_weight_normalization,
                                             supports_masking=True)([query_matrix, keys_matrix])

    tower_in = combined_dnn_input([Flatten()(concat_func([context_matrix, interest]))], dense_values)
    hidden = DNN(dnn_hidden_units, dnn_activation, l2_reg_dnn, dnn_dropout, dnn_use_bn, seed=seed)(tower_in)
    logit = Dense(1, use_bias=False)(hidden)
    return Model(inputs=list(features.values()), outputs=PredictionLayer(task)(logit))
