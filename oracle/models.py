"""TEST INFRASTRUCTURE - CPU oracle of the five builders (torch-CPU fp32), composed from oracle/ops.py.

Follows deepctr/feature_column.py:171-233 (get_linear_logit / input_from_feature_columns),
deepctr/inputs.py:101-158 and the builder bodies deepctr/models/{deepfm,xdeepfm,dcn,autoint}.py,
deepctr/models/sequence/din.py.  Feature columns are duck-typed (any object with the reference's
SparseFeat / VarLenSparseFeat / DenseFeat attributes); weights arrive as a plain dict
``{weight_name: torch tensor}`` using the reference's naming:
``<prefix>sparse_emb_<embedding_name>`` / ``..._seq_emb_...`` for tables, and per-layer dicts.
"""
from collections import OrderedDict, defaultdict
from itertools import chain

import numpy as np
import torch

from . import ops as O


def _kind(fc):
    if hasattr(fc, "sparsefeat"):
        return "varlen"
    if hasattr(fc, "vocabulary_size"):
        return "sparse"
    return "dense"


def table_name(prefix, fc, varlen_names, is_varlen_owner):
    suffix = "seq_emb" if is_varlen_owner else "emb"
    return prefix + "sparse_" + suffix + "_" + fc.embedding_name


def table_names(feature_columns, prefix=""):
    """embedding_name -> layer name, following create_embedding_dict (inputs.py:44-71)."""
    names = OrderedDict()
    sparse = [c for c in feature_columns if _kind(c) == "sparse"]
    varlen = [c for c in feature_columns if _kind(c) == "varlen"]
    for fc in sparse:
        if fc.embedding_name not in names:
            names[fc.embedding_name] = prefix + "sparse_emb_" + fc.embedding_name
    for fc in varlen:
        if fc.embedding_name not in names:
            names[fc.embedding_name] = prefix + "sparse_seq_emb_" + fc.embedding_name
    return names


def mask_zero_tables(feature_columns, seq_mask_zero=True):
    varlen_names = set(c.embedding_name for c in feature_columns if _kind(c) == "varlen")
    return varlen_names if seq_mask_zero else set()


def _ids(fc, x, mask_zero):
    a = np.asarray(x)
    if fc.use_hash:
        vocab = O.load_vocabulary(fc.vocabulary_path) if fc.vocabulary_path else None
        a = O.hash_layer(a, fc.vocabulary_size, mask_zero, vocab)
    return torch.as_tensor(a.astype(np.int64))


def input_from_feature_columns(inputs, feature_columns, tables, prefix="", support_group=False,
                               mask_feat_list=()):
    """-> (group dict | flat list of [B,1,E], dense list).  ``tables``: layer name -> [V,E] tensor."""
    names = table_names(feature_columns, prefix)
    mz = mask_zero_tables(feature_columns)
    groups = defaultdict(list)
    for fc in feature_columns:
        if _kind(fc) != "sparse":
            continue
        idx = _ids(fc, inputs[fc.name], fc.name in mask_feat_list).reshape(-1, 1)
        groups[fc.group_name].append(O.embedding_lookup(tables[names[fc.embedding_name]], idx))
    vgroups = defaultdict(list)
    for fc in feature_columns:
        if _kind(fc) != "varlen":
            continue
        idx = _ids(fc, inputs[fc.name], True)
        seq = O.embedding_lookup(tables[names[fc.embedding_name]], idx)          # [B,T,E]
        keras_mask = (idx != 0) if fc.embedding_name in mz else None
        if fc.length_name is not None:
            lens = torch.as_tensor(np.asarray(inputs[fc.length_name]).reshape(-1).astype(np.int64))
            if fc.weight_name is not None:
                seq = O.weighted_sequence(seq, torch.as_tensor(np.asarray(inputs[fc.weight_name],
                                                                          dtype=np.float32)),
                                          fc.weight_norm, lengths=lens)
            vec = O.sequence_pooling(seq, fc.combiner, lengths=lens)
        else:
            if keras_mask is None:
                raise ValueError("When supports_masking=True,input must support masking")
            if fc.weight_name is not None:
                seq = O.weighted_sequence(seq, torch.as_tensor(np.asarray(inputs[fc.weight_name],
                                                                          dtype=np.float32)),
                                          fc.weight_norm, mask=keras_mask)
            vec = O.sequence_pooling(seq, fc.combiner, mask=keras_mask)
        vgroups[fc.group_name].append(vec)
    merged = defaultdict(list)
    for k, v in groups.items():
        merged[k].extend(v)
    for k, v in vgroups.items():
        merged[k].extend(v)
    dense = [torch.as_tensor(np.asarray(inputs[fc.name], dtype=np.float32)).reshape(-1, fc.dimension)
             for fc in feature_columns if _kind(fc) == "dense"]
    if not support_group:
        return list(chain.from_iterable(merged.values())), dense
    return merged, dense


def linear_logit(inputs, feature_columns, tables, linear_kernel=None, bias=None, prefix="linear0"):
    """get_linear_logit with units=1 (feature_column.py:171-210): dim-1 copies of every column."""

    class _One(object):
        def __init__(self, fc):
            self._fc = fc

        def __getattr__(self, k):
            if k == "embedding_dim":
                return 1
            return getattr(self._fc, k)

    cols = []
    for fc in feature_columns:
        if _kind(fc) == "varlen":
            w = _One(fc)
            w.sparsefeat = fc.sparsefeat
            cols.append(w)
        elif _kind(fc) == "sparse":
            cols.append(_One(fc))
        else:
            cols.append(fc)
    embs, dense = input_from_feature_columns(inputs, cols, tables, prefix=prefix)
    if not embs and not dense:
        return torch.zeros(1, 1)
    sparse_in = torch.cat(embs, dim=-1) if embs else None
    dense_in = torch.cat(dense, dim=-1) if dense else None
    return O.linear(sparse_in, dense_in, linear_kernel, bias)


def combined_dnn_input(embs, dense):
    parts = []
    if embs:
        parts.append(torch.cat(embs, dim=-1).flatten(1))
    if dense:
        parts.append(torch.cat(dense, dim=-1).flatten(1))
    return torch.cat(parts, dim=-1)


def deepfm(inputs, linear_cols, dnn_cols, W, fm_group=("default_group",), task="binary",
           dnn_activation="relu"):
    """W keys: 'tables' {layer name: tensor}, 'linear_kernel', 'dnn_kernels', 'dnn_biases',
    'dense_kernel' [H,1], 'global_bias' [1]."""
    lin = linear_logit(inputs, linear_cols, W["tables"], W.get("linear_kernel"))
    groups, dense = input_from_feature_columns(inputs, dnn_cols, W["tables"], support_group=True)
    fm_logits = [O.fm(torch.cat(v, dim=1)) for k, v in groups.items() if k in fm_group]
    x = combined_dnn_input(list(chain.from_iterable(groups.values())), dense)
    h = O.dnn(x, W["dnn_kernels"], W["dnn_biases"], dnn_activation)
    logit = lin + h @ W["dense_kernel"]
    for f in fm_logits:
        logit = logit + f
    return logit, O.prediction(logit, W.get("global_bias"), task)


def xdeepfm(inputs, linear_cols, dnn_cols, W, cin_layer_size=(128, 128), cin_split_half=True,
            cin_activation="relu", task="binary"):
    lin = linear_logit(inputs, linear_cols, W["tables"], W.get("linear_kernel"))
    embs, dense = input_from_feature_columns(inputs, dnn_cols, W["tables"])
    h = O.dnn(combined_dnn_input(embs, dense), W["dnn_kernels"], W["dnn_biases"], "relu")
    logit = lin + h @ W["dense_kernel"]
    if len(cin_layer_size) > 0:
        c = O.cin(torch.cat(embs, dim=1), W["cin_filters"], W["cin_biases"], cin_layer_size, cin_activation,
                  cin_split_half)
        logit = logit + c @ W["cin_dense_kernel"]
    return logit, O.prediction(logit, W.get("global_bias"), task)


def dcn(inputs, linear_cols, dnn_cols, W, cross_num=2, parameterization="vector", use_dnn=True, task="binary"):
    lin = linear_logit(inputs, linear_cols, W["tables"], W.get("linear_kernel"))
    embs, dense = input_from_feature_columns(inputs, dnn_cols, W["tables"])
    x = combined_dnn_input(embs, dense)
    parts = []
    if cross_num > 0:
        parts.append(O.crossnet(x, W["cross_kernels"], W["cross_biases"], parameterization))
    if use_dnn:
        parts.append(O.dnn(x, W["dnn_kernels"], W["dnn_biases"], "relu"))
    logit = torch.cat(parts, dim=-1) @ W["dense_kernel"] + lin
    return logit, O.prediction(logit, W.get("global_bias"), task)


def autoint(inputs, linear_cols, dnn_cols, W, att_layer_num=3, att_embedding_size=8, att_head_num=2,
            att_res=True, use_dnn=True, task="binary"):
    lin = linear_logit(inputs, linear_cols, W["tables"], W.get("linear_kernel"))
    embs, dense = input_from_feature_columns(inputs, dnn_cols, W["tables"])
    att = torch.cat(embs, dim=1)
    for i in range(att_layer_num):
        lw = W["att"][i]
        att = O.interacting(att, lw["query"], lw["key"], lw["value"], lw.get("res"), att_head_num,
                            att_embedding_size, att_res)
    parts = []
    if att_layer_num > 0:
        parts.append(att.flatten(1))
    if use_dnn:
        parts.append(O.dnn(combined_dnn_input(embs, dense), W["dnn_kernels"], W["dnn_biases"], "relu"))
    logit = torch.cat(parts, dim=-1) @ W["dense_kernel"] + lin
    return logit, O.prediction(logit, W.get("global_bias"), task)


def din(inputs, dnn_cols, history_feature_list, W, att_activation="sigmoid", att_weight_normalization=False,
        task="binary", training=False):
    """deepctr/models/sequence/din.py:43-96.  W['lau'] = dict(dnn_kernels, dnn_biases, kernel, bias[, act_params])."""
    tables = W["tables"]
    names = table_names(dnn_cols, "")
    sparse = [c for c in dnn_cols if _kind(c) == "sparse"]
    dense_cols = [c for c in dnn_cols if _kind(c) == "dense"]
    varlen = [c for c in dnn_cols if _kind(c) == "varlen"]
    hist_names = ["hist_" + f for f in history_feature_list]
    history = [c for c in varlen if c.name in hist_names]
    others = [c for c in varlen if c.name not in hist_names]
    mz = mask_zero_tables(dnn_cols)

    def look(fc, mask_zero_hash):
        idx = _ids(fc, inputs[fc.name], mask_zero_hash)
        if idx.dim() == 1:
            idx = idx.reshape(-1, 1)
        emb = O.embedding_lookup(tables[names[fc.embedding_name]], idx)
        m = (idx != 0) if fc.embedding_name in mz else None
        return emb, m

    q, qm = zip(*[look(c, True) for c in sparse if c.name in history_feature_list])
    k, km = zip(*[look(c, True) for c in history])
    deep = [look(c, c.name in history_feature_list)[0] for c in sparse]
    if others:
        pooled, _ = input_from_feature_columns(inputs, others, tables)
        deep += pooled
    dense = [torch.as_tensor(np.asarray(inputs[c.name], dtype=np.float32)).reshape(-1, c.dimension)
             for c in dense_cols]
    keys = torch.cat(k, dim=-1)
    key_mask = None
    for m in km:                               # Concat.compute_mask: AND across features
        if m is not None:
            key_mask = m if key_mask is None else (key_mask & m)
    query = torch.cat(q, dim=-1)
    hist = O.attention_sequence_pooling(query, keys, key_mask, W["lau"], att_activation,
                                        att_weight_normalization, training=training)
    deep_in = torch.cat([torch.cat(deep, dim=-1), hist], dim=-1).flatten(1)
    x = torch.cat([deep_in] + dense, dim=-1) if dense else deep_in
    h = O.dnn(x, W["dnn_kernels"], W["dnn_biases"], "relu")
    logit = h @ W["dense_kernel"]
    return logit, O.prediction(logit, W.get("global_bias"), task)
