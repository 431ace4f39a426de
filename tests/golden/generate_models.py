#!/usr/bin/env python
"""Generate tests/golden/models/*.npz: MODEL-level golden vectors produced by the REFERENCE's own,
unmodified composition code - /root/reference/deepctr/feature_column.py, inputs.py, layers/*.py and
the five builders models/{deepfm,xdeepfm,dcn,autoint}.py, models/sequence/din.py - executed eagerly
under the torch-backed ``tensorflow`` stand-in of tf_torch_shim.py.

    python tests/golden/generate_models.py          (build container only: needs /root/reference)

What is the reference's: which tables exist and how they are named and shared, mask_zero rules, the
Hash / lookup / pooling / attention dispatch, group handling, the order of the DNN input columns, every
layer's op sequence, the way the logits are added.  What is torch's: the arithmetic inside each op
(TensorFlow's kernels are not installable here - DESIGN.md section 7).  FarmHash buckets come from
oracle/farmhash.py (TensorFlow's is unavailable; string-hash cases pin the COMPOSITION, not the hash).

Every fixture holds: ``meta`` (json: builder, kwargs, feature columns, training flag), the inputs
``x_<feature>``, labels ``y``, every weight ``w_<key>`` as the reference layers created and named
them, the model output ``out`` ([B,1]), the pre-activation ``logit``, the Keras binary-crossentropy /
mse ``loss`` (SURVEY.md App. C formula) and ``g_<key>`` = d loss / d weight by torch autograd THROUGH
the reference's graph.

Weight keys: ``<top-level layer name>/<attribute path>/<weight name>``; top-level layers are the ones
the builder / feature_column code creates directly (Keras-style unique names: ``dnn``, ``dense``,
``dense_1``, ``sparse_emb_C1``, ``linear0sparse_emb_C1`` ...), nested layers are addressed by the
reference's attribute names (``local_att/dnn/activation_layers0/bn``).
"""
import json
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "models")

import tf_torch_shim as S  # noqa: E402

REF = "/root/reference"


def _collect(layer, prefix):
    """(key, tensor) for every weight below ``layer``, following the reference's attribute names."""
    out = []
    for name, t in layer._w:
        out.append((prefix + name, t))
    for key, v in vars(layer).items():
        if key.startswith("_last"):
            continue
        vs = v if isinstance(v, (list, tuple)) else [v]
        for i, e in enumerate(vs):
            if isinstance(e, S.Layer) and e is not layer:
                sub = "%s%s%s/" % (prefix, key, "" if not isinstance(v, (list, tuple)) else str(i))
                out.extend(_collect(e, sub))
                if isinstance(e, S.BatchNormalization) and e.moving_mean is not None:
                    out.append((sub + "moving_mean", e.moving_mean))
                    out.append((sub + "moving_variance", e.moving_variance))
    return out


def _col_meta(fc, FC):
    if isinstance(fc, FC.SparseFeat):
        return {"kind": "sparse", "name": fc.name, "vocabulary_size": int(fc.vocabulary_size),
                "embedding_dim": int(fc.embedding_dim), "use_hash": bool(fc.use_hash),
                "vocabulary_path": os.path.basename(fc.vocabulary_path) if fc.vocabulary_path else None,
                "dtype": fc.dtype, "embedding_name": fc.embedding_name, "group_name": fc.group_name,
                "trainable": bool(fc.trainable)}
    if isinstance(fc, FC.VarLenSparseFeat):
        return {"kind": "varlen", "sparsefeat": _col_meta(fc.sparsefeat, FC), "maxlen": int(fc.maxlen),
                "combiner": fc.combiner, "length_name": fc.length_name, "weight_name": fc.weight_name,
                "weight_norm": bool(fc.weight_norm)}
    return {"kind": "dense", "name": fc.name, "dimension": int(fc.dimension), "dtype": fc.dtype}


def keras_loss(task, y, p):
    """tf.keras binary_crossentropy on probabilities / mean squared error."""
    y = torch.as_tensor(np.asarray(y, dtype=np.float32)).reshape(-1, 1)
    if task == "binary":
        eps = 1e-7
        pc = torch.clamp(p, eps, 1.0 - eps)
        return -(y * torch.log(pc + eps) + (1.0 - y) * torch.log(1.0 - pc + eps)).mean()
    return ((p - y) ** 2).mean()


def run_case(name, builder, make_args, x, y, seed, training=None, task="binary", note="", scale=1.0):
    """make_args(FC) -> (positional args, kwargs, linear columns, dnn columns) built with the
    REFERENCE's feature-column classes."""
    S.CTX.reset()
    S.CTX.feed = dict(x)
    S.CTX.rng = np.random.RandomState(seed)
    S.CTX.grad = True
    S.CTX.training = training
    S.CTX.scale = scale
    FC = sys.modules["deepctr.feature_column"]
    import importlib
    mod = importlib.import_module({"DeepFM": "deepctr.models.deepfm", "xDeepFM": "deepctr.models.xdeepfm",
                                   "DCN": "deepctr.models.dcn", "AutoInt": "deepctr.models.autoint",
                                   "DIN": "deepctr.models.sequence.din"}[builder])
    args, kwargs, cols_meta = make_args(FC)
    model = getattr(mod, builder)(*args, **kwargs)
    out = model.outputs
    pred_layers = [l for l in model.layers if l.__class__.__name__ == "PredictionLayer"]
    logit = pred_layers[-1]._last_in
    loss = keras_loss(task, y, out)
    weights = []
    for l in model.layers:
        if not l._nested:
            weights.extend(_collect(l, l.name + "/"))
    leaves = [t for _, t in weights if t.requires_grad]
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    gmap = {id(t): g for t, g in zip(leaves, grads)}
    d = {"meta": np.array(json.dumps({"builder": builder, "kwargs": kwargs_json(kwargs), "columns": cols_meta,
                                      "training": training, "task": task, "note": note,
                                      "layers": [[l.__class__.__name__, l.name] for l in model.layers
                                                 if not l._nested]}))}
    for k, v in x.items():
        a = np.asarray(v)
        d["x_" + k] = a.astype(str) if a.dtype.kind in "OUS" else a
    d["y"] = np.asarray(y, dtype=np.float32)
    for k, t in weights:
        d["w_" + k] = t.detach().numpy().copy()
        g = gmap.get(id(t))
        if t.requires_grad:
            d["g_" + k] = (g if g is not None else torch.zeros_like(t)).detach().numpy().copy()
    d["out"] = out.detach().numpy().reshape(-1, 1)
    d["logit"] = logit.detach().numpy().reshape(-1, 1)
    d["loss"] = np.asarray(float(loss.detach()), dtype=np.float64)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("%-28s out%s logit[min %.3f max %.3f] loss %.5f  %d weights" %
          (name, tuple(d["out"].shape), float(d["logit"].min()), float(d["logit"].max()), float(loss.detach()), len(weights)))


def kwargs_json(kw):
    out = {}
    for k, v in kw.items():
        out[k] = list(v) if isinstance(v, tuple) else v
    return out


def criteo_like(rng, n, n_sparse, n_dense, dim, vocab0=20):
    def make(FC):
        cols = [FC.SparseFeat("C%d" % (i + 1), vocab0 + 3 * i, dim) for i in range(n_sparse)]
        cols += [FC.DenseFeat("I%d" % (i + 1), 1) for i in range(n_dense)]
        return cols
    x = {"C%d" % (i + 1): rng.randint(0, vocab0 + 3 * i, size=n).astype(np.int32) for i in range(n_sparse)}
    x.update({"I%d" % (i + 1): rng.rand(n).astype(np.float32) for i in range(n_dense)})
    y = (rng.rand(n) < 0.3).astype(np.float32)
    return make, x, y


def main():
    S.install(REF)
    import importlib
    importlib.import_module("deepctr.feature_column")
    rng = np.random.RandomState(20260923)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)

    # ---- DeepFM, Criteo shape (examples/run_classification_criteo.py:28-44) ---------------------
    mk, x, y = criteo_like(rng, 32, 6, 3, 8)

    def deepfm_args(FC):
        cols = mk(FC)
        return (cols, cols), dict(dnn_hidden_units=(32, 16), l2_reg_linear=0, l2_reg_embedding=0), \
            {"linear": [_col_meta(c, FC) for c in cols], "dnn": [_col_meta(c, FC) for c in cols]}
    run_case("deepfm_criteo", "DeepFM", deepfm_args, x, y, 1)

    # ---- DeepFM over groups + pooled VarLen features (length / mask, weighted, sum / mean / max) + an
    #      integer feature hashed on the fly (tests/utils.py:38-105 style columns)
    n, T = 24, 5
    lens = rng.randint(0, T + 1, size=n).astype(np.int32)
    lens[:3] = (0, 1, T)

    def padded(vocab, ln=None):
        ln = lens if ln is None else ln
        a = rng.randint(1, vocab, size=(n, T)).astype(np.int32)
        a[np.arange(T)[None, :] >= ln[:, None]] = 0
        return a
    # (max pooling of an empty bag is -1e9, sequence.py:103: "cats" keeps >= 1 valid id per row)
    xg = {"user": rng.randint(0, 30, size=n).astype(np.int32), "item": rng.randint(0, 40, size=n).astype(np.int32),
          "city": rng.randint(0, 100000, size=n).astype(np.int32), "price": rng.rand(n, 2).astype(np.float32),
          "tags": padded(17), "tags_len": lens.copy(), "tags_w": rng.rand(n, T, 1).astype(np.float32),
          "cats": padded(11, np.maximum(lens, 1)), "hist": padded(40), "kw": padded(23), "kw_w": rng.rand(n, T, 1).astype(np.float32)}
    yg = (rng.rand(n) < 0.4).astype(np.float32)

    def groups_args(FC):
        cols = [FC.SparseFeat("user", 30, 8, group_name="g_user"),
                FC.SparseFeat("item", 40, 8, group_name="g_item"),
                FC.SparseFeat("city", 50, 8, use_hash=True, group_name="g_user"),
                FC.DenseFeat("price", 2),
                FC.VarLenSparseFeat(FC.SparseFeat("tags", 17, 8, group_name="g_item"), maxlen=T, combiner="sum",
                                    length_name="tags_len", weight_name="tags_w", weight_norm=True),
                FC.VarLenSparseFeat(FC.SparseFeat("cats", 11, 8, group_name="g_item"), maxlen=T, combiner="max"),
                FC.VarLenSparseFeat(FC.SparseFeat("hist", 40, 8, embedding_name="item", group_name="g_user"),
                                    maxlen=T, combiner="mean", length_name="tags_len"),
                FC.VarLenSparseFeat(FC.SparseFeat("kw", 23, 8, group_name="g_other"), maxlen=T, combiner="mean",
                                    weight_name="kw_w", weight_norm=False)]
        return (cols, cols), dict(fm_group=("g_user", "g_item"), dnn_hidden_units=(16, 8), l2_reg_linear=0,
                                  l2_reg_embedding=0), \
            {"linear": [_col_meta(c, FC) for c in cols], "dnn": [_col_meta(c, FC) for c in cols]}
    run_case("deepfm_groups_varlen", "DeepFM", groups_args, xg, yg, 2,
             note="hashed int feature: bucket ids from oracle/farmhash.py (TF FarmHash unavailable)")

    # ---- DeepFM regression on string ids hashed on the fly + a vocabulary file
    #      (examples/run_multivalue_movielens_vocab_hash.py:27-47, rows of examples/movielens_sample.txt)
    import pandas as pd
    data = pd.read_csv(os.path.join(REF, "examples", "movielens_sample.txt")).iloc[:40]
    sparse_features = ["movie_id", "user_id", "gender", "age", "occupation", "zip"]
    data[sparse_features] = data[sparse_features].astype(str)
    genres_list = [g.split("|") for g in data["genres"].values]
    max_len = max(len(g) for g in genres_list)
    genres = np.array([g + ["0"] * (max_len - len(g)) for g in genres_list], dtype=object).astype(str)
    os.makedirs(OUT, exist_ok=True)
    shutil.copy(os.path.join(REF, "examples", "movielens_age_vocabulary.csv"),
                os.path.join(OUT, "movielens_age_vocabulary.csv"))
    nun = {f: int(data[f].nunique()) for f in sparse_features}
    xm = {f: data[f].values.astype(str) for f in sparse_features}
    xm["genres"] = genres
    ym = data["rating"].values.astype(np.float32)

    def movielens_args(FC):
        fix = [FC.SparseFeat(f, nun[f] * 5, embedding_dim=4, use_hash=True,
                             vocabulary_path=os.path.join(OUT, "movielens_age_vocabulary.csv") if f == "age" else None,
                             dtype="string") for f in sparse_features]
        var = [FC.VarLenSparseFeat(FC.SparseFeat("genres", vocabulary_size=100, embedding_dim=4, use_hash=True,
                                                 dtype="string"), maxlen=max_len, combiner="mean")]
        cols = fix + var
        return (cols, cols), dict(task="regression", dnn_hidden_units=(16, 8), l2_reg_linear=0, l2_reg_embedding=0), \
            {"linear": [_col_meta(c, FC) for c in cols], "dnn": [_col_meta(c, FC) for c in cols]}
    run_case("deepfm_movielens_hash", "DeepFM", movielens_args, xm, ym, 3, task="regression",
             note="string ids: FarmHash buckets from oracle/farmhash.py; 'age' through the reference's vocabulary file")

    # ---- xDeepFM (models/xdeepfm.py:50-65; tests/models/xDeepFM_test.py:7-25 parameter sets) ------------
    mk, x, y = criteo_like(rng, 32, 5, 2, 4)

    def xdeepfm_args(cin, split, act):
        def f(FC):
            cols = mk(FC)
            return (cols, cols), dict(dnn_hidden_units=(16, 8), cin_layer_size=cin, cin_split_half=split,
                                      cin_activation=act, l2_reg_linear=0, l2_reg_embedding=0), \
                {"linear": [_col_meta(c, FC) for c in cols], "dnn": [_col_meta(c, FC) for c in cols]}
        return f
    run_case("xdeepfm_split_relu", "xDeepFM", xdeepfm_args((8, 6), True, "relu"), x, y, 4)
    run_case("xdeepfm_nosplit_linear", "xDeepFM", xdeepfm_args((6, 5, 4), False, "linear"), x, y, 5)
    run_case("xdeepfm_nocin", "xDeepFM", xdeepfm_args((), True, "relu"), x, y, 6)

    # ---- DCN (models/dcn.py:45-76; tests/models/DCN_test.py incl. test_DCN_2: empty linear columns) -----
    def dcn_args(cross_num, param, hidden, empty_linear=False):
        def f(FC):
            cols = mk(FC)
            lin = [] if empty_linear else cols
            return (lin, cols), dict(cross_num=cross_num, cross_parameterization=param, dnn_hidden_units=hidden,
                                     l2_reg_linear=0, l2_reg_embedding=0, l2_reg_cross=0), \
                {"linear": [_col_meta(c, FC) for c in lin], "dnn": [_col_meta(c, FC) for c in cols]}
        return f
    run_case("dcn_vector2", "DCN", dcn_args(2, "vector", (16, 8)), x, y, 7)
    run_case("dcn_matrix1", "DCN", dcn_args(1, "matrix", (16, 8)), x, y, 8)
    run_case("dcn_crossonly", "DCN", dcn_args(2, "vector", ()), x, y, 9)
    run_case("dcn_empty_linear", "DCN", dcn_args(1, "vector", (8,), True), x, y, 10)

    # ---- AutoInt (models/autoint.py:46-82) ---------------------------------------------------------
    def autoint_args(layers, heads, res, hidden):
        def f(FC):
            cols = mk(FC)
            return (cols, cols), dict(att_layer_num=layers, att_embedding_size=4, att_head_num=heads, att_res=res,
                                      dnn_hidden_units=hidden, l2_reg_linear=0, l2_reg_embedding=0), \
                {"linear": [_col_meta(c, FC) for c in cols], "dnn": [_col_meta(c, FC) for c in cols]}
        return f
    run_case("autoint_2x2_res", "AutoInt", autoint_args(2, 2, True, (16, 8)), x, y, 11)
    run_case("autoint_attonly", "AutoInt", autoint_args(1, 1, False, ()), x, y, 12)

    # ---- DIN on the reference's own 3-row batch (tests/models/DIN_test.py:10-36 = examples/run_din.py:7-33)
    xd = {"user": np.array([0, 1, 2]), "gender": np.array([0, 1, 0]), "item_id": np.array([1, 2, 3]),
          "cate_id": np.array([1, 2, 2]), "pay_score": np.array([0.1, 0.2, 0.3], dtype=np.float32),
          "hist_item_id": np.array([[1, 2, 3, 0], [3, 2, 1, 0], [1, 2, 0, 0]]),
          "hist_cate_id": np.array([[1, 2, 2, 0], [2, 2, 1, 0], [1, 2, 0, 0]]), "seq_length": np.array([3, 3, 2])}
    yd = np.array([1, 0, 1], dtype=np.float32)

    def din_cols(FC):
        cols = [FC.SparseFeat("user", 3, embedding_dim=10), FC.SparseFeat("gender", 2, embedding_dim=4),
                FC.SparseFeat("item_id", 3 + 1, embedding_dim=8), FC.SparseFeat("cate_id", 2 + 1, embedding_dim=4),
                FC.DenseFeat("pay_score", 1)]
        cols += [FC.VarLenSparseFeat(FC.SparseFeat("hist_item_id", vocabulary_size=3 + 1, embedding_dim=8,
                                                   embedding_name="item_id"), maxlen=4, length_name="seq_length"),
                 FC.VarLenSparseFeat(FC.SparseFeat("hist_cate_id", 2 + 1, embedding_dim=4, embedding_name="cate_id"),
                                     maxlen=4, length_name="seq_length")]
        return cols

    def din_args(act, wn, hidden=(4, 4, 4)):
        def f(FC):
            cols = din_cols(FC)
            return (cols, ["item_id", "cate_id"]), dict(dnn_hidden_units=hidden, att_activation=act,
                                                        att_weight_normalization=wn, l2_reg_embedding=0), \
                {"dnn": [_col_meta(c, FC) for c in cols], "linear": []}
        return f
    run_case("din_ref_batch_sigmoid", "DIN", din_args("sigmoid", False), xd, yd, 13, scale=2.0)
    run_case("din_ref_batch_dice", "DIN", din_args("dice", False), xd, yd, 14, training=False, scale=2.0)
    run_case("din_ref_batch_dice_wn", "DIN", din_args("dice", True), xd, yd, 15, training=False, scale=2.0)
    run_case("din_ref_batch_dice_train", "DIN", din_args("dice", False), xd, yd, 16, training=True, scale=2.0)

    # ---- DIN, a wider batch with an extra pooled VarLen feature and ragged histories -----------------
    n, T = 20, 6
    hl = rng.randint(1, T + 1, size=n)
    hi = rng.randint(1, 30, size=(n, T))
    hc = rng.randint(1, 9, size=(n, T))
    dead = np.arange(T)[None, :] >= hl[:, None]
    hi[dead] = 0
    hc[dead] = 0
    tg = rng.randint(1, 12, size=(n, T))
    tl = rng.randint(0, T + 1, size=n)
    tg[np.arange(T)[None, :] >= tl[:, None]] = 0
    xw = {"user": rng.randint(0, 15, size=n), "item_id": rng.randint(1, 30, size=n), "cate_id": rng.randint(1, 9, size=n),
          "pay_score": rng.rand(n).astype(np.float32), "hist_item_id": hi, "hist_cate_id": hc, "seq_length": hl,
          "tags": tg}
    yw = (rng.rand(n) < 0.5).astype(np.float32)

    def din_wide(FC):
        cols = [FC.SparseFeat("user", 15, embedding_dim=6), FC.SparseFeat("item_id", 30, embedding_dim=8),
                FC.SparseFeat("cate_id", 9, embedding_dim=4), FC.DenseFeat("pay_score", 1),
                FC.VarLenSparseFeat(FC.SparseFeat("hist_item_id", 30, embedding_dim=8, embedding_name="item_id"),
                                    maxlen=T, length_name="seq_length"),
                FC.VarLenSparseFeat(FC.SparseFeat("hist_cate_id", 9, embedding_dim=4, embedding_name="cate_id"),
                                    maxlen=T, length_name="seq_length"),
                FC.VarLenSparseFeat(FC.SparseFeat("tags", 12, embedding_dim=4), maxlen=T, combiner="sum")]
        return (cols, ["item_id", "cate_id"]), dict(dnn_hidden_units=(16, 8), att_hidden_size=(12, 6),
                                                    att_activation="dice", att_weight_normalization=False,
                                                    l2_reg_embedding=0), \
            {"dnn": [_col_meta(c, FC) for c in cols], "linear": []}
    run_case("din_wide_dice", "DIN", din_wide, xw, yw, 17, training=False, scale=2.0)


if __name__ == "__main__":
    main()
