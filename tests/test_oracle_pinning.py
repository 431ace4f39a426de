"""CPU: pin the oracle.

1. against the golden vectors produced by the REFERENCE's own layer code (tests/golden/*.npz, made by
   tests/golden/generate.py running /root/reference/deepctr/layers/*.py under a torch-backed
   ``tensorflow`` stand-in), including the reference's only known-answer vector
   (tests/layers/utils_test.py:20-22);
2. against the closed-form identities of SURVEY.md section 8c;
3. FarmHash: the published empty-string value and agreement between the oracle's restatement and the
   product's independent host copy (bucket values themselves stay unpinned: no TensorFlow here).
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import farmhash
from oracle import ops as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(d["meta"]))
    ins = [torch.as_tensor(d["in_%d" % i]) for i in range(len([k for k in d.files if k.startswith("in_")]))]
    w = {k[2:]: torch.as_tensor(d[k]) for k in d.files if k.startswith("w_")}
    return meta, ins, w, torch.as_tensor(d["out"])


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, prefix + "*.npz")))


def _mask_from(meta, ins, T):
    m = meta["extra"].get("mask")
    if m is None:
        return None
    if isinstance(m, list) and m and (m[0] is None or isinstance(m[0], list) and isinstance(m[0][0], list)):
        m = [x for x in m if x is not None][0]
    return torch.as_tensor(np.asarray(m, dtype=bool))


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.float(), b.float(), rtol=rtol, atol=atol)


def test_golden_set_is_complete():
    assert len(glob.glob(os.path.join(GOLD, "*.npz"))) >= 36


def test_fm():
    meta, ins, w, out = load("fm")
    close(O.fm(ins[0]), out)


@pytest.mark.parametrize("name", names("crossnet"))
def test_crossnet(name):
    meta, ins, w, out = load(name)
    n = meta["kwargs"].get("layer_num", 2)
    ks = [w["kernel%d" % i] for i in range(n)]
    bs = [w["bias%d" % i] for i in range(n)]
    close(O.crossnet(ins[0], ks, bs, meta["kwargs"].get("parameterization", "vector")), out)


@pytest.mark.parametrize("name", names("cin"))
def test_cin(name):
    meta, ins, w, out = load(name)
    kw = meta["kwargs"]
    n = len(kw["layer_size"])
    got = O.cin(ins[0], [w["filter%d" % i] for i in range(n)], [w["bias%d" % i] for i in range(n)],
                tuple(kw["layer_size"]), kw["activation"], kw["split_half"])
    close(got, out, 1e-5, 1e-5)


@pytest.mark.parametrize("name", names("interacting"))
def test_interacting(name):
    meta, ins, w, out = load(name)
    kw = meta["kwargs"]
    got = O.interacting(ins[0], w["query"], w["key"], w["value"], w.get("res"), kw["head_num"],
                        kw["att_embedding_size"], kw.get("use_res", True), kw.get("scaling", False))
    close(got, out)


@pytest.mark.parametrize("name", names("seqpool"))
def test_sequence_pooling(name):
    meta, ins, w, out = load(name)
    if meta["kwargs"]["supports_masking"]:
        got = O.sequence_pooling(ins[0], meta["kwargs"]["mode"], mask=_mask_from(meta, ins, ins[0].shape[1]))
    else:
        got = O.sequence_pooling(ins[0], meta["kwargs"]["mode"], lengths=ins[1].reshape(-1))
    # the ORDER of the fp32 accumulation inside reduce_sum is internal to TensorFlow (torch here): the
    # oracle fixes it to ascending t, so this comparison is to fp32 rounding, not to the last bit
    close(got, out, 1e-5, 1e-6)


@pytest.mark.parametrize("name", names("weightedseq"))
def test_weighted_sequence(name):
    meta, ins, w, out = load(name)
    norm = meta["kwargs"]["weight_normalization"]
    if meta["kwargs"]["supports_masking"]:
        got = O.weighted_sequence(ins[0], ins[1], norm, mask=_mask_from(meta, ins, ins[0].shape[1]))
    else:
        got = O.weighted_sequence(ins[0], ins[2], norm, lengths=ins[1].reshape(-1))
    close(got, out)


def _lau_weights(w, prefix="local_att/"):
    n = len([k for k in w if k.startswith(prefix + "dnn/kernel")])
    d = {"dnn_kernels": [w[prefix + "dnn/kernel%d" % i] for i in range(n)],
         "dnn_biases": [w[prefix + "dnn/bias%d" % i] for i in range(n)],
         "kernel": w[prefix + "kernel"], "bias": w[prefix + "bias"]}
    if (prefix + "dnn/activation_layers0/dice_alpha") in w:
        d["act_params"] = [{"alphas": w[prefix + "dnn/activation_layers%d/dice_alpha" % i],
                            "moving_mean": w[prefix + "dnn/activation_layers%d/bn/moving_mean" % i],
                            "moving_var": w[prefix + "dnn/activation_layers%d/bn/moving_variance" % i]}
                           for i in range(n)]
    return d


@pytest.mark.parametrize("name", names("din_att"))
def test_attention_sequence_pooling(name):
    meta, ins, w, out = load(name)
    kw = meta["kwargs"]
    T = ins[1].shape[1]
    if kw.get("supports_masking"):
        mask = _mask_from(meta, ins, T)
    else:
        mask = O.sequence_mask(ins[2].reshape(-1), T)
    got = O.attention_sequence_pooling(ins[0], ins[1], mask, _lau_weights(w), kw["att_activation"],
                                       kw.get("weight_normalization", False), kw.get("return_score", False))
    close(got, out, 1e-5, 1e-6)


def test_local_activation_unit():
    meta, ins, w, out = load("lau_sigmoid")
    got = O.local_activation_unit(ins[0], ins[1], act="sigmoid", **_lau_weights(w, ""))
    close(got, out)


@pytest.mark.parametrize("name", names("dnn"))
def test_dnn(name):
    meta, ins, w, out = load(name)
    kw = meta["kwargs"]
    n = len(kw["hidden_units"])
    params = None
    if kw["activation"] == "dice":
        params = [{"alphas": w["activation_layers%d/dice_alpha" % i],
                   "moving_mean": w["activation_layers%d/bn/moving_mean" % i],
                   "moving_var": w["activation_layers%d/bn/moving_variance" % i]} for i in range(n)]
    got = O.dnn(ins[0], [w["kernel%d" % i] for i in range(n)], [w["bias%d" % i] for i in range(n)],
                kw["activation"], kw.get("output_activation"), params, bool(meta["extra"].get("training")))
    close(got, out, 1e-5, 1e-6)


def test_dice_training_statistics():
    meta, ins, w, out = load("dice_train")
    close(O.dice(ins[0], w["dice_alpha"], training=True), out, 1e-5, 1e-6)


@pytest.mark.parametrize("name", names("prediction"))
def test_prediction(name):
    meta, ins, w, out = load(name)
    close(O.prediction(ins[0], w["global_bias"], meta["kwargs"]["task"]), out)


@pytest.mark.parametrize("name", names("linear"))
def test_linear(name):
    meta, ins, w, out = load(name)
    mode = meta["kwargs"]["mode"]
    sp = ins[0] if mode in (0, 2) else None
    dn = ins[0] if mode == 1 else (ins[1] if mode == 2 else None)
    got = O.linear(sp, dn, w.get("linear_kernel"), w.get("linear_bias"))
    close(got.reshape(-1), out.reshape(-1))        # mode 0 is [B,1,1] in the reference, [B,1] here (App. F.3)


def test_hash_vocabulary_known_answer(tmp_path):
    d = np.load(os.path.join(GOLD, "hash_vocab_kat.npz"))
    p = tmp_path / "vocab.csv"
    p.write_text(str(d["vocab"]))
    got = O.hash_layer(np.array([[k] for k in d["keys"]]), 4, vocabulary=O.load_vocabulary(str(p)))
    assert got.tolist() == d["out"].tolist() == [[1], [3], [0]]


# ---- closed forms ----------------------------------------------------------------
def test_closed_forms():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(6, 5, 4, generator=g, dtype=torch.float64)
    pair = sum((x[:, i] * x[:, j]).sum(-1) for i in range(5) for j in range(i + 1, 5))
    assert (O.fm(x)[:, 0] - pair).abs().max() < 1e-12
    xf = x.float()
    W = torch.randn(1, 25, 7, generator=g)
    b = torch.zeros(7)
    lit = O.cin(xf, [W], [b], (7,), "linear", False)
    ein = torch.einsum("bid,bjd,ijn->bdn", xf, xf, W[0].reshape(5, 5, 7)).sum(dim=1)
    assert (lit - ein).abs().max() < 1e-4
    x2 = torch.randn(9, 6, generator=g)
    w = torch.randn(6, 1, generator=g)
    bb = torch.randn(6, 1, generator=g)
    want = x2 * (x2 @ w) + bb[:, 0] + x2
    assert torch.equal(O.crossnet(x2, [w], [bb], "vector"), want) or \
        (O.crossnet(x2, [w], [bb], "vector") - want).abs().max() < 1e-6


# ---- FarmHash ---------------------------------------------------------------------------------------------
def test_farmhash_restatements_agree():
    from deepctr_b200.layers.utils import _fingerprint64, host_hash_array
    assert farmhash.fingerprint64(b"") == 0x9ae16a3b2f90404f          # k2: the published empty-string value
    rng = np.random.RandomState(0)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789-_", dtype=np.uint8)
    for n in list(range(0, 65)) * 2:
        s = bytes(rng.choice(alphabet, size=n).tolist())
        assert farmhash.fingerprint64(s) == _fingerprint64(s)
    ids = rng.randint(0, 10 ** 9, size=200)
    for nb, mz in [(1000, False), (1000, True)]:
        assert np.array_equal(O.hash_layer(ids, nb, mz), host_hash_array(ids, nb, mz))
    assert O.hash_layer(np.array([0, 5]), 10, True)[0] == 0             # mask_zero keeps id 0 at bucket 0
    assert 1 <= O.hash_layer(np.array([0, 5]), 10, True)[1] <= 9
