"""Shared test helpers: synthetic CTR data and model -> oracle weight extraction."""
import numpy as np
import torch


def criteo_like(rng, n, n_sparse=6, n_dense=3, vocab=50, dim=8, dtype="int32"):
    from deepctr_b200.feature_column import SparseFeat, DenseFeat
    cols = [SparseFeat("C%d" % i, vocab + i, dim, dtype=dtype) for i in range(n_sparse)]
    cols += [DenseFeat("I%d" % i, 1) for i in range(n_dense)]
    x = {"C%d" % i: rng.randint(0, vocab + i, size=n).astype(np.int32 if dtype == "int32" else np.int64)
         for i in range(n_sparse)}
    x.update({"I%d" % i: rng.rand(n).astype(np.float32) for i in range(n_dense)})
    y = (rng.rand(n) < 0.3).astype(np.float32)
    return cols, x, y


def randomize_weights(model, rng, std=0.1):
    """Replace the reference initialisers (zeros / 1e-4) by O(0.1) values so every term of the logit
    carries signal in the parity check."""
    for w in model.weights:
        if w.name.endswith("moving_variance"):
            w.set_value(np.abs(rng.normal(1.0, 0.1, size=w.shape)).astype(np.float32))
        else:
            w.set_value(rng.normal(0, std, size=w.shape).astype(np.float32))


def oracle_weights(model, requires_grad=False):
    """Collect the model's weights in the structure oracle/models.py expects."""
    from deepctr_b200.inputs import Embedding
    from deepctr_b200.layers.core import DNN, PredictionLayer
    from deepctr_b200.layers.interaction import CIN, CrossNet, InteractingLayer
    from deepctr_b200.layers.sequence import AttentionSequencePoolingLayer
    from deepctr_b200.layers.utils import Linear
    from deepctr_b200.engine import Dense

    def t(w):
        return torch.tensor(w.value(), requires_grad=requires_grad)

    W = {"tables": {}, "att": []}
    denses = []
    for l in model.layers:
        if isinstance(l, Embedding):
            W["tables"][l.name] = t(l.embeddings)
        elif isinstance(l, AttentionSequencePoolingLayer):
            lau = l.local_att
            d = {"dnn_kernels": [t(k) for k in lau.dnn.kernels], "dnn_biases": [t(b) for b in lau.dnn.bias],
                 "kernel": t(lau.kernel), "bias": t(lau.bias)}
            acts = []
            for al in lau.dnn.activation_layers:
                if al is not None and al.__class__.__name__ == "Dice":
                    acts.append({"alphas": t(al.alphas), "moving_mean": t(al.moving_mean),
                                 "moving_var": t(al.moving_variance)})
                else:
                    acts.append(None)
            if any(a is not None for a in acts):
                d["act_params"] = acts
            W["lau"] = d
        elif isinstance(l, DNN):
            W["dnn_kernels"] = [t(k) for k in l.kernels]
            W["dnn_biases"] = [t(b) for b in l.bias]
        elif isinstance(l, Linear):
            if l.mode in (1, 2):
                W["linear_kernel"] = t(l.kernel)
        elif isinstance(l, PredictionLayer):
            if l.use_bias:
                W["global_bias"] = t(l.global_bias)
        elif isinstance(l, CIN):
            W["cin_filters"] = [t(f) for f in l.filters]
            W["cin_biases"] = [t(b) for b in l.bias]
        elif isinstance(l, CrossNet):
            W["cross_kernels"] = [t(k) for k in l.kernels]
            W["cross_biases"] = [t(b) for b in l.bias]
        elif isinstance(l, InteractingLayer):
            d = {"query": t(l.W_Query), "key": t(l.W_key), "value": t(l.W_Value)}
            if l.use_res:
                d["res"] = t(l.W_Res)
            W["att"].append(d)
        elif isinstance(l, Dense):
            denses.append(l)
    W["_dense_layers"] = denses
    if denses:
        W["dense_kernel"] = t(denses[0].kernel)
    if len(denses) > 1:
        W["cin_dense_kernel"] = t(denses[1].kernel)
    return W


def flat_params(W):
    """name -> leaf tensor for every oracle weight (for gradient comparison)."""
    out = {}
    for k, v in W.items():
        if k.startswith("_"):
            continue
        if isinstance(v, torch.Tensor):
            out[k] = v
        elif isinstance(v, dict):
            for k2, v2 in v.items():
                if isinstance(v2, torch.Tensor):
                    out["%s/%s" % (k, k2)] = v2
                elif isinstance(v2, list):
                    for i, e in enumerate(v2):
                        if isinstance(e, torch.Tensor):
                            out["%s/%s/%d" % (k, k2, i)] = e
                        elif isinstance(e, dict):
                            for k3, v3 in e.items():
                                if isinstance(v3, torch.Tensor):
                                    out["%s/%s/%d/%s" % (k, k2, i, k3)] = v3
        elif isinstance(v, list):
            for i, e in enumerate(v):
                if isinstance(e, torch.Tensor):
                    out["%s/%d" % (k, i)] = e
                elif isinstance(e, dict):
                    for k2, v2 in e.items():
                        out["%s/%d/%s" % (k, i, k2)] = v2
    return out


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-6)))
