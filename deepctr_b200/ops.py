"""Differentiable ops on ``engine.Var``: each launches libb2ctr kernels for the forward and pushes a
closure on the active tape that launches the backward kernels.  No torch arithmetic anywhere.
"""
import numpy as np
import torch

from . import _lib as L
from . import kernels as K
from . import engine as E

# precision mode used by every dense GEMM (DNN / attention MLP / projections); see DESIGN.md
GEMM_PRECISION = L.GEMM_FP32


def set_gemm_precision(mode):
    """'fp32' (exact FFMA) or 'bf16x3' (tcgen05 split-bf16, ~2^-17 relative)."""
    global GEMM_PRECISION
    GEMM_PRECISION = {"fp32": L.GEMM_FP32, "bf16x3": L.GEMM_BF16X3}[mode]


def _empty(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


def _split_k(m_out, n_out, kred):
    """wgrad-style GEMMs reduce over the batch: give every SM a slice."""
    if kred < 4096:
        return 1
    tiles = ((m_out + 127) // 128) * ((n_out + 127) // 128)
    want = (2 * 148 + tiles - 1) // tiles
    return int(max(1, min(64, want, kred // 1024)))


def _as2d(x):
    t2, ld = x.as2d()
    if t2 is None:
        t = E.contiguous(x)
        t2 = t.reshape(-1, t.shape[-1])
        ld = t2.stride(0)
    return t2, ld


# ---- GEMM based --------------------------------------------------------------------------------
def dense(x, w, b=None, activation=None):
    """y = act(x @ w + b) over the last axis (tf.tensordot(x, w, axes=(-1, 0)) + bias_add)."""
    act = L.ACT_BY_NAME.get(activation, None)
    fused_act = act if act is not None else L.ACT_NONE
    x2, _ = _as2d(x)
    m, kdim = x2.shape
    n = w.shape[1]
    wd = w.materialize() if isinstance(w, E.Weight) else w.data
    bd = (b.materialize() if isinstance(b, E.Weight) else b.data) if b is not None else None
    y = K.gemm(x2, wd, bias=bd, act=fused_act, precision=GEMM_PRECISION, m=m, n=n, k=kdim)
    out = E.Var(y.reshape(tuple(x.data.shape[:-1]) + (n,)))
    if act is None and activation is not None:
        raise ValueError("activation %r cannot be fused into dense(); apply it as a layer" % activation)

    def bwd(grads):
        dy = grads[0].reshape(m, n)
        if not dy.is_contiguous():
            dy = dy.contiguous()
        need_db = b is not None and b.requires_grad
        if fused_act != L.ACT_NONE or need_db:
            dz, db = K.bias_act_bwd(dy, y, fused_act, want_dz=fused_act != L.ACT_NONE, want_dbias=need_db)
            if dz is None:
                dz = dy
        else:
            dz, db = dy, None
        if x.requires_grad:
            base = x.base
            if (base is not None and x.col0 == 0 and x.ncols != -1 and base.data.dim() == 2
                    and base.data.shape[0] == m and x.ncols == kdim):
                # x is the leading window of a K-padded buffer: write dx with the same ld so that
                # the buffer's gradient is adopted without a copy
                ld = base.data.stride(0)
                buf = _empty((m, ld), dy)
                dxw = buf[:, :kdim]
                K.gemm(dz, wd, c=dxw, trans_b=True, precision=GEMM_PRECISION, m=m, n=kdim, k=n)
                E.add_grad(x, dxw)
            else:
                dx = K.gemm(dz, wd, trans_b=True, precision=GEMM_PRECISION, m=m, n=kdim, k=n)
                E.add_grad(x, dx.reshape(x.data.shape))
        if w.requires_grad:
            dw = K.gemm(x2, dz, trans_a=True, precision=GEMM_PRECISION, split_k=_split_k(kdim, n, m),
                        m=kdim, n=n, k=m)
            E.add_grad(w, dw)
        if need_db:
            E.add_grad(b, db)

    E.record([out], [x, w, b], bwd)
    return out


def activation(x, name):
    act = L.ACT_BY_NAME[name]
    if act == L.ACT_NONE:
        return x
    xt = E.contiguous(x)
    y = K.act_fwd(xt, act)
    out = E.Var(y)

    def bwd(grads):
        dz, _ = K.bias_act_bwd(grads[0].reshape(-1, y.shape[-1]).contiguous(), y.reshape(-1, y.shape[-1]),
                               act, want_dz=True, want_dbias=False)
        E.add_grad(x, dz.reshape(x.data.shape))

    E.record([out], [x], bwd)
    return out


# ---- shape plumbing (zero-copy whenever the operands alias one buffer) --------------------------
def _window(base, col0, ncols, shape, mask=None):
    bt = base.data
    b = bt.shape[0]
    dense_strides = []
    acc = 1
    for s in reversed(shape[1:]):
        dense_strides.append(acc)
        acc *= s
    dense_strides = list(reversed(dense_strides))
    data = bt.as_strided((b,) + tuple(shape[1:]), (bt.stride(0),) + tuple(dense_strides),
                         bt.storage_offset() + col0)
    return E.Var(data, requires_grad=base.requires_grad, mask=mask, base=base, col0=col0, ncols=ncols,
                 owner=base.owner)


def flatten(x):
    b = x.data.shape[0]
    w = int(np.prod(x.data.shape[1:]))
    if x.base is not None and x.ncols != -1:
        return _window(x.base, x.col0, x.ncols, (b, w))
    t = E.contiguous(x)
    out = E.Var(t.reshape(b, w))
    out.base, out.col0, out.ncols = x if x.base is None else x.base, 0, -1
    out.requires_grad = x.requires_grad
    return out


def reshape(x, shape):
    """Per-sample reshape (batch dim kept)."""
    b = x.data.shape[0]
    shape = (b,) + tuple(shape[1:])
    if x.base is not None and x.ncols != -1:
        return _window(x.base, x.col0, x.ncols, shape, x.mask)
    t = E.contiguous(x)
    out = E.Var(t.reshape(shape), mask=x.mask)
    out.base, out.col0, out.ncols = x if x.base is None else x.base, 0, -1
    out.requires_grad = x.requires_grad
    return out


def _flat_concat_ok(shapes, axis):
    nd = len(shapes[0])
    ax = axis if axis >= 0 else nd + axis
    return ax, all(all(s[d] == 1 for d in range(1, ax)) for s in shapes)


def concat(xs, axis=-1):
    xs = list(xs)
    if len(xs) == 1:
        return xs[0]
    shapes = [tuple(v.data.shape) for v in xs]
    ax, flat_ok = _flat_concat_ok(shapes, axis)
    out_shape = list(shapes[0])
    out_shape[ax] = sum(s[ax] for s in shapes)
    b = shapes[0][0]
    widths = [int(np.prod(s[1:])) for s in shapes]
    if flat_ok:
        # zero-copy: adjacent windows of the same buffer, in order
        base = xs[0].base
        if base is not None and xs[0].ncols != -1:
            col = xs[0].col0
            ok = True
            for v, w in zip(xs, widths):
                if v.base is not base or v.ncols == -1 or v.col0 != col or v.ncols != w:
                    ok = False
                    break
                col += w
            if ok:
                return _window(base, xs[0].col0, col - xs[0].col0, tuple(out_shape))
        out = _empty((b, sum(widths)), xs[0].data)
        col = 0
        for v, w in zip(xs, widths):
            src, ld = v.flat2d()
            if src is None:
                src = E.contiguous(v).reshape(b, w)
                ld = w
            K.copy2d(src, ld, out, out.stride(0), b, w, dst_off=col)
            col += w
        res = E.Var(out.reshape(out_shape))

        def bwd(grads):
            g = grads[0].reshape(b, -1)
            c = 0
            for v, w in zip(xs, widths):
                if v.requires_grad:
                    gv = _empty((b, w), g)
                    K.copy2d(g, g.stride(0), gv, w, b, w, src_off=c)
                    E.add_grad(v, gv.reshape(v.data.shape))
                c += w

        E.record([res], xs, bwd)
        return res
    # general case: concatenate along `ax` with non-unit leading dims -> rows = prod(dims < ax)
    lead = int(np.prod(shapes[0][:ax]))
    tails = [int(np.prod(s[ax:])) for s in shapes]
    out = _empty((lead, sum(tails)), xs[0].data)
    col = 0
    for v, w in zip(xs, tails):
        t = E.contiguous(v).reshape(lead, w)
        K.copy2d(t, w, out, out.stride(0), lead, w, dst_off=col)
        col += w
    res = E.Var(out.reshape(out_shape))

    def bwd2(grads):
        g = grads[0].reshape(lead, -1)
        c = 0
        for v, w in zip(xs, tails):
            if v.requires_grad:
                gv = _empty((lead, w), g)
                K.copy2d(g, g.stride(0), gv, w, lead, w, src_off=c)
                E.add_grad(v, gv.reshape(v.data.shape))
            c += w

    E.record([res], xs, bwd2)
    return res


def slice_cols(x, col0, ncols, shape=None):
    """x[:, col0:col0+ncols] of the per-sample flattening."""
    b = x.data.shape[0]
    shape = shape or (b, ncols)
    if x.base is not None and x.ncols != -1:
        return _window(x.base, x.col0 + col0, ncols, shape)
    if x.data.dim() == 2 and x.base is None:
        return _window(x, col0, ncols, shape)
    src, ld = x.flat2d()
    out = _empty((b, ncols), x.data)
    K.copy2d(src, ld, out, ncols, b, ncols, src_off=col0)
    res = E.Var(out.reshape(shape))

    def bwd(grads):
        full = _empty((b, src.shape[1]), out)
        K.fill(full, 0.0)
        K.copy2d(grads[0].reshape(b, ncols), ncols, full, full.stride(0), b, ncols, dst_off=col0)
        E.add_grad(x, full.reshape(x.data.shape))

    E.record([res], [x], bwd)
    return res


def add_n(xs):
    """Keras Add on tensors with the same number of elements per sample ([B,1] / [B,1,1])."""
    xs = [v for v in xs]
    if len(xs) == 1:
        return xs[0]
    n = xs[0].data.numel()
    for v in xs:
        if v.data.numel() != n:
            raise ValueError("add_n: operands must have the same number of elements, got %s" %
                             [tuple(v.data.shape) for v in xs])
    ts = [E.contiguous(v).reshape(-1) for v in xs]
    out_t, i = None, 0
    while i < len(ts):                  # the kernel sums up to 8 operands per launch
        if out_t is None:
            chunk, i = ts[i:i + 8], i + 8
        else:
            chunk, i = [out_t] + ts[i:i + 7], i + 7
        out_t = K.add_n(chunk)
    best = max(xs, key=lambda v: v.data.dim())
    res = E.Var(out_t.reshape(best.data.shape) if best.data.dim() <= 2 else out_t.reshape(-1, 1))

    def bwd(grads):
        g = grads[0]
        for v in xs:
            if v.requires_grad:
                E.add_grad(v, g.reshape(v.data.shape))

    E.record([res], xs, bwd)
    return res


def rowsum(x):
    """sum over all non-batch axes -> [B,1]  (Linear mode 0/2, layers/utils.py:160-171)."""
    src, ld = x.flat2d()
    if src is None:
        t = E.contiguous(x)
        src = t.reshape(t.shape[0], -1)
        ld = src.stride(0)
    b, w = src.shape
    out = K.rowsum(src, b, w, ld)
    res = E.Var(out.reshape(b, 1))

    def bwd(grads):
        g = grads[0].reshape(b, 1)
        ones = _empty((1, w), g)
        K.fill(ones, 1.0)
        gx = K.gemm(g, ones, m=b, n=w, k=1)      # broadcast as an outer product
        E.add_grad(x, gx.reshape(x.data.shape))

    E.record([res], [x], bwd)
    return res


def zeros_like_batch(x, cols=1):
    b = x.data.shape[0]
    out = torch.empty((b, cols), dtype=torch.float32, device=x.data.device)
    K.fill(out, 0.0)
    return E.Var(out)


# ---- FM ------------------------------------------------------------------------------------------
def fm(x):
    """[B,F,E] -> [B,1]  (layers/interaction.py:597-602)."""
    b, f, e = x.data.shape
    src, ld = x.flat2d()
    if src is None:
        t = E.contiguous(x)
        src, ld = t.reshape(b, f * e), f * e
    out = K.fm_fwd(src, f, e, ld)
    res = E.Var(out.reshape(b, 1))

    def bwd(grads):
        g = grads[0].reshape(b).contiguous()
        dx = _empty((b, f * e), g)
        L.check(L.lib().b2ctr_fm_bwd(K.ptr(src), ld, f, e, K.ptr(g), K.ptr(dx), f * e, 0, b, K.stream()),
                "fm_bwd")
        E.add_grad(x, dx.reshape(x.data.shape))

    E.record([res], [x], bwd)
    return res


def add_bias(x, b):
    """x + b for a scalar / per-column bias (Linear use_bias, layers/utils.py:172-173)."""
    zeros = zeros_like_batch(x, 1)
    ones = E.Var(K.fill(torch.empty((1, 1), dtype=torch.float32, device=x.data.device), 1.0))
    # [B,1] = x + 1 * b  via the GEMM epilogue: dense([B,1] of ones) would cost the same; use add_n on a
    # broadcast built by the gemm kernel (outer product ones[B,1] x b[1,1])
    bd = b.materialize() if isinstance(b, E.Weight) else b.data
    col = torch.empty((x.data.shape[0], 1), dtype=torch.float32, device=x.data.device)
    K.fill(col, 1.0)
    colv = E.Var(col)
    bvar = b
    return add_n([x, dense(colv, _as_kernel(bvar), None, None)])


def _as_kernel(b):
    """view a [n] bias weight as a [1, n] kernel sharing storage and gradient."""
    data = b.materialize() if isinstance(b, E.Weight) else b.data
    v = E.Var(data.reshape(1, -1), requires_grad=b.requires_grad)
    v.base, v.col0, v.ncols = b, 0, -1
    v.shape_ = (1, data.numel())
    return _KernelView(v, b)


class _KernelView(E.Var):
    __slots__ = ("shape_",)

    def __init__(self, v, b):
        E.Var.__init__(self, v.data, requires_grad=b.requires_grad, base=b, col0=0, ncols=-1)
        self.shape_ = tuple(v.data.shape)

    @property
    def shape(self):
        return self.shape_
