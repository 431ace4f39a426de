"""Differentiable ops on ``engine.Var``: each launches libb2ctr kernels for the forward and pushes a
closure on the active tape that launches the backward kernels.  No torch arithmetic anywhere.
"""
import numpy as np
import torch

from . import _lib as L
from . import kernels as K
from . import engine as E

# precision mode used by every dense GEMM (DNN / attention MLP / projections); see DESIGN.md section 4.2.
# Default: split-bf16 on the tcgen05 tensor cores (three bf16 MMAs per product, fp32 accumulation; relative
# error ~2^-16, inside the 1e-4 logit tolerance of the parity tests); 'fp32' selects the exact FFMA GEMM.
GEMM_PRECISION = L.GEMM_BF16X3


# BF16X3: split every GEMM operand into bf16 planes once per step and reuse them (needs MN-major operands,
# i.e. tc variant 3)
PLANE_REUSE = True
# the fused embedding gather also writes the planes of the first DNN operand (saves re-reading X to split it)
GATHER_PLANES = False    # measured: -11 us/step but +60 us inside the gather kernel itself; opt-in


# bumped by ops whose kernels take per-step by-value state (dropout seeds) or that need the host (string
# hashing): a model that ran one of them is never replayed as a CUDA graph (engine.Model._graph_eligible)
UNCAPTURABLE = 0


def mark_uncapturable():
    global UNCAPTURABLE
    UNCAPTURABLE += 1


def set_gemm_precision(mode):
    """'fp32' (exact FFMA) or 'bf16x3' (tcgen05 split-bf16, ~2^-17 relative)."""
    global GEMM_PRECISION
    GEMM_PRECISION = {"fp32": L.GEMM_FP32, "bf16x3": L.GEMM_BF16X3}[mode]


def _empty(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


def _split_k(m_out, n_out, kred):
    """wgrad-style GEMMs reduce over the batch: one K slice per CTA pair (74 pairs of SMs, 256 x 256 tiles)."""
    if kred < 4096:
        return 1
    if n_out <= 8:
        # skinny wgrad (the [*, 1] projections, CrossNet's w): an HBM stream over the batch handled by
        # skinny_tn_kernel, one CTA per (64-column block, K slice) - slices of ~256 rows, up to 4 CTAs per SM
        # (64 slices of 1024 rows left 57 % of the SMs idle: 86 us for a 16 MB read, ncu r2_launches_c2.csv)
        blocks = (m_out + 63) // 64
        return int(max(1, min(592 // blocks, kred // 256)))
    tiles = ((m_out + 255) // 256) * ((n_out + 255) // 256)
    return int(max(1, min(74 // tiles if tiles <= 74 else 1, kred // 1024)))


def _as2d(x):
    t2, ld = x.as2d()
    if t2 is None:
        t = E.contiguous(x)
        t2 = t.reshape(-1, t.shape[-1])
        ld = t2.stride(0)
    return t2, ld


# ---- GEMM based --------------------------------------------------------------------------------
def _planes_of(var, t2):
    """bf16 hi/lo planes of a Var's 2-D view, split once per step and shared by every GEMM that reads it
    (forward of each consumer + the wgrad GEMMs)."""
    base = var.base
    if (base is not None and base.xplanes is not None and var.col0 == 0 and base.data is not None
            and base.xplanes[0] == t2.shape[1] and t2.data_ptr() == base.data.data_ptr()
            and t2.stride(0) == base.data.stride(0)):
        return base.xplanes[1]       # the fused gather already wrote the planes of this window
    key = (t2.data_ptr(), tuple(t2.shape), t2.stride(0))
    if var.planes is None or var.planes[0] != key:
        var.planes = (key, K.split_planes(t2))
    return var.planes[1]


def dense(x, w, b=None, activation=None):
    """y = act(x @ w + b) over the last axis (tf.tensordot(x, w, axes=(-1, 0)) + bias_add)."""
    act = L.ACT_BY_NAME.get(activation, None)
    fused_act = act if act is not None else L.ACT_NONE
    x2, _ = _as2d(x)
    m, kdim = x2.shape
    n = w.shape[1]
    wd = w.materialize() if isinstance(w, E.Weight) else w.data
    bd = (b.materialize() if isinstance(b, E.Weight) else b.data) if b is not None else None
    # BF16X3: operands are split into bf16 planes once and the planes are reused by forward/dgrad/wgrad
    # skinny layers (the final [*, 1] projection) are GEMVs: exact-fp32 FFMA path, no tensor-core staging
    prec = GEMM_PRECISION if min(n, kdim) >= 16 else L.GEMM_FP32
    reuse = prec == L.GEMM_BF16X3 and PLANE_REUSE and m >= 128
    xp = _planes_of(x, x2) if reuse else None
    wp = K.split_planes(wd) if reuse else None
    y = K.gemm(x2, wd, bias=bd, act=fused_act, precision=prec, m=m, n=n, k=kdim, a_planes=xp, b_planes=wp)
    out = E.Var(y.reshape(tuple(x.data.shape[:-1]) + (n,)))
    if act is None and activation is not None:
        raise ValueError("activation %r cannot be fused into dense(); apply it as a layer" % activation)

    def bwd(grads):
        dy = grads[0].reshape(m, n)
        if not dy.is_contiguous():
            dy = dy.contiguous()
        need_db = b is not None and b.requires_grad
        need_planes = reuse and (x.requires_grad or w.requires_grad)
        dzp = None
        if fused_act != L.ACT_NONE or need_db:
            if need_planes and fused_act != L.ACT_NONE and dy.stride(0) % 4 == 0 and K.planes_fusable(m, n):
                dz, db, dzp = K.bias_act_bwd(dy, y, fused_act, want_dz=True, want_dbias=need_db, want_planes=True)
            else:
                dz, db = K.bias_act_bwd(dy, y, fused_act, want_dz=fused_act != L.ACT_NONE, want_dbias=need_db)
            if dz is None:
                dz = dy
        else:
            dz, db = dy, None
        if need_planes and dzp is None:
            dzp = K.split_planes(dz)
        if x.requires_grad:
            base = x.base
            if (base is not None and x.col0 == 0 and x.ncols != -1 and base.data.dim() == 2
                    and base.data.shape[0] == m and x.ncols == kdim):
                # x is the leading window of a K-padded buffer: write dx with the same ld so that
                # the buffer's gradient is adopted without a copy
                ld = base.data.stride(0)
                buf = _empty((m, ld), dy)
                dxw = buf[:, :kdim]
                K.gemm(dz, wd, c=dxw, trans_b=True, precision=prec, m=m, n=kdim, k=n, a_planes=dzp,
                       b_planes=wp)
                E.add_grad(x, dxw)
            else:
                dx = K.gemm(dz, wd, trans_b=True, precision=prec, m=m, n=kdim, k=n, a_planes=dzp,
                            b_planes=wp)
                E.add_grad(x, dx.reshape(x.data.shape))
        if w.requires_grad:
            dw = K.gemm(x2, dz, trans_a=True, precision=prec, split_k=_split_k(kdim, n, m),
                        m=kdim, n=n, k=m, a_planes=xp, b_planes=dzp)
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
    shapes = [tuple(v.shape) for v in xs]
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
            if ok and base.data is None:     # virtual buffer (fused away): stay virtual
                return E.Var(None, base=base, col0=xs[0].col0, ncols=col - xs[0].col0, owner=base.owner,
                             vshape=tuple(out_shape))
            if ok:
                return _window(base, xs[0].col0, col - xs[0].col0, tuple(out_shape))
    if any(v.data is None for v in xs):
        raise L.B2ctrError("a fused-away (virtual) embedding output reached a layer that needs its values")
    if flat_ok:
        # per-sample flat concatenation of non-adjacent windows: strided 2-D copies, no densifying pass
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
    """x + b over the last axis for a per-column (or scalar, when the last axis is 1) bias:
    Linear(use_bias=True) without a dense part, layers/utils.py:172-173.  Forward = copy + rank-1 update
    ones[B,1] @ b[1,n] through the skinny GEMM kernel; backward = pass-through and a column sum."""
    xt = E.contiguous(x)
    x2 = xt.reshape(-1, xt.shape[-1])
    m, n = x2.shape
    bd = b.materialize() if isinstance(b, E.Weight) else b.data
    if bd.numel() != n:
        raise ValueError("add_bias: bias has %d elements, last axis has %d" % (bd.numel(), n))
    y = K.add_n([x2])                                     # copy
    ones = K.fill(torch.empty((m, 1), dtype=torch.float32, device=y.device), 1.0)
    K.gemm(ones, bd.reshape(1, n), c=y, accumulate=True, m=m, n=n, k=1)
    out = E.Var(y.reshape(x.data.shape))

    def bwd(grads):
        dy = grads[0]
        if x.requires_grad:
            E.add_grad(x, dy)
        if b.requires_grad:
            dy2 = dy.reshape(m, n)
            if not dy2.is_contiguous():
                dy2 = dy2.contiguous()
            _, db = K.bias_act_bwd(dy2, None, L.ACT_NONE, want_dz=False, want_dbias=True)
            E.add_grad(b, db.reshape(bd.shape))

    E.record([out], [x, b], bwd)
    return out


# ==================================================================================================
# interaction operators
# ==================================================================================================
def _wdata(w):
    return w.materialize() if isinstance(w, E.Weight) else w.data


def _rows2d(x):
    """[B, d] window -> (tensor2d, ld)."""
    t, ld = x.flat2d()
    if t is None:
        t = E.contiguous(x).reshape(x.shape[0], -1)
        ld = t.stride(0)
    return t, ld


def cross_vector(x0, xl, w, bias):
    """x_{l+1} = x_0 * (x_l . w) + b + x_l   (layers/interaction.py:413-416)."""
    t0, ld0 = _rows2d(x0)
    tl, ldl = _rows2d(xl)
    b, d = t0.shape
    wd, bd = _wdata(w), _wdata(bias)
    out, s = K.cross_vector_fwd(t0, ld0, tl, ldl, wd, bd, b, d)
    res = E.Var(out)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        dx0, dxl, ds = K.cross_vector_bwd(t0, ld0, wd, g, s, b, d)
        E.add_grad(x0, dx0)
        E.add_grad(xl, dxl)
        if w.requires_grad:
            dw = K.gemm(tl, ds.reshape(b, 1), trans_a=True, split_k=_split_k(d, 1, b), m=d, n=1, k=b)
            E.add_grad(w, dw.reshape(w.shape))
        if bias.requires_grad:
            _, db = K.bias_act_bwd(g, None, L.ACT_NONE, want_dz=False, want_dbias=True)
            E.add_grad(bias, db.reshape(bias.shape))

    E.record([res], [x0, xl, w, bias], bwd)
    return res


def cross_matrix(x0, xl, w, bias):
    """x_{l+1} = x_0 * (W x_l + b) + x_l   (layers/interaction.py:417-420)."""
    t0, ld0 = _rows2d(x0)
    tl, ldl = _rows2d(xl)
    b, d = t0.shape
    wd, bd = _wdata(w), _wdata(bias).reshape(-1)
    c0 = E.contiguous(x0).reshape(b, d) if ld0 != d else t0
    cl = E.contiguous(xl).reshape(b, d) if ldl != d else tl
    u = K.gemm(cl, wd, bias=bd, trans_b=True, precision=GEMM_PRECISION, m=b, n=d, k=d)     # W x_l + b
    out = K.ewise(1, c0, u, cl)                                                           # x0*u + xl
    res = E.Var(out)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        du = K.ewise(0, g, c0)
        E.add_grad(x0, K.ewise(0, g, u))
        dxl = K.gemm(du, wd, precision=GEMM_PRECISION, m=b, n=d, k=d)                     # du @ W
        K.axpy(g, dxl, 1.0)
        E.add_grad(xl, dxl)
        if w.requires_grad:
            dw = K.gemm(du, cl, trans_a=True, precision=GEMM_PRECISION, split_k=_split_k(d, d, b), m=d, n=d, k=b)
            E.add_grad(w, dw)
        if bias.requires_grad:
            _, db = K.bias_act_bwd(du, None, L.ACT_NONE, want_dz=False, want_dbias=True)
            E.add_grad(bias, db.reshape(bias.shape))

    E.record([res], [x0, xl, w, bias], bwd)
    return res


CIN_CHUNK_BYTES = 48 << 20      # outer-product chunk kept well inside the 126 MB L2


CIN_FUSED = True              # generate the outer product inside the tcgen05 GEMM producer (b2ctr_cin_gemm)
CIN_FOLD = True               # ... and fold dZ = dY W^T onto the factors inside the GEMM epilogue (b2ctr_cin_fold)
CIN_DZ_CHUNK_BYTES = 512 << 20


def cin(x, filters, biases, layer_size, activation, split_half):
    B, m, D = x.shape
    with K.profile_tag("cin"):
        hid = [n // 2 if split_half else n for n in layer_size[:-1]]
        if (CIN_FUSED and GEMM_PRECISION == L.GEMM_BF16X3 and B * D >= 256 and D in (4, 8, 16, 32, 64, 128)
                and min(layer_size) >= 8 and all(n % 4 == 0 for n in layer_size) and all(h % 4 == 0 for h in hid)):
            return _cin_fused(x, filters, biases, layer_size, activation, split_half)
        return _cin(x, filters, biases, layer_size, activation, split_half)


def _cin_pad(h):
    return 32 if h <= 32 else (h + 63) // 64 * 64


def _cin_fused(x, filters, biases, layer_size, activation, split_half):
    """CIN (layers/interaction.py:277-325) with the outer product Z[(b,d), (i,j)] = X0(b,i,d) X_k(b,j,d) GENERATED
    by the producer warps of the tensor-core GEMM instead of being written anywhere: forward Y = act(Z W + b) and
    the filter gradient dW = Z^T dY both read only the two factors (T0 = X0 transposed to [(b,d), m], X_k as the
    previous layer's [(b,d), N] activations).  Only dZ = dY W^T of the backward pass exists in memory, in row chunks."""
    B, m, D = x.shape
    x2, ldx = x.flat2d()
    if x2 is None:
        x2 = E.contiguous(x).reshape(B, m * D)
        ldx = m * D
    act = L.ACT_BY_NAME[activation]
    nl = len(layer_size)
    hs = [m] + [size // 2 if split_half else size for size in layer_size]
    direct = [((size // 2, size // 2) if (split_half and i != nl - 1) else (0, size)) for i, size in enumerate(layer_size)]
    out_cols = sum(nc for _, nc in direct)
    out = _empty((B, out_cols), x2)
    rows = B * D
    v0 = (ldx, D, 1)
    ld0 = 32 if m <= 32 else (m + 63) // 64 * 64
    t0 = K.cin_t0(x2, v0, B, m, D, ld0)                       # [rows, ld0]
    tv0 = (D * ld0, 1, ld0)                                   # T0 seen as X0(b,i,d)
    ws2d = [_wdata(f).reshape(-1, f.shape[-1]) for f in filters]
    bs = [_wdata(bv) for bv in biases]
    hps = [_cin_pad(hs[i]) for i in range(nl)]
    wplanes, ys = [], []
    oc = 0
    for i, size in enumerate(layer_size):
        h, hp = hs[i], hps[i]
        xk, ldk = (t0, ld0) if i == 0 else (ys[-1], layer_size[i - 1])
        wp = K.cin_filter_planes(ws2d[i], m, h, hp)
        y = K.cin_gemm(0, t0, xk, ldk, rows, m, h, hp, size, wp, bias=bs[i], act=act)
        K.cin_sum_d(y, size, direct[i][0], direct[i][1], D, out, out_cols, oc, 0, B)
        oc += direct[i][1]
        wplanes.append(wp)
        ys.append(y)
    res = E.Var(out)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        gx = (m * D, D, 1)
        fold = CIN_FOLD and all(hp in (32, 64, 128) for hp in hps)
        dx = _empty((B, m * D), x2)
        if fold:
            dt0 = _empty((rows, ld0), x2)          # gradient of the factor table, all layers accumulate into it
            K.fill(dt0, 0.0)
        else:
            K.fill(dx, 0.0)
        dh = None
        col = out_cols
        for i in range(nl - 1, -1, -1):
            size, h, hp = layer_size[i], hs[i], hps[i]
            kq = m * hp
            col -= direct[i][1]
            dy = _empty((rows, size), x2)
            K.cin_expand_grad(g, out_cols, col, direct[i][0], direct[i][1], dh, hs[i + 1] if dh is not None else 0,
                              hs[i + 1] if dh is not None else 0, dy, size, D, 0, B)
            if act != L.ACT_NONE and K.planes_fusable(rows, size):
                dz_, db, dzp = K.bias_act_bwd(dy, ys[i], act, want_dz=True, want_dbias=True, want_planes=True)
            else:
                dz_, db = K.bias_act_bwd(dy, ys[i], act, want_dz=act != L.ACT_NONE, want_dbias=True)
                if dz_ is None:
                    dz_ = dy
                dzp = K.split_planes(dz_)
            E.add_grad(biases[i], db)
            xk, ldk = (t0, ld0) if i == 0 else (ys[i - 1], layer_size[i - 1])
            xkv = tv0 if i == 0 else (D * layer_size[i - 1], 1, layer_size[i - 1])
            if filters[i].requires_grad:
                dwp = K.cin_gemm(1, t0, xk, ldk, rows, m, h, hp, size, dzp, split_k=_split_k(kq, size, rows))
                E.add_grad(filters[i], K.cin_unpad_rows(dwp, m, h, hp).reshape(filters[i].shape))
            dhid = None
            if fold:
                # dZ = dY W'^T exists only as TMEM tiles: the GEMM epilogue folds it onto T0 and X_k
                if i > 0:
                    dhid = _empty((rows, h), x2)
                    K.fill(dhid, 0.0)
                K.cin_fold(t0, xk, ldk, rows, m, h, hp, size, wplanes[i], dzp, dt0, dt0 if i == 0 else dhid,
                           ld0 if i == 0 else h)
            else:
                # dZ in row chunks, folded back onto the two factors by a second kernel
                dhid = _empty((rows, h), x2) if i > 0 else None
                chunk = max(256, min(rows, CIN_DZ_CHUNK_BYTES // (4 * kq)) // 256 * 256)
                dzf = _empty((min(chunk, rows), kq), x2)
                for r0 in range(0, rows, chunk):
                    nr = min(chunk, rows - r0)
                    K.gemm(dz_[r0:r0 + nr], dz_, c=dzf[:nr], trans_b=True, precision=L.GEMM_BF16X3, m=nr, n=kq, k=size,
                           b_planes=wplanes[i])
                    b0, nbk = r0 // D, nr // D
                    if i == 0:
                        K.cin_outer_bwd(dzf, t0, tv0, t0, tv0, dx, gx, True, dx, gx, True, b0, nbk, m, h, D, hp)
                    else:
                        K.cin_outer_bwd(dzf, t0, tv0, ys[i - 1], xkv, dx, gx, True, dhid, (D * h, 1, h), False, b0, nbk,
                                        m, h, D, hp)
            dh = dhid
        if fold:
            K.cin_t0_bwd(dt0, ld0, dx, gx, False, B, m, D)
        E.add_grad(x, dx.reshape(x.shape))

    E.record([res], [x] + list(filters) + list(biases), bwd)
    return res


def _cin(x, filters, biases, layer_size, activation, split_half):
    """Compressed Interaction Network (layers/interaction.py:277-325).  Per batch chunk the outer
    product Z[(b,d), i*H+j] lives in an L2-sized scratch buffer and is contracted with the filter by
    b2ctr_gemm; layer outputs are kept as [B, D, N] so the next layer reads them through strides."""
    B, m, D = x.shape
    x2, ldx = x.flat2d()
    if x2 is None:
        x2 = E.contiguous(x).reshape(B, m * D)
        ldx = m * D
    act = L.ACT_BY_NAME[activation]
    nl = len(layer_size)
    hs = [m]
    for i, size in enumerate(layer_size):
        hs.append(size // 2 if split_half else size)
    direct = []            # (col0, ncols) of each layer's direct maps
    for i, size in enumerate(layer_size):
        if split_half and i != nl - 1:
            direct.append((size // 2, size // 2))
        else:
            direct.append((0, size))
    out_cols = sum(nc for _, nc in direct)
    out = _empty((B, out_cols), x2)
    v0 = (ldx, D, 1)
    ys = []                # per layer activations [B*D, N]
    kmax = max(m * hs[i] for i in range(nl))
    chunk = max(1, min(B, CIN_CHUNK_BYTES // (4 * D * kmax)))
    z = _empty((chunk * D, kmax), x2)
    ws = [_wdata(f).reshape(-1, f.shape[-1]) for f in filters]
    bs = [_wdata(bv) for bv in biases]
    oc = 0
    for i, size in enumerate(layer_size):
        h = hs[i]
        kdim = m * h
        y = _empty((B * D, size), x2)
        src, vk = (x2, v0) if i == 0 else (ys[-1], (D * layer_size[i - 1], 1, layer_size[i - 1]))
        for b0 in range(0, B, chunk):
            nb = min(chunk, B - b0)
            zc = z.reshape(-1)[:nb * D * kdim].reshape(nb * D, kdim)
            K.cin_outer_fwd(x2, v0, src, vk, zc, b0, nb, m, h, D)
            K.gemm(zc, ws[i], c=y[b0 * D:(b0 + nb) * D], bias=bs[i], act=act, precision=GEMM_PRECISION,
                   m=nb * D, n=size, k=kdim)
        K.cin_sum_d(y, size, direct[i][0], direct[i][1], D, out, out_cols, oc, 0, B)
        oc += direct[i][1]
        ys.append(y)
    res = E.Var(out)

    def bwd(grads):
        with K.profile_tag("cin"):
            _bwd(grads)

    def _bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        dx = _empty((B, m * D), x2)
        K.fill(dx, 0.0)
        gx = (m * D, D, 1)
        dh = None                                    # gradient wrt the hidden maps feeding layer i+1
        col = out_cols
        for i in range(nl - 1, -1, -1):
            size, h = layer_size[i], hs[i]
            kdim = m * h
            col -= direct[i][1]
            dy = _empty((B * D, size), x2)
            K.cin_expand_grad(g, out_cols, col, direct[i][0], direct[i][1], dh, hs[i + 1] if dh is not None else 0,
                              hs[i + 1] if dh is not None else 0, dy, size, D, 0, B)
            dz_, db = K.bias_act_bwd(dy, ys[i], act, want_dz=act != L.ACT_NONE, want_dbias=True)
            if dz_ is None:
                dz_ = dy
            E.add_grad(biases[i], db)
            src, vk = (x2, v0) if i == 0 else (ys[i - 1], (D * layer_size[i - 1], 1, layer_size[i - 1]))
            dhid = None
            if i > 0:
                dhid = _empty((B * D, h), x2)        # grad of the first h maps of layer i-1, [B, D, h]
            dw = _empty((kdim, size), x2)
            K.fill(dw, 0.0)
            for b0 in range(0, B, chunk):
                nb = min(chunk, B - b0)
                zc = z.reshape(-1)[:nb * D * kdim].reshape(nb * D, kdim)
                K.cin_outer_fwd(x2, v0, src, vk, zc, b0, nb, m, h, D)          # recompute Z (stays in L2)
                dzc = dz_[b0 * D:(b0 + nb) * D]
                K.gemm(zc, dzc, c=dw, trans_a=True, accumulate=True, precision=GEMM_PRECISION,
                       m=kdim, n=size, k=nb * D)
                dzf = K.gemm(dzc, ws[i], c=zc, trans_b=True, precision=GEMM_PRECISION, m=nb * D, n=kdim, k=size)
                if i == 0:
                    # X_k is X_0 itself: both factors accumulate into dx
                    K.cin_outer_bwd(dzf, x2, v0, src, vk, dx, gx, True, dx, gx, True, b0, nb, m, h, D)
                else:
                    K.cin_outer_bwd(dzf, x2, v0, src, vk, dx, gx, True, dhid, (D * h, 1, h), False, b0, nb,
                                    m, h, D)
            E.add_grad(filters[i], dw.reshape(filters[i].shape))
            dh = dhid
        E.add_grad(x, dx.reshape(x.shape))

    E.record([res], [x] + list(filters) + list(biases), bwd)
    return res


def interacting_attention(q, k, v, res, heads, dhead, scaling):
    """softmax(q_h k_h^T) v_h (+ res) -> relu, per sample and head (layers/interaction.py:760-777)."""
    B, F, HD = q.shape
    qt, kt, vt = E.contiguous(q), E.contiguous(k), E.contiguous(v)
    rt = E.contiguous(res) if res is not None else None
    out = K.interacting_fwd(qt, kt, vt, rt, B, F, heads, dhead, scaling)
    o = E.Var(out)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        dq, dk, dv, dres = K.interacting_bwd(qt, kt, vt, out, g, res is not None, B, F, heads, dhead, scaling)
        E.add_grad(q, dq)
        E.add_grad(k, dk)
        E.add_grad(v, dv)
        if res is not None:
            E.add_grad(res, dres)

    E.record([o], [q, k, v, res], bwd)
    return o


# ==================================================================================================
# sequence operators
# ==================================================================================================
def din_att_input(query, keys):
    """[q, k, q-k, q*k] along the last axis (layers/core.py:98-101)."""
    B, T, Edim = keys.shape
    qt, ldq = query.flat2d()
    if qt is None:
        qt = E.contiguous(query).reshape(B, Edim)
        ldq = Edim
    kt, ldk = keys.flat2d()
    if kt is None:
        kt = E.contiguous(keys).reshape(B, T * Edim)
        ldk = T * Edim
    out = K.din_att_input_fwd(qt, ldq, kt, ldk, B, T, Edim)
    res = E.Var(out)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        dq, dk = K.din_att_input_bwd(qt, ldq, kt, ldk, g, B, T, Edim)
        E.add_grad(query, dq.reshape(query.shape))
        E.add_grad(keys, dk.reshape(keys.shape))

    E.record([res], [query, keys], bwd)
    return res


DIN_FUSED = True     # generate [q, k, q-k, q*k] inside the first attention GEMM (b2ctr_att_gemm)


def din_att_fusable(query, keys, n_out):
    B, T, Edim = keys.shape
    return (DIN_FUSED and GEMM_PRECISION == L.GEMM_BF16X3 and Edim % 8 == 0 and B * T >= 256 and n_out >= 8
            and n_out % 4 == 0)


def din_att_first(query, keys, w, b, activation):
    """act([q, k, q-k, q*k] @ w + b) -> [B, T, n] without materialising the [B, T, 4E] attention input
    (layers/core.py:96-103).  The backward pass generates the same operand again for the kernel gradient; only
    d(input) = dZ W^T exists in memory, to be folded back onto q and k."""
    B, T, Edim = keys.shape
    qt, ldq = query.flat2d()
    if qt is None or ldq % 4 or qt.data_ptr() % 16:
        qt = E.contiguous(query).reshape(B, Edim)
        ldq = Edim
    kt, ldk = keys.flat2d()
    if kt is None or ldk % 4 or kt.data_ptr() % 16:
        kt = E.contiguous(keys).reshape(B, T * Edim)
        ldk = T * Edim
    act = L.ACT_BY_NAME[activation]
    wd, bd = _wdata(w), (_wdata(b) if b is not None else None)
    n = wd.shape[1]
    wp = K.split_planes(wd)
    y = K.att_gemm(0, qt, ldq, kt, ldk, B, T, Edim, n, wp, bias=bd, act=act)
    out = E.Var(y.reshape(B, T, n))
    rows = B * T

    def bwd(grads):
        dy = grads[0].reshape(rows, n)
        if not dy.is_contiguous():
            dy = dy.contiguous()
        need_db = b is not None and b.requires_grad
        dzp = None
        if act != L.ACT_NONE or need_db:
            if act != L.ACT_NONE and K.planes_fusable(rows, n):
                dz, db, dzp = K.bias_act_bwd(dy, y, act, want_dz=True, want_dbias=need_db, want_planes=True)
            else:
                dz, db = K.bias_act_bwd(dy, y, act, want_dz=act != L.ACT_NONE, want_dbias=need_db)
            if dz is None:
                dz = dy
        else:
            dz, db = dy, None
        if dzp is None:
            dzp = K.split_planes(dz)
        if w.requires_grad:
            dw = K.att_gemm(1, qt, ldq, kt, ldk, B, T, Edim, n, dzp, split_k=_split_k(4 * Edim, n, rows))
            E.add_grad(w, dw)
        if need_db:
            E.add_grad(b, db)
        if query.requires_grad or keys.requires_grad:
            da = K.gemm(dz, wd, trans_b=True, precision=L.GEMM_BF16X3, m=rows, n=4 * Edim, k=n, a_planes=dzp, b_planes=wp)
            dq, dk = K.din_att_input_bwd(qt, ldq, kt, ldk, da, B, T, Edim)
            E.add_grad(query, dq.reshape(query.shape))
            E.add_grad(keys, dk.reshape(keys.shape))

    E.record([out], [query, keys, w, b], bwd)
    return out


def din_attention_pool(score, keys, mask_u8, weight_normalization, return_score):
    """masked fill -> [softmax] -> score @ keys   (layers/sequence.py:278-291)."""
    B, T, Edim = keys.shape
    st = E.contiguous(score).reshape(B, T)
    kt, ldk = keys.flat2d()
    if kt is None:
        kt = E.contiguous(keys).reshape(B, T * Edim)
        ldk = T * Edim
    out, w = K.din_pool_fwd(st, kt, ldk, mask_u8, B, T, Edim, weight_normalization, return_score)
    res = E.Var(out)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        dscore, dkeys = K.din_pool_bwd(w, kt, ldk, mask_u8, g, B, T, Edim, weight_normalization, return_score,
                                       want_dkeys=keys.requires_grad)
        E.add_grad(score, dscore.reshape(score.shape))
        if dkeys is not None:
            E.add_grad(keys, dkeys.reshape(keys.shape))

    E.record([res], [score, keys], bwd)
    return res


def _len_i32(lengths):
    t = lengths.data
    if t.dtype != torch.int32:
        raise ValueError("sequence lengths must be int32")
    return dense_i32(t).reshape(-1)


def dense_i32(t):
    """Contiguous copy of a strided int32 [B, W] window through the copy2d kernel (bit-exact 4-byte moves)."""
    if t.is_contiguous():
        return t
    t2 = t.reshape(t.shape[0], -1) if t.dim() != 2 else t
    out = torch.empty(t2.shape, dtype=torch.int32, device=t.device)
    K.copy2d(t2.view(torch.float32), t2.stride(0), out.view(torch.float32), out.stride(0), t2.shape[0], t2.shape[1])
    return out


def seqpool(seq, mode, mask_u8=None, lengths=None):
    B, T, Edim = seq.shape
    xt = E.contiguous(seq)
    ln = _len_i32(lengths) if lengths is not None else None
    code = L.POOL_BY_NAME[mode]
    out = K.seqpool_fwd(xt, mask_u8, ln, B, T, Edim, code)
    res = E.Var(out)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        E.add_grad(seq, K.seqpool_bwd(xt, mask_u8, ln, g, B, T, Edim, code))

    E.record([res], [seq], bwd)
    return res


def weighted_seq(seq, weights, normalize, mask_u8=None, lengths=None):
    B, T, Edim = seq.shape
    xt = E.contiguous(seq)
    wt_in = E.contiguous(weights).reshape(B, T)
    ln = _len_i32(lengths) if lengths is not None else None
    wt = K.seqweight(wt_in, mask_u8, ln, B, T, normalize)
    out = K.seqscale(xt, wt, B * T, Edim)
    res = E.Var(out, mask=seq.mask)

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        E.add_grad(seq, K.seqscale(g, wt, B * T, Edim))     # weights are inputs: no gradient needed

    E.record([res], [seq], bwd)
    return res


def _stats_for(x2, m, n, mean_w, var_w, training, momentum):
    if training:
        stats = K.colstats(x2, n, m, n)
        K.moving_update(mean_w.materialize(), stats[0], momentum)
        K.moving_update(var_w.materialize(), stats[1], momentum)
        return stats[0], stats[1]
    return mean_w.materialize(), var_w.materialize()


def dice(x, alphas, moving_mean, moving_var, eps, training, momentum=0.99):
    """Dice (layers/activation.py:59-64) over the last axis; batch statistics over all leading axes."""
    xt = E.contiguous(x)
    n = xt.shape[-1]
    m = xt.numel() // n
    x2 = xt.reshape(m, n)
    mean, var = _stats_for(x2, m, n, moving_mean, moving_var, training, momentum)
    al = alphas.materialize()
    y = K.dice_fwd(x2, mean, var, al, m, n, eps)
    res = E.Var(y.reshape(xt.shape))

    def bwd(grads):
        g = grads[0].reshape(m, n)
        if not g.is_contiguous():
            g = g.contiguous()
        dx, dalpha = K.dice_bwd(x2, mean, var, al, g, m, n, eps, training)
        E.add_grad(x, dx.reshape(x.shape))
        E.add_grad(alphas, dalpha)

    E.record([res], [x, alphas], bwd)
    return res


def batchnorm(x, gamma, beta, moving_mean, moving_var, eps, training, momentum=0.99):
    xt = E.contiguous(x)
    n = xt.shape[-1]
    m = xt.numel() // n
    x2 = xt.reshape(m, n)
    mean, var = _stats_for(x2, m, n, moving_mean, moving_var, training, momentum)
    gd = gamma.materialize() if gamma is not None else None
    bd = beta.materialize() if beta is not None else None
    y = K.bn_apply(x2, mean, var, gd, bd, m, n, eps)
    res = E.Var(y.reshape(xt.shape))

    def bwd(grads):
        g = grads[0].reshape(m, n)
        if not g.is_contiguous():
            g = g.contiguous()
        dx, dgamma, dbeta = K.bn_bwd(x2, mean, var, gd, g, m, n, eps, training)
        E.add_grad(x, dx.reshape(x.shape))
        if gamma is not None:
            E.add_grad(gamma, dgamma)
        if beta is not None:
            E.add_grad(beta, dbeta)

    E.record([res], [x, gamma, beta], bwd)
    return res


def dropout(x, rate, seed):
    mark_uncapturable()
    xt = E.contiguous(x)
    res = E.Var(K.dropout(xt, rate, seed))

    def bwd(grads):
        g = grads[0]
        if not g.is_contiguous():
            g = g.contiguous()
        E.add_grad(x, K.dropout(g, rate, seed))

    E.record([res], [x], bwd)
    return res


def reduce(x, kind, axis, keep_dims):
    """reduce_sum / reduce_mean / reduce_max shims of layers/utils.py:245-303 for the common cases."""
    if kind == "sum" and (axis in (-1, x.data.dim() - 1)) and x.data.dim() == 2:
        r = rowsum(x)
        return r if keep_dims else reshape(r, (x.shape[0],))
    raise NotImplementedError("reduce_%s over axis %r of a %d-D tensor" % (kind, axis, x.data.dim()))


def div(x, y):
    raise NotImplementedError("div is only used inside SequencePoolingLayer, which has its own kernel")


def softmax(x, dim=-1):
    raise NotImplementedError("softmax is only used inside attention layers, which have their own kernels")
