"""CPU, world_size 2, gloo: the communication schedule of the row-sharded embedding exchange
(deepctr_b200/parallel.py) with the device kernels replaced by a torch-CPU emulation of their contracts
(b2ctr_shard_bucketize / _fill / _gather_rows / _scatter_rows, include/b2ctr.h section 7).
Checks routing (every lookup returns the right row of the right table), the gradient return path
(owner-side SGD equals the update of the unsharded table) and the dense-gradient bucket all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class CpuBackend(object):
    """Emulates the four shard kernels on CPU tensors.  `feats` is a list of int64 id tensors [B]."""

    @staticmethod
    def shard_bucketize(feats, batch, world):
        ids = torch.stack(feats, dim=1).reshape(-1)              # item = b * F + f
        owner = ids % world
        counts = torch.bincount(owner, minlength=world).to(torch.int32)
        rank = torch.zeros_like(ids)
        for d in range(world):
            m = owner == d
            rank[m] = torch.arange(int(m.sum()))
        return counts, (owner << 32) | rank

    @staticmethod
    def shard_fill(feats, batch, world, counts, slot):
        F = len(feats)
        ids = torch.stack(feats, dim=1).reshape(-1)
        off = torch.cumsum(counts.to(torch.int64), 0) - counts
        owner, rank = slot >> 32, slot & 0xffffffff
        p = off[owner] + rank
        f = torch.arange(batch * F) % F
        keys = torch.empty(batch * F, dtype=torch.int64)
        keys[p] = (f << 40) | (ids // world)
        return keys, p.to(torch.int32).reshape(batch, F)

    @staticmethod
    def shard_gather_rows(tables, lin_tables, dim, keys, n):
        rows = torch.zeros((max(n, 1), dim))
        lin = torch.zeros((max(n, 1),)) if lin_tables is not None else None
        for i in range(n):
            f, r = int(keys[i] >> 40), int(keys[i] & ((1 << 40) - 1))
            rows[i] = tables[f][r]
            if lin is not None:
                lin[i] = lin_tables[f][r]
        return rows, lin

    @staticmethod
    def shard_scatter_rows(tables, lin_tables, dim, keys, n, grows, glin, scale, lin_scale):
        for i in range(n):
            f, r = int(keys[i] >> 40), int(keys[i] & ((1 << 40) - 1))
            tables[f][r] += scale * grows[i]
            if glin is not None:
                lin_tables[f][r] += lin_scale * glin[i]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepctr_b200 import parallel
        ctx = parallel.DistContext()
        assert (ctx.rank, ctx.world, ctx.backend) == (rank, world, "gloo")
        B, F, V, E = 37 + rank, 3, 23, 4                      # ragged: different local batch per rank
        g = torch.Generator().manual_seed(7)
        full = [torch.randn(V, E, generator=g) for _ in range(F)]
        full_lin = [torch.randn(V, generator=g) for _ in range(F)]
        tabs = [parallel.shard_rows(t, rank, world).clone() for t in full]
        lins = [parallel.shard_rows(t, rank, world).clone() for t in full_lin]
        assert tabs[0].shape[0] == parallel.shard_size(V, rank, world)
        gi = torch.Generator().manual_seed(100 + rank)
        ids = [torch.randint(0, V, (B,), generator=gi) for _ in range(F)]
        ex = parallel.ShardedExchange(ctx, CpuBackend)
        st = ex.route(ids, B)
        assert st["n_send"] == B * F and sum(st["recv"]) == st["n_recv"]
        rows, lin = ex.fetch(st, tabs, lins, E)
        pos = st["pos"].long()
        for f in range(F):                                    # every lookup got the right row back
            assert torch.equal(rows[pos[:, f]], full[f][ids[f]])
            assert torch.equal(lin[pos[:, f]], full_lin[f][ids[f]])
        # backward: per-lookup gradient rows, owner-side SGD
        gg = torch.Generator().manual_seed(200 + rank)
        g_bf = torch.randn(B, F, E, generator=gg)
        gl_bf = torch.randn(B, F, generator=gg)
        grows = torch.zeros(B * F, E)
        glin = torch.zeros(B * F)
        for f in range(F):
            grows[pos[:, f]] = g_bf[:, f]
            glin[pos[:, f]] = gl_bf[:, f]
        lr = 0.1
        ex.push(st, tabs, lins, E, grows, glin, -lr, -lr)
        # reference: the unsharded tables updated with the gradients of BOTH ranks
        all_ids = [None] * world
        all_g = [None] * world
        all_gl = [None] * world
        dist.all_gather_object(all_ids, [t.tolist() for t in ids])
        dist.all_gather_object(all_g, g_bf.tolist())
        dist.all_gather_object(all_gl, gl_bf.tolist())
        for f in range(F):
            want, want_l = full[f].clone(), full_lin[f].clone()
            for r in range(world):
                idr = torch.tensor(all_ids[r][f])
                want.index_add_(0, idr, torch.tensor(all_g[r])[:, f], alpha=-lr)
                want_l.index_add_(0, idr, torch.tensor(all_gl[r])[:, f], alpha=-lr)
            torch.testing.assert_close(tabs[f], parallel.shard_rows(want, rank, world), rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(lins[f], parallel.shard_rows(want_l, rank, world), rtol=1e-5, atol=1e-6)

        # dense gradients: one flat-bucket all-reduce, averaged
        class W(object):
            pass
        ws = []
        for shape in [(3, 2), (5,), (1,)]:
            w = W()
            w.grad = torch.full(shape, float(rank + 1))
            ws.append(w)
        none = W()
        none.grad = None
        parallel.reduce_dense_grads(ctx, ws + [none],
                                    lambda src, flat, off: flat[off:off + src.numel()].copy_(src.reshape(-1)),
                                    lambda flat, f: flat.mul_(f))
        for w, shape in zip(ws, [(3, 2), (5,), (1,)]):
            assert tuple(w.grad.shape) == shape
            assert torch.allclose(w.grad, torch.full(shape, (1 + world) / 2.0))
        out.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_exchange_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]


def test_shard_helpers():
    from deepctr_b200 import parallel
    full = np.arange(23)
    parts = [parallel.shard_rows(full, r, 4) for r in range(4)]
    assert [len(p) for p in parts] == [parallel.shard_size(23, r, 4) for r in range(4)] == [6, 6, 6, 5]
    for r, p in enumerate(parts):
        assert all(v % 4 == r for v in p) and list(p // 4) == list(range(len(p)))


def _policy_worker(rank, world, port, out):
    """compile() under a world-2 process group decides, per table, between 'row-sharded + row-wise update' and
    'replicated + dense all-reduced gradient' - never 'replicated + row-wise local update' (replicas would drift)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepctr_b200 import engine as E
        from deepctr_b200.engine import SGD
        from deepctr_b200.feature_column import SparseFeat, DenseFeat, VarLenSparseFeat
        from deepctr_b200.inputs import Embedding
        from deepctr_b200.models import DeepFM
        # (1) Criteo shape: every table (and its dim-1 linear twin) is row-sharded
        E.clear_session()
        cols = [SparseFeat("C%d" % i, 41 + i, 8) for i in range(4)] + [DenseFeat("I0", 1)]
        m = DeepFM(cols, cols, dnn_hidden_units=(8,), l2_reg_linear=0, l2_reg_embedding=0)
        m.compile(SGD(0.1), "binary_crossentropy", embedding_update="sparse")
        assert m.planner.sharded
        for l in m.layers:
            if isinstance(l, Embedding):
                w = l.embeddings
                assert w.sparse_grad and w.opt_state["shard"][:2] == (rank, world), l.name
                assert w.shape[0] == (l.input_dim - rank + world - 1) // world, (l.name, w.shape)
        # (2) a pooled VarLen feature next to them: the fast tables are sharded, the sequence table (generic
        #     gather path) stays replicated and therefore takes the dense, all-reduced gradient path
        E.clear_session()
        cols2 = cols + [VarLenSparseFeat(SparseFeat("tags", 30, 8), maxlen=5, combiner="mean")]
        m2 = DeepFM(cols2, cols2, dnn_hidden_units=(8,), l2_reg_linear=0, l2_reg_embedding=0)
        m2.compile(SGD(0.1), "binary_crossentropy", embedding_update="sparse")
        kinds = {}
        for l in m2.layers:
            if isinstance(l, Embedding):
                w = l.embeddings
                sharded = "shard" in w.opt_state
                assert sharded == w.sparse_grad, (l.name, sharded, w.sparse_grad)     # never replicated + sparse
                kinds[l.name] = sharded
        assert kinds["sparse_seq_emb_tags"] is False and kinds["linear0sparse_seq_emb_tags"] is False
        dense = [w.name for w in m2.trainable_weights if not w.sparse_grad]
        assert "sparse_seq_emb_tags/embeddings" in dense          # part of the all-reduced bucket
        # (3) sparse + L2 on the tables is refused, 'auto' warns
        import warnings
        m3 = DeepFM(cols, cols, dnn_hidden_units=(8,))            # builder default l2 = 1e-5
        try:
            m3.compile(SGD(0.1), "binary_crossentropy", embedding_update="sparse")
            raise AssertionError("sparse update with an L2-regularised table must raise")
        except ValueError as exc:
            assert "L2" in str(exc)
        out.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_table_placement_policy_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_policy_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(world)) == [(0, "ok"), (1, "ok")]
