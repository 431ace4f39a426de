"""Multi-GPU data path: one process per GPU, torch.distributed for the plumbing.

* dense part (FM / CIN / Cross / attention / DNN / logit): data parallel - weights replicated, gradients
  averaged with ONE all-reduce over a flat bucket (`reduce_dense_grads`);
* embedding tables of the fused fast path: ROW-SHARDED - row r of every table (and of its dim-1 linear
  twin) lives on rank r % G as local row r // G.  Two transports:
    - 'peer' (default on one NVSwitch box, world a power of two): every rank maps every other rank's shards
      through CUDA IPC (`PeerTables`) and the ordinary fused gather / scatter kernels address the owner's
      shard directly - rows are read with NVLink peer loads, gradient rows are applied with red.add at the
      owner's L2.  No bucketing, no staging buffers, no host synchronisation; the only collective on the
      embedding path is a tiny all-reduce that separates "everyone has gathered" from "anyone scatters".
    - 'a2a' (fallback; B2CTR_SHARD_MODE=a2a): the lookups travel to their owners and the rows travel back
      with NCCL all-to-alls (ids, rows), the gradient rows return with a third, and the owner applies them
      with the fused SGD scatter (`ShardedExchange`).
  No collective is issued on a single GPU.

The reference has none of this (no sharding, no collectives: SURVEY.md section 2.1).

`ShardedExchange` holds only the communication schedule and bookkeeping; the device work is delegated to a
`backend` object (`deepctr_b200.kernels` in the product, a CPU emulation in tests/test_parallel_gloo.py so
the schedule is covered by world_size-2 gloo tests without a GPU).
"""
import torch
import torch.distributed as dist


class DistContext(object):
    def __init__(self, group=None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)


class PeerTables(object):
    """Device arrays of shard pointers [n_tables * world] (entry t * world + g = rank g's shard of table t)
    for the b2ctr_uniform_gather_t.peer_tables / peer_lin_tables fields.  Peer shards are opened through
    torch's CUDA-IPC storage sharing (cudaIpcOpenMemHandle underneath); the mapped tensors are kept alive
    here, the owners keep their shards alive as model weights."""

    def __init__(self, ctx, shards, lib):
        import os
        self.ctx = ctx
        dev = shards[0].device
        metas = []
        for t in shards:
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
            metas.append((t.untyped_storage()._share_cuda_(), t.storage_offset(), tuple(t.shape)))
        every = [None] * ctx.world
        dist.all_gather_object(every, (dev.index, metas), group=ctx.group)
        self.mapped = []          # keeps the IPC mappings alive
        ptrs = []
        for t_i, own in enumerate(shards):
            for g in range(ctx.world):
                if g == ctx.rank:
                    ptrs.append(own.data_ptr())
                    continue
                peer_dev, peer_metas = every[g]
                handle, offset, shape = peer_metas[t_i]
                if peer_dev != dev.index:
                    lib.check(lib.lib().b2ctr_enable_peer_access(peer_dev), "enable_peer_access")
                # open the handle with THIS rank's device current (first tuple element): the runtime then maps
                # the owner's memory into this device's address space (cudaIpcMemLazyEnablePeerAccess)
                storage = torch.UntypedStorage._new_shared_cuda(dev.index, *handle[1:])
                view = torch.empty(0, dtype=torch.float32, device=storage.device).set_(
                    storage, offset, shape, _contig_strides(shape))
                self.mapped.append(view)
                ptrs.append(view.data_ptr())
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        dist.barrier(group=ctx.group)     # nobody proceeds (or frees) before every rank has mapped everything

    def close(self):
        self.mapped = []
        self.table = None


def _contig_strides(shape):
    out, acc = [], 1
    for s_ in reversed(shape):
        out.append(acc)
        acc *= s_
    return tuple(reversed(out))


PEER_MAPPED_LIMIT = 384 << 30


def choose_transport(shard_bytes, world):
    """'peer' or 'a2a' for row-sharded tables of `shard_bytes` per rank.

    Peer transport maps every other rank's shards and reads / updates their rows in place over NVLink.  Measured
     with 166 GB of shards per rank: full rate on 2 GPUs (166 GB peer-mapped per rank), ~50 GB/s
    on 8 GPUs (1.16 TB peer-mapped: random rows over that footprint miss the address-translation caches on every
    access), where the NCCL all-to-all transport - owners gather / update locally, rows travel as bulk messages - is
    4.2x faster.  The switch-over sits between the two measured points."""
    return "a2a" if shard_bytes * (world - 1) > PEER_MAPPED_LIMIT else "peer"


def device_barrier(ctx, token):
    """Stream-ordered cross-rank barrier: a 4-byte all-reduce completes only when every rank has reached it
    on its stream (no host synchronisation)."""
    dist.all_reduce(token, group=ctx.group)


def shard_rows(full, rank, world):
    """Rows of a full [V, ...] table owned by `rank` (row r -> rank r % world, local row r // world)."""
    return full[rank::world]


def shard_size(vocab, rank, world):
    return (vocab - rank + world - 1) // world


class ShardedExchange(object):
    """One step's routing state for a set of F same-dim single-valued features."""

    def __init__(self, ctx, backend):
        self.ctx, self.k = ctx, backend

    # ---- forward ---------------------------------------------------------------------------------
    def route(self, feats, batch):
        """Bucket the B x F lookups by owner and exchange the keys.
        Returns a state dict with send/recv counts, the received keys and pos [B, F]."""
        G = self.ctx.world
        counts, slot = self.k.shard_bucketize(feats, batch, G)
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=self.ctx.group)
        # split sizes must be known on the host: the only synchronisation point of the step
        both = torch.stack([counts, recv_counts]).cpu()
        send = [int(v) for v in both[0]]
        recv = [int(v) for v in both[1]]
        keys, pos = self.k.shard_fill(feats, batch, G, counts, slot)
        n_recv = sum(recv)
        recv_keys = torch.empty((max(n_recv, 1),), dtype=torch.int64, device=keys.device)[:n_recv]
        dist.all_to_all_single(recv_keys, keys, output_split_sizes=recv, input_split_sizes=send,
                               group=self.ctx.group)
        return {"send": send, "recv": recv, "keys": keys, "recv_keys": recv_keys, "pos": pos,
                "n_send": sum(send), "n_recv": n_recv}

    def fetch(self, st, tables, lin_tables, dim):
        """Owners serve the received keys; rows (and linear values) travel back to the requesters.
        Returns (rows [n_send, dim], lin [n_send] | None) in the order of st['keys'] (= pos indexing)."""
        rows, lin = self.k.shard_gather_rows(tables, lin_tables, dim, st["recv_keys"], st["n_recv"])
        n_send = st["n_send"]
        back = torch.empty((max(n_send, 1), dim), dtype=torch.float32, device=rows.device)[:n_send]
        dist.all_to_all_single(back, rows[:st["n_recv"]], output_split_sizes=st["send"],
                               input_split_sizes=st["recv"], group=self.ctx.group)
        back_lin = None
        if lin is not None:
            back_lin = torch.empty((max(n_send, 1),), dtype=torch.float32, device=rows.device)[:n_send]
            dist.all_to_all_single(back_lin, lin[:st["n_recv"]], output_split_sizes=st["send"],
                                   input_split_sizes=st["recv"], group=self.ctx.group)
        return back, back_lin

    # ---- backward --------------------------------------------------------------------------------
    def push(self, st, tables, lin_tables, dim, grows, glin, scale, lin_scale):
        """Gradient rows (ordered like st['keys']) return to their owners, which apply
        table[row] += scale * g (fused SGD when scale = -lr / world)."""
        n_recv = st["n_recv"]
        g_recv = torch.empty((max(n_recv, 1), dim), dtype=torch.float32, device=grows.device)[:n_recv]
        dist.all_to_all_single(g_recv, grows, output_split_sizes=st["recv"], input_split_sizes=st["send"],
                               group=self.ctx.group)
        gl_recv = None
        if glin is not None:
            gl_recv = torch.empty((max(n_recv, 1),), dtype=torch.float32, device=grows.device)[:n_recv]
            dist.all_to_all_single(gl_recv, glin, output_split_sizes=st["recv"], input_split_sizes=st["send"],
                                   group=self.ctx.group)
        self.k.shard_scatter_rows(tables, lin_tables, dim, st["recv_keys"], n_recv, g_recv, gl_recv, scale,
                                  lin_scale)


def reduce_dense_grads(ctx, weights, copy_into, scale_into):
    """Average the gradients of the replicated weights with a single all-reduce over a flat bucket.
    `copy_into(src, flat, offset)` and `scale_into(flat, factor)` are kernel wrappers (no torch math)."""
    ws = [w for w in weights if w.grad is not None]
    if not ws or ctx.world == 1:
        return
    total = sum(w.grad.numel() for w in ws)
    flat = torch.empty((total,), dtype=torch.float32, device=ws[0].grad.device)
    off = 0
    for w in ws:
        copy_into(w.grad, flat, off)
        off += w.grad.numel()
    if ctx.backend == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=ctx.group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=ctx.group)
        scale_into(flat, 1.0 / ctx.world)
    off = 0
    for w in ws:
        n = w.grad.numel()
        w.grad = flat[off:off + n].view(w.grad.shape)
        off += n
