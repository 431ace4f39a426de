"""2+ ranks (torchrun): step-by-step probe of the peer-mapped sharded tables: IPC mapping, peer reads by a
torch kernel, the fused gather / scatter kernels over peer pointers, NCCL inside a captured graph."""
import faulthandler
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
faulthandler.dump_traceback_later(40, exit=True)


def say(rank, *a):
    print("[r%d %.2fs]" % (rank, time.time() - T0), *a, flush=True)


T0 = time.time()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl")
    from deepctr_b200 import _lib as L, kernels as K, parallel
    ctx = parallel.DistContext()
    V, E, F, B = 1000, 8, 3, 256
    full = [torch.arange(V * E, dtype=torch.float32).reshape(V, E) * (f + 1) for f in range(F)]
    shards = [t[rank::world].contiguous().to(dev) for t in full]
    lin_full = [torch.arange(V, dtype=torch.float32) * 0.5 * (f + 1) for f in range(F)]
    lins = [t[rank::world].contiguous().to(dev) for t in lin_full]
    torch.cuda.synchronize()
    say(rank, "mapping peers")
    pt = parallel.PeerTables(ctx, shards, L)
    pl = parallel.PeerTables(ctx, lins, L)
    say(rank, "mapped; torch kernel reads a peer shard:", float(pt.mapped[0].sum()), "expected",
        float(full[0][(rank + 1) % world::world].sum()) if world == 2 else "n/a")
    g = torch.Generator().manual_seed(7 + rank)
    ids = torch.randint(0, V, (B, F), generator=g).to(torch.int32).to(dev)
    x = torch.zeros((B, F * E), device=dev)
    linear = torch.zeros((B,), device=dev)
    feats = [K.make_feature(shards[f], ids[:, f], x, out_col=f * E, out_ld=F * E) for f in range(F)]
    plan = K.UniformPlan(feats, None, None, x, linear, None, 0)
    plan.set_peers(world, pt.table, pl.table)
    K.embed_gather_uniform_fwd(plan, B)
    torch.cuda.synchronize()
    want = torch.cat([full[f][ids[:, f].cpu().long()] for f in range(F)], 1)
    want_lin = sum(lin_full[f][ids[:, f].cpu().long()] for f in range(F))
    say(rank, "gather over peers: max err", float((x.cpu() - want).abs().max()),
        "lin err", float((linear.cpu() - want_lin).abs().max()))
    dist.barrier()
    # scatter: every rank adds its gradient rows at the owners
    dx = torch.ones((B, F * E), device=dev) * (rank + 1)
    dlin = torch.ones((B,), device=dev) * (rank + 1)
    K.embed_scatter_uniform_bwd(plan, dx, None, dlin, 1.0, 1.0, B)
    torch.cuda.synchronize()
    dist.barrier()
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    exp = [t.clone() for t in full]
    exp_lin = [t.clone() for t in lin_full]
    for r in range(world):
        idr = all_ids[r].cpu().long()
        for f in range(F):
            exp[f].index_add_(0, idr[:, f], torch.ones(B, E) * (r + 1))
            exp_lin[f].index_add_(0, idr[:, f], torch.ones(B) * (r + 1))
    err = max(float((shards[f].cpu() - exp[f][rank::world]).abs().max()) for f in range(F))
    errl = max(float((lins[f].cpu() - exp_lin[f][rank::world]).abs().max()) for f in range(F))
    say(rank, "scatter (remote red.add): max err", err, "lin", errl)
    # NCCL inside a captured graph
    tok = torch.ones((1,), device=dev)
    dist.all_reduce(tok)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, capture_error_mode="thread_local"):
        dist.all_reduce(tok)
        K.embed_gather_uniform_fwd(plan, B)
    say(rank, "captured graph with all_reduce + peer gather")
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    say(rank, "replayed: token", float(tok))
    # timing at bench shape
    V2, E2, F2, B2 = 1 << 20, 32, 26, 65536
    sh2 = [torch.randn((V2 // world, E2), device=dev) for _ in range(F2)]
    ln2 = [torch.randn((V2 // world,), device=dev) for _ in range(F2)]
    p2, l2 = parallel.PeerTables(ctx, sh2, L), parallel.PeerTables(ctx, ln2, L)
    ids2 = torch.randint(0, V2, (B2, F2), device=dev, dtype=torch.int32)
    x2 = torch.empty((B2, 848), device=dev)
    lin2 = torch.empty((B2,), device=dev)
    fm2 = torch.empty((B2,), device=dev)
    feats2 = [K.make_feature(sh2[f], ids2[:, f], x2, out_col=f * E2, out_ld=848) for f in range(F2)]
    plan2 = K.UniformPlan(feats2, None, None, x2, lin2, fm2, (1 << F2) - 1)
    plan2.set_peers(world, p2.table, l2.table)
    dx2 = torch.randn((B2, 848), device=dev)
    for name, fn in (("gather", lambda: K.embed_gather_uniform_fwd(plan2, B2)),
                     ("scatter", lambda: K.embed_scatter_uniform_bwd(plan2, dx2, fm2, lin2, -0.01, -0.01, B2))):
        for _ in range(2):
            fn()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        say(rank, "%s over %d peers at the C2 shape: %.3f ms" % (name, world, e0.elapsed_time(e1) / 10))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
