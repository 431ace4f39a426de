"""Run under torchrun on >= 2 GPUs (tests/test_sharded_gpu.py launches it):
DeepFM with row-sharded tables + data-parallel dense part vs the single-process CPU oracle on the
GLOBAL batch.  Exit code 0 = parity."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    import b2_helpers as H
    from oracle import models as OM, ops as O
    from deepctr_b200.engine import SGD
    from deepctr_b200.models import DeepFM

    n = 64
    rng = np.random.RandomState(5)
    cols, _, _ = H.criteo_like(rng, n, n_sparse=5, n_dense=3, vocab=37, dim=8)
    model = DeepFM(cols, cols, dnn_hidden_units=(16, 8), l2_reg_linear=0, l2_reg_embedding=0)
    H.randomize_weights(model, np.random.RandomState(6), 0.2)        # identical on every rank
    full = H.oracle_weights(model)                                   # unsharded copies
    lr = 0.1
    model.compile(SGD(lr), "binary_crossentropy", embedding_update="sparse")
    assert model.planner.sharded, "fast-path tables should be row-sharded"
    batches = []
    for r in range(world):
        rr = np.random.RandomState(100 + r)
        _, x, y = H.criteo_like(rr, n, n_sparse=5, n_dense=3, vocab=37, dim=8)
        batches.append((x, y))
    x, y = batches[rank]
    for step in range(6):     # steps >= 2 replay the captured step graph (peer mode)
        # forward parity on the local batch against the oracle with the FULL tables
        _, want = OM.deepfm(x, cols, cols, full)
        got = model.predict(x, batch_size=n)
        err = H.rel_err(got, want.detach().numpy())
        assert err < 1e-4, (rank, step, err)
        # one global step: the oracle keeps its own unsharded state and takes the same SGD step
        leaves = H.flat_params(full)
        for t in leaves.values():
            t.requires_grad_(True)
        gx = {k: np.concatenate([b[0][k] for b in batches]) for k in x}
        gy = np.concatenate([b[1] for b in batches])
        _, pred = OM.deepfm(gx, cols, cols, full)
        loss = O.binary_crossentropy(gy, pred)
        loss.backward()
        model.train_on_batch(x, y)
        scales = {}
        with torch.no_grad():
            for name, t in leaves.items():
                if t.grad is not None:
                    scales[name] = float((lr * t.grad).abs().max()) + 1e-9
                    t -= lr * t.grad
                    t.grad = None
                t.requires_grad_(False)
        mine = H.flat_params(H.oracle_weights(model))
        for name, want_t in leaves.items():
            got_t = mine[name]
            if got_t.shape != want_t.shape:                          # sharded table: compare my rows
                want_t = want_t[rank::world]
            e = float((got_t - want_t).abs().max()) / scales.get(name, 1.0)
            assert e < 3e-3, (rank, step, name, e)
    dist.barrier()
    if rank == 0:
        mode = "peer" if getattr(model.planner, "peer_mode", False) else "a2a"
        print("dist_check OK: world %d, mode %s, step graphs %d, sharded tables + DP dense match the global-batch "
              "oracle" % (world, mode, len(model._step_graphs)))
    sys.stdout.flush()
    model.close()
    dist.barrier()
    os._exit(0)       # skip NCCL / IPC teardown ordering issues at interpreter exit: parity was asserted above


if __name__ == "__main__":
    main()
