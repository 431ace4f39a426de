"""Per-tile vs per-k-block cost of the persistent tcgen05 GEMM: C[M, N] = A[M, K] B[K, N] at M = 65536 (256 CTA-pair row
tiles), K swept.  Run once per B2CTR_TC_DEBUG value (knock-outs: 1 no epilogue stores, 2 no tcgen05.ld, 4 no MMAs,
8 no operand loads - results are wrong then, only the time matters) to attribute the cost.
usage: [B2CTR_TC_DEBUG=d] python tools/gemm_sweep.py"""
import os
import sys

import torch
sys.path.insert(0, ".")
from deepctr_b200 import _lib as L, kernels as K

SHAPES = [(65536, 128, k) for k in (64, 128, 256, 512, 1024, 2048)] + \
         [(65536, 256, k) for k in (64, 256, 1024)] + \
         [(65536, 845, 256), (409600, 40, 80), (409600, 80, 256), (65536, 64, 128)]


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dbg = os.environ.get("B2CTR_TC_DEBUG", "0")
    out = []
    for m, n, k in SHAPES:
        a = torch.randn((m, k), device=dev)
        b = torch.randn((k, n), device=dev)
        ap, bp = K.split_planes(a), K.split_planes(b)
        cbuf = torch.empty((m, (n + 7) // 8 * 8), device=dev)
        cview = cbuf[:, :n]

        def run():
            return K.gemm(a, b, c=cview, precision=L.GEMM_BF16X3, m=m, n=n, k=k, a_planes=ap, b_planes=bp)
        for _ in range(3):
            run()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        out.append("%dx%dx%d=%.1f" % (m, n, k, ts[len(ts) // 2]))
    print("debug=%s us: %s" % (dbg, "  ".join(out)))


if __name__ == "__main__":
    main()
