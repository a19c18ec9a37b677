"""Per-shape timing of toc3d_gemm_bf16 (CUDA events, L2 flushed between launches) for the shapes of the
ToC3D_fast forward + one large square reference shape.  Diagnostic only (not a bench line).

    python tools/gemm_bench.py [--reps 20]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from toc3d_b200 import lib as L  # noqa: E402
from toc3d_b200.backbone import hidden_pad, interleave_w12  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--no-flush", action="store_true")
args = ap.parse_args()
dev = "cuda"
L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.reps):
        if not args.no_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


def report(name, M, N, K, fn):
    med, best = timeit(fn)
    fl = 2.0 * M * N * K
    print("%-28s M=%6d N=%5d K=%5d  median %7.1f us %7.1f TF/s   best %7.1f us %7.1f TF/s" % (
        name, M, N, K, med, fl / med / 1e6, best, fl / best / 1e6), flush=True)


C, Hd = 1024, 2730
Hp = hidden_pad(Hd)
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
for M in (8640, 6192, 6000, 5058, 3618, 12288, 7200):
    A = rn(M, C).bfloat16()
    Wqkv = (rn(3 * C, C) * 0.02).bfloat16(); bq = rn(3 * C)
    qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    cos = rn(16, 16); sin = rn(16, 16)
    report("qkv+rope", M, 3 * C, C, lambda: L.gemm(A, Wqkv, L.EPI_QKV_ROPE, bias=bq, out=qkv, rope_slots=256,
                                                   rope_ft=16, rope_cols=2 * C, q_scale=0.125, cos_axis=cos, sin_axis=sin))
    Wp = (rn(C, C) * 0.02).bfloat16(); bp = rn(C)
    X = rn(M, C); T = torch.empty(M, C, device=dev)
    report("proj+resid", M, C, C, lambda: L.gemm(A, Wp, L.EPI_RESID, bias=bp, out=T, resid=X))
    W12 = (rn(2 * Hp, C) * 0.02).bfloat16(); b12 = rn(2 * Hp)
    hid = torch.empty(M, Hp, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(M, 2, device=dev, dtype=torch.int64)
    report("w12 swiglu+stats", M, 2 * Hp, C, lambda: L.gemm(A, W12, L.EPI_SWIGLU, bias=b12, out=hid, row_stats=stats))
    W3 = (rn(C, Hp) * 0.02).bfloat16(); u3 = rn(C)
    report("w3 ln-fold+resid", M, C, Hp, lambda: L.gemm(hid, W3, L.EPI_RESID, bias=bp, out=X, resid=T, row_stats=stats,
                                                       ln_u=u3, ln_n=Hd, ln_eps=1e-6))
    ob = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    report("linear bf16 (N=1024)", M, C, C, lambda: L.gemm(A, Wp, L.EPI_LINEAR, bias=bp, out=ob))
for n in (4096, 8192):
    A = rn(n, n).bfloat16(); B = rn(n, n).bfloat16(); o = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
    report("square linear", n, n, n, lambda: L.gemm(A, B, L.EPI_LINEAR, out=o))
    report("torch.matmul (cuBLAS)", n, n, n, lambda: torch.matmul(A, B.t()))
