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
from bench import ClockSampler  # noqa: E402
from toc3d_b200.backbone import hidden_pad, interleave_w12  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--tiles", default="0", help="comma list of tile_n values to sweep (0 = auto)")
ap.add_argument("--ms", default="8640,6192,6000,5058,3618,12288,7200")
ap.add_argument("--square", action="store_true")
args = ap.parse_args()
dev = "cuda"
if os.environ.get("TOC3D_LIB"):                # A/B of two builds of the library
    L.LIB_PATH = os.environ["TOC3D_LIB"]
L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    """Median over `reps` trains of 10 back-to-back launches (host launch latency hidden behind a device-side
    sleep), per-launch time in us; the L2 flush (if enabled) happens once before each train."""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.reps):
        if not args.no_flush:
            flush.zero_()
        torch.cuda._sleep(2_000_000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / 10)
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


TILES = [int(t) for t in args.tiles.split(",")]


def report(name, M, N, K, fn):
    """fn(tile_n) launches the GEMM."""
    fl = 2.0 * M * N * K
    cells = []
    for t in TILES:
        if t and ((name.startswith("w12") and t % 64) or (t > 256 and "resid" not in name)):
            cells.append("%3d:    -    " % t)
            continue
        med, best = timeit(lambda: fn(t))
        cells.append("%3d:%6.1fus %4.0f" % (t, med, fl / med / 1e6))
    print("%-22s M=%6d N=%5d K=%5d  %s" % (name, M, N, K, " | ".join(cells)), flush=True)


clk = ClockSampler(0).__enter__()      # same nvidia-smi sampler as bench.py: the clocks the numbers were taken at
C, Hd = 1024, 2730
Hp = hidden_pad(Hd)
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
for M in [int(m) for m in args.ms.split(",")]:
    A = rn(M, C).bfloat16()
    Wqkv = (rn(3 * C, C) * 0.02).bfloat16(); bq = rn(3 * C)
    qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    cos = rn(16, 16); sin = rn(16, 16)
    report("qkv+rope", M, 3 * C, C, lambda t: L.gemm(A, Wqkv, L.EPI_QKV_ROPE, bias=bq, out=qkv, rope_slots=256, tile_n=t,
                                                     rope_ft=16, rope_cols=2 * C, q_scale=0.125, cos_axis=cos, sin_axis=sin))
    Wp = (rn(C, C) * 0.02).bfloat16(); bp = rn(C)
    X = rn(M, C); T = torch.empty(M, C, device=dev)
    report("proj+resid", M, C, C, lambda t: L.gemm(A, Wp, L.EPI_RESID, bias=bp, out=T, resid=X, tile_n=t))
    W12 = (rn(2 * Hp, C) * 0.02).bfloat16(); b12 = rn(2 * Hp)
    hid = torch.empty(M, Hp, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(M, 2, device=dev, dtype=torch.int64)
    report("w12 swiglu+stats", M, 2 * Hp, C, lambda t: L.gemm(A, W12, L.EPI_SWIGLU, bias=b12, out=hid, row_stats=stats, tile_n=t))
    W3 = (rn(C, Hp) * 0.02).bfloat16(); u3 = rn(C)
    report("w3 ln-fold+resid", M, C, Hp, lambda t: L.gemm(hid, W3, L.EPI_RESID, bias=bp, out=X, resid=T, ln_stats=stats,
                                                         ln_u=u3, ln_n=Hd, ln_eps=1e-6, tile_n=t))
    ob = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    report("linear bf16 (N=1024)", M, C, C, lambda t: L.gemm(A, Wp, L.EPI_LINEAR, bias=bp, out=ob, tile_n=t))
for n in ((4096, 8192) if args.square else ()):
    A = rn(n, n).bfloat16(); B = rn(n, n).bfloat16(); o = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
    report("square linear", n, n, n, lambda t: L.gemm(A, B, L.EPI_LINEAR, out=o, tile_n=t))
    report("torch.matmul (cuBLAS)", n, n, n, lambda t: torch.matmul(A, B.t()))
clk.__exit__()
print("clocks:", clk.summary(), flush=True)
