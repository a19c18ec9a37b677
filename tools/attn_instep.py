"""The attention launches of one ToC3D_fast / dense step with their REAL arguments (q_rows, item_order, analytic pad
keys), timed per shape in trains of launches (CUDA events, clocks recorded), and - with `trace` - the clock64 timeline of
CTA 0 of the persistent kernel (<= 256 keys) from the -DTOC3D_ATTN_TRACE build (tools/probes/attn_trace.py build; the
stamps themselves cost 50-150 clk each in the traced warps).  Diagnostic only.

    python tools/attn_instep.py                 # per-shape timing, product library
    python tools/attn_instep.py trace 48 129    # timeline of one shape, trace library
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from toc3d_b200 import lib as L  # noqa: E402
from toc3d_b200.backbone import balanced_item_order  # noqa: E402

HEADS, C = 16, 1024


def real_tokens(V, H, W, ws):
    """real tokens per window of a V x H x W token grid cut into ws x ws windows (row-major windows per view)"""
    out = []
    for _ in range(V):
        for r0 in range(0, H, ws):
            for c0 in range(0, W, ws):
                out.append(min(ws, H - r0) * min(ws, W - c0))
    return torch.tensor(out, dtype=torch.int32)


def step_shapes(V=6, H=20, W=50):
    """(name, launches per step, nW, seq, q_rows, kv_rows) of the ToC3D_fast step"""
    shapes = []
    for ws, n_dense, n_s0, n_s12 in ((16, 4, 4, 8), (20, 2, 2, 4)):
        real = real_tokens(V, H, W, ws)
        n = ws * ws
        shapes.append(("dense ws%d" % ws, n_dense, len(real), n, real, real))
        for ratio, cnt in ((0.7, n_s0), (0.5, n_s12)):
            k = int(n * ratio)
            q = torch.minimum(real, torch.tensor(k, dtype=torch.int32)) + 1
            shapes.append(("toc3d ws%d k=%d" % (ws, k), cnt, len(real), k + 1, q, None))
    return shapes


def make_args(nW, seq, q_rows, kv_rows, dev="cuda"):
    qkv = torch.randn(nW * seq, 3 * C, device=dev).bfloat16()
    out = torch.empty(nW * seq, C, device=dev, dtype=torch.bfloat16)
    kw = dict(q_rows=q_rows.to(dev), item_order=balanced_item_order(q_rows, HEADS).to(dev))
    if kv_rows is not None:
        kw.update(kv_rows=kv_rows.to(dev), pad_v=torch.randn(C, device=dev))
    return qkv, out, kw


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        L.LIB_PATH = os.path.join(ROOT, "tools", "probes", "libtoc3d_trace.so")
    if os.environ.get("TOC3D_LIB"):                        # A/B of two builds of the library
        L.LIB_PATH = os.environ["TOC3D_LIB"]
    so = L.load()
    shapes = step_shapes()
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        want = (int(sys.argv[2]), int(sys.argv[3]))
        for name, cnt, nW, seq, q, kv in shapes:
            if (nW, seq) != want:
                continue
            qkv, out, kw = make_args(nW, seq, q, kv)
            for _ in range(3):
                L.window_attention(qkv, out, nW, seq, HEADS, **kw)
            torch.cuda.synchronize()
            buf = (ctypes.c_ulonglong * (4 * 32 * 8))()
            so.toc3d_attn_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
            assert so.toc3d_attn_trace_read(buf, 4 * 32 * 8) == 0
            cb = (ctypes.c_ulonglong * 64)()
            so.toc3d_attn_trace_chunks.argtypes = [ctypes.c_void_p]
            assert so.toc3d_attn_trace_chunks(cb) == 0
            t = [[[buf[(r * 32 + u) * 8 + k] for k in range(8)] for u in range(32)] for r in range(4)]
            t0 = min(x for r in t for u in r for x in u if x)
            rel = lambda x: (x - t0) if x else -1
            print("%s nW=%d seq=%d  (clock64 relative to the first stamp)" % (name, nW, seq))
            print("softmax warps (lane quarter 0): slot, unit of the slot | wait S | S ready | max pass done | exp pass done | P arrived")
            for slot in range(2):
                for n in range(12):
                    if t[slot][n][0]:
                        print("slot %d n=%d  " % (slot, n) + "  ".join("%7d" % rel(t[slot][n][k]) for k in range(5)))
            print("TMA producer: item | before EMPTY wait | EMPTY ok | loads issued")
            for n in range(12):
                if t[3][n][0]:
                    print("i=%2d  " % n + "  ".join("%7d" % rel(t[3][n][k]) for k in range(3)))
            print("MMA thread: unit | top | FULL / OFREE ok | QK committed | after preparing the next unit and P V (previous unit)")
            for u in range(24):
                if t[2][u][0]:
                    print("u=%2d  " % u + "  ".join("%7d" % rel(t[2][u][k]) for k in range(4)))
            fb = (ctypes.c_ulonglong * 512)()
            so.toc3d_attn_trace_fin.argtypes = [ctypes.c_void_p]
            assert so.toc3d_attn_trace_fin(fb) == 0
            print("epilogue warp (quarter 0), unit n: start | O ready | O loaded | OFREE arrived | staged | stored")
            for n in range(16):
                st = [fb[n * 8 + k] for k in range(6)]
                if st[0]:
                    print("u=%2d  " % n + "  ".join("%7d" % rel(x) for x in st))
            for slot in range(2):
                for ps in range(2):
                    st = [cb[(slot * 2 + ps) * 16 + c] for c in range(16)]
                    st = [x for x in st if x]
                    print("slot %d unit 1 pass %d chunk stamps (delta):" % (slot, ps + 1), " ".join(
                        "%d" % (b - a) for a, b in zip(st, st[1:])), " first rel", rel(st[0]) if st else -1)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "one":        # a few launches of one shape (for ncu)
        want = (int(sys.argv[2]), int(sys.argv[3]))
        for name, cnt, nW, seq, q, kv in shapes:
            if (nW, seq) == want:
                qkv, out, kw = make_args(nW, seq, q, kv)
                for _ in range(4):
                    L.window_attention(qkv, out, nW, seq, HEADS, **kw)
                torch.cuda.synchronize()
        return
    from bench import ClockSampler
    clk = ClockSampler(0).__enter__()
    total = 0.0
    for name, cnt, nW, seq, q, kv in shapes:
        qkv, out, kw = make_args(nW, seq, q, kv)
        for _ in range(3):
            L.window_attention(qkv, out, nW, seq, HEADS, **kw)
        ts = []
        for _ in range(7):
            torch.cuda._sleep(2_000_000)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                L.window_attention(qkv, out, nW, seq, HEADS, **kw)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / 10 * 1e3)
        ts.sort()
        units = int(((q + 127) // 128).sum()) * HEADS
        total += cnt * ts[3]
        print("%-18s nW=%3d seq=%3d  x%d  %7.1f us   (%d tile units)" % (name, nW, seq, cnt, ts[3], units), flush=True)
    print("sum over the step's 24 launches: %.1f us" % total)
    clk.__exit__()
    print("clocks:", clk.summary(), flush=True)


if __name__ == "__main__":
    main()
