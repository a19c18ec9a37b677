"""Per-shape timing of toc3d_window_attention (trains of launches, CUDA events).  Diagnostic only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from toc3d_b200 import lib as L  # noqa: E402
from bench import ClockSampler  # noqa: E402

L.load()
clk = ClockSampler(0).__enter__()      # same nvidia-smi sampler as bench.py
dev = "cuda"
heads, C = 16, 1024
for nW, seq in [(48, 256), (18, 400), (48, 180), (18, 281), (48, 129), (18, 201), (48, 103), (18, 161), (168, 256), (90, 400)]:
    qkv = torch.randn(nW * seq, 3 * C, device=dev).bfloat16()
    out = torch.empty(nW * seq, C, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.window_attention(qkv, out, nW, seq, heads)
    ts = []
    for _ in range(7):
        torch.cuda._sleep(2_000_000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            L.window_attention(qkv, out, nW, seq, heads)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / 10 * 1e3)
    ts.sort()
    fl = 4.0 * nW * heads * seq * seq * 64
    print("nW=%3d seq=%3d  %7.1f us  %6.1f TF/s" % (nW, seq, ts[3], fl / ts[3] / 1e6), flush=True)
clk.__exit__()
print("clocks:", clk.summary(), flush=True)
