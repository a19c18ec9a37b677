"""Probe: does a torch-captured CUDA graph keep programmatic dependent launch edges?  Times a train of 20
dependent GEMMs (M=3618, N=K=1024) as (a) stream launches, (b) graph replay; run with and without TOC3D_NO_PDL=1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from toc3d_b200 import lib as L  # noqa: E402

L.load()
dev = "cuda"
M, C = 3618, 1024
A = torch.randn(M, C, device=dev).bfloat16()
W = (torch.randn(C, C, device=dev) * 0.02).bfloat16()
bufs = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(2)]


def train():
    cur = A
    for i in range(20):
        L.gemm(cur, W, L.EPI_LINEAR, out=bufs[i % 2])
        cur = bufs[i % 2]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(2_000_000)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps / 20 * 1e3


print("PDL", "off" if os.environ.get("TOC3D_NO_PDL") else "on")
print("  stream: %.2f us per GEMM" % timeit(train))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    train()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        train()
torch.cuda.synchronize()
print("  graph : %.2f us per GEMM" % timeit(g.replay))
