"""Where a GEMM launch spends its time (diagnostic): -DTOC3D_GEMM_TRACE build of the library (tools/probes/libtoc3d_gtrace.so,
never the product library) stamps %globaltimer in pair 0 of the last 8 launches of a train; this prints the phases.

    python tools/probes/gemm_trace.py build        # here (nvcc)
    python tools/probes/gemm_trace.py              # on the GPU box
"""
import ctypes
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libtoc3d_gtrace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    src = sorted(glob.glob(os.path.join(ROOT, "toc3d_b200", "csrc", "*.cu")))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-DTOC3D_PRECISE_MATH", "-DTOC3D_GEMM_TRACE", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", LIB] + src)
    print(LIB)
    sys.exit(0)

import torch  # noqa: E402

sys.path.insert(0, ROOT)
from toc3d_b200 import lib as L  # noqa: E402

L.LIB_PATH = os.environ.get("TOC3D_LIB", LIB)
so = L.load()
so.toc3d_gemm_trace_read.argtypes = [ctypes.c_void_p]
dev = "cuda"
names = ["entry", "prologue done", "dependency ok", "first operands", "last MMA issued", "(unused)", "last epilogue done", "before exit"]
for label, M, N, K, kind in (("proj+resid", 3744, 1024, 1024, "resid"), ("proj+resid", 6000, 1024, 1024, "resid"),
                             ("qkv-like linear", 6000, 3072, 1024, "linear"), ("w3-like resid", 6000, 1024, 2752, "resid"),
                             ("one tile per pair, 1 k-block", 3744, 1024, 64, "linear")):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    if kind == "resid":
        X = torch.randn(M, N, device=dev)
        T = torch.empty(M, N, device=dev)
        fn = lambda: L.gemm(A, W, L.EPI_RESID, out=T, resid=X)
    else:
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda: L.gemm(A, W, L.EPI_LINEAR, out=o)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(2_000_000)
    for _ in range(16):
        fn()
    buf = (ctypes.c_ulonglong * 64)()
    assert so.toc3d_gemm_trace_read(buf) == 0
    eb = (ctypes.c_ulonglong * 128)()
    so.toc3d_gemm_epi_trace_read.argtypes = [ctypes.c_void_p]
    assert so.toc3d_gemm_epi_trace_read(eb) == 0
    epi = {buf[i * 8]: [eb[i * 16 + k] for k in range(16)] for i in range(8)}
    rows = sorted([[buf[i * 8 + k] for k in range(8)] for i in range(8)], key=lambda r: r[0])
    print("%s  M=%d N=%d K=%d   (ns; consecutive launches of a train, pair 0)" % (label, M, N, K))
    print("   launch period | " + " | ".join(names[1:]) + "   (each relative to this launch's entry)")
    for a, b in zip(rows[2:], rows[3:]):
        print("   %6d        | " % (b[0] - a[0]) + " | ".join("%6d" % (a[k] - a[0]) for k in range(1, 8))
              + "   | next entry - this exit: %d" % (b[0] - a[7]))
        e = epi[a[0]]
        print("        last epilogue of warp 2: entry %d | prefetch issued %d | accumulator seen %d | chunks done %s" % (
            e[0] - a[0], e[1] - a[0], e[2] - a[0], " ".join(str(x - a[0]) for x in e[3:] if x > e[2])))
