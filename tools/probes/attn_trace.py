"""Timeline of one CTA of the ping-pong attention kernel (diagnostic).  Builds a -DTOC3D_ATTN_TRACE copy of the library
(tools/probes/libtoc3d_trace.so; never the product library), runs one launch and prints the clock64 stamps.

    python tools/probes/attn_trace.py build        # here (nvcc)
    python tools/probes/attn_trace.py [nW seq]     # on the GPU box
"""
import ctypes
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TRACE_LIB = os.path.join(HERE, "libtoc3d_trace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    src = sorted(glob.glob(os.path.join(ROOT, "toc3d_b200", "csrc", "*.cu")))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-DTOC3D_PRECISE_MATH", "-DTOC3D_ATTN_TRACE", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
                           "-o", TRACE_LIB] + src)
    print(TRACE_LIB)
    sys.exit(0)

import torch  # noqa: E402

sys.path.insert(0, ROOT)
from toc3d_b200 import lib as L  # noqa: E402

L.LIB_PATH = TRACE_LIB
so = L.load()
nW, seq = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (48, 180)
heads, C = 16, 1024
qkv = torch.randn(nW * seq, 3 * C, device="cuda").bfloat16()
out = torch.empty(nW * seq, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    L.window_attention(qkv, out, nW, seq, heads)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (3 * 32 * 8))()
so.toc3d_attn_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = so.toc3d_attn_trace_read(buf, 3 * 32 * 8)
assert rc == 0, rc
t = [[[buf[(r * 32 + u) * 8 + k] for k in range(8)] for u in range(32)] for r in range(3)]
t0 = min(x for r in t for u in r for x in u if x)
rel = lambda x: (x - t0) if x else -1
print("nW=%d seq=%d  (clock64 relative to the first stamp)" % (nW, seq))
print("softmax warps: unit | wait S | S ready | max done | exp done | P arrived | tile done")
for slot in range(2):
    for n in range(8):
        if t[slot][n][0]:
            print("slot %d n=%d  " % (slot, n) + "  ".join("%7d" % rel(t[slot][n][k]) for k in range(6)))
print("MMA thread: unit | top | FULL ok | QK committed | after PV(prev) | [P ready of this unit (stamp 4)]")
for u in range(16):
    if t[2][u][0]:
        print("u=%2d  " % u + "  ".join("%7d" % rel(t[2][u][k]) for k in range(5)))
