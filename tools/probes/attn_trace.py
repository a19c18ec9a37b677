"""Builds the -DTOC3D_ATTN_TRACE copy of the library (tools/probes/libtoc3d_trace.so; never the product library) whose
persistent attention kernel stamps clock64 into device arrays (CTA 0: softmax warps of lane quarter 0, MMA thread, TMA
producer, epilogue warp of quarter 0).  The timeline is printed by tools/attn_instep.py:

    python tools/probes/attn_trace.py build              # here (nvcc)
    python tools/attn_instep.py trace 48 129             # on the GPU box: one of the step's shapes (windows, keys)
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TRACE_LIB = os.path.join(HERE, "libtoc3d_trace.so")

if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] != "build":
        sys.exit(__doc__)
    src = sorted(glob.glob(os.path.join(ROOT, "toc3d_b200", "csrc", "*.cu")))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-DTOC3D_PRECISE_MATH", "-DTOC3D_ATTN_TRACE", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
                           "-o", TRACE_LIB] + src)
    print(TRACE_LIB)
