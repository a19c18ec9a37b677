"""In-tree build of libtoc3d_b200.so (nvcc, sm_100a only).  `python -m toc3d_b200.build`."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtoc3d_b200.so")
SRC = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
DEPS = SRC + sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [
    os.path.join(os.path.dirname(HERE), "include", "toc3d_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--use_fast_math" if os.environ.get("TOC3D_FAST_MATH") else "-DTOC3D_PRECISE_MATH",
         "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SRC
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libtoc3d_b200.so")
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
