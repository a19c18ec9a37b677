"""The built library really is Blackwell-native: SASS of the hot kernels (cuobjdump, no GPU needed) holds tcgen05 MMAs
(UTCHMMA), TMA tensor loads (UTMALDG), TMEM loads / stores (LDTM / STTM), no legacy HMMA in the tensor-core kernels, and
no per-instruction election loops (BRA.U.ANY) around the single-lane roles."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass(tmp_path_factory):
    from sass_diff import functions
    from toc3d_b200 import lib as L
    if not os.path.exists(CUOBJDUMP) or not os.path.exists(L.LIB_PATH):
        pytest.skip("cuobjdump or the built library is missing")
    out = tmp_path_factory.mktemp("sass") / "lib.sass"
    with open(out, "w") as f:
        subprocess.run([CUOBJDUMP, "-sass", L.LIB_PATH], stdout=f, check=True)
    return functions(str(out))


def _count(body, mnemonic):
    return sum(mnemonic in line for line in body)


def test_only_sm_100a_code_is_embedded():
    from toc3d_b200 import lib as L
    if not os.path.exists(CUOBJDUMP) or not os.path.exists(L.LIB_PATH):
        pytest.skip("cuobjdump or the built library is missing")
    out = subprocess.run([CUOBJDUMP, "-lelf", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if "ELF file" in l}
    assert archs == {"sm_100a"}, archs


def test_gemm_kernels_use_tcgen05_and_tma(sass):
    gemms = {k: v for k, v in sass.items() if "gemm_kernel" in k}
    assert len(gemms) >= 5          # LINEAR, QKV_ROPE, RESID, RESID + sub-LN fold, SWIGLU
    for name, body in gemms.items():
        assert _count(body, "UTCHMMA") >= 4, name          # tcgen05.mma, four per k-block
        assert _count(body, "UTMALDG") >= 3, name          # TMA loads of A and B
        assert _count(body, "LDTM") >= 2, name             # tcgen05.ld in the epilogue
        assert _count(body, "UTCBAR") >= 2, name           # tcgen05.commit
        assert _count(body, "HMMA") == _count(body, "UTCHMMA"), name       # no legacy mma.sync
        assert _count(body, "BRA.U.ANY") == 0, name        # single-lane roles branch on elect.sync


def test_attention_kernels_use_tcgen05_with_tmem_operands(sass):
    tc = {k: v for k, v in sass.items() if "attn_tc" in k}
    assert len(tc) >= 2             # persistent ping-pong kernel (<= 256 keys), single-slot split-softmax kernel (<= 448)
    for name, body in tc.items():
        assert _count(body, "UTCHMMA") >= 8, name
        assert _count(body, "UTMALDG") >= 3 and _count(body, "LDTM") >= 2 and _count(body, "STTM") >= 1, name
        assert _count(body, "MUFU.EX2") >= 32 and _count(body, "BRA.U.ANY") == 0, name
