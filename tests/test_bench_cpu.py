"""bench.py's host logic (no GPU): the algorithmic FLOP / byte model equals SURVEY.md 8d, the roofline denominators are
read from MEASURED_PEAKS.json, both arms print the same `config` object, the reference arm prefers the reference's own
modules and only rank 0 runs it."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from toc3d_b200.configs import CONFIGS  # noqa: E402


@pytest.mark.parametrize("name,tf", [("toc3d_fast", 4.041), ("toc3d_faster", 3.200), ("eva_vit_l", 4.858),
                                     ("toc3d_fast_1600", 16.260), ("toc3d_faster_1600", 13.025), ("eva_vit_l_1600", 21.506)])
def test_algorithmic_flops_match_the_survey(name, tf):
    kind, cfg, hw = CONFIGS[name]
    w = bench.algorithmic_work(cfg, kind, hw, 6)
    assert abs(w["flops"] / 1e12 - tf) < 0.002 * tf + 5e-4, (name, w["flops"] / 1e12)


@pytest.mark.parametrize("name,gb", [("toc3d_fast", 1.226), ("toc3d_faster", 1.095), ("toc3d_fast_1600", 4.708), ("toc3d_faster_1600", 4.205)])
def test_prune_gather_bytes_match_the_survey(name, gb):
    kind, cfg, hw = CONFIGS[name]
    assert abs(bench.algorithmic_work(cfg, kind, hw, 6)["gather_bytes"] / 1e9 - gb) < 0.002 * gb + 5e-4


def test_peaks_come_from_measured_file(tmp_path, monkeypatch):
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    pk = bench.load_peaks()
    if os.path.exists(p):
        d = json.load(open(p))
        assert pk["src"] == "measured" and pk["hbm"] == d["hbm_gbs"] and pk["tf_sust"] == d["bf16_tflops_sustained"]
        assert pk["tf_burst"] == d["bf16_tflops"]
    else:
        assert pk["src"].startswith("fallback") and pk["hbm"] == 6650.0


def test_both_arms_share_the_config_object():
    a = bench.workload_config("toc3d_fast", 1, 1)
    assert a["workload"] == "toc3d_fast" and a["views"] == 6 and a["image_hw"] == [320, 800] and "neck" in a
    assert bench.workload_config("toc3d_fast", 1, 8)["parallelism"].startswith("dp8")
    assert bench.reference_views("toc3d_fast") == 6 and bench.reference_views("toc3d_faster_1600") == 1


def test_reference_arm_runs_on_rank_0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
