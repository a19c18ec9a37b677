"""Host logic of the plugin without a GPU (tests/dryrun.py): the launch sequence of the default path and of the tested
options (view_groups)."""
import torch

from tests import dryrun
from tests.helpers import build_model
from toc3d_b200 import TINY
from toc3d_b200 import lib as L
from toc3d_b200.synthetic import make_inputs


def _inputs(hw=(160, 352), views=2):
    inp = make_inputs(1, views, hw, seed=3, pose="random")
    inp["prev_exists"] = True
    return inp


def _toc3d():
    return build_model("toc3d", TINY)


def test_default_sequence_per_block():
    m = _toc3d()
    seq = dryrun.names(dryrun.run(m, _inputs()))
    depth = len(m.blocks)
    n_acc = sum(b.accelerate for b in m.blocks)
    # every block: q/k/v, attention, proj, norm2, w1/w2, w3; dense blocks start with layernorm_rows, accelerated
    # ones with the fused gather + merge + norm1 launch and end with the fast-token update
    assert seq.count("gemm:%d" % L.EPI_QKV_ROPE) == depth and seq.count("window_attention") == depth
    assert seq.count("gemm:%d" % L.EPI_SWIGLU) == depth
    assert seq.count("gemm:%d" % L.EPI_RESID) == 2 * depth + 1                     # proj + w3 per block, patch embed
    # fast-token updates are deferred into the next block's first launch inside a stage: one launch per stage remains
    assert seq.count("ln_gather_merge") == n_acc and seq.count("fast_token_update") == len(m.pruning_loc)
    m2 = _toc3d()
    seq2 = dryrun.names(dryrun.run(m2, _inputs(), defer_fast_update=False))
    assert seq2.count("fast_token_update") == n_acc and len(seq2) == len(seq) + n_acc - len(m.pruning_loc)
    assert seq.count("layernorm_rows") == depth + (depth - n_acc)                 # norm2 everywhere + norm1 of dense blocks
    assert seq.count("motion_queries_fold") == 1                                   # all three stages in one call
    i = seq.index("window_attention")
    assert seq[i + 1:i + 5] == ["gemm:%d" % L.EPI_RESID, "layernorm_rows", "gemm:%d" % L.EPI_SWIGLU, "gemm:%d" % L.EPI_RESID]


def test_dense_model_sequences():
    m = build_model("dense", dict(TINY_DENSE))
    inp = {"x": _inputs()["x"]}
    seq = dryrun.names(dryrun.run(m, inp))
    depth = len(m.blocks)
    assert seq.count("window_attention") == depth and seq.count("layernorm_rows") == 2 * depth


TINY_DENSE = {k: v for k, v in TINY.items() if k not in ("pc_range", "pruning_num_queries", "pruning_loc", "accelerate_global",
                                                         "token_ratio", "token_selection_loss", "rope_acc")}


def test_view_groups_launch_every_group():
    m = _toc3d()
    base = dryrun.names(dryrun.run(m, _inputs()))
    seq = dryrun.names(dryrun.run(m, _inputs(), view_groups=2))
    assert seq.count("window_attention") == 2 * base.count("window_attention")
    assert seq.count("motion_queries_fold") == 1
