"""Host logic of the plugin without a GPU (tests/dryrun.py): launch sequence of the default path and of the opt-in
chained launches, and the consistency of what a chained launch is given (problems, schedule, counters)."""
import torch

from tests import dryrun
from tests.helpers import build_model
from toc3d_b200 import TINY, chain_plan
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
    assert seq.count("ln_gather_merge") == n_acc and seq.count("fast_token_update") == n_acc
    assert seq.count("layernorm_rows") == depth + (depth - n_acc)                 # norm2 everywhere + norm1 of dense blocks
    assert "gemm_chain:2" not in seq and "gemm_chain:3" not in seq
    i = seq.index("window_attention")
    assert seq[i + 1:i + 5] == ["gemm:%d" % L.EPI_RESID, "layernorm_rows", "gemm:%d" % L.EPI_SWIGLU, "gemm:%d" % L.EPI_RESID]


def _check_chain_call(call, nprob):
    (probs, M, sched, sync), _ = call[3]
    assert len(probs) == nprob
    shapes = []
    for q, (A, B, kind, epi) in enumerate(probs):
        N, K = B.shape
        shapes.append((N, K, epi.get("tile_n") or 256))
        assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.shape[0] >= M
        if q:                                      # A and the folded statistics come from the previous problem
            pA, pB, pkind, pe = probs[q - 1]
            src = pe["out"] if pkind == L.EPI_SWIGLU else pe["a_out"]
            assert A.data_ptr() == src.data_ptr() and K == (pB.shape[0] // 2 if pkind == L.EPI_SWIGLU else pB.shape[0])
            assert epi["ln_stats"].data_ptr() == pe["row_stats"].data_ptr()
    kinds = [p[2] for p in probs]
    assert kinds == ([L.EPI_SWIGLU, L.EPI_RESID] if nprob == 2 else [L.EPI_RESID, L.EPI_SWIGLU, L.EPI_RESID])
    # the uploaded schedule covers exactly the tiles of these shapes and cannot deadlock
    sh = chain_plan.chain_shape(M, shapes)
    lists = [[g for g in row.tolist() if g >= 0] for row in sched]
    chain_plan.verify(sh, lists)
    assert all(row[-1] == -1 for row in sched.tolist()) and sched.shape[0] <= 74 and sched.shape[1] <= 256
    assert sync.dtype == torch.int32 and sync.numel() >= 2 * (nprob - 1) * sh.num_m and int(sync.abs().sum()) == 0


def test_fuse_mlp_replaces_the_two_mlp_launches():
    m = _toc3d()
    base = dryrun.names(dryrun.run(m, _inputs()))
    calls = dryrun.run(m, _inputs(), fuse_mlp=True)
    seq = dryrun.names(calls)
    depth = len(m.blocks)
    assert seq.count("gemm_chain:2") == depth and seq.count("gemm:%d" % L.EPI_SWIGLU) == 0
    assert seq.count("gemm:%d" % L.EPI_RESID) == depth + 1 and seq.count("layernorm_rows") == base.count("layernorm_rows")
    # same sequence as the default with [w1/w2, w3] collapsed into the chain
    collapsed, i = [], 0
    while i < len(base):
        if base[i] == "gemm:%d" % L.EPI_SWIGLU:
            assert base[i + 1] == "gemm:%d" % L.EPI_RESID
            collapsed.append("gemm_chain:2")
            i += 2
        else:
            collapsed.append(base[i])
            i += 1
    assert seq == collapsed
    for c in calls:
        if c[0] == "gemm_chain":
            _check_chain_call(c, 2)


def test_fuse_block_tail_chains_proj_and_mlp_without_norm2_launch():
    m = _toc3d()
    calls = dryrun.run(m, _inputs(), fuse_block_tail=True)
    seq = dryrun.names(calls)
    depth = len(m.blocks)
    n_acc = sum(b.accelerate for b in m.blocks)
    assert seq.count("gemm_chain:3") == depth and seq.count("gemm:%d" % L.EPI_SWIGLU) == 0
    assert seq.count("gemm:%d" % L.EPI_RESID) == 1                                   # patch embed only
    assert seq.count("layernorm_rows") == depth - n_acc                              # norm1 of the dense blocks only
    i = seq.index("window_attention")
    assert seq[i + 1] == "gemm_chain:3"
    for c in calls:
        if c[0] == "gemm_chain":
            _check_chain_call(c, 3)
            probs = c[3][0][0]
            assert probs[0][3].get("out_map") is None and probs[0][3]["zero_stats"].data_ptr() == probs[1][3]["row_stats"].data_ptr()


def test_dense_model_sequences():
    m = build_model("dense", dict(TINY_DENSE))
    inp = {"x": _inputs()["x"]}
    seq = dryrun.names(dryrun.run(m, inp))
    depth = len(m.blocks)
    assert seq.count("window_attention") == depth and seq.count("layernorm_rows") == 2 * depth
    seq3 = dryrun.names(dryrun.run(m, inp, fuse_block_tail=True))
    assert seq3.count("gemm_chain:3") == depth and seq3.count("layernorm_rows") == depth


TINY_DENSE = {k: v for k, v in TINY.items() if k not in ("pc_range", "pruning_num_queries", "pruning_loc", "accelerate_global",
                                                         "token_ratio", "token_selection_loss", "rope_acc")}


def test_chained_launches_refuse_concurrent_view_groups():
    """A chained launch is deadlock-free only when its whole grid is co-resident: two of them on concurrent streams
    (view_groups > 1) are refused before anything is launched."""
    import pytest
    m = _toc3d()
    with pytest.raises(NotImplementedError, match="co-resident"):
        dryrun.run(m, _inputs(), fuse_mlp=True, view_groups=2)
    m = _toc3d()
    assert "gemm_chain:2" not in dryrun.names(dryrun.run(m, _inputs(), view_groups=2))      # default path still runs
