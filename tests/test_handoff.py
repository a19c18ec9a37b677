"""Query hand-off row: host mirror of petr3d.py:116-143 (CPU; no kernel involved)."""
import types

import pytest
import torch

from toc3d_b200.handoff import memory_queries


def _head(B=2, M=640, dim=256, grad=True):
    g = torch.Generator().manual_seed(0)
    h = types.SimpleNamespace(embed_dims=dim)
    h.memory_embedding = torch.randn(B, M, dim, generator=g, requires_grad=grad)
    h.memory_reference_point = torch.rand(B, M, 3, generator=g)
    h.memory_timestamp = torch.rand(B, M, 1, generator=g, dtype=torch.float64)   # float64 in the reference (streampetr_head.py)
    h.memory_egopose = torch.eye(4).expand(B, M, 4, 4).clone()
    h.memory_velo = torch.randn(B, M, 2, generator=g)
    return h


def test_mid_frame_takes_the_leading_memory_entries_detached():
    h = _head()
    kw = memory_queries(h, torch.tensor([[1.0], [1.0]]), 2, 64, "cpu")
    assert kw["prev_exists"] is True
    assert torch.equal(kw["temp_queries"], h.memory_embedding[:, :64].detach()) and not kw["temp_queries"].requires_grad
    assert kw["temp_queries"].data_ptr() == h.memory_embedding.data_ptr()        # a view, not a copy (reference semantics)
    assert torch.equal(kw["temp_ref_points"], h.memory_reference_point[:, :64])
    assert torch.equal(kw["temp_timestamp"], h.memory_timestamp[:, :64]) and kw["temp_timestamp"].dtype == torch.float64
    assert torch.equal(kw["temp_ego_pose"], h.memory_egopose[:, :64]) and kw["temp_ego_pose"].shape == (2, 64, 4, 4)
    assert torch.equal(kw["temp_vel"], h.memory_velo[:, :64])
    assert memory_queries(h, True, 2, 64, "cpu")["prev_exists"] is True          # host bool: no tensor read


@pytest.mark.parametrize("prev", [False, torch.tensor([0.0, 1.0])])
def test_first_frame_or_empty_memory_gives_zeros(prev):
    kw = memory_queries(_head(), prev, 2, 64, "cpu")
    assert kw["prev_exists"] is False
    shapes = {k: tuple(v.shape) for k, v in kw.items() if torch.is_tensor(v)}
    assert shapes == {"temp_queries": (2, 64, 256), "temp_ref_points": (2, 64, 3), "temp_timestamp": (2, 64, 1),
                      "temp_ego_pose": (2, 64, 4, 4), "temp_vel": (2, 64, 2)}
    assert all(v.dtype == torch.float32 and not v.any() for v in kw.values() if torch.is_tensor(v))
    empty = _head(); empty.memory_embedding = None
    kw = memory_queries(empty, True, 1, 64, "cpu")
    assert kw["prev_exists"] is True and kw["temp_queries"].shape == (1, 64, 256) and not kw["temp_queries"].any()
    with pytest.raises(AssertionError):
        memory_queries(_head(), None, 2, 64, "cpu")


REF_PETR3D = "/root/reference/projects/mmdet3d_plugin/models/detectors/petr3d.py"


@pytest.mark.skipif(not __import__("os").path.exists(REF_PETR3D), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("prev,empty", [(1.0, False), (0.0, False), (1.0, True)])
def test_mirror_equals_the_reference_lines_executed(prev, empty):
    """The reference's OWN statements (the `if self.query_backbone_selection:` block of Petr3D.extract_img_feat,
    petr3d.py:116-143 - the detector class itself cannot be imported without mmdet3d) are cut out of the source file and
    executed against a stand-in `self`; the mirror must hand the backbone the very same tensors."""
    import textwrap
    src = open(REF_PETR3D).read().split("\n")
    a = next(i for i, l in enumerate(src) if "if self.query_backbone_selection:" in l)
    b = next(i for i in range(a, len(src)) if src[i].strip() == "mid_frame = False")
    block = textwrap.dedent("\n".join(src[a:b + 1]))
    head = _head()
    if empty:
        head.memory_embedding = None
    me = types.SimpleNamespace(query_backbone_selection=True, pts_bbox_head=head,
                               img_backbone=types.SimpleNamespace(pruning_num_queries=64))
    B = 2
    ns = {"self": me, "torch": torch, "prev_exists": torch.full((B, 1), prev), "B": B, "img": torch.zeros(1)}
    exec(compile(block, REF_PETR3D, "exec"), ns)
    kw = memory_queries(head, ns["prev_exists"], B, 64, "cpu")
    assert kw["prev_exists"] == ns["mid_frame"]
    for ours, theirs in (("temp_queries", "mem_queries"), ("temp_ref_points", "mem_reference_point"), ("temp_timestamp", "mem_timestamp"),
                         ("temp_ego_pose", "mem_egopose"), ("temp_vel", "mem_velo")):
        x, y = kw[ours], ns[theirs]
        assert x.shape == y.shape and x.dtype == y.dtype and torch.equal(x, y) and x.requires_grad == y.requires_grad, ours
        if prev and not empty:
            assert x.data_ptr() == y.data_ptr()            # both are views of the head's memory


@pytest.mark.gpu
def test_backbone_consumes_the_handoff():
    from tests.helpers import build_model
    from toc3d_b200 import TINY
    from toc3d_b200.synthetic import randomize_state_dict
    model = build_model("toc3d", TINY)
    model.load_state_dict(randomize_state_dict(model.state_dict(), seed=2, bias_std=0.05))
    model = model.cuda()
    h = _head(B=1, grad=False)
    for a in ("memory_embedding", "memory_reference_point", "memory_timestamp", "memory_egopose", "memory_velo"):
        setattr(h, a, getattr(h, a).cuda())
    x = torch.randn(2, 3, 160, 352, device="cuda")
    eye = torch.eye(4, device="cuda")[None]
    with torch.no_grad():
        for prev in (False, True, True):
            kw = memory_queries(h, prev, 1, model.pruning_num_queries, "cuda")
            out = model(x=x, ego_pose_inv=eye, **kw)
            assert torch.isfinite(out.img_feats["last_feat"]).all() and out.keep_idx[0].shape[0] == 2
