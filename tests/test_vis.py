"""Row f4: the token-selection visualiser mirror (toc3d_b200/vis.py) against the reference's own function
(token_select_vis.py:8-79) run here with mmcv stubbed through the cv2 calls it makes, and against a committed golden of
that run; plus the consumer contract on the outputs of the oracle (same shapes / dtypes as the plugin's)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from tests.golden.ref_import import reference_available
from toc3d_b200 import vis

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "token_vis.npz")
NORM = dict(mean=np.array([103.530, 116.280, 123.675], dtype=np.float32),
            std=np.array([57.375, 57.120, 58.395], dtype=np.float32), to_rgb=False)


def vis_case():
    g = torch.Generator().manual_seed(11)
    V, H, W, ps = 2, 32, 48, 16
    N = (H // ps) * (W // ps)
    imgs = torch.randn(V, 3, H, W, generator=g)
    masks = [torch.rand(V, H // ps, W // ps, 1, generator=g) for _ in range(3)]
    keeps, drops = [], []
    for k in (4, 3, 2):
        perm = torch.stack([torch.randperm(N, generator=g) for _ in range(V)])
        keeps.append(perm[:, :k].contiguous()); drops.append(perm[:, k:].contiguous())
    return imgs, masks, keeps, drops


def run_reference(imgs, masks, keeps, drops):
    """The unmodified reference function; mmcv.imdenormalize / imwrite replaced by their cv2 bodies / a recorder."""
    import cv2
    import importlib.util
    written = {}
    mm = types.ModuleType("mmcv")

    def imdenormalize(img, mean, std, to_bgr=True):
        assert img.dtype != np.uint8
        mean = mean.reshape(1, -1).astype(np.float64)
        std = std.reshape(1, -1).astype(np.float64)
        img = cv2.multiply(img, std)
        cv2.add(img, mean, img)
        if to_bgr:
            cv2.cvtColor(img, cv2.COLOR_RGB2BGR, img)
        return img
    mm.imdenormalize = imdenormalize
    mm.imwrite = lambda arr, path: written.__setitem__(os.path.basename(path), np.array(arr, copy=True))
    old = sys.modules.get("mmcv")
    sys.modules["mmcv"] = mm
    try:
        spec = importlib.util.spec_from_file_location(
            "_ref_token_select_vis", "/root/reference/projects/mmdet3d_plugin/models/utils/token_select_vis.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.token_selection_vis(imgs, masks, keeps, drops, NORM, "/tmp/unused/")
    finally:
        if old is None:
            del sys.modules["mmcv"]
        else:
            sys.modules["mmcv"] = old
    return written


def _same(a, b):
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_overlays_match_the_live_reference():
    imgs, masks, keeps, drops = vis_case()
    _same(vis.token_selection_overlays(imgs, masks, keeps, drops, NORM), run_reference(imgs, masks, keeps, drops))
    _same(vis.token_selection_overlays(imgs, masks, None, None, NORM), run_reference(imgs, masks, None, None))


def test_overlays_match_the_committed_golden():
    imgs, masks, keeps, drops = vis_case()
    got = vis.token_selection_overlays(imgs, masks, keeps, drops, NORM)
    ref = dict(np.load(GOLDEN))
    _same(got, ref)


def test_writer_and_oracle_outputs(tmp_path):
    """The consumer accepts what the backbone returns (oracle outputs have the plugin's shapes and dtypes)."""
    from tests.helpers import case_setup, run_oracle
    fx, kind, cfg, model, sd, inp, gn = case_setup("tiny_prev_small")
    out = run_oracle(kind, cfg, sd, inp, gn)
    masks = [m.reshape(m.shape[0], 10, 22, 1) for m in out["token_masks"]]
    ov = vis.token_selection_overlays(inp["x"], masks, out["keep_idx"], out["drop_idx"], NORM)
    assert len(ov) == inp["x"].shape[0] * (2 * 3 + 3)
    assert ov["view0_layer0.png"].shape == (160, 352, 4) and ov["view1_layer2_keepidx.png"].shape == (160, 352, 4)
    a = ov["view0_layer1_keepidx.png"][..., 3]
    assert set(np.unique(a).round(2)) <= {76.5, 255.0}                   # min_alpha * 255 and max_alpha * 255
    assert int((a[::16, ::16] == 255).sum()) == out["keep_idx"][1].shape[1]
    vis.token_selection_vis(inp["x"], masks, out["keep_idx"], out["drop_idx"], NORM, str(tmp_path / "v"))
    assert len(os.listdir(tmp_path / "v")) == len(ov)
    with pytest.raises(AssertionError):
        vis.token_selection_overlays(inp["x"], [m[:, :5] for m in masks], None, None, NORM)
