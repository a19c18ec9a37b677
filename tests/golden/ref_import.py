"""Import the UNMODIFIED reference backbone from /root/reference on CPU.

Test infrastructure only (SURVEY.md §8c recipe).  The reference tree needs
mmcv/mmdet/mmdet3d/detectron2/fvcore/timm/fairscale, none of which exist in
this image, so the heavy package ``__init__`` files are bypassed with shell
packages and the handful of imported third-party names are stubbed.  None of
the stubs carries arithmetic that is on the backbone path (DropPath and
checkpoint_wrapper are identities in eval).

Two pins are applied, each a choice among behaviours the reference leaves
unspecified:
  pin 1  torch.sort inside toc3d_utils runs with stable=True
         (order = score descending, index ascending);
  pin 2  F.gumbel_softmax inside toc3d_utils consumes injected noise for the
         image-level calls (last dim 2) and returns ones for the window-level
         calls (last dim 1, result discarded by the caller).

This file is never imported by the product.  It exists to generate
tests/golden/*.pt, to validate oracle/ against the real reference in this
container, and - through the staged copy baseline/_ref/ (tools/stage_ref.sh,
git-ignored, travels with the gpurun snapshot) - to let bench.py time the
reference's own modules on the GPU box, where /root/reference is absent.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _find_root():
    """TOC3D_REFERENCE_ROOT, else /root/reference (build container), else the git-ignored staged copy
    baseline/_ref/ (tools/stage_ref.sh; travels to the GPU box with the gpurun snapshot)."""
    cands = [os.environ.get("TOC3D_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "projects", "mmdet3d_plugin", "models", "backbones")):
            return c
    return cands[1]


REF_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "projects", "mmdet3d_plugin", "models", "backbones"))


class _Registry:
    def register_module(self, *a, **k):
        def deco(cls):
            return cls
        return deco


def _shell(name, path=None):
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


_loaded = {}


def load_reference():
    """Returns a namespace with ToC3DEVAViT, EVA_ViT, toc3d_utils, pins."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    plug = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin")
    _shell("projects", os.path.join(REF_ROOT, "projects"))
    _shell("projects.mmdet3d_plugin", plug)
    _shell("projects.mmdet3d_plugin.models", os.path.join(plug, "models"))
    _shell("projects.mmdet3d_plugin.models.backbones", os.path.join(plug, "models", "backbones"))
    _shell("projects.mmdet3d_plugin.models.utils", os.path.join(plug, "models", "utils"))

    # third-party stubs ------------------------------------------------------
    _shell("mmdet", "/nonexistent"); _shell("mmdet.models", "/nonexistent")
    b = _shell("mmdet.models.builder"); b.BACKBONES = _Registry()
    _shell("mmdet.models.utils", "/nonexistent")
    t = _shell("mmdet.models.utils.transformer"); t.inverse_sigmoid = lambda x, eps=1e-5: x
    c = _shell("mmdet.core"); c.bbox_xyxy_to_cxcywh = lambda x: x
    _shell("mmdet3d", "/nonexistent"); _shell("mmdet3d.models", "/nonexistent")
    b3 = _shell("mmdet3d.models.builder"); b3.build_loss = lambda cfg: None
    _shell("detectron2", "/nonexistent")
    dl = _shell("detectron2.layers")
    dl.CNNBlockBase = nn.Module; dl.Conv2d = nn.Conv2d
    dl.get_norm = lambda *a, **k: None
    dl.ShapeSpec = lambda **k: types.SimpleNamespace(**k)
    _shell("detectron2.modeling", "/nonexistent"); _shell("detectron2.modeling.backbone", "/nonexistent")
    f = _shell("detectron2.modeling.backbone.fpn"); f._assert_strides_are_log2_contiguous = lambda s: None
    _shell("fvcore", "/nonexistent"); fn = _shell("fvcore.nn", "/nonexistent")
    wi = _shell("fvcore.nn.weight_init"); fn.weight_init = wi
    _shell("timm", "/nonexistent"); _shell("timm.models", "/nonexistent")
    tl = _shell("timm.models.layers")

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x
    tl.DropPath = DropPath
    _shell("fairscale", "/nonexistent"); _shell("fairscale.nn", "/nonexistent")
    fc = _shell("fairscale.nn.checkpoint"); fc.checkpoint_wrapper = lambda m, *a, **k: m

    bb = "projects.mmdet3d_plugin.models.backbones."
    tu = importlib.import_module(bb + "toc3d_utils")
    tv = importlib.import_module(bb + "toc3d_eva_vit")
    ev = importlib.import_module(bb + "eva_vit")

    # pin 1: stable sort ------------------------------------------------------
    class _TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def sort(x, dim=-1, descending=False, **kw):
            return torch.sort(x, dim=dim, descending=descending, stable=True)
    tu.torch = _TorchProxy()

    # pin 2: injected gumbel noise -------------------------------------------
    state = {"noise": [], "calls": 0}

    class _FProxy:
        def __getattr__(self, k):
            return getattr(torch.nn.functional, k)

        @staticmethod
        def gumbel_softmax(logits, tau=1, hard=False, dim=-1):
            if logits.shape[-1] == 1:
                return torch.ones_like(logits)
            g = state["noise"][state["calls"]]
            state["calls"] += 1
            return ((logits + g.to(logits.dtype).view_as(logits)) / tau).softmax(dim)
    tu.F = _FProxy()

    def set_gumbel(noise_list):
        state["noise"] = list(noise_list)
        state["calls"] = 0

    # neck: cp_fpn.py needs mmcv's ConvModule / BaseModule / auto_fp16 and mmdet's NECKS.
    # With the shipped config (norm_cfg=None, act_cfg=None) ConvModule is a plain nn.Conv2d with bias,
    # registered as sub-module `conv` (state-dict keys lateral_convs.0.conv.weight, ...); auto_fp16 is the
    # identity unless fp16_enabled is set (it is False).
    _shell("projects.mmdet3d_plugin.models.necks", os.path.join(plug, "models", "necks"))
    _shell("mmcv", "/nonexistent")
    mc = _shell("mmcv.cnn")

    class ConvModule(nn.Module):
        def __init__(self, cin, cout, k, stride=1, padding=0, conv_cfg=None, norm_cfg=None, act_cfg=None, inplace=False):
            super().__init__()
            assert conv_cfg is None and norm_cfg is None and act_cfg is None, "only the shipped CPFPN config is stubbed"
            self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=True)

        def forward(self, x):
            return self.conv(x)
    mc.ConvModule = ConvModule
    mr = _shell("mmcv.runner")

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
    mr.BaseModule = BaseModule
    mr.auto_fp16 = lambda *a, **k: (lambda fn: fn)
    sys.modules["mmdet.models"].NECKS = _Registry()
    nk = importlib.import_module("projects.mmdet3d_plugin.models.necks.cp_fpn")

    ns = types.SimpleNamespace(ToC3DEVAViT=tv.ToC3DEVAViT, EVA_ViT=ev.EVA_ViT, toc3d_utils=tu,
                               toc3d_eva_vit=tv, eva_vit=ev, set_gumbel=set_gumbel, CPFPN=nk.CPFPN)
    _loaded["ns"] = ns
    return ns
