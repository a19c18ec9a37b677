"""Generate tests/golden/*.pt from the UNMODIFIED reference (build container only).

    python -m tests.golden.make_golden

Each fixture holds the outputs of the real reference backbone (imported via
tests/golden/ref_import.py with the two documented pins) on seeded synthetic
inputs and weights that are a pure function of (seed, key) — see
toc3d_b200/synthetic.py — so the oracle and the CUDA path can rebuild the
identical inputs on the GPU box where /root/reference does not exist.
"""
import contextlib
import io
import os

import torch

from tests.golden.ref_import import load_reference
from toc3d_b200.configs import CONFIGS, TINY
from toc3d_b200.synthetic import make_gumbel, make_inputs, randomize_state_dict

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (kind, cfg, hw, views, bias_std, prev_exists, seed, subsample)
CASES = {
    "tiny_prev": ("toc3d", TINY, (320, 800), 1, 0.1, True, 0, 1),
    "tiny_first": ("toc3d", TINY, (160, 352), 1, 0.1, False, 1, 1),
    "tiny_prev_small": ("toc3d", TINY, (160, 352), 2, 0.0, True, 2, 1),
    "tiny_dense": ("dense", {k: v for k, v in TINY.items() if k in CONFIGS["eva_vit_l"][1]},
                   (160, 352), 1, 0.1, True, 3, 1),
    "vitl_faster_1view": ("toc3d", CONFIGS["toc3d_faster"][1], (320, 800), 1, 0.1, True, 4, 16),
    # the 1600x800 token grid (50 x 100: 28 ragged ws16 windows and 15 ws20 windows per view, N = 5000)
    "tiny_prev_1600": ("toc3d", dict(TINY, token_ratio=[0.5, 0.4, 0.3]), (800, 1600), 1, 0.1, True, 5, 8),
}


def build_case(name):
    kind, cfg, hw, views, bias_std, prev, seed, sub = CASES[name]
    ns = load_reference()
    cls = ns.ToC3DEVAViT if kind == "toc3d" else ns.EVA_ViT
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(**cfg).eval()
    sd = randomize_state_dict(m.state_dict(), seed=seed, bias_std=bias_std)
    m.load_state_dict(sd)
    inp = make_inputs(1, views, hw, seed=seed, pose="random")
    inp["prev_exists"] = prev
    N = (hw[0] // 16) * (hw[1] // 16)
    gn = make_gumbel(views, N, seed=seed + 100)
    ns.set_gumbel(gn)
    with torch.no_grad():
        r = m(**inp)
    fx = dict(meta=dict(kind=kind, hw=list(hw), views=views, bias_std=bias_std, prev_exists=prev, seed=seed,
                        subsample=sub, n_keys=len(sd), torch=str(torch.__version__)))
    if kind == "dense":
        fx["last_feat"] = r["last_feat"][:, ::sub].contiguous()
    else:
        fx["last_feat"] = r.img_feats["last_feat"][:, ::sub].contiguous()
        fx["token_masks"] = [t.contiguous() for t in r.token_masks]
        fx["keep_idx"] = [t.to(torch.int16) for t in r.keep_idx]
        fx["drop_idx"] = [t.to(torch.int16) for t in r.drop_idx]
        assert r.attn_scores is None
    return fx


NECK_CASE = dict(views=2, hw=(10, 14), in_ch=1024, out_ch=256, seed=7)


def neck_inputs():
    """Seeded input map and weights of the neck golden (rebuilt identically by the tests)."""
    c = NECK_CASE
    g = torch.Generator(); g.manual_seed(c["seed"])
    x = torch.randn(c["views"], c["in_ch"], c["hw"][0], c["hw"][1], generator=g) * 3.0
    tmpl = {"lateral_convs.0.conv.weight": torch.empty(c["out_ch"], c["in_ch"], 1, 1),
            "lateral_convs.0.conv.bias": torch.empty(c["out_ch"]),
            "fpn_convs.0.conv.weight": torch.empty(c["out_ch"], c["out_ch"], 3, 3),
            "fpn_convs.0.conv.bias": torch.empty(c["out_ch"])}
    return x, randomize_state_dict(tmpl, seed=c["seed"], bias_std=0.1)


def build_neck_case():
    """Outputs of the real reference CPFPN (necks/cp_fpn.py) with the shipped config."""
    ns = load_reference()
    c = NECK_CASE
    m = ns.CPFPN(in_channels=[c["in_ch"]], out_channels=c["out_ch"], num_outs=2).eval()
    x, sd = neck_inputs()
    m.load_state_dict(sd)
    with torch.no_grad():
        outs = m([x])
    return dict(meta=dict(NECK_CASE, torch=str(torch.__version__), keys=sorted(m.state_dict().keys())),
                outs=[o.contiguous() for o in outs])


PREPROCESS_CASES = [  # (views, Hs, Ws, to_rgb)
    (2, 64, 96, False), (2, 64, 96, True), (1, 50, 70, False), (1, 33, 40, True)]
IMG_NORM = dict(mean=[103.530, 116.280, 123.675], std=[57.375, 57.120, 58.395])      # ToC3D_fast.py:13-14


def preprocess_input(case):
    """Seeded uint8 HWC crops (every byte value occurs); rebuilt identically by the tests."""
    V, Hs, Ws, _ = case
    g = torch.Generator(); g.manual_seed(V * 1000 + Hs * 10 + Ws)
    x = torch.randint(0, 256, (V, Hs, Ws, 3), generator=g, dtype=torch.uint8)
    x.view(-1)[:256] = torch.arange(256, dtype=torch.uint8)
    return x


def build_preprocess_case():
    """NormalizeMultiviewImage + PadMultiViewImage through the very OpenCV calls mmcv 1.6.0 makes
    (imnormalize_: cvtColor / cv2.subtract / cv2.multiply with float64 scalars; impad: cv2.copyMakeBorder).
    mmcv itself is not installed; cv2 is."""
    import cv2
    import numpy as np
    mean = np.array(IMG_NORM["mean"], dtype=np.float32)
    std = np.array(IMG_NORM["std"], dtype=np.float32)
    outs = []
    for case in PREPROCESS_CASES:
        x = preprocess_input(case)
        per_view = []
        for v in range(case[0]):
            img = x[v].numpy().astype(np.float32)                  # to_float32=True
            img = img.copy().astype(np.float32)                    # imnormalize
            mean64 = np.float64(mean.reshape(1, -1))
            stdinv = 1 / np.float64(std.reshape(1, -1))
            if case[3]:
                cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
            cv2.subtract(img, mean64, img)
            cv2.multiply(img, stdinv, img)
            Hi, Wi = -(-img.shape[0] // 32) * 32, -(-img.shape[1] // 32) * 32   # impad_to_multiple(32)
            img = cv2.copyMakeBorder(img, 0, Hi - img.shape[0], 0, Wi - img.shape[1], cv2.BORDER_CONSTANT, value=0)
            per_view.append(torch.from_numpy(img).permute(2, 0, 1))
        outs.append(torch.stack(per_view).contiguous())
    return dict(meta=dict(cases=PREPROCESS_CASES, cv2=cv2.__version__, **IMG_NORM), outs=outs)


def build_vis_case():
    """Row f4: the arrays the reference's token_selection_vis (models/utils/token_select_vis.py:8-79) hands to
    mmcv.imwrite for the seeded case of tests/test_vis.py (mmcv stubbed through the cv2 calls it makes)."""
    from tests.test_vis import run_reference, vis_case
    return run_reference(*vis_case())


def main():
    import numpy as np
    np.savez_compressed(os.path.join(HERE, "token_vis.npz"), **build_vis_case())
    print("token_vis written")
    fx = build_preprocess_case()
    torch.save(fx, os.path.join(HERE, "preprocess_cv2.pt"))
    print("preprocess_cv2", [tuple(o.shape) for o in fx["outs"]])
    fx = build_neck_case()
    torch.save(fx, os.path.join(HERE, "neck_cpfpn.pt"))
    print("neck_cpfpn", [tuple(o.shape) for o in fx["outs"]])
    for name in CASES:
        fx = build_case(name)
        path = os.path.join(HERE, name + ".pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB", tuple(fx["last_feat"].shape))


if __name__ == "__main__":
    main()
