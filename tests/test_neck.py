"""CPFPN neck: oracle vs the committed reference output (CPU), CUDA plugin vs oracle (GPU)."""
import os

import pytest
import torch

from oracle import toc3d_oracle as O
from tests.golden.make_golden import NECK_CASE, neck_inputs
from tests.golden.ref_import import reference_available
from toc3d_b200 import CPFPN

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "neck_cpfpn.pt")


def test_oracle_neck_matches_reference_golden():
    fx = torch.load(GOLDEN)
    x, sd = neck_inputs()
    outs = O.neck_cpfpn(sd, x)
    assert len(outs) == 2 and outs[1].shape[-2:] == (5, 7)
    for a, b in zip(outs, fx["outs"]):
        assert a.shape == b.shape and (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item())
    assert torch.equal(outs[1], outs[0][:, :, ::2, ::2])                  # max_pool2d(k=1, s=2) is a subsample


def test_neck_state_dict_keys_and_config_guard():
    fx = torch.load(GOLDEN)
    m = CPFPN(in_channels=[NECK_CASE["in_ch"]], out_channels=NECK_CASE["out_ch"], num_outs=2)
    assert sorted(m.state_dict().keys()) == fx["meta"]["keys"]
    m.load_state_dict(neck_inputs()[1])                                    # strict
    with pytest.raises(NotImplementedError):
        CPFPN(in_channels=[256, 512], out_channels=256, num_outs=2)
    with pytest.raises(NotImplementedError):
        CPFPN(in_channels=[1024], out_channels=256, num_outs=2, add_extra_convs="on_output")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.eval()([torch.zeros(1, 1024, 2, 2)])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_neck_matches_live_reference():
    from tests.golden.ref_import import load_reference
    ns = load_reference()
    m = ns.CPFPN(in_channels=[1024], out_channels=256, num_outs=2).eval()
    x = torch.randn(1, 1024, 7, 9)
    with torch.no_grad():
        r = m([x])
    o = O.neck_cpfpn(m.state_dict(), x)
    assert all((a - b).abs().max().item() < 1e-5 for a, b in zip(r, o))


@pytest.mark.gpu
@pytest.mark.parametrize("V,H,W", [(2, 10, 14), (6, 20, 50), (1, 50, 100)])
def test_neck_cuda_matches_oracle(V, H, W):
    g = torch.Generator().manual_seed(V * H)
    x = torch.randn(V, H, W, 1024, generator=g) * 3.0                      # NHWC storage, like the backbone output
    _, sd = neck_inputs()
    ref = O.neck_cpfpn(sd, x.permute(0, 3, 1, 2))
    m = CPFPN(in_channels=[1024], out_channels=256, num_outs=2).eval()
    m.load_state_dict(sd)
    m = m.cuda()
    outs = m([x.cuda().permute(0, 3, 1, 2)])
    torch.cuda.synchronize()
    assert len(outs) == 2
    for got, want in zip(outs, ref):
        assert got.shape == want.shape and got.dtype == torch.float32
        err = (got.cpu() - want).abs().max().item()
        # two chained bf16-operand GEMMs (K = 1024 and 2304), fp32 accumulate
        assert err < 1.5e-2 * max(1.0, want.abs().max().item()), err


@pytest.mark.gpu
def test_backbone_then_neck_contract():
    """ToC3DEVAViT -> CPFPN the way Petr3D chains them (petr3d.py:159-190): list(img_feats.values()) -> neck."""
    from toc3d_b200 import TINY, ToC3DEVAViT
    from toc3d_b200.synthetic import make_inputs
    torch.manual_seed(0)
    bb = ToC3DEVAViT(**TINY).eval().cuda()
    neck = CPFPN(in_channels=[TINY["embed_dim"]], out_channels=64, num_outs=2).eval().cuda()
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in make_inputs(1, 2, (160, 352), seed=3).items()}
    out = bb(**inp)
    feats = neck(list(out.img_feats.values()))
    assert feats[0].shape == (2, 64, 10, 22) and feats[1].shape == (2, 64, 5, 11)
    assert all(torch.isfinite(f).all() for f in feats)


@pytest.mark.gpu
def test_fused_neck_matches_separate_call_and_keeps_state_dict():
    """fuse_neck: the neck's launches run inside the backbone's CUDA graph; `neck(list(img_feats.values()))` returns those
    maps (bit-identical to the separate call), the backbone's state-dict keys are unchanged, reloading the neck's
    weights invalidates the captured graph."""
    from toc3d_b200 import TINY, EVA_ViT, ToC3DEVAViT
    from toc3d_b200 import lib as L
    from toc3d_b200.synthetic import make_gumbel, make_inputs
    torch.manual_seed(0)
    bb = ToC3DEVAViT(**TINY).eval().cuda()
    keys = sorted(bb.state_dict().keys())
    neck = CPFPN(in_channels=[TINY["embed_dim"]], out_channels=64, num_outs=2).eval().cuda()
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in make_inputs(1, 2, (160, 352), seed=3).items()}
    gn = make_gumbel(2, 220, seed=4)
    out = bb(**inp, gumbel_noise=gn)
    sep = neck(list(out.img_feats.values()))
    bb.fuse_neck(neck)
    assert sorted(bb.state_dict().keys()) == keys
    out2 = bb(**inp, gumbel_noise=gn)                       # eager path (injected noise)
    n0 = L.launch_count
    fused = neck(list(out2.img_feats.values()))
    assert L.launch_count == n0, "the fused neck must not launch anything in its own forward"
    assert torch.equal(out.img_feats["last_feat"], out2.img_feats["last_feat"])
    assert all(torch.equal(a, b) for a, b in zip(sep, fused)) and fused[1].shape == (2, 64, 5, 11)
    # graph path: replay twice, results finite and identical in shape; a plain tensor (not from this backbone) still works
    o3 = bb(**inp); f3 = neck(list(o3.img_feats.values()))
    o4 = bb(**inp); f4 = neck(list(o4.img_feats.values()))
    assert f3[0].shape == sep[0].shape and torch.isfinite(f4[0]).all() and f3[0].data_ptr() != f4[0].data_ptr()
    again = neck([out.img_feats["last_feat"].clone()])
    assert all(torch.equal(a, b) for a, b in zip(sep, again))
    assert len(bb._graphs) > 0
    neck.load_state_dict(neck.state_dict())
    assert len(bb._graphs) == 0
    # dense backbone
    d = EVA_ViT(**{k: v for k, v in TINY.items() if k not in ("pc_range", "pruning_num_queries", "pruning_loc", "accelerate_global",
                                                               "token_ratio", "token_selection_loss", "rope_acc")}).eval().cuda()
    x = inp["x"]
    a = neck([d(x)["last_feat"]])
    d.fuse_neck(neck)
    b = neck([d(x)["last_feat"]])
    assert all(torch.equal(u, v) for u, v in zip(a, b))


@pytest.mark.gpu
@pytest.mark.parametrize("V,H,W,ci,co", [(1, 5, 7, 64, 64), (3, 20, 50, 256, 256), (2, 13, 9, 128, 72)])
def test_implicit_conv3x3_gemm(V, H, W, ci, co):
    """The GEMM's implicit 3x3 convolution mode (toc3d_epilogue.conv_*: nine row-shifted TMA views of a zero-padded NHWC
    matrix, no im2col buffer) against F.conv2d on the same bf16-rounded operands, fp32 accumulation."""
    import torch.nn.functional as F
    from toc3d_b200 import lib as L
    g = torch.Generator().manual_seed(V * H * W + ci)
    x = (torch.randn(V, H, W, ci, generator=g) * 2).to(torch.bfloat16).float()
    w = (torch.randn(co, ci, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float()
    b = torch.randn(co, generator=g)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1).reshape(-1, co)
    Hp, Wp = H + 2, W + 2
    pad = torch.zeros(V, Hp, Wp, ci)
    pad[:, 1:-1, 1:-1] = x
    vv, yy, xx = torch.meshgrid(torch.arange(V), torch.arange(H), torch.arange(W), indexing="ij")
    to_pad = ((vv * Hp + yy + 1) * Wp + xx + 1).reshape(-1)
    from_pad = torch.full((V * Hp * Wp,), -1, dtype=torch.int32)
    from_pad[to_pad] = torch.arange(V * H * W, dtype=torch.int32)
    A = pad.reshape(-1, ci).cuda().bfloat16()
    Bw = w.permute(0, 2, 3, 1).reshape(co, 9 * ci).cuda().bfloat16().contiguous()
    out = torch.full((V * H * W, co), float("nan"), device="cuda")
    shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    if co % 8:
        pytest.skip("N must be a multiple of 8")
    L.gemm(A, Bw, L.EPI_LINEAR, M=A.shape[0], bias=b.cuda(), out=out, out_f32=True, out_map=from_pad.cuda(), conv_cin=ci,
           conv_row_shift=shifts)
    got = out.cpu()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    with pytest.raises(RuntimeError, match="9 \\* conv_cin"):
        L.gemm(A, Bw[:, : 8 * ci].contiguous(), L.EPI_LINEAR, M=A.shape[0], out=out, out_f32=True, conv_cin=ci, conv_row_shift=shifts)
