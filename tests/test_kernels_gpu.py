"""Kernel-level parity (GPU): every C-ABI entry point against the CPU oracle / fp32 torch math
on identical inputs.  Integer outputs must be bit-exact; bf16 tensor-core outputs carry the
tolerance written next to each check."""
import math

import pytest
import torch

from oracle import toc3d_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (200, 520, 192), (1000, 1024, 768),
                                   (8640, 3072, 1024), (333, 1024, 2752), (6000, 256, 1024)])
@pytest.mark.parametrize("f32", [False, True])
def test_gemm_linear(lib, M, N, K, f32):
    g = torch.Generator().manual_seed(M + N + K)
    A = bf16_round(torch.randn(M, K, generator=g))
    B = bf16_round(torch.randn(N, K, generator=g) * 0.05)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ B.double().t() + bias.double()
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32 if f32 else torch.bfloat16)
    lib.gemm(A.to(DEV).bfloat16(), B.to(DEV).bfloat16(), lib.EPI_LINEAR, bias=bias.to(DEV), out=out, out_f32=f32)
    torch.cuda.synchronize()
    got = out.float().cpu().double()
    assert torch.isfinite(got).all()
    tol = 2e-5 if f32 else 6e-3          # fp32 accumulate; bf16 output rounding 2^-8 relative
    assert rel_err(got, ref) < tol, rel_err(got, ref)


@pytest.mark.parametrize("tile_n", [64, 96, 128, 160, 192, 224, 256])
@pytest.mark.parametrize("M,N,K", [(700, 1024, 256), (129, 328, 64), (3000, 3072, 128)])
def test_gemm_tile_width_sweep(lib, tile_n, M, N, K):
    """Every legal pair-tile width gives the same result (runtime BN, CTA-pair tiles with row / column tails)."""
    g = torch.Generator().manual_seed(M + tile_n)
    A = bf16_round(torch.randn(M, K, generator=g))
    B = bf16_round(torch.randn(N, K, generator=g) * 0.05)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ B.double().t() + bias.double()
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
    lib.gemm(A.to(DEV).bfloat16(), B.to(DEV).bfloat16(), lib.EPI_LINEAR, bias=bias.to(DEV), out=out, out_f32=True,
             tile_n=tile_n)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) < 2e-5


def test_gemm_back_to_back_stream_order(lib):
    """Programmatic dependent launch: a chain of GEMMs where each consumes the previous output in place of A
    must equal the serial result (the device-side wait orders them)."""
    g = torch.Generator().manual_seed(11)
    M, C = 2000, 256
    x = bf16_round(torch.randn(M, C, generator=g))
    Ws = [bf16_round(torch.randn(C, C, generator=g) * 0.06) for _ in range(6)]
    ref = x.clone()
    for W in Ws:
        ref = bf16_round(ref @ W.t())
    a = x.to(DEV).bfloat16()
    bufs = [torch.empty(M, C, device=DEV, dtype=torch.bfloat16) for _ in range(2)]
    Wd = [W.to(DEV).bfloat16() for W in Ws]
    cur = a
    for i, W in enumerate(Wd):
        lib.gemm(cur, W, lib.EPI_LINEAR, out=bufs[i % 2])
        cur = bufs[i % 2]
    torch.cuda.synchronize()
    assert rel_err(cur.float().cpu(), ref) < 2e-2


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_linear_act(lib, act):
    g = torch.Generator().manual_seed(7)
    M, N, K = 300, 512, 128
    A = bf16_round(torch.randn(M, K, generator=g)); B = bf16_round(torch.randn(N, K, generator=g) * 0.1)
    bias = torch.randn(N, generator=g)
    pre = A @ B.t() + bias
    ref = torch.nn.functional.gelu(pre) if act == 1 else torch.relu(pre)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    lib.gemm(A.to(DEV).bfloat16(), B.to(DEV).bfloat16(), lib.EPI_LINEAR, bias=bias.to(DEV), out=out, out_f32=True, act=act)
    assert (out.cpu() - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("ft,use_rows", [(16, False), (20, False), (16, True), (20, True)])
def test_gemm_qkv_rope(lib, ft, use_rows):
    g = torch.Generator().manual_seed(ft)
    heads, C, K = 4, 256, 128
    n = ft * ft
    nW = 3
    seq = n if not use_rows else 101
    M = nW * seq
    A = bf16_round(torch.randn(M, K, generator=g)); Wt = bf16_round(torch.randn(3 * C, K, generator=g) * 0.1)
    bias = torch.randn(3 * C, generator=g); bias[C:2 * C] = 0
    cos, sin = O.rope_table(ft, 32, 16)
    if use_rows:
        rows = torch.stack([torch.randperm(n, generator=g)[:seq] for _ in range(nW)]).reshape(-1)
        rows[seq - 1::seq] = seq - 1
    else:
        rows = torch.arange(M) % n
    y = (A @ Wt.t() + bias).reshape(M, 3, heads, 64)
    q = O.apply_rope(y[:, 0], cos[rows][:, None], sin[rows][:, None]) * 0.125
    k = O.apply_rope(y[:, 1], cos[rows][:, None], sin[rows][:, None])
    ref = torch.stack([q, k, y[:, 2]], dim=1).reshape(M, 3 * C)
    cos_axis = cos.reshape(ft, ft, 64)[:, 0, 0:32:2].contiguous().to(DEV)
    sin_axis = sin.reshape(ft, ft, 64)[:, 0, 0:32:2].contiguous().to(DEV)
    out = torch.empty(M, 3 * C, device=DEV, dtype=torch.bfloat16)
    lib.gemm(A.to(DEV).bfloat16(), Wt.to(DEV).bfloat16(), lib.EPI_QKV_ROPE, bias=bias.to(DEV), out=out,
             rope_rows=rows.int().to(DEV) if use_rows else None, rope_slots=n, rope_ft=ft, rope_cols=2 * C,
             q_scale=0.125, cos_axis=cos_axis, sin_axis=sin_axis)
    assert rel_err(out.float().cpu(), ref) < 6e-3


def test_gemm_resid_maps(lib):
    g = torch.Generator().manual_seed(3)
    M, N, K, R = 300, 256, 128, 500
    A = bf16_round(torch.randn(M, K, generator=g)); Wt = bf16_round(torch.randn(N, K, generator=g) * 0.1)
    bias = torch.randn(N, generator=g)
    x = torch.randn(R, N, generator=g)
    alt = torch.randn(M, N, generator=g)
    perm = torch.randperm(R, generator=g)[:M].int()
    rmap = perm.clone(); rmap[::7] = -1; rmap[5::11] = -2
    omap = perm.clone(); omap[3::5] = -1; omap[5::11] = -2
    acc = A @ Wt.t() + bias
    ref_x, ref_alt = x.clone(), alt.clone()
    for m in range(M):
        r = x[rmap[m]] if rmap[m] >= 0 else (alt[m] if rmap[m] == -2 else torch.zeros(N))
        if omap[m] >= 0:
            ref_x[omap[m]] = r + acc[m]
        elif omap[m] == -2:
            ref_alt[m] = r + acc[m]
    xd, altd = x.to(DEV), alt.to(DEV)
    lib.gemm(A.to(DEV).bfloat16(), Wt.to(DEV).bfloat16(), lib.EPI_RESID, bias=bias.to(DEV), out=xd, resid=xd,
             resid_map=rmap.to(DEV), out_map=omap.to(DEV), out_alt=altd)
    assert (xd.cpu() - ref_x).abs().max().item() < 2e-4
    assert (altd.cpu() - ref_alt).abs().max().item() < 2e-4
    # identity maps + resid_mod (abs-pos broadcast)
    pos = torch.randn(100, N, generator=g)
    out = torch.empty(M, N, device=DEV)
    lib.gemm(A.to(DEV).bfloat16(), Wt.to(DEV).bfloat16(), lib.EPI_RESID, bias=bias.to(DEV), out=out,
             resid=pos.to(DEV), resid_mod=100)
    assert (out.cpu() - (acc + pos[torch.arange(M) % 100])).abs().max().item() < 2e-4


@pytest.mark.parametrize("tile_n", [0, 64, 128, 192, 256])
@pytest.mark.parametrize("M,Hd", [(257, 2730), (1000, 341)])
def test_gemm_swiglu(lib, M, Hd, tile_n):
    from toc3d_b200.backbone import interleave_w12, hidden_pad
    g = torch.Generator().manual_seed(Hd)
    K = 128
    Hp = hidden_pad(Hd)
    A = bf16_round(torch.randn(M, K, generator=g))
    w1 = bf16_round(torch.randn(Hd, K, generator=g) * 0.1); w2 = bf16_round(torch.randn(Hd, K, generator=g) * 0.1)
    b1 = torch.randn(Hd, generator=g); b2 = torch.randn(Hd, generator=g)
    ref = torch.nn.functional.silu(A @ w1.t() + b1) * (A @ w2.t() + b2)
    W12, b12 = interleave_w12(w1, b1, w2, b2, Hp)
    out = torch.full((M, Hp), float("nan"), device=DEV, dtype=torch.bfloat16)
    lib.gemm(A.to(DEV).bfloat16(), W12.to(DEV).bfloat16(), lib.EPI_SWIGLU, bias=b12.to(DEV), out=out, tile_n=tile_n)
    got = out.float().cpu()
    assert rel_err(got[:, :Hd], ref) < 8e-3
    assert (got[:, Hd:] == 0).all()


@pytest.mark.parametrize("M,Hd,C", [(300, 2730, 256), (1000, 341, 128)])
def test_gemm_swiglu_folded_subln(lib, M, Hd, C):
    """SwiGLU.forward (eva_vit.py:44-51) with the sub-LN folded into the two GEMM epilogues:
    w3(LN(h)) == rstd * (h @ (W3*gamma)^T) - rstd * mean * (W3 gamma) + (W3 beta + b3)."""
    from toc3d_b200.backbone import interleave_w12, hidden_pad
    F = torch.nn.functional
    g = torch.Generator().manual_seed(Hd + 1)
    K, Hp, eps = 128, hidden_pad(Hd), 1e-6
    A = bf16_round(torch.randn(M, K, generator=g))
    w1 = bf16_round(torch.randn(Hd, K, generator=g) * 0.1); w2 = bf16_round(torch.randn(Hd, K, generator=g) * 0.1)
    b1 = torch.randn(Hd, generator=g) * 0.5; b2 = torch.randn(Hd, generator=g) * 0.5 + 0.3
    gamma = 1 + 0.2 * torch.randn(Hd, generator=g); beta = 0.2 * torch.randn(Hd, generator=g)
    w3 = torch.randn(C, Hd, generator=g) * 0.05; b3 = torch.randn(C, generator=g)
    resid = torch.randn(M, C, generator=g)
    h = F.silu(A @ w1.t() + b1) * (A @ w2.t() + b2)
    ref = resid + F.layer_norm(h, (Hd,), gamma, beta, eps) @ w3.t() + b3
    W12, b12 = interleave_w12(w1, b1, w2, b2, Hp)
    hid = torch.empty(M, Hp, device=DEV, dtype=torch.bfloat16)
    stats = torch.full((M, 2), 7, device=DEV, dtype=torch.int64)
    # the LN launch that precedes the MLP zeroes the statistics rows
    lib.layernorm_rows(torch.randn(M, 128, device=DEV), torch.ones(128, device=DEV), torch.zeros(128, device=DEV),
                       torch.empty(M, 128, device=DEV, dtype=torch.bfloat16), M, 128, 1e-6, zero_stats=stats)
    assert (stats == 0).all()
    lib.gemm(A.to(DEV).bfloat16(), W12.to(DEV).bfloat16(), lib.EPI_SWIGLU, bias=b12.to(DEV), out=hid, row_stats=stats)
    hb = hid.float().cpu()[:, :Hd]
    st = stats.cpu().double()
    assert rel_err(st[:, 0] / 2 ** 30, hb.double().sum(1)) < 1e-5 and rel_err(st[:, 1] / 2 ** 26, hb.double().pow(2).sum(1)) < 1e-5
    stats2 = torch.zeros_like(stats)                                  # integer accumulation is order-independent
    lib.gemm(A.to(DEV).bfloat16(), W12.to(DEV).bfloat16(), lib.EPI_SWIGLU, bias=b12.to(DEV), out=hid, row_stats=stats2)
    assert torch.equal(stats, stats2)
    w3g = torch.nn.functional.pad(w3 * gamma[None, :], (0, Hp - Hd)).to(DEV).bfloat16().contiguous()
    u3 = (w3 @ gamma).to(DEV); c3 = (w3 @ beta + b3).to(DEV)
    out = torch.empty(M, C, device=DEV)
    lib.gemm(hid, w3g, lib.EPI_RESID, bias=c3, out=out, resid=resid.to(DEV), ln_stats=stats, ln_u=u3, ln_n=Hd, ln_eps=eps)
    err = (out.cpu() - ref).abs().max().item()
    print("folded sub-LN MLP max-abs %.4g (|ref| max %.3g)" % (err, ref.abs().max()))
    assert err < 8e-2 and ((out.cpu() - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item() < 4e-3


# ------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("seq", [1, 16, 64, 77, 103, 121, 129, 161, 180, 192, 193, 201, 256, 257, 281, 400, 401, 448, 449, 600])
def test_window_attention(lib, seq):
    """seq <= 448: tcgen05/TMEM kernel (256 or 512 TMEM columns, 1 or 2 S halves); above: mma.sync fallback."""
    g = torch.Generator().manual_seed(seq)
    nW, heads = 5, 3
    C = heads * 64
    qkv = bf16_round(torch.randn(nW * seq, 3 * C, generator=g))
    q, k, v = qkv.reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW * seq, C)
    out = torch.full((nW * seq, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    lib.window_attention(qkv.to(DEV).bfloat16(), out, nW, seq, heads)
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 2e-2, (got - ref).abs().max().item()   # P rounded to bf16, |v|~3


# ------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("C", [128, 1024])
@pytest.mark.parametrize("pad_mode", [0, 1])
def test_layernorm_rows(lib, C, pad_mode):
    g = torch.Generator().manual_seed(C)
    R, M = 300, 411
    x = torch.randn(R, C, generator=g) * 3 + 1
    alt = torch.randn(M, C, generator=g)
    gam = torch.randn(C, generator=g); bet = torch.randn(C, generator=g)
    rmap = torch.randint(0, R, (M,), generator=g).int(); rmap[::5] = -1; rmap[3::17] = -2
    src = torch.zeros(M, C)
    for m in range(M):
        src[m] = x[rmap[m]] if rmap[m] >= 0 else (alt[m] if rmap[m] == -2 else 0)
    ref = torch.nn.functional.layer_norm(src, (C,), gam, bet, 1e-6)
    if pad_mode == 0:
        ref[rmap == -1] = 0
    out = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    lib.layernorm_rows(x.to(DEV), gam.to(DEV), bet.to(DEV), out, M, C, 1e-6, row_map=rmap.to(DEV), alt=alt.to(DEV),
                       pad_mode=pad_mode)
    assert (out.float().cpu() - bf16_round(ref)).abs().max().item() <= 0.04   # 1 bf16 ulp at |y|<8
    out2 = torch.empty(R, C, device=DEV, dtype=torch.bfloat16)
    lib.layernorm_rows(x.to(DEV), gam.to(DEV), bet.to(DEV), out2, R, C, 1e-6)
    ref2 = torch.nn.functional.layer_norm(x, (C,), gam, bet, 1e-6)
    assert (out2.float().cpu() - bf16_round(ref2)).abs().max().item() <= 0.04


@pytest.mark.parametrize("Hd,ld", [(2730, 2752), (341, 352), (64, 64)])
def test_subln(lib, Hd, ld):
    g = torch.Generator().manual_seed(Hd)
    M = 257
    h = torch.zeros(M, ld); h[:, :Hd] = torch.randn(M, Hd, generator=g) * 2 + 0.3
    h = bf16_round(h)
    gam = torch.randn(Hd, generator=g); bet = torch.randn(Hd, generator=g)
    ref = torch.nn.functional.layer_norm(h[:, :Hd], (Hd,), gam, bet, 1e-6)
    gp = torch.zeros(ld); gp[:Hd] = gam; bp = torch.zeros(ld); bp[:Hd] = bet
    out = torch.full((M, ld), float("nan"), device=DEV, dtype=torch.bfloat16)
    lib.subln(h.to(DEV).bfloat16(), out, gp.to(DEV), bp.to(DEV), M, Hd, ld, 1e-6)
    got = out.float().cpu()
    assert (got[:, :Hd] - bf16_round(ref)).abs().max().item() <= 0.04
    assert (got[:, Hd:] == 0).all()


# ------------------------------------------------------------------------------------ selection
def _adversarial_scores(V, H, W, kind, g):
    s = -torch.rand(V, H, W, generator=g) * 5
    if kind == "equal":
        s[:] = -0.5
    elif kind == "dups":
        s = (s * 4).round() / 4
    elif kind == "zeros":
        s = torch.where(torch.rand(V, H, W, generator=g) < 0.5, torch.zeros(()), -torch.zeros(()))
    elif kind == "ramp_up":
        s = torch.arange(V * H * W, dtype=torch.float32).reshape(V, H, W) * 1e-3 - 10
    elif kind == "ramp_down":
        s = -torch.arange(V * H * W, dtype=torch.float32).reshape(V, H, W) * 1e-3
    elif kind == "pads_tie":
        s[:, ::2] = O.PAD_SCORE      # real tokens that tie with the pad value
    elif kind == "special":          # NaN ranks first in torch.sort(descending=True); +-inf, denormals, +-0 ties
        r = torch.rand(V, H, W, generator=g)
        for lo, val in ((0.00, float("nan")), (0.08, float("inf")), (0.16, float("-inf")), (0.24, 1e-42),
                        (0.32, -1e-42), (0.40, 0.0), (0.48, -0.0)):
            s = torch.where((r >= lo) & (r < lo + 0.08), torch.full((), val), s)
    return s


@pytest.mark.parametrize("kind", ["random", "equal", "dups", "zeros", "ramp_up", "ramp_down", "pads_tie", "special"])
@pytest.mark.parametrize("H,W,ws,ratio", [(20, 50, 16, 0.7), (20, 50, 20, 0.5), (20, 50, 20, 0.4), (20, 50, 16, 0.3),
                                          (50, 100, 16, 0.5), (50, 100, 20, 0.3), (7, 9, 16, 0.5)])
def test_window_topk_bit_exact(lib, kind, H, W, ws, ratio):
    g = torch.Generator().manual_seed(H * W + ws)
    V = 3
    s = _adversarial_scores(V, H, W, kind, g)
    n = ws * ws
    k = int(n * ratio)
    sw, _ = O.window_partition(s[..., None], ws, pad_value=O.PAD_SCORE)
    sw = sw.reshape(-1, n)
    _, fast_s, slow_idx, fast_idx = O.sample(sw, ratio)
    nW = sw.shape[0]
    d = dict(slow_idx=torch.full((nW, k), -9, dtype=torch.int32, device=DEV),
             fast_idx=torch.full((nW, n - k), -9, dtype=torch.int32, device=DEV),
             fast_score=torch.zeros(nW, n - k, device=DEV),
             tok_map=torch.full((nW * (k + 1),), -9, dtype=torch.int32, device=DEV),
             rope_rows=torch.full((nW * (k + 1),), -9, dtype=torch.int32, device=DEV),
             fast_map=torch.full((nW, n - k), -9, dtype=torch.int32, device=DEV))
    lib.window_topk(s.to(DEV), V, H, W, ws, k, **d)
    assert torch.equal(d["slow_idx"].cpu().long(), slow_idx)
    assert torch.equal(d["fast_idx"].cpu().long(), fast_idx)
    assert torch.equal(d["fast_score"].cpu().view(torch.int32), fast_s.contiguous().view(torch.int32))   # bit copy
    # derived maps: slot -> image row
    rows = torch.arange(V * H * W, dtype=torch.float32).reshape(V, H, W, 1)
    rw, _ = O.window_partition(rows, ws, pad_value=-1.0)
    rw = rw.reshape(nW, n).long()
    tok = torch.cat([torch.gather(rw, 1, slow_idx), torch.full((nW, 1), -2)], 1).reshape(-1)
    assert torch.equal(d["tok_map"].cpu().long(), tok)
    rr = torch.cat([slow_idx, torch.full((nW, 1), k)], 1).reshape(-1)
    assert torch.equal(d["rope_rows"].cpu().long(), rr)
    assert torch.equal(d["fast_map"].cpu().long(), torch.gather(rw, 1, fast_idx))


@pytest.mark.parametrize("kind", ["random", "equal", "dups", "zeros", "special"])
@pytest.mark.parametrize("N,ratio", [(1000, 0.7), (1000, 0.3), (5000, 0.5), (5000, 0.4), (63, 0.5)])
def test_topk_split_bit_exact(lib, kind, N, ratio):
    g = torch.Generator().manual_seed(N)
    B = 6
    s = _adversarial_scores(B, 1, N, kind, g).reshape(B, N)
    k = int(N * ratio)
    _, _, keep, drop = O.sample(s, ratio)
    ki = torch.full((B, k), -9, dtype=torch.int64, device=DEV)
    di = torch.full((B, N - k), -9, dtype=torch.int64, device=DEV)
    lib.topk_split(s.to(DEV), B, N, k, ki, di)
    assert torch.equal(ki.cpu(), keep) and torch.equal(di.cpu(), drop)


def test_merge_and_fast_update(lib):
    g = torch.Generator().manual_seed(11)
    V, H, W, ws, C, ratio = 2, 20, 50, 16, 256, 0.7
    n, k = ws * ws, int(ws * ws * ratio)
    x = torch.randn(V, H, W, C, generator=g)
    s = -torch.rand(V, H, W, generator=g) * 4
    xw, _ = O.window_partition(x, ws); xw = xw.reshape(-1, n, C)
    sw, _ = O.window_partition(s[..., None], ws, pad_value=O.PAD_SCORE); sw = sw.reshape(-1, n)
    _, fast_s, slow_idx, fast_idx = O.sample(sw, ratio)
    rep_ref = O.merge_tokens(O.batch_index_select(xw, fast_idx), fast_s)[:, 0]
    nW, nf = sw.shape[0], n - k
    fs = torch.empty(nW, nf, device=DEV); fm = torch.empty(nW, nf, dtype=torch.int32, device=DEV)
    lib.window_topk(s.to(DEV), V, H, W, ws, k, fast_score=fs, fast_map=fm)
    xd = x.reshape(-1, C).to(DEV).contiguous()
    rep = torch.empty(nW, C, device=DEV); packed = torch.zeros(nW * (k + 1), C, device=DEV)
    lib.merge_fast_tokens(xd, fm, fs, nW, nf, k, C, rep, packed)
    assert (rep.cpu() - rep_ref).abs().max().item() < 1e-5 * max(1.0, rep_ref.abs().max().item())
    assert torch.equal(packed.reshape(nW, k + 1, C)[:, k], rep)
    # fast update: packed rep rows hold t2_rep
    delta = torch.randn(nW, C, generator=g)
    packed.reshape(nW, k + 1, C)[:, k] = rep + delta.to(DEV)
    lib.fast_token_update(xd, fm, packed, rep, nW, nf, k, C)
    out_w = xw.clone()
    fi = fast_idx[..., None].expand(-1, -1, C)
    out_w.scatter_(1, fi, torch.gather(xw, 1, fi) + delta[:, None])
    ref = O.window_unpartition(out_w.reshape(-1, ws, ws, C), ws, (32, 64), (H, W)).reshape(-1, C)
    assert (xd.cpu() - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("C,ws,ratio", [(1024, 16, 0.7), (1024, 20, 0.5), (128, 16, 0.3), (256, 20, 0.4)])
def test_ln_gather_merge_fused(lib, C, ws, ratio):
    """One launch = batch_index_select (slow + fast), merge_tokens and norm1 of the packed rows
    (toc3d_eva_vit.py:421-427, :371; pad slots are zero vectors whose LayerNorm is the bias)."""
    g = torch.Generator().manual_seed(C + ws)
    V, H, W = 2, 20, 50
    n, k = ws * ws, int(ws * ws * ratio)
    x = torch.randn(V, H, W, C, generator=g) * 3 + 0.5
    s = -torch.rand(V, H, W, generator=g) * 4
    gamma = 1 + 0.1 * torch.randn(C, generator=g); beta = 0.1 * torch.randn(C, generator=g)
    xw, _ = O.window_partition(x, ws); xw = xw.reshape(-1, n, C)
    sw, _ = O.window_partition(s[..., None], ws, pad_value=O.PAD_SCORE); sw = sw.reshape(-1, n)
    _, fast_s, slow_idx, fast_idx = O.sample(sw, ratio)
    rep_ref = O.merge_tokens(O.batch_index_select(xw, fast_idx), fast_s)
    t_ref = torch.cat([O.batch_index_select(xw, slow_idx), rep_ref], dim=1)          # (nW, k+1, C)
    ln_ref = torch.nn.functional.layer_norm(t_ref, (C,), gamma, beta, 1e-6).reshape(-1, C)
    nW, nf = sw.shape[0], n - k
    i32 = dict(dtype=torch.int32, device=DEV)
    fs = torch.empty(nW, nf, device=DEV); fm = torch.empty(nW, nf, **i32)
    tok = torch.empty(nW * (k + 1), **i32)
    lib.window_topk(s.to(DEV), V, H, W, ws, k, fast_score=fs, fast_map=fm, tok_map=tok)
    xd = x.reshape(-1, C).to(DEV).contiguous()
    out = torch.full((nW * (k + 1), C), float("nan"), device=DEV, dtype=torch.bfloat16)
    rep = torch.empty(nW, C, device=DEV); packed = torch.zeros(nW * (k + 1), C, device=DEV)
    stats = torch.ones(nW * (k + 1), 2, device=DEV, dtype=torch.int64)
    cnt = torch.zeros(nW, device=DEV, dtype=torch.int32)
    lib.ln_gather_merge(xd, tok, fm, fs, gamma.to(DEV), beta.to(DEV), out, rep, packed, nW, k, nf, C, 1e-6, zero_stats=stats,
                        counters=cnt)
    assert (cnt == 0).all()
    assert (rep.cpu() - rep_ref[:, 0]).abs().max().item() < 1e-5 * max(1.0, rep_ref.abs().max().item())
    assert torch.equal(packed.reshape(nW, k + 1, C)[:, k], rep)
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    assert (got - ln_ref).abs().max().item() < 2e-2 * max(1.0, ln_ref.abs().max().item() / 4)      # bf16 output
    assert (stats == 0).all()


def test_compact_rows_and_pad_fill(lib):
    """Compact row space of an accelerated block + constant k / v of dense pad slots."""
    g = torch.Generator().manual_seed(5)
    V, H, W, ws, ratio = 2, 20, 50, 16, 0.7
    n, k = ws * ws, int(ws * ws * ratio)
    s = -torch.rand(V, H, W, generator=g) * 4
    nW = V * 2 * 4
    i32 = dict(dtype=torch.int32, device=DEV)
    tok = torch.empty(nW * (k + 1), **i32)
    lib.window_topk(s.to(DEV), V, H, W, ws, k, tok_map=tok)
    real = torch.tensor([256, 256, 256, 32, 64, 64, 64, 8] * V, dtype=torch.int32)
    rcap = torch.minimum(real, torch.tensor(k, dtype=torch.int32))
    coff = torch.cumsum(rcap + 1, 0, dtype=torch.int32) - (rcap + 1)
    Mc = int((rcap + 1).sum())
    rope = torch.empty(nW * (k + 1), **i32)
    lib.window_topk(s.to(DEV), V, H, W, ws, k, tok_map=tok, rope_rows=rope)
    cmap = torch.full((nW * (k + 1),), -7, **i32); ctok = torch.full((Mc,), -7, **i32); rep_row = torch.full((nW,), -7, **i32)
    cinv = torch.full((Mc,), -7, **i32); crope = torch.full((Mc,), -7, **i32); prope = torch.full((nW * (k + 1),), -7, **i32)
    lib.compact_rows(tok, coff.to(DEV), rcap.to(DEV), nW, k, cmap, ctok, rep_row, rope_rows=rope, cinv=cinv, crope=crope, prope=prope)
    tok_c, rope_c = tok.cpu().view(nW, k + 1), rope.cpu().view(nW, k + 1)
    cmap_c, prope_c, ctok_c, cinv_c, crope_c = cmap.cpu().view(nW, k + 1), prope.cpu().view(nW, k + 1), ctok.cpu(), cinv.cpu(), crope.cpu()
    for w in range(nW):
        r = int(rcap[w]); c0 = int(coff[w])
        assert (tok_c[w, :r] >= 0).all() and (tok_c[w, r:k] == -1).all()          # real rows outrank pads
        # packed layout [real | rep | pads]
        assert cmap_c[w, :r].tolist() == list(range(c0, c0 + r)) and cmap_c[w, r] == c0 + r and (cmap_c[w, r + 1:] == -1).all()
        assert torch.equal(prope_c[w, :r], rope_c[w, :r]) and prope_c[w, r] == k and torch.equal(prope_c[w, r + 1:], rope_c[w, r:k])
        assert rep_row[w].item() == c0 + r and ctok_c[c0 + r] == -2
        assert torch.equal(ctok_c[c0:c0 + r], tok_c[w, :r]) and torch.equal(crope_c[c0:c0 + r], rope_c[w, :r]) and crope_c[c0 + r] == k
        assert cinv_c[c0:c0 + r + 1].tolist() == list(range(w * (k + 1), w * (k + 1) + r + 1))
    # dense pad slots
    C = 128
    qkv = torch.full((40, 3 * C), 7.0, device=DEV, dtype=torch.bfloat16)
    pads = torch.tensor([3, 17, 39], **i32)
    vb = torch.randn(C, generator=g).to(DEV)
    lib.fill_pad_kv(qkv, pads, vb, C)
    q = qkv.float().cpu()
    assert (q[[3, 17, 39], C:2 * C] == 0).all() and torch.equal(q[[3, 17, 39], 2 * C:], bf16_round(vb.cpu()).expand(3, C))
    assert (q[[3, 17, 39], :C] == 7).all() and (q[[0, 1, 2, 4]] == 7).all()


@pytest.mark.parametrize("ws_prev,ws", [(16, 16), (16, 20), (20, 16)])
def test_deferred_fast_update_in_gather_merge(lib, ws_prev, ws):
    """toc3d_ln_gather_merge with a pending update == toc3d_fast_token_update of the previous block followed by the plain
    launch, bit for bit (residual stream, LayerNorm rows, representative tokens); also window_topk's fast_win table."""
    g = torch.Generator().manual_seed(ws_prev * 31 + ws)
    V, H, W, C, ratio = 2, 20, 50, 256, 0.6
    N = H * W
    i32 = dict(dtype=torch.int32, device=DEV)
    s = (-torch.rand(V, H, W, generator=g) * 4).to(DEV)

    def tables(ws_):
        n, k = ws_ * ws_, int(ws_ * ws_ * ratio)
        nWh, nWw = -(-H // ws_), -(-W // ws_)
        nW = V * nWh * nWw
        real = torch.zeros(nW, dtype=torch.int32)
        for v in range(V):
            for a in range(nWh):
                for b in range(nWw):
                    real[(v * nWh + a) * nWw + b] = min(ws_, H - a * ws_) * min(ws_, W - b * ws_)
        rcap = torch.minimum(real, torch.tensor(k, dtype=torch.int32))
        coff = torch.cumsum(rcap + 1, 0, dtype=torch.int32) - (rcap + 1)
        Mc = int((rcap + 1).sum())
        t = dict(n=n, k=k, nf=n - k, nW=nW, Mc=Mc, tok=torch.empty(nW * (k + 1), **i32), rope=torch.empty(nW * (k + 1), **i32),
                 fmap=torch.empty(nW, n - k, **i32), fsc=torch.empty(nW, n - k, device=DEV), fwin=torch.full((V * N,), -9, **i32),
                 cmap=torch.empty(nW * (k + 1), **i32), ctok=torch.empty(Mc, **i32), rep_row=torch.empty(nW, **i32))
        lib.window_topk(s, V, H, W, ws_, k, fast_score=t["fsc"], tok_map=t["tok"], rope_rows=t["rope"], fast_map=t["fmap"], fast_win=t["fwin"])
        lib.compact_rows(t["tok"], coff.to(DEV), rcap.to(DEV), nW, k, t["cmap"], t["ctok"], t["rep_row"], rope_rows=t["rope"])
        return t
    tp, tn = tables(ws_prev), tables(ws)
    # fast_win: image row -> window of its fast_map entry, -1 for slow rows, every row written
    fw = tp["fwin"].cpu(); fm = tp["fmap"].cpu()
    want = torch.full((V * N,), -1, dtype=torch.int32)
    for w_ in range(tp["nW"]):
        rows = fm[w_][fm[w_] >= 0].long()
        want[rows] = w_
    assert torch.equal(fw, want)
    x0 = torch.randn(V * N, C, generator=g).to(DEV) * 3
    T_prev = torch.randn(tp["Mc"], C, generator=g).to(DEV)          # previous block's compact rows after its MLP
    rep_prev = torch.randn(tp["nW"], C, generator=g).to(DEV)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV); beta = (0.1 * torch.randn(C, generator=g)).to(DEV)
    cnt = torch.zeros(tn["nW"], **i32)

    def run(deferred):
        x = x0.clone()
        out = torch.zeros(tn["Mc"], C, device=DEV, dtype=torch.bfloat16)
        rep = torch.zeros(tn["nW"], C, device=DEV); T = torch.zeros(tn["Mc"], C, device=DEV)
        if not deferred:
            lib.fast_token_update(x, tp["fmap"], T_prev, rep_prev, tp["nW"], tp["nf"], tp["k"], C, rep_row=tp["rep_row"])
        lib.ln_gather_merge(x, tn["ctok"], tn["fmap"], tn["fsc"], gamma, beta, out, rep, T, tn["nW"], tn["k"], tn["nf"], C, 1e-6,
                            rep_row=tn["rep_row"], compact_rows=tn["Mc"], counters=cnt,
                            pending=(tp["fwin"], T_prev, tp["rep_row"], rep_prev) if deferred else None)
        torch.cuda.synchronize()
        return x, out, rep, T
    a, b = run(False), run(True)
    assert not torch.equal(a[0], x0)                                  # the update did something
    for u, v, name in zip(a, b, ("x", "ln rows", "rep", "packed")):
        assert torch.equal(u, v), name
    with pytest.raises(RuntimeError, match="ping-pong"):
        Tn = torch.zeros(max(tn["Mc"], tp["Mc"]), C, device=DEV)
        lib.ln_gather_merge(x0.clone(), tn["ctok"], tn["fmap"], tn["fsc"], gamma, beta, torch.zeros(tn["Mc"], C, device=DEV, dtype=torch.bfloat16),
                            torch.zeros(tn["nW"], C, device=DEV), Tn, tn["nW"], tn["k"], tn["nf"], C, 1e-6, rep_row=tn["rep_row"],
                            compact_rows=tn["Mc"], counters=cnt, pending=(tp["fwin"], Tn, tp["rep_row"], rep_prev))


@pytest.mark.parametrize("ft", [16, 20])
def test_fill_pad_kv_rope_matches_qkv_gemm(lib, ft):
    """Pad rows of an accelerated block (norm1(0) = beta): k / v from the block constants + per-slot rotation equal
    what the QKV GEMM epilogue produces for an A row holding bf16(beta) (up to one bf16 rounding of k)."""
    g = torch.Generator().manual_seed(ft)
    heads, C = 2, 128
    n = ft * ft
    Mp = 200
    beta = bf16_round(torch.randn(C, generator=g) * 0.3)
    Wt = bf16_round(torch.randn(3 * C, C, generator=g) * 0.1)
    bias = torch.randn(3 * C, generator=g); bias[C:2 * C] = 0
    rows = torch.randint(0, n, (Mp,), generator=g).int()
    cos, sin = O.rope_table(ft, 32, 16)
    ca = cos.reshape(ft, ft, 64)[:, 0, 0:32:2].contiguous().to(DEV); sa = sin.reshape(ft, ft, 64)[:, 0, 0:32:2].contiguous().to(DEV)
    A = beta.expand(Mp, C).contiguous().to(DEV).bfloat16()
    ref = torch.empty(Mp, 3 * C, device=DEV, dtype=torch.bfloat16)
    lib.gemm(A, Wt.to(DEV).bfloat16(), lib.EPI_QKV_ROPE, bias=bias.to(DEV), out=ref, rope_rows=rows.to(DEV), rope_ft=ft,
             rope_cols=2 * C, q_scale=0.125, cos_axis=ca, sin_axis=sa)
    kpad = (Wt[C:2 * C] @ beta).to(DEV); vpad = (Wt[2 * C:] @ beta + bias[2 * C:]).to(DEV)
    cmap = torch.full((Mp,), -1, dtype=torch.int32); cmap[::3] = 5          # every third row is "real": left alone
    out = torch.full((Mp, 3 * C), 9.0, device=DEV, dtype=torch.bfloat16)
    lib.fill_pad_kv_rope(out, cmap.to(DEV), rows.to(DEV), Mp, kpad, vpad, ca, sa, ft, C)
    got, want = out.float().cpu(), ref.float().cpu()
    pad = cmap == -1
    assert (got[~pad] == 9).all() and (got[pad][:, :C] == 9).all()
    assert (got[pad][:, C:] - want[pad][:, C:]).abs().max().item() <= 2 ** -7 * want[pad][:, C:].abs().max().item()


@pytest.mark.parametrize("seq", [129, 130, 180, 256])
def test_attention_tail_row_129(lib, seq):
    """Windows that need exactly 128 + 1 query rows (k = 128 slow tokens + representative: the second query tile holds ONE
    row), mixed with windows needing 1, 128, 130 and all rows, through compact out maps, many items per CTA (ring
    reuse), with and without the balanced item order, against the fp32 reference."""
    g = torch.Generator().manual_seed(seq)
    nW, heads = 40, 4
    C = heads * 64
    qkv = bf16_round(torch.randn(nW * seq, 3 * C, generator=g))
    q, k, v = qkv.reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW, seq, C)
    choices = [129, 129, 129, 1, 128, min(seq, 130), seq, 129]
    qr = torch.tensor([min(seq, choices[i % len(choices)]) for i in range(nW)], dtype=torch.int32)
    omap = torch.full((nW * seq,), -1, dtype=torch.int32)
    n = 0
    for w in range(nW):                                   # compact destination rows: only the needed query rows
        for r in range(int(qr[w])):
            omap[w * seq + r] = n
            n += 1
    order = torch.argsort(((qr + 127) // 128).repeat_interleave(heads), descending=True, stable=True).int()
    for io in (None, order):
        out = torch.full((n, C), float("nan"), device=DEV, dtype=torch.bfloat16)
        lib.window_attention(qkv.to(DEV).bfloat16(), out, nW, seq, heads, out_map=omap.to(DEV), q_rows=qr.to(DEV),
                             item_order=None if io is None else io.to(DEV))
        got = out.float().cpu()
        assert torch.isfinite(got).all()
        for w in range(nW):
            r = int(qr[w])
            rows = omap[w * seq: w * seq + r].long()
            err = (got[rows] - ref[w, :r]).abs().max().item()
            assert err < 3e-2, (w, r, err)


@pytest.mark.parametrize("analytic", [False, True])
@pytest.mark.parametrize("seq", [103, 129, 144, 180, 192, 201, 224, 256])
def test_attention_many_items_per_cta(lib, seq, analytic):
    """Persistent kernel under load: ~6 (window, head) items per CTA (ring of item buffers reused several times, both
    TMEM buffers alternating through units of different key counts and tile counts), ragged q_rows / kv_rows, balanced
    item order - against the fp32 reference (computed on the GPU in fp32)."""
    g = torch.Generator().manual_seed(seq + 1000 * analytic)
    nW, heads = 150, 6
    C = heads * 64
    rows = [seq, 1, 33, min(seq, 65), seq, min(seq, 129), 9, seq, min(seq, 128), min(seq, 17), seq, min(seq, 161)]
    qr = torch.tensor([rows[i % len(rows)] for i in range(nW)], dtype=torch.int32)
    qkv = bf16_round(torch.randn(nW, seq, 3 * C, generator=g))
    ref_in = qkv.clone()
    vb = torch.randn(C, generator=g) * 0.5
    if analytic:                                           # keys beyond kv = q_rows are pad slots: k = 0, v = v_bias
        for w in range(nW):
            ref_in[w, int(qr[w]):, C:2 * C] = 0.0
            ref_in[w, int(qr[w]):, 2 * C:] = vb
            qkv[w, int(qr[w]):, C:] = 7.0                  # finite garbage that must not be read
    q, k, v = ref_in.to(DEV).reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW, seq, C).cpu()
    order = torch.argsort(((qr + 127) // 128).repeat_interleave(heads), descending=True, stable=True).int()
    kw = dict(kv_rows=qr.to(DEV), pad_v=vb.to(DEV)) if analytic else {}
    for io in (None, order):
        out = torch.full((nW * seq, C), float("nan"), device=DEV, dtype=torch.bfloat16)
        lib.window_attention(qkv.reshape(nW * seq, 3 * C).to(DEV).bfloat16(), out, nW, seq, heads, q_rows=qr.to(DEV),
                             item_order=None if io is None else io.to(DEV), **kw)
        got = out.float().cpu().view(nW, seq, C)
        for w in range(nW):
            r = int(qr[w])
            assert torch.isfinite(got[w, :r]).all(), (w, r)
            err = (got[w, :r] - ref[w, :r]).abs().max().item()
            assert err < 3e-2, (w, r, err)


def test_attention_more_items_than_the_smem_item_list(lib):
    """> 64 (window, head) items per CTA: the persistent kernel stages the metadata of a CTA's first 64 items in shared
    memory and fetches the rest from the index tables on the fly."""
    g = torch.Generator().manual_seed(7)
    seq, nW, heads = 33, 1700, 6                           # 10 200 items: 69 per CTA on 148 SMs
    C = heads * 64
    qr = torch.tensor([(1, 33, 17, 32)[i % 4] for i in range(nW)], dtype=torch.int32)
    qkv = bf16_round(torch.randn(nW, seq, 3 * C, generator=g))
    q, k, v = qkv.to(DEV).reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW, seq, C).cpu()
    out = torch.full((nW * seq, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    lib.window_attention(qkv.reshape(nW * seq, 3 * C).to(DEV).bfloat16(), out, nW, seq, heads, q_rows=qr.to(DEV))
    got = out.float().cpu().view(nW, seq, C)
    mask = torch.arange(seq)[None, :] < qr[:, None]
    assert torch.isfinite(got[mask]).all()
    assert (got[mask] - ref[mask]).abs().max().item() < 3e-2


@pytest.mark.parametrize("seq", [64, 180, 256, 400, 448])
def test_attention_analytic_pad_keys(lib, seq):
    """Dense-block pad slots (k = 0, v = v_bias, eva_vit.py:249-254) as ONE closed-form softmax term: kv_rows[w] real keys
    are staged, the seq - kv_rows[w] pads are never read (their qkv rows hold finite garbage here) - against the fp32
    attention over the explicitly padded window."""
    g = torch.Generator().manual_seed(seq)
    nW, heads = 12, 3
    C = heads * 64
    vb = torch.randn(C, generator=g) * 0.5
    kvs = [seq, 1, 8, 32, min(seq, 64), min(seq, 65), seq // 2, seq - 1, min(seq, 200), 17, seq, min(seq, 129)]
    kv = torch.tensor([max(1, k) for k in kvs], dtype=torch.int32)
    qkv = bf16_round(torch.randn(nW, seq, 3 * C, generator=g))
    ref_in = qkv.clone()
    for w in range(nW):                                    # the reference sees the true pad slots
        ref_in[w, int(kv[w]):, C:2 * C] = 0.0
        ref_in[w, int(kv[w]):, 2 * C:] = vb
        qkv[w, int(kv[w]):, C:] = bf16_round(torch.randn(seq - int(kv[w]), 2 * C, generator=g) * 30)   # garbage: must not matter
    q, k, v = ref_in.reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW, seq, C)
    out = torch.zeros(nW * seq, C, device=DEV, dtype=torch.bfloat16)
    lib.window_attention(qkv.reshape(nW * seq, 3 * C).to(DEV).bfloat16(), out, nW, seq, heads, q_rows=kv.to(DEV), kv_rows=kv.to(DEV),
                         pad_v=vb.to(DEV))
    got = out.float().cpu().reshape(nW, seq, C)
    for w in range(nW):
        r = int(kv[w])
        err = (got[w, :r] - ref[w, :r]).abs().max().item()
        assert err < 3e-2, (w, r, err)
    with pytest.raises(RuntimeError, match="go together"):
        lib.window_attention(qkv.reshape(nW * seq, 3 * C).to(DEV).bfloat16(), out, nW, seq, heads, kv_rows=kv.to(DEV))


@pytest.mark.parametrize("seq", [77, 180, 256, 300, 401, 500])
def test_attention_q_rows_prefix(lib, seq):
    """q_rows: only the leading query rows of each window are computed / stored; all rows remain keys."""
    g = torch.Generator().manual_seed(seq)
    nW, heads = 6, 2
    C = heads * 64
    qkv = bf16_round(torch.randn(nW * seq, 3 * C, generator=g))
    q, k, v = qkv.reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW, seq, C)
    qr = torch.tensor([1, seq, min(seq, 128), min(seq, 129), seq // 2, min(seq, 65)], dtype=torch.int32)
    out = torch.full((nW * seq, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    lib.window_attention(qkv.to(DEV).bfloat16(), out, nW, seq, heads, q_rows=qr.to(DEV))
    got = out.float().cpu().view(nW, seq, C)
    for w in range(nW):
        n = int(qr[w])
        assert torch.isfinite(got[w, :n]).all()
        assert (got[w, :n] - ref[w, :n]).abs().max().item() < 2e-2
        # whole 128-row (64 for the fallback kernel) query tiles beyond the prefix are skipped
        tile = 128 if seq <= 448 else 64
        first_skipped = -(-n // tile) * tile
        assert torch.isnan(got[w, first_skipped:]).all()


def test_attention_and_qkv_row_maps(lib):
    """out_map of the attention (rows stored in compact order, -1 skipped) and of the QKV epilogue (rows scattered)."""
    g = torch.Generator().manual_seed(9)
    nW, seq, heads = 4, 77, 2
    C = heads * 64
    qkv = bf16_round(torch.randn(nW * seq, 3 * C, generator=g))
    q, k, v = qkv.reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW * seq, C)
    keep = torch.rand(nW * seq, generator=g) > 0.4
    omap = torch.full((nW * seq,), -1, dtype=torch.int32)
    omap[keep] = torch.randperm(int(keep.sum()), generator=g).int()
    out = torch.full((int(keep.sum()), C), float("nan"), device=DEV, dtype=torch.bfloat16)
    lib.window_attention(qkv.to(DEV).bfloat16(), out, nW, seq, heads, out_map=omap.to(DEV))
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    assert (got[omap[keep].long()] - ref[keep]).abs().max().item() < 2e-2
    # GEMM bf16 rows through out_map
    M, N, K = 300, 256, 128
    A = bf16_round(torch.randn(M, K, generator=g)); Wt = bf16_round(torch.randn(N, K, generator=g) * 0.1)
    perm = torch.randperm(400, generator=g)[:M].int(); perm[::9] = -1
    o = torch.zeros(400, N, device=DEV, dtype=torch.bfloat16)
    lib.gemm(A.to(DEV).bfloat16(), Wt.to(DEV).bfloat16(), lib.EPI_LINEAR, out=o, out_map=perm.to(DEV))
    oc = o.float().cpu(); full = bf16_round(A @ Wt.t())
    ok = perm >= 0
    assert (oc[perm[ok].long()] - full[ok]).abs().max().item() < 2e-2
    untouched = torch.ones(400, dtype=torch.bool); untouched[perm[ok].long()] = False
    assert (oc[untouched] == 0).all()


# ------------------------------------------------------------------------------------ scorer
def _selector_case(seed, Bf, Q=64, C=1024, f64_time=True, bias_std=0.1):
    """A randomised MotionAwareQueryGuidedTokenSelector parameter set + history-query inputs (rigid random poses)."""
    from toc3d_b200.backbone import _Selector, pack_motion_blob
    from toc3d_b200.configs import PC_RANGE
    from toc3d_b200.synthetic import make_inputs, randomize_state_dict
    torch.manual_seed(seed)
    sel = _Selector(C, Q, 0.7, PC_RANGE)
    sd = randomize_state_dict(sel.state_dict(), seed=seed, bias_std=bias_std, weight_std=0.05)
    sd["ego_pose_pe.gamma.weight"] = sd["ego_pose_pe.gamma.weight"] * 2          # zero-init in the reference: make them count
    sel.load_state_dict(sd)
    p = {"score_predictor.0." + k: v for k, v in sel.state_dict().items()}
    inp = make_inputs(Bf, 1, (32, 32), seed=seed, num_queries=Q, pose="random")
    if not f64_time:
        inp["temp_timestamp"] = inp["temp_timestamp"].float()
    inp["temp_timestamp"] = inp["temp_timestamp"] * 3.0 - 0.5                     # beyond [0, 1): several periods of the embedding
    return sel, p, inp, pack_motion_blob(sel, Q, C)


@pytest.mark.parametrize("Bf,Q,f64_time", [(1, 64, True), (2, 64, False), (3, 30, True)])
def test_motion_queries_and_fold(lib, Bf, Q, f64_time):
    """toc3d_motion_queries_fold (row a12): encoded queries against the oracle's get_motion_aware_queries restatement
    (toc3d_utils.py:334-360) at 1e-5 relative, folded (A, c) against the fp64 product; two stages in one call."""
    C = 1024
    sel0, p0, inp, blob0 = _selector_case(41, Bf, Q, C, f64_time)
    sel1, p1, _, blob1 = _selector_case(42, Bf, Q, C, f64_time)
    blob = torch.stack([blob0, blob1]).to(DEV)
    q_out = torch.empty(2, Bf, Q, 256, device=DEV); A = torch.empty(2, Bf, 2, C, device=DEV); c = torch.empty(2, Bf, 2, device=DEV)
    d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in inp.items()}
    lib.motion_queries_fold(blob, d["temp_queries"], d["temp_ref_points"], d["temp_vel"], d["temp_timestamp"], d["temp_ego_pose"],
                            d["ego_pose_inv"], sel0.scale, C, q_out, A, c)
    for j, p in enumerate((p0, p1)):
        ref = O.motion_aware_queries(p, 0, inp["temp_queries"], inp["temp_ref_points"], inp["temp_vel"], inp["temp_timestamp"],
                                     inp["temp_ego_pose"], inp["ego_pose_inv"])
        err = (q_out[j].cpu() - ref).abs().max().item() / ref.abs().max().item()
        print("motion queries stage %d: max-abs %.3g rel %.3g (|ref| max %.2f)" % (j, (q_out[j].cpu() - ref).abs().max().item(), err, ref.abs().max().item()))
        assert err < 1e-5, err
        pre = "score_predictor.0."
        P = p[pre + "aggregate.0.weight"].double() @ q_out[j].cpu().double()                           # [Bf, 2, 256]
        A_ref = sel0.scale * (P @ p[pre + "input_proj.0.weight"].double())
        c_ref = sel0.scale * (P @ p[pre + "input_proj.0.bias"].double()) + p[pre + "aggregate.0.bias"].double()
        assert (A[j].cpu().double() - A_ref).abs().max().item() < 1e-5 * max(1.0, A_ref.abs().max().item())
        assert (c[j].cpu().double() - c_ref).abs().max().item() < 1e-5 * max(1.0, c_ref.abs().max().item())


def test_scorer_fold_and_tokens(lib):
    """Folded query scorer end to end (rows a12 + a13): toc3d_motion_queries_fold + toc3d_score_tokens against the oracle's
    motion_aware_queries + query_based_score (toc3d_utils.py:232-252, 334-360), fp32, 2e-4 on the log-probabilities."""
    g = torch.Generator().manual_seed(5)
    Bf, views, H, W, C, Q = 2, 3, 10, 22, 1024, 64
    V, N = Bf * views, H * W
    sel, p, inp, blob = _selector_case(43, Bf, Q, C)
    x = torch.randn(V, H, W, C, generator=g) * 2
    mask = torch.rand(V, H, W, 1, generator=g)
    gn = -torch.log(-torch.log(torch.rand(V, N, 2, generator=g).clamp(1e-9, 1 - 1e-7)))
    queries = O.motion_aware_queries(p, 0, inp["temp_queries"], inp["temp_ref_points"], inp["temp_vel"], inp["temp_timestamp"],
                                     inp["temp_ego_pose"], inp["ego_pose_inv"])
    pred_ref = O.query_based_score(x, mask, queries, p, 0)
    mask_ref = O.gumbel_mask(pred_ref, gn)[..., 0]
    q_out = torch.empty(1, Bf, Q, 256, device=DEV); A = torch.empty(1, Bf, 2, C, device=DEV); c = torch.empty(1, Bf, 2, device=DEV)
    d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in inp.items()}
    lib.motion_queries_fold(blob[None].to(DEV), d["temp_queries"], d["temp_ref_points"], d["temp_vel"], d["temp_timestamp"],
                            d["temp_ego_pose"], d["ego_pose_inv"], sel.scale, C, q_out, A, c)
    A, c = A[0].contiguous(), c[0].contiguous()
    pred = torch.empty(V, N, 2, device=DEV); score = torch.empty(V, N, device=DEV); mo = torch.empty(V, N, device=DEV)
    lib.score_tokens(x.to(DEV), mask.reshape(V, N).to(DEV).contiguous(), A, c, V, N, C, views, gn.to(DEV), 0, pred, score, mo)
    assert (pred.cpu() - pred_ref).abs().max().item() < 2e-4       # fp32, re-associated (folded) product
    assert torch.equal(score.cpu(), pred.cpu()[..., 0])
    assert (mo.cpu() - mask_ref).abs().max().item() < 2e-4
    # device-drawn noise: masks must lie in (0,1) and differ between seeds
    m1 = torch.empty(V, N, device=DEV); m2 = torch.empty(V, N, device=DEV)
    lib.score_tokens(x.to(DEV), None, A, c, V, N, C, views, None, 1, None, None, m1)
    lib.score_tokens(x.to(DEV), None, A, c, V, N, C, views, None, 2, None, None, m2)
    assert ((m1 > 0) & (m1 < 1)).all() and not torch.equal(m1, m2)


def test_mask_rows(lib):
    """x * mask (toc3d_utils.py:117,234: the previous stage's soft mask multiplies the scorer input), fp32, exact."""
    g = torch.Generator().manual_seed(21)
    M, C = 6000, 1024
    x = torch.randn(M, C, generator=g) * 3
    m = torch.rand(M, generator=g)
    out = torch.full((M, C), float("nan"), device=DEV)
    lib.mask_rows(x.to(DEV), m.to(DEV), out, M, C)
    assert torch.equal(out.cpu(), x * m[:, None])


@pytest.mark.parametrize("V,N,C", [(1, 77, 256), (2, 1000, 1024), (1, 5000, 1024)])
def test_global_half_mean(lib, V, N, C):
    """cat[y[:, :, :C/2], mean_N(y[:, :, C/2:])] (toc3d_utils.py:122-126) in place on the bf16 activations: the lower half
    is untouched bit for bit, the upper half holds the token mean (fp32 accumulation, one bf16 rounding)."""
    g = torch.Generator().manual_seed(V + N + C)
    y = bf16_round(torch.randn(V, N, C, generator=g) + 0.3)
    d = y.to(DEV).bfloat16().contiguous()
    lib.global_half_mean(d, V, N, C)
    got = d.float().cpu()
    assert torch.equal(got[:, :, : C // 2], y[:, :, : C // 2])
    mean = y[:, :, C // 2:].double().mean(dim=1, keepdim=True).float()
    assert (got[:, :, C // 2:] - mean).abs().max().item() <= 2.0 ** -8 * mean.abs().max().item() + 1e-6
    assert (got[:, :, C // 2:] == got[:, :1, C // 2:]).all()          # one value per (view, channel)


def test_score_finish(lib):
    """LogSoftmax over the 2 logits + pinned Gumbel mask (toc3d_utils.py:112,147): fp32, against the oracle."""
    g = torch.Generator().manual_seed(23)
    M = 6000
    logits = torch.randn(M, 2, generator=g) * 4
    logits[::97] = torch.tensor([60.0, -60.0]); logits[1::97] = torch.tensor([-80.0, 75.0])     # saturated rows
    gn = -torch.log(-torch.log(torch.rand(M, 2, generator=g).clamp(1e-9, 1 - 1e-7)))
    pred_ref = torch.log_softmax(logits, -1)
    mask_ref = O.gumbel_mask(pred_ref, gn)[..., 0]
    pred = torch.empty(M, 2, device=DEV); score = torch.empty(M, device=DEV); mo = torch.empty(M, device=DEV)
    lib.score_finish(logits.to(DEV), M, gn.to(DEV), 0, pred, score, mo)
    assert (pred.cpu() - pred_ref).abs().max().item() < 1e-5
    assert torch.equal(score.cpu(), pred.cpu()[:, 0])
    assert (mo.cpu() - mask_ref).abs().max().item() < 1e-5
    m1 = torch.empty(M, device=DEV); m2 = torch.empty(M, device=DEV)       # device-drawn noise
    lib.score_finish(logits.to(DEV), M, None, 1, None, None, m1)
    lib.score_finish(logits.to(DEV), M, None, 2, None, None, m2)
    assert ((m1 >= 0) & (m1 <= 1)).all() and not torch.equal(m1, m2)


def test_first_frame_scorer_full_width(lib):
    """The first-frame scorer (ScoreBasedTokenSelector.score, toc3d_utils.py:114-129) at EVA-ViT-L width through the
    engine's launch sequence (mask_rows, LayerNorm 1e-5, 1024x1024 GELU GEMM, global_half_mean, 1024-512-256-2 GEMMs,
    score_finish) against the oracle on identical fp32 input.  bf16 GEMM operands: log-prob tolerance 2e-2."""
    from toc3d_b200 import CONFIGS, ToC3DEVAViT
    from toc3d_b200.synthetic import randomize_state_dict
    from toc3d_b200.backbone import _Engine
    kind, cfg, hw = CONFIGS["toc3d_fast"]
    torch.manual_seed(0)
    model = ToC3DEVAViT(**dict(cfg, depth=6, global_attn_indexes=(2, 5), pruning_loc=[3], token_ratio=[0.7])).eval()
    sd = randomize_state_dict(model.state_dict(), seed=31, bias_std=0.1)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(32)
    V, H, W, C = 2, 20, 50, 1024
    x = torch.randn(V, H, W, C, generator=g) * 2.5
    mask = torch.rand(V, H, W, 1, generator=g)
    gn = -torch.log(-torch.log(torch.rand(V, H * W, 2, generator=g).clamp(1e-9, 1 - 1e-7)))
    pred_ref = O.first_frame_score(x, mask, sd, 0)
    lib.load()
    eng = _Engine(model, torch.device("cuda", 0))
    wsp = eng.workspace(V, H, W)
    for mk in (None, mask):
        ref = pred_ref if mk is not None else O.first_frame_score(x, torch.ones_like(mask), sd, 0)
        pred, score, m_out = eng.score_stage(0, x.reshape(V * H * W, C).to(DEV), None if mk is None else mk.reshape(-1).to(DEV),
                                             wsp, None, gn.to(DEV))
        d = (pred.cpu() - ref).abs().max().item()
        print("first-frame scorer full width max-abs log-prob diff %.5f (|ref| max %.2f)" % (d, ref.abs().max().item()))
        assert d < 2e-2, d
        assert torch.equal(score.cpu().reshape(-1), pred.cpu()[..., 0].reshape(-1))
        assert (m_out.cpu().reshape(-1) - O.gumbel_mask(ref, gn)[..., 0].reshape(-1)).abs().max().item() < 1e-2


def test_im2col_patch_embed(lib):
    g = torch.Generator().manual_seed(9)
    V, Hi, Wi, C = 2, 64, 96, 256
    img = torch.randn(V, 3, Hi, Wi, generator=g)
    p = {"patch_embed.proj.weight": bf16_round(torch.randn(C, 3, 16, 16, generator=g) * 0.05),
         "patch_embed.proj.bias": torch.randn(C, generator=g)}
    ref = O.patch_embed(bf16_round(img), p, 16).reshape(-1, C)
    cols = torch.empty(V * (Hi // 16) * (Wi // 16), 768, device=DEV, dtype=torch.bfloat16)
    lib.im2col_patch16(img.to(DEV), cols, V, Hi, Wi)
    out = torch.empty(cols.shape[0], C, device=DEV)
    lib.gemm(cols, p["patch_embed.proj.weight"].reshape(C, 768).to(DEV).bfloat16(), lib.EPI_LINEAR,
             bias=p["patch_embed.proj.bias"].to(DEV), out=out, out_f32=True)
    assert (out.cpu() - ref).abs().max().item() < 1e-3


def test_bad_arguments_fail_loudly(lib):
    A = torch.zeros(8, 60, device=DEV, dtype=torch.bfloat16)
    B = torch.zeros(8, 60, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(8, 8, device=DEV)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        lib.gemm(A, B, lib.EPI_LINEAR, out=out, out_f32=True)
    with pytest.raises(RuntimeError, match="multiple of the 16x16 patch"):
        lib.im2col_patch16(torch.zeros(1, 3, 30, 32, device=DEV), torch.zeros(4, 768, device=DEV, dtype=torch.bfloat16), 1, 30, 32)
