"""CPU oracle for the ToC3D image-backbone hot path.  TEST INFRASTRUCTURE ONLY.

A plain fp32 PyTorch restatement (functional, state-dict driven) of the
reference algorithm, used as the checker by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  The product package
(toc3d_b200/) never imports this file.

Parity status: the reference ships NO tests or golden vectors for this path
(SURVEY.md §4), so "parity unpinned by the reference's own tests".  Instead
the oracle is pinned against outputs of the reference itself: tests/golden/
holds vectors produced by importing the unmodified reference in the build
container (tests/golden/make_golden.py), and tests/test_oracle_cpu.py /
tests/test_neck.py re-check the oracle against the live reference whenever
/root/reference is present.

Two pins (choices where the reference leaves behaviour unspecified):
  pin 1  sort tie-break = (score descending, index ascending)  [torch.sort stable=True]
  pin 2  the eval-time gumbel_softmax mask consumes injected noise g:
         mask = softmax(logp + g)[..., 0]   (g = -log(Exp(1)) in the reference)

Reference files restated (paths under projects/mmdet3d_plugin/models/):
  backbones/toc3d_eva_vit.py   ToC3DEVAViT.forward :230-310, ToC3DEVAViTBlock :329-473,
                               ToC3DEVAAttention :480-518
  backbones/toc3d_utils.py     batch_index_select/fill :28-62, merge_tokens :65-70,
                               ScoreBasedTokenSelector :92-158, NaiveQuery... :196-274,
                               MotionAware... :294-422
  backbones/eva_vit.py         SwiGLU :27-51, Attention :54-119, Block :200-268, EVA_ViT :409-428
  backbones/eva_utils.py       window_partition/unpartition :89-133, get_abs_pos :229-258,
                               PatchEmbed :261-287, VisionRotaryEmbeddingFast(+WithSelection) :325-402
  necks/cp_fpn.py              CPFPN.forward :157-208 (shipped config: in_channels=[1024], out 256, num_outs=2)
  datasets/pipelines/transform_3d.py  (under projects/mmdet3d_plugin/) PadMultiViewImage :21-68,
                               NormalizeMultiviewImage :72-104 -> mmcv.imnormalize / impad_to_multiple.  mmcv is a
                               dependency ABSENT from /root/reference and this image (README.md:52 pins mmcv-full
                               1.6.0): normalize_pad_images restates its published algorithm and is pinned against the
                               same OpenCV calls (cv2 is installed; tests/test_preprocess.py, tests/golden/preprocess_cv2.pt)
  utils/misc.py                MLN :154-188, transform_reference_points :191-200
  utils/positional_encoding.py pos2posemb3d :14-26, pos2posemb1d :28-37, nerf_positional_encoding :39-81
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-6          # norm_layer=partial(nn.LayerNorm, eps=1e-6), toc3d_eva_vit.py:38
PAD_SCORE = -1e6       # toc3d_eva_vit.py:415


# --------------------------------------------------------------------------- layout
def window_partition(x, ws, pad_value=0.0):
    """eva_utils.py:89-110.  (B,H,W,C) -> (B*nWh*nWw, ws, ws, C), (Hp, Wp)."""
    B, H, W, C = x.shape
    ph, pw = (-H) % ws, (-W) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph), value=pad_value)
    Hp, Wp = H + ph, W + pw
    x = x.reshape(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(w, ws, pad_hw, hw):
    """eva_utils.py:113-133."""
    Hp, Wp = pad_hw
    H, W = hw
    B = w.shape[0] // ((Hp // ws) * (Wp // ws))
    x = w.reshape(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :].contiguous()


# --------------------------------------------------------------------------- RoPE
def rope_table(ft_seq_len, half_head_dim=32, pt_seq_len=16, theta=10000.0):
    """eva_utils.py:325-371: cos/sin tables (ft*ft, 2*half_head_dim), fp32.

    Column layout per row (r, c) of the ft x ft grid: the first half_head_dim
    columns carry the row-coordinate angles, the last half_head_dim the
    column-coordinate angles; each of the half_head_dim/2 frequencies appears
    twice in adjacent columns (pair rotation).
    """
    d = half_head_dim
    freqs = 1.0 / (theta ** (torch.arange(0, d, 2)[: d // 2].float() / d))
    t = torch.arange(ft_seq_len) / ft_seq_len * pt_seq_len
    ang = (t[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)        # (ft, d)
    full = torch.cat([ang[:, None, :].expand(ft_seq_len, ft_seq_len, d),
                      ang[None, :, :].expand(ft_seq_len, ft_seq_len, d)], dim=-1)
    full = full.reshape(ft_seq_len * ft_seq_len, 2 * d)
    return full.cos(), full.sin()


def rotate_pairs(x):
    """eva_utils.py:318-322 rotate_half: (x0,x1,x2,x3,..) -> (-x1,x0,-x3,x2,..)."""
    a, b = x[..., 0::2], x[..., 1::2]
    return torch.stack((-b, a), dim=-1).flatten(-2)


def apply_rope(t, cos, sin):
    return t * cos + rotate_pairs(t) * sin


# --------------------------------------------------------------------------- ViT pieces
def _lin(x, p, name, bias=True):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"] if bias else None)


def attention(x, p, pre, heads, cos, sin):
    """eva_vit.py:86-119 / toc3d_eva_vit.py:484-518.  x (B,N,C); cos/sin (N,64) or (B,1,N,64)."""
    B, N, C = x.shape
    q = F.linear(x, p[pre + "q_proj.weight"], p.get(pre + "q_bias"))
    k = F.linear(x, p[pre + "k_proj.weight"], None)
    v = F.linear(x, p[pre + "v_proj.weight"], p.get(pre + "v_bias"))
    q, k, v = (t.reshape(B, N, heads, -1).permute(0, 2, 1, 3) for t in (q, k, v))
    if cos is not None:
        q = apply_rope(q, cos, sin)
        k = apply_rope(k, cos, sin)
    q = q * (q.shape[-1] ** -0.5)
    a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, -1)
    return _lin(o, p, pre + "proj")


def swiglu(x, p, pre):
    """eva_vit.py:44-51 with subln=True."""
    h = F.silu(_lin(x, p, pre + "w1")) * _lin(x, p, pre + "w2")
    h = F.layer_norm(h, (h.shape[-1],), p[pre + "ffn_ln.weight"], p[pre + "ffn_ln.bias"], LN_EPS)
    return _lin(h, p, pre + "w3")


def _ln(x, p, name, eps=LN_EPS):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


def dense_block(x, p, i, ws, heads, rope):
    """eva_vit.py:247-268: pads are added AFTER norm1 (zero rows)."""
    pre = "blocks.%d." % i
    B, H, W, C = x.shape
    y = _ln(x, p, pre + "norm1")
    yw, pad_hw = window_partition(y, ws)
    cos, sin = rope if rope is not None else (None, None)
    a = attention(yw.reshape(-1, ws * ws, C), p, pre + "attn.", heads, cos, sin)
    a = window_unpartition(a.reshape(-1, ws, ws, C), ws, pad_hw, (H, W))
    x = x + a
    return x + swiglu(_ln(x, p, pre + "norm2"), p, pre + "mlp.")


# --------------------------------------------------------------------------- selection
def stable_sort_desc(score):
    """pin 1.  score (B,n) -> (sorted_score, sorted_idx)."""
    return torch.sort(score, dim=1, descending=True, stable=True)


def sample(score, ratio):
    """toc3d_utils.py:131-144 (index part).  score (B,n)."""
    k = int(score.shape[1] * ratio)
    s, idx = stable_sort_desc(score)
    return s[:, :k], s[:, k:], idx[:, :k], idx[:, k:]


def batch_index_select(x, idx):
    """toc3d_utils.py:28-44."""
    return torch.gather(x, 1, idx[..., None].expand(-1, -1, x.shape[-1])) if x.dim() == 3 \
        else torch.gather(x, 1, idx)


def merge_tokens(x_drop, score):
    """toc3d_utils.py:65-70."""
    w = score / score.sum(dim=1, keepdim=True)
    return (w[..., None] * x_drop).sum(dim=1, keepdim=True)


def toc3d_block(x, scores, p, i, ws, ratio, heads, rope, forced=None, tap=None):
    """toc3d_eva_vit.py:395-473 (accelerated branch), use_represent_tokens=True.

    forced: optional (slow_idx, fast_idx) per-window index tensors that replace
    the sort (teacher forcing for kernel/block parity tests).
    """
    pre = "blocks.%d." % i
    B, H, W, C = x.shape
    n = ws * ws
    xw, pad_hw = window_partition(x, ws)                                   # pads BEFORE norm1
    sw, _ = window_partition(scores[..., None], ws, pad_value=PAD_SCORE)
    xw = xw.reshape(-1, n, C)
    sw = sw.reshape(-1, n)
    if forced is None:
        _, fast_s, slow_idx, fast_idx = sample(sw, ratio)
    else:
        slow_idx, fast_idx = forced
        fast_s = torch.gather(sw, 1, fast_idx)
    k = slow_idx.shape[1]
    slow = batch_index_select(xw, slow_idx)
    fast = batch_index_select(xw, fast_idx)
    if fast.shape[1] == 0:
        raise NotImplementedError("ratio=1.0 is a latent bug in the reference (toc3d_eva_vit.py:462-463)")
    rep = merge_tokens(fast, fast_s)
    t = torch.cat([slow, rep], dim=1)
    cos = sin = None
    if rope is not None:
        rows = torch.cat([slow_idx, torch.full_like(slow_idx[:, :1], k)], dim=1)   # rep token -> table row k
        cos = rope[0][rows][:, None]                                               # (nW,1,k+1,64)
        sin = rope[1][rows][:, None]
    raw1 = attention(_ln(t, p, pre + "norm1"), p, pre + "attn.", heads, cos, sin)
    t1 = t + raw1
    raw2 = swiglu(_ln(t1, p, pre + "norm2"), p, pre + "mlp.")
    t2 = t1 + raw2
    slow_o = t2[:, :k]
    fast_o = fast + raw1[:, k:k + 1] + raw2[:, k:k + 1]
    out = torch.zeros_like(xw)
    out.scatter_(1, slow_idx[..., None].expand(-1, -1, C), slow_o)
    out.scatter_(1, fast_idx[..., None].expand(-1, -1, C), fast_o)
    if tap is not None:
        tap.update(slow_idx=slow_idx, fast_idx=fast_idx, rep=rep, t=t, raw1=raw1, raw2=raw2, t2=t2)
    return window_unpartition(out.reshape(-1, ws, ws, C), ws, pad_hw, (H, W))


# --------------------------------------------------------------------------- query encoder
def pos2posemb(pos, num_pos_feats, temperature=10000):
    """positional_encoding.py:14-37 per-coordinate sin/cos embedding; pos (..., D) -> list of (..., F)."""
    pos = pos * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    out = []
    for c in range(pos.shape[-1]):
        a = pos[..., c, None] / dim_t
        out.append(torch.stack((a[..., 0::2].sin(), a[..., 1::2].cos()), dim=-1).flatten(-2))
    return out


def pos2posemb3d(pos):
    ex, ey, ez = pos2posemb(pos, 128)
    return torch.cat((ey, ex, ez), dim=-1)        # y first: positional_encoding.py:25


def pos2posemb1d(pos):
    return pos2posemb(pos[..., :1], 256)[0]


def nerf_encoding(t, n_fn=6):
    """positional_encoding.py:39-81 (include_input=False, log_sampling=True)."""
    bands = 2.0 ** torch.linspace(0.0, n_fn - 1, n_fn, dtype=t.dtype)
    enc = []
    for f in bands:
        enc += [torch.sin(t * f), torch.cos(t * f)]
    return torch.cat(enc, dim=-1)


def mln(x, c, p, pre):
    """misc.py:154-188."""
    x = F.layer_norm(x, (x.shape[-1],))
    c = F.relu(_lin(c, p, pre + "reduce.0"))
    return _lin(c, p, pre + "gamma") * x + _lin(c, p, pre + "beta")


def motion_aware_queries(p, j, temp_queries, temp_ref_points, temp_vel, temp_timestamp,
                         temp_ego_pose, ego_pose_inv):
    """toc3d_utils.py:334-360."""
    pre = "score_predictor.%d." % j
    assert ego_pose_inv is not None
    ref = torch.cat([temp_ref_points, torch.ones_like(temp_ref_points[..., :1])], dim=-1)
    ref = (ego_pose_inv.unsqueeze(1) @ ref.unsqueeze(-1)).squeeze(-1)[..., :3]      # misc.py:191-200
    pc = p[pre + "pc_range"]
    ref = (ref - pc[:3]) / (pc[3:6] - pc[0:3])
    pos = _lin(F.relu(_lin(pos2posemb3d(ref), p, pre + "query_embedding.0")), p, pre + "query_embedding.2")
    motion = torch.cat([temp_vel, temp_timestamp, temp_ego_pose[..., :3, :].flatten(-2)], dim=-1).float()
    motion = nerf_encoding(motion)
    pos = mln(pos, motion, p, pre + "ego_pose_pe.")
    te = _lin(pos2posemb1d(temp_timestamp).float(), p, pre + "time_embedding.0")
    pos = pos + F.layer_norm(te, (te.shape[-1],), p[pre + "time_embedding.1.weight"],
                             p[pre + "time_embedding.1.bias"])
    return mln(temp_queries, motion, p, pre + "ego_pose_queries.") + pos


def query_based_score(x, mask, queries, p, j):
    """toc3d_utils.py:232-252 (score_type='attention', attn_scale=True).  -> (V,N,2) log-probs."""
    pre = "score_predictor.%d." % j
    V = x.shape[0]
    xin = (x * mask).flatten(1, 2)
    xin = _lin(xin, p, pre + "input_proj.0")
    q = queries.repeat_interleave(V // queries.shape[0], dim=0)
    att = torch.einsum("bnc,bqc->bnq", xin, q) * (q.shape[-1] ** -0.5)
    return F.log_softmax(_lin(att, p, pre + "aggregate.0"), dim=-1)


def first_frame_score(x, mask, p, j):
    """toc3d_utils.py:114-129 (used in eval when prev_exists is False, :270-271)."""
    pre = "score_predictor.%d." % j
    V, H, W, C = x.shape
    xin = (x * mask).reshape(V, H * W, C)
    y = F.gelu(_lin(F.layer_norm(xin, (C,), p[pre + "in_conv.0.weight"], p[pre + "in_conv.0.bias"]),
                    p, pre + "in_conv.1"))
    g = y[:, :, C // 2:].mean(dim=1, keepdim=True)
    y = torch.cat([y[:, :, : C // 2], g.expand(V, H * W, C // 2)], dim=2)
    y = F.gelu(_lin(y, p, pre + "out_conv.0"))
    y = F.gelu(_lin(y, p, pre + "out_conv.2"))
    return F.log_softmax(_lin(y, p, pre + "out_conv.4"), dim=-1)


def gumbel_mask(pred_score, g):
    """pin 2: F.gumbel_softmax(pred_score, hard=False)[..., 0:1] with injected g (toc3d_utils.py:147)."""
    return (pred_score + g).softmax(dim=-1)[..., 0:1]


# --------------------------------------------------------------------------- stem
def abs_pos(pos_embed, hw, has_cls=True):
    """eva_utils.py:229-258."""
    h, w = hw
    a = pos_embed[:, 1:] if has_cls else pos_embed
    size = int(math.sqrt(a.shape[1]))
    assert size * size == a.shape[1]
    if size != h or size != w:
        a = F.interpolate(a.reshape(1, size, size, -1).permute(0, 3, 1, 2), size=(h, w),
                          mode="bicubic", align_corners=False)
        return a.permute(0, 2, 3, 1)
    return a.reshape(1, h, w, -1)


def patch_embed(img, p, patch):
    """eva_utils.py:283-287."""
    y = F.conv2d(img, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=patch)
    return y.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------- full forwards
def _cfg(cfg):
    c = dict(img_size=320, patch_size=16, window_size=16, global_window_size=20, embed_dim=1024,
             depth=24, num_heads=16, global_attn_indexes=(2, 5, 8, 11, 14, 17, 20, 23),
             pruning_loc=(6, 12, 18), token_ratio=(0.7, 0.5, 0.5), pt_hw_seq_len=16,
             rope=True, rope_acc=True, accelerate_global=True)
    c.update(cfg or {})
    return c


def _ropes(c):
    hh = c["embed_dim"] // c["num_heads"] // 2
    return (rope_table(c["window_size"], hh, c["pt_hw_seq_len"]),
            rope_table(c["img_size"] // c["patch_size"], hh, c["pt_hw_seq_len"]))


def forward_dense(p, cfg, img, tap=None):
    """EVA_ViT.forward eva_vit.py:409-428 -> {'last_feat': (V,C,H,W)}."""
    c = _cfg(cfg)
    x = patch_embed(img.float(), p, c["patch_size"])
    if "pos_embed" in p:
        x = x + abs_pos(p["pos_embed"], x.shape[1:3])
    win, glb = _ropes(c)
    for i in range(c["depth"]):
        g = i in c["global_attn_indexes"]
        if tap is not None:
            tap.setdefault("block_in", []).append(x)
        x = dense_block(x, p, i, c["global_window_size"] if g else c["window_size"], c["num_heads"],
                        glb if g else win)
        if tap is not None:
            tap.setdefault("block_out", []).append(x)
    return {"last_feat": x.permute(0, 3, 1, 2)}


def forward_toc3d(p, cfg, img, temp_queries, temp_ref_points, temp_vel, temp_timestamp, temp_ego_pose,
                  ego_pose_inv, prev_exists=True, gumbel_noise=None, tap=None, forced_scores=None):
    """ToC3DEVAViT.forward toc3d_eva_vit.py:230-310 in eval mode.

    gumbel_noise: list of 3 tensors (V,N,2) (pin 2).  Returns a dict with
    last_feat (V,C,H,W), token_masks [3x(V,H,W,1)], keep_idx, drop_idx, scores [3x(V,H,W)].
    forced_scores (test hook, mirrors the plugin's teacher_scores): per-stage (V,H,W) scores that replace the
    predicted ones for the sort / window selection (the masks still come from the predicted scores).
    """
    c = _cfg(cfg)
    assert not set(c["pruning_loc"]) & set(c["global_attn_indexes"])          # toc3d_eva_vit.py:141
    x = patch_embed(img.float(), p, c["patch_size"])
    if "pos_embed" in p:
        x = x + abs_pos(p["pos_embed"], x.shape[1:3])
    V, H, W, C = x.shape
    win, glb = _ropes(c)
    masks = torch.ones(V, H, W, 1)
    scores = None
    stage = -1
    out = dict(token_masks=[], keep_idx=[], drop_idx=[], scores=[])
    if tap is not None:
        tap["block_in"] = []
        tap["block_out"] = []
        tap["stem"] = x
    for i in range(c["depth"]):
        if i in c["pruning_loc"]:
            stage += 1
            if prev_exists:
                q = motion_aware_queries(p, stage, temp_queries, temp_ref_points, temp_vel, temp_timestamp,
                                         temp_ego_pose, ego_pose_inv)
                pred = query_based_score(x, masks, q, p, stage)
            else:
                pred = first_frame_score(x, masks, p, stage)
            sel = pred[:, :, 0] if forced_scores is None else forced_scores[stage].reshape(V, H * W).float()
            _, _, keep_idx, drop_idx = sample(sel, c["token_ratio"][stage])
            masks = gumbel_mask(pred, gumbel_noise[stage].reshape(pred.shape)).reshape(V, H, W, 1)
            scores = sel.reshape(V, H, W)
            out["token_masks"].append(masks)
            out["keep_idx"].append(keep_idx)
            out["drop_idx"].append(drop_idx)
            out["scores"].append(scores)
        g = i in c["global_attn_indexes"]
        ws = c["global_window_size"] if g else c["window_size"]
        accel = i >= c["pruning_loc"][0] and (c["accelerate_global"] or not g)   # toc3d_eva_vit.py:178-180
        if tap is not None:
            tap["block_in"].append(x)
        if accel:
            rope = None
            if c["rope"] and c["rope_acc"]:
                rope = glb if g else win
            x = toc3d_block(x, scores, p, i, ws, c["token_ratio"][stage], c["num_heads"], rope)
        else:
            x = dense_block(x, p, i, ws, c["num_heads"], (glb if g else win) if c["rope"] else None)
        if tap is not None:
            tap["block_out"].append(x)
    out["last_feat"] = x.permute(0, 3, 1, 2)
    return out


# --------------------------------------------------------------------------- neck
def neck_cpfpn(p, last_feat, prefix=""):
    """necks/cp_fpn.py:157-208 with the shipped config (ToC3D_fast.py:70-74: in_channels=[1024],
    out_channels=256, num_outs=2, no norm / activation / extra convs): lateral 1x1 conv (:114-122, :163-166),
    3x3 conv pad 1 on level 0 (:123-133, :182-184), then max_pool2d(kernel 1, stride 2) = subsample (:190-191).
    last_feat (V,C,H,W) -> [(V,256,H,W), (V,256,ceil(H/2),ceil(W/2))]."""
    lat = F.conv2d(last_feat, p[prefix + "lateral_convs.0.conv.weight"], p[prefix + "lateral_convs.0.conv.bias"])
    out0 = F.conv2d(lat, p[prefix + "fpn_convs.0.conv.weight"], p[prefix + "fpn_convs.0.conv.bias"], padding=1)
    return [out0, F.max_pool2d(out0, 1, stride=2)]


def normalize_pad_images(imgs_u8, mean, std, to_rgb=True, size_divisor=32):
    """NormalizeMultiviewImage + PadMultiViewImage (transform_3d.py:38-48,95-96) + the NCHW stacking of the format
    bundle, on uint8 HWC camera crops (V, Hs, Ws, 3) -> fp32 (V, 3, Hi, Wi).

    mmcv.imnormalize_ (mmcv/image/photometric.py): mean = float64(mean), stdinv = 1/float64(std), optional BGR->RGB,
    then cv2.subtract(img, mean, img); cv2.multiply(img, stdinv, img) on the float32 image.  OpenCV evaluates both in
    double per element and stores float32 after each call.  mmcv.impad_to_multiple pads bottom / right with 0."""
    import numpy as np
    img = imgs_u8.numpy().astype(np.float32)                     # LoadMultiViewImageFromFiles(to_float32=True)
    if to_rgb:
        img = img[..., ::-1]
    mean64 = np.asarray(mean, dtype=np.float32).astype(np.float64).reshape(1, 1, 1, 3)
    stdinv = 1.0 / np.asarray(std, dtype=np.float32).astype(np.float64).reshape(1, 1, 1, 3)
    img = (img.astype(np.float64) - mean64).astype(np.float32)
    img = (img.astype(np.float64) * stdinv).astype(np.float32)
    V, Hs, Ws, _ = img.shape
    Hi, Wi = -(-Hs // size_divisor) * size_divisor, -(-Ws // size_divisor) * size_divisor
    out = np.zeros((V, Hi, Wi, 3), dtype=np.float32)
    out[:, :Hs, :Ws] = img
    return torch.from_numpy(out).permute(0, 3, 1, 2).contiguous()
