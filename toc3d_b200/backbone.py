"""Drop-in `img_backbone` plugins: ToC3DEVAViT and EVA_ViT on hand-written sm_100a kernels.

Host-side mirror of the reference plugin interface (same registry names, constructor
kwargs, keyword `forward` signature, return type and state-dict keys):
  projects/mmdet3d_plugin/models/backbones/toc3d_eva_vit.py:25-310  (ToC3DEVAViT)
  projects/mmdet3d_plugin/models/backbones/eva_vit.py:270-428       (EVA_ViT)
  caller: projects/mmdet3d_plugin/models/detectors/petr3d.py:145-179

nn.Module is used as the parameter container (so reference checkpoints load with
`load_state_dict`); all arithmetic on the path runs in libtoc3d_b200.so through
toc3d_b200.lib (ctypes, raw device pointers).  Inference only.  There is no CPU or
PyTorch fallback: CPU inputs or a missing library raise.
"""
import contextlib
import math
import os
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as L
from .preprocess import ImagePreprocess

LN_EPS = 1e-6

try:  # register with mmdet when it is importable (reference: toc3d_eva_vit.py:25, eva_vit.py:270)
    from mmdet.models.builder import BACKBONES as _BACKBONES

    def _register(cls):
        return _BACKBONES.register_module(force=True)(cls)
except Exception:  # mmdet absent: plain nn.Module
    def _register(cls):
        return cls

try:  # inside the reference tree Petr3D isinstance-checks this exact class (petr3d.py:17,159)
    from projects.mmdet3d_plugin.models.backbones.toc3d_utils import ToC3DViTReturnType
except Exception:
    class ToC3DViTReturnType:
        """toc3d_utils.py:10-25."""

        def __init__(self, img_feats=None, token_masks=None, attn_scores=None, keep_idx=None, drop_idx=None,
                     aux_outputs=None):
            self.img_feats = img_feats
            self.token_masks = token_masks
            self.attn_scores = attn_scores
            self.keep_idx = keep_idx
            self.drop_idx = drop_idx
            self.aux_outputs = aux_outputs

try:
    from projects.mmdet3d_plugin.models.utils.gpu_timer import GLOBAL_TIMER
except Exception:
    class _NoTimer:
        def event_start(self, name):
            pass

        def event_end(self, name):
            pass
    GLOBAL_TIMER = _NoTimer()


# ------------------------------------------------------------------------------- weight repacking
def balanced_item_order(q_rows, heads):
    """(window, head) items of the persistent attention kernel, heaviest (most 128-row query tiles) first, so that
    dealing them round-robin to the CTAs balances the tile units.  q_rows: int tensor [nW] on the host."""
    tiles = (q_rows.to(torch.int64) + 127) // 128
    per_item = tiles.repeat_interleave(heads)
    return torch.sort(per_item, descending=True, stable=True).indices.to(torch.int32)


def hidden_pad(hd):
    """SwiGLU hidden width padded to the GEMM K block (2730 -> 2752): TMA needs 16-byte row strides."""
    return (hd + 63) // 64 * 64


def interleave_w12(w1, b1, w2, b2, hp):
    """[w1; w2] -> rows interleaved in blocks of 32 so one accumulator tile holds both SwiGLU factors.
    Row 64*b + j = w1[32*b + j], row 64*b + 32 + j = w2[32*b + j] (zero beyond the true width)."""
    hd, k = w1.shape
    W = w1.new_zeros(hp // 32, 2, 32, k)
    B = b1.new_zeros(hp // 32, 2, 32)
    w1p = F.pad(w1, (0, 0, 0, hp - hd)); w2p = F.pad(w2, (0, 0, 0, hp - hd))
    W[:, 0] = w1p.reshape(hp // 32, 32, k); W[:, 1] = w2p.reshape(hp // 32, 32, k)
    B[:, 0] = F.pad(b1, (0, hp - hd)).reshape(-1, 32); B[:, 1] = F.pad(b2, (0, hp - hd)).reshape(-1, 32)
    return W.reshape(2 * hp, k).contiguous(), B.reshape(2 * hp).contiguous()


def pack_motion_blob(sel, Q, C):
    """Per-stage parameter blob of toc3d_motion_queries_fold (layout: include/toc3d_b200.h).  Linear weights are
    transposed to [in][out]; the positional-encoding temperature tables are computed with torch exactly as the
    reference does (positional_encoding.py:17,31: temperature ** (2 * (i // 2) / F) in fp32)."""
    f = lambda t: t.detach().to(torch.float32).cpu().reshape(-1)
    wT = lambda lin: lin.weight.detach().to(torch.float32).cpu().t().contiguous().reshape(-1)

    def dim_t(n):
        d = torch.arange(n, dtype=torch.float32)
        return 10000 ** (2 * torch.div(d, 2, rounding_mode="floor") / n)

    def mln(m):
        return [wT(m.reduce[0]), f(m.reduce[0].bias), wT(m.gamma), f(m.gamma.bias), wT(m.beta), f(m.beta.bias)]
    parts = [dim_t(128), dim_t(256), F.pad(f(sel.pc_range), (0, 2)),
             wT(sel.query_embedding[0]), f(sel.query_embedding[0].bias), wT(sel.query_embedding[2]), f(sel.query_embedding[2].bias)]
    parts += mln(sel.ego_pose_pe) + mln(sel.ego_pose_queries)
    parts += [wT(sel.time_embedding[0]), f(sel.time_embedding[0].bias), f(sel.time_embedding[1].weight), f(sel.time_embedding[1].bias),
              f(sel.input_proj[0].weight), f(sel.input_proj[0].bias), f(sel.aggregate[0].weight), F.pad(f(sel.aggregate[0].bias), (0, 2))]
    blob = torch.cat(parts)
    assert blob.numel() == L.motion_blob_floats(Q, C), (blob.numel(), L.motion_blob_floats(Q, C))
    return blob


# ------------------------------------------------------------------------------- parameter containers
class _Rope(nn.Module):
    """eva_utils.py:325-371 buffers (freqs_cos/freqs_sin of shape (ft*ft, 2*dim))."""

    def __init__(self, dim, pt_seq_len, ft_seq_len, theta=10000.0):
        super().__init__()
        ft = ft_seq_len if ft_seq_len is not None else pt_seq_len
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        t = torch.arange(ft) / ft * pt_seq_len
        ang = (t[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)
        full = torch.cat([ang[:, None, :].expand(ft, ft, dim), ang[None, :, :].expand(ft, ft, dim)], -1)
        full = full.reshape(ft * ft, 2 * dim)
        self.ft = ft
        self.register_buffer("freqs_cos", full.cos())
        self.register_buffer("freqs_sin", full.sin())


class _Attention(nn.Module):
    def __init__(self, dim, heads, qkv_bias, rope):
        super().__init__()
        self.num_heads = heads
        self.q_proj = nn.Linear(dim, dim, bias=False)
        self.k_proj = nn.Linear(dim, dim, bias=False)
        self.v_proj = nn.Linear(dim, dim, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            self.q_bias = self.v_bias = None
        self.rope = rope
        self.proj = nn.Linear(dim, dim)


class _SwiGLU(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden)
        self.w2 = nn.Linear(dim, hidden)
        self.ffn_ln = nn.LayerNorm(hidden, eps=LN_EPS)
        self.w3 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, qkv_bias, window_size, rope, accelerate):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = _Attention(dim, heads, qkv_bias, rope)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = _SwiGLU(dim, int(dim * mlp_ratio))
        self.window_size = window_size
        self.accelerate = accelerate


class _PatchEmbed(nn.Module):
    def __init__(self, patch, in_chans, dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)


class _MLN(nn.Module):
    """Parameters of misc.py:154-188 (the arithmetic runs in toc3d_motion_queries_fold)."""

    def __init__(self, c_dim, f_dim=256):
        super().__init__()
        self.reduce = nn.Sequential(nn.Linear(c_dim, f_dim), nn.ReLU())
        self.gamma = nn.Linear(f_dim, f_dim)
        self.beta = nn.Linear(f_dim, f_dim)
        nn.init.zeros_(self.gamma.weight); nn.init.zeros_(self.beta.weight)
        nn.init.ones_(self.gamma.bias); nn.init.zeros_(self.beta.bias)


class _Selector(nn.Module):
    """Parameter container of MotionAwareQueryGuidedTokenSelector (toc3d_utils.py:92-112,196-230,294-332).  The per-frame
    query encoder (toc3d_utils.py:334-360) runs in toc3d_motion_queries_fold from a packed copy (pack_motion_blob)."""

    def __init__(self, embed_dim, num_queries, ratio, pc_range, query_dim=256):
        super().__init__()
        self.ratio = ratio
        self.num_queries = num_queries
        self.scale = query_dim ** -0.5
        self.in_conv = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, embed_dim), nn.GELU())
        self.out_conv = nn.Sequential(nn.Linear(embed_dim, embed_dim // 2), nn.GELU(),
                                      nn.Linear(embed_dim // 2, embed_dim // 4), nn.GELU(),
                                      nn.Linear(embed_dim // 4, 2), nn.LogSoftmax(dim=-1))
        self.input_proj = nn.Sequential(nn.Linear(embed_dim, query_dim))
        self.aggregate = nn.Sequential(nn.Linear(num_queries, 2), nn.LogSoftmax(dim=-1))
        self.pc_range = nn.Parameter(torch.tensor(pc_range, dtype=torch.float32), requires_grad=False)
        self.query_embedding = nn.Sequential(nn.Linear(query_dim * 3 // 2, query_dim), nn.ReLU(),
                                             nn.Linear(query_dim, query_dim))
        self.ego_pose_pe = _MLN(180)
        self.ego_pose_queries = _MLN(180)
        self.time_embedding = nn.Sequential(nn.Linear(query_dim, query_dim), nn.LayerNorm(query_dim))


# ------------------------------------------------------------------------------- device engine
class _Workspace:
    """Per-(V, H, W) device buffers and static window maps, sized for the largest block."""

    def __init__(self, eng, V, H, W):
        dev, C = eng.device, eng.C
        self.V, self.H, self.W, self.N = V, H, W, H * W
        rows = 0
        self.win = {}
        for ws in sorted(set(eng.block_ws)):
            nWh, nWw = -(-H // ws), -(-W // ws)
            nW, n = V * nWh * nWw, ws * ws
            idx = torch.arange(V * H * W, dtype=torch.float32).reshape(V, H, W)
            idx = F.pad(idx, (0, nWw * ws - W, 0, nWh * ws - H), value=-1.0)
            idx = idx.reshape(V, nWh, ws, nWw, ws).permute(0, 1, 3, 2, 4).reshape(-1)
            m0 = idx.to(torch.int32).view(nW, n)               # row-major window slot -> image row | -1 (pad)
            real0 = m0 >= 0
            # storage order of a window in the qkv buffer: real tokens first (row-major among themselves), pad slots
            # after them - attention is order-free over keys, RoPE uses the original slot position, and the query
            # rows that are needed become a prefix (query tiles of pure padding are skipped)
            order = torch.sort((~real0).to(torch.int8), dim=1, stable=True).indices        # [nW, n] original slot per storage slot
            m = torch.gather(m0, 1, order).reshape(-1)
            pos = order.to(torch.int32).reshape(-1)            # storage slot -> original position inside the window
            slots = torch.arange(nW * n, dtype=torch.int32)
            real = m >= 0
            inv = torch.empty(V * H * W, dtype=torch.int32)
            inv[m[real].long()] = slots[real]                  # image row -> storage slot (global index)
            rope_of_row = torch.empty(V * H * W, dtype=torch.int32)
            rope_of_row[m[real].long()] = pos[real]
            self.win[ws] = dict(nW=nW, n=n, map=m.to(dev),
                                # dense blocks run q/k/v only over real tokens and scatter them to their window slots
                                slot_of_row=inv.to(dev), rope_slot=rope_of_row.to(dev), pad_rows=slots[~real].to(dev),
                                real_per_window=real0.sum(1).to(torch.int32), q_rows=real0.sum(1).to(torch.int32).to(dev),
                                item_order=balanced_item_order(real0.sum(1), eng.heads).to(dev))
            rows = max(rows, nW * n)
        self.rows = rows
        bf = dict(device=dev, dtype=torch.bfloat16)
        self.cols = torch.empty(V * self.N, 768, **bf) if eng.patch == 16 else None
        self.a = torch.empty(rows, C, **bf)            # LN outputs / attention outputs (GEMM A operands)
        self.qkv = torch.zeros(rows, 3 * C, **bf)      # zeroed once: q of pad slots is never written, must stay finite
        self.ao = torch.empty(rows, C, **bf)
        self.hid = torch.empty(rows, eng.Hp, **bf)
        # compact slow + rep residual rows of an accelerated block, ping-ponged between consecutive blocks: the deferred
        # fast-token update of block i reads block i's buffer while block i + 1 already fills the other one
        self.T2 = [torch.empty(rows, C, device=dev, dtype=torch.float32) for _ in range(2)]
        self.T = self.T2[0]
        self.stats = torch.zeros(rows, 2, device=dev, dtype=torch.int64)   # sub-LN fixed-point [sum, sum sq] per MLP row
        self.merge_cnt = torch.zeros(max(v["nW"] for v in self.win.values()), device=dev, dtype=torch.int32)
        # side branches of this group of views (one workspace per view group, so concurrent groups never queue behind each
        # other): pad-slot k / v constants of the dense blocks; image-level sort + tables of the second window size
        self.fill_stream = torch.cuda.Stream(device=dev)
        self.side = torch.cuda.Stream(device=dev)
        self.stage = {}                                # (stage, ws) -> selection tables


class _Engine:
    """Device-resident, repacked (bf16, padded, interleaved) weights + the launch sequence."""

    def __init__(self, model, device):
        self.device = device
        m = model
        self.C, self.heads, self.patch = m.embed_dim, m.num_heads, m.patch_size
        self.block_ws = [b.window_size for b in m.blocks]
        self.block_acc = [b.accelerate for b in m.blocks]
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=device, dtype=torch.float32).to(torch.bfloat16).contiguous()
        C = self.C
        self.Hd = m.blocks[0].mlp.w1.out_features
        self.Hp = hidden_pad(self.Hd)
        self.w_pe = b16(m.patch_embed.proj.weight.reshape(C, -1))
        self.b_pe = f32(m.patch_embed.proj.bias)
        self.pos_embed = f32(m.pos_embed) if m.pos_embed is not None else None
        self.pos_cache = {}
        self.blocks = []
        for b in m.blocks:
            a = b.attn
            zeros = torch.zeros(C, device=device)
            qb = f32(a.q_bias) if a.q_bias is not None else zeros
            vb = f32(a.v_bias) if a.v_bias is not None else zeros
            w1, w2 = f32(b.mlp.w1.weight), f32(b.mlp.w2.weight)
            w12, b12 = interleave_w12(w1, f32(b.mlp.w1.bias), w2, f32(b.mlp.w2.bias), self.Hp)
            ft = a.rope.ft
            self.blocks.append(dict(
                n1w=f32(b.norm1.weight), n1b=f32(b.norm1.bias), n2w=f32(b.norm2.weight), n2b=f32(b.norm2.bias),
                wqkv=b16(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                bqkv=torch.cat([qb, zeros, vb]).contiguous(), vb=vb.contiguous(),
                # k / v of a pad slot selected as slow token in an accelerated block: norm1(0) = beta through
                # k_proj / v_proj (bf16 operands like the GEMM), before the per-slot rotation
                kpad=(b16(a.k_proj.weight).float() @ b16(b.norm1.bias).float()).contiguous(),
                vpad=(b16(a.v_proj.weight).float() @ b16(b.norm1.bias).float() + vb).contiguous(),
                wproj=b16(a.proj.weight), bproj=f32(a.proj.bias),
                w12=w12.to(torch.bfloat16).contiguous(), b12=b12,
                # SwiGLU sub-LN (eva_vit.py:48) folded into the w3 GEMM: w3g = W3 * gamma (columns),
                # u3 = W3 gamma, c3 = W3 beta + b3; the epilogue applies rstd * acc - rstd * mean * u3 + c3
                w3=F.pad(f32(b.mlp.w3.weight) * f32(b.mlp.ffn_ln.weight)[None, :],
                         (0, self.Hp - self.Hd)).to(torch.bfloat16).contiguous(),
                u3=(f32(b.mlp.w3.weight) @ f32(b.mlp.ffn_ln.weight)).contiguous(),
                b3=(f32(b.mlp.w3.weight) @ f32(b.mlp.ffn_ln.bias) + f32(b.mlp.w3.bias)).contiguous(), ft=ft,
                cos=f32(a.rope.freqs_cos).reshape(ft, ft, -1)[:, 0, 0:32:2].contiguous(),
                sin=f32(a.rope.freqs_sin).reshape(ft, ft, -1)[:, 0, 0:32:2].contiguous(),
            ))
        self.sel = []
        for s in getattr(m, "score_predictor", []):
            self.sel.append(dict(
                w_in=f32(s.input_proj[0].weight), b_in=f32(s.input_proj[0].bias),
                w_agg=f32(s.aggregate[0].weight), b_agg=f32(s.aggregate[0].bias),
                ln_w=f32(s.in_conv[0].weight), ln_b=f32(s.in_conv[0].bias),
                w_ic=b16(s.in_conv[1].weight), b_ic=f32(s.in_conv[1].bias),
                w_o0=b16(s.out_conv[0].weight), b_o0=f32(s.out_conv[0].bias),
                w_o2=b16(s.out_conv[2].weight), b_o2=f32(s.out_conv[2].bias),
                # 256 -> 2 head padded to 8 output rows (GEMM N granularity)
                w_o4=F.pad(b16(s.out_conv[4].weight), (0, 0, 0, 6)).contiguous(),
                b_o4=F.pad(f32(s.out_conv[4].bias), (0, 6)).contiguous(),
            ))
        self.sel_blob = self.sel_scale = None
        if len(self.sel):
            preds = list(m.score_predictor)
            self.sel_Q = preds[0].num_queries
            self.sel_blob = torch.stack([pack_motion_blob(s, self.sel_Q, C) for s in preds]).to(device).contiguous()
            self.sel_scale = preds[0].scale
        self.ws_cache = {}
        self._gstreams = []
        self.side = torch.cuda.Stream(device=device)      # history-query encoding + folding overlaps the stem and the dense blocks
        self.seed_t = torch.zeros(1, device=device, dtype=torch.int64)   # forward counter (Gumbel seed, graph-safe)

    # -- helpers ---------------------------------------------------------------------------
    def workspace(self, V, H, W, group=0):
        key = (V, H, W, group)
        if key not in self.ws_cache:
            self.ws_cache[key] = _Workspace(self, V, H, W)
        return self.ws_cache[key]

    def group_streams(self, G):
        while len(self._gstreams) < G:
            self._gstreams.append(torch.cuda.Stream(device=self.device))
        return self._gstreams

    def abs_pos(self, H, W):
        """eva_utils.py:229-258, cached per token grid (the reference re-interpolates every call)."""
        if self.pos_embed is None:
            return None
        if (H, W) not in self.pos_cache:
            a = self.pos_embed[:, 1:] if self.has_cls else self.pos_embed
            size = int(math.sqrt(a.shape[1]))
            assert size * size == a.shape[1]
            if size != H or size != W:
                a = F.interpolate(a.reshape(1, size, size, -1).permute(0, 3, 1, 2), size=(H, W), mode="bicubic",
                                  align_corners=False).permute(0, 2, 3, 1)
            self.pos_cache[(H, W)] = a.reshape(H * W, self.C).contiguous()
        return self.pos_cache[(H, W)]

    # -- stem ------------------------------------------------------------------------------
    def stem(self, img, wsp, X_out=None, pre=None):
        """img: fp32 NCHW (the reference's input), or - row f3 - the u8 HWC camera crop with `pre` (ImagePreprocess)."""
        V = img.shape[0]
        img = img if img.is_contiguous() else img.contiguous()
        if img.dtype == torch.uint8:
            L.preprocess_patch16_u8(img, pre.lut(self.device), wsp.cols, V, img.shape[1], img.shape[2], wsp.H * 16, wsp.W * 16,
                                    pre.to_rgb)
        else:
            L.im2col_patch16(img, wsp.cols, V, wsp.H * 16, wsp.W * 16)
        X = X_out if X_out is not None else torch.empty(V * wsp.N, self.C, device=self.device, dtype=torch.float32)
        pos = self.abs_pos(wsp.H, wsp.W)
        if pos is not None:
            L.gemm(wsp.cols, self.w_pe, L.EPI_RESID, bias=self.b_pe, out=X, resid=pos, resid_mod=wsp.N)
        else:
            L.gemm(wsp.cols, self.w_pe, L.EPI_LINEAR, bias=self.b_pe, out=X, out_f32=True)
        return X

    # -- MLP shared by both block kinds ----------------------------------------------------------
    def _mlp(self, bp, wsp, M, **resid_kw):
        """eva_vit.py:44-51.  wsp.a holds the bf16 norm2 rows; wsp.stats rows [0, M) must be zero on entry (zeroed by the
        norm2 launch)."""
        L.gemm(wsp.a, bp["w12"], L.EPI_SWIGLU, M=M, bias=bp["b12"], out=wsp.hid, row_stats=wsp.stats)
        L.gemm(wsp.hid, bp["w3"], L.EPI_RESID, M=M, bias=bp["b3"], ldo=self.C, ln_stats=wsp.stats, ln_u=bp["u3"],
               ln_n=self.Hd, ln_eps=LN_EPS, **resid_kw)

    def _qkv_attn(self, bp, wsp, M, nW, seq, rope_rows, rope_slots, qkv_out_map=None, attn_out_map=None, q_rows=None,
                  item_order=None, join=None, kv_rows=None, pad_v=None):
        """q/k/v for the M rows of wsp.a (scattered to window slots through qkv_out_map when given), then attention
        over nW windows of seq slots; attn_out_map sends the rows that are used afterwards to compact positions."""
        C = self.C
        L.gemm(wsp.a, bp["wqkv"], L.EPI_QKV_ROPE, M=M, bias=bp["bqkv"], out=wsp.qkv, rope_rows=rope_rows,
               rope_slots=rope_slots, rope_ft=bp["ft"], rope_cols=2 * C, q_scale=64 ** -0.5,
               cos_axis=bp["cos"], sin_axis=bp["sin"], out_map=qkv_out_map)
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)
        L.window_attention(wsp.qkv, wsp.ao, nW, seq, self.heads, out_map=attn_out_map, q_rows=q_rows,
                           item_order=item_order, kv_rows=kv_rows, pad_v=pad_v)

    def dense_block(self, i, X, wsp):
        """eva_vit.py:247-268.  The reference pads the normalised map to whole windows and runs q/k/v, attention and
        proj on every slot; the pad slots are exact zeros after norm1, so their k is 0 and their v is v_bias, and
        their own outputs are cropped by window_unpartition.  Here q/k/v and proj run over the real tokens only
        (image-row order) and the pad slots are never materialised: with k = 0 their score is 0 for every query, so the
        attention kernel adds them as one closed-form softmax term (analytic pad keys: kv_rows + pad_v).  Windows longer
        than the tcgen05 kernels take (> 448 slots, no shipped config) fall back to writing the constants (fill_pad_kv)."""
        bp, C = self.blocks[i], self.C
        w = wsp.win[self.block_ws[i]]
        VN = wsp.V * wsp.N
        analytic = w["n"] <= 448 and not os.environ.get("TOC3D_NO_ANALYTIC_PADS")     # env: A/B diagnostic only (tools/)
        fs = None
        if not analytic:
            # the pad slots' constants touch only pad rows of the qkv buffer: written on a second stream, concurrently with
            # norm1 and the q/k/v GEMM (which writes the real slots), joined before the attention
            cur, fs = torch.cuda.current_stream(), wsp.fill_stream
            fs.wait_stream(cur)                                  # the previous attention has finished reading qkv
            with torch.cuda.stream(fs):
                L.fill_pad_kv(wsp.qkv, w["pad_rows"], bp["vb"], C)
        L.layernorm_rows(X, bp["n1w"], bp["n1b"], wsp.a, VN, C, LN_EPS)
        self._qkv_attn(bp, wsp, VN, w["nW"], w["n"], w["rope_slot"], 0, qkv_out_map=w["slot_of_row"], attn_out_map=w["map"],
                       q_rows=w["q_rows"], item_order=w["item_order"], join=fs,
                       kv_rows=w["q_rows"] if analytic else None, pad_v=bp["vb"] if analytic else None)
        L.gemm(wsp.ao, bp["wproj"], L.EPI_RESID, M=VN, bias=bp["bproj"], out=X, ldo=C, resid=X)
        L.layernorm_rows(X, bp["n2w"], bp["n2b"], wsp.a, VN, C, LN_EPS, zero_stats=wsp.stats)
        self._mlp(bp, wsp, VN, out=X, resid=X)

    def select_windows(self, stage, score, ratio, wsp, first_ws=None):
        """Per-window stable top-k tables for every window size used after this stage.  The reference
        re-sorts inside each of the 6 blocks of a stage (toc3d_eva_vit.py:419); the indices only depend
        on (stage scores, window size), so they are computed once per (stage, ws).  The tables of the window size the
        next block needs (first_ws) are built on the launch stream, the others on the side stream (they are first
        needed two blocks later); toc3d_block waits for their event."""
        dev = self.device
        cur = torch.cuda.current_stream()
        order = sorted(wsp.win.keys(), key=lambda ws: (ws != first_ws, ws))
        for ws in order:
            w = wsp.win[ws]
            n, nW = w["n"], w["nW"]
            k = int(n * ratio)                      # toc3d_utils.py:136
            if k >= n:
                raise NotImplementedError("token_ratio=1.0 hits a latent bug in the reference "
                                          "(toc3d_eva_vit.py:462-463) and is not supported")
            t = wsp.stage.get((stage, ws))
            if t is None or t["k"] != k:
                i32 = dict(device=dev, dtype=torch.int32)
                # compact row space (real slow rows + representative per window); its layout is static: a window
                # keeps min(k, #real tokens) real rows, because real scores (log-probabilities) outrank the -1e6 pads
                rcap = torch.minimum(w["real_per_window"], torch.tensor(k, dtype=torch.int32))
                coff = torch.cumsum(rcap + 1, 0, dtype=torch.int32) - (rcap + 1)
                Mc = int((rcap + 1).sum())
                t = dict(k=k, nf=n - k, tok_map=torch.empty(nW * (k + 1), **i32),
                         rope_rows=torch.empty(nW * (k + 1), **i32), fast_map=torch.empty(nW, n - k, **i32),
                         fast_score=torch.empty(nW, n - k, device=dev),
                         rep2=[torch.empty(nW, self.C, device=dev) for _ in range(2)],
                         fast_win=torch.empty(wsp.V * wsp.N, **i32),
                         Mc=Mc, rcap=rcap.to(dev), coff=coff.to(dev), cmap=torch.empty(nW * (k + 1), **i32),
                         ctok=torch.empty(Mc, **i32), rep_row=torch.empty(nW, **i32), cinv=torch.empty(Mc, **i32),
                         crope=torch.empty(Mc, **i32), prope=torch.empty(nW * (k + 1), **i32), q_rows=(rcap + 1).to(dev),
                         item_order=balanced_item_order(rcap + 1, self.heads).to(dev))
                wsp.stage[(stage, ws)] = t
            on_side = first_ws is not None and ws != first_ws
            if on_side:
                wsp.side.wait_stream(cur)
            with torch.cuda.stream(wsp.side) if on_side else contextlib.nullcontext():
                L.window_topk(score, wsp.V, wsp.H, wsp.W, ws, k, fast_score=t["fast_score"], tok_map=t["tok_map"],
                              rope_rows=t["rope_rows"], fast_map=t["fast_map"], fast_win=t["fast_win"])
                L.compact_rows(t["tok_map"], t["coff"], t["rcap"], nW, k, t["cmap"], t["ctok"], t["rep_row"],
                               rope_rows=t["rope_rows"], cinv=t["cinv"], crope=t["crope"], prope=t["prope"])
                t["ready"] = None
                if on_side:
                    t["ready"] = torch.cuda.Event()
                    t["ready"].record(wsp.side)

    def toc3d_block(self, i, X, wsp, stage, pending=None, defer=False):
        """toc3d_eva_vit.py:395-473 (accelerated branch).  The packed set of a window (k slow rows + rep) contains pad
        slots whenever the window holds fewer than k real tokens.  A pad row is norm1(0) = beta: it matters only as an
        attention key / value (block constants up to the RoPE rotation, written by fill_pad_kv_rope); its own
        attention / proj / norm2 / MLP results are cropped by window_unpartition (toc3d_eva_vit.py:459-461).  So
        norm1, q/k/v, proj, norm2 and the MLP run on the COMPACT rows (real slow rows + rep) only; the attention
        reads the window-packed qkv buffer and writes compact rows.

        pending: the previous accelerated block's deferred fast-token update (applied by this block's first launch, which
        reads every real row anyway).  defer=True: leave THIS block's fast-token update (toc3d_eva_vit.py:452-461) to the
        next block and return its description; else run it here and return None."""
        bp, C = self.blocks[i], self.C
        ws = self.block_ws[i]
        w, t = wsp.win[ws], wsp.stage[(stage, ws)]
        if t.get("ready") is not None:              # tables built on the side stream (select_windows)
            torch.cuda.current_stream().wait_event(t["ready"])
            t["ready"] = None
        nW, k, nf, Mc = w["nW"], t["k"], t["nf"], t["Mc"]
        Mp = nW * (k + 1)
        T, rep = wsp.T2[i % 2], t["rep2"][i % 2]
        # one launch: representative token (-> T[rep_row]) + norm1 of the compact rows + k / v of the pad rows
        L.ln_gather_merge(X, t["ctok"], t["fast_map"], t["fast_score"], bp["n1w"], bp["n1b"], wsp.a, rep,
                          T, nW, k, nf, C, LN_EPS,
                          rep_row=t["rep_row"], compact_rows=Mc,
                          pad_fill=(wsp.qkv, t["cmap"], t["prope"], Mp, bp["kpad"], bp["vpad"], bp["cos"], bp["sin"], bp["ft"]),
                          counters=wsp.merge_cnt, pending=pending)
        self._qkv_attn(bp, wsp, Mc, nW, k + 1, t["crope"], 0, qkv_out_map=t["cinv"], attn_out_map=t["cmap"],
                       q_rows=t["q_rows"], item_order=t["item_order"])
        L.gemm(wsp.ao, bp["wproj"], L.EPI_RESID, M=Mc, bias=bp["bproj"], out=T, ldo=C, resid=X,
               resid_map=t["ctok"], out_alt=T)                                          # t1 = t + attn
        L.layernorm_rows(T, bp["n2w"], bp["n2b"], wsp.a, Mc, C, LN_EPS, zero_stats=wsp.stats)
        self._mlp(bp, wsp, Mc, out=X, resid=T, out_map=t["ctok"], out_alt=T)             # t2 -> image rows
        if defer:
            return (t["fast_win"], T, t["rep_row"], rep)
        L.fast_token_update(X, t["fast_map"], T, rep, nW, nf, k, C, rep_row=t["rep_row"])
        return None

    # -- scorers ----------------------------------------------------------------------------
    def fold_all_queries(self, q_kw, V):
        """Motion-aware query encoding (toc3d_utils.py:334-360) + folding of the query bank into (A, c) for ALL stages in
        one C-ABI call (two launches).  Depends only on the history queries, not on the image, so the caller runs it on
        the side stream.  -> ([(A_j, c_j)], q_all [S,Bf,Q,256])."""
        tq = q_kw["temp_queries"]
        Bf, Q = tq.shape[0], tq.shape[1]
        assert V % Bf == 0, "views*batch must be a multiple of the query batch (toc3d_utils.py:240)"
        assert Q == self.sel_Q and tq.shape[2] == 256
        S = len(self.sel)
        dev = self.device
        f32c = lambda t: t.to(torch.float32).contiguous()
        ts = q_kw["temp_timestamp"]
        ts = ts.contiguous() if ts.dtype == torch.float64 else f32c(ts)
        q_all = torch.empty(S, Bf, Q, 256, device=dev)
        A = torch.empty(S, Bf, 2, self.C, device=dev); c = torch.empty(S, Bf, 2, device=dev)
        L.motion_queries_fold(self.sel_blob, f32c(tq), f32c(q_kw["temp_ref_points"]), f32c(q_kw["temp_vel"]), ts,
                              f32c(q_kw["temp_ego_pose"]), f32c(q_kw["ego_pose_inv"]), self.sel_scale, self.C, q_all, A, c)
        return [(A[j], c[j]) for j in range(S)], q_all

    def score_stage(self, j, X, mask_prev, wsp, folded, gumbel, seed=None):
        """folded = (A, c) from fold_queries when the previous frame exists, else None (first-frame scorer)."""
        dev, V, N, C = self.device, wsp.V, wsp.N, self.C
        sp = self.sel[j]
        pred = torch.empty(V, N, 2, device=dev)
        score = torch.empty(V, N, device=dev)
        mask = torch.empty(V, N, device=dev)
        seed = j if seed is None else seed
        if folded is not None:
            A, c = folded
            L.score_tokens(X, mask_prev, A, c, V, N, C, V // A.shape[0], gumbel, seed, pred, score, mask,
                           seed_dev=self.seed_t)
        else:
            VN = V * N
            src = X
            if mask_prev is not None:
                src = torch.empty_like(X)
                L.mask_rows(X, mask_prev, src, VN, C)
            L.layernorm_rows(src, sp["ln_w"], sp["ln_b"], wsp.a, VN, C, 1e-5)
            y = wsp.ao[:VN]
            L.gemm(wsp.a, sp["w_ic"], L.EPI_LINEAR, M=VN, bias=sp["b_ic"], out=y, act=L.ACT_GELU)
            L.global_half_mean(y, V, N, C)
            y1 = wsp.qkv.view(-1)[: VN * (C // 2)].view(VN, C // 2)
            L.gemm(y, sp["w_o0"], L.EPI_LINEAR, M=VN, bias=sp["b_o0"], out=y1, act=L.ACT_GELU)
            y2 = wsp.hid.view(-1)[: VN * (C // 4)].view(VN, C // 4)
            L.gemm(y1, sp["w_o2"], L.EPI_LINEAR, M=VN, bias=sp["b_o2"], out=y2, act=L.ACT_GELU)
            logits8 = torch.empty(VN, 8, device=dev)
            L.gemm(y2, sp["w_o4"], L.EPI_LINEAR, M=VN, bias=sp["b_o4"], out=logits8, out_f32=True)
            logits = logits8[:, :2].contiguous()
            L.score_finish(logits, VN, gumbel, seed, pred, score, mask, seed_dev=self.seed_t)
        return pred, score, mask


class _EvaBase(nn.Module):
    """Shared construction / engine plumbing of the two backbones."""

    def _build_common(self, img_size, patch_size, in_chans, embed_dim, num_heads, use_abs_pos, pretrain_img_size,
                      pretrain_use_cls_token, pt_hw_seq_len, intp_freq, window_size):
        if patch_size != 16 or in_chans != 3:
            raise NotImplementedError("the sm_100a stem kernel is specialised for 16x16 RGB patches")
        if embed_dim // num_heads != 64:
            raise NotImplementedError("attention / RoPE kernels are specialised for head_dim 64")
        if embed_dim not in (128, 256, 512, 768, 1024):
            raise NotImplementedError("the fused norm1 / merge kernel supports embed_dim 128, 256, 512, 768, 1024")
        self.embed_dim, self.num_heads, self.patch_size = embed_dim, num_heads, patch_size
        self.pretrain_use_cls_token = pretrain_use_cls_token
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        if use_abs_pos:
            npos = (pretrain_img_size // patch_size) ** 2 + (1 if pretrain_use_cls_token else 0)
            self.pos_embed = nn.Parameter(torch.zeros(1, npos, embed_dim))
        else:
            self.pos_embed = None
        hh = embed_dim // num_heads // 2
        self.rope_win = _Rope(hh, pt_hw_seq_len, window_size if intp_freq else None)
        self.rope_glb = _Rope(hh, pt_hw_seq_len, img_size // patch_size if intp_freq else None)
        self._engine = None
        self._graphs = {}
        self.use_cuda_graph = True     # replay the whole forward as one CUDA graph per (shape, prev_exists)
        # views are independent: with G > 1 the forward runs G groups of views on their own streams inside the
        # one CUDA graph, so the tail / epilogue of one group's kernels overlaps the other group's kernels
        self.view_groups = 1
        # fast-token updates of an accelerated block are applied by the next block's first launch (15 of 18 launches less
        # in the shipped configs; bit-identical results).  False: one toc3d_fast_token_update launch per block.
        self.defer_fast_update = True
        self.graph_outputs = "clone"   # "static": return the graph's own buffers (overwritten by the next call)

    def _init_weights(self):
        """toc3d_eva_vit.py:214-228 / eva_vit.py:396-407."""
        if self.pos_embed is not None:
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    # any change of the parameters invalidates the repacked device copies
    def _apply(self, fn, *a, **k):
        self.refresh_weights()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.refresh_weights()
        return super().load_state_dict(*a, **k)

    def _load_from_state_dict(self, *a, **k):
        # mmcv's load_checkpoint (what the reference uses) calls this on every module directly, bypassing load_state_dict
        self.refresh_weights()
        return super()._load_from_state_dict(*a, **k)

    def _option_key(self):
        """Options that change the captured launch sequence: part of every graph key."""
        return (self.view_groups, self.defer_fast_update, self.graph_outputs, id(getattr(self, "_fused_neck", None)),
                tuple(getattr(self, "token_ratio", ()) or ()))

    def refresh_weights(self):
        """Call after mutating parameters in place (the engine holds repacked bf16 copies, the captured
        graphs hold pointers into it)."""
        self._engine = None
        self._graphs = {}

    def _graphed(self, key, inputs, core):
        """Run core(static_inputs) -> tuple of tensors through a CUDA graph captured once per `key`.
        inputs: dict name -> CUDA tensor (copied into the graph's static input buffers before replay)."""
        entry = self._graphs.get(key)
        if entry is None:
            static_in = {k: v.clone() for k, v in inputs.items()}
            core(static_in)                         # eager warm-up: lazy inits, workspaces, selection tables
            g = torch.cuda.CUDAGraph()
            l0 = L.launch_count
            with torch.cuda.graph(g):
                out = core(static_in)
            entry = (g, static_in, out, L.launch_count - l0)
            self._graphs[key] = entry
        g, static_in, out, n_launch = entry
        for k, v in inputs.items():
            static_in[k].copy_(v, non_blocking=True)
        g.replay()
        L.launch_count += n_launch
        if self.graph_outputs == "static":
            return out
        return tuple(t.clone() for t in out)

    def _get_engine(self, x):
        if not x.is_cuda:
            raise RuntimeError("toc3d_b200 backbones run on CUDA (sm_100a) only; got a %s tensor. "
                               "There is no CPU fallback." % x.device)
        if self.training:
            raise RuntimeError("toc3d_b200 backbones are inference-only; call .eval()")
        if self._engine is None or self._engine.device != x.device:
            L.load()
            self._engine = _Engine(self, x.device)
            self._engine.has_cls = self.pretrain_use_cls_token
        return self._engine

    def fuse_neck(self, neck):
        """Run `neck` (toc3d_b200.CPFPN) at the end of this backbone's own launch sequence / CUDA graph.  The plugin
        boundary is unchanged: forward returns what it always returns, and `neck(list(img_feats.values()))` - the call
        Petr3D makes (petr3d.py:188-190) - hands back the maps computed here instead of launching anything.  None undoes it."""
        import weakref
        object.__setattr__(self, "_fused_neck", neck)      # NOT a sub-module: the state-dict keys stay the reference's
        if neck is not None:
            if not hasattr(neck, "_fused_into"):
                neck._fused_into = []
            neck._fused_into.append(weakref.ref(self))
        self._graphs = {}

    def _neck_tail(self, X, V, H, W):
        """-> tuple with the fused neck's level-0 rows (fp32 [V*H*W, out_channels]) or ()."""
        neck = getattr(self, "_fused_neck", None)
        return () if neck is None else (neck.launch(X, V, H, W),)

    def _attach_neck(self, last_feat, extra, V, H, W):
        neck = getattr(self, "_fused_neck", None)
        if neck is not None:
            last_feat._toc3d_fused_neck = (neck, neck.levels(extra[0], V, H, W))
        return last_feat

    def set_image_preprocess(self, mean, std, to_rgb=True, size_divisor=32, size=None):
        """Row f3: take over `NormalizeMultiviewImage(**img_norm_cfg)` + `PadMultiViewImage(size_divisor=32)`
        (transform_3d.py:21-104).  Afterwards `forward(x=...)` also accepts the uint8 HWC camera crops
        (B*views, Hs, Ws, 3) that the reference pipeline would have normalised on the CPU."""
        self.img_preprocess = ImagePreprocess(mean, std, to_rgb, size_divisor, size)
        self._graphs = {}

    def _prep_img(self, x):
        """-> (x, V, Hi, Wi) with Hi x Wi the (padded) image size the patch grid is built on."""
        if x.dim() != 4:
            raise ValueError("expected images of shape (B*views, 3, H, W)")
        if x.dtype == torch.uint8:
            pre = getattr(self, "img_preprocess", None)
            if pre is None:
                raise RuntimeError("uint8 images need set_image_preprocess(mean, std, to_rgb) (the img_norm_cfg of the config)")
            if x.shape[-1] != 3:
                raise ValueError("uint8 images are camera crops of shape (B*views, H, W, 3)")
            Hi, Wi = pre.padded_hw(x.shape[1], x.shape[2])
            return x.contiguous(), x.shape[0], Hi, Wi
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise ValueError("image %dx%d must be a multiple of the 16x16 patch" % (x.shape[2], x.shape[3]))
        return x.float().contiguous(), x.shape[0], x.shape[2], x.shape[3]


@_register
class EVA_ViT(_EvaBase):
    """Dense EVA-02 ViT backbone (eva_vit.py:270-428); returns {'last_feat': (V,C,H/16,W/16)}."""

    def __init__(self, img_size=1024, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4 * 2 / 3, qkv_bias=True, drop_path_rate=0.0, norm_layer=None, act_layer=None,
                 use_abs_pos=True, use_rel_pos=False, rope=True, pt_hw_seq_len=16, intp_freq=True, window_size=0,
                 global_window_size=20, use_checkpoint=True, global_attn_indexes=(), residual_block_indexes=(),
                 use_act_checkpoint=False, pretrain_img_size=224, pretrain_use_cls_token=True,
                 return_intermediate=False, out_feature="last_feat", xattn=True):
        super().__init__()
        if len(residual_block_indexes) or return_intermediate or window_size <= 0:
            raise NotImplementedError("residual blocks / return_intermediate / window_size=0 are not used by any "
                                      "shipped config and are not implemented")
        self._build_common(img_size, patch_size, in_chans, embed_dim, num_heads, use_abs_pos, pretrain_img_size,
                           pretrain_use_cls_token, pt_hw_seq_len, intp_freq, window_size)
        self.blocks = nn.ModuleList()
        for i in range(depth):
            g = i in global_attn_indexes
            self.blocks.append(_Block(embed_dim, num_heads, mlp_ratio, qkv_bias,
                                      global_window_size if g else window_size,
                                      self.rope_glb if g else self.rope_win, accelerate=False))
        self._out_features = [out_feature]
        self._out_feature_channels = {out_feature: embed_dim}
        self._out_feature_strides = {out_feature: patch_size}
        self._init_weights()

    @torch.no_grad()
    def forward(self, x, *args, tap=None, **kwargs):
        """tap: test hook (dict) - collects per-block outputs; tap["inject_block_in"] replaces each block's input."""
        x, V, Hi, Wi = self._prep_img(x)
        eng = self._get_engine(x)
        pre = getattr(self, "img_preprocess", None)

        def core(t):
            wsp = eng.workspace(V, Hi // 16, Wi // 16)
            X = eng.stem(t["x"], wsp, pre=pre)
            if tap is not None:
                tap["stem"] = X.clone()
                tap["block_out"] = []
            for i in range(len(self.blocks)):
                if tap is not None and "inject_block_in" in tap:
                    X.copy_(tap["inject_block_in"][i].reshape(X.shape))
                eng.dense_block(i, X, wsp)
                if tap is not None:
                    tap["block_out"].append(X.clone())
            return (X,) + self._neck_tail(X, V, Hi // 16, Wi // 16)

        GLOBAL_TIMER.event_start("StreamPETR-EVA-ViT/backbone")
        if self.use_cuda_graph and tap is None:
            X, *extra = self._graphed(("dense", V, Hi, Wi, tuple(x.shape), x.dtype, self._option_key()), {"x": x}, core)
        else:
            X, *extra = core({"x": x})
        GLOBAL_TIMER.event_end("StreamPETR-EVA-ViT/backbone")
        lf = X.view(V, Hi // 16, Wi // 16, -1).permute(0, 3, 1, 2)
        return {self._out_features[0]: self._attach_neck(lf, extra, V, Hi // 16, Wi // 16)}


@_register
class ToC3DEVAViT(_EvaBase):
    """EVA-02 ViT with ToC3D motion-aware token compression (toc3d_eva_vit.py:25-326)."""

    def __init__(self, img_size=1024, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4 * 2 / 3, qkv_bias=True, drop_path_rate=0.0, norm_layer=None, act_layer=None,
                 use_abs_pos=True, use_rel_pos=False, rope=True, rope_acc=False, pt_hw_seq_len=16, intp_freq=True,
                 window_size=0, global_window_size=20, use_checkpoint=True, global_attn_indexes=(),
                 residual_block_indexes=(), use_act_checkpoint=False, pretrain_img_size=224,
                 pretrain_use_cls_token=True, out_feature="last_feat", return_intermediate=False, xattn=True,
                 pruning_loc=None, pruning_score_type="attention", score_mask=True, pruning_attn_scale=True,
                 pruning_num_queries=256, accelerate_global=True, token_ratio=None, use_represent_tokens=True,
                 pc_range=None, token_selection_loss=None):
        super().__init__()
        if pruning_score_type != "attention":
            raise NotImplementedError("Not supported score type: %s, only support: ['attention']" % pruning_score_type)
        unsupported = dict(residual_block_indexes=len(residual_block_indexes) > 0, return_intermediate=return_intermediate,
                           rope=not rope, rope_acc=not rope_acc, accelerate_global=not accelerate_global,
                           use_represent_tokens=not use_represent_tokens, score_mask=not score_mask,
                           pruning_attn_scale=not pruning_attn_scale, window_size=window_size <= 0)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError("options %s take code paths no shipped ToC3D config reaches "
                                      "(SURVEY.md §8a) and are not implemented" % bad)
        assert pruning_loc is not None and token_ratio is not None and pc_range is not None
        assert len(set(pruning_loc) & set(global_attn_indexes)) == 0, \
            "The pruning score calculation layer cannot be the global attention layer"      # toc3d_eva_vit.py:141
        assert list(pruning_loc) == sorted(pruning_loc) and len(token_ratio) >= len(pruning_loc)
        self._build_common(img_size, patch_size, in_chans, embed_dim, num_heads, use_abs_pos, pretrain_img_size,
                           pretrain_use_cls_token, pt_hw_seq_len, intp_freq, window_size)
        self.pruning_loc = pruning_loc
        self.pruning_num_queries = pruning_num_queries
        self.token_ratio = token_ratio
        self.use_represent_tokens = use_represent_tokens
        self.accelerate_global = accelerate_global
        self.token_selection_loss = None
        if token_selection_loss is not None:
            try:
                from mmdet3d.models.builder import build_loss
                self.token_selection_loss = build_loss(token_selection_loss)
            except ImportError:
                # the loss is a training-time consumer (toc3d_eva_vit.py:312-326); the shipped configs always pass it, and
                # inference never uses it, so a stand-alone (mmdet3d-less) construction keeps going
                import warnings
                warnings.warn("token_selection_loss ignored: mmdet3d is not importable (the loss is training-only)")
        self.score_predictor = nn.ModuleList([
            _Selector(embed_dim, pruning_num_queries, token_ratio[i], pc_range) for i in range(len(pruning_loc))])
        hh = embed_dim // num_heads // 2
        self.rope_win_acc = _Rope(hh, pt_hw_seq_len, window_size if intp_freq else None)
        self.rope_glb_acc = _Rope(hh, pt_hw_seq_len, img_size // patch_size if intp_freq else None)
        self.blocks = nn.ModuleList()
        for i in range(depth):
            g = i in global_attn_indexes
            acc = len(pruning_loc) > 0 and i >= pruning_loc[0]                               # toc3d_eva_vit.py:178-180
            rope_m = (self.rope_glb_acc if g else self.rope_win_acc) if acc else (self.rope_glb if g else self.rope_win)
            self.blocks.append(_Block(embed_dim, num_heads, mlp_ratio, qkv_bias,
                                      global_window_size if g else window_size, rope_m, accelerate=acc))
        self._out_features = [out_feature]
        self._out_feature_channels = {out_feature: embed_dim}
        self._out_feature_strides = {out_feature: patch_size}
        self.gumbel_seed = 0
        self._init_weights()

    @torch.no_grad()
    def forward(self, x, temp_queries=None, prev_exists=None, temp_ref_points=None, temp_vel=None,
                temp_timestamp=None, temp_ego_pose=None, ego_pose_inv=None, *args, gumbel_noise=None,
                teacher_scores=None, tap=None, **kwargs):
        """Same keyword contract as the reference (extra kwargs such as gt_bboxes are ignored).

        gumbel_noise: optional list of per-stage (V,N,2) tensors (parity pin 2); default draws
        -log(-log(u)) on device.  teacher_scores / tap are test hooks (teacher forcing, intermediates;
        tap["inject_block_in"] = per-block inputs replaces the residual stream before every block).
        """
        x, V, Hi, Wi = self._prep_img(x)
        eng = self._get_engine(x)
        H, W = Hi // 16, Wi // 16
        prev = bool(prev_exists) if prev_exists is not None else False
        q_names = ("temp_queries", "temp_ref_points", "temp_vel", "temp_timestamp", "temp_ego_pose", "ego_pose_inv")
        q_vals = (temp_queries, temp_ref_points, temp_vel, temp_timestamp, temp_ego_pose, ego_pose_inv)
        tensors = {"x": x}
        if prev:
            assert ego_pose_inv is not None                                              # toc3d_utils.py:345
            tensors.update({k: v.contiguous() for k, v in zip(q_names, q_vals)})
        nst = len(self.pruning_loc)

        def core(t):
            q_kw = {k: t[k] for k in q_names} if prev else None
            flat = self._forward_core(eng, t["x"], (H, W), q_kw, gumbel_noise, teacher_scores, tap)
            return tuple(flat) + self._neck_tail(flat[0], V, H, W)

        GLOBAL_TIMER.event_start("ToC3D-StreamPETR-EVAViT/backbone")
        if self.use_cuda_graph and gumbel_noise is None and teacher_scores is None and tap is None:
            key = ("toc3d", prev, self._option_key()) + tuple((k, tuple(v.shape), v.dtype) for k, v in tensors.items())
            flat = self._graphed(key, tensors, core)
        else:
            flat = core(tensors)
        GLOBAL_TIMER.event_end("ToC3D-StreamPETR-EVAViT/backbone")
        X, masks, keeps, drops = flat[0], list(flat[1:1 + nst]), list(flat[1 + nst:1 + 2 * nst]), list(flat[1 + 2 * nst:1 + 3 * nst])
        outputs = {self._out_features[0]: self._attach_neck(X.view(V, H, W, -1).permute(0, 3, 1, 2), flat[1 + 3 * nst:], V, H, W)}
        none_if_empty = lambda l: l if len(l) else None
        return ToC3DViTReturnType(outputs, none_if_empty([m.view(V, H, W, 1) for m in masks]), None,
                                  keep_idx=none_if_empty(keeps), drop_idx=none_if_empty(drops), aux_outputs=None)

    def _forward_core(self, eng, x, grid, q_kw, gumbel_noise, teacher_scores, tap):
        """The launch sequence of one forward (toc3d_eva_vit.py:243-310); capturable in a CUDA graph.
        Returns (X, *masks, *keep_idx, *drop_idx).

        With view_groups = G > 1 the views are cut into G groups that run the whole block sequence on their own
        streams (every op is independent per image): the ragged tail wave / exposed epilogue of one group's
        kernel is filled by the other group's next kernel."""
        V = x.shape[0]
        H, W = grid
        N = H * W
        cur, side = torch.cuda.current_stream(), eng.side
        nst = len(self.pruning_loc)
        # side stream: the history queries do not depend on the image -> encode + fold all stages up front
        side.wait_stream(cur)
        folded = [None] * nst
        with torch.cuda.stream(side):
            eng.seed_t.add_(1)
            if q_kw is not None:
                folded, q_all = eng.fold_all_queries(q_kw, V)
                if tap is not None:
                    tap["motion_queries"] = q_all
            ev_q = torch.cuda.Event()
            ev_q.record(side)
        G = self.view_groups if (tap is None and self.view_groups > 1 and V % self.view_groups == 0) else 1
        if G > 1 and q_kw is not None:
            Bf = q_kw["temp_queries"].shape[0]
            vg = V // G
            if not (vg % (V // Bf) == 0 or (V // Bf) % vg == 0):     # a group must not straddle frames unevenly
                G = 1
        if G == 1:
            out = self._forward_views(eng, x, eng.workspace(V, H, W), folded, ev_q, gumbel_noise, teacher_scores, tap, 0, 0)
            cur.wait_stream(side)
            return out
        vg = V // G
        X = torch.empty(V * N, eng.C, device=x.device, dtype=torch.float32)
        streams = eng.group_streams(G)
        outs = []
        for g in range(G):
            st = streams[g]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                sl = slice(g * vg, (g + 1) * vg)
                gn = None if gumbel_noise is None else [t[sl] for t in gumbel_noise]
                ts = None if teacher_scores is None else [t.reshape(V, -1)[sl] for t in teacher_scores]
                fold_g = folded
                if q_kw is not None:            # this group's frames only (views are frame-major)
                    vpf = V // folded[0][0].shape[0]
                    f0, f1 = (g * vg) // vpf, ((g + 1) * vg - 1) // vpf + 1
                    fold_g = [(A[f0:f1].contiguous(), c[f0:f1].contiguous()) for A, c in folded]
                outs.append(self._forward_views(eng, x[sl], eng.workspace(vg, H, W, g), fold_g, ev_q, gn, ts, None, g,
                                                g * vg, X_out=X[g * vg * N:(g + 1) * vg * N]))
        for st in streams[:G]:
            cur.wait_stream(st)
        cur.wait_stream(side)
        cat = [torch.cat([o[1 + j] for o in outs], dim=0) for j in range(3 * nst)]
        return (X, *cat)

    def _forward_views(self, eng, x, wsp, folded, ev_q, gumbel_noise, teacher_scores, tap, group, view0, X_out=None):
        """Stem + all blocks for a contiguous group of views on the current stream."""
        V = x.shape[0]
        H, W = wsp.H, wsp.W
        N = H * W
        cur, side = torch.cuda.current_stream(), wsp.side
        X = eng.stem(x, wsp, X_out, pre=getattr(self, "img_preprocess", None))
        masks, keep_idxes, drop_idxes, scores_l = [], [], [], []
        mask_prev, stage = None, -1
        pending = None
        if tap is not None:
            tap["stem"] = X.clone()
            tap["block_out"] = []
        for i, blk in enumerate(self.blocks):
            if i in self.pruning_loc:
                stage += 1
                g = None
                if gumbel_noise is not None:
                    g = gumbel_noise[stage].to(device=x.device, dtype=torch.float32).contiguous()
                if stage == 0:
                    cur.wait_event(ev_q)
                if tap is not None and "inject_block_in" in tap:      # test hook: per-block isolation (block i starts from
                    X.copy_(tap["inject_block_in"][i].reshape(X.shape))  # the checker's input, so errors do not accumulate)
                pred, score, mask = eng.score_stage(stage, X, mask_prev, wsp, folded[stage], g, seed=stage + 16 * group)
                if tap is not None:
                    tap.setdefault("scores_raw", []).append(score.view(V, H, W).clone())
                if teacher_scores is not None:
                    score = teacher_scores[stage].to(x.device).reshape(V, N).contiguous()
                k = int(N * self.token_ratio[stage])
                keep = torch.empty(V, k, device=x.device, dtype=torch.int64)
                drop = torch.empty(V, N - k, device=x.device, dtype=torch.int64)
                # the image-level sort only feeds the returned keep/drop lists -> off the critical path
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    L.topk_split(score, V, N, k, keep, drop)
                eng.select_windows(stage, score, self.token_ratio[stage], wsp, first_ws=blk.window_size)
                mask_prev = mask
                masks.append(mask)
                keep_idxes.append(keep)
                drop_idxes.append(drop)
                scores_l.append(score.view(V, H, W))
            if tap is not None and "inject_block_in" in tap:
                X.copy_(tap["inject_block_in"][i].reshape(X.shape))
            if blk.accelerate:
                # the fast-token update of this block is left to the next block's first launch when that is an
                # accelerated block of the same stage (nothing reads the residual stream in between)
                nxt = i + 1
                defer = (self.defer_fast_update and tap is None and nxt < len(self.blocks) and self.blocks[nxt].accelerate
                         and nxt not in self.pruning_loc)
                pending = eng.toc3d_block(i, X, wsp, stage, pending, defer)
            else:
                eng.dense_block(i, X, wsp)
            if tap is not None:
                tap["block_out"].append(X.clone())
        if tap is not None:
            tap["scores"] = scores_l
        cur.wait_stream(side)                      # image-level index lists
        return (X, *masks, *keep_idxes, *drop_idxes)

    def loss(self, pred_masks, gt_bboxes, *args, **kwargs):
        """toc3d_eva_vit.py:312-326 (training only; delegates to the mmdet3d-built loss when present)."""
        losses = dict()
        if self.token_selection_loss is not None:
            losses.update(self.token_selection_loss(pred_mask=pred_masks, gt_bboxes=gt_bboxes))
        return losses
