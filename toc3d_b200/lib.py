"""ctypes binding of libtoc3d_b200.so (include/toc3d_b200.h).

torch is used only for device memory and streams: every wrapper passes raw
device pointers + sizes + the current CUDA stream.  There is NO fallback: if the
shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtoc3d_b200.so")

EPI_LINEAR, EPI_QKV_ROPE, EPI_RESID, EPI_SWIGLU = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2

_c_void_p, _c_int, _c_i64, _c_float, _c_u64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                                               ctypes.c_uint64)


class Epilogue(ctypes.Structure):
    _fields_ = [
        ("bias", _c_void_p), ("out", _c_void_p), ("ldo", _c_int), ("out_f32", _c_int), ("act", _c_int),
        ("resid", _c_void_p), ("resid_map", _c_void_p), ("resid_mod", _c_int), ("out_map", _c_void_p),
        ("out_alt", _c_void_p), ("rope_rows", _c_void_p), ("rope_slots", _c_int), ("rope_ft", _c_int),
        ("rope_cols", _c_int), ("q_scale", _c_float), ("cos_axis", _c_void_p), ("sin_axis", _c_void_p),
        ("row_stats", _c_void_p), ("ln_stats", _c_void_p), ("ln_u", _c_void_p), ("ln_n", _c_int), ("ln_eps", _c_float),
        ("tile_n", _c_int), ("conv_cin", _c_int), ("conv_row_shift", _c_int * 9),
    ]


class PendingUpdate(ctypes.Structure):
    _fields_ = [("fast_win", _c_void_p), ("packed", _c_void_p), ("rep_row", _c_void_p), ("rep", _c_void_p)]


class PadFill(ctypes.Structure):
    _fields_ = [("qkv", _c_void_p), ("cmap", _c_void_p), ("rope_rows", _c_void_p), ("Mp", _c_int), ("kpad", _c_void_p),
                ("vpad", _c_void_p), ("cos_axis", _c_void_p), ("sin_axis", _c_void_p), ("ft", _c_int)]


_SIGS = {
    "toc3d_abi_version": ([], _c_int),
    "toc3d_last_error": ([], ctypes.c_char_p),
    "toc3d_gemm_bf16": ([_c_void_p, _c_i64, _c_void_p, _c_i64, _c_int, _c_int, _c_int, _c_int,
                         ctypes.POINTER(Epilogue), _c_void_p], _c_int),
    "toc3d_window_attention": ([_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                _c_void_p], _c_int),
    "toc3d_layernorm_rows": ([_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                              _c_float, _c_int, _c_void_p, _c_void_p], _c_int),
    "toc3d_subln_bf16": ([_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p],
                         _c_int),
    "toc3d_window_topk": ([_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                           _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p], _c_int),
    "toc3d_compact_rows": ([_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                            _c_void_p, _c_void_p, _c_void_p], _c_int),
    "toc3d_fill_pad_kv_rope": ([_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                _c_void_p], _c_int),
    "toc3d_fill_pad_kv": ([_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p], _c_int),
    "toc3d_topk_split": ([_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p], _c_int),
    "toc3d_merge_fast_tokens": ([_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                 _c_void_p, _c_void_p], _c_int),
    "toc3d_fast_token_update": ([_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                                 _c_void_p, _c_void_p], _c_int),
    "toc3d_ln_gather_merge": ([_c_void_p] * 9 + [_c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_void_p, _c_int,
                               ctypes.POINTER(PadFill), _c_void_p, ctypes.POINTER(PendingUpdate), _c_void_p], _c_int),
    "toc3d_motion_blob_floats": ([_c_int, _c_int], _c_i64),
    "toc3d_motion_queries_fold": ([_c_void_p, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                   _c_int, _c_void_p, _c_void_p, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p], _c_int),
    "toc3d_score_tokens": ([_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                            _c_u64, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p], _c_int),
    "toc3d_score_finish": ([_c_void_p, _c_int, _c_void_p, _c_u64, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                            _c_void_p], _c_int),
    "toc3d_im2col_patch16": ([_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p], _c_int),
    "toc3d_preprocess_patch16_u8": ([_c_void_p, _c_void_p, _c_void_p] + [_c_int] * 6 + [_c_void_p], _c_int),
    "toc3d_cast_f32_to_bf16": ([_c_void_p, _c_void_p, _c_i64, _c_void_p], _c_int),
    "toc3d_mask_rows": ([_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p], _c_int),
    "toc3d_global_half_mean": ([_c_void_p, _c_int, _c_int, _c_int, _c_void_p], _c_int),
}

EXPORTS = tuple(_SIGS.keys())
_lib = None
launch_count = 0   # number of kernel-launching C-ABI calls issued (bench.py's gpu_launches)


def load():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "toc3d_b200: %s is missing - build it with `python -m toc3d_b200.build` "
                "(there is no CPU or PyTorch fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
        if lib.toc3d_abi_version() != 14:
            raise RuntimeError("toc3d_b200: ABI version mismatch")
        _lib = lib
    return _lib


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "toc3d_b200 kernels need contiguous CUDA tensors"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(rc, name):
    global launch_count
    if rc != 0:
        msg = load().toc3d_last_error().decode()
        raise RuntimeError("%s failed (rc=%d): %s" % (name, rc, msg))
    launch_count += 1


def _want(t, dtype, name):
    assert t.dtype == dtype, "%s must be %s, got %s" % (name, dtype, t.dtype)


# ------------------------------------------------------------------------------- wrappers
def _epilogue(*, bias=None, out=None, ldo=None, out_f32=False, act=ACT_NONE, resid=None,
              resid_map=None, resid_mod=0, out_map=None, out_alt=None, rope_rows=None, rope_slots=0, rope_ft=0,
              rope_cols=0, q_scale=1.0, cos_axis=None, sin_axis=None, row_stats=None, ln_u=None, ln_n=0, ln_eps=0.0,
              tile_n=0, ln_stats=None, conv_cin=0, conv_row_shift=None):
    e = Epilogue()
    e.bias = _p(bias); e.out = _p(out); e.ldo = out.shape[-1] if ldo is None else ldo
    e.out_f32 = int(out_f32); e.act = act
    e.resid = _p(resid); e.resid_map = _p(resid_map); e.resid_mod = resid_mod
    e.out_map = _p(out_map); e.out_alt = _p(out_alt)
    e.rope_rows = _p(rope_rows); e.rope_slots = rope_slots; e.rope_ft = rope_ft; e.rope_cols = rope_cols
    e.q_scale = q_scale; e.cos_axis = _p(cos_axis); e.sin_axis = _p(sin_axis)
    e.row_stats = _p(row_stats); e.ln_stats = _p(ln_stats)
    e.ln_u = _p(ln_u); e.ln_n = ln_n; e.ln_eps = ln_eps; e.tile_n = tile_n
    e.conv_cin = conv_cin
    if conv_cin:
        assert len(conv_row_shift) == 9
        for i, v in enumerate(conv_row_shift):
            e.conv_row_shift[i] = int(v)
    return e


def gemm(A, B, kind, M=None, **epi):
    """C = A[M,K] @ B[N,K]^T with fused epilogue `kind` (see include/toc3d_b200.h; keywords = toc3d_epilogue fields)."""
    _want(A, torch.bfloat16, "A"); _want(B, torch.bfloat16, "B")
    M = A.shape[0] if M is None else M
    N, K = B.shape
    assert (A.shape[1] == K or epi.get("conv_cin")) and A.stride(1) == 1 and B.stride(1) == 1
    e = _epilogue(**epi)
    out = epi.get("out")
    rc = load().toc3d_gemm_bf16(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, kind,
                                ctypes.byref(e), _stream())
    _check(rc, "toc3d_gemm_bf16")
    return out


def window_attention(qkv, out, n_windows, seq_len, heads, out_map=None, q_rows=None, item_order=None, kv_rows=None, pad_v=None):
    """kv_rows (int32 [n_windows]) + pad_v (fp32 [C]): analytic pad keys of the dense blocks (include/toc3d_b200.h)."""
    _want(qkv, torch.bfloat16, "qkv"); _want(out, torch.bfloat16, "out")
    _check(load().toc3d_window_attention(_p(qkv), _p(out), n_windows, seq_len, heads, _p(out_map), _p(q_rows), _p(item_order),
                                         _p(kv_rows), _p(pad_v), _stream()),
           "toc3d_window_attention")
    return out


def layernorm_rows(x, gamma, beta, out, M, C, eps, row_map=None, alt=None, pad_mode=0, zero_stats=None):
    _want(x, torch.float32, "x"); _want(out, torch.bfloat16, "out")
    _check(load().toc3d_layernorm_rows(_p(x), _p(row_map), _p(alt), _p(gamma), _p(beta), _p(out), M, C, eps,
                                       pad_mode, _p(zero_stats), _stream()), "toc3d_layernorm_rows")
    return out


def subln(h, out, gamma, beta, M, Hd, ld, eps):
    _check(load().toc3d_subln_bf16(_p(h), _p(out), _p(gamma), _p(beta), M, Hd, ld, eps, _stream()),
           "toc3d_subln_bf16")
    return out


def window_topk(scores, V, H, W, ws, k, slow_idx=None, fast_idx=None, fast_score=None, tok_map=None,
                rope_rows=None, fast_map=None, fast_win=None):
    _want(scores, torch.float32, "scores")
    _check(load().toc3d_window_topk(_p(scores), V, H, W, ws, k, _p(slow_idx), _p(fast_idx), _p(fast_score),
                                    _p(tok_map), _p(rope_rows), _p(fast_map), _p(fast_win), _stream()), "toc3d_window_topk")


def topk_split(scores, B, N, k, keep_idx, drop_idx):
    _want(scores, torch.float32, "scores"); _want(keep_idx, torch.int64, "keep_idx")
    _check(load().toc3d_topk_split(_p(scores), B, N, k, _p(keep_idx), _p(drop_idx), _stream()), "toc3d_topk_split")


def merge_fast_tokens(x, fast_map, fast_score, nW, n_fast, k, C, rep_out, packed=None):
    _check(load().toc3d_merge_fast_tokens(_p(x), _p(fast_map), _p(fast_score), nW, n_fast, k, C, _p(rep_out),
                                          _p(packed), _stream()), "toc3d_merge_fast_tokens")


def compact_rows(tok_map, coff, rcap, nW, k, cmap, ctok, rep_row, rope_rows=None, cinv=None, crope=None, prope=None):
    _check(load().toc3d_compact_rows(_p(tok_map), _p(rope_rows), _p(coff), _p(rcap), nW, k, _p(cmap), _p(ctok), _p(rep_row),
                                     _p(cinv), _p(crope), _p(prope), _stream()), "toc3d_compact_rows")


def fill_pad_kv_rope(qkv, cmap, rope_rows, Mp, kpad, vpad, cos_axis, sin_axis, ft, C):
    _check(load().toc3d_fill_pad_kv_rope(_p(qkv), _p(cmap), _p(rope_rows), Mp, _p(kpad), _p(vpad), _p(cos_axis), _p(sin_axis),
                                         ft, C, _stream()), "toc3d_fill_pad_kv_rope")


def fill_pad_kv(qkv, pad_rows, v_bias, C):
    _check(load().toc3d_fill_pad_kv(_p(qkv), _p(pad_rows), pad_rows.numel(), _p(v_bias), C, _stream()), "toc3d_fill_pad_kv")


def fast_token_update(x, fast_map, packed, rep, nW, n_fast, k, C, rep_row=None):
    _check(load().toc3d_fast_token_update(_p(x), _p(fast_map), _p(packed), _p(rep), nW, n_fast, k, C, _p(rep_row), _stream()),
           "toc3d_fast_token_update")


def ln_gather_merge(x, tok_map, fast_map, fast_score, gamma, beta, out, rep_out, packed, nW, k, n_fast, C, eps,
                    zero_stats=None, rep_row=None, compact_rows=0, pad_fill=None, counters=None, pending=None):
    """pad_fill = (qkv, cmap, rope_rows, Mp, kpad, vpad, cos_axis, sin_axis, ft): also write the pad rows' k / v.
    pending = (fast_win, packed, rep_row, rep) of the previous accelerated block: apply its deferred fast-token update."""
    pu = None
    if pending is not None:
        fw, pk, rr_, rp = pending
        pu = PendingUpdate(_p(fw), _p(pk), _p(rr_), _p(rp))
    pf = None
    if pad_fill is not None:
        q, cm, rr, Mp, kp, vp, ca, sa, ft = pad_fill
        pf = PadFill(_p(q), _p(cm), _p(rr), Mp, _p(kp), _p(vp), _p(ca), _p(sa), ft)
    _want(x, torch.float32, "x"); _want(out, torch.bfloat16, "out")
    _check(load().toc3d_ln_gather_merge(_p(x), _p(tok_map), _p(fast_map), _p(fast_score), _p(gamma), _p(beta), _p(out),
                                        _p(rep_out), _p(packed), nW, k, n_fast, C, eps, _p(zero_stats), _p(rep_row), compact_rows,
                                        ctypes.byref(pf) if pf is not None else None, _p(counters),
                                        ctypes.byref(pu) if pu is not None else None, _stream()),
           "toc3d_ln_gather_merge")


def motion_blob_floats(Q, C):
    n = load().toc3d_motion_blob_floats(Q, C)
    if n <= 0:
        raise RuntimeError("toc3d_motion_blob_floats: bad shape Q=%d C=%d" % (Q, C))
    return n


def motion_queries_fold(blob, temp_queries, ref_points, vel, timestamp, ego_pose, ego_pose_inv, scale, C, q_out, A_out, c_out):
    """Motion-aware query encoder + scorer folding of all stages (toc3d_motion_queries_fold).  blob fp32 [S, stride]
    (layout: include/toc3d_b200.h); timestamp fp32 or fp64 [Bf, Q(, 1)]; outputs q_out [S,Bf,Q,256], A_out [S,Bf,2,C],
    c_out [S,Bf,2]."""
    for t, n in ((blob, "blob"), (temp_queries, "temp_queries"), (ref_points, "ref_points"), (vel, "vel"),
                 (ego_pose, "ego_pose"), (ego_pose_inv, "ego_pose_inv"), (q_out, "q_out"), (A_out, "A_out"), (c_out, "c_out")):
        _want(t, torch.float32, n)
    assert timestamp.dtype in (torch.float32, torch.float64), "timestamp must be fp32 or fp64"
    S, stride = blob.shape
    Bf, Q, D = temp_queries.shape
    assert D == 256 and tuple(q_out.shape) == (S, Bf, Q, 256) and tuple(A_out.shape) == (S, Bf, 2, C)
    assert timestamp.numel() == Bf * Q and ref_points.numel() == Bf * Q * 3 and vel.numel() == Bf * Q * 2
    assert ego_pose.numel() == Bf * Q * 16 and ego_pose_inv.numel() == Bf * 16 and c_out.numel() == S * Bf * 2
    rc = load().toc3d_motion_queries_fold(_p(blob), stride, S, Bf, Q, C, _p(temp_queries), _p(ref_points), _p(vel), _p(timestamp),
                                          int(timestamp.dtype == torch.float64), _p(ego_pose), _p(ego_pose_inv), scale,
                                          _p(q_out), _p(A_out), _p(c_out), _stream())
    _check(rc, "toc3d_motion_queries_fold")
    global launch_count
    launch_count += 1          # two kernels behind this entry point (encoder, fold)


def score_tokens(x, mask_in, A, c, V, N, C, views_per_frame, gumbel, seed, pred, score, mask_out, seed_dev=None):
    _check(load().toc3d_score_tokens(_p(x), _p(mask_in), _p(A), _p(c), V, N, C, views_per_frame, _p(gumbel), seed,
                                     _p(seed_dev), _p(pred), _p(score), _p(mask_out), _stream()),
           "toc3d_score_tokens")


def score_finish(logits, M, gumbel, seed, pred, score, mask_out, seed_dev=None):
    _check(load().toc3d_score_finish(_p(logits), M, _p(gumbel), seed, _p(seed_dev), _p(pred), _p(score),
                                     _p(mask_out), _stream()), "toc3d_score_finish")


def im2col_patch16(img, out, V, Hi, Wi):
    _want(img, torch.float32, "img")
    _check(load().toc3d_im2col_patch16(_p(img), _p(out), V, Hi, Wi, _stream()), "toc3d_im2col_patch16")


def preprocess_patch16_u8(img, lut, out, V, Hs, Ws, Hi, Wi, to_rgb):
    _want(img, torch.uint8, "img"); _want(lut, torch.float32, "lut"); _want(out, torch.bfloat16, "out")
    _check(load().toc3d_preprocess_patch16_u8(_p(img), _p(lut), _p(out), V, Hs, Ws, Hi, Wi, int(bool(to_rgb)), _stream()),
           "toc3d_preprocess_patch16_u8")


def cast_bf16(src, dst):
    _want(src, torch.float32, "src"); _want(dst, torch.bfloat16, "dst")
    _check(load().toc3d_cast_f32_to_bf16(_p(src), _p(dst), src.numel(), _stream()), "toc3d_cast_f32_to_bf16")
    return dst


def mask_rows(x, mask, out, M, C):
    _check(load().toc3d_mask_rows(_p(x), _p(mask), _p(out), M, C, _stream()), "toc3d_mask_rows")


def global_half_mean(y, V, N, C):
    _check(load().toc3d_global_half_mean(_p(y), V, N, C, _stream()), "toc3d_global_half_mean")
