/*
 * toc3d_b200 — C-ABI of the B200-native (sm_100a) ToC3D image-backbone hot path.
 *
 * Drop-in boundary: the reference (DYZhang09/ToC3D) is pure Python/PyTorch and has
 * no FFI of its own; every entry point below replaces the ATen call sequence of one
 * reference call site (cited per function, paths under
 * projects/mmdet3d_plugin/models/ of the reference tree).  The Python plugin
 * toc3d_b200/backbone.py (registry names ToC3DEVAViT / EVA_ViT) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted;
 *   - the caller owns all memory; the library allocates nothing that outlives a call;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), does
 *     not synchronise, and is re-entrant per stream (CUDA-graph capturable);
 *   - return value 0 = success; >0 = cudaError_t; <0 = argument error.  No C++
 *     exception crosses the boundary; toc3d_last_error() returns the message;
 *   - bf16 tensors are raw uint16 storage (nv_bfloat16); activations row-major.
 */
#ifndef TOC3D_B200_H_
#define TOC3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOC3D_B200_ABI_VERSION 14

int toc3d_abi_version(void);
/* Thread-local message of the last failing call ("" if none). Host pointer. */
const char* toc3d_last_error(void);

/* ------------------------------------------------------------------ GEMM (tcgen05 + TMA + TMEM)
 * C[M,N] = A[M,K] * B[N,K]^T with bf16 operands and fp32 accumulation in TMEM, followed by a
 * fused epilogue.  B has the nn.Linear weight layout ([out_features, in_features]).
 * Replaces F.linear / nn.Linear / nn.Conv2d(k=s=16) at: eva_vit.py:97-99,113 (q/k/v/proj),
 * eva_vit.py:45-49 (w1,w2,w3), eva_utils.py:284 (patch conv as im2col GEMM),
 * toc3d_utils.py:122,127 (first-frame scorer MLP).
 * Requirements: K % 8 == 0, lda % 8 == 0, ldb % 8 == 0, A/B 16-byte aligned.
 * Runs on CTA pairs (tcgen05 cta_group::2, 256 x tile_n pair tiles).  Like every kernel of the library it is
 * launched with programmatic stream serialization and waits for its stream predecessor on the device.
 */
enum toc3d_epilogue_kind {
  TOC3D_EPI_LINEAR = 0, /* out = act(acc + bias), bf16 or fp32                                   */
  TOC3D_EPI_QKV_ROPE = 1, /* bf16 out = [rope(acc+bias)*q_scale | rope(acc) | acc+bias]           */
  TOC3D_EPI_RESID = 2,  /* fp32 out[out_map[m]] = resid[resid_map[m]] + acc + bias               */
  TOC3D_EPI_SWIGLU = 3  /* bf16 out[:, h] = silu(acc1+b1) * (acc2+b2), B rows interleaved 32/32  */
};

typedef struct toc3d_epilogue {
  const float* bias;      /* [N] fp32 or NULL                                                   */
  void* out;              /* output base (bf16 or fp32, see kind)                               */
  int32_t ldo;            /* leading dimension of out, in elements                              */
  int32_t out_f32;        /* LINEAR: 1 -> fp32 output, 0 -> bf16                                */
  int32_t act;            /* LINEAR: 0 none, 1 GELU(erf), 2 ReLU                                */
  /* RESID: row maps are int32 per GEMM row m (NULL = identity).
   *   resid_map[m] >= 0 : residual row in `resid`;  -1 : zero residual;  -2 : out_alt row m
   *   out_map[m]   >= 0 : destination row in `out`; -1 : row dropped;    -2 : out_alt row m
   *   resid_mod > 0     : residual row = m % resid_mod (abs-pos broadcast over views)         */
  const float* resid;
  const int32_t* resid_map;
  int32_t resid_mod;
  const int32_t* out_map;
  float* out_alt;         /* fp32 [M, ldo] packed side buffer (may alias out)                   */
  /* QKV_ROPE: columns [0, rope_cols) are rotated (q then k), q columns [0, rope_cols/2) are
   * also multiplied by q_scale.  Head dim is 64: 32 row-axis + 32 column-axis channels.
   *   rope_rows[m]  table row of GEMM row m (NULL -> m % rope_slots)
   *   cos_axis/sin_axis  fp32 [rope_ft, 16]: per-axis position x frequency tables
   *                       (= freqs_cos[p * rope_ft, 0:32:2] of eva_utils.py:364-371)           */
  const int32_t* rope_rows;
  int32_t rope_slots;
  int32_t rope_ft;
  int32_t rope_cols;
  float q_scale;
  const float* cos_axis;
  const float* sin_axis;
  /* Folded SwiGLU sub-LayerNorm (ffn_ln, eva_vit.py:48: no separate normalisation pass; statistics are int64 fixed
   * point [sum * 2^30, sum of squares * 2^26] per row, accumulated with integer atomics = order-independent, hence
   * deterministic; int64 [rows,2], 16-byte aligned).
   *   OUTPUT statistics (SWIGLU, row_stats != NULL): of the bf16-rounded hidden row m;
   *   INPUT statistics (RESID, ln_stats != NULL): the A rows are UN-normalised, B is W * gamma (column-scaled) and
   *     the epilogue applies   y = rstd_m * acc - rstd_m * mean_m * ln_u[n] + bias[n]
   *     with mean/rstd of row m from ln_stats over ln_n true columns (eps = ln_eps), ln_u = W * gamma (fp32 [N]),
   *     bias = W * beta + b.
   * (The same fold for norm2 - proj epilogue emitting bf16 rows + statistics, SwiGLU epilogue applying them - was
   * built, tested and measured twice on B200: 4.29 vs 4.03 ms per forward, slower; removed.) */
  int64_t* row_stats;
  const int64_t* ln_stats;
  const float* ln_u;
  int32_t ln_n;
  float ln_eps;
  /* Tile width override (tuning / tests): 0 = chosen per launch to minimise wave quantisation on the
   * 74 CTA pairs; else a multiple of 32 (64 for SWIGLU) in [64, 256]. */
  int32_t tile_n;
  /* Implicit 3 x 3 convolution (LINEAR epilogue; necks/cp_fpn.py:123-133 without a materialised im2col): conv_cin > 0
   * makes A a [M, conv_cin] matrix of pixels in a spatially ZERO-PADDED NHWC layout (rows of (H+2) x (W+2) pixels per
   * image, border pixels zero), K = 9 * conv_cin with the weight columns ordered (ky, kx, ci), and the k-blocks of tap
   * t = 3 ky + kx read the A rows m + conv_row_shift[t] (= (ky-1) * (W+2) + (kx-1)) by TMA - the shifted rows of a
   * padded layout are again contiguous.  Rows outside A read zeros.  Border rows of the output are garbage: drop them
   * with out_map = -1.  conv_cin % 64 == 0. */
  int32_t conv_cin;
  int32_t conv_row_shift[9];
} toc3d_epilogue;

int toc3d_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N, int32_t K,
                    int32_t epilogue_kind, const toc3d_epilogue* epi /* host */, void* stream);

/* ------------------------------------------------------------------ windowed attention
 * softmax(q k^T) v per (window, head); q already rotated and scaled by the QKV epilogue.
 * qkv bf16 [n_windows*seq_len, 3*C] (q | k | v), out bf16 [n_windows*seq_len, C]; head dim 64.
 * Replaces eva_vit.py:109-111 and toc3d_eva_vit.py:509-511 (bmm, softmax, bmm). seq_len <= 1024.
 * out_map (optional, int32 [n_windows*seq_len]): destination row of each query row in `out`, -1 = row not
 * stored.  Rows that are window PADDING matter only as keys / values; with out_map the attention writes the rows
 * that are used afterwards in compact order, so that the row-wise GEMMs after it skip the padding.
 * q_rows (optional, int32 [n_windows]): only the first q_rows[w] query rows of window w are needed (the callers
 * put the needed rows first); query tiles beyond them are not computed.  All seq_len rows still act as keys.
 * item_order (optional, int32 [n_windows*heads], a permutation of the (window*heads + head) items): processing
 * order of the persistent kernel, e.g. sorted by query-tile count so that its round-robin deal is balanced.
 * kv_rows + pad_v (optional, together; seq_len <= 448): ANALYTIC pad keys of the dense blocks.  The reference zero-pads
 * the normalised map to whole windows (eva_vit.py:249-254), so a pad slot has k = 0 exactly (k_proj has no bias,
 * RoPE keeps zero) and v = v_bias: its score is 0 for every query and its value is one common vector.  With
 * kv_rows[w] (int32 [n_windows]) = number of REAL keys of window w (stored first) and pad_v (fp32 [C]) = v_bias, only the
 * real keys are staged and multiplied; the seq_len - kv_rows[w] pad keys enter the softmax as one closed-form term
 * (row max includes 0, row sum += n_pad * exp(0 - max), O += n_pad * exp(0 - max) * pad_v).  The pad rows of qkv are
 * then never read, so nobody has to write them.
 * The index tables q_rows, item_order and kv_rows are launch parameters in device memory, not data of the stream: the
 * kernel reads them before its grid dependency on the preceding kernel resolves (programmatic dependent launch), so they
 * must not be written by the kernel launched immediately before this call on the same stream (copies and earlier
 * kernels are fine).  qkv, out_map and pad_v have no such restriction.
 */
int toc3d_window_attention(const void* qkv, void* out, int32_t n_windows, int32_t seq_len, int32_t heads,
                           const int32_t* out_map, const int32_t* q_rows, const int32_t* item_order,
                           const int32_t* kv_rows, const float* pad_v, void* stream);

/* ------------------------------------------------------------------ LayerNorm over gathered rows
 * out_bf16[m] = LN(row(m)) * gamma + beta over C channels (C % 128 == 0, C <= 4096), m in [0,M).
 *   row_map NULL: row(m) = x[m];  row_map[m] >= 0: x[row_map[m]];  -2: alt[m];
 *   -1: pad slot -> pad_mode 0: output zeros (dense Block pads AFTER norm1, eva_vit.py:249-254)
 *                   pad_mode 1: LN of a zero vector = beta (ToC3D block pads BEFORE norm1,
 *                               toc3d_eva_vit.py:412-415 then :369)
 * zero_stats (optional, int64 [M,2]): rows are zeroed as a side effect -- the accumulator the SwiGLU
 * GEMM epilogue adds its folded sub-LN statistics to (toc3d_epilogue.row_stats).
 * Replaces nn.LayerNorm at eva_vit.py:249,263; toc3d_eva_vit.py:371,379; toc3d_utils.py:99.
 */
int toc3d_layernorm_rows(const float* x, const int32_t* row_map, const float* alt, const float* gamma,
                         const float* beta, void* out_bf16, int32_t M, int32_t C, float eps, int32_t pad_mode,
                         int64_t* zero_stats, void* stream);

/* SwiGLU sub-LayerNorm (eva_vit.py:48, ffn_ln over the true hidden width `Hd`) on the padded bf16
 * hidden buffer [M, ld].  Contract: columns >= Hd of h are exactly zero on input and gamma/beta are
 * zero-padded to ld; they are written as zero.  In place allowed (out == h). */
int toc3d_subln_bf16(const void* h, void* out, const float* gamma, const float* beta, int32_t M, int32_t Hd,
                     int32_t ld, float eps, void* stream);

/* ------------------------------------------------------------------ token selection
 * Per-window stable top-k (toc3d_eva_vit.py:412-419 + toc3d_utils.py:131-144 with the
 * tie-break pin "score descending, index ascending"): scores fp32 [V,H,W] are window-partitioned
 * with pad value -1e6; for every window the ws*ws slots are ranked and split into the first
 * k (slow) and the rest (fast), both in rank order.
 * Outputs (any may be NULL): slow_idx int32 [nW,k], fast_idx int32 [nW,n-k] (slot indices, the
 * values torch.sort returns), fast_score fp32 [nW,n-k],
 *   tok_map  int32 [nW*(k+1)] packed row -> image row v*H*W+r*W+c | -1 (pad slot) | -2 (the
 *            representative token, last row of each window),
 *   rope_rows int32 [nW*(k+1)] RoPE table row (slot index; k for the representative token,
 *            toc3d_eva_vit.py:434-435),
 *   fast_map int32 [nW,n-k] image row of each fast token or -1,
 *   fast_win int32 [V*H*W] image row -> window in which it is a FAST token | -1 (slow token); every entry is written.
 * nW = V*ceil(H/ws)*ceil(W/ws); window order view-major, row, col (eva_utils.py:108-109).
 */
int toc3d_window_topk(const float* scores, int32_t V, int32_t H, int32_t W, int32_t ws, int32_t k,
                      int32_t* slow_idx, int32_t* fast_idx, float* fast_score, int32_t* tok_map,
                      int32_t* rope_rows, int32_t* fast_map, int32_t* fast_win, void* stream);

/* Compact row space of an accelerated block (toc3d_eva_vit.py:421-461): of the selected rows [k slow | rep] of a
 * window, the slow rows that are pad slots (tok_map = -1) are needed as attention keys / values only - norm1, q/k/v,
 * proj, norm2 and the SwiGLU MLP are row-wise and window_unpartition crops the pads' results - so those run on the
 * compact rows (real slow rows + rep).  coff / rcap int32 [nW] are host-static (rcap = min(k, #real tokens of the
 * window), coff = exclusive prefix sum of rcap + 1).  The window-packed qkv buffer the attention reads is laid out
 * [real slow rows | rep | pad rows] per window ("packed position"), so the needed query rows are a prefix.
 * Inputs tok_map / rope_rows int32 [nW*(k+1)] in rank order (toc3d_window_topk).  Outputs:
 *   cmap  int32 [nW*(k+1)]  packed position -> compact row | -1 (pad)      prope  same shape: RoPE table row (NULL ok)
 *   ctok  int32 [sum(rcap+1)]  compact -> image row | -2 (representative) | -1 (unused)
 *   cinv / crope (NULL ok)     compact -> packed position (-1 unused) / RoPE table row
 *   rep_row int32 [nW]         compact row of the representative. */
int toc3d_compact_rows(const int32_t* tok_map, const int32_t* rope_rows, const int32_t* coff, const int32_t* rcap,
                       int32_t nW, int32_t k, int32_t* cmap, int32_t* ctok, int32_t* rep_row, int32_t* cinv,
                       int32_t* crope, int32_t* prope, void* stream);

/* Accelerated blocks pad BEFORE norm1 (toc3d_eva_vit.py:412-415 then :369), so a pad slot selected as a slow token
 * is the vector norm1(0) = beta: key = RoPE(W_k beta, slot), value = W_v beta + v_bias.  kpad / vpad fp32 [C] are
 * those block constants before the rotation; every packed qkv row m with cmap[m] == -1 gets them (slot =
 * rope_rows[m], per-axis tables as in toc3d_epilogue).  q/k/v of the real rows then come from a GEMM over the
 * compact rows only. */
int toc3d_fill_pad_kv_rope(void* qkv_bf16, const int32_t* cmap, const int32_t* rope_rows, int32_t Mp, const float* kpad,
                           const float* vpad, const float* cos_axis, const float* sin_axis, int32_t ft, int32_t C,
                           void* stream);

/* Dense blocks (eva_vit.py:247-268) pad AFTER norm1, so pad slots are exact zeros: k = 0 (k_proj has no bias, RoPE
 * keeps zero), v = v_bias.  Writes those constants into the listed slot rows of the bf16 qkv buffer [.., 3C]
 * (pad_rows int32 [n_pad]); the QKV GEMM then only runs over real tokens. */
int toc3d_fill_pad_kv(void* qkv_bf16, const int32_t* pad_rows, int32_t n_pad, const float* v_bias, int32_t C,
                      void* stream);

/* Image-level stable descending sort split (toc3d_utils.py:137-144): scores fp32 [B,N] ->
 * keep_idx int64 [B,k], drop_idx int64 [B,N-k].  N <= 12288. */
int toc3d_topk_split(const float* scores, int32_t B, int32_t N, int32_t k, int64_t* keep_idx, int64_t* drop_idx,
                     void* stream);

/* Representative token (toc3d_utils.py:65-70 merge_tokens on the gathered fast set,
 * toc3d_eva_vit.py:424-427): rep[w] = sum_j (s_j / sum s) * x[fast_map[w,j]] (pad slots are zero
 * vectors but their scores count).  Written to rep_out fp32 [nW,C] and to
 * packed[(w*(k+1)+k)*C] when packed != NULL. */
int toc3d_merge_fast_tokens(const float* x, const int32_t* fast_map, const float* fast_score, int32_t nW,
                            int32_t n_fast, int32_t k, int32_t C, float* rep_out, float* packed, void* stream);

/* Fast-token update (toc3d_eva_vit.py:452-456): x[fast_map[w,j]] += packed[(w*(k+1)+k)] - rep[w],
 * i.e. the representative token's attention + MLP residual deltas. */
int toc3d_fast_token_update(float* x, const int32_t* fast_map, const float* packed, const float* rep, int32_t nW,
                            int32_t n_fast, int32_t k, int32_t C, const int32_t* rep_row /* NULL: w*(k+1)+k */,
                            void* stream);

/* Arguments of toc3d_fill_pad_kv_rope as a struct, so that toc3d_ln_gather_merge can do the same work in extra thread
 * blocks of its own launch (pad_fill != NULL). */
typedef struct toc3d_pad_fill {
  void* qkv;                 /* bf16 [Mp, 3C] packed qkv buffer */
  const int32_t* cmap;       /* [Mp] packed row -> compact row | -1 (pad: gets filled) */
  const int32_t* rope_rows;  /* [Mp] RoPE table row of each packed row */
  int32_t Mp;
  const float* kpad;         /* fp32 [C] W_k beta (before rotation) */
  const float* vpad;         /* fp32 [C] W_v beta + v_bias */
  const float* cos_axis;     /* fp32 [ft,16] */
  const float* sin_axis;
  int32_t ft;
} toc3d_pad_fill;

/* Deferred fast-token update (toc3d_eva_vit.py:452-461) of the PREVIOUS accelerated block, applied by
 * toc3d_ln_gather_merge while it reads the rows anyway (every real row is read exactly once: as a slow row or as a fast
 * row of the new block): x[r] += packed[rep_row[w]] - rep[w] for w = fast_win[r] >= 0, written back to x.  Same
 * expression as toc3d_fast_token_update, so the result is bit-identical to calling that in between.  `packed` / `rep`
 * must not be the buffers the same launch writes (ping-pong them between consecutive blocks). */
typedef struct toc3d_pending_update {
  const int32_t* fast_win;   /* [rows of x] previous block: window in which the row was a fast token | -1 (toc3d_window_topk) */
  const float* packed;       /* previous block's packed / compact rows (representative AFTER the block) */
  const int32_t* rep_row;    /* previous block: row of `packed` of window w's representative */
  const float* rep;          /* previous block: representative BEFORE the block, fp32 [nW_prev, C] */
} toc3d_pending_update;

/* Fused front end of an accelerated block (toc3d_eva_vit.py:421-427 gather + merge_tokens, then norm1 at
 * :371): in ONE launch, (1) rep[w] = sum_j (s_j / sum s) x[fast_map[w,j]] -> rep_out[w] and packed row
 * w*(k+1)+k (fp32), LayerNorm(rep[w]) -> out row w*(k+1)+k; (2) LayerNorm of every gathered slow row
 * m (tok_map[m] >= 0: x row; -1: pad slot = zero vector -> beta; -2: the representative row, see (1)) ->
 * bf16 out [nW*(k+1), C].  C in {128, 256, 512, 1024}.  zero_stats as in toc3d_layernorm_rows. */
int toc3d_ln_gather_merge(float* x /* written only with a pending update */, const int32_t* tok_map, const int32_t* fast_map, const float* fast_score,
                          const float* gamma, const float* beta, void* out_bf16, float* rep_out, float* packed,
                          int32_t nW, int32_t k, int32_t n_fast, int32_t C, float eps, int64_t* zero_stats,
                          const int32_t* rep_row /* row of `packed` for the representative; NULL: w*(k+1)+k */,
                          int32_t compact_rows /* > 0: tok_map lists that many COMPACT rows (ctok; -1 skipped) and the
                                                  LayerNorm output uses compact rows (rep at rep_row[w]); 0: packed */,
                          const toc3d_pad_fill* pad_fill /* host pointer or NULL */,
                          int32_t* counters /* int32 [nW], zeroed once by the caller: the representative token of a window
                                               is merged by C/256 thread blocks (256-channel slices), the last one to
                                               arrive normalises the row and clears the counter; NULL allowed for C = 128 */,
                          const toc3d_pending_update* pending /* host pointer or NULL */,
                          void* stream);

/* ------------------------------------------------------------------ history-query scorer
 * Motion-aware query encoder + scorer folding for ALL S selector stages of one forward (row a12 + a13), two launches.
 * Replaces MotionAwareQueryGuidedTokenSelector.get_motion_aware_queries (toc3d_utils.py:334-360) with its helpers
 * transform_reference_points (misc.py:191-200), MLN (misc.py:154-188), pos2posemb3d / pos2posemb1d /
 * nerf_positional_encoding (positional_encoding.py:14-81), fp32 throughout (the timestamp embedding in fp64 when the
 * timestamps are fp64, as in the reference), and folds the result: toc3d_utils.py:232-252 is linear in the token up
 * to the LogSoftmax, so per (stage, frame)
 *   A = scale * W_agg (2xQ) * q (Qx256) * W_in (256xC),   c = scale * W_agg * q * b_in + b_agg.
 * blob: fp32 [S, blob_stride] per-stage parameters, packed by the caller in this order (Linear weights TRANSPOSED to
 *   [in][out]; D = 256):  dimt128[128] dimt256[256] (temperature ** (2 * (i // 2) / F), positional_encoding.py:17,31)
 *   pc_range[8: 6 used]  query_embedding.0 wT[384][D] b[D]  query_embedding.2 wT[D][D] b[D]
 *   ego_pose_pe:      reduce.0 wT[180][D] b[D]  gamma wT[D][D] b[D]  beta wT[D][D] b[D]
 *   ego_pose_queries: reduce.0 wT[180][D] b[D]  gamma wT[D][D] b[D]  beta wT[D][D] b[D]
 *   time_embedding.0 wT[D][D] b[D]  time_embedding.1 weight[D] bias[D]
 *   input_proj.0 weight[D][C] (NOT transposed) bias[D]  aggregate.0 weight[2][Q] bias[4: 2 used];
 *   toc3d_motion_blob_floats(Q, C) returns the number of floats (host-only helper, no CUDA call).
 * temp_queries fp32 [Bf,Q,256]; ref_points [Bf,Q,3]; vel [Bf,Q,2]; timestamp [Bf,Q] fp32 or fp64 (timestamp_is_f64);
 * ego_pose [Bf,Q,4,4]; ego_pose_inv [Bf,4,4].  Outputs: q_out [S,Bf,Q,256] (the encoded queries), A_out [S,Bf,2,C],
 * c_out [S,Bf,2].  Q even, C % 4 == 0. */
int64_t toc3d_motion_blob_floats(int32_t Q, int32_t C);
int toc3d_motion_queries_fold(const float* blob, int64_t blob_stride, int32_t S, int32_t Bf, int32_t Q, int32_t C,
                              const float* temp_queries, const float* ref_points, const float* vel, const void* timestamp,
                              int32_t timestamp_is_f64, const float* ego_pose, const float* ego_pose_inv, float scale,
                              float* q_out, float* A_out, float* c_out, void* stream);

/* Per token: logit = mask_in * (x . A[f]) + c[f]; pred = log_softmax(logit) (fp32 [V*N,2]);
 * mask_out = softmax(pred + g)[0] (toc3d_utils.py:147, pin 2) with g = gumbel [V*N,2] or, when
 * gumbel == NULL, -log(-log(u)) drawn on device from (seed, token index); when seed_dev != NULL the
 * effective seed is seed + 1000003 * *seed_dev (a device-resident call counter, so that a captured
 * CUDA graph draws fresh noise on every replay).
 * mask_in NULL = all ones.  views_per_frame = V / Bf (repeat_interleave, toc3d_utils.py:240). */
int toc3d_score_tokens(const float* x, const float* mask_in, const float* A, const float* c, int32_t V, int32_t N,
                       int32_t C, int32_t views_per_frame, const float* gumbel, uint64_t seed,
                       const uint64_t* seed_dev, float* pred, float* score, float* mask_out, void* stream);

/* Same tail for the first-frame scorer (toc3d_utils.py:114-129) whose logits come from its four MLP
 * GEMMs: logits fp32 [M,2] -> pred, score, mask_out. */
int toc3d_score_finish(const float* logits, int32_t M, const float* gumbel, uint64_t seed,
                       const uint64_t* seed_dev, float* pred, float* score, float* mask_out, void* stream);

/* ------------------------------------------------------------------ stem
 * im2col for the 16x16/stride-16 patch conv (eva_utils.py:283-287): img fp32 NCHW [V,3,Hi,Wi] ->
 * bf16 [V*(Hi/16)*(Wi/16), 768], column = c*256 + ky*16 + kx (= conv weight flattening). */
int toc3d_im2col_patch16(const float* img, void* out_bf16, int32_t V, int32_t Hi, int32_t Wi, void* stream);

/* Next row f3: the step before the path.  NormalizeMultiviewImage + PadMultiViewImage
 * (datasets/pipelines/transform_3d.py:21-104, i.e. mmcv.imnormalize + impad_to_multiple; config
 * ToC3D_fast.py:13-14,209-210) fused with the patch im2col above, so the host uploads the u8 camera crop
 * (4x fewer bytes than the normalised fp32 NCHW image) and no fp32 image is ever materialised.
 * img u8 HWC [V,Hs,Ws,3]; lut fp32 [3,256]: lut[c*256+b] = normalised value of byte b in output channel c
 * (the host builds it in cv2's arithmetic: fp32(fp64(fp32(b - mean_c)) * (1/fp64(std_c)))); output channel c
 * reads input channel (to_rgb ? 2-c : c); Hi >= Hs, Wi >= Ws are the padded dims (multiples of 16), pad = 0.
 * Output identical (bit-exact) to toc3d_im2col_patch16 of the normalised, padded fp32 NCHW image. */
int toc3d_preprocess_patch16_u8(const uint8_t* img, const float* lut, void* out_bf16, int32_t V, int32_t Hs,
                                int32_t Ws, int32_t Hi, int32_t Wi, int32_t to_rgb, void* stream);

/* Neck (next row: CPFPN, necks/cp_fpn.py:157-208): the 1x1 lateral conv (cp_fpn.py:114-122) is a toc3d_gemm_bf16 call
 * (LINEAR) whose out_map scatters the rows into a zero-padded NHWC layout, the 3x3 fpn conv (cp_fpn.py:123-133,182-184)
 * a toc3d_gemm_bf16 call in implicit-convolution mode over that layout (toc3d_epilogue.conv_*); no im2col buffer.
 *
 * Row-wise helpers used by the neck, the first-frame scorer and weight repacking. */
int toc3d_cast_f32_to_bf16(const float* in, void* out_bf16, int64_t n, void* stream);
/* x[m,:] * mask[m] -> LN -> bf16 is toc3d_layernorm_rows on a pre-masked buffer; this masks. */
int toc3d_mask_rows(const float* x, const float* mask, float* out, int32_t M, int32_t C, void* stream);
/* Half-channel token mean (toc3d_utils.py:124-126): y bf16 [V,N,C] -> y[:, :, C/2:] = mean_n y[:, n, C/2:]. */
int toc3d_global_half_mean(void* y_bf16, int32_t V, int32_t N, int32_t C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOC3D_B200_H_ */
