// HBM-bound token kernels of the ToC3D path: LayerNorm over gathered rows, SwiGLU sub-LN, per-window
// stable top-k, representative-token merge, fast-token update, the folded history-query scorer,
// im2col for the patch stem.  Warp-per-row, 16-byte vector accesses, fp32 math.
// Reference call sites per function: include/toc3d_b200.h.
#include "common.cuh"
#include "../../include/toc3d_b200.h"

#include <stdarg.h>
#include <string.h>

namespace toc3d {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr float PAD_SCORE = -1e6f;  // toc3d_eva_vit.py:415

// ------------------------------------------------------------------------------------------------
// LayerNorm over gathered fp32 rows -> bf16.  One warp per row, C/128 float4 per lane.
template <int VPL>  // float4 per lane: C = VPL * 128
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, const int* __restrict__ row_map, const float* __restrict__ alt,
                      const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                      int M, float eps, int pad_mode, long long* __restrict__ zero_stats) {
  pdl_wait();
  pdl_launch_dependents();
  const int C = VPL * 128;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  if (zero_stats != nullptr && lane == 0) *reinterpret_cast<longlong2*>(zero_stats + 2 * (size_t)m) = make_longlong2(0, 0);
  int src = row_map ? row_map[m] : m;
  const float* p = nullptr;
  if (src >= 0) p = x + (size_t)src * C;
  else if (src == -2) p = alt + (size_t)m * C;
  uint2* o = reinterpret_cast<uint2*>(out + (size_t)m * C);
  if (p == nullptr && pad_mode == 0) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) o[lane + 32 * i] = make_uint2(0u, 0u);
    return;
  }
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = p ? reinterpret_cast<const float4*>(p)[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
    const float4 b = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
    uint2 u;
    u.x = pack_bf16((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
    u.y = pack_bf16((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    o[lane + 32 * i] = u;
  }
}

// SwiGLU sub-LN over bf16 hidden rows [M, ld], true width Hd (<= ld, ld % 8 == 0).  Contract: the pad
// columns [Hd, ld) of h are exactly zero (they are: zero weight rows and biases in the SwiGLU GEMM) and
// gamma/beta are zero there, so sums run unmasked over ld and the variance is corrected analytically.
__global__ void __launch_bounds__(256)
subln_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out, const float* __restrict__ gamma,
             const float* __restrict__ beta, int M, int Hd, int ld, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const uint4* p = reinterpret_cast<const uint4*>(h + (size_t)m * ld);
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)m * ld);
  const int nvec = ld >> 3;
  constexpr int MAXV = 12;  // supports ld <= 3072
  float x[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      const uint4 v = p[j];
      const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
      x[i][0] = a.x; x[i][1] = a.y; x[i][2] = b.x; x[i][3] = b.y;
      x[i][4] = c.x; x[i][5] = c.y; x[i][6] = d.x; x[i][7] = d.y;
      s += ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
    }
  }
  const float mean = warp_sum(s) / (float)Hd;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = x[i][e] - mean;
        q += t * t;
      }
    }
  }
  q = warp_sum(q) - (float)(ld - Hd) * mean * mean;      // pad columns each contributed mean^2
  const float rstd = rsqrtf(fmaxf(q, 0.f) / (float)Hd + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * j);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * j + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * j);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * j + 1);
      uint4 r;
      r.x = pack_bf16((x[i][0] - mean) * rstd * g0.x + b0.x, (x[i][1] - mean) * rstd * g0.y + b0.y);
      r.y = pack_bf16((x[i][2] - mean) * rstd * g0.z + b0.z, (x[i][3] - mean) * rstd * g0.w + b0.w);
      r.z = pack_bf16((x[i][4] - mean) * rstd * g1.x + b1.x, (x[i][5] - mean) * rstd * g1.y + b1.y);
      r.w = pack_bf16((x[i][6] - mean) * rstd * g1.z + b1.z, (x[i][7] - mean) * rstd * g1.w + b1.w);
      o[j] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Total order of torch.sort(descending=True, stable=True): NaN first, then larger values, ties
// (incl. -0.0 == +0.0) by lower index.  order_key maps a float to a uint32 that is larger for
// elements that rank earlier, so "a ranks before b" == key_a > key_b || (key_a == key_b && ia < ib).
__device__ __forceinline__ uint32_t order_key(float a) {
  if (a != a) return 0xFFFFFFFFu;
  if (a == 0.0f) a = 0.0f;                       // -0.0 ties with +0.0
  const uint32_t b = __float_as_uint(a);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// rank of element i (key ki) among keys[0, n): number of elements ranking strictly before it.
// n4 = ceil(n / 4) uint4 groups; slots in [n, 4*n4) hold key 0, which never ranks before a real element
// (order_key >= 0x007FFFFF for every float, -inf included).
__device__ __forceinline__ int rank_of(const uint32_t* __restrict__ keys, int n4, uint32_t ki, int i) {
  int rank = 0;
  const uint4* k4 = reinterpret_cast<const uint4*>(keys);
#pragma unroll 4
  for (int g = 0; g < n4; ++g) {
    const uint4 kj = k4[g];
    const int j = 4 * g;
    rank += (kj.x > ki) || (kj.x == ki && j + 0 < i);
    rank += (kj.y > ki) || (kj.y == ki && j + 1 < i);
    rank += (kj.z > ki) || (kj.z == ki && j + 2 < i);
    rank += (kj.w > ki) || (kj.w == ki && j + 3 < i);
  }
  return rank;
}

// One CTA per window, one thread per slot (n = ws*ws <= 1024).
__global__ void window_topk_kernel(const float* __restrict__ scores, int V, int H, int W, int ws, int k,
                                   int* __restrict__ slow_idx, int* __restrict__ fast_idx,
                                   float* __restrict__ fast_score, int* __restrict__ tok_map,
                                   int* __restrict__ rope_rows, int* __restrict__ fast_map, int* __restrict__ fast_win) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) uint32_t s_key[1024];
  const int n = ws * ws;
  const int nWw = (W + ws - 1) / ws, nWh = (H + ws - 1) / ws;
  const int w = blockIdx.x;
  const int v = w / (nWh * nWw);
  const int wr = (w / nWw) % nWh, wc = w % nWw;
  const int i = threadIdx.x;
  int img_row = -1;
  float sc = PAD_SCORE;
  if (i < n) {
    const int r = wr * ws + i / ws, c = wc * ws + i % ws;
    if (r < H && c < W) {
      img_row = (v * H + r) * W + c;
      sc = scores[img_row];
    }
  }
  const uint32_t ki = order_key(sc);
  s_key[i] = i < n ? ki : 0u;                    // blockDim.x = n rounded up to 32 (a multiple of 4)
  __syncthreads();
  if (i >= n) return;
  const int rank = rank_of(s_key, (n + 3) >> 2, ki, i);
  const int nf = n - k;
  if (rank < k) {
    if (slow_idx) slow_idx[(size_t)w * k + rank] = i;
    const size_t m = (size_t)w * (k + 1) + rank;
    if (tok_map) tok_map[m] = img_row;
    if (rope_rows) rope_rows[m] = i;
  } else {
    const size_t f = (size_t)w * nf + (rank - k);
    if (fast_idx) fast_idx[f] = i;
    if (fast_score) fast_score[f] = sc;
    if (fast_map) fast_map[f] = img_row;
  }
  // image row -> window whose representative token carries this row through the block (fast token), -1 for slow rows;
  // every real token belongs to exactly one window, so the whole [V*H*W] table is rewritten by each launch
  if (fast_win && img_row >= 0) fast_win[img_row] = rank < k ? -1 : w;
  if (i == 0) {
    const size_t m = (size_t)w * (k + 1) + k;
    if (tok_map) tok_map[m] = -2;
    if (rope_rows) rope_rows[m] = k;
  }
}

// Compact row space of an accelerated block.  The selected set of a window is [k slow rows in rank order | rep];
// slow rows that are PAD slots (tok_map = -1) matter only as attention keys / values: everything after the
// attention (proj, norm2, SwiGLU) is row-wise and their results are cropped by window_unpartition
// (toc3d_eva_vit.py:459-461), so those GEMMs (and norm1 / q,k,v) run on the compact rows = real slow rows + rep.
// coff[w] / rcap[w] (host-static: rcap = min(k, #real tokens of the window)) give the window's compact range.
// The window-packed qkv buffer the attention reads is laid out [real slow rows | rep | pad rows] per window, so the
// query rows that are needed form a prefix and whole query tiles of padding can be skipped (keys are order-free).
//   cmap[w*(k+1)+p]  = compact row of packed position p (or -1 for a pad)     prope[..] = RoPE table row of p
//   ctok[c] = image row | -2 (rep) | -1 (unused)        cinv[c] = packed position        crope[c] = RoPE row
//   rep_row[w] = compact row of the representative token.                     One CTA per window.
__global__ void __launch_bounds__(1024)
compact_rows_kernel(const int* __restrict__ tok_map, const int* __restrict__ rope_rows, const int* __restrict__ coff,
                    const int* __restrict__ rcap, int k, int* __restrict__ cmap, int* __restrict__ ctok,
                    int* __restrict__ rep_row, int* __restrict__ cinv, int* __restrict__ crope, int* __restrict__ prope) {
  __shared__ int s_warp[32];
  pdl_wait();
  pdl_launch_dependents();
  const int w = blockIdx.x, t = threadIdx.x;
  const int lane = t & 31, wid = t >> 5;
  const size_t base = (size_t)w * (k + 1);
  const int src = t < k ? tok_map[base + t] : -1;
  const int rope = (t < k && rope_rows != nullptr) ? rope_rows[base + t] : 0;
  const bool real = t < k && src >= 0;
  const unsigned bal = __ballot_sync(0xffffffffu, real);
  if (lane == 0) s_warp[wid] = __popc(bal);
  __syncthreads();
  int before = __popc(bal & ((1u << lane) - 1u));          // real rows ranking before this one
  int total = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
    const int c = s_warp[i];
    if (i < wid) before += c;
    total += c;
  }
  const int c0 = coff[w], cap = rcap[w];
  if (t < k) {
    const bool ok = real && before < cap;
    const int pos = ok ? before : total + 1 + (t - before);      // pads (and overflow rows) go after the rep
    cmap[base + pos] = ok ? c0 + before : -1;
    if (prope) prope[base + pos] = rope;
    if (ok) {
      ctok[c0 + before] = src;
      if (cinv) cinv[c0 + before] = (int)(base + pos);
      if (crope) crope[c0 + before] = rope;
    }
  }
  if (t >= total && t < cap) {                             // degenerate: fewer real slow rows than the static capacity
    ctok[c0 + t] = -1;
    if (cinv) cinv[c0 + t] = -1;
    if (crope) crope[c0 + t] = 0;
  }
  if (t == 0) {
    const int rrope = rope_rows != nullptr ? rope_rows[base + k] : k;
    cmap[base + total] = c0 + cap;
    if (prope) prope[base + total] = rrope;
    ctok[c0 + cap] = -2;
    rep_row[w] = c0 + cap;
    if (cinv) cinv[c0 + cap] = (int)(base + total);
    if (crope) crope[c0 + cap] = rrope;
  }
}

// Dense blocks compute q/k/v only for real tokens; the window slots that are padding hold constants:
// k = 0 (k_proj has no bias and the padded norm1 output is zero, eva_vit.py:249-254,97-99; RoPE keeps zero) and
// v = v_bias.  rows: slot rows of the qkv buffer [.., 3C]; thread = 8 channels.
// Accelerated blocks pad BEFORE norm1 (toc3d_eva_vit.py:412-415), so a pad slot that was selected as a slow token
// is the vector norm1(0) = beta: its key is RoPE(W_k beta, slot), its value W_v beta + v_bias - block constants
// (kpad / vpad, fp32 [C]) up to the rotation.  Fills the packed qkv rows m with cmap[m] == -1; thread = 8 channels.
// Pad rows of an accelerated block as attention keys / values (see toc3d_fill_pad_kv_rope).
struct PadFill {
  __nv_bfloat16* qkv;          // packed qkv buffer [Mp, 3C]; nullptr = no fill work
  const int* cmap;             // packed row -> compact row | -1 (pad)
  const int* rope_rows;        // packed row -> RoPE table row (window slot)
  int Mp;
  const float* kpad;           // fp32 [C] W_k beta
  const float* vpad;           // fp32 [C] W_v beta + v_bias
  const float* cos_axis;
  const float* sin_axis;
  int ft;
};

// one warp per packed row: real rows leave at once, a pad row writes its 2 * C bf16 (k then v), 8 channels per lane-step
__device__ __forceinline__ void fill_pad_row(const PadFill& f, int m, int lane, int C) {
  if (m >= f.Mp || f.cmap[m] != -1) return;
  const int slot = f.rope_rows[m];
  const int r = slot / f.ft, c = slot - r * f.ft;
  const int cv = C >> 3;
  __nv_bfloat16* dst = f.qkv + (size_t)m * 3 * C + C;
  for (int i = lane; i < 2 * cv; i += 32) {
    const int part = i >= cv;                              // 0: k, 1: v
    const int ch = (part ? i - cv : i) * 8;
    float x[8];
    const float* src = (part == 0 ? f.kpad : f.vpad) + ch;
    *reinterpret_cast<float4*>(x) = __ldg(reinterpret_cast<const float4*>(src));
    *reinterpret_cast<float4*>(x + 4) = __ldg(reinterpret_cast<const float4*>(src + 4));
    if (part == 0) {
      const int d = ch & 63;                               // channel inside the 64-wide head
      const int p = d >= 32 ? c : r;                       // second 32 channels use the column coordinate
      const int j0 = (d & 31) >> 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float cs = __ldg(f.cos_axis + p * 16 + j0 + j), sn = __ldg(f.sin_axis + p * 16 + j0 + j);
        const float x0 = x[2 * j], x1 = x[2 * j + 1];
        x[2 * j] = x0 * cs - x1 * sn;
        x[2 * j + 1] = x1 * cs + x0 * sn;
      }
    }
    uint4 val;
    val.x = pack_bf16(x[0], x[1]); val.y = pack_bf16(x[2], x[3]); val.z = pack_bf16(x[4], x[5]); val.w = pack_bf16(x[6], x[7]);
    *reinterpret_cast<uint4*>(dst + (size_t)part * C + ch) = val;
  }
}

__global__ void __launch_bounds__(256)
fill_pad_kv_rope_kernel(const PadFill f, int C) {
  pdl_wait();
  pdl_launch_dependents();
  fill_pad_row(f, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), threadIdx.x & 31, C);
}

__global__ void __launch_bounds__(256)
fill_pad_kv_kernel(__nv_bfloat16* __restrict__ qkv, const int* __restrict__ pad_rows, int n_pad, const float* __restrict__ v_bias, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int cv = C >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_pad * 2 * cv) return;
  const int c8 = (int)(idx % cv);
  const int part = (int)((idx / cv) & 1);                  // 0: k, 1: v
  const int row = pad_rows[idx / (2 * cv)];
  uint4 val = make_uint4(0u, 0u, 0u, 0u);
  if (part == 1 && v_bias != nullptr) {
    const float4 a = *reinterpret_cast<const float4*>(v_bias + c8 * 8), b = *reinterpret_cast<const float4*>(v_bias + c8 * 8 + 4);
    val.x = pack_bf16(a.x, a.y); val.y = pack_bf16(a.z, a.w); val.z = pack_bf16(b.x, b.y); val.w = pack_bf16(b.z, b.w);
  }
  *reinterpret_cast<uint4*>(qkv + (size_t)row * 3 * C + (size_t)(1 + part) * C + c8 * 8) = val;
}

// Image-level stable sort split by rank counting; grid (ceil(N/32), B) x 256 threads: the keys of one row sit
// in smem, 8 lanes share one element (lane s counts key groups s, s+8, ...) and combine with shuffles.
__global__ void __launch_bounds__(256)
topk_split_kernel(const float* __restrict__ scores, int N, int k, long long* __restrict__ keep_idx,
                  long long* __restrict__ drop_idx) {
  extern __shared__ __align__(16) uint32_t s_keys[];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const float* row = scores + (size_t)b * N;
  const int n4 = (N + 3) >> 2;
  for (int j = threadIdx.x; j < 4 * n4; j += blockDim.x) s_keys[j] = j < N ? order_key(row[j]) : 0u;
  __syncthreads();
  const int i = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int sub = threadIdx.x & 7;
  const bool ok = i < N;
  const uint32_t ki = ok ? s_keys[i] : 0u;
  int rank = 0;
  if (ok) {
    const uint4* k4 = reinterpret_cast<const uint4*>(s_keys);
#pragma unroll 4
    for (int g = sub; g < n4; g += 8) {
      const uint4 kj = k4[g];
      const int j = 4 * g;
      rank += (kj.x > ki) || (kj.x == ki && j + 0 < i);
      rank += (kj.y > ki) || (kj.y == ki && j + 1 < i);
      rank += (kj.z > ki) || (kj.z == ki && j + 2 < i);
      rank += (kj.w > ki) || (kj.w == ki && j + 3 < i);
    }
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  rank += __shfl_xor_sync(0xffffffffu, rank, 4);
  if (ok && sub == 0) {
    if (rank < k) keep_idx[(size_t)b * k + rank] = i;
    else drop_idx[(size_t)b * (N - k) + (rank - k)] = i;
  }
}

// ------------------------------------------------------------------------------------------------
// rep[w] = sum_j (s_j / sum s) * x[fast_map[w,j]];  grid (nW, C/(128*LV)), 256 threads: warp q owns
// rows j = q (mod 8) with LV float4 per lane (128*LV channels per CTA), 4 rows in flight; partials meet in smem.
template <int LV>
__global__ void __launch_bounds__(256)
merge_fast_kernel(const float* __restrict__ x, const int* __restrict__ fast_map, const float* __restrict__ fast_score,
                  int n_fast, int k, int C, float* __restrict__ rep_out, float* __restrict__ packed) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float s_part[8];
  __shared__ float s_wgt[1024];
  __shared__ int s_row[1024];
  __shared__ float4 s_acc[8][32 * LV];
  constexpr int CH = 128 * LV;
  const int w = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* fs = fast_score + (size_t)w * n_fast;
  const int* fm = fast_map + (size_t)w * n_fast;
  float part = 0.f;
  for (int j = threadIdx.x; j < n_fast; j += 256) {
    const float s = fs[j];
    s_wgt[j] = s;
    s_row[j] = fm[j];
    part += s;
  }
  part = warp_sum(part);
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  const float total = ((s_part[0] + s_part[1]) + (s_part[2] + s_part[3])) + ((s_part[4] + s_part[5]) + (s_part[6] + s_part[7]));
  const float4* xb = reinterpret_cast<const float4*>(x + blockIdx.y * CH) + lane;
  const size_t ld4 = (size_t)C >> 2;
  float4 a[LV];
#pragma unroll
  for (int i = 0; i < LV; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = warp; j0 < n_fast; j0 += 32) {
    float4 v[4][LV];
    float wg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 8 * u;
      const int r = j < n_fast ? s_row[j] : -1;
      wg[u] = j < n_fast ? s_wgt[j] / total : 0.f;             // weight = score / sum(score)
#pragma unroll
      for (int i = 0; i < LV; ++i) v[u][i] = r >= 0 ? xb[(size_t)r * ld4 + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < LV; ++i) {
        a[i].x += wg[u] * v[u][i].x; a[i].y += wg[u] * v[u][i].y;
        a[i].z += wg[u] * v[u][i].z; a[i].w += wg[u] * v[u][i].w;
      }
  }
#pragma unroll
  for (int i = 0; i < LV; ++i) s_acc[warp][lane + 32 * i] = a[i];
  __syncthreads();
  if (threadIdx.x < CH) {
    const float* sa = reinterpret_cast<const float*>(s_acc);
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += sa[q * CH + threadIdx.x];
    const int ch = blockIdx.y * CH + threadIdx.x;
    rep_out[(size_t)w * C + ch] = acc;
    if (packed) packed[((size_t)w * (k + 1) + k) * C + ch] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused norm1 front end of an accelerated block (toc3d_eva_vit.py:421-427 + :371): one launch does
//   blocks [0, nW)      : representative token of window w = sum_j (s_j / sum s) x[fast_map[w,j]] (fp32, written
//                         to rep_out[w] and packed[w*(k+1)+k]) followed by its LayerNorm -> out row w*(k+1)+k;
//   blocks [nW, ...)    : LayerNorm of the gathered slow rows (8 rows per block, warp per row).
// Merge block: warp q accumulates rows j = q (mod 8) over all C channels (VPL float4 per lane, two rows in
// flight), the eight partials meet in shared memory in a fixed order (deterministic).
// Deferred fast-token update of the PREVIOUS accelerated block (toc3d_eva_vit.py:452-461): its fast tokens still
// miss  x += packed_prev[rep row of their window] - rep_prev[window].  This launch reads every real row exactly once
// (slow rows in the LayerNorm blocks, fast rows in the merge blocks), so it adds the pending delta on the way and writes
// the row back - same expression as fast_update_kernel, hence bit-identical to running that kernel in between.
struct Pending {
  const int* fast_win;         // [rows of x] window of the previous block in which the row was a fast token | -1; nullptr = none
  const float* packed;         // previous block's packed / compact residual rows (its representative AFTER the block)
  const int* rep_row;          // previous block: row of `packed` holding window w's representative
  const float* rep;            // previous block: representative BEFORE the block, [nW_prev, C]
};

template <int VPL>
__global__ void __launch_bounds__(256)
ln_gather_merge_kernel(float* x, const int* __restrict__ tok_map, const int* __restrict__ fast_map,
                       const float* __restrict__ fast_score, const float* __restrict__ gamma,
                       const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, float* __restrict__ rep_out,
                       float* __restrict__ packed, int nW, int k, int n_fast, float eps, long long* __restrict__ zero_stats,
                       const int* __restrict__ rep_row, int ln_rows, int compact, const PadFill fill,
                       int* __restrict__ counters, const Pending pend) {
  constexpr int C = VPL * 128;
  constexpr int MSLICES = C >= 256 ? C / 256 : 1;
  const int mblocks = nW * MSLICES;       // merge blocks come first (they are the long pole), then LN, then pad fill
  __shared__ __align__(16) float s_acc[8][C >= 256 ? 256 : 128];
  __shared__ float s_wgt[1024];
  __shared__ int s_row[1024];
  __shared__ int s_pw[1024];
  __shared__ float s_red[8];
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ln_blocks = (ln_rows + 7) >> 3;
  if ((int)blockIdx.x >= mblocks + ln_blocks) {
    // ---- blocks after the LayerNorm ones: constant k / v of the pad rows of the packed qkv buffer
    fill_pad_row(fill, ((int)blockIdx.x - mblocks - ln_blocks) * 8 + warp, lane, C);
    return;
  }
  if ((int)blockIdx.x >= mblocks) {
    // ---- LayerNorm of gathered slow rows (pad slots are zero vectors -> beta), rep rows belong to the merge blocks
    // compact = 0: tok_map lists the packed rows (-1 = pad slot -> LN(0) = beta, -2 = rep, handled by a merge block);
    // compact = 1: tok_map lists the compact rows (ctok; -1 = unused row, skipped), out / packed use compact rows
    const int m = ((int)blockIdx.x - mblocks) * 8 + warp;
    if (m >= ln_rows) return;
    if (zero_stats != nullptr && lane == 0) *reinterpret_cast<longlong2*>(zero_stats + 2 * (size_t)m) = make_longlong2(0, 0);
    const int src = tok_map[m];
    if (src == -2 || (compact && src < 0)) return;
    float4* p = src >= 0 ? reinterpret_cast<float4*>(x + (size_t)src * C) : nullptr;
    const int pw = (pend.fast_win != nullptr && src >= 0) ? pend.fast_win[src] : -1;
    float4 v[VPL];
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = p ? p[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (pw >= 0) {                                     // fast token of the previous block: apply its pending update
      const float4* t2 = reinterpret_cast<const float4*>(pend.packed + (size_t)pend.rep_row[pw] * C) + lane;
      const float4* t0 = reinterpret_cast<const float4*>(pend.rep + (size_t)pw * C) + lane;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 a = __ldg(t2 + 32 * i), b = __ldg(t0 + 32 * i);
        v[i].x += a.x - b.x; v[i].y += a.y - b.y; v[i].z += a.z - b.z; v[i].w += a.w - b.w;
        p[lane + 32 * i] = v[i];
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) sm += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sm) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    uint2* o = reinterpret_cast<uint2*>(out + (size_t)m * C);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
      const float4 b = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
      uint2 u;
      u.x = pack_bf16((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      u.y = pack_bf16((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      o[lane + 32 * i] = u;
    }
    return;
  }
  // ---- representative token of window w: MS channel slices per window (one CTA each), the last one to finish
  //      normalises the whole row
  constexpr int MS = C >= 256 ? C / 256 : 1;          // slices per window
  constexpr int SW = C / MS;                          // channels per slice (256, or 128 for C = 128)
  constexpr int LV = SW / 128;                        // float4 per lane
  const int w = blockIdx.x / MS, part = blockIdx.x - w * MS;
  const float* fs = fast_score + (size_t)w * n_fast;
  const int* fm = fast_map + (size_t)w * n_fast;
  float part_s = 0.f;
  for (int j = threadIdx.x; j < n_fast; j += 256) {
    const float sc = fs[j];
    const int r = fm[j];
    s_wgt[j] = sc;
    s_row[j] = r;
    s_pw[j] = (pend.fast_win != nullptr && r >= 0) ? pend.fast_win[r] : -1;
    part_s += sc;
  }
  part_s = warp_sum(part_s);
  if (lane == 0) s_red[warp] = part_s;
  __syncthreads();
  const float total = ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + ((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]));
  float4 a[LV];
#pragma unroll
  for (int i = 0; i < LV; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4* xb = reinterpret_cast<float4*>(x + part * SW) + lane;
  for (int j0 = warp; j0 < n_fast; j0 += 32) {            // warp q owns rows j = q (mod 8), four rows in flight
    float4 v[4][LV];
    float wg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 8 * u;
      const int r = j < n_fast ? s_row[j] : -1;
      wg[u] = j < n_fast ? s_wgt[j] / total : 0.f;             // weight = score / sum(score)
#pragma unroll
      for (int i = 0; i < LV; ++i)
        v[u][i] = r >= 0 ? xb[(size_t)r * (C / 4) + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (pend.fast_win != nullptr) {                        // pending update of the previous block (this slice of the row)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 8 * u;
        const int pw = j < n_fast ? s_pw[j] : -1;
        if (pw >= 0) {
          const int r = s_row[j];
          const float4* t2 = reinterpret_cast<const float4*>(pend.packed + (size_t)pend.rep_row[pw] * C + part * SW) + lane;
          const float4* t0 = reinterpret_cast<const float4*>(pend.rep + (size_t)pw * C + part * SW) + lane;
#pragma unroll
          for (int i = 0; i < LV; ++i) {
            const float4 a = __ldg(t2 + 32 * i), b = __ldg(t0 + 32 * i);
            v[u][i].x += a.x - b.x; v[u][i].y += a.y - b.y; v[u][i].z += a.z - b.z; v[u][i].w += a.w - b.w;
            xb[(size_t)r * (C / 4) + 32 * i] = v[u][i];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < LV; ++i) {
        a[i].x += wg[u] * v[u][i].x; a[i].y += wg[u] * v[u][i].y;
        a[i].z += wg[u] * v[u][i].z; a[i].w += wg[u] * v[u][i].w;
      }
  }
  __syncthreads();                        // s_red is reused below
#pragma unroll
  for (int i = 0; i < LV; ++i) reinterpret_cast<float4*>(s_acc[warp])[lane + 32 * i] = a[i];
  __syncthreads();
  const size_t prow = (size_t)(rep_row != nullptr ? rep_row[w] : w * (k + 1) + k);
  if ((int)threadIdx.x < SW) {
    float r = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) r += s_acc[q][threadIdx.x];
    const int ch = part * SW + threadIdx.x;
    rep_out[(size_t)w * C + ch] = r;
    packed[prow * C + ch] = r;
  }
  if (MS > 1) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counters + w, 1) == MS - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) counters[w] = 0;             // clean for the next launch (CUDA-graph replays included)
  }
  // ---- LayerNorm of the finished row (thread t owns channels t, t + 256, ...)
  constexpr int PER = (C + 255) / 256;
  float r[PER];
  float sm = 0.f;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int ch = threadIdx.x + 256 * e;
    r[e] = ch < C ? __ldcg(rep_out + (size_t)w * C + ch) : 0.f;
    sm += r[e];
  }
  sm = warp_sum(sm);
  if (lane == 0) s_red[warp] = sm;
  __syncthreads();
  const float mean = (((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + ((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]))) / (float)C;
  float q2 = 0.f;
#pragma unroll
  for (int e = 0; e < PER; ++e)
    if (threadIdx.x + 256 * e < C) q2 += (r[e] - mean) * (r[e] - mean);
  q2 = warp_sum(q2);
  __syncthreads();
  if (lane == 0) s_red[warp] = q2;
  __syncthreads();
  const float var = (((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + ((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]))) / (float)C;
  const float rstd = rsqrtf(var + eps);
  __nv_bfloat16* o = out + (size_t)(compact && rep_row != nullptr ? rep_row[w] : w * (k + 1) + k) * C;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int ch = threadIdx.x + 256 * e;
    if (ch < C) o[ch] = __float2bfloat16_rn((r[e] - mean) * rstd * gamma[ch] + beta[ch]);
  }
}

// x[fast_map[w,j]] += packed[rep row of w] - rep[w];  one warp per fast token, VPL float4 per lane, all of the
// row's loads issued before the first use (the two representative rows are shared by the window: L1/L2 hits).
template <int VPL>
__global__ void __launch_bounds__(256)
fast_update_kernel(float* __restrict__ x, const int* __restrict__ fast_map, const float* __restrict__ packed,
                   const float* __restrict__ rep, int total_fast, int n_fast, int k, const int* __restrict__ rep_row) {
  constexpr int C = VPL * 128;
  pdl_wait();
  pdl_launch_dependents();
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= total_fast) return;
  const int row = fast_map[f];
  if (row < 0) return;
  const int w = f / n_fast;
  const int lane = threadIdx.x & 31;
  const float4* t2 = reinterpret_cast<const float4*>(packed + (size_t)(rep_row != nullptr ? rep_row[w] : w * (k + 1) + k) * C) + lane;
  const float4* t0 = reinterpret_cast<const float4*>(rep + (size_t)w * C) + lane;
  float4* xr = reinterpret_cast<float4*>(x + (size_t)row * C) + lane;
  float4 v[VPL], a[VPL], b[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[32 * i];
    a[i] = __ldg(t2 + 32 * i);
    b[i] = __ldg(t0 + 32 * i);
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i].x += a[i].x - b[i].x; v[i].y += a[i].y - b[i].y; v[i].z += a[i].z - b[i].z; v[i].w += a[i].w - b[i].w;
    xr[32 * i] = v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// (the query encoder and the scorer folding live in motion_queries.cu)
__device__ __forceinline__ float gumbel_from_hash(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;   // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = ((float)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  return -logf(-logf(u));
}

// effective noise seed = seed + 1000003 * *seed_dev (device-resident call counter, so a captured CUDA
// graph draws fresh noise on every replay)
__device__ __forceinline__ uint64_t mix_seed(uint64_t seed, const uint64_t* seed_dev) {
  return seed_dev ? seed + 1000003ull * (*seed_dev) : seed;
}

__device__ __forceinline__ void score_tail(float l0, float l1, size_t tok, const float* gumbel, uint64_t seed,
                                           float* pred, float* score, float* mask_out) {
  const float mx = fmaxf(l0, l1);
  const float lse = mx + logf(expf(l0 - mx) + expf(l1 - mx));
  const float p0 = l0 - lse, p1 = l1 - lse;
  if (pred) { pred[tok * 2] = p0; pred[tok * 2 + 1] = p1; }
  if (score) score[tok] = p0;
  if (mask_out) {
    const float g0 = gumbel ? gumbel[tok * 2] : gumbel_from_hash(seed, tok * 2);
    const float g1 = gumbel ? gumbel[tok * 2 + 1] : gumbel_from_hash(seed, tok * 2 + 1);
    const float a = p0 + g0, b = p1 + g1;
    const float m2 = fmaxf(a, b);
    const float ea = expf(a - m2), eb = expf(b - m2);
    mask_out[tok] = ea / (ea + eb);
  }
}

// one warp per token: logit_o = mask * (x . A[f][o]) + c[f][o]
__global__ void __launch_bounds__(256)
score_tokens_kernel(const float* __restrict__ x, const float* __restrict__ mask_in, const float* __restrict__ A,
                    const float* __restrict__ cvec, int V, int N, int C, int vpf, const float* __restrict__ gumbel,
                    uint64_t seed, const uint64_t* __restrict__ seed_dev, float* __restrict__ pred, float* __restrict__ score, float* __restrict__ mask_out) {
  pdl_wait();
  pdl_launch_dependents();
  const size_t tok = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= (size_t)V * N) return;
  const int lane = threadIdx.x & 31;
  const int f = (int)(tok / N) / vpf;
  const float4* xr = reinterpret_cast<const float4*>(x + tok * C);
  const float4* a0 = reinterpret_cast<const float4*>(A + ((size_t)f * 2) * C);
  const float4* a1 = reinterpret_cast<const float4*>(A + ((size_t)f * 2 + 1) * C);
  float d0 = 0.f, d1 = 0.f;
  for (int i = lane; i < (C >> 2); i += 32) {
    const float4 xv = xr[i], u = __ldg(a0 + i), w = __ldg(a1 + i);
    d0 += (xv.x * u.x + xv.y * u.y) + (xv.z * u.z + xv.w * u.w);
    d1 += (xv.x * w.x + xv.y * w.y) + (xv.z * w.z + xv.w * w.w);
  }
  d0 = warp_sum(d0);
  d1 = warp_sum(d1);
  if (lane == 0) {
    const float mk = mask_in ? mask_in[tok] : 1.0f;
    score_tail(mk * d0 + cvec[f * 2], mk * d1 + cvec[f * 2 + 1], tok, gumbel, mix_seed(seed, seed_dev), pred, score,
               mask_out);
  }
}

__global__ void score_finish_kernel(const float* __restrict__ logits, int M, const float* __restrict__ gumbel,
                                    uint64_t seed, const uint64_t* __restrict__ seed_dev, float* pred, float* score,
                                    float* mask_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int tok = blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= M) return;
  score_tail(logits[tok * 2], logits[tok * 2 + 1], (size_t)tok, gumbel, mix_seed(seed, seed_dev), pred, score, mask_out);
}

// ------------------------------------------------------------------------------------------------
// im2col for the 16x16 stride-16 stem: thread = 8 consecutive kx of one (v, c, y, patch-col).
__global__ void __launch_bounds__(256)
im2col_patch16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int V, int Hi, int Wi) {
  pdl_wait();
  pdl_launch_dependents();
  const int Wp = Wi >> 4, Hp = Hi >> 4;
  const int halves = Wp * 2;                       // 8-pixel chunks per image row
  const size_t total = (size_t)V * 3 * Hi * halves;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int hx = (int)(idx % halves);
  size_t r = idx / halves;
  const int y = (int)(r % Hi); r /= Hi;
  const int c = (int)(r % 3);
  const int v = (int)(r / 3);
  const float4* src = reinterpret_cast<const float4*>(img + (((size_t)v * 3 + c) * Hi + y) * Wi + hx * 8);
  const float4 a = src[0], b = src[1];
  const int pw = hx >> 1, kx0 = (hx & 1) * 8, ph = y >> 4, ky = y & 15;
  const size_t row = ((size_t)v * Hp + ph) * Wp + pw;
  uint4 u;
  u.x = pack_bf16(a.x, a.y); u.y = pack_bf16(a.z, a.w);
  u.z = pack_bf16(b.x, b.y); u.w = pack_bf16(b.z, b.w);
  *reinterpret_cast<uint4*>(out + row * 768 + c * 256 + ky * 16 + kx0) = u;
}

// Camera crop -> patch matrix in one pass (next row f3): NormalizeMultiviewImage + PadMultiViewImage
// (transform_3d.py:21-104; mmcv.imnormalize / impad_to_multiple) fused with the stem im2col.
// img u8 HWC [V,Hs,Ws,3]; lut fp32 [3,256] = normalised value of byte b in OUTPUT channel c (built on the host in
// cv2's arithmetic); output channel c reads input channel (to_rgb ? 2-c : c); pixels outside Hs x Ws are the pad
// value 0 of the normalised image.  Thread = 8 consecutive pixels of one image row (24 bytes in, 3 x 16 bytes out).
__global__ void __launch_bounds__(256)
preprocess_patch16_u8_kernel(const uint8_t* __restrict__ img, const float* __restrict__ lut, __nv_bfloat16* __restrict__ out,
                             int V, int Hs, int Ws, int Hi, int Wi, int to_rgb) {
  __shared__ float s_lut[3 * 256];
  pdl_wait();
  pdl_launch_dependents();
  for (int j = threadIdx.x; j < 3 * 256; j += blockDim.x) s_lut[j] = lut[j];
  __syncthreads();
  const int Wp = Wi >> 4, Hp = Hi >> 4;
  const int halves = Wp * 2;
  const size_t total = (size_t)V * Hi * halves;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int hx = (int)(idx % halves);
  size_t r = idx / halves;
  const int y = (int)(r % Hi);
  const int v = (int)(r / Hi);
  const int x0 = hx * 8;
  uint8_t px[24];
  if (y < Hs && x0 + 8 <= Ws && (Ws & 7) == 0 && (reinterpret_cast<uintptr_t>(img) & 7) == 0) {
    const uint2* src = reinterpret_cast<const uint2*>(img + (((size_t)v * Hs + y) * Ws + x0) * 3);
    const uint2 a = src[0], b = src[1], c = src[2];
    const uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
    for (int j = 0; j < 24; ++j) px[j] = (uint8_t)(w[j >> 2] >> ((j & 3) * 8));
  } else {
    const uint8_t* src = img + (((size_t)v * Hs + (y < Hs ? y : 0)) * Ws) * 3;
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const int x = x0 + j / 3;
      px[j] = (y < Hs && x < Ws) ? src[(size_t)x * 3 + j % 3] : 0;
    }
  }
  const int pw = hx >> 1, kx0 = (hx & 1) * 8, ph = y >> 4, ky = y & 15;
  const size_t row = ((size_t)v * Hp + ph) * Wp + pw;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int ci = to_rgb ? 2 - c : c;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (y < Hs && x0 + j < Ws) ? s_lut[c * 256 + px[j * 3 + ci]] : 0.0f;
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
    u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + row * 768 + c * 256 + ky * 16 + kx0) = u;
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 u;
    u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = u;
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

__global__ void __launch_bounds__(256)
mask_rows_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ out, int M, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const float mk = mask[m];
  const float4* p = reinterpret_cast<const float4*>(x + (size_t)m * C);
  float4* o = reinterpret_cast<float4*>(out + (size_t)m * C);
  for (int i = threadIdx.x & 31; i < (C >> 2); i += 32) {
    float4 v = p[i];
    v.x *= mk; v.y *= mk; v.z *= mk; v.w *= mk;
    o[i] = v;
  }
}

// y[v, :, C/2:] <- mean over tokens; grid (V, C/2/128) x 128 threads; two passes over N.
__global__ void __launch_bounds__(128)
global_half_mean_kernel(__nv_bfloat16* __restrict__ y, int N, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int v = blockIdx.x;
  const int ch = C / 2 + blockIdx.y * 128 + threadIdx.x;
  if (ch >= C) return;
  __nv_bfloat16* base = y + (size_t)v * N * C + ch;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc += __bfloat162float(base[(size_t)n * C]);
  const __nv_bfloat16 m = __float2bfloat16_rn(acc / (float)N);
  for (int n = 0; n < N; ++n) base[(size_t)n * C] = m;
}

}  // namespace toc3d

// =================================================================================================
using namespace toc3d;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int toc3d_abi_version(void) { return TOC3D_B200_ABI_VERSION; }
extern "C" const char* toc3d_last_error(void) { return toc3d::g_err; }

extern "C" int toc3d_layernorm_rows(const float* x, const int32_t* row_map, const float* alt, const float* gamma,
                                    const float* beta, void* out, int32_t M, int32_t C, float eps, int32_t pad_mode,
                                    int64_t* zero_stats, void* stream) {
  TOC3D_REQUIRE(x && gamma && beta && out, kErrBadArg, "toc3d_layernorm_rows: null pointer");
  TOC3D_REQUIRE(M > 0 && C % 128 == 0 && C >= 128 && C <= 4096, kErrBadArg, "toc3d_layernorm_rows: bad shape M=%d C=%d", M, C);
  const int rows_per_block = 8;
  dim3 grid((M + rows_per_block - 1) / rows_per_block), block(256);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define LN_CASE(V)                                                                                                  \
  case V: TOC3D_CHECK_CUDA(launch_pdl(layernorm_rows_kernel<V>, dim3(grid), dim3(block), 0, ST(stream), 1, x, row_map, alt, gamma, beta, o, M, eps, pad_mode, reinterpret_cast<long long*>(zero_stats))); break;
  switch (C / 128) {
    LN_CASE(1) LN_CASE(2) LN_CASE(4) LN_CASE(6) LN_CASE(8) LN_CASE(10) LN_CASE(12) LN_CASE(16) LN_CASE(32)
    default: TOC3D_REQUIRE(false, kErrBadArg, "toc3d_layernorm_rows: unsupported C=%d", C);
  }
#undef LN_CASE
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_subln_bf16(const void* h, void* out, const float* gamma, const float* beta, int32_t M, int32_t Hd,
                                int32_t ld, float eps, void* stream) {
  TOC3D_REQUIRE(h && out && gamma && beta, kErrBadArg, "toc3d_subln_bf16: null pointer");
  TOC3D_REQUIRE(M > 0 && Hd > 0 && Hd <= ld && ld % 8 == 0 && ld <= 3072, kErrBadArg,
                "toc3d_subln_bf16: bad shape M=%d Hd=%d ld=%d", M, Hd, ld);
  TOC3D_CHECK_CUDA(launch_pdl(subln_kernel, dim3((M + 7) / 8), dim3(256), 0, ST(stream), 1, reinterpret_cast<const __nv_bfloat16*>(h),
                                                    reinterpret_cast<__nv_bfloat16*>(out), gamma, beta, M, Hd, ld, eps));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_window_topk(const float* scores, int32_t V, int32_t H, int32_t W, int32_t ws, int32_t k,
                                 int32_t* slow_idx, int32_t* fast_idx, float* fast_score, int32_t* tok_map,
                                 int32_t* rope_rows, int32_t* fast_map, int32_t* fast_win, void* stream) {
  TOC3D_REQUIRE(scores, kErrBadArg, "toc3d_window_topk: null scores");
  const int n = ws * ws;
  TOC3D_REQUIRE(V > 0 && H > 0 && W > 0 && ws > 0 && n <= 1024 && k >= 0 && k <= n, kErrBadArg,
                "toc3d_window_topk: bad shape V=%d H=%d W=%d ws=%d k=%d", V, H, W, ws, k);
  const int nW = V * ((H + ws - 1) / ws) * ((W + ws - 1) / ws);
  const int threads = ((n + 31) / 32) * 32;
  TOC3D_CHECK_CUDA(launch_pdl(window_topk_kernel, dim3(nW), dim3(threads), 0, ST(stream), 1, scores, V, H, W, ws, k, slow_idx, fast_idx,
                                                                      fast_score, tok_map, rope_rows, fast_map, fast_win));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_compact_rows(const int32_t* tok_map, const int32_t* rope_rows, const int32_t* coff, const int32_t* rcap,
                                  int32_t nW, int32_t k, int32_t* cmap, int32_t* ctok, int32_t* rep_row, int32_t* cinv,
                                  int32_t* crope, int32_t* prope, void* stream) {
  TOC3D_REQUIRE(tok_map && coff && rcap && cmap && ctok && rep_row, kErrBadArg, "toc3d_compact_rows: null pointer");
  TOC3D_REQUIRE(nW > 0 && k >= 1 && k <= 1023, kErrBadArg, "toc3d_compact_rows: bad shape nW=%d k=%d", nW, k);
  const int threads = ((k + 1 + 31) / 32) * 32;
  TOC3D_CHECK_CUDA(launch_pdl(compact_rows_kernel, dim3(nW), dim3(threads), 0, ST(stream), 1, tok_map, rope_rows, coff, rcap, k, cmap, ctok, rep_row,
                              cinv, crope, prope));
  return 0;
}

extern "C" int toc3d_fill_pad_kv(void* qkv, const int32_t* pad_rows, int32_t n_pad, const float* v_bias, int32_t C, void* stream) {
  TOC3D_REQUIRE(qkv && (pad_rows || n_pad == 0), kErrBadArg, "toc3d_fill_pad_kv: null pointer");
  TOC3D_REQUIRE(n_pad >= 0 && C > 0 && C % 8 == 0, kErrBadArg, "toc3d_fill_pad_kv: bad shape n_pad=%d C=%d", n_pad, C);
  if (n_pad == 0) return 0;
  const size_t total = (size_t)n_pad * 2 * (C / 8);
  TOC3D_CHECK_CUDA(launch_pdl(fill_pad_kv_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ST(stream), 1,
                              reinterpret_cast<__nv_bfloat16*>(qkv), pad_rows, n_pad, v_bias, C));
  return 0;
}

extern "C" int toc3d_fill_pad_kv_rope(void* qkv, const int32_t* cmap, const int32_t* rope_rows, int32_t Mp, const float* kpad,
                                      const float* vpad, const float* cos_axis, const float* sin_axis, int32_t ft, int32_t C,
                                      void* stream) {
  TOC3D_REQUIRE(qkv && cmap && rope_rows && kpad && vpad && cos_axis && sin_axis, kErrBadArg, "toc3d_fill_pad_kv_rope: null pointer");
  TOC3D_REQUIRE(Mp > 0 && ft > 0 && C > 0 && C % 64 == 0, kErrBadArg, "toc3d_fill_pad_kv_rope: bad shape Mp=%d ft=%d C=%d", Mp, ft, C);
  PadFill f{reinterpret_cast<__nv_bfloat16*>(qkv), cmap, rope_rows, Mp, kpad, vpad, cos_axis, sin_axis, ft};
  TOC3D_CHECK_CUDA(launch_pdl(fill_pad_kv_rope_kernel, dim3((unsigned)((Mp + 7) / 8)), dim3(256), 0, ST(stream), 1, f, C));
  return 0;
}

extern "C" int toc3d_topk_split(const float* scores, int32_t B, int32_t N, int32_t k, int64_t* keep_idx,
                                int64_t* drop_idx, void* stream) {
  TOC3D_REQUIRE(scores && keep_idx && drop_idx, kErrBadArg, "toc3d_topk_split: null pointer");
  TOC3D_REQUIRE(B > 0 && N > 0 && N <= 12288 && k >= 0 && k <= N, kErrBadArg, "toc3d_topk_split: bad shape B=%d N=%d k=%d", B, N, k);
  dim3 grid((N + 31) / 32, B);
  TOC3D_CHECK_CUDA(launch_pdl(topk_split_kernel, dim3(grid), dim3(256), ((N + 3) / 4) * 16, ST(stream), 1, scores, N, k, reinterpret_cast<long long*>(keep_idx),
                                                                   reinterpret_cast<long long*>(drop_idx)));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_merge_fast_tokens(const float* x, const int32_t* fast_map, const float* fast_score, int32_t nW,
                                       int32_t n_fast, int32_t k, int32_t C, float* rep_out, float* packed, void* stream) {
  TOC3D_REQUIRE(x && fast_map && fast_score && rep_out, kErrBadArg, "toc3d_merge_fast_tokens: null pointer");
  TOC3D_REQUIRE(nW > 0 && n_fast > 0 && n_fast <= 1024 && C % 128 == 0, kErrBadArg,
                "toc3d_merge_fast_tokens: bad shape nW=%d n_fast=%d C=%d", nW, n_fast, C);
  if (C % 256 == 0)
    TOC3D_CHECK_CUDA(launch_pdl(merge_fast_kernel<2>, dim3(nW, C / 256), dim3(256), 0, ST(stream), 1, x, fast_map, fast_score, n_fast, k, C, rep_out, packed));
  else
    TOC3D_CHECK_CUDA(launch_pdl(merge_fast_kernel<1>, dim3(nW, C / 128), dim3(256), 0, ST(stream), 1, x, fast_map, fast_score, n_fast, k, C, rep_out, packed));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_fast_token_update(float* x, const int32_t* fast_map, const float* packed, const float* rep,
                                       int32_t nW, int32_t n_fast, int32_t k, int32_t C, const int32_t* rep_row, void* stream) {
  TOC3D_REQUIRE(x && fast_map && packed && rep, kErrBadArg, "toc3d_fast_token_update: null pointer");
  TOC3D_REQUIRE(nW > 0 && n_fast > 0 && C % 128 == 0, kErrBadArg, "toc3d_fast_token_update: bad shape");
  const int total = nW * n_fast;
  dim3 grid((total + 7) / 8), block(256);
#define UPD_CASE(V)                                                                                                  \
  case V: TOC3D_CHECK_CUDA(launch_pdl(fast_update_kernel<V>, grid, block, 0, ST(stream), 1, x, fast_map, packed, rep, total, n_fast, k, rep_row)); break;
  switch (C / 128) {
    UPD_CASE(1) UPD_CASE(2) UPD_CASE(4) UPD_CASE(6) UPD_CASE(8) UPD_CASE(16)
    default: TOC3D_REQUIRE(false, kErrBadArg, "toc3d_fast_token_update: unsupported C=%d", C);
  }
#undef UPD_CASE
  return 0;
}

extern "C" int toc3d_ln_gather_merge(float* x, const int32_t* tok_map, const int32_t* fast_map,
                                     const float* fast_score, const float* gamma, const float* beta, void* out,
                                     float* rep_out, float* packed, int32_t nW, int32_t k, int32_t n_fast, int32_t C,
                                     float eps, int64_t* zero_stats, const int32_t* rep_row, int32_t compact_rows,
                                     const toc3d_pad_fill* pf, int32_t* counters, const toc3d_pending_update* pu, void* stream) {
  TOC3D_REQUIRE(x && tok_map && fast_map && fast_score && gamma && beta && out && rep_out && packed, kErrBadArg,
                "toc3d_ln_gather_merge: null pointer");
  TOC3D_REQUIRE(nW > 0 && k >= 0 && n_fast > 0 && n_fast <= 1024, kErrBadArg,
                "toc3d_ln_gather_merge: bad shape nW=%d k=%d n_fast=%d", nW, k, n_fast);
  TOC3D_REQUIRE(compact_rows >= 0 && (compact_rows == 0 || rep_row != nullptr), kErrBadArg,
                "toc3d_ln_gather_merge: compact_rows needs rep_row");
  const int M = compact_rows > 0 ? compact_rows : nW * (k + 1);
  const int compact = compact_rows > 0 ? 1 : 0;
  PadFill fill{};
  int fill_blocks = 0;
  if (pf != nullptr && pf->qkv != nullptr) {
    TOC3D_REQUIRE(pf->cmap && pf->rope_rows && pf->kpad && pf->vpad && pf->cos_axis && pf->sin_axis && pf->Mp > 0 && pf->ft > 0 &&
                  C % 64 == 0, kErrBadArg, "toc3d_ln_gather_merge: incomplete pad-fill description");
    fill = PadFill{reinterpret_cast<__nv_bfloat16*>(pf->qkv), pf->cmap, pf->rope_rows, pf->Mp, pf->kpad, pf->vpad, pf->cos_axis,
                   pf->sin_axis, pf->ft};
    fill_blocks = (pf->Mp + 7) / 8;
  }
  Pending pend{};
  if (pu != nullptr && pu->fast_win != nullptr) {
    TOC3D_REQUIRE(pu->packed && pu->rep_row && pu->rep, kErrBadArg, "toc3d_ln_gather_merge: incomplete pending-update description");
    TOC3D_REQUIRE(pu->packed != packed && pu->rep != rep_out, kErrBadArg,
                  "toc3d_ln_gather_merge: the pending update must read buffers this launch does not write (ping-pong packed / rep)");
    pend = Pending{pu->fast_win, pu->packed, pu->rep_row, pu->rep};
  }
  const int mslices = C >= 256 ? C / 256 : 1;
  TOC3D_REQUIRE(mslices == 1 || counters != nullptr, kErrBadArg, "toc3d_ln_gather_merge: counters (int32 [nW], zeroed) required for C >= 256");
  dim3 grid(nW * mslices + (M + 7) / 8 + fill_blocks), block(256);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  long long* zs = reinterpret_cast<long long*>(zero_stats);
#define LGM_CASE(V)                                                                                                  \
  case V: TOC3D_CHECK_CUDA(launch_pdl(ln_gather_merge_kernel<V>, grid, block, 0, ST(stream), 1, x, tok_map, fast_map, fast_score, gamma, beta, o, rep_out, packed, nW, k, n_fast, eps, zs, rep_row, M, compact, fill, counters, pend)); break;
  switch (C % 128 == 0 ? C / 128 : 0) {
    LGM_CASE(1) LGM_CASE(2) LGM_CASE(4) LGM_CASE(6) LGM_CASE(8)
    default: TOC3D_REQUIRE(false, kErrBadArg, "toc3d_ln_gather_merge: C must be 128, 256, 512, 768 or 1024 (got %d)", C);
  }
#undef LGM_CASE
  return 0;
}

extern "C" int toc3d_score_tokens(const float* x, const float* mask_in, const float* A, const float* c, int32_t V,
                                  int32_t N, int32_t C, int32_t views_per_frame, const float* gumbel, uint64_t seed,
                                  const uint64_t* seed_dev, float* pred, float* score, float* mask_out, void* stream) {
  TOC3D_REQUIRE(x && A && c, kErrBadArg, "toc3d_score_tokens: null pointer");
  TOC3D_REQUIRE(V > 0 && N > 0 && C % 4 == 0 && views_per_frame > 0 && V % views_per_frame == 0, kErrBadArg,
                "toc3d_score_tokens: bad shape V=%d N=%d C=%d vpf=%d", V, N, C, views_per_frame);
  const long long toks = (long long)V * N;
  TOC3D_CHECK_CUDA(launch_pdl(score_tokens_kernel, dim3((unsigned)((toks + 7) / 8)), dim3(256), 0, ST(stream), 1, x, mask_in, A, c, V, N, C, views_per_frame, gumbel,
                                                                          seed, seed_dev, pred, score, mask_out));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_score_finish(const float* logits, int32_t M, const float* gumbel, uint64_t seed,
                                  const uint64_t* seed_dev, float* pred, float* score, float* mask_out, void* stream) {
  TOC3D_REQUIRE(logits && M > 0, kErrBadArg, "toc3d_score_finish: bad args");
  TOC3D_CHECK_CUDA(launch_pdl(score_finish_kernel, dim3((M + 255) / 256), dim3(256), 0, ST(stream), 1, logits, M, gumbel, seed, seed_dev, pred, score, mask_out));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_im2col_patch16(const float* img, void* out, int32_t V, int32_t Hi, int32_t Wi, void* stream) {
  TOC3D_REQUIRE(img && out, kErrBadArg, "toc3d_im2col_patch16: null pointer");
  TOC3D_REQUIRE(V > 0 && Hi > 0 && Wi > 0 && Hi % 16 == 0 && Wi % 16 == 0, kErrBadArg,
                "toc3d_im2col_patch16: image %dx%d must be a multiple of the 16x16 patch", Hi, Wi);
  const size_t total = (size_t)V * 3 * Hi * (Wi / 8);
  TOC3D_CHECK_CUDA(launch_pdl(im2col_patch16_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ST(stream), 1, img, reinterpret_cast<__nv_bfloat16*>(out), V, Hi, Wi));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_preprocess_patch16_u8(const uint8_t* img, const float* lut, void* out, int32_t V, int32_t Hs, int32_t Ws,
                                           int32_t Hi, int32_t Wi, int32_t to_rgb, void* stream) {
  TOC3D_REQUIRE(img && lut && out, kErrBadArg, "toc3d_preprocess_patch16_u8: null pointer");
  TOC3D_REQUIRE(V > 0 && Hs > 0 && Ws > 0 && Hi >= Hs && Wi >= Ws && Hi % 16 == 0 && Wi % 16 == 0, kErrBadArg,
                "toc3d_preprocess_patch16_u8: crop %dx%d must fit the padded image %dx%d (a multiple of the 16x16 patch)", Hs, Ws, Hi, Wi);
  const size_t total = (size_t)V * Hi * (Wi / 8);
  TOC3D_CHECK_CUDA(launch_pdl(preprocess_patch16_u8_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ST(stream), 1,
                              img, lut, reinterpret_cast<__nv_bfloat16*>(out), V, Hs, Ws, Hi, Wi, to_rgb));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  TOC3D_REQUIRE(in && out && n > 0, kErrBadArg, "toc3d_cast_f32_to_bf16: bad args");
  TOC3D_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0, kErrBadArg, "toc3d_cast_f32_to_bf16: unaligned");
  const long long threads = (n + 3) / 4;
  TOC3D_CHECK_CUDA(launch_pdl(cast_bf16_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, ST(stream), 1, in, reinterpret_cast<__nv_bfloat16*>(out), n));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_mask_rows(const float* x, const float* mask, float* out, int32_t M, int32_t C, void* stream) {
  TOC3D_REQUIRE(x && mask && out && M > 0 && C % 4 == 0, kErrBadArg, "toc3d_mask_rows: bad args");
  TOC3D_CHECK_CUDA(launch_pdl(mask_rows_kernel, dim3((M + 7) / 8), dim3(256), 0, ST(stream), 1, x, mask, out, M, C));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int toc3d_global_half_mean(void* y, int32_t V, int32_t N, int32_t C, void* stream) {
  TOC3D_REQUIRE(y && V > 0 && N > 0 && C % 2 == 0, kErrBadArg, "toc3d_global_half_mean: bad args");
  dim3 grid(V, (C / 2 + 127) / 128);
  TOC3D_CHECK_CUDA(launch_pdl(global_half_mean_kernel, dim3(grid), dim3(128), 0, ST(stream), 1, reinterpret_cast<__nv_bfloat16*>(y), N, C));
  TOC3D_CHECK_CUDA(cudaGetLastError());
  return 0;
}
