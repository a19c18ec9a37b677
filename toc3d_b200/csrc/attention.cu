// Windowed multi-head attention, head dim 64, sequence = one window (<= 401 tokens in the shipped
// configs).  Three kernels behind toc3d_window_attention:
//   attn_tc::window_attention_pp_kernel (seq <= 256): tcgen05 / TMEM / TMA, persistent, one CTA per SM looping over
//            (window, head) items.  Q, K, V of an item are staged by TMA (64-row boxes, 128B swizzle) into a ring of
//            smem buffers while earlier items are processed.  TMEM holds two 256-column slots; per 128-row query
//            tile S = Q K^T is one tcgen05.mma chain into a slot, the slot's four softmax warps (thread = row) read
//            S with tcgen05.ld, take the exact row max over the whole window (no online rescaling at these
//            lengths), write P = exp2(.) as packed bf16 back over S with tcgen05.st, and O = P V runs as
//            tcgen05.mma with the A operand in TMEM and V as an MN-major smem operand.  The MMA thread issues
//            S(u) then P V(u-1), so one slot's MMAs overlap the other slot's softmax.
//   attn_tc::window_attention_tc2_kernel (256 < seq <= 448): same math, one CTA per (window, head), one slot, two
//            softmax warps per TMEM lane quarter splitting the key columns.
//   attn::window_attention_kernel (seq > 448, not reached by any shipped config): flash-style mma.sync fallback.
// Measured and NOT adopted (round 2, profiles/r02h_attn_bench_tail_warp.txt): an 11th warp computing the 129th query row
// of 128 + 1 row windows on CUDA cores (so that those windows need one tensor tile instead of two) - correct, but the
// extra warp slowed every shape (48 x 256 keys: 35.3 -> 41.7 us) and the row itself took longer than the tile it replaced.
// All take an optional out_map (rows stored in compact order, padding rows skipped) and q_rows (only the leading
// query rows of a window are needed; the rest is padding that only serves as keys / values).
//
// Fallback kernel: one CTA = 64 query rows of one (window, head); K/V tiles of 64 keys are
// double-buffered with cp.async; S = QK^T and O += PV run on mma.sync m16n8k16 (bf16 in, fp32
// accumulate); softmax is online in fp32.  q arrives rotated and pre-scaled (QKV epilogue).
//
// Replaces eva_vit.py:109-111 / toc3d_eva_vit.py:509-511 (q@k^T, softmax, @v, head merge).
#include "common.cuh"
#include "../../include/toc3d_b200.h"

#include <mutex>

namespace toc3d {

// Diagnostic build only (-DTOC3D_ATTN_TRACE, tools/probes/attn_trace.py): clock64 stamps of CTA 0 of the ping-pong kernel,
// [role = slot 0 | slot 1 | MMA thread][unit of that role][8 stamps].  Compiled out of the product library.
#ifdef TOC3D_ATTN_TRACE
__device__ unsigned long long g_attn_trace[3 * 32 * 8];
#define ATRACE(cond, role, unit, k)                                                                    \
  do {                                                                                                 \
    if ((cond) && blockIdx.x == 0 && (unit) < 32) g_attn_trace[((role) * 32 + (unit)) * 8 + (k)] = clock64(); \
  } while (0)
#else
#define ATRACE(cond, role, unit, k) do {} while (0)
#endif

namespace attn {

constexpr int D = 64;          // head dim
constexpr int BQ = 64;         // query rows per CTA (16 per warp)
constexpr int BKV = 64;        // keys per tile
constexpr int LDS = D + 8;     // padded smem row (144 B) -> conflict-free ldmatrix
constexpr int NTHREADS = 128;

struct Smem {
  __nv_bfloat16 q[BQ][LDS];
  __nv_bfloat16 k[2][BKV][LDS];
  __nv_bfloat16 v[2][BKV][LDS];
};

// copy `rows` x 64 bf16 (row stride ld elements in gmem) into a padded smem tile; rows >= valid are zero-filled
__device__ __forceinline__ void load_tile(__nv_bfloat16 (*dst)[LDS], const __nv_bfloat16* src, int64_t ld, int valid) {
  // 64 rows x 8 chunks of 16 B = 512 chunks / 128 threads
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = threadIdx.x + i * NTHREADS;
    const int r = c >> 3, ch = c & 7;
    const bool ok = r < valid;
    cp_async16(&dst[r][ch * 8], src + (ok ? (int64_t)r * ld + ch * 8 : 0), ok);
  }
}

__global__ void __launch_bounds__(NTHREADS)
window_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int seq, int heads,
                        const int* __restrict__ out_map, const int* __restrict__ q_rows) {
  __shared__ __align__(16) Smem sm;
  pdl_wait();
  pdl_launch_dependents();
  if (q_rows != nullptr && (int)blockIdx.x * BQ >= q_rows[blockIdx.z]) return;      // padding-only query tile
  const int C = heads * D;
  const int64_t ld = 3 * (int64_t)C;
  const int qt = blockIdx.x, h = blockIdx.y, w = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const int64_t row0 = (int64_t)w * seq;
  const __nv_bfloat16* qbase = qkv + (row0 + qt * BQ) * ld + h * D;
  const __nv_bfloat16* kbase = qkv + row0 * ld + C + h * D;
  const __nv_bfloat16* vbase = qkv + row0 * ld + 2 * C + h * D;
  const int q_valid = min(BQ, seq - qt * BQ);
  const int n_tiles = (seq + BKV - 1) / BKV;

  load_tile(sm.q, qbase, ld, q_valid);
  load_tile(sm.k[0], kbase, ld, min(BKV, seq));
  load_tile(sm.v[0], vbase, ld, min(BKV, seq));
  cp_async_commit();

  uint32_t qf[4][4];            // Q fragments: 4 k-steps of 16 along d
  float o[8][4];                // O accumulators: 8 d-tiles of 8
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

  constexpr float LOG2E = 1.4426950408889634f;

  for (int kt = 0; kt < n_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_tiles) {
      const int valid = min(BKV, seq - (kt + 1) * BKV);
      load_tile(sm.k[buf ^ 1], kbase + (int64_t)(kt + 1) * BKV * ld, ld, valid);
      load_tile(sm.v[buf ^ 1], vbase + (int64_t)(kt + 1) * BKV * ld, ld, valid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(qf[ks], &sm.q[warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + (lane >> 4) * 8]);
    }

    // ---- S = Q K^T : 16 x 64 per warp
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {   // pairs of 8-key tiles
        uint32_t kf[4];
        ldmatrix_x4(kf, &sm.k[buf][np * 16 + (lane & 7) + (lane >> 4) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
        mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
      }
    }

    // ---- mask the tail keys, online softmax (rows g and g+8 of this warp's 16)
    const int key0 = kt * BKV;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + nt * 8 + 2 * t + (j & 1);
        if (key >= seq) s[nt][j] = -INFINITY;
        mx[j >> 1] = fmaxf(mx[j >> 1], s[nt][j]);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);            // finite: every tile holds at least one valid key
      corr[r] = exp2f((m_run[r] - mnew[r]) * LOG2E);
      m_run[r] = mnew[r];
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];   // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float p0 = exp2f((s[nt][0] - mnew[0]) * LOG2E);
      float p1 = exp2f((s[nt][1] - mnew[0]) * LOG2E);
      float p2 = exp2f((s[nt][2] - mnew[1]) * LOG2E);
      float p3 = exp2f((s[nt][3] - mnew[1]) * LOG2E);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      const int ks = nt >> 1, hi = nt & 1;
      pf[ks][hi * 2 + 0] = pack_bf16(p0, p1);
      pf[ks][hi * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= corr[0]; o[dt][1] *= corr[0];
      o[dt][2] *= corr[1]; o[dt][3] *= corr[1];
    }

    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide d tiles
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, &sm.v[buf][ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dp * 16 + (lane >> 4) * 8]);
        mma_bf16_16816(o[2 * dp], pf[ks], vf[0], vf[1]);
        mma_bf16_16816(o[2 * dp + 1], pf[ks], vf[2], vf[3]);
      }
    }
    __syncthreads();   // all warps done with buf before it is refilled two iterations later
  }

  // ---- finalize: quad-reduce the row sums, normalise, store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int r0 = qt * BQ + warp * 16 + g, r1 = r0 + 8;
  // destination rows: packed order, or through out_map (-1 = row not needed)
  int64_t d0 = r0 < seq ? (out_map ? out_map[row0 + r0] : row0 + r0) : -1;
  int64_t d1 = r1 < seq ? (out_map ? out_map[row0 + r1] : row0 + r1) : -1;
  __nv_bfloat16* obase = out + h * D;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = dt * 8 + 2 * t;
    if (d0 >= 0) *reinterpret_cast<uint32_t*>(obase + d0 * C + col) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
    if (d1 >= 0) *reinterpret_cast<uint32_t*>(obase + d1 * C + col) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

}  // namespace attn

namespace attn_tc {

constexpr int D = 64;
constexpr int BOX_ROWS = 64;                    // TMA box: 64 rows x 64 bf16 (128 B) = 8 KB
constexpr int BOX_BYTES = BOX_ROWS * D * 2;
constexpr int MAX_SEQ = 448;                    // S (<= 448 fp32 columns) + O (64) fill the 512 TMEM columns
constexpr int PP_MAX_SEQ = 256;                 // ping-pong kernel: two 256-column slots
constexpr int PP_THREADS = 320;                 // warp 0 TMA, warp 1 MMA issue, warps 2-5 / 6-9 softmax of slot 0 / 1
constexpr float LOG2E = 1.4426950408889634f;

// kind::f16 instruction descriptor: D=f32, A=B=bf16, M=128; b_mn = 1 selects an MN-major B operand
__device__ __forceinline__ uint32_t idesc_m128(int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Softmax of one 128-row query tile whose scores S sit in TMEM (thread = row, `lane_base` = TMEM address of
// this warp's lane quarter at the first S column):
//   pass 1: exact row max over the window (no online rescaling at these lengths);
//   pass 2: p = exp2(s log2e - max log2e), fp32 row sum, packed bf16 P written over S with tcgen05.st.
// The TMEM load of chunk c+1 is in flight while chunk c is processed (two named register buffers;
// tcgen05.wait::ld covers every outstanding load); max and sum run on four / two independent accumulators so
// the fixed-latency FMNMX / FADD chains do not serialise a warp that shares its scheduler with one other warp.
// Returns the row sum (0 for warps whose rows are all beyond the window: `active` false, warp-uniform).
// Analytic pad keys (dense blocks, eva_vit.py:249-254): npad further keys of the window have k = 0 exactly (score 0)
// and one common value vector; they are not staged or multiplied: the row max includes the score 0, the row sum
// npad * exp(0 - max), and `corr` (= npad * exp(0 - max)) is the factor of that value vector in the O epilogue.
__device__ __forceinline__ float softmax_rows(uint32_t lane_base, int seq, bool active, float npad, float& corr, int tr_role = -1,
                                              int tr_unit = 0) {
  corr = 0.f;
  if (!active) return 0.f;
  const int nchunks = (seq + 31) >> 5;
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  uint32_t va[32], vb[32];
  auto max_chunk = [&](const uint32_t (&v)[32], int c) {
    const int lim = seq - c * 32;                       // valid columns in this chunk (>= 1)
    if (lim >= 32) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(v[i]));
        m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) m0 = fmaxf(m0, i < lim ? __uint_as_float(v[i]) : -INFINITY);
    }
  };
  tmem_ld_32x32(lane_base, va);
  for (int c = 0; c < nchunks; c += 2) {
    tmem_ld_wait();
    if (c + 1 < nchunks) tmem_ld_32x32(lane_base + (uint32_t)((c + 1) * 32), vb);
    max_chunk(va, c);
    if (c + 1 < nchunks) {
      tmem_ld_wait();
      if (c + 2 < nchunks) tmem_ld_32x32(lane_base + (uint32_t)((c + 2) * 32), va);
      max_chunk(vb, c + 1);
    }
  }
  float mrow = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  if (npad > 0.f) mrow = fmaxf(mrow, 0.f);
  const float mneg = -mrow * LOG2E;
  ATRACE(tr_role >= 0, tr_role, tr_unit, 2);
  float s0 = 0.f, s1 = 0.f;
  auto exp_chunk = [&](const uint32_t (&v)[32], int c) {
    uint32_t pk[16];
    const int lim = seq - c * 32;
    if (lim >= 32) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), LOG2E, mneg));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), LOG2E, mneg));
        s0 += p0;
        s1 += p1;
        pk[i] = pack_bf16(p0, p1);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = 2 * i < lim ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), LOG2E, mneg)) : 0.f;
        const float p1 = 2 * i + 1 < lim ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), LOG2E, mneg)) : 0.f;
        s0 += p0;
        s1 += p1;
        pk[i] = pack_bf16(p0, p1);
      }
    }
    tmem_st_32x16(lane_base + (uint32_t)(c * 16), pk);
  };
  tmem_ld_32x32(lane_base, va);
  for (int c = 0; c < nchunks; c += 2) {
    tmem_ld_wait();
    if (c + 1 < nchunks) tmem_ld_32x32(lane_base + (uint32_t)((c + 1) * 32), vb);
    exp_chunk(va, c);
    if (c + 1 < nchunks) {
      tmem_ld_wait();
      if (c + 2 < nchunks) tmem_ld_32x32(lane_base + (uint32_t)((c + 2) * 32), va);
      exp_chunk(vb, c + 1);
    }
  }
  tmem_st_wait();
  if (npad > 0.f) corr = npad * ex2_approx(mneg);
  return s0 + s1 + corr;
}

// O epilogue of a tile, in two halves so that the slot can be released between them: read the 64 fp32 O columns
// of this thread's row from TMEM ...
__device__ __forceinline__ void load_o(uint32_t o_addr, bool active, uint32_t (&o0)[32], uint32_t (&o1)[32]) {
  if (active) {
    tmem_ld_32x32(o_addr, o0);
    tmem_ld_32x32(o_addr + 32u, o1);
    tmem_ld_wait();
  }
}
// ... and store (O + corr * pad_v) / sum as 64 bf16 (128 bytes of one output row); pad_v = the pad keys' common value
// vector for this head (fp32 [64], nullptr / corr = 0: no analytic pad keys).
__device__ __forceinline__ void store_o(uint32_t (&o0)[32], uint32_t (&o1)[32], float sum, __nv_bfloat16* out_row, float corr,
                                        const float* __restrict__ pad_v) {
  const float inv = 1.0f / sum;
  if (pad_v != nullptr && corr != 0.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(pad_v) + j), b = __ldg(reinterpret_cast<const float4*>(pad_v) + 8 + j);
      o0[4 * j + 0] = __float_as_uint(fmaf(corr, a.x, __uint_as_float(o0[4 * j + 0])));
      o0[4 * j + 1] = __float_as_uint(fmaf(corr, a.y, __uint_as_float(o0[4 * j + 1])));
      o0[4 * j + 2] = __float_as_uint(fmaf(corr, a.z, __uint_as_float(o0[4 * j + 2])));
      o0[4 * j + 3] = __float_as_uint(fmaf(corr, a.w, __uint_as_float(o0[4 * j + 3])));
      o1[4 * j + 0] = __float_as_uint(fmaf(corr, b.x, __uint_as_float(o1[4 * j + 0])));
      o1[4 * j + 1] = __float_as_uint(fmaf(corr, b.y, __uint_as_float(o1[4 * j + 1])));
      o1[4 * j + 2] = __float_as_uint(fmaf(corr, b.z, __uint_as_float(o1[4 * j + 2])));
      o1[4 * j + 3] = __float_as_uint(fmaf(corr, b.w, __uint_as_float(o1[4 * j + 3])));
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out_row);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(o0[8 * j + 0]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
    u.y = pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
    u.z = pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
    u.w = pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
    dst[j] = u;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(o1[8 * j + 0]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
    u.y = pack_bf16(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
    u.z = pack_bf16(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
    u.w = pack_bf16(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
    dst[4 + j] = u;
  }
}

// O epilogue of a tile: await O = P V (bar_o), read it, release the O columns (bar_ofree; and bar_p when the caller
// hands over the next P at the same moment), store the row.
__device__ __forceinline__ void finish_tile(uint32_t o_addr, bool active, bool store, float sum, __nv_bfloat16* out_row,
                                            uint64_t* bar_o, uint64_t* bar_ofree, uint64_t* bar_p, uint32_t parity, float corr,
                                            const float* __restrict__ pad_v) {
  uint32_t o0[32], o1[32];
  mbar_wait(bar_o, parity);
  tcgen05_fence_after();
  load_o(o_addr, active, o0, o1);
  tcgen05_fence_before();
  mbar_arrive(bar_ofree);
  if (bar_p != nullptr) mbar_arrive(bar_p);
  if (store) store_o(o0, o1, sum, out_row, corr, pad_v);
}

// S(tile) = Q_tile K^T into TMEM columns [s_col, s_col + spad): one MMA chain for the first 256 keys, a second
// for the rest.
__device__ __forceinline__ void issue_qk(uint32_t s_addr, const uint8_t* q_tile, const uint8_t* k_base, int spad) {
  const int n1 = spad < 256 ? spad : 256, n2 = spad - n1;
  const uint64_t q_desc = umma_desc_k_sw128(smem_u32(q_tile));
  const uint64_t k_desc = umma_desc_k_sw128(smem_u32(k_base));
  const uint32_t id1 = idesc_m128(n1, 0);
#pragma unroll
  for (int k = 0; k < D / 16; ++k) umma_bf16_ss(s_addr, q_desc + (uint64_t)(2 * k), k_desc + (uint64_t)(2 * k), id1, k);
  if (n2 > 0) {
    const uint64_t k_desc2 = umma_desc_k_sw128(smem_u32(k_base + 256 * 128));
    const uint32_t id2 = idesc_m128(n2, 0);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) umma_bf16_ss(s_addr + 256u, q_desc + (uint64_t)(2 * k), k_desc2 + (uint64_t)(2 * k), id2, k);
  }
}
// O = P V: P packed bf16 in TMEM at p_addr (8 columns per 16 keys), V MN-major in smem (8-key groups 1024 B apart)
__device__ __forceinline__ void issue_pv(uint32_t o_addr, uint32_t p_addr, const uint8_t* v_base, int spad) {
  const uint64_t v_desc = umma_desc_k_sw128(smem_u32(v_base));
  const uint32_t id = idesc_m128(D, 1);
  const int ksteps = spad >> 4;
  for (int kk = 0; kk < ksteps; ++kk) umma_bf16_ts(o_addr, p_addr + (uint32_t)(8 * kk), v_desc + (uint64_t)(128 * kk), id, kk);
}

enum { BAR_QK = 0, BAR_V, BAR_S, BAR_P, BAR_O, BAR_OFREE, NUM_BARS };

// ---------------------------------------------------------------------------------------------------
// Ping-pong kernel (seq <= 256): persistent CTAs, one per SM, looping over (window, head) items.
//   * the Q/K/V boxes of the next items are prefetched by TMA into a ring of item buffers while the
//     current ones are processed;
//   * TMEM holds two 256-column slots; unit u = (item, query tile) runs on slot u & 1, each slot has its own
//     four softmax warps, so the MMAs / barrier round trips of one slot overlap the softmax of the other;
//   * the MMA thread issues S(u) and then P V of unit u-1 (software pipeline of depth 1).
enum { PB_FULL = 0 /* +nbuf */, PB_EMPTY = 4, PB_S = 8 /* +slot */, PB_P = 10, PB_O = 12, PB_OFREE = 14, PP_NUM_BARS = 16 };

template <bool deferred>
__global__ void __launch_bounds__(PP_THREADS, 1)
window_attention_pp_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ out, int seq, int heads,
                           int n_items, int nbuf, const int* __restrict__ out_map, const int* __restrict__ q_rows,
                           const int* __restrict__ item_order, const int* __restrict__ kv_rows, const float* __restrict__ pad_v) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int T = (seq + 127) >> 7;               // 1 or 2 query tiles
  const int nb = (seq + BOX_ROWS - 1) / BOX_ROWS;
  const int item_bytes = (2 * T + 2 * nb) * BOX_BYTES;      // [Q tiles | K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nbuf * item_bytes);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + PP_NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = heads * D;
  const int spad = (seq + 15) & ~15;
  const uint32_t o_off = 192u;                  // O columns inside a 256-column slot (aliases dead S columns if spad > 192)
  // deferred (host: spad <= 192, O disjoint from S / P): the O epilogue runs behind the next tile's softmax
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // query tiles of item i: only the leading q_rows[w] query rows of a window are needed afterwards (the rest are
  // window padding, used as keys / values only), so a window may need fewer tiles than its key count suggests
  auto item_tiles = [&](int i, int& w, int& h, int& need) {
    // item_order (optional): (window, head) items sorted by query-tile count, so that the round-robin deal to the
    // persistent CTAs balances the tile units
    const int idx = (int)blockIdx.x + i * (int)gridDim.x;
    const int it = item_order != nullptr ? item_order[idx] : idx;
    w = it / heads;
    h = it - w * heads;
    need = q_rows != nullptr ? min(seq, max(1, q_rows[w])) : seq;
    return (need + 127) >> 7;
  };
  // keys of window w that are staged and multiplied: all seq slots, or only the leading kv_rows[w] (the rest are the
  // analytic pad keys, see softmax_rows)
  auto item_kv = [&](int w) {
    if (kv_rows == nullptr) return seq;
    const int kv = kv_rows[w];
    return kv >= 1 && kv < seq ? kv : seq;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    for (int b = 0; b < 4; ++b) {
      mbar_init(&bars[PB_FULL + b], 1);
      mbar_init(&bars[PB_EMPTY + b], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[PB_S + s], 1);
      mbar_init(&bars[PB_P + s], 128);
      mbar_init(&bars[PB_O + s], 1);
      mbar_init(&bars[PB_OFREE + s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (elect_one_sync()) {
      // ------------------------------------------------------------------ TMA producer: ring of item buffers
      for (int i = 0; i < my_items; ++i) {
        int w, h, need;
        const int Ti = item_tiles(i, w, h, need);
        const int row0 = w * seq;
        const int buf = i % nbuf;
        const uint32_t round = (uint32_t)(i / nbuf);
        mbar_wait(&bars[PB_EMPTY + buf], (round & 1) ^ 1);
        uint8_t* base = smem + buf * item_bytes;
        uint8_t* sK = base + 2 * T * BOX_BYTES;
        uint8_t* sV = sK + nb * BOX_BYTES;
        const int nbi = (item_kv(w) + BOX_ROWS - 1) / BOX_ROWS;      // K / V boxes of this window
        mbar_arrive_expect_tx(&bars[PB_FULL + buf], (uint32_t)((2 * Ti + 2 * nbi) * BOX_BYTES));
        for (int b = 0; b < nbi; ++b) tma_load_2d(&tm, &bars[PB_FULL + buf], sK + b * BOX_BYTES, C + h * D, row0 + b * BOX_ROWS);
        for (int b = 0; b < 2 * Ti; ++b) tma_load_2d(&tm, &bars[PB_FULL + buf], base + b * BOX_BYTES, h * D, row0 + b * BOX_ROWS);
        for (int b = 0; b < nbi; ++b) tma_load_2d(&tm, &bars[PB_FULL + buf], sV + b * BOX_BYTES, 2 * C + h * D, row0 + b * BOX_ROWS);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ------------------------------------------------------------------ MMA issuer
      // unit u = (item, query tile), enumerated item by item; unit u runs on slot u & 1.  Software pipeline of
      // depth 1: issue S(u), then P V of unit u - 1.
      struct Unit { int s; uint32_t n; const uint8_t* sV; int last_buf; int spad; };   // last_buf >= 0: last tile of its item
      auto do_pv = [&](const Unit& un) {
        mbar_wait(&bars[PB_P + un.s], un.n & 1);
        tcgen05_fence_after();
        ATRACE(true, 2, (int)(2 * un.n) + un.s, 4);        // unit index of `un` = 2 n + slot
        issue_pv(tmem_base + (uint32_t)(un.s * 256) + o_off, tmem_base + (uint32_t)(un.s * 256), un.sV, un.spad);
        tcgen05_commit(&bars[PB_O + un.s]);
        if (un.last_buf >= 0) tcgen05_commit(&bars[PB_EMPTY + un.last_buf]);   // all MMAs reading this item are issued
      };
      Unit prev{0, 0, nullptr, -1, spad};
      bool have_prev = false;
      int u = 0;
      for (int i = 0; i < my_items; ++i) {
        int w, h, need;
        const int Ti = item_tiles(i, w, h, need);
        const int buf = i % nbuf;
        const uint8_t* base = smem + buf * item_bytes;
        const uint8_t* sK = base + 2 * T * BOX_BYTES;
        const uint8_t* sV = sK + nb * BOX_BYTES;
        const int spad_i = (item_kv(w) + 15) & ~15;            // keys of this window, padded to the MMA granularity
        for (int t = 0; t < Ti; ++t, ++u) {
          const int s = u & 1;
          const uint32_t n = (uint32_t)(u >> 1);              // how many units this slot has seen before
          ATRACE(true, 2, u, 0);
          if (t == 0) mbar_wait(&bars[PB_FULL + buf], (uint32_t)((i / nbuf) & 1));
          ATRACE(true, 2, u, 1);
          // in-order mode: the slot (S / P and the O columns that may alias them) must be drained by its softmax
          // warps.  Deferred mode (O never aliases S): S(u) may follow P V (u - 2) directly - the tensor pipe runs
          // this thread's MMAs in issue order - and the O columns are guarded by the P barrier (see below)
          if (!deferred && n > 0) mbar_wait(&bars[PB_OFREE + s], (n - 1) & 1);
          tcgen05_fence_after();
          issue_qk(tmem_base + (uint32_t)(s * 256), base + t * 2 * BOX_BYTES, sK, spad_i);
          tcgen05_commit(&bars[PB_S + s]);
          ATRACE(true, 2, u, 2);
          if (have_prev) do_pv(prev);
          ATRACE(true, 2, u, 3);
          prev = Unit{s, n, sV, t == Ti - 1 ? buf : -1, spad_i};
          have_prev = true;
        }
      }
      if (have_prev) do_pv(prev);
    }
  } else {
    // -------------------------------------------------------------------- softmax warps: slot = (warp - 2) / 4
    const int slot = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 256);
    uint64_t* bar_s = &bars[PB_S + slot];
    uint64_t* bar_p = &bars[PB_P + slot];
    uint64_t* bar_o = &bars[PB_O + slot];
    uint64_t* bar_ofree = &bars[PB_OFREE + slot];
    // In-order mode: S -> softmax -> P -> (P V) -> O epilogue, tile by tile.  Deferred mode: the epilogue of the
    // slot's previous tile runs after the softmax of the current one, so the slot never idles through P V + O read
    // + S(next): it only waits for S, which was issued right behind P V.
    bool pend = false, pend_act = false, pend_ok = false;
    float pend_sum = 0.f, pend_corr = 0.f;
    __nv_bfloat16* pend_row = out;
    const float* pend_pv = nullptr;
    uint32_t pend_par = 0;
    auto finish = [&](bool then_p) {
      finish_tile(lane_base + o_off, pend_act, pend_act && pend_ok, pend_sum, pend_row, bar_o, bar_ofree, then_p ? bar_p : nullptr,
                  pend_par, pend_corr, pend_pv);
    };
    int u = 0;
    for (int i = 0; i < my_items; ++i) {
      int w, h, need;
      const int Ti = item_tiles(i, w, h, need);
      const int kv = item_kv(w);
      for (int t = 0; t < Ti; ++t, ++u) {
        if ((u & 1) != slot) continue;
        const int q = t * 128 + quarter * 32 + lane;
        int dst = q < need ? w * seq + q : -1;
        if (dst >= 0 && out_map != nullptr) dst = out_map[dst];
        const bool active = t * 128 + quarter * 32 < need;      // a warp whose 32 query rows are all unneeded skips the tile
        const uint32_t parity = (uint32_t)((u >> 1) & 1);
        const bool tr = quarter == 0 && lane == 0;
        ATRACE(tr, slot, u >> 1, 0);
        mbar_wait(bar_s, parity);
        tcgen05_fence_after();
        ATRACE(tr, slot, u >> 1, 1);
        float corr;
        const float sum = softmax_rows(lane_base, kv, active, (float)(seq - kv), corr, tr ? slot : -1, u >> 1);
        ATRACE(tr, slot, u >> 1, 3);
        if (deferred) {
          if (pend) {
            finish(true);                                // O(previous) completed long ago: its P V ran before this S
          } else {
            tcgen05_fence_before();
            mbar_arrive(bar_p);
          }
        } else {
          tcgen05_fence_before();
          mbar_arrive(bar_p);
        }
        ATRACE(tr, slot, u >> 1, 4);
        pend = true; pend_act = active; pend_ok = dst >= 0; pend_sum = sum; pend_par = parity; pend_corr = corr;
        pend_pv = pad_v != nullptr ? pad_v + h * D : nullptr;
        pend_row = out + (size_t)(dst < 0 ? 0 : dst) * C + h * D;
        if (!deferred) {
          finish(false);
          pend = false;
        }
        ATRACE(tr, slot, u >> 1, 5);
      }
    }
    if (pend) finish(false);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------
// Split softmax (used by the single-slot kernel below).  With ONE softmax warp per TMEM lane quarter the ex2 pass is
// latency-bound (MUFU 33 % busy).  Here TWO warps share a lane quarter (warp_id % 4 selects both the TMEM
// lane quarter and the scheduler, so they sit on the same sub-partition and fill each other's MUFU / TMEM-load
// latencies) and split the KEY columns:
//   half 0: 32-key chunks [0, c0) in ascending order,  P chunk c packed at columns [16 c, 16 c + 16)
//   half 1: chunks [c0, n) in DESCENDING order,         P chunk c packed at columns [16 (n + c), 16 (n + c) + 16)
// (n = ceil(seq / 32), c0 = ceil(n / 2)).  A P chunk always lands on score columns its own warp has already read
// (half 0: below 32 (c + 1); half 1: at or above 32 c) and never in the other half's score range, so the two warps
// need no ordering between their passes except the exchange of the row max (and of the row sum) through shared
// memory + a 64-thread named barrier.  Each half reads and stores 32 of the 64 O columns of its rows.  (Measured on
// B200, profiles/r02a_attn_bench_*.txt: 18 x 400 keys 55.5 -> 42.6 us, 18 x 281 33.0 -> 26.6 us.  The same split on the
// two 256-column slots of the ping-pong kernel was 0-12 % SLOWER - two slots already give each scheduler two softmax
// warps - and was removed.)

__device__ __forceinline__ uint32_t p_col_split(int c, int c0, int n) { return c < c0 ? 16u * (uint32_t)c : 16u * (uint32_t)(n + c); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// O = P V with the split P layout: k-step kk covers keys [16 kk, 16 kk + 16) = half a 32-key chunk
__device__ __forceinline__ void issue_pv_split(uint32_t o_addr, uint32_t slot_addr, const uint8_t* v_base, int spad, int c0, int n) {
  const uint64_t v_desc = umma_desc_k_sw128(smem_u32(v_base));
  const uint32_t id = idesc_m128(D, 1);
  const int ksteps = spad >> 4;
  for (int kk = 0; kk < ksteps; ++kk)
    umma_bf16_ts(o_addr, slot_addr + p_col_split(kk >> 1, c0, n) + (uint32_t)(8 * (kk & 1)), v_desc + (uint64_t)(128 * kk), id, kk);
}

// One warp's share of the softmax of a 128-row tile (thread = row): its key chunks only; row max and row sum are
// combined with the partner warp of the lane quarter through xmax / xsum (indexed [half][lane]) and barrier bar_id.
// Returns the full row sum.  `active` is the same for both partners (same rows).
__device__ __forceinline__ float softmax_half(uint32_t lane_base, int seq, int n, int c0, int half, int lane, bool active,
                                              float* xmax, float* xsum, int bar_id, float npad, float& corr) {
  corr = 0.f;
  if (!active) return 0.f;
  const int cb = half ? c0 : 0, ce = half ? n : c0;
  uint32_t v[32];
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  for (int c = cb; c < ce; ++c) {
    tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    const int lim = seq - c * 32;
    if (lim >= 32) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(v[i]));
        m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) m0 = fmaxf(m0, i < lim ? __uint_as_float(v[i]) : -INFINITY);
    }
  }
  float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  xmax[half * 32 + lane] = mx;
  named_bar_sync(bar_id, 64);
  mx = fmaxf(mx, xmax[(half ^ 1) * 32 + lane]);          // finite: chunk 0 holds at least one valid key
  if (npad > 0.f) mx = fmaxf(mx, 0.f);                   // analytic pad keys (score 0), see softmax_rows
  const float mneg = -mx * LOG2E;
  float s0 = 0.f, s1 = 0.f;
  for (int j = cb; j < ce; ++j) {
    const int c = half ? (ce - 1 - (j - cb)) : j;        // half 1 walks its chunks from the top
    tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    uint32_t pk[16];
    const int lim = seq - c * 32;
    if (lim >= 32) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), LOG2E, mneg));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), LOG2E, mneg));
        s0 += p0;
        s1 += p1;
        pk[i] = pack_bf16(p0, p1);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = 2 * i < lim ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), LOG2E, mneg)) : 0.f;
        const float p1 = 2 * i + 1 < lim ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), LOG2E, mneg)) : 0.f;
        s0 += p0;
        s1 += p1;
        pk[i] = pack_bf16(p0, p1);
      }
    }
    tmem_st_32x16(lane_base + p_col_split(c, c0, n), pk);
  }
  tmem_st_wait();
  const float sum = s0 + s1;
  xsum[half * 32 + lane] = sum;
  named_bar_sync(bar_id, 64);                           // also: both halves of P are in TMEM
  if (npad > 0.f) corr = npad * ex2_approx(mneg);
  return sum + xsum[(half ^ 1) * 32 + lane] + corr;
}

// this warp's 32 of the 64 O columns of a row -> 64 bytes of the output row
__device__ __forceinline__ void store_o_half(uint32_t (&o)[32], float sum, __nv_bfloat16* out_half_row, float corr,
                                             const float* __restrict__ pad_v_half) {
  const float inv = 1.0f / sum;
  if (pad_v_half != nullptr && corr != 0.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(pad_v_half) + j);
      o[4 * j + 0] = __float_as_uint(fmaf(corr, a.x, __uint_as_float(o[4 * j + 0])));
      o[4 * j + 1] = __float_as_uint(fmaf(corr, a.y, __uint_as_float(o[4 * j + 1])));
      o[4 * j + 2] = __float_as_uint(fmaf(corr, a.z, __uint_as_float(o[4 * j + 2])));
      o[4 * j + 3] = __float_as_uint(fmaf(corr, a.w, __uint_as_float(o[4 * j + 3])));
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out_half_row);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
    u.y = pack_bf16(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
    u.z = pack_bf16(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
    u.w = pack_bf16(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
    dst[j] = u;
  }
}

// Single-slot kernel (256 < seq <= 448): one CTA per (window, head), query tiles in sequence, split softmax: two warps
// share a lane quarter and split the key columns (layout and exchange as above); S occupies [0, spad) <= 448 and O the
// last 64 of the 512 columns, so nothing aliases.  warp 0: TMA + MMA issue; warps 1-4: half 0; warps 5-8: half 1.
constexpr int TC2_THREADS = 32 + 8 * 32;
constexpr int TC2_XCH_BYTES = 2 * 4 * 2 * 32 * 4;   // {max, sum} x quarter x half x lane

__global__ void __launch_bounds__(TC2_THREADS)
window_attention_tc2_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ out, int seq, int heads,
                            const int* __restrict__ out_map, const int* __restrict__ q_rows, const int* __restrict__ kv_rows,
                            const float* __restrict__ pad_v) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int Tmax = (seq + 127) >> 7;
  const int nb = (seq + BOX_ROWS - 1) / BOX_ROWS;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * Tmax * BOX_BYTES;
  uint8_t* sV = sK + nb * BOX_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + nb * BOX_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  float* xch = reinterpret_cast<float*>(sV + nb * BOX_BYTES + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, w = blockIdx.y;
  const int need = q_rows != nullptr ? min(seq, max(1, q_rows[w])) : seq;      // query rows that are used afterwards
  const int T = (need + 127) >> 7;
  const int C = heads * D;
  const int row0 = w * seq;
  constexpr int tmem_cols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(&bars[BAR_QK], 1);
    mbar_init(&bars[BAR_V], 1);
    mbar_init(&bars[BAR_S], 1);
    mbar_init(&bars[BAR_P], 256);
    mbar_init(&bars[BAR_O], 1);
    mbar_init(&bars[BAR_OFREE], 256);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, (uint32_t)tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch_dependents();

  // keys that are staged and multiplied: all seq slots, or the leading kv_rows[w] (the rest: analytic pad keys)
  int kv = seq;
  if (kv_rows != nullptr) {
    const int r = kv_rows[w];
    if (r >= 1 && r < seq) kv = r;
  }
  const float npad = (float)(seq - kv);
  const int nbk = (kv + BOX_ROWS - 1) / BOX_ROWS;
  const int spad = (kv + 15) & ~15;
  const int n = (kv + 31) >> 5;
  const int c0 = (n + 1) >> 1;
  const uint32_t o_col = (uint32_t)tmem_cols - 64u;          // host: spad <= 448, so O never aliases S / P

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&bars[BAR_QK], (uint32_t)((2 * T + nbk) * BOX_BYTES));
      for (int b = 0; b < nbk; ++b) tma_load_2d(&tm, &bars[BAR_QK], sK + b * BOX_BYTES, C + h * D, row0 + b * BOX_ROWS);
      for (int b = 0; b < 2 * T; ++b) tma_load_2d(&tm, &bars[BAR_QK], sQ + b * BOX_BYTES, h * D, row0 + b * BOX_ROWS);
      mbar_arrive_expect_tx(&bars[BAR_V], (uint32_t)(nbk * BOX_BYTES));
      for (int b = 0; b < nbk; ++b) tma_load_2d(&tm, &bars[BAR_V], sV + b * BOX_BYTES, 2 * C + h * D, row0 + b * BOX_ROWS);
      mbar_wait(&bars[BAR_QK], 0);
      tcgen05_fence_after();
      issue_qk(tmem_base, sQ, sK, spad);
      tcgen05_commit(&bars[BAR_S]);
      for (int t = 0; t < T; ++t) {
        mbar_wait(&bars[BAR_P], t & 1);                     // both halves of P(t) are in TMEM (over S(t))
        if (t == 0) mbar_wait(&bars[BAR_V], 0);
        else mbar_wait(&bars[BAR_OFREE], (t - 1) & 1);      // O(t-1) has been read out
        tcgen05_fence_after();
        issue_pv_split(tmem_base + o_col, tmem_base, sV, spad, c0, n);
        tcgen05_commit(&bars[BAR_O]);
        if (t + 1 < T) {
          issue_qk(tmem_base, sQ + (t + 1) * 2 * BOX_BYTES, sK, spad);   // executes after PV(t): may overwrite P(t)
          tcgen05_commit(&bars[BAR_S]);
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 1) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* xmax = xch + (quarter * 2) * 32;
    float* xsum = xmax + 4 * 2 * 32;
    for (int t = 0; t < T; ++t) {
      const int q = t * 128 + quarter * 32 + lane;
      int dst = q < need ? row0 + q : -1;
      if (dst >= 0 && out_map != nullptr) dst = out_map[dst];
      const bool active = t * 128 + quarter * 32 < need;       // warps whose 32 query rows are all unneeded skip the tile
      const uint32_t parity = (uint32_t)(t & 1);
      mbar_wait(&bars[BAR_S], parity);
      tcgen05_fence_after();
      float corr;
      const float sum = softmax_half(lane_base, kv, n, c0, half, lane, active, xmax, xsum, 1 + quarter, npad, corr);
      tcgen05_fence_before();
      mbar_arrive(&bars[BAR_P]);
      mbar_wait(&bars[BAR_O], parity);
      tcgen05_fence_after();
      uint32_t o[32];
      if (active) {
        tmem_ld_32x32(lane_base + o_col + (uint32_t)(half * 32), o);
        tmem_ld_wait();
      }
      tcgen05_fence_before();
      mbar_arrive(&bars[BAR_OFREE]);
      if (active && dst >= 0)
        store_o_half(o, sum, out + (size_t)dst * C + h * D + half * 32, corr, pad_v != nullptr ? pad_v + h * D + half * 32 : nullptr);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

}  // namespace attn_tc

#ifdef TOC3D_ATTN_TRACE
extern "C" int toc3d_attn_trace_read(unsigned long long* host, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(unsigned long long) * (size_t)n);
}
#endif
}  // namespace toc3d

extern "C" int toc3d_window_attention(const void* qkv, void* out, int32_t n_windows, int32_t seq_len, int32_t heads,
                                      const int32_t* out_map, const int32_t* q_rows, const int32_t* item_order,
                                      const int32_t* kv_rows, const float* pad_v, void* stream) {
  using namespace toc3d;
  TOC3D_REQUIRE(qkv && out, kErrBadArg, "toc3d_window_attention: null pointer");
  TOC3D_REQUIRE((kv_rows == nullptr) == (pad_v == nullptr), kErrBadArg, "toc3d_window_attention: kv_rows and pad_v go together");
  TOC3D_REQUIRE(kv_rows == nullptr || (seq_len <= attn_tc::MAX_SEQ && ((uintptr_t)pad_v & 15) == 0), kErrBadArg,
                "toc3d_window_attention: analytic pad keys need seq <= %d and a 16-byte aligned pad_v", attn_tc::MAX_SEQ);
  TOC3D_REQUIRE(n_windows > 0 && seq_len > 0 && seq_len <= 1024 && heads > 0 && heads <= 65535, kErrBadArg,
                "toc3d_window_attention: bad shape nW=%d seq=%d heads=%d", n_windows, seq_len, heads);
  TOC3D_REQUIRE(n_windows <= 65535, kErrBadArg, "toc3d_window_attention: too many windows (%d)", n_windows);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (seq_len <= attn_tc::MAX_SEQ) {
    // one-time setup per device ordinal (several GPUs in one process)
    static bool configured_dev[64] = {};
    static int n_sm_dev[64] = {};
    static std::mutex cfg_mutex;
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (dev_ < 0 || dev_ >= 64) dev_ = 0;
    int n_sm = 148;
    {
    std::lock_guard<std::mutex> lock(cfg_mutex);
    bool& configured = configured_dev[dev_];
    if (!configured) {
      TOC3D_CHECK_CUDA(cudaFuncSetAttribute(attn_tc::window_attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            227 * 1024));
      TOC3D_CHECK_CUDA(cudaFuncSetAttribute(attn_tc::window_attention_pp_kernel<false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      TOC3D_CHECK_CUDA(cudaFuncSetAttribute(attn_tc::window_attention_pp_kernel<true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      int v = 0;
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev_);
      n_sm_dev[dev_] = v > 0 ? v : 148;
      configured = true;
    }
    n_sm = n_sm_dev[dev_];
    }
    const int C = heads * attn_tc::D;
    CUtensorMap tm;
    int rc = make_tmap_bf16_2d(&tm, qkv, (int64_t)n_windows * seq_len, 3 * (int64_t)C, 3 * (int64_t)C, attn_tc::BOX_ROWS);
    if (rc) return rc;
    const int T = (seq_len + 127) / 128, nb = (seq_len + 63) / 64;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (seq_len <= attn_tc::PP_MAX_SEQ) {
      const int item_bytes = (2 * T + 2 * nb) * attn_tc::BOX_BYTES;
      const int n_items = n_windows * heads;
      const int grid = n_items < n_sm ? n_items : n_sm;
      int nbuf = (226 * 1024 - 1024 - 256) / item_bytes;
      nbuf = nbuf > 4 ? 4 : nbuf;                       // >= 2 for seq <= 256 (96 KB per item)
      const size_t smem = (size_t)nbuf * item_bytes + 1024 + 256;
      // keys (padded to 16) <= 192: the O columns do not alias S, the epilogue is deferred behind the next softmax
      if (((seq_len + 15) & ~15) <= 192) {
        TOC3D_CHECK_CUDA(launch_pdl(attn_tc::window_attention_pp_kernel<true>, dim3(grid), dim3(attn_tc::PP_THREADS), smem, st, 1,
                                    tm, o, seq_len, heads, n_items, nbuf, out_map, q_rows, item_order, kv_rows, pad_v));
      } else {
        TOC3D_CHECK_CUDA(launch_pdl(attn_tc::window_attention_pp_kernel<false>, dim3(grid), dim3(attn_tc::PP_THREADS), smem, st, 1,
                                    tm, o, seq_len, heads, n_items, nbuf, out_map, q_rows, item_order, kv_rows, pad_v));
      }
      return 0;
    }
    const size_t smem2 = (size_t)(2 * T + 2 * nb) * attn_tc::BOX_BYTES + 1024 + 128 + attn_tc::TC2_XCH_BYTES;
    TOC3D_CHECK_CUDA(launch_pdl(attn_tc::window_attention_tc2_kernel, dim3(heads, n_windows), dim3(attn_tc::TC2_THREADS), smem2,
                                st, 1, tm, o, seq_len, heads, out_map, q_rows, kv_rows, pad_v));
    return 0;
  }
  dim3 grid((seq_len + attn::BQ - 1) / attn::BQ, heads, n_windows);
  TOC3D_CHECK_CUDA(launch_pdl(attn::window_attention_kernel, grid, dim3(attn::NTHREADS), 0, st, 1,
                              reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), seq_len, heads, out_map, q_rows));
  return 0;
}
