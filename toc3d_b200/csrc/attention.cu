// Windowed multi-head attention, head dim 64, sequence = one window (<= 401 tokens in the shipped
// configs).  Three kernels behind toc3d_window_attention:
//   attn_tc::window_attention_pp_kernel (seq <= 256): tcgen05 / TMEM / TMA, persistent, one CTA per SM looping over
//            (window, head) items.  Q, K, V of an item are staged by TMA (128B swizzle) into a ring of smem buffers while
//            earlier items are processed.  TMEM holds two 256-column slots; per 128-row query tile S = Q K^T is one
//            tcgen05.mma chain into a slot, the slot's four softmax warps (thread = row) read S with tcgen05.ld, take the
//            exact row max over the whole window (no online rescaling at these lengths), write P = exp2(.) as packed bf16
//            back over S with tcgen05.st, and O = P V runs as tcgen05.mma with the A operand in TMEM and V as an MN-major
//            smem operand.  The MMA thread issues S(u) then P V(u-1), so one slot's MMAs overlap the other slot's
//            softmax; four epilogue warps read O and store the rows.
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
__device__ unsigned long long g_attn_trace[4 * 32 * 8];      // role 3 = TMA producer (per item)
#define ATRACE(cond, role, unit, k)                                                                    \
  do {                                                                                                 \
    if ((cond) && blockIdx.x == 0 && (unit) < 32) g_attn_trace[((role) * 32 + (unit)) * 8 + (k)] = clock64(); \
  } while (0)
// per-chunk stamps of the softmax passes of unit 1 of each slot: [slot][pass][chunk]
__device__ unsigned long long g_attn_chunk[2 * 2 * 16];
#define ATRACE_CHUNK(cond, role, pass, c)                                                               \
  do {                                                                                                 \
    if ((cond) && blockIdx.x == 0 && (c) < 16) g_attn_chunk[((role) * 2 + (pass)) * 16 + (c)] = clock64(); \
  } while (0)
// stamps inside the epilogue of the pipelined kernel: [half][unit][8]
__device__ unsigned long long g_attn_fin[2 * 32 * 8];
#define ATRACE_FIN(cond, role, unit, k)                                                                \
  do {                                                                                                 \
    if ((cond) && blockIdx.x == 0 && (unit) < 32) g_attn_fin[((role) * 32 + (unit)) * 8 + (k)] = clock64(); \
  } while (0)
#else
#define ATRACE(cond, role, unit, k) do {} while (0)
#define ATRACE_CHUNK(cond, role, pass, c) do {} while (0)
#define ATRACE_FIN(cond, role, unit, k) do {} while (0)
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

// exp2 on the FMA pipe (Cody-Waite reduction + cubic minimax, 7.5e-5 relative error: 26 x below the bf16 rounding of P):
// the MUFU pipe does 4 ex2 per clock and scheduler, and with two softmax warps per scheduler the exp pass runs at that
// limit (probe: 295 clk per 32-key chunk alone, 530 with two warps).  Every fourth probability takes this path instead, as
// in FlashAttention-4.  x <= 0 here (scores minus the row max); below -126 the result is clamped to 2^-126.
// Measured (profiles/r02an): the single-slot kernel (two softmax warps per scheduler, nothing else) gains 3-5 %
// (18 x 400 keys 42.0 -> 40.1 us); the persistent kernel, whose schedulers also carry an epilogue warp, is issue-bound
// and LOSES 9 % (48 x 129: 18.2 -> 19.9 us), so only softmax_half uses it.
#ifndef TOC3D_EXP_POLY
#define TOC3D_EXP_POLY 1
#endif
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;           // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float r = x - (t - 12582912.f);     // [-0.5, 0.5]
  float p = fmaf(0.0551716648f, r, 0.2426111251f);
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float ex2_mixed(float x, int i) {      // i: compile-time position in an unrolled loop
  return TOC3D_EXP_POLY && (i & 3) == 3 ? ex2_poly(x) : ex2_approx(x);
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
  if (npad > 0.f) mx = fmaxf(mx, 0.f);                   // analytic pad keys: score 0
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
        const float p1 = ex2_mixed(fmaf(__uint_as_float(v[2 * i + 1]), LOG2E, mneg), 2 * i + 1);
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

// ---------------------------------------------------------------------------------------------------
// Ping-pong kernel (seq <= 256): persistent CTAs, one per SM, looping over (window, head) items.
//   * 14 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 / 6-9 softmax of TMEM slot 0 / 1 (thread = row,
//     quarter = warp % 4), warps 10-13 epilogue (one per TMEM lane quarter, both slots).
//   * TMEM holds two 256-column slots; unit u = (item, 128-row query tile) runs on slot u & 1, so the MMAs and barrier round
//     trips of one slot overlap the softmax of the other.  The MMA thread issues S(u), then P V of unit u - 1.
//   * the Q / K / V rows of the next items are prefetched by TMA into a ring of item buffers; an item holds whole query
//     tiles and ceil16(seq) rows of K and V, loaded as one box per matrix when all rows are wanted and as 64- / 16-row
//     boxes when only the leading rows are (needed query rows, real keys of a ragged window);
//   * the softmax warps do nothing but the two passes (exact row max, then p = exp2(.) packed over S) and hand the row
//     sum to the epilogue warp of their quarter through shared memory; a last chunk of <= 16 keys (129 = 128 slow tokens
//     + representative) takes a 16-column load / 8-column store instead of a predicated 32-column pass;
//   * the epilogue warps wait for O = P V, read its 64 fp32 columns, release them (OFREE), and store (O + corr pad_v) / sum
//     as bf16 through a swizzled staging tile, so that every store instruction writes 4 complete 128-byte row segments;
//   * item metadata (window, head, needed rows, keys) is staged in shared memory by all threads BEFORE the grid dependency
//     resolves: no role chases item_order -> q_rows -> kv_rows through global memory per unit.
// O lives in columns [192, 256) of its slot.  With more than 192 key columns it aliases dead score columns (P occupies the
// first half of S by then), and S(u) has to wait for the epilogue of unit u - 2; otherwise only P V(u) does.
// History (clock64 timelines of CTA 0 and A/B timings on the step's real shapes, profiles/r02a*_attn_*): the round-1
// kernel ran the O epilogue (1300-1700 clk per unit, scattered 16-byte stores) and three dependent global loads of item
// metadata (~1000 clk) in the softmax warps' chain: 9000 clk per slot and unit at 129 keys.  Moving both out gave 48 x 129
// keys 21.1 -> 18.2 us, 48 x 180 23.8 -> 22.5, 48 x 256 (ragged dense windows) 23.0 -> 21.7, 18 x 201 15.5 -> 13.7.
// Built, verified and NOT adopted on the way: all eight softmax warps on ONE unit with the key columns split between the
// two warps of a quarter and S(u + 1) issued a unit ahead (no S wait, but a max exchange per unit and every warp in every
// unit: 20.9-26.0 us, no better than round 1); rotating partial tiles to the least-used lane quarters (a unit's P V still
// needs all quarters of its slot's previous unit: 21.1 -> 23.4 us); exp2 on the FMA pipe for every fourth probability
// (issue-bound here: 18.2 -> 19.9 us; kept in the single-slot kernel, where it wins).

// 32-column chunks first .. first + cnt - 1 of a row, the load of the next chunk in flight behind the current one
template <typename F>
__device__ __forceinline__ void for_chunks(uint32_t lane_base, int cnt, F&& body, int tr_role = -1, int tr_pass = 0) {
  if (cnt <= 0) return;
  uint32_t va[32], vb[32];
  ATRACE_CHUNK(tr_role >= 0, tr_role, tr_pass, 0);
  tmem_ld_32x32(lane_base, va);
  for (int c = 0; c < cnt; c += 2) {
    tmem_ld_wait();
    if (c + 1 < cnt) tmem_ld_32x32(lane_base + (uint32_t)((c + 1) * 32), vb);
    body(va, c);
    ATRACE_CHUNK(tr_role >= 0, tr_role, tr_pass, c + 1);
    if (c + 1 < cnt) {
      tmem_ld_wait();
      if (c + 2 < cnt) tmem_ld_32x32(lane_base + (uint32_t)((c + 2) * 32), va);
      body(vb, c + 1);
      ATRACE_CHUNK(tr_role >= 0, tr_role, tr_pass, c + 2);
    }
  }
}

struct RowChunks {         // the key columns of a row as 32-column chunks
  int wide;                // chunks read with a 32-column load (the last of them may be partly valid)
  int tail;                // index of a narrow (<= 16 keys) last chunk, or -1
  int tail_lim;            // its valid columns
};
__device__ __forceinline__ RowChunks row_chunks(int kv) {
  const int n = (kv + 31) >> 5;
  const int tl = kv - (n - 1) * 32;                         // keys in the last chunk (1 .. 32)
  RowChunks rc;
  rc.tail = tl <= 16 ? n - 1 : -1;
  rc.tail_lim = tl;
  rc.wide = tl <= 16 ? n - 1 : n;
  return rc;
}

// exact row max over the kv key columns of this thread's row (no online rescaling at these lengths)
__device__ __forceinline__ float row_max(uint32_t lane_base, int kv, const RowChunks& rc, int tr_role = -1) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  if (rc.tail >= 0) {
    uint32_t v[16];
    tmem_ld_32x16(lane_base + (uint32_t)(rc.tail * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) m0 = fmaxf(m0, i < rc.tail_lim ? __uint_as_float(v[i]) : -INFINITY);
  }
  for_chunks(lane_base, rc.wide, [&](const uint32_t (&v)[32], int c) {
    const int lim = kv - c * 32;                            // valid columns in this chunk (>= 1)
    if (lim >= 32) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {                     // four accumulators: the FMNMX chains do not serialise
        m0 = fmaxf(m0, __uint_as_float(v[i]));
        m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) m0 = fmaxf(m0, i < lim ? __uint_as_float(v[i]) : -INFINITY);
    }
  }, tr_role, 0);
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// p = exp2(s log2e + mneg), packed bf16 P written over S with tcgen05.st (chunk c -> columns [16 c, 16 c + 16), i.e. over
// scores that are already in registers); returns the fp32 row sum.  The narrow tail goes last: its P lands on score
// columns of chunk (n - 1) / 2.
__device__ __forceinline__ float row_exp(uint32_t lane_base, int kv, const RowChunks& rc, float mneg, int tr_role = -1) {
  float s0 = 0.f, s1 = 0.f;
  for_chunks(lane_base, rc.wide, [&](const uint32_t (&v)[32], int c) {
    uint32_t pk[16];
    const int lim = kv - c * 32;
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
  }, tr_role, 1);
  if (rc.tail >= 0) {
    uint32_t v[16], pk[8];
    tmem_ld_32x16(lane_base + (uint32_t)(rc.tail * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = 2 * i < rc.tail_lim ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), LOG2E, mneg)) : 0.f;
      const float p1 = 2 * i + 1 < rc.tail_lim ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), LOG2E, mneg)) : 0.f;
      s0 += p0;
      s1 += p1;
      pk[i] = pack_bf16(p0, p1);
    }
    tmem_st_32x8(lane_base + (uint32_t)(rc.tail * 16), pk);
  }
  tmem_st_wait();
  return s0 + s1;
}

enum { PB_FULL = 0 /* +nbuf */, PB_EMPTY = 4, PB_S = 8 /* +slot */, PB_P = 10, PB_O = 12, PB_OFREE = 14, PP_NUM_BARS = 16 };
constexpr int PP_MAX_ITEMS = 64;                // items of a CTA whose metadata is staged in shared memory at kernel start
constexpr int PP_CTL_HEAD = 256 + PP_MAX_ITEMS * 16;   // barriers (128 B), TMEM base (16 B), item list
constexpr int PP_THREADS = 14 * 32;
// exchange area (floats): row sum and pad-key factor, [unit & 3][quarter][lane] each
constexpr int PP_XSUM = 0, PP_XCORR = 4 * 128;
constexpr int PP_XCH_BYTES = (PP_XCORR + 4 * 128) * 4;
constexpr uint32_t PP_O_OFF = 192u;             // O columns inside a 256-column slot
constexpr int PP_STAGE_BYTES = 4 * 32 * 128;    // per epilogue warp: 32 rows x 128 B of bf16 output
constexpr int PP_TAIL_ROWS = 16;                // second tensor map: 16-row boxes for ragged tails

// bytes that pp_load_rows will make arrive
__device__ __forceinline__ uint32_t pp_row_bytes(int rows, int seq) {
  if (rows == seq) return (uint32_t)(((seq + 15) & ~15) * 128);
  const int nfull = rows >> 6, rem = rows & 63;
  if (rem > BOX_ROWS - PP_TAIL_ROWS) return (uint32_t)((nfull + 1) * BOX_BYTES);
  return (uint32_t)(nfull * BOX_BYTES + ((rem + 15) >> 4) * (PP_TAIL_ROWS * 128));
}
// rows [0, rows) of one of Q / K / V of an item.  All seq rows: ONE box of ceil16(seq) rows (third tensor map).  Fewer
// (needed query rows / real keys of a ragged window): 64-row boxes, then 16-row boxes for the tail (the smem region of an
// item holds ceil16(seq) rows per matrix, so a tail is never rounded up to a 64-row box unless it fills one).
__device__ __forceinline__ void pp_load_rows(const CUtensorMap* tm64, const CUtensorMap* tm16, const CUtensorMap* tmF, uint64_t* bar,
                                             uint8_t* dst, int col, int row0, int rows, int seq) {
  if (rows == seq) {
    tma_load_2d(tmF, bar, dst, col, row0);
    return;
  }
  const int nfull = rows >> 6, rem = rows & 63;
  int r = 0;
  for (int b = 0; b < nfull; ++b, r += BOX_ROWS) tma_load_2d(tm64, bar, dst + r * 128, col, row0 + r);
  if (rem > BOX_ROWS - PP_TAIL_ROWS) {
    tma_load_2d(tm64, bar, dst + r * 128, col, row0 + r);
  } else {
    for (; r < rows; r += PP_TAIL_ROWS) tma_load_2d(tm16, bar, dst + r * 128, col, row0 + r);
  }
}

__global__ void __launch_bounds__(PP_THREADS, 1)
window_attention_pp_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm16,
                           const __grid_constant__ CUtensorMap tmF, __nv_bfloat16* __restrict__ out, int seq, int heads, int n_items, int nbuf,
                           const int* __restrict__ out_map, const int* __restrict__ q_rows, const int* __restrict__ item_order,
                           const int* __restrict__ kv_rows, const float* __restrict__ pad_v) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int mat_bytes = ((seq + 15) & ~15) * 128;           // K or V of an item: ceil16(seq) rows of 128 B
  const int q_bytes = ((seq + 127) >> 7) * 128 * 128;       // Q: whole 128-row tiles (the MMA reads all 128 rows of a tile)
  const int item_bytes = q_bytes + 2 * mat_bytes;           // [Q | K | V]
  uint8_t* ctl = smem + nbuf * item_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctl);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + PP_NUM_BARS);
  int4* items = reinterpret_cast<int4*>(ctl + 256);                      // {window, head, needed rows, keys} of this CTA's first items
  float* xch = reinterpret_cast<float*>(ctl + PP_CTL_HEAD);
  uint8_t* stage_all = ctl + PP_CTL_HEAD + PP_XCH_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = heads * D;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // {window, head, needed query rows, staged keys} of item i of this CTA
  auto item_meta = [&](int i) {
    const int idx = (int)blockIdx.x + i * (int)gridDim.x;
    const int it = item_order != nullptr ? item_order[idx] : idx;
    const int w = it / heads, h = it - w * heads;
    const int need = q_rows != nullptr ? min(seq, max(1, q_rows[w])) : seq;      // leading query rows that are used
    int kv = seq;                                                               // keys staged and multiplied
    if (kv_rows != nullptr) {
      const int r = kv_rows[w];
      if (r >= 1 && r < seq) kv = r;
    }
    return make_int4(w, h, need, kv);
  };
  // The item tables (item_order, q_rows, kv_rows) are inputs that no kernel writes (include/toc3d_b200.h): they are read
  // BEFORE the grid dependency resolves, one item per thread, so the three dependent global loads per item (~2000 clk)
  // overlap the previous kernel's tail instead of sitting in the TMA producer's loop.
  if ((int)threadIdx.x < my_items && threadIdx.x < PP_MAX_ITEMS) items[threadIdx.x] = item_meta((int)threadIdx.x);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm16);
    tma_prefetch_desc(&tmF);
    for (int b = 0; b < 4; ++b) {
      mbar_init(&bars[PB_FULL + b], 1);
      mbar_init(&bars[PB_EMPTY + b], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[PB_S + s], 1);
      mbar_init(&bars[PB_P + s], 128);           // the four softmax warps of the slot
      mbar_init(&bars[PB_O + s], 1);
      mbar_init(&bars[PB_OFREE + s], 128);       // the four epilogue warps
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
      // ------------------------------------------------------------------ TMA producer: ring of item buffers + metadata
      for (int i = 0; i < my_items; ++i) {
        const int4 md = i < PP_MAX_ITEMS ? items[i] : item_meta(i);
        const int w = md.x, h = md.y, need = md.z, kv = md.w;
        const int row0 = w * seq;
        const int buf = i % nbuf;
        ATRACE(true, 3, i, 0);
        mbar_wait(&bars[PB_EMPTY + buf], (uint32_t)(((i / nbuf) & 1) ^ 1));
        ATRACE(true, 3, i, 1);
        uint8_t* sQ = smem + buf * item_bytes;
        uint8_t* sK = sQ + q_bytes;
        uint8_t* sV = sK + mat_bytes;
        uint64_t* full = &bars[PB_FULL + buf];
        mbar_arrive_expect_tx(full, 2 * pp_row_bytes(kv, seq) + pp_row_bytes(need, seq));
        pp_load_rows(&tm, &tm16, &tmF, full, sK, C + h * D, row0, kv, seq);
        pp_load_rows(&tm, &tm16, &tmF, full, sQ, h * D, row0, need, seq);
        pp_load_rows(&tm, &tm16, &tmF, full, sV, 2 * C + h * D, row0, kv, seq);
        ATRACE(true, 3, i, 2);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ------------------------------------------------------------------ MMA issuer: S(u), then P V of unit u - 1.
      // A slot's next S can only be issued behind the P V that consumes its P, so the softmax warps of the slot idle from
      // their P arrival until that S lands: everything the next S needs (item metadata, the item's FULL barrier, operand
      // addresses) is therefore prepared BEFORE this thread blocks on P, and nothing but the two MMA chains sits between
      // the P arrival and the S commit.
      struct Unit { int s; uint32_t k; const uint8_t* q; const uint8_t* sK; const uint8_t* sV; int spad, last_buf, buf; uint32_t full_par; bool full_ok; };
      int i = 0, t = 0, u = 0, Ti = 0;
      int4 m = make_int4(0, 0, 0, 0);
      auto load_item = [&]() {
        m = i < PP_MAX_ITEMS ? items[i] : item_meta(i);
        Ti = (m.z + 127) >> 7;
      };
      auto prepare = [&](Unit& q) {                           // unit (i, t); advances to the next one
        const int buf = i % nbuf;
        const uint8_t* sQ = smem + buf * item_bytes;
        q.s = u & 1;
        q.k = (uint32_t)(u >> 1);                            // how many units this slot has seen before
        q.q = sQ + t * 128 * 128;
        q.sK = sQ + q_bytes;
        q.sV = q.sK + mat_bytes;
        q.spad = (m.w + 15) & ~15;
        q.last_buf = t == Ti - 1 ? buf : -1;
        q.buf = buf;
        q.full_par = (uint32_t)((i / nbuf) & 1);
        q.full_ok = t != 0 || mbar_test_wait(&bars[PB_FULL + buf], q.full_par);
        ++u;
        if (++t == Ti) {
          t = 0;
          if (++i < my_items) load_item();
        }
      };
      auto do_pv = [&](const Unit& un) {
        mbar_wait(&bars[PB_P + un.s], un.k & 1);
        // O(u - 2) sits in the O columns of this slot until the epilogue warps have read it (already awaited before
        // S(u) if that aliased the O columns)
        if (un.spad <= (int)PP_O_OFF && un.k > 0) mbar_wait(&bars[PB_OFREE + un.s], (un.k - 1) & 1);
        tcgen05_fence_after();
        const uint32_t slot_addr = tmem_base + (uint32_t)(un.s * 256);
        issue_pv(slot_addr + PP_O_OFF, slot_addr, un.sV, un.spad);
        tcgen05_commit(&bars[PB_O + un.s]);
        if (un.last_buf >= 0) tcgen05_commit(&bars[PB_EMPTY + un.last_buf]);   // all MMAs reading this item are issued
      };
      Unit cur, prev;
      bool have_prev = false, more = my_items > 0;
      if (more) {
        load_item();
        prepare(cur);
      }
      while (more) {
        ATRACE(true, 2, (int)(2 * cur.k) + cur.s, 0);
        if (!cur.full_ok) mbar_wait(&bars[PB_FULL + cur.buf], cur.full_par);
        // more than 192 key columns: S aliases the slot's O columns, O(u - 2) must have been read out
        if (cur.spad > (int)PP_O_OFF && cur.k > 0) mbar_wait(&bars[PB_OFREE + cur.s], (cur.k - 1) & 1);
        tcgen05_fence_after();
        ATRACE(true, 2, (int)(2 * cur.k) + cur.s, 1);
        issue_qk(tmem_base + (uint32_t)(cur.s * 256), cur.q, cur.sK, cur.spad);
        tcgen05_commit(&bars[PB_S + cur.s]);
        ATRACE(true, 2, (int)(2 * cur.k) + cur.s, 2);
        Unit nxt = cur;
        more = i < my_items;
        if (more) prepare(nxt);
        if (have_prev) do_pv(prev);
        ATRACE(true, 2, (int)(2 * cur.k) + cur.s, 3);
        prev = cur;
        have_prev = true;
        cur = nxt;
      }
      if (have_prev) do_pv(prev);
    }
  } else if (warp < 10) {
    // -------------------------------------------------------------------- softmax warps: slot = (warp - 2) / 4, quarter = warp % 4
    // thread = row; a slot's warps take every other unit.  No epilogue here: the row sum and the pad-key factor go to the
    // epilogue warp of the quarter through shared memory (ordered by its wait for O = P V, which follows every P arrival).
    const int quarter = warp & 3;
    const int slot = (warp - 2) >> 2;
    float* xsum = xch + PP_XSUM + quarter * 32;                      // + (unit & 3) * 128: [lane]
    float* xcorr = xch + PP_XCORR + quarter * 32;                    // + (unit & 3) * 128: [lane]
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 256);
    const bool tr = quarter == 0 && lane == 0;
    (void)tr;
    int u = 0;
    for (int i = 0; i < my_items; ++i) {
      const int4 m = i < PP_MAX_ITEMS ? items[i] : item_meta(i);
      const int need = m.z, kv = m.w;
      const int Ti = (need + 127) >> 7;
      const float npad = (float)(seq - kv);
      const RowChunks rc = row_chunks(kv);
      for (int t = 0; t < Ti; ++t, ++u) {
        if ((u & 1) != slot) continue;
        const bool active = t * 128 + quarter * 32 < need;            // warps whose 32 query rows are all unneeded skip the tile
        ATRACE(tr, slot, u >> 1, 0);
        mbar_wait(&bars[PB_S + slot], (uint32_t)((u >> 1) & 1));
        tcgen05_fence_after();
        ATRACE(tr, slot, u >> 1, 1);
        if (active) {
          float mx = row_max(lane_base, kv, rc, tr && (u >> 1) == 1 ? slot : -1);
          if (npad > 0.f) mx = fmaxf(mx, 0.f);              // analytic pad keys: score 0
          const float mneg = -mx * LOG2E;
          ATRACE(tr, slot, u >> 1, 2);
          const float part = row_exp(lane_base, kv, rc, mneg, tr && (u >> 1) == 1 ? slot : -1);
          const float corr = npad > 0.f ? npad * ex2_approx(mneg) : 0.f;
          xsum[(u & 3) * 128 + lane] = part + corr;
          xcorr[(u & 3) * 128 + lane] = corr;
        }
        ATRACE(tr, slot, u >> 1, 3);
        tcgen05_fence_before();
        mbar_arrive(&bars[PB_P + slot]);
        ATRACE(tr, slot, u >> 1, 4);
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps: quarter = warp % 4, thread = row.
    // O = P V of unit u completes while the softmax warps are already in unit u + 1; read the 64 fp32 columns of this
    // thread's row, release them (OFREE), and store (O + corr pad_v) / sum as bf16 through a swizzled staging tile, so that
    // every store instruction writes 4 complete 128-byte row segments.
    const int quarter = warp & 3;
    const float* xsum = xch + PP_XSUM + quarter * 32;
    const float* xcorr = xch + PP_XCORR + quarter * 32;
    uint8_t* stage = stage_all + (warp - 10) * (32 * 128);
    const uint32_t quarter_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool tr = quarter == 0 && lane == 0;
    (void)tr;
    int u = 0;
    for (int i = 0; i < my_items; ++i) {
      const int4 m = i < PP_MAX_ITEMS ? items[i] : item_meta(i);
      const int w = m.x, h = m.y, need = m.z, kv = m.w;
      const int Ti = (need + 127) >> 7;
      const uint32_t o_off = PP_O_OFF;
      const float4* pv = pad_v != nullptr && kv < seq ? reinterpret_cast<const float4*>(pad_v + h * D) : nullptr;
      for (int t = 0; t < Ti; ++t, ++u) {
        const int s = u & 1;
        const int q = t * 128 + quarter * 32 + lane;
        int dst = q < need ? w * seq + q : -1;
        if (dst >= 0 && out_map != nullptr) dst = out_map[dst];
        const bool active = t * 128 + quarter * 32 < need;
        ATRACE_FIN(tr, 0, u, 0);
        mbar_wait(&bars[PB_O + s], (uint32_t)((u >> 1) & 1));
        tcgen05_fence_after();
        ATRACE_FIN(tr, 0, u, 1);
        uint32_t o0[32], o1[32];
        float sum = 1.f, corr = 0.f;
        if (active) {
          tmem_ld_32x32(quarter_base + (uint32_t)(s * 256) + o_off, o0);
          tmem_ld_32x32(quarter_base + (uint32_t)(s * 256) + o_off + 32u, o1);
          corr = xcorr[(u & 3) * 128 + lane];
          sum = xsum[(u & 3) * 128 + lane];
          tmem_ld_wait();
        }
        ATRACE_FIN(tr, 0, u, 2);
        tcgen05_fence_before();
        mbar_arrive(&bars[PB_OFREE + s]);
        ATRACE_FIN(tr, 0, u, 3);
        if (active) {
          const float inv = 1.0f / sum;
          if (pv != nullptr && corr != 0.f) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 a = __ldg(pv + j), b = __ldg(pv + 8 + j);
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
          uint8_t* srow = stage + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o0[8 * j + 0]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
            *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = v;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o1[8 * j + 0]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
            *reinterpret_cast<uint4*>(srow + (((4 + j) ^ (lane & 7)) << 4)) = v;
          }
          __syncwarp();
          ATRACE_FIN(tr, 0, u, 4);
          __nv_bfloat16* obase = out + h * D;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = (lane >> 3) + 4 * k, c = lane & 7;
            const int d = __shfl_sync(0xffffffffu, dst, r);
            const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 128 + ((c ^ (r & 7)) << 4));
            st_global_if(reinterpret_cast<uint4*>(obase + (size_t)d * C + c * 8), v, d >= 0);     // branch-free loop body
          }
          __syncwarp();
          ATRACE_FIN(tr, 0, u, 5);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
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
extern "C" int toc3d_attn_trace_fin(unsigned long long* host) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, g_attn_fin, sizeof(unsigned long long) * 512);
}
extern "C" int toc3d_attn_trace_chunks(unsigned long long* host) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, g_attn_chunk, sizeof(unsigned long long) * 64);
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
      TOC3D_CHECK_CUDA(cudaFuncSetAttribute(attn_tc::window_attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            227 * 1024));
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
      const int n_items = n_windows * heads;
      const int grid = n_items < n_sm ? n_items : n_sm;
      // ragged tails arrive through a second map with 16-row boxes, all rows of an item's Q / K / V through a third in one box
      CUtensorMap tm16, tmF;
      rc = make_tmap_bf16_2d(&tm16, qkv, (int64_t)n_windows * seq_len, 3 * (int64_t)C, 3 * (int64_t)C, attn_tc::PP_TAIL_ROWS);
      if (rc) return rc;
      rc = make_tmap_bf16_2d(&tmF, qkv, (int64_t)n_windows * seq_len, 3 * (int64_t)C, 3 * (int64_t)C, (seq_len + 15) & ~15);
      if (rc) return rc;
      const int ctl_bytes = attn_tc::PP_CTL_HEAD + attn_tc::PP_XCH_BYTES + attn_tc::PP_STAGE_BYTES;
      const int item_bytes = T * 128 * 128 + 2 * ((seq_len + 15) & ~15) * 128;     // Q in whole tiles, K and V exact
      int nbuf = (227 * 1024 - 1024 - ctl_bytes) / item_bytes;
      nbuf = nbuf > 4 ? 4 : nbuf;                       // >= 2 for seq <= 256 (96 KB per item)
      const size_t smem = (size_t)nbuf * item_bytes + 1024 + ctl_bytes;
      TOC3D_CHECK_CUDA(launch_pdl(attn_tc::window_attention_pp_kernel, dim3(grid), dim3(attn_tc::PP_THREADS), smem, st, 1,
                                  tm, tm16, tmF, o, seq_len, heads, n_items, nbuf, out_map, q_rows, item_order, kv_rows, pad_v));
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
