// Persistent warp-specialised bf16 GEMM for sm_100a on CTA PAIRS: TMA -> 128B-swizzled smem ->
// tcgen05.mma.cta_group::2 (256 x BN x 16 per pair, fp32 accumulators double-buffered in TMEM) ->
// fused epilogues read with tcgen05.ld.
//
//   cluster = 2 CTAs (one TPC).  CTA r owns rows [128 r, 128 r + 128) of the 256-row pair tile and
//   stages its own A rows plus HALF of the B tile (BN/2 weight rows): 32 KB per k-block per CTA
//   instead of 48 KB for a single-CTA 128x256 tile, so the same 192 KB of smem holds 6 stages and the
//   L2 -> SM traffic per FLOP drops by a third.
//   warp 0 : TMA producer (one elected lane, both CTAs; completion bytes land on the leader's barrier)
//   warp 1 : TMEM allocator (both CTAs) + tcgen05.mma issuer (leader CTA only, one elected lane)
//   warps 2-9 : epilogue; warp w reads TMEM lane quarter (w % 4) and every other 32-column chunk
//   BN (tile width) is a runtime multiple of 32 in [64, 256], picked per launch to minimise wave
//   quantisation on the 74 CTA pairs.
//
// Call sites replaced: see include/toc3d_b200.h (toc3d_gemm_bf16).
#include "common.cuh"
#include "../../include/toc3d_b200.h"

#include <mutex>

namespace toc3d {
namespace gemm {

constexpr int BM = 128;                  // rows per CTA; the pair tile has 2 * BM rows
constexpr int BN_MAX = 256, BK = 64, UMMA_K = 16;
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
constexpr int B_BYTES_MAX = (BN_MAX / 2) * BK * 2;   // 16 KB: this CTA's half of the B tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int NUM_THREADS = 64 + 8 * 32;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int TMEM_COLS = 512;                    // 2 accumulator buffers x 256 fp32 columns
constexpr int ROPE_MAX_FT = 256;
constexpr int SMEM_TILES = STAGES * STAGE_BYTES;  // 196608
constexpr int SMEM_AUX = 256 + 8 * 32 * 32 * 4;   // barriers + epilogue staging (8 warps x 32 rows x 32 fp32)
constexpr int SMEM_BYTES = SMEM_TILES + SMEM_AUX + 1024;  // + alignment slack (230656 <= 232448)

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6), A=bf16 [7,10),
// B=bf16 [10,13), A/B K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).  M = 256 over the pair.
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}

struct EpiParams {
  const float* bias;
  void* out;
  int ldo;
  int out_f32;
  int act;
  const float* resid;
  const int* resid_map;
  int resid_mod;
  const int* out_map;
  float* out_alt;
  const int* rope_rows;
  int rope_slots;
  int rope_ft;
  int rope_cols;
  float q_scale;
  const float* cos_axis;
  const float* sin_axis;
  long long* row_stats;          // OUTPUT statistics accumulator (SWIGLU: per-row sum / sum of squares of the hidden rows)
  const long long* ln_stats;     // INPUT statistics of the A rows for a folded LayerNorm (RESID)
  const float* ln_u;
  int ln_n;
  float ln_eps;
  // implicit 3 x 3 convolution (A operand addressing only): k-block kb reads A columns (kb % conv_kmod) * BK of the rows
  // m + conv_shift[kb / conv_kmod]; conv_kmod = 0: plain GEMM
  int conv_kmod;
  int conv_shift[9];
};

// Folded-LayerNorm row statistics are accumulated as int64 fixed point: sum * 2^30 (|sum| < 8.6e9,
// step 9e-10) and sum of squares * 2^26 (< 1.4e11, step 1.5e-8; the LN eps is 1e-6 * n).
constexpr double STAT_SUM_SCALE = 1073741824.0;
constexpr double STAT_SQ_SCALE = 67108864.0;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// Diagnostic build only (-DTOC3D_GEMM_TRACE, tools/probes/gemm_trace.py): %globaltimer stamps (ns) of pair 0 / CTA 0 of the
// last 8 launches, [launch & 7][8 stamps]: 0 kernel entry, 1 prologue done, 2 dependency resolved (producer), 3 first
// operands landed (MMA thread), 4 last MMA issued, 5 last accumulator complete (epilogue warp 2), 6 last epilogue done,
// 7 before exit.  Compiled out of the product library.
#ifdef TOC3D_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[8 * 8];
__device__ unsigned int g_gemm_launch;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// [launch & 7][16]: inside the last epilogue of warp 2: 0 entry, 1 maps / residual prefetch issued, 2 accumulator seen,
// 3 + c: chunk c done
__device__ unsigned long long g_epi_trace[8 * 16];
#define GTRACE(cond, k) do { if ((cond) && blockIdx.x == 0) g_gemm_trace[(trace_slot & 7) * 8 + (k)] = gtime(); } while (0)
#define ETRACE(tr, k) do { if ((tr) >= 0 && lane == 0 && (k) < 16) g_epi_trace[((tr) & 7) * 16 + (k)] = gtime(); } while (0)
#else
#define GTRACE(cond, k) do {} while (0)
#define ETRACE(tr, k) do {} while (0)
#endif

// ---------------------------------------------------------------------------------------------
// Epilogue: 8 warps; warp w owns TMEM lane quarter (w % 4) = 32 rows of this CTA's 128 x BN
// accumulator and the 32-column chunks ch = half, half + 2, ... (half = (w - 2) / 4).  Each chunk
// goes TMEM -> registers (thread = row) -> XOR-swizzled smem staging (4 KB per warp, conflict-free
// both ways) -> "coalesced domain": 8 lanes x float4 cover the 32 columns of one row, 4 rows per
// instruction, 8 independent iterations per chunk.  Bias / RoPE / residual / activation are applied
// there, so global memory sees contiguous 128-byte (fp32) or 64-byte (bf16) row segments.
// Everything that does not depend on the accumulator (row maps, folded-LN coefficients, the first
// chunk's residual) is fetched BEFORE waiting for the MMA warp, and the TMEM / residual loads of
// the next chunk are issued before the current chunk is processed.
constexpr int CHUNK = 32;          // columns per staged chunk
constexpr int EPI_WARPS = 8;
constexpr int STAGE_FLOATS = 32 * CHUNK;   // per warp

__device__ __forceinline__ void stage_rows32(float* stage, int lane, const float (&f)[32]) {
  float4* d = reinterpret_cast<float4*>(stage) + lane * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j ^ (lane & 7)] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
}
__device__ __forceinline__ void stage_write(float* stage, int row, int cseg, float4 v) {
  reinterpret_cast<float4*>(stage)[row * 8 + (cseg ^ (row & 7))] = v;
}
__device__ __forceinline__ float4 stage_read(const float* stage, int row, int cseg) {
  return reinterpret_cast<const float4*>(stage)[row * 8 + (cseg ^ (row & 7))];
}
__device__ __forceinline__ void tmem_ld_f32x32(uint32_t taddr, float (&f)[32]) {
  uint32_t v[32];
  tmem_ld_32x32(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
}

// taddr: accumulator base of this warp's lane quarter (column 0 of the tile); m0: first global row
// of the quarter; nt0: first global column of the tile; bn: tile width.  full_bar/parity: the
// "accumulator complete" barrier this warp must observe before its first TMEM load.
// per-row LayerNorm fold coefficients from the fixed-point statistics: y = a * acc + b * u[col] + bias[col]
__device__ __forceinline__ void ln_fold_coeffs(const long long* stats_row, int n, float eps, float& a, float& b) {
  const longlong2 st = *reinterpret_cast<const longlong2*>(stats_row);
  const double inv_n = 1.0 / (double)n;
  const double mean = (double)st.x * (1.0 / STAT_SUM_SCALE) * inv_n;
  const double var = fmax((double)st.y * (1.0 / STAT_SQ_SCALE) * inv_n - mean * mean, 0.0);
  a = rsqrtf((float)var + eps);
  b = -a * (float)mean;
}

template <int EPI, bool LNF>
__device__ __forceinline__ void epilogue_warp_tile(const EpiParams& ep, uint32_t taddr, int m0, int nt0, int bn, int half,
                                                   int M, int N, float* stage, int lane, uint64_t* full_bar,
                                                   uint32_t parity, int tr = -1) {
  ETRACE(tr, 0);
  // coalesced-domain coordinates: rows {rin, rin+4, ..., rin+28}, columns 4*cseg..4*cseg+3 of the chunk
  const int rin = lane >> 3;
  const int cseg = lane & 7;
  const bool my_row_ok = (m0 + lane) < M;

  if constexpr (EPI == TOC3D_EPI_SWIGLU) {
    // 64 GEMM columns = [32 x w1 | 32 x w2] -> 32 hidden columns; this warp takes every other 64-block
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out);
    // this row's sum / sum of squares of the bf16-rounded hidden values, accumulated per 32-column block
    // in fixed point so that the result does not depend on the tile width or on which warp owns a block
    long long st_sum = 0, st_sq = 0;
    mbar_wait(full_bar, parity);
    tcgen05_fence_after();
#pragma unroll 1
    for (int blk = half; blk * 64 < bn; blk += 2) {
      const int col1 = nt0 + blk * 64;                   // GEMM column of the w1 part
      if (col1 >= N) break;                              // warp-uniform
      float h[32];
      {
        uint32_t v1[32], v2[32];
        tmem_ld_32x32(taddr + blk * 64, v1);
        tmem_ld_32x32(taddr + blk * 64 + 32, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f), b2 = b1;
          if (ep.bias != nullptr) {
            b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col1) + j);
            b2 = __ldg(reinterpret_cast<const float4*>(ep.bias + col1 + 32) + j);
          }
          h[4 * j + 0] = silu(__uint_as_float(v1[4 * j + 0]) + b1.x) * (__uint_as_float(v2[4 * j + 0]) + b2.x);
          h[4 * j + 1] = silu(__uint_as_float(v1[4 * j + 1]) + b1.y) * (__uint_as_float(v2[4 * j + 1]) + b2.y);
          h[4 * j + 2] = silu(__uint_as_float(v1[4 * j + 2]) + b1.z) * (__uint_as_float(v2[4 * j + 2]) + b2.z);
          h[4 * j + 3] = silu(__uint_as_float(v1[4 * j + 3]) + b1.w) * (__uint_as_float(v2[4 * j + 3]) + b2.w);
        }
      }
      if (ep.row_stats != nullptr) {
        float bs = 0.f, bq = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          h[i] = __bfloat162float(__float2bfloat16_rn(h[i]));   // statistics of exactly what the next GEMM reads
          bs += h[i];
          bq += h[i] * h[i];
        }
        st_sum += __float2ll_rn(bs * (float)STAT_SUM_SCALE);
        st_sq += __float2ll_rn(bq * (float)STAT_SQ_SCALE);
      }
      stage_rows32(stage, lane, h);
      __syncwarp();
      const int hcol = (col1 >> 1) + 4 * cseg;            // hidden column of this lane
      const bool hcol_ok = hcol < ep.ldo;
#pragma unroll
      for (int it = 0; it < 8; ++it) {                    // branch-free body (st_global_if): the iterations overlap
        const int row = it * 4 + rin;
        const float4 a = stage_read(stage, row, cseg);
        uint2 u;
        u.x = pack_bf16(a.x, a.y);
        u.y = pack_bf16(a.z, a.w);
        st_global_if(reinterpret_cast<uint2*>(out + (size_t)(m0 + row) * ep.ldo + hcol), u, hcol_ok && m0 + row < M);
      }
      __syncwarp();
    }
    if (ep.row_stats != nullptr && my_row_ok) {
      // fixed-point (integer) atomics: the accumulated statistics do not depend on arrival order
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(ep.row_stats + 2 * (size_t)(m0 + lane));
      atomicAdd(dst, (unsigned long long)st_sum);
      atomicAdd(dst + 1, (unsigned long long)st_sq);
    }
    return;
  }

  const int ch0 = half * CHUNK;       // this warp's first chunk (tile-relative column)
  if constexpr (EPI == TOC3D_EPI_RESID) {
    // per-row maps and folded-LN coefficients, computed by the lane that owns the row, then
    // redistributed to the coalesced-domain owners (8 rows per lane)
    int rr_t = -1, or_t = -1;
    float lnA_t = 1.0f, lnB_t = 0.0f;    // y = lnA * acc + lnB * u[col] + bias[col]
    if (my_row_ok) {
      const int row = m0 + lane;
      rr_t = ep.resid_mod > 0 ? (row % ep.resid_mod) : (ep.resid_map ? ep.resid_map[row] : row);
      or_t = ep.out_map ? ep.out_map[row] : row;
      if constexpr (LNF) ln_fold_coeffs(ep.ln_stats + 2 * (size_t)row, ep.ln_n, ep.ln_eps, lnA_t, lnB_t);
    }
    // 32-bit row indices + masks instead of 16 pointers (register pressure)
    int o_row[8], r_row[8];
    uint32_t o_ok = 0, o_alt = 0, r_ok = 0, r_alt = 0;
    float la[8], lb[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + rin;
      const int orow = __shfl_sync(0xffffffffu, or_t, row);
      const int rrow = __shfl_sync(0xffffffffu, rr_t, row);
      if constexpr (LNF) {
        la[it] = __shfl_sync(0xffffffffu, lnA_t, row);
        lb[it] = __shfl_sync(0xffffffffu, lnB_t, row);
      } else {
        la[it] = 1.0f;
        lb[it] = 0.0f;
      }
      o_row[it] = orow >= 0 ? orow : m0 + row;
      r_row[it] = rrow >= 0 ? rrow : m0 + row;
      if (orow != -1) {
        o_ok |= 1u << it;
        if (orow == -2) o_alt |= 1u << it;
        if (rrow != -1) r_ok |= 1u << it;
        if (rrow == -2) r_alt |= 1u << it;
      }
    }
    auto load_resid = [&](float4 (&r)[8], int col) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < N && ((r_ok >> it) & 1u)) {
          const float* base = ((r_alt >> it) & 1u) ? ep.out_alt : ep.resid;
          r[it] = __ldcg(reinterpret_cast<const float4*>(base + (size_t)r_row[it] * ep.ldo + col));   // one-shot: skip L1
        }
      }
    };
    auto load_cols = [&](float4& b, float4& u4, int col) {      // per-column bias and folded-LN vector
      b = make_float4(0.f, 0.f, 0.f, 0.f);
      u4 = b;
      if (col < N) {
        if (ep.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
        if constexpr (LNF) u4 = __ldg(reinterpret_cast<const float4*>(ep.ln_u + col));
      }
    };
    float4 r_cur[8], r_nxt[8];
    float4 b, u4, b_nxt, u_nxt;
    if (ch0 < bn) {                                               // in flight while the MMAs finish
      load_resid(r_cur, nt0 + ch0 + 4 * cseg);
      load_cols(b, u4, nt0 + ch0 + 4 * cseg);
    }
    ETRACE(tr, 1);
    mbar_wait(full_bar, parity);
    tcgen05_fence_after();
    ETRACE(tr, 2);
    if (ch0 >= bn || nt0 + ch0 >= N) return;                      // warp-uniform
    float f[32];
    tmem_ld_f32x32(taddr + ch0, f);
#pragma unroll 1
    for (int c = ch0; c < bn; c += 2 * CHUNK) {
      const int col0 = nt0 + c;
      if (col0 >= N) break;                               // warp-uniform
      stage_rows32(stage, lane, f);
      __syncwarp();
      const bool more = c + 2 * CHUNK < bn && col0 + 2 * CHUNK < N;
      if (more) {
        load_resid(r_nxt, col0 + 2 * CHUNK + 4 * cseg);
        load_cols(b_nxt, u_nxt, col0 + 2 * CHUNK + 4 * cseg);
        tmem_ld_f32x32(taddr + c + 2 * CHUNK, f);
      }
      const int col = col0 + 4 * cseg;
      const bool col_ok = col < N;                        // N % 4 == 0
#pragma unroll
      for (int it = 0; it < 8; ++it) {                    // branch-free body (st_global_cg_if): the iterations overlap
        const float4 a = stage_read(stage, it * 4 + rin, cseg);
        float4 o;
        o.x = r_cur[it].x + (fmaf(la[it], a.x, lb[it] * u4.x) + b.x);
        o.y = r_cur[it].y + (fmaf(la[it], a.y, lb[it] * u4.y) + b.y);
        o.z = r_cur[it].z + (fmaf(la[it], a.z, lb[it] * u4.z) + b.z);
        o.w = r_cur[it].w + (fmaf(la[it], a.w, lb[it] * u4.w) + b.w);
        float* base = ((o_alt >> it) & 1u) ? ep.out_alt : reinterpret_cast<float*>(ep.out);
        st_global_cg_if(reinterpret_cast<float4*>(base + (size_t)o_row[it] * ep.ldo + col), o, col_ok && ((o_ok >> it) & 1u));
      }
      __syncwarp();
      ETRACE(tr, 3 + (c - ch0) / (2 * CHUNK));
      if (more) {
#pragma unroll
        for (int it = 0; it < 8; ++it) r_cur[it] = r_nxt[it];
        b = b_nxt;
        u4 = u_nxt;
      }
    }
    return;
  }

  // ---- QKV_ROPE and LINEAR
  // Everything below the accumulator is fetched before the wait.  A warp's chunks are 64 columns apart
  // (half = chunk parity), i.e. always the same half of a 64-wide head: the same RoPE axis and the same
  // 16 frequencies for all of them, so its cos/sin values are loaded once per tile (8 rows x 2 pairs).
  float2 rc[8], rs[8];
  if constexpr (EPI == TOC3D_EPI_QKV_ROPE) {
    int pos_t = 0;
    if (my_row_ok) {
      const int row = m0 + lane;
      const int t = ep.rope_rows ? ep.rope_rows[row] : (row % ep.rope_slots);
      const int r = t / ep.rope_ft;
      pos_t = (r << 16) | (t - r * ep.rope_ft);
    }
    const bool col_axis = ((nt0 + ch0) >> 5) & 1;         // second 32 channels of a head use the column coordinate
    const int j0 = 2 * cseg;                              // frequency index of this lane's first pair
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int pp = __shfl_sync(0xffffffffu, pos_t, it * 4 + rin);
      const int p = col_axis ? (pp & 0xffff) : (pp >> 16);
      rc[it] = __ldg(reinterpret_cast<const float2*>(ep.cos_axis + p * 16 + j0));
      rs[it] = __ldg(reinterpret_cast<const float2*>(ep.sin_axis + p * 16 + j0));
    }
  }
  // destination rows (QKV_ROPE / LINEAR): identity or out_map (-1 = row dropped), per coalesced-domain row
  int d_row[8];
  {
    int dr_t = my_row_ok ? m0 + lane : -1;
    if (dr_t >= 0 && ep.out_map != nullptr) dr_t = ep.out_map[dr_t];
#pragma unroll
    for (int it = 0; it < 8; ++it) d_row[it] = __shfl_sync(0xffffffffu, dr_t, it * 4 + rin);
  }
  float4 bias4[BN_MAX / (2 * CHUNK)];
#pragma unroll
  for (int i = 0; i < BN_MAX / (2 * CHUNK); ++i) {
    const int col = nt0 + ch0 + i * 2 * CHUNK + 4 * cseg;
    bias4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr && ch0 + i * 2 * CHUNK < bn && col < N) bias4[i] = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
  }
  ETRACE(tr, 1);
  mbar_wait(full_bar, parity);
  tcgen05_fence_after();
  ETRACE(tr, 2);
  if (ch0 >= bn || nt0 + ch0 >= N) return;                        // warp-uniform
  float f[32];
  tmem_ld_f32x32(taddr + ch0, f);
#pragma unroll
  for (int i = 0; i < BN_MAX / (2 * CHUNK); ++i) {
    const int c = ch0 + i * 2 * CHUNK;
    const int col0 = nt0 + c;
    if (c >= bn || col0 >= N) break;                      // warp-uniform
    stage_rows32(stage, lane, f);
    __syncwarp();
    if (c + 2 * CHUNK < bn && col0 + 2 * CHUNK < N) tmem_ld_f32x32(taddr + c + 2 * CHUNK, f);
    const int col = col0 + 4 * cseg;
    const bool col_ok = col < N;                          // N % 4 == 0
    const float4 b = bias4[i];
    if constexpr (EPI == TOC3D_EPI_QKV_ROPE) {
      const bool rot = col0 < ep.rope_cols;               // warp-uniform (rope_cols % 128 == 0)
      const float sc_q = (col0 < (ep.rope_cols >> 1)) ? ep.q_scale : 1.0f;
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out);
      // the warp-uniform `rot` is decided once per chunk and the stores are predicated: branch-free loop bodies, so the
      // eight iterations overlap instead of running as eight convergence regions
      if (rot) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float4 a = stage_read(stage, it * 4 + rin, cseg);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
          const float2 c2 = rc[it], sn = rs[it];
          const float x0 = a.x, x1 = a.y, x2 = a.z, x3 = a.w;
          a.x = (x0 * c2.x - x1 * sn.x) * sc_q; a.y = (x1 * c2.x + x0 * sn.x) * sc_q;
          a.z = (x2 * c2.y - x3 * sn.y) * sc_q; a.w = (x3 * c2.y + x2 * sn.y) * sc_q;
          uint2 u;
          u.x = pack_bf16(a.x, a.y);
          u.y = pack_bf16(a.z, a.w);
          st_global_if(reinterpret_cast<uint2*>(out + (size_t)d_row[it] * ep.ldo + col), u, col_ok && d_row[it] >= 0);
        }
      } else {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float4 a = stage_read(stage, it * 4 + rin, cseg);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
          uint2 u;
          u.x = pack_bf16(a.x, a.y);
          u.y = pack_bf16(a.z, a.w);
          st_global_if(reinterpret_cast<uint2*>(out + (size_t)d_row[it] * ep.ldo + col), u, col_ok && d_row[it] >= 0);
        }
      }
    } else {  // TOC3D_EPI_LINEAR
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + rin;
        float4 a = stage_read(stage, row, cseg);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        if (ep.act == 1) { a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w); }
        else if (ep.act == 2) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        const bool ok = col_ok && d_row[it] >= 0;
        uint2 u;
        u.x = pack_bf16(a.x, a.y);
        u.y = pack_bf16(a.z, a.w);
        st_global_if(reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)d_row[it] * ep.ldo + col), a, ok && ep.out_f32);
        st_global_if(reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)d_row[it] * ep.ldo + col), u, ok && !ep.out_f32);
      }
    }
    __syncwarp();
    ETRACE(tr, 3 + i);
  }
}

// (registers are allocated per 4 warps: 10 warps count as 12, hence the 168-register cap)
template <int EPI, bool LNF>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
            int BN, const EpiParams ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle; pointer arithmetic (not an integer round trip) keeps the
  // shared address space visible to the compiler (LDS/STS instead of generic loads in the epilogue).
  // Both CTAs of the pair compute the same offset (same kernel, same static layout), which the
  // cta_group::2 MMA and the multicast commits rely on.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_TILES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stage = reinterpret_cast<float*>(smem + SMEM_TILES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef TOC3D_GEMM_TRACE
  __shared__ unsigned int s_trace_slot;
  if (threadIdx.x == 0 && blockIdx.x == 0) s_trace_slot = atomicAdd(&g_gemm_launch, 1u);
  const unsigned long long t_entry = gtime();
#endif
  const uint32_t rank = cluster_ctarank();          // rank in the pair (= cluster), 0 = leader
  const uint32_t lead = 0u;                         // cluster rank of the pair's leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int mrows = 2 * BM;                         // rows covered by one pair tile
  const int num_m = (M + mrows - 1) / mrows;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;
  const int b_rows = BN >> 1;                       // weight rows this CTA feeds to the pair MMA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * EPI_WARPS);     // the epilogue warps of BOTH CTAs release the leader's MMA
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr, TMEM_COLS);
  tcgen05_fence_before();
  cluster_sync_all();                               // barriers + TMEM of both CTAs are ready
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();            // the next kernel may start its prologue; it waits for this grid itself
#ifdef TOC3D_GEMM_TRACE
  const unsigned int trace_slot = blockIdx.x == 0 ? s_trace_slot : 0u;
  if (threadIdx.x == 0 && blockIdx.x == 0) g_gemm_trace[(trace_slot & 7) * 8 + 0] = t_entry;
  GTRACE(threadIdx.x == 0, 1);
#endif

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = 2u * (uint32_t)(A_BYTES + b_rows * BK * 2);    // bytes landing in both CTAs of the pair
      auto load_b = [&](int stage, int kb, int n_idx) {
        const uint32_t bar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;             // pair leader's barrier (peer bit cleared)
        uint8_t* sb = smem + stage * STAGE_BYTES + A_BYTES;
        tma_load_2d_2sm(&tmB, bar, sb, kb * BK, n_idx);
      };
      // The weights (B) do not depend on the previous kernel in the stream: the first ring of B loads
      // is issued before the programmatic-dependency wait and overlaps that kernel's tail.
      int pre = 0;
      if (pair < num_tiles) {
        const int n_idx = (pair / num_m) * BN + (int)rank * b_rows;
        pre = num_k < STAGES ? num_k : STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[kb], stage_tx);
          load_b(kb, kb, n_idx);
        }
      }
      pdl_wait();
      GTRACE(true, 2);
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_idx = (tile % num_m) * mrows + (int)rank * BM;
        const int n_idx = (tile / num_m) * BN + (int)rank * b_rows;
        for (int kb = 0; kb < num_k; ++kb) {
          const uint32_t lead_full = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
          uint8_t* sa = smem + stage * STAGE_BYTES;
          if (pre > 0) {                              // B of this slot is already in flight
            --pre;
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
            load_b(stage, kb, n_idx);
          }
          if (ep.conv_kmod > 0) {    // tap (ky, kx) of an implicit 3 x 3 convolution: the same rows, shifted (OOB rows read 0)
            const int tap = kb / ep.conv_kmod;
            tma_load_2d_2sm(&tmA, lead_full, sa, (kb - tap * ep.conv_kmod) * BK, m_idx + ep.conv_shift[tap]);
          } else {
            tma_load_2d_2sm(&tmA, lead_full, sa, kb * BK, m_idx);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc(BN);
      const uint16_t full_mask = (uint16_t)(3u << lead);
      const uint16_t empty_mask = full_mask;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN_MAX);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          GTRACE(tile == pair && kb == 0, 3);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_desc = umma_desc_k_sw128(sa);
          const uint64_t b_desc = umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the 128B swizzle row: +2 in the >>4 address field
            umma_bf16_ss_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                             (kb | k) != 0 ? 1u : 0u);
          }
          tcgen05_commit_2sm(&empty_bar[stage], empty_mask);   // slot reusable in every CTA that may refill it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_2sm(&tmem_full[acc], full_mask);      // accumulator complete -> both epilogues of the pair
        GTRACE(tile + num_pairs >= num_tiles, 4);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps per CTA)
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read (warp_id % 4)
    const int half = (warp - 2) >> 2;      // even / odd 32-column chunks
    float* stage_buf = s_stage + (warp - 2) * STAGE_FLOATS;
    int acc = 0;
    uint32_t acc_phase = 0;
    pdl_wait();                            // residual / maps / statistics come from the previous kernels
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_idx = (tile % num_m) * mrows + (int)rank * BM;
      const int n_idx = (tile / num_m) * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN_MAX);
      int tr = -1;
#ifdef TOC3D_GEMM_TRACE
      if (warp == 2 && tile + num_pairs >= num_tiles && blockIdx.x == 0) tr = (int)(trace_slot & 7);
#endif
      epilogue_warp_tile<EPI, LNF>(ep, taddr, m_idx + quarter * 32, n_idx, BN, half, M, N, stage_buf, lane,
                                   &tmem_full[acc], acc_phase, tr);
      GTRACE(warp == 2 && lane == 0 && tile + num_pairs >= num_tiles, 6);
      // release this accumulator buffer to the leader's MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), lead));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                      // nobody may still signal a barrier / read TMEM of an exited peer
  GTRACE(threadIdx.x == 0, 7);
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

#ifdef TOC3D_GEMM_TRACE
extern "C" int toc3d_gemm_trace_read(unsigned long long* host) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, g_gemm_trace, sizeof(unsigned long long) * 64);
}
extern "C" int toc3d_gemm_epi_trace_read(unsigned long long* host) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, g_epi_trace, sizeof(unsigned long long) * 128);
}
#endif


// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace gemm

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  using namespace gemm;
  EncodeTiledFn enc = get_encode_fn();
  TOC3D_REQUIRE(enc != nullptr, kErrNoDriver, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TOC3D_REQUIRE(r == CUDA_SUCCESS, kErrBadArg, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld ld=%lld",
                (int)r, (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

namespace gemm {

constexpr int kMaxDevices = 64;       // one-time kernel setup is kept per device ordinal (several GPUs in one process)
static std::mutex g_cfg_mutex;

static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

static int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

// Tile width.  The mainloop is L2 -> SM bandwidth bound (measured: 8192^3 runs at 1.49 PFLOP/s with 256-wide
// tiles and time per tile scales with the A + B bytes staged per k-block), so narrower tiles re-read A more
// often and only pay off when they remove a mostly empty last wave:
//   cost(BN) = waves * (k_blocks * (A + B bytes per CTA) + c_tile) + c_epi * BN      (last epilogue is exposed)
static int pick_tile_n(int M, int N, int K, int kind, int units) {
  const int mrows = 2 * BM;
  const int num_m = (M + mrows - 1) / mrows;
  const int step = kind == TOC3D_EPI_SWIGLU ? 64 : 32;
  const double kb = (double)((K + BK - 1) / BK);
  double best = 1e30;
  int best_bn = BN_MAX;
  for (int bn = BN_MAX; bn >= 128; bn -= step) {
    const long tiles = (long)num_m * ((N + bn - 1) / bn);
    const long waves = (tiles + units - 1) / units;
    const double cost = (double)waves * (kb * (256.0 + (double)bn) + 1024.0) + 16.0 * bn;
    if (cost < best * 0.97) { best = cost; best_bn = bn; }     // prefer the widest tile unless clearly better
  }
  return best_bn;
}

template <int EPI, bool LNF = false>
static int launch(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int tile_n,
                  const EpiParams& ep, cudaStream_t st) {
  static bool configured_dev[kMaxDevices] = {};
  static int max_pairs_dev[kMaxDevices] = {};     // co-resident CTA pairs (GPC boundaries can strand SMs)
  const int dev = current_device();
  {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    if (!configured_dev[dev]) {
      TOC3D_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<EPI, LNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3((unsigned)(sm_count() / 2 * 2)); cfg.blockDim = dim3(NUM_THREADS);
      cfg.dynamicSmemBytes = SMEM_BYTES; cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_kernel<EPI, LNF>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = sm_count() / 2;
      }
      max_pairs_dev[dev] = n < sm_count() / 2 ? n : sm_count() / 2;
      configured_dev[dev] = true;
    }
  }
  // Built, verified and measured slower on B200, then removed: two pairs per cluster sharing the weight
  // tile by TMA multicast (3 - 10 % slower: multicast across <= 4 CTAs does not reduce the L2 reads on this part, and
  // 4-CTA clusters strand SMs at GPC boundaries; profiles/r01l_gemm_bench_cluster_pairs.txt); wide tiles (one 352 / 512
  // column accumulator, two MMAs per k-step, a single wave for the N = 1024 GEMMs: 6 - 10 % slower than the two balanced
  // waves of 176 / 192-wide tiles the cost model picks, whose first epilogue hides behind the second mainloop;
  // profiles/r02e_gemm_bench_wide_tiles_experiment.txt); stream-K; chained launches.
  const int max_units = max_pairs_dev[dev];
  const int bn = tile_n > 0 ? tile_n : pick_tile_n(M, N, K, EPI, max_units);
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, A, M, ep.conv_kmod > 0 ? ep.conv_kmod * BK : K, lda, BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, B, N, K, ldb, bn / 2);
  if (rc) return rc;
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + bn - 1) / bn);
  const int units = tiles < max_units ? tiles : max_units;
  TOC3D_CHECK_CUDA(launch_pdl(gemm_kernel<EPI, LNF>, dim3(2 * units), dim3(NUM_THREADS), SMEM_BYTES, st, 2, ta, tb, M, N, K, bn, ep));
  return 0;
}

}  // namespace gemm
}  // namespace toc3d

namespace toc3d {
namespace gemm {

// C-ABI epilogue -> kernel parameters, with the argument checks of toc3d_gemm_bf16
static int to_params(const char* fn, const toc3d_epilogue* e, int kind, int N, EpiParams& ep) {
  ep.bias = e->bias; ep.out = e->out; ep.ldo = e->ldo; ep.out_f32 = e->out_f32; ep.act = e->act;
  ep.resid = e->resid; ep.resid_map = e->resid_map; ep.resid_mod = e->resid_mod; ep.out_map = e->out_map;
  ep.out_alt = e->out_alt; ep.rope_rows = e->rope_rows; ep.rope_slots = e->rope_slots; ep.rope_ft = e->rope_ft;
  ep.rope_cols = e->rope_cols; ep.q_scale = e->q_scale; ep.cos_axis = e->cos_axis; ep.sin_axis = e->sin_axis;
  ep.row_stats = reinterpret_cast<long long*>(e->row_stats); ep.ln_u = e->ln_u; ep.ln_n = e->ln_n; ep.ln_eps = e->ln_eps;
  ep.ln_stats = reinterpret_cast<const long long*>(e->ln_stats);
  ep.conv_kmod = 0;
  for (int i = 0; i < 9; ++i) ep.conv_shift[i] = 0;
  if (e->conv_cin > 0) {
    TOC3D_REQUIRE(kind == TOC3D_EPI_LINEAR && e->conv_cin % BK == 0, kErrBadArg,
                  "%s: implicit 3x3 convolution needs a LINEAR epilogue and conv_cin %% 64 == 0 (got %d)", fn, e->conv_cin);
    ep.conv_kmod = e->conv_cin / BK;
    for (int i = 0; i < 9; ++i) ep.conv_shift[i] = e->conv_row_shift[i];
  }
  const int tile_n = e->tile_n;
  TOC3D_REQUIRE(tile_n == 0 || (tile_n >= 64 && tile_n <= BN_MAX && tile_n % 32 == 0 &&
                                (kind != TOC3D_EPI_SWIGLU || tile_n % 64 == 0)), kErrBadArg,
                "%s: tile_n must be 0 (auto) or a multiple of 32 (64 for SWIGLU) in [64, 256], got %d", fn, tile_n);
  const bool lnf = ep.ln_stats != nullptr;
  if (lnf)
    TOC3D_REQUIRE(kind == TOC3D_EPI_RESID && ep.ln_u != nullptr && ep.ln_n > 0 &&
                  ((uintptr_t)ep.ln_u & 15) == 0 && ((uintptr_t)ep.ln_stats & 15) == 0,
                  kErrBadArg, "%s: folded LayerNorm (RESID) needs ln_stats + ln_u (16-byte aligned), ln_n > 0", fn);
  TOC3D_REQUIRE(ep.ldo > 0 && ep.ldo % 4 == 0 && N % 4 == 0, kErrBadArg,
                "%s: N and ldo must be positive multiples of 4 (vector epilogue), got N=%d ldo=%d", fn, N, ep.ldo);
  TOC3D_REQUIRE(((uintptr_t)ep.out & 15) == 0 && ((uintptr_t)ep.bias & 15) == 0 && ((uintptr_t)ep.resid & 15) == 0 &&
                ((uintptr_t)ep.out_alt & 15) == 0, kErrBadArg, "%s: epilogue pointers must be 16-byte aligned", fn);
  if (kind == TOC3D_EPI_QKV_ROPE) {
    TOC3D_REQUIRE(ep.cos_axis && ep.sin_axis && ep.rope_ft > 0 && ep.rope_ft <= ROPE_MAX_FT, kErrBadArg,
                  "%s: bad RoPE tables (ft=%d)", fn, ep.rope_ft);
    TOC3D_REQUIRE(ep.rope_rows || ep.rope_slots > 0, kErrBadArg, "%s: rope_rows or rope_slots required", fn);
    TOC3D_REQUIRE(ep.rope_cols % 128 == 0 && ep.rope_cols <= N, kErrBadArg, "%s: rope_cols %d", fn, ep.rope_cols);
  }
  if (kind == TOC3D_EPI_RESID) {
    TOC3D_REQUIRE(ep.resid_mod > 0 ? ep.resid != nullptr : true, kErrBadArg, "%s: resid_mod needs resid", fn);
    TOC3D_REQUIRE(ep.resid != nullptr || ep.resid_map != nullptr, kErrBadArg, "%s: RESID needs resid", fn);
  }
  if (kind == TOC3D_EPI_SWIGLU) TOC3D_REQUIRE(N % 64 == 0, kErrBadArg, "%s: SWIGLU needs N %% 64 == 0", fn);
  return 0;
}


}  // namespace gemm
}  // namespace toc3d

extern "C" int toc3d_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N,
                               int32_t K, int32_t kind, const toc3d_epilogue* e, void* stream) {
  using namespace toc3d;
  using namespace toc3d::gemm;
  TOC3D_REQUIRE(A && B && e && e->out, kErrBadArg, "toc3d_gemm_bf16: null pointer");
  TOC3D_REQUIRE(M > 0 && N > 0 && K > 0, kErrBadArg, "toc3d_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  TOC3D_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, kErrBadArg,
                "toc3d_gemm_bf16: K, lda, ldb must be multiples of 8 (16-byte TMA strides)");
  TOC3D_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, kErrBadArg, "toc3d_gemm_bf16: unaligned operand");
  TOC3D_REQUIRE(kind >= TOC3D_EPI_LINEAR && kind <= TOC3D_EPI_SWIGLU, kErrBadArg, "toc3d_gemm_bf16: unknown epilogue kind %d", kind);
  EpiParams ep;
  int rc = to_params("toc3d_gemm_bf16", e, kind, N, ep);
  if (rc) return rc;
  TOC3D_REQUIRE(ep.conv_kmod == 0 || K == 9 * e->conv_cin, kErrBadArg,
                "toc3d_gemm_bf16: implicit 3x3 convolution needs K == 9 * conv_cin (K=%d, conv_cin=%d)", K, e->conv_cin);
  const int tile_n = e->tile_n;
  const bool lnf = ep.ln_stats != nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (kind) {
    case TOC3D_EPI_LINEAR: return launch<TOC3D_EPI_LINEAR>(A, lda, B, ldb, M, N, K, tile_n, ep, st);
    case TOC3D_EPI_QKV_ROPE: return launch<TOC3D_EPI_QKV_ROPE>(A, lda, B, ldb, M, N, K, tile_n, ep, st);
    case TOC3D_EPI_RESID:
      return lnf ? launch<TOC3D_EPI_RESID, true>(A, lda, B, ldb, M, N, K, tile_n, ep, st)
                 : launch<TOC3D_EPI_RESID, false>(A, lda, B, ldb, M, N, K, tile_n, ep, st);
    default:
      return launch<TOC3D_EPI_SWIGLU>(A, lda, B, ldb, M, N, K, tile_n, ep, st);
  }
}
