// Windowed multi-head attention, head dim 64, sequence = one window (<= 401 tokens in the shipped
// configs).  Flash-style: one CTA = 64 query rows of one (window, head); K/V tiles of 64 keys are
// double-buffered with cp.async; S = QK^T and O += PV run on mma.sync m16n8k16 (bf16 in, fp32
// accumulate); softmax is online in fp32.  q arrives rotated and pre-scaled (QKV epilogue).
//
// Replaces eva_vit.py:109-111 / toc3d_eva_vit.py:509-511 (q@k^T, softmax, @v, head merge).
#include "common.cuh"
#include "../../include/toc3d_b200.h"

namespace toc3d {
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
window_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int seq, int heads) {
  __shared__ __align__(16) Smem sm;
  pdl_wait();
  pdl_launch_dependents();
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
  __nv_bfloat16* obase = out + row0 * C + h * D;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = dt * 8 + 2 * t;
    if (r0 < seq) *reinterpret_cast<uint32_t*>(obase + (int64_t)r0 * C + col) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
    if (r1 < seq) *reinterpret_cast<uint32_t*>(obase + (int64_t)r1 * C + col) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

}  // namespace attn
}  // namespace toc3d

extern "C" int toc3d_window_attention(const void* qkv, void* out, int32_t n_windows, int32_t seq_len, int32_t heads,
                                      void* stream) {
  using namespace toc3d;
  TOC3D_REQUIRE(qkv && out, kErrBadArg, "toc3d_window_attention: null pointer");
  TOC3D_REQUIRE(n_windows > 0 && seq_len > 0 && seq_len <= 1024 && heads > 0 && heads <= 65535, kErrBadArg,
                "toc3d_window_attention: bad shape nW=%d seq=%d heads=%d", n_windows, seq_len, heads);
  TOC3D_REQUIRE(n_windows <= 65535, kErrBadArg, "toc3d_window_attention: too many windows (%d)", n_windows);
  dim3 grid((seq_len + attn::BQ - 1) / attn::BQ, heads, n_windows);
  TOC3D_CHECK_CUDA(launch_pdl(attn::window_attention_kernel, grid, dim3(attn::NTHREADS), 0,
                              reinterpret_cast<cudaStream_t>(stream), 1, reinterpret_cast<const __nv_bfloat16*>(qkv),
                              reinterpret_cast<__nv_bfloat16*>(out), seq_len, heads));
  return 0;
}
