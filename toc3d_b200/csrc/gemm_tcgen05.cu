// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem -> tcgen05.mma
// (128x256x16, fp32 accumulators double-buffered in TMEM) -> fused epilogues read with tcgen05.ld.
//
//   warp 0 : TMA producer (one elected lane)
//   warp 1 : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2-5 : epilogue, one TMEM lane quarter each (lane quarter = warp_id % 4)
//
// Call sites replaced: see include/toc3d_b200.h (toc3d_gemm_bf16).
#include "common.cuh"
#include "../../include/toc3d_b200.h"

#include <mutex>

namespace toc3d {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;                    // 2 accumulator buffers x 256 fp32 columns
constexpr int ROPE_MAX_FT = 32;
constexpr int SMEM_TILES = STAGES * STAGE_BYTES;  // 196608
constexpr int SMEM_AUX = 256 + 2 * ROPE_MAX_FT * 16 * 4;
constexpr int SMEM_BYTES = SMEM_TILES + SMEM_AUX + 1024;  // + alignment slack

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6), A=bf16 [7,10),
// B=bf16 [10,13), A/B K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

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
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&f)[32], int ncols) {
  if (ncols >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(f[8 * i + 0], f[8 * i + 1]);
      u.y = pack_bf16(f[8 * i + 2], f[8 * i + 3]);
      u.z = pack_bf16(f[8 * i + 4], f[8 * i + 5]);
      u.w = pack_bf16(f[8 * i + 6], f[8 * i + 7]);
      d4[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[i] = __float2bfloat16_rn(f[i]);
  }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&f)[32], int ncols) {
  if (ncols >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[i] = f[i];
  }
}
__device__ __forceinline__ void load_f32x32(const float* src, float (&f)[32], int ncols) {
  if (ncols >= 32 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = s4[i];
      f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (i < ncols) ? src[i] : 0.0f;
  }
}
// bias is warp-uniform per column chunk -> broadcast loads
__device__ __forceinline__ void add_bias32(float (&f)[32], const float* bias, int col0, int ncols) {
  if (bias == nullptr) return;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < ncols) f[i] += __ldg(bias + col0 + i);
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
            const EpiParams ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_TILES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_cos = reinterpret_cast<float*>(smem + SMEM_TILES + 256);
  float* s_sin = s_cos + ROPE_MAX_FT * 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  if (EPI == TOC3D_EPI_QKV_ROPE && warp >= 2) {
    for (int i = threadIdx.x - 64; i < ep.rope_ft * 16; i += 128) {
      s_cos[i] = ep.cos_axis[i];
      s_sin[i] = ep.sin_axis[i];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_idx = (tile % num_m) * BM;
        const int n_idx = (tile / num_m) * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m_idx);
          tma_load_2d(&tmB, &full_bar[stage], sa + A_BYTES, kb * BK, n_idx);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_desc = umma_desc_k_sw128(sa);
          const uint64_t b_desc = umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the 128B swizzle row: +2 in the >>4 address field
            umma_bf16_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), IDESC,
                         (kb | k) != 0 ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tmem_full[acc]);      // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_idx = (tile % num_m) * BM;
      const int n_idx = (tile / num_m) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int row = m_idx + quarter * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);

      if constexpr (EPI == TOC3D_EPI_SWIGLU) {
        // B rows are interleaved [32 x w1 | 32 x w2]; each chunk pair yields 32 hidden columns.
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out);
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          uint32_t v1[32], v2[32];
          tmem_ld_32x32(taddr + c * 64, v1);
          tmem_ld_32x32(taddr + c * 64 + 32, v2);
          tmem_ld_wait();
          const int col1 = n_idx + c * 64;             // GEMM column of the w1 part
          const int hcol = (n_idx >> 1) + c * 32;      // hidden column
          if (col1 < N) {
            float h[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float a = __uint_as_float(v1[i]), g = __uint_as_float(v2[i]);
              if (ep.bias != nullptr) {
                a += __ldg(ep.bias + col1 + i);
                g += __ldg(ep.bias + col1 + 32 + i);
              }
              h[i] = silu(a) * g;
            }
            if (row_ok) store_bf16x32(out + (size_t)row * ep.ldo + hcol, h, min(32, ep.ldo - hcol));
          }
        }
      } else {
        int rope_r = 0, rope_c = 0;
        if constexpr (EPI == TOC3D_EPI_QKV_ROPE) {
          int t = 0;
          if (row_ok) t = ep.rope_rows ? ep.rope_rows[row] : (row % ep.rope_slots);
          rope_r = t / ep.rope_ft;
          rope_c = t - rope_r * ep.rope_ft;
        }
        int rrow = -1, orow = -1;
        if constexpr (EPI == TOC3D_EPI_RESID) {
          if (row_ok) {
            rrow = ep.resid_mod > 0 ? (row % ep.resid_mod) : (ep.resid_map ? ep.resid_map[row] : row);
            orow = ep.out_map ? ep.out_map[row] : row;
          }
        }
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          const int col0 = n_idx + c * 32;
          if (col0 >= N) continue;   // warp-uniform
          const int ncols = min(32, N - col0);
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          add_bias32(f, ep.bias, col0, ncols);

          if constexpr (EPI == TOC3D_EPI_LINEAR) {
            if (ep.act == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
            } else if (ep.act == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
            }
            if (row_ok) {
              if (ep.out_f32) store_f32x32(reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0, f, ncols);
              else store_bf16x32(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + col0, f, ncols);
            }
          } else if constexpr (EPI == TOC3D_EPI_QKV_ROPE) {
            if (col0 < ep.rope_cols) {
              // head-dim 64: chunk parity selects the row-axis (first 32) or column-axis (last 32) angles
              const int pos = ((col0 >> 5) & 1) ? rope_c : rope_r;
              const float* cs = s_cos + pos * 16;
              const float* sn = s_sin + pos * 16;
              const float sc = (col0 < (ep.rope_cols >> 1)) ? ep.q_scale : 1.0f;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float x0 = f[2 * j], x1 = f[2 * j + 1];
                const float cj = cs[j], sj = sn[j];
                f[2 * j] = (x0 * cj - x1 * sj) * sc;
                f[2 * j + 1] = (x1 * cj + x0 * sj) * sc;
              }
            }
            if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + col0, f, ncols);
          } else if constexpr (EPI == TOC3D_EPI_RESID) {
            if (row_ok && orow != -1) {
              float r[32];
              if (rrow >= 0) load_f32x32(ep.resid + (size_t)rrow * ep.ldo + col0, r, ncols);
              else if (rrow == -2) load_f32x32(ep.out_alt + (size_t)row * ep.ldo + col0, r, ncols);
              else {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0.0f;
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = r[i] + f[i];
              float* dst = (orow >= 0) ? reinterpret_cast<float*>(ep.out) + (size_t)orow * ep.ldo
                                       : ep.out_alt + (size_t)row * ep.ldo;
              store_f32x32(dst + col0, f, ncols);
            }
          }
        }
      }
      // release this accumulator buffer to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

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

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols], 128B swizzle.
static int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
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

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int EPI>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep,
                  cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    TOC3D_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ta, tb, M, N, K, ep);
  TOC3D_CHECK_CUDA(cudaGetLastError());
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
  EpiParams ep;
  ep.bias = e->bias; ep.out = e->out; ep.ldo = e->ldo; ep.out_f32 = e->out_f32; ep.act = e->act;
  ep.resid = e->resid; ep.resid_map = e->resid_map; ep.resid_mod = e->resid_mod; ep.out_map = e->out_map;
  ep.out_alt = e->out_alt; ep.rope_rows = e->rope_rows; ep.rope_slots = e->rope_slots; ep.rope_ft = e->rope_ft;
  ep.rope_cols = e->rope_cols; ep.q_scale = e->q_scale; ep.cos_axis = e->cos_axis; ep.sin_axis = e->sin_axis;
  TOC3D_REQUIRE(ep.ldo > 0, kErrBadArg, "toc3d_gemm_bf16: ldo must be positive");
  if (kind == TOC3D_EPI_QKV_ROPE) {
    TOC3D_REQUIRE(ep.cos_axis && ep.sin_axis && ep.rope_ft > 0 && ep.rope_ft <= ROPE_MAX_FT, kErrBadArg,
                  "toc3d_gemm_bf16: bad RoPE tables (ft=%d)", ep.rope_ft);
    TOC3D_REQUIRE(ep.rope_rows || ep.rope_slots > 0, kErrBadArg, "toc3d_gemm_bf16: rope_rows or rope_slots required");
    TOC3D_REQUIRE(ep.rope_cols % 128 == 0 && ep.rope_cols <= N, kErrBadArg, "toc3d_gemm_bf16: rope_cols %d", ep.rope_cols);
  }
  if (kind == TOC3D_EPI_RESID) {
    TOC3D_REQUIRE(ep.resid_mod > 0 ? ep.resid != nullptr : true, kErrBadArg, "toc3d_gemm_bf16: resid_mod needs resid");
    TOC3D_REQUIRE(ep.resid != nullptr || ep.resid_map != nullptr, kErrBadArg, "toc3d_gemm_bf16: RESID needs resid");
  }
  if (kind == TOC3D_EPI_SWIGLU) TOC3D_REQUIRE(N % 64 == 0, kErrBadArg, "toc3d_gemm_bf16: SWIGLU needs N %% 64 == 0");
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, M, K, lda, BM);
  if (rc) return rc;
  rc = make_tmap(&tb, B, N, K, ldb, BN);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (kind) {
    case TOC3D_EPI_LINEAR: return launch<TOC3D_EPI_LINEAR>(ta, tb, M, N, K, ep, st);
    case TOC3D_EPI_QKV_ROPE: return launch<TOC3D_EPI_QKV_ROPE>(ta, tb, M, N, K, ep, st);
    case TOC3D_EPI_RESID: return launch<TOC3D_EPI_RESID>(ta, tb, M, N, K, ep, st);
    case TOC3D_EPI_SWIGLU: return launch<TOC3D_EPI_SWIGLU>(ta, tb, M, N, K, ep, st);
    default: TOC3D_REQUIRE(false, kErrBadArg, "toc3d_gemm_bf16: unknown epilogue kind %d", kind);
  }
  return 0;
}
