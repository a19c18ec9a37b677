// Probe (diagnostic, not part of the library): issue rates that bound the softmax warps of the attention kernels.
//   A  MUFU.EX2 warp-instruction rate with 1 / 2 / 4 warps per SM sub-partition
//   B  the exp chunk of softmax_rows (32 x {FFMA, EX2, FADD} + 16 F2FP) on register data, 1 / 2 warps per sub-partition
//   C  tcgen05.ld 32x32b.x32: issue-to-data latency and back-to-back rate, 1 / 2 warps per TMEM lane quarter
//   D  max pass + exp pass over TMEM exactly as softmax_rows does them, 192 columns, 4 / 8 warps
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o softmax_probe softmax_probe.cu
#include "../../toc3d_b200/csrc/common.cuh"
#include <cstdio>
using namespace toc3d;

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// mode 0: A, mode 1: B, mode 2: C latency, mode 3: C throughput, mode 4: D
__global__ void probe(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tbase = 0;
  if (mode >= 2) {
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    tbase = tmem_slot;
    // fill this warp's lane quarter with finite values
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(0.001f * (float)(i + lane));
    for (int c = 0; c < 512; c += 16) tmem_st_32x16(tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c, z);
    tmem_st_wait();
    tcgen05_fence_before();
  }
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
  float acc = 0.f;
  long long t0 = clock64();
  if (mode == 0) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = -0.01f * (float)(lane + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = ex2a(x[i]) - 1.5f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += x[i];
  } else if (mode == 1) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.01f * (float)(lane + i);
    float s0 = 0.f, s1 = 0.f;
    uint32_t pk = 0;
    for (int it = 0; it < iters; ++it) {
      const float mneg = -0.5f - 1e-6f * (float)it;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ex2a(fmaf(v[2 * i], 1.4426950408889634f, mneg));
        const float p1 = ex2a(fmaf(v[2 * i + 1], 1.4426950408889634f, mneg));
        s0 += p0;
        s1 += p1;
        pk ^= pack_bf16(p0, p1);
      }
    }
    acc = s0 + s1 + __uint_as_float(pk & 0xffu);
  } else if (mode == 2) {
    uint32_t v[32];
    for (int it = 0; it < iters; ++it) {
      tmem_ld_32x32(lane_base + (uint32_t)((it & 3) * 32), v);
      tmem_ld_wait();
      acc += __uint_as_float(v[it & 31]);
    }
  } else if (mode == 3) {
    uint32_t a[32], b[32], c[32], d[32];
    for (int it = 0; it < iters; it += 4) {
      tmem_ld_32x32(lane_base, a);
      tmem_ld_32x32(lane_base + 32u, b);
      tmem_ld_32x32(lane_base + 64u, c);
      tmem_ld_32x32(lane_base + 96u, d);
      tmem_ld_wait();
      acc += __uint_as_float(a[it & 31]) + __uint_as_float(b[it & 31]) + __uint_as_float(c[it & 31]) + __uint_as_float(d[it & 31]);
    }
  } else if (mode == 4) {
    // as softmax_rows: max pass, exp pass, 6 chunks of 32 columns, P stored over S
    for (int it = 0; it < iters; ++it) {
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      uint32_t va[32], vb[32];
      const int nchunks = 6;
      tmem_ld_32x32(lane_base, va);
      for (int c = 0; c < nchunks; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(lane_base + (uint32_t)((c + 1) * 32), vb);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(va[i])); m1 = fmaxf(m1, __uint_as_float(va[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(va[i + 2])); m3 = fmaxf(m3, __uint_as_float(va[i + 3]));
        }
        tmem_ld_wait();
        if (c + 2 < nchunks) tmem_ld_32x32(lane_base + (uint32_t)((c + 2) * 32), va);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(vb[i])); m1 = fmaxf(m1, __uint_as_float(vb[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(vb[i + 2])); m3 = fmaxf(m3, __uint_as_float(vb[i + 3]));
        }
      }
      const float mneg = -fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * 1.4426950408889634f;
      if (it == 0 && lane == 0 && warp == 0) out[64] = clock64() - t0;
      float s0 = 0.f, s1 = 0.f;
      tmem_ld_32x32(lane_base, va);
      for (int c = 0; c < nchunks; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(lane_base + (uint32_t)((c + 1) * 32), vb);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2a(fmaf(__uint_as_float(va[2 * i]), 1.4426950408889634f, mneg));
          const float p1 = ex2a(fmaf(__uint_as_float(va[2 * i + 1]), 1.4426950408889634f, mneg));
          s0 += p0; s1 += p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st_32x16(lane_base + (uint32_t)(c * 16), pk);
        tmem_ld_wait();
        if (c + 2 < nchunks) tmem_ld_32x32(lane_base + (uint32_t)((c + 2) * 32), va);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2a(fmaf(__uint_as_float(vb[2 * i]), 1.4426950408889634f, mneg));
          const float p1 = ex2a(fmaf(__uint_as_float(vb[2 * i + 1]), 1.4426950408889634f, mneg));
          s0 += p0; s1 += p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st_32x16(lane_base + (uint32_t)((c + 1) * 16), pk);
      }
      tmem_st_wait();
      acc += s0 + s1;
      // restore finite scores for the next iteration (P overwrote the first 96 columns)
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(0.001f * (float)(i + lane));
      for (int c = 0; c < 96; c += 16) tmem_st_32x16(lane_base + (uint32_t)c, z);
      tmem_st_wait();
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  if (mode >= 2) {
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
      tcgen05_fence_after();
      tmem_dealloc(tbase, 512);
    }
  }
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 128 * sizeof(long long));
  cudaMalloc(&sink, 1024 * sizeof(float));
  long long h[128];
  const char* names[] = {"A ex2 only (16 per iter)", "B exp chunk on registers (32 ex2 per iter)", "C tcgen05.ld x32 latency (1 per iter)",
                         "C tcgen05.ld x32 4 in flight (per load)", "D max + exp pass over 192 TMEM columns (per unit)"};
  for (int mode = 0; mode < 5; ++mode) {
    const int iters = mode == 4 ? 50 : 400;
    for (int warps : {1, 4, 8, 16}) {
      if (mode >= 2 && warps > 8) continue;          // two 256-column slots
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(out, 0, 128 * sizeof(long long));
        probe<<<1, warps * 32>>>(mode, iters, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, out, 128 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("%-52s warps/SM %2d (per sub-partition %d): %8.1f clk per iter", names[mode], warps, (warps + 3) / 4, (double)mx / iters);
      if (mode == 4) printf("   [first max pass: %lld clk]", h[64]);
      printf("\n");
    }
  }
  return 0;
}
