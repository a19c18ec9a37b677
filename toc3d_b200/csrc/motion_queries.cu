// Motion-aware history-query encoder + scorer folding for ALL selector stages of a forward, two launches
// (SURVEY.md 8a row a12).  Replaces the reference's ~35 small ATen ops per stage:
//   MotionAwareQueryGuidedTokenSelector.get_motion_aware_queries   backbones/toc3d_utils.py:334-360
//   transform_reference_points                                     utils/misc.py:191-200
//   MLN.forward                                                    utils/misc.py:154-188
//   pos2posemb3d / pos2posemb1d / nerf_positional_encoding         utils/positional_encoding.py:14-81
// and, because the query scorer is linear up to its log-softmax (toc3d_utils.py:232-252), folds the encoded queries
// into a 2 x C matrix per (stage, frame) right away:  P = w_agg q,  A = scale P w_in,  c = scale P b_in + b_agg.
//
// Everything is fp32 (the reference runs this in fp32; the 1-D timestamp embedding in fp64 when the timestamps arrive
// as float64, streampetr_head.py:354,371).  The work is tiny (64 queries x 256 channels, 0.6 MFLOP per query and
// stage); what matters is that it is ONE dependency-free side branch of the graph instead of a hundred nodes.
//
// Per-stage parameter blob (fp32, packed once per load_state_dict by toc3d_b200/backbone.py; Linear weights transposed
// to [in][out] so that thread = output channel reads them coalesced).  Offsets in floats, D = 256 (query_dim):
//   dimt128[128] dimt256[256] pc_range[8: 6 used]
//   qe0_wT[384][D] qe0_b[D]  qe2_wT[D][D] qe2_b[D]                                       query_embedding.{0,2}
//   pe_red_wT[180][D] pe_red_b[D] pe_gam_wT[D][D] pe_gam_b[D] pe_bet_wT[D][D] pe_bet_b[D]   ego_pose_pe (MLN)
//   qs_red_wT[180][D] qs_red_b[D] qs_gam_wT[D][D] qs_gam_b[D] qs_bet_wT[D][D] qs_bet_b[D]   ego_pose_queries (MLN)
//   te_wT[D][D] te_b[D] te_lnw[D] te_lnb[D]                                              time_embedding.{0,1}
//   w_in[D][C] b_in[D] w_agg[2][Q] b_agg[4: 2 used]                                      input_proj.0, aggregate.0
#include "../../include/toc3d_b200.h"
#include "common.cuh"

namespace toc3d {
namespace mq {

constexpr int D = 256;          // query_dim (toc3d_utils.py:300, never overridden by a config)
constexpr int QPB = 4;          // queries per thread block
constexpr int NERF = 180;       // 15 motion scalars x 6 bands x {sin, cos}
constexpr float LN_EPS_DEFAULT = 1e-5f;

struct Layout {
  int dimt128, dimt256, pc, qe0_w, qe0_b, qe2_w, qe2_b;
  int mln[2][6];  // red_w red_b gam_w gam_b bet_w bet_b for ego_pose_pe, ego_pose_queries
  int te_w, te_b, te_lnw, te_lnb, w_in, b_in, w_agg, b_agg, total;
};

__host__ __device__ inline Layout make_layout(int Q, int C) {
  Layout L;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  L.dimt128 = take(128); L.dimt256 = take(256); L.pc = take(8);
  L.qe0_w = take(384 * D); L.qe0_b = take(D); L.qe2_w = take(D * D); L.qe2_b = take(D);
  for (int m = 0; m < 2; ++m) {
    L.mln[m][0] = take(NERF * D); L.mln[m][1] = take(D);
    L.mln[m][2] = take(D * D); L.mln[m][3] = take(D);
    L.mln[m][4] = take(D * D); L.mln[m][5] = take(D);
  }
  L.te_w = take(D * D); L.te_b = take(D); L.te_lnw = take(D); L.te_lnb = take(D);
  L.w_in = take(D * C); L.b_in = take(D); L.w_agg = take(2 * Q); L.b_agg = take(4);
  L.total = o;
  return L;
}

// out[q][t] = b[t] + sum_k in[q][k] * wT[k][t]   for the QPB queries of the block; thread t = output channel
template <int LD>
__device__ __forceinline__ void dense(const float* __restrict__ wT, const float* __restrict__ b, const float (*in)[LD], int K,
                                      float (&acc)[QPB]) {
  const int t = threadIdx.x;
  const float bv = b[t];
#pragma unroll
  for (int q = 0; q < QPB; ++q) acc[q] = bv;
  int k = 0;
  for (; k + 16 <= K; k += 16) {
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = __ldg(wT + (size_t)(k + j) * D + t);
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int q = 0; q < QPB; ++q) acc[q] = fmaf(in[q][k + j], w[j], acc[q]);
  }
  for (; k < K; ++k) {
    const float w = __ldg(wT + (size_t)k * D + t);
#pragma unroll
    for (int q = 0; q < QPB; ++q) acc[q] = fmaf(in[q][k], w, acc[q]);
  }
}

// mean / rstd over the D channels of each of the QPB rows held one value per thread (blockDim.x == D)
__device__ __forceinline__ void row_stats(const float (&v)[QPB], float (*red)[QPB][2], float (&mean)[QPB], float (&rstd)[QPB]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < QPB; ++q) {
    const float s = warp_sum(v[q]);
    if (lane == 0) red[warp][q][0] = s;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < QPB; ++q) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < D / 32; ++w) s += red[w][q][0];
    mean[q] = s * (1.0f / D);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < QPB; ++q) {
    const float d = v[q] - mean[q];
    const float s = warp_sum(d * d);
    if (lane == 0) red[warp][q][1] = s;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < QPB; ++q) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < D / 32; ++w) s += red[w][q][1];
    rstd[q] = rsqrtf(s * (1.0f / D) + LN_EPS_DEFAULT);     // biased variance, like F.layer_norm
  }
  __syncthreads();
}

// grid (ceil(Q / QPB), Bf, S), D threads
template <bool TS_F64>
__global__ void __launch_bounds__(D)
motion_queries_kernel(const float* __restrict__ blob, int blob_stride, int Q, int C, const float* __restrict__ temp_queries,
                      const float* __restrict__ ref_points, const float* __restrict__ vel, const void* __restrict__ timestamp,
                      const float* __restrict__ ego_pose, const float* __restrict__ ego_pose_inv, int Bf,
                      float* __restrict__ q_out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float s_pe[QPB][384];
  __shared__ float s_nerf[QPB][NERF + 4];
  __shared__ float s_tpe[QPB][D];
  __shared__ float s_h[QPB][D];
  __shared__ float s_motion[QPB][16];
  __shared__ float s_ref[QPB][4];
  __shared__ float s_red[D / 32][QPB][2];
  const int t = threadIdx.x;
  const int q0 = blockIdx.x * QPB, f = blockIdx.y, s = blockIdx.z;
  const Layout L = make_layout(Q, C);
  const float* W = blob + (size_t)s * blob_stride;

  // ---- raw inputs of the block's queries (rows beyond Q are clamped; their results are not stored)
  if (t < QPB * 3) {
    const int q = t / 3, c = t % 3;
    const int qi = min(q0 + q, Q - 1);
    const float* r = ref_points + ((size_t)f * Q + qi) * 3;
    const float* m = ego_pose_inv + (size_t)f * 16 + c * 4;
    // (matrix @ [x y z 1])[c], misc.py:191-200, then (p - pc[:3]) / (pc[3:6] - pc[:3]), toc3d_utils.py:348
    const float p = fmaf(m[0], r[0], fmaf(m[1], r[1], fmaf(m[2], r[2], m[3])));
    const float* pc = W + L.pc;
    s_ref[q][c] = (p - pc[c]) / (pc[3 + c] - pc[c]);
  }
  if (t >= 32 && t < 32 + QPB * 15) {
    const int q = (t - 32) / 15, d = (t - 32) % 15;
    const int qi = min(q0 + q, Q - 1);
    const size_t row = (size_t)f * Q + qi;
    float v;
    if (d < 2) v = vel[row * 2 + d];
    else if (d == 2) v = TS_F64 ? (float)reinterpret_cast<const double*>(timestamp)[row] : reinterpret_cast<const float*>(timestamp)[row];
    else v = ego_pose[row * 16 + (d - 3)];          // rows 0..2 of the 4x4, flattened (12 values)
    s_motion[q][d] = v;
  }
  __syncthreads();
  // pos2posemb3d: [y | x | z] x 128, element i: (i even ? sin : cos)(p * 2pi / dim_t[i])
  for (int idx = t; idx < QPB * 384; idx += D) {
    const int q = idx / 384, i = idx % 384;
    const int grp = i >> 7, j = i & 127;
    const float p = s_ref[q][grp == 0 ? 1 : (grp == 1 ? 0 : 2)];
    const float a = (p * 6.283185307179586f) / W[L.dimt128 + j];
    s_pe[q][i] = (j & 1) ? cosf(a) : sinf(a);
  }
  // nerf encoding of the 15 motion scalars: [sin(m f0) cos(m f0) sin(m f1) ...], f = 1, 2, 4, 8, 16, 32
  for (int idx = t; idx < QPB * NERF; idx += D) {
    const int q = idx / NERF, e = idx % NERF;
    const int band = e / 30, sc = (e / 15) & 1, d = e % 15;
    const float v = s_motion[q][d] * (float)(1 << band);
    s_nerf[q][e] = sc ? cosf(v) : sinf(v);
  }
  // pos2posemb1d of the timestamp (fp64 when the caller's timestamps are fp64, then rounded once)
  for (int q = 0; q < QPB; ++q) {
    const int qi = min(q0 + q, Q - 1);
    const size_t row = (size_t)f * Q + qi;
    if (TS_F64) {
      const double ts = reinterpret_cast<const double*>(timestamp)[row];
      const double a = (ts * 6.283185307179586) / (double)W[L.dimt256 + t];
      s_tpe[q][t] = (float)((t & 1) ? cos(a) : sin(a));
    } else {
      const float ts = reinterpret_cast<const float*>(timestamp)[row];
      const float a = (ts * 6.283185307179586f) / W[L.dimt256 + t];
      s_tpe[q][t] = (t & 1) ? cosf(a) : sinf(a);
    }
  }
  __syncthreads();

  float acc[QPB], pos[QPB], gam[QPB], bet[QPB], mean[QPB], rstd[QPB];
  // temp_pos = query_embedding(posemb3d)
  dense<384>(W + L.qe0_w, W + L.qe0_b, s_pe, 384, acc);
#pragma unroll
  for (int q = 0; q < QPB; ++q) s_h[q][t] = fmaxf(acc[q], 0.f);
  __syncthreads();
  dense<D>(W + L.qe2_w, W + L.qe2_b, s_h, D, pos);
  __syncthreads();
  // temp_pos = MLN_pe(temp_pos, motion)
  dense<NERF + 4>(W + L.mln[0][0], W + L.mln[0][1], s_nerf, NERF, acc);
#pragma unroll
  for (int q = 0; q < QPB; ++q) s_h[q][t] = fmaxf(acc[q], 0.f);
  __syncthreads();
  dense<D>(W + L.mln[0][2], W + L.mln[0][3], s_h, D, gam);
  dense<D>(W + L.mln[0][4], W + L.mln[0][5], s_h, D, bet);
  row_stats(pos, s_red, mean, rstd);
#pragma unroll
  for (int q = 0; q < QPB; ++q) pos[q] = gam[q] * ((pos[q] - mean[q]) * rstd[q]) + bet[q];
  // temp_pos += time_embedding(posemb1d(t))
  dense<D>(W + L.te_w, W + L.te_b, s_tpe, D, acc);
  row_stats(acc, s_red, mean, rstd);
  {
    const float g = W[L.te_lnw + t], b = W[L.te_lnb + t];
#pragma unroll
    for (int q = 0; q < QPB; ++q) pos[q] += (acc[q] - mean[q]) * rstd[q] * g + b;
  }
  // queries = MLN_q(temp_queries, motion) + temp_pos
  dense<NERF + 4>(W + L.mln[1][0], W + L.mln[1][1], s_nerf, NERF, acc);
  __syncthreads();                               // everyone is done reading s_h (gamma / beta of the first MLN)
#pragma unroll
  for (int q = 0; q < QPB; ++q) s_h[q][t] = fmaxf(acc[q], 0.f);
  __syncthreads();
  dense<D>(W + L.mln[1][2], W + L.mln[1][3], s_h, D, gam);
  dense<D>(W + L.mln[1][4], W + L.mln[1][5], s_h, D, bet);
  float x[QPB];
#pragma unroll
  for (int q = 0; q < QPB; ++q) x[q] = temp_queries[((size_t)f * Q + min(q0 + q, Q - 1)) * D + t];
  row_stats(x, s_red, mean, rstd);
#pragma unroll
  for (int q = 0; q < QPB; ++q) {
    if (q0 + q < Q)
      q_out[(((size_t)s * Bf + f) * Q + q0 + q) * D + t] = gam[q] * ((x[q] - mean[q]) * rstd[q]) + bet[q] + pos[q];
  }
}

// Folding: grid (C / 128, Bf, S), 128 threads.  P = w_agg (2 x Q) q (Q x D) in shared memory, then
// A[o][ch] = scale * sum_c P[o][c] w_in[c][ch],  c[o] = scale * P[o] . b_in + b_agg[o].
__global__ void __launch_bounds__(128)
fold_queries_kernel(const float* __restrict__ blob, int blob_stride, int Q, int C, const float* __restrict__ q_all, int Bf,
                    float scale, float* __restrict__ A_out, float* __restrict__ c_out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float s_p[2][D];
  const int f = blockIdx.y, s = blockIdx.z;
  const Layout L = make_layout(Q, C);
  const float* W = blob + (size_t)s * blob_stride;
  const float* q = q_all + ((size_t)s * Bf + f) * Q * D;
  const float* w_agg = W + L.w_agg;
  for (int c = threadIdx.x; c < D; c += 128) {
    float p0 = 0.f, p1 = 0.f;
    int j = 0;
    for (; j + 8 <= Q; j += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = q[(size_t)(j + u) * D + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) { p0 = fmaf(w_agg[j + u], v[u], p0); p1 = fmaf(w_agg[Q + j + u], v[u], p1); }
    }
    for (; j < Q; ++j) { const float v = q[(size_t)j * D + c]; p0 = fmaf(w_agg[j], v, p0); p1 = fmaf(w_agg[Q + j], v, p1); }
    s_p[0][c] = p0;
    s_p[1][c] = p1;
  }
  __syncthreads();
  const int ch = blockIdx.x * 128 + threadIdx.x;
  const float* w_in = W + L.w_in;
  if (ch < C) {
    float a0 = 0.f, a1 = 0.f;
    for (int c = 0; c < D; c += 16) {
      float w[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) w[u] = __ldg(w_in + (size_t)(c + u) * C + ch);
#pragma unroll
      for (int u = 0; u < 16; ++u) { a0 = fmaf(s_p[0][c + u], w[u], a0); a1 = fmaf(s_p[1][c + u], w[u], a1); }
    }
    A_out[(((size_t)s * Bf + f) * 2 + 0) * C + ch] = a0 * scale;
    A_out[(((size_t)s * Bf + f) * 2 + 1) * C + ch] = a1 * scale;
  }
  if (blockIdx.x == 0 && threadIdx.x < 2) {
    float acc = 0.f;
    for (int c = 0; c < D; ++c) acc = fmaf(s_p[threadIdx.x][c], W[L.b_in + c], acc);
    c_out[((size_t)s * Bf + f) * 2 + threadIdx.x] = acc * scale + W[L.b_agg + threadIdx.x];
  }
}

}  // namespace mq
}  // namespace toc3d

using namespace toc3d;

extern "C" int64_t toc3d_motion_blob_floats(int32_t Q, int32_t C) {
  if (Q <= 0 || C <= 0) return -1;
  return mq::make_layout(Q, C).total;
}

extern "C" int toc3d_motion_queries_fold(const float* blob, int64_t blob_stride, int32_t S, int32_t Bf, int32_t Q, int32_t C,
                                         const float* temp_queries, const float* ref_points, const float* vel,
                                         const void* timestamp, int32_t timestamp_is_f64, const float* ego_pose,
                                         const float* ego_pose_inv, float scale, float* q_out, float* A_out, float* c_out,
                                         void* stream) {
  TOC3D_REQUIRE(blob && temp_queries && ref_points && vel && timestamp && ego_pose && ego_pose_inv && q_out && A_out && c_out,
                kErrBadArg, "toc3d_motion_queries_fold: null pointer");
  TOC3D_REQUIRE(S > 0 && Bf > 0 && Q > 0 && Q % 2 == 0 && C > 0 && C % 4 == 0, kErrBadArg,
                "toc3d_motion_queries_fold: bad shape (S=%d Bf=%d Q=%d C=%d; Q must be even, C a multiple of 4)", S, Bf, Q, C);
  TOC3D_REQUIRE(blob_stride >= mq::make_layout(Q, C).total && blob_stride < (1ll << 31), kErrBadArg,
                "toc3d_motion_queries_fold: blob stride %lld smaller than the layout (%d floats)", (long long)blob_stride,
                mq::make_layout(Q, C).total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 g1((Q + mq::QPB - 1) / mq::QPB, Bf, S);
  if (timestamp_is_f64)
    TOC3D_CHECK_CUDA(launch_pdl(mq::motion_queries_kernel<true>, g1, dim3(mq::D), 0, st, 1, blob, (int)blob_stride, Q, C, temp_queries,
                                ref_points, vel, timestamp, ego_pose, ego_pose_inv, Bf, q_out));
  else
    TOC3D_CHECK_CUDA(launch_pdl(mq::motion_queries_kernel<false>, g1, dim3(mq::D), 0, st, 1, blob, (int)blob_stride, Q, C, temp_queries,
                                ref_points, vel, timestamp, ego_pose, ego_pose_inv, Bf, q_out));
  const dim3 g2((C + 127) / 128, Bf, S);
  TOC3D_CHECK_CUDA(launch_pdl(mq::fold_queries_kernel, g2, dim3(128), 0, st, 1, blob, (int)blob_stride, Q, C, (const float*)q_out, Bf,
                              scale, A_out, c_out));
  return 0;
}
