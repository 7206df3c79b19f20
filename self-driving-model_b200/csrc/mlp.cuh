// Small-batch MLP building blocks shared by the gate and policy-head kernels:
// a CTA owns GATE_FT frames whose activations live in shared memory; weight rows are
// streamed from global memory once per CTA with coalesced 16-byte loads.
#pragma once
#include "common.cuh"

constexpr int GATE_FT = 4;       // frames per CTA
constexpr int GATE_THREADS = 512;

__host__ __device__ inline int64_t al4(int64_t n) { return (n + 3) & ~(int64_t)3; }

// y[f][o] = act(b[o] + sum_i W[o][i] * x[f][i]),  x,y in shared memory.
// A warp owns MLP_RPW output rows at a time so that MLP_RPW independent 16-byte weight loads per lane
// are in flight (the kernel is bound by the L2->SM latency of streaming the weight rows).
constexpr int MLP_RPW = 4;

__device__ __forceinline__ void linear_ft(const float* __restrict__ Wg, const float* __restrict__ bg,
                                          const float* x, int x_ld, int in_dim, float* y, int y_ld,
                                          int out_dim, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const bool vec = (in_dim & 3) == 0 && (x_ld & 3) == 0;
  for (int o0 = warp * MLP_RPW; o0 < out_dim; o0 += nwarp * MLP_RPW) {
    float acc[MLP_RPW][GATE_FT];
#pragma unroll
    for (int r = 0; r < MLP_RPW; ++r)
#pragma unroll
      for (int f = 0; f < GATE_FT; ++f) acc[r][f] = 0.f;
    if (vec) {
      const int n4 = in_dim >> 2;
      for (int i = lane; i < n4; i += 32) {
        float4 w4[MLP_RPW];
#pragma unroll
        for (int r = 0; r < MLP_RPW; ++r)
          w4[r] = (o0 + r < out_dim) ? __ldg(reinterpret_cast<const float4*>(Wg + (int64_t)(o0 + r) * in_dim) + i)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int f = 0; f < GATE_FT; ++f) {
          const float4 x4 = *reinterpret_cast<const float4*>(x + f * x_ld + (i << 2));
#pragma unroll
          for (int r = 0; r < MLP_RPW; ++r) {
            acc[r][f] = fmaf(w4[r].x, x4.x, acc[r][f]);
            acc[r][f] = fmaf(w4[r].y, x4.y, acc[r][f]);
            acc[r][f] = fmaf(w4[r].z, x4.z, acc[r][f]);
            acc[r][f] = fmaf(w4[r].w, x4.w, acc[r][f]);
          }
        }
      }
    } else {
      for (int i = lane; i < in_dim; i += 32) {
        float wv[MLP_RPW];
#pragma unroll
        for (int r = 0; r < MLP_RPW; ++r) wv[r] = (o0 + r < out_dim) ? __ldg(Wg + (int64_t)(o0 + r) * in_dim + i) : 0.f;
#pragma unroll
        for (int f = 0; f < GATE_FT; ++f) {
          const float xv = x[f * x_ld + i];
#pragma unroll
          for (int r = 0; r < MLP_RPW; ++r) acc[r][f] = fmaf(wv[r], xv, acc[r][f]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < MLP_RPW; ++r)
#pragma unroll
      for (int f = 0; f < GATE_FT; ++f) acc[r][f] = warp_sum(acc[r][f]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < MLP_RPW; ++r) {
        if (o0 + r < out_dim) {
          const float bv = __ldg(bg + o0 + r);
#pragma unroll
          for (int f = 0; f < GATE_FT; ++f) {
            float v = acc[r][f] + bv;
            y[f * y_ld + o0 + r] = relu ? fmaxf(v, 0.f) : v;
          }
        }
      }
    }
  }
  __syncthreads();
}

// nn.LayerNorm(dim) (eps 1e-5, biased variance) in place on x[f][0..dim); warp f handles frame f
__device__ __forceinline__ void layernorm_ft(float* x, int x_ld, int dim, const float* __restrict__ g,
                                             const float* __restrict__ b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < GATE_FT) {
    float* xr = x + warp * x_ld;
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) s += xr[i];
    float mean = warp_sum(s) / (float)dim;
    float v = 0.f;
    for (int i = lane; i < dim; i += 32) {
      float d = xr[i] - mean;
      v = fmaf(d, d, v);
    }
    float rstd = rsqrtf(warp_sum(v) / (float)dim + 1e-5f);
    for (int i = lane; i < dim; i += 32) xr[i] = (xr[i] - mean) * rstd * __ldg(g + i) + __ldg(b + i);
  }
  __syncthreads();
}

__device__ __forceinline__ void store_rows(float* dst, int64_t dst_ld, const float* src, int src_ld,
                                           int dim, int f0, int B) {
  for (int i = threadIdx.x; i < GATE_FT * dim; i += blockDim.x) {
    int f = i / dim, c = i - f * dim;
    if (f0 + f < B) dst[(int64_t)(f0 + f) * dst_ld + c] = src[f * src_ld + c];
  }
}

