// Small-batch MLP building blocks shared by the gate and policy-head kernels:
// a CTA owns GATE_FT frames whose activations live in shared memory; weight rows are
// streamed from global memory once per CTA with coalesced 16-byte loads.
#pragma once
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

constexpr int GATE_FT = 4;       // frames per CTA (single-CTA variant)
constexpr int GATE_THREADS = 512;
// Cluster variant (large batches): a cluster of CL_RANKS CTAs owns CL_FT frames.  Every CTA keeps the
// activations of all CL_FT frames in its shared memory, computes only 1/CL_RANKS of each layer's output
// rows (so it streams 1/8 of the weights instead of all of them) and broadcasts its rows into the other
// CTAs' copies through distributed shared memory; barrier.cluster separates the layers.
constexpr int CL_RANKS = 8;
constexpr int CL_FT = 16;

__host__ __device__ inline int64_t al4(int64_t n) { return (n + 3) & ~(int64_t)3; }

// Cluster variant: opt-in (AMOE_MLP_CLUSTER=1, batches >= 64).  Measured on B200 at batch 256 it is SLOWER than
// the single-CTA kernels (gate 257 vs 169 us, policy head 217 vs 183 us): with 16 frames per cluster the
// shared-memory reads of the activations (one per FMA group) and the cluster barriers outweigh the 8x smaller
// weight stream; kept because it is correct (tested against the single-CTA kernels) and is the starting point
// for a tensor-core version of these layers.
inline bool mlp_use_cluster(int B) {
  const char* e = getenv("AMOE_MLP_CLUSTER");
  return (e ? atoi(e) : 0) != 0 && B >= 64;
}

// y[f][o] = act(b[o] + sum_i W[o][i] * x[f][i]),  x,y in shared memory.
// The kernels built on this are bound by the L2->SM latency of streaming every weight row once per
// CTA, so the mapping maximises independent loads in flight: 8 lanes share one output row (each
// lane owns every 8th 16-byte piece -> a row is read as full 128-byte lines), a warp works on 4
// rows at once, and a lane issues up to 8 independent 16-byte loads before the first FMA.
constexpr int MLP_LPR = 8;             // lanes per output row
constexpr int MLP_RPW = 32 / MLP_LPR;  // rows per warp pass
constexpr int MLP_BATCH = 8;           // 16-byte weight loads in flight per lane

template <bool CL>
__device__ __forceinline__ void mlp_sync() {
  if (CL) cg::this_cluster().sync();
  else __syncthreads();
}

template <int FT = GATE_FT, bool CL = false>
__device__ __forceinline__ void linear_ft(const float* __restrict__ Wg, const float* __restrict__ bg,
                                          const float* x, int x_ld, int in_dim, float* y, int y_ld,
                                          int out_dim, bool relu) {
  constexpr int GATE_FT = FT;   // (shadows the single-CTA constant inside this function)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int sub = lane / MLP_LPR, l = lane % MLP_LPR;
  const bool vec = (in_dim & 3) == 0 && (x_ld & 3) == 0;
  // rows this CTA computes: everything, or its contiguous slice of the cluster's split
  int r_begin = 0, r_end = out_dim;
  if (CL) {
    const int per = ((out_dim + CL_RANKS - 1) / CL_RANKS + 3) & ~3;   // multiple of 4: the all-gather moves float4s
    r_begin = min(out_dim, (int)cg::this_cluster().block_rank() * per);
    r_end = min(out_dim, r_begin + per);
  }
  for (int o0 = r_begin + warp * MLP_RPW; o0 < r_end; o0 += nwarp * MLP_RPW) {
    const int o = o0 + sub;
    const bool row_ok = o < r_end;
    float acc[GATE_FT];
#pragma unroll
    for (int f = 0; f < GATE_FT; ++f) acc[f] = 0.f;
    if (vec) {
      const int n4 = in_dim >> 2;
      const float4* wr = reinterpret_cast<const float4*>(Wg + (int64_t)(row_ok ? o : 0) * in_dim);
      for (int j0 = l; j0 < n4; j0 += MLP_LPR * MLP_BATCH) {
        float4 w4[MLP_BATCH];
#pragma unroll
        for (int b = 0; b < MLP_BATCH; ++b) {
          const int j = j0 + b * MLP_LPR;
          w4[b] = (row_ok && j < n4) ? __ldg(wr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < MLP_BATCH; ++b) {
          const int j = j0 + b * MLP_LPR;
          if (j < n4) {
#pragma unroll
            for (int f = 0; f < GATE_FT; ++f) {
              const float4 x4 = *reinterpret_cast<const float4*>(x + f * x_ld + (j << 2));
              acc[f] = fmaf(w4[b].x, x4.x, acc[f]);
              acc[f] = fmaf(w4[b].y, x4.y, acc[f]);
              acc[f] = fmaf(w4[b].z, x4.z, acc[f]);
              acc[f] = fmaf(w4[b].w, x4.w, acc[f]);
            }
          }
        }
      }
    } else {
      const float* wr = Wg + (int64_t)(row_ok ? o : 0) * in_dim;
      for (int i = l; i < in_dim; i += MLP_LPR) {
        const float wv = row_ok ? __ldg(wr + i) : 0.f;
#pragma unroll
        for (int f = 0; f < GATE_FT; ++f) acc[f] = fmaf(wv, x[f * x_ld + i], acc[f]);
      }
    }
    // reduce over the 8 lanes of the row
#pragma unroll
    for (int f = 0; f < GATE_FT; ++f) {
#pragma unroll
      for (int s = MLP_LPR / 2; s > 0; s >>= 1) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], s);
    }
    if (l == 0 && row_ok) {
      const float bv = __ldg(bg + o);
#pragma unroll
      for (int f = 0; f < GATE_FT; ++f) {
        float v = acc[f] + bv;
        y[f * y_ld + o] = relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
  if (CL) {
    // all-gather: this CTA's column slice [r_begin, r_end) of every frame goes to the 7 other ranks as
    // coalesced 16-byte distributed-shared-memory stores (scattered 4-byte remote stores cost ~10x more)
    __syncthreads();
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int c_end = min((out_dim + 3) & ~3, (r_end + 3) & ~3);     // 4-float chunks; padding columns ride along
    const int n4 = r_begin < c_end ? (c_end - r_begin) >> 2 : 0;
    const int per_rank = GATE_FT * n4;
    for (int i = threadIdx.x; i < per_rank * (CL_RANKS - 1); i += blockDim.x) {
      const int rr = i / per_rank, k = i - rr * per_rank;
      const int dst_rank = rr + (rr >= rank ? 1 : 0);
      const int f = k / n4, c4 = k - f * n4;
      const int off = f * y_ld + r_begin + (c4 << 2);
      const float4 v = *reinterpret_cast<const float4*>(y + off);
      *reinterpret_cast<float4*>(cl.map_shared_rank(y, (unsigned)dst_rank) + off) = v;
    }
  }
  mlp_sync<CL>();
}

// nn.LayerNorm(dim) (eps 1e-5, biased variance) in place on x[f][0..dim); warp f handles frame f
template <int FT = GATE_FT, bool CL = false>
__device__ __forceinline__ void layernorm_ft(float* x, int x_ld, int dim, const float* __restrict__ g,
                                             const float* __restrict__ b) {
  constexpr int GATE_FT = FT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(FT <= GATE_THREADS / 32, "one warp per frame");
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
  __syncthreads();   // cluster variant: every rank normalises its own copy (identical results); no remote access
}

// cluster variant: rank r stores frames r, r+8, ... (all ranks hold identical copies)
template <int FT = GATE_FT, bool CL = false>
__device__ __forceinline__ void store_rows(float* dst, int64_t dst_ld, const float* src, int src_ld,
                                           int dim, int f0, int B) {
  const int rank = CL ? (int)cg::this_cluster().block_rank() : 0;
  for (int i = threadIdx.x; i < FT * dim; i += blockDim.x) {
    int f = i / dim, c = i - f * dim;
    if (CL && (f % CL_RANKS) != rank) continue;
    if (f0 + f < B) dst[(int64_t)(f0 + f) * dst_ld + c] = src[f * src_ld + c];
  }
}

