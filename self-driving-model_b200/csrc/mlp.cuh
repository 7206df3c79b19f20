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

// every tensor of a flat parameter buffer starts on an 8-element boundary (32 B as fp32, 16 B in the bf16 copy)
__host__ __device__ inline int64_t al8(int64_t n) { return (n + 7) & ~(int64_t)7; }

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

// ---- tensor-core variant (bf16 inference mode, 16 frames per CTA) -------------------------------------------
// The same layer as one warp-level GEMM: M = 16 frames, N = 8 output rows per warp pass, K in steps of 8 on
// mma.sync m16n8k8 TF32 with fp32 accumulation.  These layers are a dependent chain of skinny GEMMs whose cost is
// streaming the weights once per CTA through L2 latency with ~64 KB in flight per SM - not math - so: 16 frames
// per CTA instead of 4 (4x fewer weight passes), weights stored as bf16 (half the bytes; what the reference's
// autocast multiplies with, and exact in TF32), a fragment layout that needs no shuffle reductions, and two
// batches of MMA_U independent 16-byte weight loads in flight per lane.  Activations stay fp32 in shared memory;
// the tensor core reads their upper 19 bits (TF32 truncation, <= 2^-10 relative, ~8x below the bf16 rounding of
// the convolution features these layers consume).  tcgen05 would need M = 128 rows per CTA - two CTAs for the whole
// 256-frame batch, serialising the weight stream on two SMs; mma.sync is the right size for M = 16.
// Measured on B200 (ncu, batch 256): cvt.rna.tf32.f32 expands to ~10 instructions on sm_100a and made the first
// version issue-bound (11.4 M instructions for the gate); a cp.async shared-memory ring with a CTA barrier per
// 16 KB stage was slower (210-240 us per kernel) than register prefetch (140-170 us) and was removed.
// Inside a 32-wide K block the K index is permuted identically for both operands (lane q owns the eight
// consecutive elements 8q..8q+7 of its row: elements 2s, 2s+1 feed k8 step s), so activations and weights are
// both read as plain 16-byte vectors.
constexpr int MMA_FT = 16;   // frames per CTA of the tensor-core variant
constexpr int MMA_U = 4;     // 32-wide K blocks per prefetch batch

__host__ __device__ inline int ld_tc(int n) { return ((n + 31) & ~31) + 4; }   // row stride = 4 mod 32 floats: conflict-free 16-byte fragment loads

__device__ __forceinline__ void mma_m16n8k8_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// one 32-wide K block: fp32 activations of frames g / g+8 at shared addresses xa / xb (8 floats each),
// w = the lane's 8 bf16 weights (two per 32-bit word, lower address in the low half)
__device__ __forceinline__ void mma_block32(float (&c)[4], uint32_t xa, uint32_t xb, const uint4& w) {
  const uint4 lo0 = lds128u(xa), lo1 = lds128u(xa + 16), hi0 = lds128u(xb), hi1 = lds128u(xb + 16);
  mma_m16n8k8_tf32(c, lo0.x, hi0.x, lo0.y, hi0.y, w.x << 16, w.x & 0xffff0000u);
  mma_m16n8k8_tf32(c, lo0.z, hi0.z, lo0.w, hi0.w, w.y << 16, w.y & 0xffff0000u);
  mma_m16n8k8_tf32(c, lo1.x, hi1.x, lo1.y, hi1.y, w.z << 16, w.z & 0xffff0000u);
  mma_m16n8k8_tf32(c, lo1.z, hi1.z, lo1.w, hi1.w, w.w << 16, w.w & 0xffff0000u);
}

// y[f][o] = act(b[o] + sum_i W[o][i] * x[f][i]) for f < 16; Wb = W as bf16 (same [out][in] layout),
// in_dim % 32 == 0, x_ld % 4 == 0, x and the rows of Wb 16-byte aligned
static __device__ __noinline__ void linear_mma16(const __nv_bfloat16* __restrict__ Wb, const float* __restrict__ bg,
                                                 const float* x, int x_ld, int in_dim, float* y, int y_ld,
                                                 int out_dim, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int ngroups = (out_dim + 7) >> 3, nblk = in_dim >> 5;
  const int nfull = nblk / MMA_U;                  // full prefetch batches
  const uint32_t xa = (uint32_t)__cvta_generic_to_shared(x + g * x_ld + 8 * q);   // frame g
  const uint32_t xb = xa + (uint32_t)(8 * x_ld) * 4u;                             // frame g + 8
  for (int grp = warp; grp < ngroups; grp += nwarp) {
    const int n = grp * 8 + g;   // rows past out_dim re-read the last row; their outputs are not stored
    const uint4* wr = reinterpret_cast<const uint4*>(Wb + (int64_t)(n < out_dim ? n : out_dim - 1) * in_dim) + q;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    uint4 w[MMA_U];
    if (nfull > 0) {
#pragma unroll
      for (int u = 0; u < MMA_U; ++u) w[u] = __ldg(wr + u * 4);
    }
    for (int bt = 0; bt < nfull; ++bt) {
      uint4 wn[MMA_U];
      if (bt + 1 < nfull) {
#pragma unroll
        for (int u = 0; u < MMA_U; ++u) wn[u] = __ldg(wr + ((bt + 1) * MMA_U + u) * 4);
      }
      const uint32_t o = (uint32_t)(bt * MMA_U) * 128u;
#pragma unroll
      for (int u = 0; u < MMA_U; ++u) mma_block32(c, xa + o + u * 128, xb + o + u * 128, w[u]);
#pragma unroll
      for (int u = 0; u < MMA_U; ++u) w[u] = wn[u];
    }
    for (int b = nfull * MMA_U; b < nblk; ++b) {   // K tail (K % 128 != 0)
      const uint4 wt = __ldg(wr + b * 4);
      mma_block32(c, xa + b * 128, xb + b * 128, wt);
    }
    // accumulator fragment: c0,c1 = frame g, outputs 2q, 2q+1 of this group; c2,c3 = frame g+8
    const int o = grp * 8 + 2 * q;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (o + j < out_dim) {
        const float bv = __ldg(bg + o + j);
        const float v0 = c[j] + bv, v1 = c[2 + j] + bv;
        y[g * y_ld + o + j] = relu ? fmaxf(v0, 0.f) : v0;
        y[(g + 8) * y_ld + o + j] = relu ? fmaxf(v1, 0.f) : v1;
      }
    }
  }
  __syncthreads();
}

template <int FT = GATE_FT, bool CL = false, bool TC = false>
__device__ __forceinline__ void linear_ft(const float* __restrict__ Wg, const float* __restrict__ bg,
                                          const float* x, int x_ld, int in_dim, float* y, int y_ld,
                                          int out_dim, bool relu, const __nv_bfloat16* Wb = nullptr) {
  constexpr int GATE_FT = FT;   // (shadows the single-CTA constant inside this function)
  if (TC) {
    static_assert(!TC || (FT == MMA_FT && !CL), "tensor-core variant: 16 frames per CTA, no cluster");
    if (Wb != nullptr && (in_dim & 31) == 0 && out_dim >= 8 && (x_ld & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0) {   // CTA-uniform
      linear_mma16(Wb, bg, x, x_ld, in_dim, y, y_ld, out_dim, relu);
      return;
    }
  }
  constexpr int BATCH = TC ? 2 : MLP_BATCH;   // the tensor-core kernels only send their few tiny layers here
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int sub = lane / MLP_LPR, l = lane % MLP_LPR;
  const bool vec = (in_dim & 3) == 0 && (x_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
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
      for (int j0 = l; j0 < n4; j0 += MLP_LPR * BATCH) {
        float4 w4[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          const int j = j0 + b * MLP_LPR;
          w4[b] = (row_ok && j < n4) ? __ldg(wr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
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

