// Training-step kernels of the policy backbone (EasyBackbone, models/policy/trajectory_head.py:5-33):
// train-mode BatchNorm2d (batch statistics, running-stat update) forward/backward, convolution
// backward (data and weight gradients) and global-average-pool forward/backward.  fp32, NHWC, CUDA
// cores: four small stride-2 convolutions at a per-GPU batch of 32; the forward
// convolution itself is amoe_conv2d_fwd(dtype = f32).
//
// All reductions are two-stage with a fixed summation order, so a training step is bit-reproducible.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int RED_THREADS = 256;
constexpr int RED_ROWS_MAX = 1024;
// rows of [M, C] per CTA of the column reductions: enough CTAs to cover the GPU a few times (a fixed 1024 rows left the
// deep layers - a few thousand rows of 512 channels - to two CTAs: 84 us per BatchNorm statistics pass on average)
static inline int red_rows(int64_t M) {
  int64_t r = (M + 591) / 592;
  if (r < 16) r = 16;
  if (r > RED_ROWS_MAX) r = RED_ROWS_MAX;
  return (int)r;
}

// Column sums of up to two per-element quantities over the rows of an [M, C] fp32 matrix.
//   MODE 0: s0 = sum x,           s1 = sum x*x                       (BatchNorm statistics)
//   MODE 1: s0 = sum g,           s1 = sum g * (x - mean) * rstd      (BatchNorm backward)
//           with g = dy * [y > 0] when y != nullptr (ReLU fused behind the norm) else dy
//   MODE 2: s0 = sum x                                                (bias gradient / pooling)
// partial: [gridDim.x][2][C]
template <int MODE>
__global__ __launch_bounds__(RED_THREADS) void colreduce_partial_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ dy,
                                                                        const float* __restrict__ y,
                                                                        const float* __restrict__ mean,
                                                                        const float* __restrict__ rstd, int64_t M, int C,
                                                                        double* __restrict__ partial, int rows_per_cta) {
  __shared__ double sh0[RED_THREADS], sh1[RED_THREADS];
  const int cb = min(C, RED_THREADS);          // channels handled per pass (C is a multiple of cb or < 256)
  const int rl = threadIdx.x / cb, nrl = RED_THREADS / cb;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  // grouped launches (MODE 0 only, amoe_bn_train_fwd_grouped): blockIdx.y = group, x = [G][M][C], partial = [G][gridDim.x][2][C]
  x += (int64_t)blockIdx.y * M * C;
  partial += (int64_t)blockIdx.y * gridDim.x * 2 * C;
  for (int c0 = 0; c0 < C; c0 += cb) {
    const int c = c0 + threadIdx.x % cb;
    // double accumulators: these sums feed differences (variance; g - mean(g) in the BatchNorm backward,
    // where the pooled policy gradient is almost constant over a channel), so fp32 rounding would be
    // amplified by the cancellation.  The reductions are memory-bound; the fp64 adds are free.
    double s0 = 0.0, s1 = 0.0;
    if (c < C && rl < nrl) {
      const float mu = MODE == 1 ? mean[c] : 0.f, rs = MODE == 1 ? rstd[c] : 0.f;
      for (int64_t r = r0 + rl; r < r1; r += nrl) {
        const int64_t i = r * C + c;
        if (MODE == 0) {
          const double v = (double)x[i];
          s0 += v;
          s1 += v * v;
        } else if (MODE == 1) {
          float g = dy[i];
          if (y && !(y[i] > 0.f)) g = 0.f;
          s0 += (double)g;
          s1 += (double)g * (double)((x[i] - mu) * rs);
        } else {
          s0 += (double)x[i];
        }
      }
    }
    sh0[threadIdx.x] = s0;
    sh1[threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.x < cb && c < C) {
      double t0 = 0.0, t1 = 0.0;
      for (int j = 0; j < nrl; ++j) {      // fixed order
        t0 += sh0[j * cb + threadIdx.x];
        t1 += sh1[j * cb + threadIdx.x];
      }
      partial[((int64_t)blockIdx.x * 2 + 0) * C + c] = t0;
      partial[((int64_t)blockIdx.x * 2 + 1) * C + c] = t1;
    }
    __syncthreads();
  }
}

// BatchNorm statistics from the partial sums: mean, rstd (biased variance), running-stat update
// (momentum; running_var gets the unbiased variance, as nn.BatchNorm2d does).
// One WARP per channel: lane l sums the partials l, l+32, ... in order, then a fixed shuffle tree (bit-reproducible); a
// thread per channel walking up to 592 partials serially cost 65 us per call.
__device__ __forceinline__ void warp_sum2(double& a, double& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

__global__ void bn_stats_final_kernel(const double* __restrict__ partial, int nblk, int64_t M, int C, float eps,
                                      float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                      float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, ss = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    s += partial[((int64_t)b * 2 + 0) * C + c];
    ss += partial[((int64_t)b * 2 + 1) * C + c];
  }
  warp_sum2(s, ss);
  if (lane != 0) return;
  const double mu = s / (double)M;
  double var = ss / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}
// out0[c] = sum of partial s0, out1[c] = sum of partial s1 (either may be NULL); scale applied
__global__ void colreduce_final_kernel(const double* __restrict__ partial, int nblk, int C, float scale,
                                       float* __restrict__ out0, float* __restrict__ out1) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, t = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    s += partial[((int64_t)b * 2 + 0) * C + c];
    t += partial[((int64_t)b * 2 + 1) * C + c];
  }
  warp_sum2(s, t);
  if (lane != 0) return;
  if (out0) out0[c] = (float)(s * (double)scale);
  if (out1) out1[c] = (float)(t * (double)scale);
}

// y = act((x - mean[c]) * rstd[c] * gamma[c] + beta[c])
// Four 16-byte elements per thread, all loads issued before the arithmetic (one element per thread left large activations
// at 1.8 TB/s).
constexpr int BN_APPLY_PER = 4;
__global__ __launch_bounds__(256) void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ y, int64_t n4, int C,
                                                       int relu) {
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
  float4* __restrict__ y4 = reinterpret_cast<float4*>(y);
  const int64_t i0 = (int64_t)blockIdx.x * (256 * BN_APPLY_PER) + threadIdx.x;
  float4 v[BN_APPLY_PER];
#pragma unroll
  for (int u = 0; u < BN_APPLY_PER; ++u) {
    const int64_t i = i0 + u * 256;
    if (i < n4) v[u] = x4[i];
  }
#pragma unroll
  for (int u = 0; u < BN_APPLY_PER; ++u) {
    const int64_t i = i0 + u * 256;
    if (i >= n4) continue;
    const int c = (int)((i * 4) % C);
    float o[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = (o[j] - mean[c + j]) * rstd[c + j] * gamma[c + j] + beta[c + j];
      if (relu) o[j] = fmaxf(o[j], 0.f);
    }
    y4[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}
// The same two kernels for G independent BatchNorm layers of one shape in ONE launch each (blockIdx.y = layer): the frozen
// experts of a gating-training step run the same ResNet-18 layer by layer, so their statistics passes are grouped like their
// convolutions.  Same arithmetic in the same order as the single-layer kernels (results are bit-identical).
constexpr int BN_MAX_GROUPS = 4;
struct BnGroupPtrs {
  const float* gamma[BN_MAX_GROUPS];
  const float* beta[BN_MAX_GROUPS];
  float* running_mean[BN_MAX_GROUPS];
  float* running_var[BN_MAX_GROUPS];
};

__global__ void bn_stats_final_grouped_kernel(const double* __restrict__ partial, int nblk, int64_t M, int C, float eps,
                                              float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                              const BnGroupPtrs gp) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, g = blockIdx.y;
  if (c >= C) return;
  partial += (int64_t)g * nblk * 2 * C;
  double s = 0.0, ss = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    s += partial[((int64_t)b * 2 + 0) * C + c];
    ss += partial[((int64_t)b * 2 + 1) * C + c];
  }
  warp_sum2(s, ss);
  if (lane != 0) return;
  const double mu = s / (double)M;
  double var = ss / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[g * C + c] = (float)mu;
  rstd[g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  float* rm = gp.running_mean[g];
  float* rv = gp.running_var[g];
  if (rm) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    rm[c] = (1.f - momentum) * rm[c] + momentum * (float)mu;
    rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unbiased;
  }
}

// (four 16-byte elements per thread as in bn_apply_kernel: 113 -> 57 us on the 100 MB layer1 activations of three experts)
__global__ __launch_bounds__(256) void bn_apply_grouped_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const BnGroupPtrs gp,
                                                               float* __restrict__ y, int64_t n4, int C, int relu) {
  const int g = blockIdx.y;
  const float* __restrict__ gamma = gp.gamma[g];
  const float* __restrict__ beta = gp.beta[g];
  mean += g * C;
  rstd += g * C;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x) + (int64_t)g * n4;
  float4* __restrict__ y4 = reinterpret_cast<float4*>(y) + (int64_t)g * n4;
  const int64_t i0 = (int64_t)blockIdx.x * (256 * BN_APPLY_PER) + threadIdx.x;
  float4 v[BN_APPLY_PER];
#pragma unroll
  for (int u = 0; u < BN_APPLY_PER; ++u) {
    const int64_t i = i0 + u * 256;
    if (i < n4) v[u] = __ldcs(x4 + i);
  }
#pragma unroll
  for (int u = 0; u < BN_APPLY_PER; ++u) {
    const int64_t i = i0 + u * 256;
    if (i >= n4) continue;
    const int c = (int)((i * 4) % C);
    float o[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = (o[j] - mean[c + j]) * rstd[c + j] * gamma[c + j] + beta[c + j];
      if (relu) o[j] = fmaxf(o[j], 0.f);
    }
    y4[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// train: dx = gamma*rstd*(g - sum_g/M - xhat*sum_gx/M);  eval (batch_stats = 0): dx = gamma*rstd*g
// (two 16-byte elements per thread and all loads first, as in bn_apply_kernel: up to six independent loads in flight)
constexpr int BN_BWD_PER = 2;
__global__ __launch_bounds__(256) void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ sum_g, const float* __restrict__ sum_gx,
                                                           float* __restrict__ dx, int64_t n4, int C, float inv_M,
                                                           int batch_stats) {
  const int64_t i0 = (int64_t)blockIdx.x * (256 * BN_BWD_PER) + threadIdx.x;
  float4 d4[BN_BWD_PER], x4[BN_BWD_PER], y4[BN_BWD_PER];
#pragma unroll
  for (int u = 0; u < BN_BWD_PER; ++u) {
    const int64_t i = i0 + u * 256;
    if (i < n4) {
      d4[u] = reinterpret_cast<const float4*>(dy)[i];
      x4[u] = reinterpret_cast<const float4*>(x)[i];
      if (y) y4[u] = reinterpret_cast<const float4*>(y)[i];
    }
  }
#pragma unroll
  for (int u = 0; u < BN_BWD_PER; ++u) {
    const int64_t i = i0 + u * 256;
    if (i >= n4) continue;
    const int c = (int)((i * 4) % C);
    float g[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
    const float xv[4] = {x4[u].x, x4[u].y, x4[u].z, x4[u].w};
    if (y) {
      const float yv[4] = {y4[u].x, y4[u].y, y4[u].z, y4[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (!(yv[j] > 0.f)) g[j] = 0.f;
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float k = gamma[c + j] * rstd[c + j];
      if (batch_stats) {
        const float xh = (xv[j] - mean[c + j]) * rstd[c + j];
        o[j] = k * (g[j] - sum_g[c + j] * inv_M - xh * sum_gx[c + j] * inv_M);
      } else {
        o[j] = k * g[j];
      }
    }
    reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// Convolution backward, NHWC fp32, implicit GEMMs on 64x64x16 tiles (4x4 outputs per thread).
struct ConvBwdParams {
  const float* dy;   // [B,Ho,Wo,Cout]
  const float* x;    // [B,H,W,Cin]            (weight gradient)
  const float* w;    // [Cout][KH][KW][Cin]     (data gradient)
  float* out;        // dx [B,H,W,Cin]  or  partial dW [S][Cout][KH*KW*Cin]
  int B, H, W, Cin, Cout, KH, KW, sh, sw, ph, pw, Ho, Wo;
  int rows_per_slice;  // weight gradient: output pixels per K slice
};
constexpr int TBM = 64, TBN = 64, TBK = 16;

// dx[m=(n,ih,iw), ci] = sum_{kh,kw,co} dy[n, (ih+ph-kh)/sh, (iw+pw-kw)/sw, co] * w[co,kh,kw,ci]
__global__ __launch_bounds__(256) void conv_bwd_data_kernel(ConvBwdParams p) {
  __shared__ __align__(16) float As[TBK][TBM + 4];
  __shared__ __align__(16) float Bs[TBK][TBN + 4];
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int Ktot = p.KH * p.KW * p.Cout;   // reduction index k = tap*Cout + co
  const int Kw = p.KH * p.KW * p.Cin;      // row length of the packed weights
  const int64_t m0 = (int64_t)blockIdx.x * TBM;
  const int n0 = blockIdx.y * TBN;
  const int tid = threadIdx.x, lrow = tid >> 2, lk = (tid & 3) << 2, ty = tid >> 4, tx = tid & 15;
  const int64_t m = m0 + lrow;
  const bool m_ok = m < M;
  int n_img = 0, ih = 0, iw = 0;
  if (m_ok) {
    n_img = (int)(m / ((int64_t)p.H * p.W));
    const int r = (int)(m - (int64_t)n_img * p.H * p.W);
    ih = r / p.W;
    iw = r - ih * p.W;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // k -> (kh, kw, co) once per CTA (two integer divisions per gathered element otherwise), strides 1 / 2 as shifts
  constexpr int KTAB = 4608;              // 3x3 x 512 channels
  __shared__ int s_kinfo[KTAB];
  const bool use_tab = Ktot <= KTAB && p.Cout <= 4096;
  if (use_tab) {
    for (int k = tid; k < Ktot; k += 256) {
      const int tap = k / p.Cout, co = k - tap * p.Cout;
      const int kh = tap / p.KW, kw = tap - kh * p.KW;
      s_kinfo[k] = (kh << 20) | (kw << 12) | co;
    }
    __syncthreads();
  }
  const bool pow2 = (p.sh == 1 || p.sh == 2) && (p.sw == 1 || p.sw == 2);
  const int shh = p.sh - 1, sww = p.sw - 1;   // shift amounts / parity masks when pow2
  // Software pipeline (as in conv_bwd_weight_kernel): operands of k-block i+1 go to registers before the FMAs of block i.
  constexpr int B_PER = TBK * TBN / 256;
  float a_reg[4], b_reg[B_PER];
  auto fetch = [&](int k0) {
    // A: this thread's pixel, 4 consecutive k (one 16-byte load when Cout % 4 == 0: same tap, aligned)
    if ((p.Cout & 3) == 0) {
      const int k = k0 + lk;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m_ok && k < Ktot) {
        int co, kh, kw;
        if (use_tab) { const int info = s_kinfo[k]; kh = info >> 20; kw = (info >> 12) & 255; co = info & 4095; }
        else { const int tap = k / p.Cout; co = k - tap * p.Cout; kh = tap / p.KW; kw = tap - kh * p.KW; }
        const int th = ih + p.ph - kh, tw = iw + p.pw - kw;
        if (th >= 0 && tw >= 0 && (pow2 ? ((th & shh) | (tw & sww)) == 0 : (th % p.sh == 0 && tw % p.sw == 0))) {
          const int oh = pow2 ? th >> shh : th / p.sh, ow = pow2 ? tw >> sww : tw / p.sw;
          if (oh < p.Ho && ow < p.Wo)
            a = *reinterpret_cast<const float4*>(p.dy + (((int64_t)n_img * p.Ho + oh) * p.Wo + ow) * p.Cout + co);
        }
      }
      a_reg[0] = a.x; a_reg[1] = a.y; a_reg[2] = a.z; a_reg[3] = a.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lk + j;
        float a = 0.f;
        if (m_ok && k < Ktot) {
          const int tap = k / p.Cout, co = k - tap * p.Cout;
          const int kh = tap / p.KW, kw = tap - kh * p.KW;
          const int th = ih + p.ph - kh, tw = iw + p.pw - kw;
          if (th >= 0 && tw >= 0 && th % p.sh == 0 && tw % p.sw == 0) {
            const int oh = th / p.sh, ow = tw / p.sw;
            if (oh < p.Ho && ow < p.Wo) a = p.dy[(((int64_t)n_img * p.Ho + oh) * p.Wo + ow) * p.Cout + co];
          }
        }
        a_reg[j] = a;
      }
    }
    // B: 16 k x 64 ci; element (k, ci) = w[co][tap][ci]; consecutive threads walk ci
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * 256;
      const int kk = e / TBN, c = e - kk * TBN;
      const int k = k0 + kk, ci = n0 + c;
      float v = 0.f;
      if (k < Ktot && ci < p.Cin) {
        int tap, co;
        if (use_tab) { const int info = s_kinfo[k]; co = info & 4095; tap = (info >> 20) * p.KW + ((info >> 12) & 255); }
        else { tap = k / p.Cout; co = k - tap * p.Cout; }
        v = p.w[(int64_t)co * Kw + tap * p.Cin + ci];
      }
      b_reg[i] = v;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < Ktot; k0 += TBK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[lk + j][lrow] = a_reg[j];
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * 256;
      Bs[e / TBN][e % TBN] = b_reg[i];
    }
    __syncthreads();
    if (k0 + TBK < Ktot) fetch(k0 + TBK);
#pragma unroll
    for (int k = 0; k < TBK; ++k) {
      // one 16-byte shared-memory load per operand (rows are 16-byte aligned: pitch TBM + 4 floats)
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = n0 + tx * 4 + j;
      if (ci < p.Cin) p.out[mm * p.Cin + ci] = acc[i][j];
    }
  }
}

// partial dW[s][co][k=(kh,kw,ci)] = sum over the output pixels of slice s of dy[pix,co] * x[window(pix,k)]
__global__ __launch_bounds__(256) void conv_bwd_weight_kernel(ConvBwdParams p) {
  __shared__ __align__(16) float As[TBK][TBM + 4];   // [pixel][co]
  __shared__ __align__(16) float Bs[TBK][TBN + 4];   // [pixel][k]
  const int Kw = p.KH * p.KW * p.Cin;
  const int64_t Mg = (int64_t)p.B * p.Ho * p.Wo;
  const int co0 = blockIdx.x * TBM, k0 = blockIdx.y * TBN, s = blockIdx.z;
  const int64_t r0 = (int64_t)s * p.rows_per_slice, r1 = min(Mg, r0 + p.rows_per_slice);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // The gather of the x window used four integer divisions per loaded element (pixel -> image/row/column, k -> tap/channel);
  // a thread always loads the same k column (c = tid % TBN), so its tap decomposition is hoisted, and the 16 pixels of an
  // iteration are decomposed once, by 16 threads, into shared memory.  Same elements, same order: the sums are unchanged.
  static_assert(256 % TBN == 0 && TBK % (256 / TBN) == 0, "gather mapping: fixed k column per thread");
  __shared__ int s_img[2][TBK], s_ih0[2][TBK], s_iw0[2][TBK];
  const int c_fix = tid % TBN, rr0 = tid / TBN;
  const int kfix = k0 + c_fix;
  const bool k_ok = kfix < Kw;
  const int tap_f = k_ok ? kfix / p.Cin : 0, ci_f = k_ok ? kfix - tap_f * p.Cin : 0;
  const int kh_f = tap_f / p.KW, kw_f = tap_f - kh_f * p.KW;
  const int HoWo = p.Ho * p.Wo;
  constexpr int A_PER = TBK * TBM / 256, B_PER = TBK / (256 / TBN);
  // Software pipeline: the global loads of iteration i+1 (dy tile, gathered x window) are issued into registers before the
  // FMAs of iteration i, so their latency overlaps the arithmetic instead of sitting between two barriers (284-505 us per
  // policy-backbone weight gradient before).  Same elements, same order of accumulation: results unchanged.
  auto decompose = [&](int buf, int64_t rb) {          // threads 0..TBK-1: pixel -> (image, top-left input row / column)
    const int64_t r = rb + tid;
    int n_img = -1, ih0 = 0, iw0 = 0;
    if (r < r1) {
      n_img = (int)(r / HoWo);
      const int rem = (int)(r - (int64_t)n_img * HoWo);
      const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
      ih0 = oh * p.sh - p.ph; iw0 = ow * p.sw - p.pw;
    }
    s_img[buf][tid] = n_img; s_ih0[buf][tid] = ih0; s_iw0[buf][tid] = iw0;
  };
  float a_reg[A_PER], b_reg[B_PER];
  auto fetch = [&](int buf, int64_t rb) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * 256;
      const int rr = e / TBM, c = e - rr * TBM;
      const int64_t r = rb + rr;
      const int co = co0 + c;
      a_reg[i] = (r < r1 && co < p.Cout) ? __ldg(p.dy + r * p.Cout + co) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int rr = rr0 + i * (256 / TBN);
      const int n_img = s_img[buf][rr];
      float v = 0.f;
      if (n_img >= 0 && k_ok) {
        const int ih = s_ih0[buf][rr] + kh_f, iw = s_iw0[buf][rr] + kw_f;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = __ldg(p.x + (((int64_t)n_img * p.H + ih) * p.W + iw) * p.Cin + ci_f);
      }
      b_reg[i] = v;
    }
  };
  if (tid < TBK) decompose(0, r0);
  __syncthreads();
  if (r0 < r1) fetch(0, r0);
  int cur = 0;
  for (int64_t rb = r0; rb < r1; rb += TBK, cur ^= 1) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * 256;
      const int rr = e / TBM, c = e - rr * TBM;
      As[rr][c] = a_reg[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) Bs[rr0 + i * (256 / TBN)][c_fix] = b_reg[i];
    if (tid < TBK) decompose(cur ^ 1, rb + TBK);
    __syncthreads();
    if (rb + TBK < r1) fetch(cur ^ 1, rb + TBK);
#pragma unroll
    for (int k = 0; k < TBK; ++k) {
      // one 16-byte shared-memory load per operand (rows are 16-byte aligned: pitch TBM + 4 floats)
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = p.out + (int64_t)s * p.Cout * Kw;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < Kw) out[(int64_t)co * Kw + k] = acc[i][j];
    }
  }
}
__global__ void slice_sum_kernel(const float* __restrict__ partial, int S, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;   // a few hundred partials of mixed sign: the final sum in double costs nothing
  for (int k = 0; k < S; ++k) s += (double)partial[(int64_t)k * n + i];
  out[i] = (float)s;
}

// global average pool over HW: x [B,HW,C] -> out [B,C]; backward broadcasts dy/HW
__global__ void gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int HW, int C) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < HW; ++i) s += x[((int64_t)b * HW + i) * C + c];
    out[(int64_t)b * C + c] = s / (float)HW;
  }
}
__global__ void gap_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t n, int HW, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const int64_t b = i / ((int64_t)HW * C);
  dx[i] = dy[b * C + c] / (float)HW;
}

inline double* ws64(float* ws) { return reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 7) & ~(uintptr_t)7); }

int wgrad_slices(int sm_count, int Cout, int Kw, int64_t Mg, int* rows_per_slice) {
  const int tiles = ceil_div(Cout, TBM) * ceil_div(Kw, TBN);
  int S = std::max(1, (4 * sm_count) / tiles);
  const int64_t max_s = std::max<int64_t>(1, Mg / 256);   // at least 256 pixels per slice
  S = (int)std::min<int64_t>(S, max_s);
  int rps = (int)((Mg + S - 1) / S);
  rps = (rps + TBK - 1) / TBK * TBK;
  S = (int)((Mg + rps - 1) / rps);
  *rows_per_slice = rps;
  return S;
}

}  // namespace

extern "C" {

// partial sums are doubles: 2 floats each, +2 so the caller's float buffer can be aligned up to 8 bytes
int64_t amoe_colreduce_workspace_floats(int64_t M, int C) { const int rr = red_rows(M); return ((M + rr - 1) / rr) * 4 * (int64_t)C + 2; }

int amoe_bn_train_fwd(amoe_ctx* ctx, const float* x, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, float* y, float* save_mean, float* save_rstd,
                      float* workspace, int64_t M, int C, int relu, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && gamma && beta && y && save_mean && save_rstd && workspace, "amoe_bn_train_fwd: NULL argument");
  AMOE_REQUIRE(C % 4 == 0 && C >= 4, "amoe_bn_train_fwd: C=%d must be a multiple of 4", C);
  AMOE_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "amoe_bn_train_fwd: running stats come together");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int rr = red_rows(M), nblk = (int)((M + rr - 1) / rr);
  colreduce_partial_kernel<0><<<nblk, RED_THREADS, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, M, C, ws64(workspace), rr);
  AMOE_LAUNCH_OK(ctx);
  bn_stats_final_kernel<<<ceil_div(C, 4), 128, 0, st>>>(ws64(workspace), nblk, M, C, eps, momentum, save_mean, save_rstd,
                                                        running_mean, running_var);
  AMOE_LAUNCH_OK(ctx);
  const int64_t n4 = M * C / 4;
  bn_apply_kernel<<<(unsigned)((n4 + 256 * BN_APPLY_PER - 1) / (256 * BN_APPLY_PER)), 256, 0, st>>>(x, save_mean, save_rstd, gamma, beta, y, n4, C, relu);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_bn_train_fwd_grouped(amoe_ctx* ctx, const float* x, const float* const* gamma, const float* const* beta,
                              float* const* running_mean, float* const* running_var, float momentum, float eps, float* y,
                              float* save_mean, float* save_rstd, float* workspace, int G, int64_t M, int C, int relu,
                              void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && gamma && beta && y && save_mean && save_rstd && workspace, "amoe_bn_train_fwd_grouped: NULL argument");
  AMOE_REQUIRE(G >= 1 && G <= BN_MAX_GROUPS, "amoe_bn_train_fwd_grouped: G=%d out of [1,%d]", G, BN_MAX_GROUPS);
  AMOE_REQUIRE(C % 4 == 0 && C >= 4, "amoe_bn_train_fwd_grouped: C=%d must be a multiple of 4", C);
  AMOE_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "amoe_bn_train_fwd_grouped: running stats come together");
  BnGroupPtrs gp{};
  for (int g = 0; g < G; ++g) {
    AMOE_REQUIRE(gamma[g] && beta[g], "amoe_bn_train_fwd_grouped: NULL gamma/beta of group %d", g);
    gp.gamma[g] = gamma[g];
    gp.beta[g] = beta[g];
    gp.running_mean[g] = running_mean ? running_mean[g] : nullptr;
    gp.running_var[g] = running_var ? running_var[g] : nullptr;
    AMOE_REQUIRE((gp.running_mean[g] == nullptr) == (gp.running_var[g] == nullptr), "amoe_bn_train_fwd_grouped: running stats come together");
  }
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int rr = red_rows(M), nblk = (int)((M + rr - 1) / rr);
  colreduce_partial_kernel<0><<<dim3(nblk, G), RED_THREADS, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, M, C, ws64(workspace), rr);
  AMOE_LAUNCH_OK(ctx);
  bn_stats_final_grouped_kernel<<<dim3(ceil_div(C, 4), G), 128, 0, st>>>(ws64(workspace), nblk, M, C, eps, momentum, save_mean,
                                                                        save_rstd, gp);
  AMOE_LAUNCH_OK(ctx);
  const int64_t n4 = M * C / 4;
  bn_apply_grouped_kernel<<<dim3((unsigned)((n4 + 256 * BN_APPLY_PER - 1) / (256 * BN_APPLY_PER)), G), 256, 0, st>>>(x, save_mean, save_rstd, gp, y, n4, C, relu);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_bn_apply_fwd(amoe_ctx* ctx, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, float* y, int64_t M, int C, int relu, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && mean && rstd && gamma && beta && y, "amoe_bn_apply_fwd: NULL argument");
  AMOE_REQUIRE(C % 4 == 0, "amoe_bn_apply_fwd: C=%d must be a multiple of 4", C);
  if (M == 0) return 0;
  const int64_t n4 = M * C / 4;
  bn_apply_kernel<<<(unsigned)((n4 + 256 * BN_APPLY_PER - 1) / (256 * BN_APPLY_PER)), 256, 0, (cudaStream_t)stream>>>(x, mean, rstd, gamma, beta, y, n4, C, relu);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_bn_bwd(amoe_ctx* ctx, const float* dy, const float* x, const float* y_relu, const float* gamma, const float* mean,
                const float* rstd, float* dx, float* dgamma, float* dbeta, float* workspace, int64_t M, int C,
                int batch_stats, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && x && gamma && mean && rstd && dgamma && dbeta && workspace, "amoe_bn_bwd: NULL argument");
  AMOE_REQUIRE(C % 4 == 0 && C >= 4, "amoe_bn_bwd: C=%d must be a multiple of 4", C);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int rr = red_rows(M), nblk = (int)((M + rr - 1) / rr);
  colreduce_partial_kernel<1><<<nblk, RED_THREADS, 0, st>>>(x, dy, y_relu, mean, rstd, M, C, ws64(workspace), rr);
  AMOE_LAUNCH_OK(ctx);
  colreduce_final_kernel<<<ceil_div(C, 4), 128, 0, st>>>(ws64(workspace), nblk, C, 1.f, dbeta, dgamma);
  AMOE_LAUNCH_OK(ctx);
  if (dx) {
    const int64_t n4 = M * C / 4;
    bn_bwd_apply_kernel<<<(unsigned)((n4 + 256 * BN_BWD_PER - 1) / (256 * BN_BWD_PER)), 256, 0, st>>>(dy, x, y_relu, mean, rstd, gamma, dbeta, dgamma, dx, n4,
                                                                    C, 1.f / (float)M, batch_stats);
    AMOE_LAUNCH_OK(ctx);
  }
  return 0;
}

int amoe_colsum(amoe_ctx* ctx, const float* x, float* out, float* workspace, int64_t M, int C, float scale, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && out && workspace, "amoe_colsum: NULL argument");
  AMOE_REQUIRE(C >= 1, "amoe_colsum: C=%d", C);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int rr = red_rows(M), nblk = (int)((M + rr - 1) / rr);
  colreduce_partial_kernel<2><<<nblk, RED_THREADS, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, M, C, ws64(workspace), rr);
  AMOE_LAUNCH_OK(ctx);
  colreduce_final_kernel<<<ceil_div(C, 4), 128, 0, st>>>(ws64(workspace), nblk, C, scale, out, nullptr);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_conv2d_bwd_data(amoe_ctx* ctx, const float* dy, const float* w, float* dx, int B, int H, int W, int Cin, int Cout,
                         int KH, int KW, int stride_h, int stride_w, int pad_h, int pad_w, int Ho, int Wo, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && w && dx, "amoe_conv2d_bwd_data: NULL argument");
  if (B == 0) return 0;
  ConvBwdParams p;
  p.dy = dy; p.x = nullptr; p.w = w; p.out = dx;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.sh = stride_h; p.sw = stride_w;
  p.ph = pad_h; p.pw = pad_w; p.Ho = Ho; p.Wo = Wo; p.rows_per_slice = 0;
  const int64_t M = (int64_t)B * H * W;
  dim3 grid((unsigned)((M + TBM - 1) / TBM), ceil_div(Cin, TBN));
  conv_bwd_data_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int64_t amoe_conv2d_bwd_weight_workspace_floats(amoe_ctx* ctx, int B, int Cin, int Cout, int KH, int KW, int Ho, int Wo) {
  AMOE_ENTER(ctx);
  if (!ctx) return -1;
  int rps;
  const int S = wgrad_slices(ctx->sm_count, Cout, KH * KW * Cin, (int64_t)B * Ho * Wo, &rps);
  return (int64_t)S * Cout * KH * KW * Cin;
}

int amoe_conv2d_bwd_weight(amoe_ctx* ctx, const float* dy, const float* x, float* dw, float* workspace,
                           int64_t workspace_floats, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride_h,
                           int stride_w, int pad_h, int pad_w, int Ho, int Wo, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && x && dw && workspace, "amoe_conv2d_bwd_weight: NULL argument");
  if (B == 0) return 0;
  ConvBwdParams p;
  p.dy = dy; p.x = x; p.w = nullptr; p.out = workspace;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.sh = stride_h; p.sw = stride_w;
  p.ph = pad_h; p.pw = pad_w; p.Ho = Ho; p.Wo = Wo;
  const int Kw = KH * KW * Cin;
  const int S = wgrad_slices(ctx->sm_count, Cout, Kw, (int64_t)B * Ho * Wo, &p.rows_per_slice);
  AMOE_REQUIRE(workspace_floats >= (int64_t)S * Cout * Kw, "amoe_conv2d_bwd_weight: workspace holds %lld floats, %lld needed",
               (long long)workspace_floats, (long long)S * Cout * Kw);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(Cout, TBM), ceil_div(Kw, TBN), S);
  conv_bwd_weight_kernel<<<grid, 256, 0, st>>>(p);
  AMOE_LAUNCH_OK(ctx);
  const int64_t n = (int64_t)Cout * Kw;
  slice_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, S, n, dw);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_gap_fwd(amoe_ctx* ctx, const float* x, float* out, int B, int HW, int C, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && out, "amoe_gap_fwd: NULL argument");
  if (B == 0) return 0;
  gap_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, out, HW, C);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_gap_bwd(amoe_ctx* ctx, const float* dy, float* dx, int B, int HW, int C, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && dx, "amoe_gap_bwd: NULL argument");
  if (B == 0) return 0;
  const int64_t n = (int64_t)B * HW * C;
  gap_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, n, HW, C);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
