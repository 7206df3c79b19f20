// Generic NHWC implicit-GEMM convolution on the CUDA cores (fp32 accumulate).
// This is the fp32-mode path (1e-4 parity, no TF32) and the bf16 path for shapes
// the tcgen05 kernel does not take (Cin not a multiple of 32).  Also: max-pool.
//
// GEMM view: D[M = G*B*Ho*Wo, N = Cout] = A[M, K = KH*KW*Cin] * W[N, K]^T
#include "common.cuh"

struct ConvSimtParams {
  const void* x;
  const void* w;
  const float* scale;
  const float* bias;
  const void* residual;
  void* y;
  int B, H, W, Cin, Cout, KH, KW, sh, sw, ph, pw, Ho, Wo, relu, x_shared;
};

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename T>
__global__ __launch_bounds__(256) void conv2d_simt_kernel(ConvSimtParams p) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int g = blockIdx.z;
  const int Ktot = p.KH * p.KW * p.Cin;
  const int Mg = p.B * p.Ho * p.Wo;  // rows per group
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
  const T* x = (const T*)p.x + (p.x_shared ? 0 : (int64_t)g * p.B * p.H * p.W * p.Cin);
  const T* w = (const T*)p.w + (int64_t)g * p.Cout * Ktot;
  const float* scale = p.scale + (int64_t)g * p.Cout;
  const float* bias = p.bias + (int64_t)g * p.Cout;
  T* y = (T*)p.y + (int64_t)g * Mg * p.Cout;
  const T* res = p.residual ? (const T*)p.residual + (int64_t)g * Mg * p.Cout : nullptr;

  const int tid = threadIdx.x;
  const int lrow = tid >> 2;        // 0..63 : tile row (pixel for A, cout for B)
  const int lk = (tid & 3) << 2;    // 0,4,8,12
  // decode this thread's A pixel once
  const int m = m0 + lrow;
  const bool m_ok = m < Mg;
  int n_img = 0, oh = 0, ow = 0;
  if (m_ok) {
    n_img = m / (p.Ho * p.Wo);
    int r = m - n_img * p.Ho * p.Wo;
    oh = r / p.Wo;
    ow = r - oh * p.Wo;
  }
  const int ih0 = oh * p.sh - p.ph, iw0 = ow * p.sw - p.pw;
  const int co = n0 + lrow;
  const bool co_ok = co < p.Cout;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // k -> (kh, kw, channel) once per CTA instead of two integer divisions per gathered element (small-Cin first layers:
  // the index arithmetic of the gather was as long as the FMAs of its tile); longer K falls back to the divisions
  constexpr int KTAB = 1024;
  __shared__ int s_kinfo[KTAB];
  const bool use_tab = Ktot <= KTAB;
  if (use_tab) {
    for (int kk = tid; kk < Ktot; kk += 256) {
      const int tap = kk / p.Cin, c = kk - tap * p.Cin;
      const int kh = tap / p.KW, kw = tap - kh * p.KW;
      s_kinfo[kk] = (kh << 20) | (kw << 12) | c;
    }
    __syncthreads();
  }

  // Software pipeline: the operands of k-block i+1 are loaded into registers before the FMAs of block i, so the gather latency
  // overlaps the arithmetic instead of sitting between the two barriers.  Same elements, same accumulation order.
  float a[4], b[4];
  auto fetch = [&](int kk0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = kk0 + lk + j;
      float av = 0.f, bv = 0.f;
      if (kk < Ktot) {
        int c, kh, kw;
        if (use_tab) {
          const int info = s_kinfo[kk];
          kh = info >> 20; kw = (info >> 12) & 255; c = info & 4095;
        } else {
          const int tap = kk / p.Cin;
          c = kk - tap * p.Cin;
          kh = tap / p.KW;
          kw = tap - kh * p.KW;
        }
        int ih = ih0 + kh, iw = iw0 + kw;
        if (m_ok && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
          av = ld_as_float<T>(x + (((int64_t)n_img * p.H + ih) * p.W + iw) * p.Cin + c);
        if (co_ok) bv = ld_as_float<T>(w + (int64_t)co * Ktot + kk);
      }
      a[j] = av;
      b[j] = bv;
    }
  };
  fetch(0);
  for (int kk0 = 0; kk0 < Ktot; kk0 += SBK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = a[j];
      Bs[lk + j][lrow] = b[j];
    }
    __syncthreads();
    if (kk0 + SBK < Ktot) fetch(kk0 + SBK);
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int mm = m0 + ty * 4 + i;
    if (mm >= Mg) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int nn = n0 + tx * 4 + j;
      if (nn >= p.Cout) continue;
      float v = fmaf(acc[i][j], scale[nn], bias[nn]);
      int64_t o = (int64_t)mm * p.Cout + nn;
      if (res) v += ld_as_float<T>(res + o);
      if (p.relu) v = fmaxf(v, 0.f);
      st_from_float<T>(y + o, v);
    }
  }
}

int amoe_conv2d_simt(amoe_ctx* ctx, const void* x, const void* w, const float* scale,
                     const float* bias, const void* residual, void* y, int G, int x_shared, int B,
                     int H, int W, int Cin, int Cout, int KH, int KW, int sh, int sw, int ph, int pw,
                     int Ho, int Wo, int relu, int dtype, cudaStream_t st) {
  AMOE_ENTER(ctx);
  ConvSimtParams p;
  p.x = x; p.w = w; p.scale = scale; p.bias = bias; p.residual = residual; p.y = y;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW;
  p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw; p.relu = relu; p.x_shared = x_shared;
  p.Ho = Ho; p.Wo = Wo;
  AMOE_REQUIRE(p.Ho > 0 && p.Wo > 0, "conv2d: empty output (%dx%d)", p.Ho, p.Wo);
  int64_t Mg = (int64_t)B * p.Ho * p.Wo;
  AMOE_REQUIRE(Mg < (1ll << 31) - SBM, "conv2d: too many output pixels per group");
  if (Mg == 0) return 0;
  dim3 grid((unsigned)((Mg + SBM - 1) / SBM), (unsigned)ceil_div(Cout, SBN), (unsigned)G);
  if (dtype == AMOE_F32)
    conv2d_simt_kernel<float><<<grid, 256, 0, st>>>(p);
  else
    conv2d_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

// ---- MaxPool2d(3, stride 2, pad 1), NHWC; VEC channels per thread.  out_pad > 0 writes into a
// physically padded output [NB][Ho+2p][Wo+2p][C] and fills the border with zeros. ----
template <typename T, int VEC>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C,
                                    int Ho, int Wo, int out_pad, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Hop = Ho + 2 * out_pad, Wop = Wo + 2 * out_pad;
  int cv = C / VEC;
  int c = (int)(i % cv) * VEC;
  int64_t r = i / cv;
  int owp = (int)(r % Wop);
  r /= Wop;
  int ohp = (int)(r % Hop);
  int64_t n = r / Hop;
  const int oh = ohp - out_pad, ow = owp - out_pad;
  const bool border = oh < 0 || oh >= Ho || ow < 0 || ow >= Wo;
  float m[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) m[v] = border ? 0.f : -INFINITY;
  if (!border) {
    for (int kh = 0; kh < 3; ++kh) {
      int ih = oh * 2 - 1 + kh;
      if (ih < 0 || ih >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        int iw = ow * 2 - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        const T* s = x + ((n * H + ih) * W + iw) * C + c;
        if constexpr (sizeof(T) * VEC == 16) {
          uint4 raw = *reinterpret_cast<const uint4*>(s);
          const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
          for (int v = 0; v < VEC; ++v) m[v] = fmaxf(m[v], ld_as_float<T>(e + v));
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) m[v] = fmaxf(m[v], ld_as_float<T>(s + v));
        }
      }
    }
  }
  T* d = y + ((n * Hop + ohp) * Wop + owp) * C + c;
  if constexpr (sizeof(T) * VEC == 16) {
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int v = 0; v < VEC; ++v) st_from_float<T>(e + v, m[v]);
    *reinterpret_cast<uint4*>(d) = raw;
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v) st_from_float<T>(d + v, m[v]);
  }
}

extern "C" int amoe_maxpool3x3s2_fwd(amoe_ctx* ctx, const void* x, void* y, int NB, int H, int W,
                                      int C, int dtype, int out_pad, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && y, "amoe_maxpool3x3s2_fwd: NULL argument");
  AMOE_REQUIRE(out_pad >= 0, "amoe_maxpool3x3s2_fwd: negative out_pad");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (NB == 0) return 0;
  const int64_t pos = (int64_t)NB * (Ho + 2 * out_pad) * (Wo + 2 * out_pad);
  if (dtype == AMOE_BF16 && C % 8 == 0) {
    int64_t total = pos * (C / 8);
    maxpool3x3s2_kernel<__nv_bfloat16, 8><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, H, W, C, Ho, Wo, out_pad, total);
  } else if (dtype == AMOE_BF16) {
    int64_t total = pos * C;
    maxpool3x3s2_kernel<__nv_bfloat16, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, H, W, C, Ho, Wo, out_pad, total);
  } else if (dtype == AMOE_F32 && C % 4 == 0) {
    int64_t total = pos * (C / 4);
    maxpool3x3s2_kernel<float, 4><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const float*)x, (float*)y, H, W, C, Ho, Wo, out_pad, total);
  } else if (dtype == AMOE_F32) {
    int64_t total = pos * C;
    maxpool3x3s2_kernel<float, 1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        (const float*)x, (float*)y, H, W, C, Ho, Wo, out_pad, total);
  } else {
    AMOE_REQUIRE(false, "amoe_maxpool3x3s2_fwd: bad dtype %d", dtype);
  }
  AMOE_LAUNCH_OK(ctx);
  return 0;
}
