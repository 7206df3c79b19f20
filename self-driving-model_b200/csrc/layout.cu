// Layout conversion and weight packing: NCHW fp32 image -> NHWC, OIHW -> packed
// [Cout][KH][KW][Cin_pad], eval-mode BatchNorm folding.
#include "common.cuh"

// one thread per pixel: reads C planes (coalesced across the warp), writes Cp
// contiguous channels (8 B for bf16 Cp=4, 16 B for Cp=8 / f32 Cp=4).
template <typename T>
__global__ void image_nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst,
                                          int C, int HW, int Cp, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t b = i / HW;
  int p = (int)(i - b * HW);
  const float* s = src + b * (int64_t)C * HW + p;
  T* d = dst + i * Cp;
  for (int c = 0; c < Cp; ++c) {
    float v = (c < C) ? __ldg(s + (int64_t)c * HW) : 0.f;
    st_from_float<T>(d + c, v);
  }
}

// specialised bf16 Cp==4: one 8-byte store per pixel
__global__ void image_nchw_to_nhwc4_bf16_kernel(const float* __restrict__ src,
                                                uint2* __restrict__ dst, int C, int HW,
                                                int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t b = i / HW;
  int p = (int)(i - b * HW);
  const float* s = src + b * (int64_t)C * HW + p;
  float v0 = __ldg(s), v1 = C > 1 ? __ldg(s + HW) : 0.f, v2 = C > 2 ? __ldg(s + 2 * (int64_t)HW) : 0.f,
        v3 = C > 3 ? __ldg(s + 3 * (int64_t)HW) : 0.f;
  dst[i] = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
}

// physically padded frame: dst[b][hp][wp][c], hp in [0,Hpad), wp in [0,Wpad): image pixel
// (hp-top, wp-left) or zero
template <typename T>
__global__ void image_nchw_to_nhwc_padded_kernel(const float* __restrict__ src, T* __restrict__ dst, int C,
                                                 int H, int W, int Cp, int left, int Wpad, int top, int Hpad,
                                                 float pad_ch, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int wp = (int)(i % Wpad);
  int64_t r = i / Wpad;
  int hp = (int)(r % Hpad);
  int64_t b = r / Hpad;
  int w = wp - left, h = hp - top;
  const bool in = w >= 0 && w < W && h >= 0 && h < H;
  const float* s = src + (b * C * H + h) * (int64_t)W + w;
  T* d = dst + i * Cp;
  for (int c = 0; c < Cp; ++c) {
    float v = c < C ? (in ? __ldg(s + (int64_t)c * H * W) : 0.f) : pad_ch;
    st_from_float<T>(d + c, v);
  }
}

// bf16, Cp == 4, even W / left / Wpad: one thread converts TWO neighbouring padded pixels - three 8-byte
// plane loads, one 16-byte store (the generic kernel issues four 2-byte stores per pixel)
__global__ void image_nchw_to_nhwc4_padded_bf16_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int C,
                                                       int H, int W, int left, int Wpad2, int top, int Hpad,
                                                       float pad_ch, int64_t total2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total2) return;
  const int wp = (int)(i % Wpad2) * 2;
  const int64_t r = i / Wpad2;
  const int hp = (int)(r % Hpad);
  const int64_t b = r / Hpad;
  const int w = wp - left, h = hp - top;
  float2 v[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  if (w >= 0 && w < W && h >= 0 && h < H) {     // w even and W even: the pair is inside or outside together
    const float* s = src + (b * C * H + h) * (int64_t)W + w;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c < C) v[c] = __ldg(reinterpret_cast<const float2*>(s + (int64_t)c * H * W));
  }
  // channel 3 is padding: pad_ch everywhere (also outside the image) - 0, or 1 when it carries folded biases
  dst[i] = make_uint4(pack_bf16x2(v[0].x, v[1].x), pack_bf16x2(v[2].x, pad_ch), pack_bf16x2(v[0].y, v[1].y),
                      pack_bf16x2(v[2].y, pad_ch));
}

template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int Cout,
                                        int Cin, int KH, int KW, int Cin_pad) {
  int64_t total = (int64_t)Cout * KH * KW * Cin_pad;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % Cin_pad);
  int64_t r = i / Cin_pad;
  int kw = (int)(r % KW);
  r /= KW;
  int kh = (int)(r % KH);
  int o = (int)(r / KH);
  float v = 0.f;
  if (c < Cin) v = w[(((int64_t)o * Cin + c) * KH + kh) * KW + kw];
  st_from_float<T>(dst + i, v);
}

__global__ void fold_bn_kernel(const float* gamma, const float* beta, const float* mean,
                               const float* var, float eps, const float* conv_bias, int C,
                               float* scale, float* bias) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float cb = conv_bias ? conv_bias[c] : 0.f;
  if (gamma == nullptr) {
    scale[c] = 1.f;
    bias[c] = cb;
  } else {
    // y = (x + cb - mean) / sqrt(var + eps) * gamma + beta
    float s = gamma[c] / sqrtf(var[c] + eps);
    scale[c] = s;
    bias[c] = (cb - mean[c]) * s + beta[c];
  }
}

extern "C" {

int amoe_image_nchw_to_nhwc(amoe_ctx* ctx, const float* src, void* dst, int B, int C, int H, int W,
                            int Cp, int dst_dtype, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && src && dst, "amoe_image_nchw_to_nhwc: NULL argument");
  AMOE_REQUIRE(Cp >= C && C >= 1, "amoe_image_nchw_to_nhwc: need Cp >= C >= 1 (C=%d Cp=%d)", C, Cp);
  cudaStream_t st = (cudaStream_t)stream;
  int HW = H * W;
  int64_t total = (int64_t)B * HW;
  if (total == 0) return 0;
  int threads = 256;
  unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (dst_dtype == AMOE_BF16 && Cp == 4) {
    image_nchw_to_nhwc4_bf16_kernel<<<blocks, threads, 0, st>>>(src, (uint2*)dst, C, HW, total);
  } else if (dst_dtype == AMOE_BF16) {
    image_nchw_to_nhwc_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(src, (__nv_bfloat16*)dst, C, HW, Cp, total);
  } else if (dst_dtype == AMOE_F32) {
    image_nchw_to_nhwc_kernel<float><<<blocks, threads, 0, st>>>(src, (float*)dst, C, HW, Cp, total);
  } else {
    AMOE_REQUIRE(false, "amoe_image_nchw_to_nhwc: bad dtype %d", dst_dtype);
  }
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_image_nchw_to_nhwc_padded(amoe_ctx* ctx, const float* src, void* dst, int B, int C, int H, int W,
                                   int Cp, int left, int Wpad, int top, int Hpad, int dst_dtype, void* stream) {
  AMOE_ENTER(ctx);
  return amoe_image_nchw_to_nhwc_padded_v(ctx, src, dst, B, C, H, W, Cp, left, Wpad, top, Hpad, dst_dtype, 0.f, stream);
}

int amoe_image_nchw_to_nhwc_padded_v(amoe_ctx* ctx, const float* src, void* dst, int B, int C, int H, int W,
                                     int Cp, int left, int Wpad, int top, int Hpad, int dst_dtype, float pad_channel_value,
                                     void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && src && dst, "amoe_image_nchw_to_nhwc_padded: NULL argument");
  AMOE_REQUIRE(Cp >= C && C >= 1 && left >= 0 && Wpad >= left + W && top >= 0 && Hpad >= top + H,
               "amoe_image_nchw_to_nhwc_padded: bad geometry");
  int64_t total = (int64_t)B * Hpad * Wpad;
  if (total == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (dst_dtype == AMOE_BF16 && Cp == 4 && C <= 3 && W % 2 == 0 && left % 2 == 0 && Wpad % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(src) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int64_t total2 = total / 2;
    image_nchw_to_nhwc4_padded_bf16_kernel<<<(unsigned)((total2 + 255) / 256), 256, 0, st>>>(
        src, (uint4*)dst, C, H, W, left, Wpad / 2, top, Hpad, pad_channel_value, total2);
  } else if (dst_dtype == AMOE_BF16)
    image_nchw_to_nhwc_padded_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(src, (__nv_bfloat16*)dst, C, H, W, Cp, left, Wpad, top, Hpad, pad_channel_value, total);
  else if (dst_dtype == AMOE_F32)
    image_nchw_to_nhwc_padded_kernel<float><<<blocks, 256, 0, st>>>(src, (float*)dst, C, H, W, Cp, left, Wpad, top, Hpad, pad_channel_value, total);
  else
    AMOE_REQUIRE(false, "amoe_image_nchw_to_nhwc_padded: bad dtype %d", dst_dtype);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_pack_conv_weight(amoe_ctx* ctx, const float* w_oihw, void* dst, int Cout, int Cin, int KH,
                          int KW, int Cin_pad, int dst_dtype, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && w_oihw && dst, "amoe_pack_conv_weight: NULL argument");
  AMOE_REQUIRE(Cin_pad >= Cin, "amoe_pack_conv_weight: Cin_pad < Cin");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)Cout * KH * KW * Cin_pad;
  int threads = 256;
  unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (dst_dtype == AMOE_BF16)
    pack_conv_weight_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(w_oihw, (__nv_bfloat16*)dst, Cout, Cin, KH, KW, Cin_pad);
  else if (dst_dtype == AMOE_F32)
    pack_conv_weight_kernel<float><<<blocks, threads, 0, st>>>(w_oihw, (float*)dst, Cout, Cin, KH, KW, Cin_pad);
  else
    AMOE_REQUIRE(false, "amoe_pack_conv_weight: bad dtype %d", dst_dtype);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_fold_bn(amoe_ctx* ctx, const float* gamma, const float* beta, const float* mean,
                 const float* var, float eps, const float* conv_bias, int C, float* scale,
                 float* bias, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && scale && bias, "amoe_fold_bn: NULL argument");
  AMOE_REQUIRE(gamma == nullptr || (beta && mean && var), "amoe_fold_bn: incomplete BN parameters");
  fold_bn_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, conv_bias, C, scale, bias);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
