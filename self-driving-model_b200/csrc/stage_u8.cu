// Input staging on the device: camera frames arrive as uint8 HWC RGB and leave as the
// normalised tensor the first convolutions read.  Replaces the per-frame CPU transform of the reference
// (inference/run_automoe.py:25-31: ToPILImage -> Resize(bilinear) -> ToTensor -> Normalize) so that a batch
// crosses PCIe as 3 bytes per pixel instead of 12.
//
//   resample_u8_kernel       one separable pass of Pillow's 8-bit ImagingResample (fixed-point coefficients
//                            prepared by the caller exactly as Pillow's precompute_coeffs/normalize_coeffs_8bpc;
//                            rounding and clipping per pass as Pillow does) - bit-exact integer work
//   stage_u8_nhwc4_kernel    (u/255 - mean)/std in IEEE fp32 (the op order of ToTensor + Normalize) through a
//                            256-entry table per channel, written as the zero-bordered bf16 NHWC4 frame of the
//                            tensor-core stem; four pixels per thread: three 4-byte loads, two 16-byte stores
//   normalize_u8_nchw_kernel same arithmetic into the reference's fp32 NCHW tensor (fp32 mode, parity tests)
#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Pillow Resample.c

__device__ __forceinline__ float norm_value(int u, float mean, float stdv) {
  // ToTensor: u8 -> float32 / 255; Normalize: (x - mean) / std  - each a correctly rounded fp32 op
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), mean), stdv);
}

__device__ __forceinline__ void fill_lut(float* lut, float m0, float m1, float m2, float s0, float s1, float s2) {
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, u = i & 255;
    lut[i] = norm_value(u, c == 0 ? m0 : (c == 1 ? m1 : m2), c == 0 ? s0 : (c == 1 ? s1 : s2));
  }
  __syncthreads();
}

// dst [B,Hpad,Wpad,4] bf16; W, left, Wpad multiples of 4.  One thread = four neighbouring padded pixels.
__global__ void __launch_bounds__(256) stage_u8_nhwc4_kernel(const uint8_t* __restrict__ src, uint4* __restrict__ dst,
                                                             int H, int W, int left, int Wpad4, int top, int Hpad,
                                                             float m0, float m1, float m2, float s0, float s1, float s2,
                                                             float pad_ch, int64_t total4) {
  __shared__ float lut[768];
  fill_lut(lut, m0, m1, m2, s0, s1, s2);
  const uint32_t padbits = pack_bf16x2(0.f, pad_ch);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wpad4) * 4;
    const int64_t r = i / Wpad4;
    const int hp = (int)(r % Hpad);
    const int64_t b = r / Hpad;
    const int w = wp - left, h = hp - top;
    uint4 o0 = make_uint4(0u, padbits, 0u, padbits), o1 = o0;
    if (w >= 0 && w < W && h >= 0 && h < H) {     // groups of four are inside or outside together
      const uint32_t* s = reinterpret_cast<const uint32_t*>(src + ((b * H + h) * (int64_t)W + w) * 3);
      const uint32_t a = __ldg(s), bb = __ldg(s + 1), c = __ldg(s + 2);   // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
      const float* lr = lut;
      const float* lg = lut + 256;
      const float* lb = lut + 512;
      o0.x = pack_bf16x2(lr[a & 255], lg[(a >> 8) & 255]);
      o0.y = pack_bf16x2(lb[(a >> 16) & 255], pad_ch);
      o0.z = pack_bf16x2(lr[a >> 24], lg[bb & 255]);
      o0.w = pack_bf16x2(lb[(bb >> 8) & 255], pad_ch);
      o1.x = pack_bf16x2(lr[(bb >> 16) & 255], lg[bb >> 24]);
      o1.y = pack_bf16x2(lb[c & 255], pad_ch);
      o1.z = pack_bf16x2(lr[(c >> 8) & 255], lg[(c >> 16) & 255]);
      o1.w = pack_bf16x2(lb[c >> 24], pad_ch);
    }
    dst[2 * i] = o0;
    dst[2 * i + 1] = o1;
  }
}

// any geometry: one thread per padded pixel
__global__ void __launch_bounds__(256) stage_u8_nhwc4_generic_kernel(const uint8_t* __restrict__ src, uint2* __restrict__ dst,
                                                                     int H, int W, int left, int Wpad, int top, int Hpad,
                                                                     float m0, float m1, float m2, float s0, float s1, float s2,
                                                                     float pad_ch, int64_t total) {
  __shared__ float lut[768];
  fill_lut(lut, m0, m1, m2, s0, s1, s2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wpad);
    const int64_t r = i / Wpad;
    const int hp = (int)(r % Hpad);
    const int64_t b = r / Hpad;
    const int w = wp - left, h = hp - top;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (w >= 0 && w < W && h >= 0 && h < H) {
      const uint8_t* s = src + ((b * H + h) * (int64_t)W + w) * 3;
      v0 = lut[__ldg(s)];
      v1 = lut[256 + __ldg(s + 1)];
      v2 = lut[512 + __ldg(s + 2)];
    }
    dst[i] = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, pad_ch));
  }
}

// dst [B,3,H,W] fp32 (the tensor the reference transform produces)
__global__ void __launch_bounds__(256) normalize_u8_nchw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int HW,
                                                                float m0, float m1, float m2, float s0, float s1, float s2,
                                                                int64_t total) {
  __shared__ float lut[768];
  fill_lut(lut, m0, m1, m2, s0, s1, s2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW;
    const int p = (int)(i - b * HW);
    const uint8_t* s = src + i * 3;
    float* d = dst + b * 3 * (int64_t)HW + p;
    d[0] = lut[__ldg(s)];
    d[HW] = lut[256 + __ldg(s + 1)];
    d[2 * (int64_t)HW] = lut[512 + __ldg(s + 2)];
  }
}

// One pass of Pillow's ImagingResampleHorizontal_8bpc / Vertical_8bpc over a tensor viewed as
// [outer][in_size][inner] -> [outer][out_size][inner] (horizontal: outer = B*H, inner = 3; vertical: outer = B,
// inner = W*3).  bounds[2*o] = first input index, bounds[2*o+1] = tap count, kk[o*ksize + t] = fixed-point weight.
__global__ void __launch_bounds__(256) resample_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          const int* __restrict__ bounds, const int* __restrict__ kk,
                                                          int ksize, int in_size, int out_size, int inner, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % inner);
    const int64_t r = i / inner;
    const int o = (int)(r % out_size);
    const int64_t outer = r / out_size;
    const int first = __ldg(bounds + 2 * o), n = __ldg(bounds + 2 * o + 1);
    const uint8_t* s = src + (outer * in_size + first) * (int64_t)inner + c;
    const int* k = kk + (int64_t)o * ksize;
    int ss = 1 << (PRECISION_BITS - 1);
    for (int t = 0; t < n; ++t) ss += (int)__ldg(s + (int64_t)t * inner) * __ldg(k + t);
    ss >>= PRECISION_BITS;            // arithmetic shift, then Pillow's clip8 lookup
    dst[i] = (uint8_t)min(max(ss, 0), 255);
  }
}

unsigned grid_for(const amoe_ctx* ctx, int64_t work_items, int threads) {
  const int64_t want = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)ctx->sm_count * 16;      // grid-stride loops: a few waves of resident CTAs
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

extern "C" {

int amoe_stage_u8_hwc_fwd(amoe_ctx* ctx, const void* src_u8, void* dst, int B, int H, int W, int left, int Wpad, int top,
                          int Hpad, const float* mean3_host, const float* std3_host, float pad_channel_value, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && mean3_host && std3_host && (B == 0 || (src_u8 && dst)), "amoe_stage_u8_hwc_fwd: NULL argument");
  AMOE_REQUIRE(left >= 0 && Wpad >= left + W && top >= 0 && Hpad >= top + H && B >= 0 && H > 0 && W > 0,
               "amoe_stage_u8_hwc_fwd: bad geometry");
  for (int c = 0; c < 3; ++c) AMOE_REQUIRE(std3_host[c] != 0.f, "amoe_stage_u8_hwc_fwd: std[%d] is zero", c);
  const int64_t total = (int64_t)B * Hpad * Wpad;
  if (total == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* m = mean3_host;
  const float* s = std3_host;
  if (W % 4 == 0 && left % 4 == 0 && Wpad % 4 == 0 && (reinterpret_cast<uintptr_t>(src_u8) & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int64_t total4 = total / 4;
    stage_u8_nhwc4_kernel<<<grid_for(ctx, total4, 256), 256, 0, st>>>((const uint8_t*)src_u8, (uint4*)dst, H, W, left, Wpad / 4,
                                                                      top, Hpad, m[0], m[1], m[2], s[0], s[1], s[2],
                                                                      pad_channel_value, total4);
  } else {
    stage_u8_nhwc4_generic_kernel<<<grid_for(ctx, total, 256), 256, 0, st>>>((const uint8_t*)src_u8, (uint2*)dst, H, W, left, Wpad,
                                                                             top, Hpad, m[0], m[1], m[2], s[0], s[1], s[2],
                                                                             pad_channel_value, total);
  }
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_normalize_u8_hwc_to_nchw_fwd(amoe_ctx* ctx, const void* src_u8, float* dst, int B, int H, int W,
                                      const float* mean3_host, const float* std3_host, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && mean3_host && std3_host && (B == 0 || (src_u8 && dst)), "amoe_normalize_u8_hwc_to_nchw_fwd: NULL argument");
  for (int c = 0; c < 3; ++c) AMOE_REQUIRE(std3_host[c] != 0.f, "amoe_normalize_u8_hwc_to_nchw_fwd: std[%d] is zero", c);
  const int64_t total = (int64_t)B * H * W;
  if (total == 0) return 0;
  const float* m = mean3_host;
  const float* s = std3_host;
  normalize_u8_nchw_kernel<<<grid_for(ctx, total, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src_u8, dst, H * W, m[0], m[1],
                                                                                        m[2], s[0], s[1], s[2], total);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_resample_u8_fwd(amoe_ctx* ctx, const void* src_u8, void* dst_u8, const int* bounds, const int* coeffs, int ksize,
                         int64_t outer, int in_size, int out_size, int inner, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && bounds && coeffs && (outer == 0 || (src_u8 && dst_u8)), "amoe_resample_u8_fwd: NULL argument");
  AMOE_REQUIRE(ksize > 0 && in_size > 0 && out_size > 0 && inner > 0 && outer >= 0, "amoe_resample_u8_fwd: bad geometry");
  const int64_t total = outer * out_size * inner;
  if (total == 0) return 0;
  resample_u8_kernel<<<grid_for(ctx, total, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src_u8, (uint8_t*)dst_u8, bounds,
                                                                                  coeffs, ksize, in_size, out_size, inner, total);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
