// Expert heads: 1x1 classifier conv + global mean, bilinear up-sampling writer
// (NHWC low-res fp32 -> NCHW full-res), and a plain NCHW spatial mean.
#include "common.cuh"

// ---- 1x1 conv + mean over pixels -------------------------------------------
// One CTA per image.  Weights [N][Cin] (fp32) and a chunk of up to HEAD_PX pixels (activations as
// fp32, filled with 16-byte global loads) live in shared memory, rows padded to an odd stride ->
// conflict-free; one thread per (pixel, n) output does a Cin-long dot product.  Phase 2:
// pooled[b,n] = mean_p low[b,p,n], summed in a fixed order (deterministic: the gate's top-1
// routing depends on it).
constexpr int HEAD_PX = 64;

template <typename T, int MAXO>
__global__ __launch_bounds__(256) void head1x1_pool_kernel(const T* __restrict__ x,
                                                           const float* __restrict__ w,
                                                           const float* __restrict__ bias,
                                                           float* __restrict__ low,
                                                           float* __restrict__ pooled, int pooled_ld,
                                                           int HW, int Cin, int N) {
  // One CTA per image.  Thread (px = tid % 64, ng = tid / 64) owns pixel px of the current 64-pixel chunk and the
  // outputs n = ng, ng+4, ... : per 4 input channels it reads 4 activations (transposed tile, conflict-free) and
  // one 16-byte weight vector per output (same address for the whole warp -> broadcast), i.e. ~0.45 shared-memory
  // reads per FMA instead of 2 (the previous thread-per-output loop took 32-39 us per launch on the critical path).
  extern __shared__ float smh[];
  const int ldw = (Cin + 3) & ~3;      // weight rows 16-byte aligned
  float* sw = smh;                     // [N][ldw]
  float* sxT = smh + N * ldw;          // [Cin][HEAD_PX] transposed activation tile
  float* s_out = sxT + Cin * HEAD_PX;  // [HEAD_PX][N] outputs of the current chunk (for the pooled mean)
  float pool_acc = 0.f;                // thread n < N: sum over the pixels, in pixel order (deterministic)
  const int b = blockIdx.x;
  // (loads are issued in batches before the dependent shared-memory stores: a plain load-store loop is one
  //  L2/HBM round trip per iteration for an in-order warp - ~1.1 us per output channel before)
  for (int i0 = threadIdx.x; i0 < N * Cin; i0 += 8 * blockDim.x) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < N * Cin ? __ldg(w + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < N * Cin) sw[(i / Cin) * ldw + (i % Cin)] = v[u];
    }
  }
  const T* xb = x + (int64_t)b * HW * Cin;
  float* lb = low + (int64_t)b * HW * N;
  constexpr int VE = 16 / sizeof(T);   // elements per 16-byte load
  const bool vec = (Cin % VE) == 0 && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0);
  const int px = threadIdx.x & (HEAD_PX - 1), ng = threadIdx.x / HEAD_PX;   // 256 threads = 64 pixels x 4 output groups
  constexpr int NG = 256 / HEAD_PX;                                          // MAXO outputs per thread and pass (4*MAXO per pass)
  const uint32_t sw_s = (uint32_t)__cvta_generic_to_shared(sw), sx_s = (uint32_t)__cvta_generic_to_shared(sxT) + px * 4u;
  for (int p0 = 0; p0 < HW; p0 += HEAD_PX) {
    const int np = min(HEAD_PX, HW - p0);
    __syncthreads();  // previous chunk consumed (also orders the weight fill before first use)
    // stage the chunk transposed: a warp covers 32 pixels of one channel group, so its stores hit 32 different banks
    if (vec) {
      const uint4* xrow = reinterpret_cast<const uint4*>(xb + (int64_t)(p0 + min(px, np - 1)) * Cin);
      for (int cg0 = ng; cg0 < Cin / VE; cg0 += 4 * NG) {     // four 16-byte loads in flight per thread
        uint4 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cg = cg0 + u * NG;
          raw[u] = cg < Cin / VE ? __ldg(xrow + cg) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cg = cg0 + u * NG;
          if (cg < Cin / VE && px < np) {
            const T* e = reinterpret_cast<const T*>(&raw[u]);
#pragma unroll
            for (int k = 0; k < VE; ++k) sxT[(cg * VE + k) * HEAD_PX + px] = ld_as_float<T>(e + k);
          }
        }
      }
    } else {
      for (int c = ng; c < Cin; c += NG)
        if (px < np) sxT[c * HEAD_PX + px] = ld_as_float<T>(xb + (int64_t)(p0 + px) * Cin + c);
    }
    __syncthreads();
    for (int n0 = 0; n0 < N; n0 += NG * MAXO) {
      float acc[MAXO];
      uint32_t wrow[MAXO];   // shared address of this thread's weight rows; rows past N re-read row N-1 (results dropped)
#pragma unroll
      for (int o = 0; o < MAXO; ++o) {
        acc[o] = 0.f;
        wrow[o] = sw_s + (uint32_t)(min(n0 + ng + o * NG, N - 1) * ldw) * 4u;
      }
      if (px < np) {
        int c = 0;
        for (; c + 3 < Cin; c += 4) {      // branch-free: 4 activations + MAXO 16-byte weight vectors + 4*MAXO FMAs
          float xv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv[k]) : "r"(sx_s + (uint32_t)((c + k) * HEAD_PX) * 4u));
#pragma unroll
          for (int o = 0; o < MAXO; ++o) {
            float4 w4;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w4.x), "=f"(w4.y), "=f"(w4.z), "=f"(w4.w) : "r"(wrow[o] + (uint32_t)c * 4u));
            acc[o] = fmaf(xv[0], w4.x, acc[o]);
            acc[o] = fmaf(xv[1], w4.y, acc[o]);
            acc[o] = fmaf(xv[2], w4.z, acc[o]);
            acc[o] = fmaf(xv[3], w4.w, acc[o]);
          }
        }
        for (; c < Cin; ++c) {
          const float xs = sxT[c * HEAD_PX + px];
#pragma unroll
          for (int o = 0; o < MAXO; ++o) acc[o] = fmaf(xs, sw[min(n0 + ng + o * NG, N - 1) * ldw + c], acc[o]);
        }
#pragma unroll
        for (int o = 0; o < MAXO; ++o) {
          const int n = n0 + ng + o * NG;
          if (n < N) {
            const float v = acc[o] + bias[n];
            lb[(int64_t)(p0 + px) * N + n] = v;
            s_out[px * N + n] = v;
          }
        }
      }
    }
    // pooled mean from the shared-memory copy (summing the image back from global memory, as before, was a chain
    // of HW dependent L2 round trips: ~25 us of the kernel's 32-39)
    __syncthreads();
    if ((int)threadIdx.x < N)
      for (int p = 0; p < np; ++p) pool_acc += s_out[p * N + threadIdx.x];
  }
  if ((int)threadIdx.x < N) pooled[(int64_t)b * pooled_ld + threadIdx.x] = pool_acc / (float)HW;
}

// ---- bilinear up-sampling writer ---------------------------------------------
// out[b,c,y,x] = bilinear(low[b,:,:,c]) with PyTorch's align_corners=False rule
// (aten upsample_bilinear2d: src = max(scale*(dst+0.5)-0.5, 0), scale = in/out).
// One thread writes 8 consecutive x (16 B in bf16).  Grid: (W/8-chunks, H, B*C).
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = min((int)s, in_size - 1);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

template <typename T>
__global__ __launch_bounds__(256) void upsample_bilinear_nchw_kernel(const float* __restrict__ low,
                                                                     T* __restrict__ out, int h,
                                                                     int w, int C, int H, int W,
                                                                     float sh, float sw_, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int wc = (W + 7) >> 3;
  int xc = (int)(i % wc);
  int64_t r = i / wc;
  int y = (int)(r % H);
  r /= H;
  int c = (int)(r % C);
  int64_t b = r / C;
  int y0, y1;
  float ly;
  src_index(sh, y, h, y0, y1, ly);
  const float hy = 1.f - ly;
  const float* l0 = low + ((b * h + y0) * w) * (int64_t)C + c;
  const float* l1p = low + ((b * h + y1) * w) * (int64_t)C + c;
  float v[8];
  int x0 = xc * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int x = x0 + j;
    int xa, xb;
    float lx;
    src_index(sw_, min(x, W - 1), w, xa, xb, lx);
    float hx = 1.f - lx;
    float v00 = __ldg(l0 + (int64_t)xa * C), v01 = __ldg(l0 + (int64_t)xb * C);
    float v10 = __ldg(l1p + (int64_t)xa * C), v11 = __ldg(l1p + (int64_t)xb * C);
    v[j] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
  T* o = out + ((b * C + c) * (int64_t)H + y) * W + x0;
  if (x0 + 8 <= W && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
    if constexpr (sizeof(T) == 2) {
      uint4 pk = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                            pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      __stcs(reinterpret_cast<uint4*>(o), pk);
    } else {
      __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
      __stcs(reinterpret_cast<float4*>(o) + 1, make_float4(v[4], v[5], v[6], v[7]));
    }
  } else {
    for (int j = 0; j < 8 && x0 + j < W; ++j) st_from_float<T>(o + j, v[j]);
  }
}

// ---- fast path: integer scale factor S with S % 16 == 0 and w <= 32 (the x32 case) -------------
// One CTA per (b,c) plane: the low-res plane sits in smem; each warp produces whole output rows.
// Per row: lanes < w blend the two source rows (vertical lerp), every lane then needs only two of
// those values (a lane's 8 consecutive pixels share one source cell because cell borders fall on
// multiples of 8), fetched by shuffle, and writes one 16-byte (bf16) / two 16-byte (fp32) stores.
template <typename T>
__global__ __launch_bounds__(256) void upsample_intscale_nchw_kernel(const float* __restrict__ low,
                                                                     T* __restrict__ out, int h, int w,
                                                                     int C, int H, int W, float sh, float sw_) {
  extern __shared__ float plane[];  // [h][w]
  const int bc = blockIdx.x;
  const int b = bc / C, c = bc - b * C;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) plane[i] = __ldg(low + ((int64_t)b * h * w + i) * C + c);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  T* obase = out + (int64_t)bc * H * W;
  if (W == 256) {
    // One 16-byte store per lane covers a row: the horizontal source cell and the eight blend weights of a lane
    // are the same for every row, so they are computed once; a row then costs two shared-memory reads, two
    // shuffles, nine FMAs, the packing and the store (the generic loop below spends ~50 instructions per 512
    // bytes and was issue-bound at 3.0 TB/s).
    const int x0 = lane * 8;
    int xa, xb;
    float lx0;
    src_index(sw_, x0, w, xa, xb, lx0);
    float lx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) lx[j] = fmaxf(sw_ * ((float)(x0 + j) + 0.5f) - 0.5f, 0.f) - (float)xa;
    const int lw = min(lane, w - 1);
    for (int y = warp; y < H; y += nwarp) {
      int y0, y1;
      float ly;
      src_index(sh, y, h, y0, y1, ly);
      const float vcol = (1.f - ly) * plane[y0 * w + lw] + ly * plane[y1 * w + lw];
      const float va = __shfl_sync(0xffffffffu, vcol, xa), vb = __shfl_sync(0xffffffffu, vcol, xb);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (1.f - lx[j]) * va + lx[j] * vb;
      T* o = obase + (int64_t)y * W + x0;
      if constexpr (sizeof(T) == 2) {
        __stcs(reinterpret_cast<uint4*>(o), make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                       pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
      } else {
        __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4*>(o) + 1, make_float4(v[4], v[5], v[6], v[7]));
      }
    }
    return;
  }
  for (int y = warp; y < H; y += nwarp) {
    int y0, y1;
    float ly;
    src_index(sh, y, h, y0, y1, ly);
    const float hy = 1.f - ly;
    float vcol = 0.f;
    if (lane < w) vcol = hy * plane[y0 * w + lane] + ly * plane[y1 * w + lane];
    for (int x0 = lane * 8; x0 < W; x0 += 256) {
      int xa, xb;
      float lx0;
      src_index(sw_, x0, w, xa, xb, lx0);
      const float va = __shfl_sync(0xffffffffu, vcol, xa), vb = __shfl_sync(0xffffffffu, vcol, xb);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = fmaxf(sw_ * ((float)(x0 + j) + 0.5f) - 0.5f, 0.f);
        float lx = s - (float)xa;
        v[j] = (1.f - lx) * va + lx * vb;
      }
      T* o = obase + (int64_t)y * W + x0;
      if constexpr (sizeof(T) == 2) {
        __stcs(reinterpret_cast<uint4*>(o), make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                       pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
      } else {
        __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4*>(o) + 1, make_float4(v[4], v[5], v[6], v[7]));
      }
    }
  }
}

// ---- mean over H*W of NCHW: one warp per (b,c) --------------------------------
template <typename T>
__global__ void mean_hw_nchw_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t BC, int HW) {
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (wid >= BC) return;
  const T* p = x + wid * HW;
  float s = 0.f;
  for (int i = lane; i < HW; i += 32) s += ld_as_float<T>(p + i);
  s = warp_sum(s);
  if (lane == 0) out[wid] = s / (float)HW;
}

extern "C" {

int amoe_head1x1_pool_fwd(amoe_ctx* ctx, const void* x, const float* w, const float* b, float* low,
                          float* pooled, int pooled_ld, int B, int HW, int Cin, int N, int x_dtype,
                          void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && w && b && low && pooled, "amoe_head1x1_pool_fwd: NULL argument");
  AMOE_REQUIRE(N >= 1 && N <= 256, "amoe_head1x1_pool_fwd: N=%d out of range [1,256]", N);
  size_t smem = ((size_t)N * ((Cin + 3) & ~3) + (size_t)Cin * HEAD_PX + (size_t)HEAD_PX * N) * sizeof(float);
  AMOE_REQUIRE(smem <= 200 * 1024, "amoe_head1x1_pool_fwd: (N+%d)*Cin too large for shared memory (N=%d Cin=%d)", HEAD_PX, N, Cin);
  AMOE_REQUIRE(x_dtype == AMOE_BF16 || x_dtype == AMOE_F32, "amoe_head1x1_pool_fwd: bad dtype %d", x_dtype);
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // outputs per thread and pass: the smallest instantiation that covers N in one pass (4 output groups per CTA)
  const int per = (N + 3) / 4;
#define AMOE_HEAD_LAUNCH(TT, MO)                                                                                          \
  do {                                                                                                                    \
    auto kern = head1x1_pool_kernel<TT, MO>;                                                                              \
    if (smem > 48 * 1024) AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<B, 256, smem, st>>>((const TT*)x, w, b, low, pooled, pooled_ld, HW, Cin, N);                                   \
  } while (0)
#define AMOE_HEAD_DISPATCH(TT)                                   \
  do {                                                           \
    if (per <= 1) AMOE_HEAD_LAUNCH(TT, 1);                       \
    else if (per <= 2) AMOE_HEAD_LAUNCH(TT, 2);                  \
    else if (per <= 4) AMOE_HEAD_LAUNCH(TT, 4);                  \
    else if (per <= 6) AMOE_HEAD_LAUNCH(TT, 6);                  \
    else AMOE_HEAD_LAUNCH(TT, 8);                                \
  } while (0)
  if (x_dtype == AMOE_BF16) AMOE_HEAD_DISPATCH(__nv_bfloat16);
  else AMOE_HEAD_DISPATCH(float);
#undef AMOE_HEAD_DISPATCH
#undef AMOE_HEAD_LAUNCH
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_upsample_bilinear_nchw_fwd(amoe_ctx* ctx, const float* low, void* out, int B, int h, int w,
                                    int C, int H, int W, int out_dtype, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && low && out, "amoe_upsample_bilinear_nchw_fwd: NULL argument");
  int64_t total = (int64_t)B * C * H * ((W + 7) / 8);
  if (total == 0) return 0;
  float sh = (float)h / (float)H, sw_ = (float)w / (float)W;
  cudaStream_t st = (cudaStream_t)stream;
  // x32-style integer up-sampling: plane-per-CTA kernel (W % 256 == 0 keeps every lane's 8-pixel run
  // inside the row and 16-byte aligned; S % 16 == 0 keeps a run inside one source cell)
  if (W % w == 0 && (W / w) % 16 == 0 && w <= 32 && W % 256 == 0 && h * w * sizeof(float) <= 48 * 1024 &&
      (out_dtype == AMOE_BF16 || out_dtype == AMOE_F32)) {
    size_t smem = (size_t)h * w * sizeof(float);
    if (out_dtype == AMOE_BF16)
      upsample_intscale_nchw_kernel<__nv_bfloat16><<<B * C, 256, smem, st>>>(low, (__nv_bfloat16*)out, h, w, C, H, W, sh, sw_);
    else
      upsample_intscale_nchw_kernel<float><<<B * C, 256, smem, st>>>(low, (float*)out, h, w, C, H, W, sh, sw_);
    AMOE_LAUNCH_OK(ctx);
    return 0;
  }
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (out_dtype == AMOE_BF16)
    upsample_bilinear_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(low, (__nv_bfloat16*)out, h, w, C, H, W, sh, sw_, total);
  else if (out_dtype == AMOE_F32)
    upsample_bilinear_nchw_kernel<float><<<blocks, 256, 0, st>>>(low, (float*)out, h, w, C, H, W, sh, sw_, total);
  else
    AMOE_REQUIRE(false, "amoe_upsample_bilinear_nchw_fwd: bad dtype %d", out_dtype);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_mean_hw_nchw_fwd(amoe_ctx* ctx, const void* x, float* out, int B, int C, int HW, int dtype,
                          void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && out, "amoe_mean_hw_nchw_fwd: NULL argument");
  int64_t BC = (int64_t)B * C;
  if (BC == 0) return 0;
  unsigned blocks = (unsigned)((BC * 32 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == AMOE_BF16)
    mean_hw_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, out, BC, HW);
  else if (dtype == AMOE_F32)
    mean_hw_nchw_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, out, BC, HW);
  else
    AMOE_REQUIRE(false, "amoe_mean_hw_nchw_fwd: bad dtype %d", dtype);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
