// Policy head: EasyBackbone global-average-pool + fc, concat with the gated context,
// and both TrajectoryPolicy MLP heads (trajectory_head.py:25-33,44-63) in one launch.
//
// Flat parameter layout (fp32, 8-float aligned tensors, PyTorch [out,in] weights):
//   backbone.fc.W[bd,Cf] b[bd]
//   head_wp.0.W[hid,bd+ctx] b | head_wp.2.W[hid,hid] b | head_wp.4.W[2*hz,hid] b
//   head_spd.0.W[hid,bd+ctx] b | head_spd.2.W[hid,hid] b | head_spd.4.W[hz,hid] b
#include "common.cuh"
#include "mlp.cuh"

struct PolicyDims {
  int B, HW, Cf, bd, ctx_dim, hid, hz;
};

template <typename T, int FT, bool CL, bool TC>
__global__ __launch_bounds__(GATE_THREADS) void policy_head_kernel(
    PolicyDims d, const T* __restrict__ x, const float* __restrict__ ctx,
    const float* __restrict__ prm, const __nv_bfloat16* __restrict__ prm16, float* __restrict__ waypoints,
    float* __restrict__ speed) {
  extern __shared__ __align__(16) float sm[];
  const int f0 = (CL ? blockIdx.x / CL_RANKS : blockIdx.x) * FT;
  const int in_dim = d.bd + d.ctx_dim;
  const int ld_c = TC ? ld_tc(d.Cf) : (int)al8(d.Cf), ld_in = TC ? ld_tc(in_dim) : (int)al8(in_dim),
            ld_h = TC ? ld_tc(d.hid) : (int)al8(d.hid);
  float* s_pool = sm;                      // [FT][Cf]
  float* s_in = s_pool + FT * ld_c;   // [FT][bd+ctx]
  float* s_h1 = s_in + FT * ld_in;    // [FT][hid]
  float* s_h2 = s_h1 + FT * ld_h;     // [FT][hid]
  float* s_o = s_h2 + FT * ld_h;      // [FT][2*hz]
  const int ld_o = (int)al8(2 * d.hz);
  // tensor-core variant: the bf16 copy of the parameter buffer has the same element offsets
  auto w16 = [&](const float* W) -> const __nv_bfloat16* { return (TC && prm16) ? prm16 + (W - prm) : nullptr; };

  // AdaptiveAvgPool2d(1): channel c of frame f summed over pixels in order (thread per (f,c);
  // consecutive threads read consecutive channels -> coalesced)
  // (cluster variant: rank r pools frames r, r+8 and writes them into every rank's copy)
  for (int i = threadIdx.x; i < FT * d.Cf; i += blockDim.x) {
    int f = i / d.Cf, c = i - f * d.Cf;
    if (CL && (f % CL_RANKS) != (int)cg::this_cluster().block_rank()) continue;
    float s = 0.f;
    if (f0 + f < d.B) {
      const T* xp = x + (int64_t)(f0 + f) * d.HW * d.Cf + c;
      for (int p = 0; p < d.HW; ++p) s += ld_as_float<T>(xp + (int64_t)p * d.Cf);
    }
    const float v = s / (float)d.HW;
    if (CL) {
      for (unsigned r = 0; r < CL_RANKS; ++r) cg::this_cluster().map_shared_rank(s_pool, r)[f * ld_c + c] = v;
    } else {
      s_pool[f * ld_c + c] = v;
    }
  }
  for (int i = threadIdx.x; i < FT * d.ctx_dim; i += blockDim.x) {
    int f = i / d.ctx_dim, c = i - f * d.ctx_dim;
    s_in[f * ld_in + d.bd + c] = (f0 + f < d.B) ? ctx[(int64_t)(f0 + f) * d.ctx_dim + c] : 0.f;
  }
  mlp_sync<CL>();

  const float* p = prm;
  const float* Wfc = p; p += al8((int64_t)d.bd * d.Cf);
  const float* bfc = p; p += al8(d.bd);
  linear_ft<FT, CL, TC>(Wfc, bfc, s_pool, ld_c, d.Cf, s_in, ld_in, d.bd, false, w16(Wfc));  // feat -> s_in[:, :bd]

  for (int head = 0; head < 2; ++head) {
    const int out_dim = head == 0 ? 2 * d.hz : d.hz;
    const float* W0 = p; p += al8((int64_t)d.hid * in_dim);
    const float* b0 = p; p += al8(d.hid);
    const float* W2 = p; p += al8((int64_t)d.hid * d.hid);
    const float* b2 = p; p += al8(d.hid);
    const float* W4 = p; p += al8((int64_t)out_dim * d.hid);
    const float* b4 = p; p += al8(out_dim);
    // tensor-core variant: the two heads are independent, so they run in different CTAs (blockIdx.y = head; pool +
    // fc are recomputed by both) - the weight stream per CTA, which bounds this kernel, shrinks from 1.46 M to 0.80 M
    if (TC && (int)blockIdx.y != head) continue;   // CTA-uniform
    linear_ft<FT, CL, TC>(W0, b0, s_in, ld_in, in_dim, s_h1, ld_h, d.hid, true, w16(W0));
    linear_ft<FT, CL, TC>(W2, b2, s_h1, ld_h, d.hid, s_h2, ld_h, d.hid, true, w16(W2));
    linear_ft<FT, CL, TC>(W4, b4, s_h2, ld_h, d.hid, s_o, ld_o, out_dim, false, w16(W4));
    store_rows<FT, CL>(head == 0 ? waypoints : speed, out_dim, s_o, ld_o, out_dim, f0, d.B);
    __syncthreads();   // s_o is rewritten two cluster barriers later at the earliest
  }
}

static int64_t policy_param_count(const PolicyDims& d) {
  int in_dim = d.bd + d.ctx_dim;
  int64_t n = al8((int64_t)d.bd * d.Cf) + al8(d.bd);
  for (int head = 0; head < 2; ++head) {
    int out_dim = head == 0 ? 2 * d.hz : d.hz;
    n += al8((int64_t)d.hid * in_dim) + al8(d.hid) + al8((int64_t)d.hid * d.hid) + al8(d.hid) +
         al8((int64_t)out_dim * d.hid) + al8(out_dim);
  }
  return n;
}

// AdaptiveAvgPool2d(1) of an NHWC tensor: out[b][c] = mean over the HW pixels, summed in a fixed order
// (8 pixel slices per CTA, each accumulated front to back, then slice 0+1+...+7) -> deterministic.
// One CTA per frame, a lane owns 8 channels (one 16-byte load per pixel), so a warp reads whole 512-byte
// pixel rows; this spreads the 33 MB read over all SMs instead of the 16 CTAs of the tensor-core head.
__global__ __launch_bounds__(256) void mean_hw_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                                 int HW, int C) {
  __shared__ float part[8][8 * 32 + 8];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (HW + 7) >> 3, p0 = warp * per, p1 = min(HW, p0 + per);
  for (int cb = 0; cb < C; cb += 256) {          // CTA-uniform trip count
    const int c0 = cb + lane * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {
      const __nv_bfloat16* xp = x + (int64_t)b * HW * C + c0;
#pragma unroll 8
      for (int p = p0; p < p1; ++p) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(xp + (int64_t)p * C));
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(v2[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    if (warp == 0 && c0 < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = lane * 8 + j;
        float t = part[0][i];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += part[q][i];
        out[(int64_t)b * C + c0 + j] = t / (float)HW;
      }
    }
    __syncthreads();
  }
}

extern "C" int amoe_mean_hw_nhwc_fwd(amoe_ctx* ctx, const void* x, float* out, int B, int HW, int C, int dtype,
                                     void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && out, "amoe_mean_hw_nhwc_fwd: NULL argument");
  AMOE_REQUIRE(dtype == AMOE_BF16 && C % 8 == 0 && HW >= 1, "amoe_mean_hw_nhwc_fwd: bf16 input with C %% 8 == 0 only (C=%d)", C);
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "amoe_mean_hw_nhwc_fwd: x must be 16-byte aligned");
  if (B == 0) return 0;
  mean_hw_nhwc_bf16_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, HW, C);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int amoe_policy_head_fwd_ex(amoe_ctx* ctx, const void* x, const float* cvec,
                                       const float* params, int64_t n_params, int B, int HW, int Cf,
                                       int backbone_dim, int ctx_dim, int hidden, int horizon,
                                       int x_dtype, const void* params_bf16, float* waypoints, float* speed,
                                       void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && params && waypoints && speed, "amoe_policy_head_fwd: NULL argument");
  AMOE_REQUIRE(ctx_dim == 0 || cvec, "amoe_policy_head_fwd: ctx is NULL but ctx_dim=%d", ctx_dim);
  PolicyDims d;
  d.B = B; d.HW = HW; d.Cf = Cf; d.bd = backbone_dim; d.ctx_dim = ctx_dim; d.hid = hidden; d.hz = horizon;
  int64_t need = policy_param_count(d);
  AMOE_REQUIRE(n_params == need, "amoe_policy_head_fwd: params has %lld floats, layout needs %lld",
               (long long)n_params, (long long)need);
  if (B == 0) return 0;
  AMOE_REQUIRE(x_dtype == AMOE_BF16 || x_dtype == AMOE_F32, "amoe_policy_head_fwd: bad dtype %d", x_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  const float* xf = (const float*)x;
  const __nv_bfloat16* p16 = (const __nv_bfloat16*)params_bf16;
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(params_bf16) & 15) == 0, "amoe_policy_head_fwd: params_bf16 must be 16-byte aligned");
  if (p16 && B >= MMA_FT) {
    // bf16 inference mode: 16 frames per CTA, layers on mma.sync TF32 with bf16 weights (mlp.cuh)
    const size_t smem = sizeof(float) * MMA_FT *
                        (size_t)(ld_tc(Cf) + ld_tc(backbone_dim + ctx_dim) + 2 * ld_tc(hidden) + al8(2 * horizon));
    AMOE_REQUIRE(smem <= 226 * 1024, "amoe_policy_head_fwd: dims too large for shared memory");
    if (x_dtype == AMOE_BF16) {
      auto kern = policy_head_kernel<__nv_bfloat16, MMA_FT, false, true>;
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<dim3(ceil_div(B, MMA_FT), 2), GATE_THREADS, smem, st>>>(d, xb, cvec, params, p16, waypoints, speed);
    } else {
      auto kern = policy_head_kernel<float, MMA_FT, false, true>;
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<dim3(ceil_div(B, MMA_FT), 2), GATE_THREADS, smem, st>>>(d, xf, cvec, params, p16, waypoints, speed);
    }
    AMOE_LAUNCH_OK(ctx);
    return 0;
  }
  const size_t per_frame = sizeof(float) * (al8(Cf) + al8(backbone_dim + ctx_dim) + 2 * al8(hidden) + al8(2 * horizon));
  AMOE_REQUIRE(per_frame * GATE_FT <= 200 * 1024, "amoe_policy_head_fwd: dims too large for shared memory");
  const bool cluster = mlp_use_cluster(B) && per_frame * CL_FT <= 200 * 1024;
  const size_t smem = per_frame * (cluster ? CL_FT : GATE_FT);
  if (cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(B, CL_FT) * CL_RANKS);
    cfg.blockDim = dim3(GATE_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL_RANKS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (x_dtype == AMOE_BF16) {
      auto kern = policy_head_kernel<__nv_bfloat16, CL_FT, true, false>;
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMOE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, d, xb, cvec, params, (const __nv_bfloat16*)nullptr, waypoints, speed));
    } else {
      auto kern = policy_head_kernel<float, CL_FT, true, false>;
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMOE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, d, xf, cvec, params, (const __nv_bfloat16*)nullptr, waypoints, speed));
    }
  } else if (x_dtype == AMOE_BF16) {
    auto kern = policy_head_kernel<__nv_bfloat16, GATE_FT, false, false>;
    if (smem > 48 * 1024) AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ceil_div(B, GATE_FT), GATE_THREADS, smem, st>>>(d, xb, cvec, params, (const __nv_bfloat16*)nullptr, waypoints, speed);
  } else {
    auto kern = policy_head_kernel<float, GATE_FT, false, false>;
    if (smem > 48 * 1024) AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ceil_div(B, GATE_FT), GATE_THREADS, smem, st>>>(d, xf, cvec, params, (const __nv_bfloat16*)nullptr, waypoints, speed);
  }
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int amoe_policy_head_fwd(amoe_ctx* ctx, const void* x, const float* cvec,
                                    const float* params, int64_t n_params, int B, int HW, int Cf,
                                    int backbone_dim, int ctx_dim, int hidden, int horizon,
                                    int x_dtype, float* waypoints, float* speed, void* stream) {
  AMOE_ENTER(ctx);
  return amoe_policy_head_fwd_ex(ctx, x, cvec, params, n_params, B, HW, Cf, backbone_dim, ctx_dim, hidden, horizon,
                                 x_dtype, nullptr, waypoints, speed, stream);
}
