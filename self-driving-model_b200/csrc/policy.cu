// Policy head: EasyBackbone global-average-pool + fc, concat with the gated context,
// and both TrajectoryPolicy MLP heads (trajectory_head.py:25-33,44-63) in one launch.
//
// Flat parameter layout (fp32, 4-float aligned tensors, PyTorch [out,in] weights):
//   backbone.fc.W[bd,Cf] b[bd]
//   head_wp.0.W[hid,bd+ctx] b | head_wp.2.W[hid,hid] b | head_wp.4.W[2*hz,hid] b
//   head_spd.0.W[hid,bd+ctx] b | head_spd.2.W[hid,hid] b | head_spd.4.W[hz,hid] b
#include "common.cuh"
#include "mlp.cuh"

struct PolicyDims {
  int B, HW, Cf, bd, ctx_dim, hid, hz;
};

template <typename T, int FT, bool CL>
__global__ __launch_bounds__(GATE_THREADS) void policy_head_kernel(
    PolicyDims d, const T* __restrict__ x, const float* __restrict__ ctx,
    const float* __restrict__ prm, float* __restrict__ waypoints, float* __restrict__ speed) {
  extern __shared__ __align__(16) float sm[];
  const int f0 = (CL ? blockIdx.x / CL_RANKS : blockIdx.x) * FT;
  const int in_dim = d.bd + d.ctx_dim;
  const int ld_c = (int)al4(d.Cf), ld_in = (int)al4(in_dim), ld_h = (int)al4(d.hid);
  float* s_pool = sm;                      // [FT][Cf]
  float* s_in = s_pool + FT * ld_c;   // [FT][bd+ctx]
  float* s_h1 = s_in + FT * ld_in;    // [FT][hid]
  float* s_h2 = s_h1 + FT * ld_h;     // [FT][hid]
  float* s_o = s_h2 + FT * ld_h;      // [FT][2*hz]
  const int ld_o = (int)al4(2 * d.hz);

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
  const float* Wfc = p; p += al4((int64_t)d.bd * d.Cf);
  const float* bfc = p; p += al4(d.bd);
  linear_ft<FT, CL>(Wfc, bfc, s_pool, ld_c, d.Cf, s_in, ld_in, d.bd, false);  // feat -> s_in[:, :bd]

  for (int head = 0; head < 2; ++head) {
    const int out_dim = head == 0 ? 2 * d.hz : d.hz;
    const float* W0 = p; p += al4((int64_t)d.hid * in_dim);
    const float* b0 = p; p += al4(d.hid);
    const float* W2 = p; p += al4((int64_t)d.hid * d.hid);
    const float* b2 = p; p += al4(d.hid);
    const float* W4 = p; p += al4((int64_t)out_dim * d.hid);
    const float* b4 = p; p += al4(out_dim);
    linear_ft<FT, CL>(W0, b0, s_in, ld_in, in_dim, s_h1, ld_h, d.hid, true);
    linear_ft<FT, CL>(W2, b2, s_h1, ld_h, d.hid, s_h2, ld_h, d.hid, true);
    linear_ft<FT, CL>(W4, b4, s_h2, ld_h, d.hid, s_o, ld_o, out_dim, false);
    store_rows<FT, CL>(head == 0 ? waypoints : speed, out_dim, s_o, ld_o, out_dim, f0, d.B);
    __syncthreads();   // s_o is rewritten two cluster barriers later at the earliest
  }
}

static int64_t policy_param_count(const PolicyDims& d) {
  int in_dim = d.bd + d.ctx_dim;
  int64_t n = al4((int64_t)d.bd * d.Cf) + al4(d.bd);
  for (int head = 0; head < 2; ++head) {
    int out_dim = head == 0 ? 2 * d.hz : d.hz;
    n += al4((int64_t)d.hid * in_dim) + al4(d.hid) + al4((int64_t)d.hid * d.hid) + al4(d.hid) +
         al4((int64_t)out_dim * d.hid) + al4(out_dim);
  }
  return n;
}

extern "C" int amoe_policy_head_fwd(amoe_ctx* ctx, const void* x, const float* cvec,
                                    const float* params, int64_t n_params, int B, int HW, int Cf,
                                    int backbone_dim, int ctx_dim, int hidden, int horizon,
                                    int x_dtype, float* waypoints, float* speed, void* stream) {
  AMOE_REQUIRE(ctx && x && params && waypoints && speed, "amoe_policy_head_fwd: NULL argument");
  AMOE_REQUIRE(ctx_dim == 0 || cvec, "amoe_policy_head_fwd: ctx is NULL but ctx_dim=%d", ctx_dim);
  PolicyDims d;
  d.B = B; d.HW = HW; d.Cf = Cf; d.bd = backbone_dim; d.ctx_dim = ctx_dim; d.hid = hidden; d.hz = horizon;
  int64_t need = policy_param_count(d);
  AMOE_REQUIRE(n_params == need, "amoe_policy_head_fwd: params has %lld floats, layout needs %lld",
               (long long)n_params, (long long)need);
  if (B == 0) return 0;
  const size_t per_frame = sizeof(float) * (al4(Cf) + al4(backbone_dim + ctx_dim) + 2 * al4(hidden) + al4(2 * horizon));
  AMOE_REQUIRE(per_frame * GATE_FT <= 200 * 1024, "amoe_policy_head_fwd: dims too large for shared memory");
  AMOE_REQUIRE(x_dtype == AMOE_BF16 || x_dtype == AMOE_F32, "amoe_policy_head_fwd: bad dtype %d", x_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const bool cluster = mlp_use_cluster(B) && per_frame * CL_FT <= 200 * 1024;
  const size_t smem = per_frame * (cluster ? CL_FT : GATE_FT);
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  const float* xf = (const float*)x;
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
      auto kern = policy_head_kernel<__nv_bfloat16, CL_FT, true>;
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMOE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, d, xb, cvec, params, waypoints, speed));
    } else {
      auto kern = policy_head_kernel<float, CL_FT, true>;
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      AMOE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, d, xf, cvec, params, waypoints, speed));
    }
  } else if (x_dtype == AMOE_BF16) {
    auto kern = policy_head_kernel<__nv_bfloat16, GATE_FT, false>;
    if (smem > 48 * 1024) AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ceil_div(B, GATE_FT), GATE_THREADS, smem, st>>>(d, xb, cvec, params, waypoints, speed);
  } else {
    auto kern = policy_head_kernel<float, GATE_FT, false>;
    if (smem > 48 * 1024) AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ceil_div(B, GATE_FT), GATE_THREADS, smem, st>>>(d, xf, cvec, params, waypoints, speed);
  }
  AMOE_LAUNCH_OK(ctx);
  return 0;
}
