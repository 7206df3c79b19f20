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

template <typename T>
__global__ __launch_bounds__(GATE_THREADS) void policy_head_kernel(
    PolicyDims d, const T* __restrict__ x, const float* __restrict__ ctx,
    const float* __restrict__ prm, float* __restrict__ waypoints, float* __restrict__ speed) {
  extern __shared__ __align__(16) float sm[];
  const int f0 = blockIdx.x * GATE_FT;
  const int in_dim = d.bd + d.ctx_dim;
  const int ld_c = (int)al4(d.Cf), ld_in = (int)al4(in_dim), ld_h = (int)al4(d.hid);
  float* s_pool = sm;                      // [FT][Cf]
  float* s_in = s_pool + GATE_FT * ld_c;   // [FT][bd+ctx]
  float* s_h1 = s_in + GATE_FT * ld_in;    // [FT][hid]
  float* s_h2 = s_h1 + GATE_FT * ld_h;     // [FT][hid]
  float* s_o = s_h2 + GATE_FT * ld_h;      // [FT][2*hz]
  const int ld_o = (int)al4(2 * d.hz);

  // AdaptiveAvgPool2d(1): channel c of frame f summed over pixels in order (thread per (f,c);
  // consecutive threads read consecutive channels -> coalesced)
  for (int i = threadIdx.x; i < GATE_FT * d.Cf; i += blockDim.x) {
    int f = i / d.Cf, c = i - f * d.Cf;
    float s = 0.f;
    if (f0 + f < d.B) {
      const T* xp = x + (int64_t)(f0 + f) * d.HW * d.Cf + c;
      for (int p = 0; p < d.HW; ++p) s += ld_as_float<T>(xp + (int64_t)p * d.Cf);
    }
    s_pool[f * ld_c + c] = s / (float)d.HW;
  }
  for (int i = threadIdx.x; i < GATE_FT * d.ctx_dim; i += blockDim.x) {
    int f = i / d.ctx_dim, c = i - f * d.ctx_dim;
    s_in[f * ld_in + d.bd + c] = (f0 + f < d.B) ? ctx[(int64_t)(f0 + f) * d.ctx_dim + c] : 0.f;
  }
  __syncthreads();

  const float* p = prm;
  const float* Wfc = p; p += al4((int64_t)d.bd * d.Cf);
  const float* bfc = p; p += al4(d.bd);
  linear_ft(Wfc, bfc, s_pool, ld_c, d.Cf, s_in, ld_in, d.bd, false);  // feat -> s_in[:, :bd]

  for (int head = 0; head < 2; ++head) {
    const int out_dim = head == 0 ? 2 * d.hz : d.hz;
    const float* W0 = p; p += al4((int64_t)d.hid * in_dim);
    const float* b0 = p; p += al4(d.hid);
    const float* W2 = p; p += al4((int64_t)d.hid * d.hid);
    const float* b2 = p; p += al4(d.hid);
    const float* W4 = p; p += al4((int64_t)out_dim * d.hid);
    const float* b4 = p; p += al4(out_dim);
    linear_ft(W0, b0, s_in, ld_in, in_dim, s_h1, ld_h, d.hid, true);
    linear_ft(W2, b2, s_h1, ld_h, d.hid, s_h2, ld_h, d.hid, true);
    linear_ft(W4, b4, s_h2, ld_h, d.hid, s_o, ld_o, out_dim, false);
    store_rows(head == 0 ? waypoints : speed, out_dim, s_o, ld_o, out_dim, f0, d.B);
    __syncthreads();
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
  size_t smem = sizeof(float) * GATE_FT *
                (al4(Cf) + al4(backbone_dim + ctx_dim) + 2 * al4(hidden) + al4(2 * horizon));
  AMOE_REQUIRE(smem <= 200 * 1024, "amoe_policy_head_fwd: dims too large for shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == AMOE_BF16) {
    if (smem > 48 * 1024)
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(policy_head_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    policy_head_kernel<__nv_bfloat16><<<ceil_div(B, GATE_FT), GATE_THREADS, smem, st>>>(
        d, (const __nv_bfloat16*)x, cvec, params, waypoints, speed);
  } else if (x_dtype == AMOE_F32) {
    if (smem > 48 * 1024)
      AMOE_CHECK_CUDA(cudaFuncSetAttribute(policy_head_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    policy_head_kernel<float><<<ceil_div(B, GATE_FT), GATE_THREADS, smem, st>>>(
        d, (const float*)x, cvec, params, waypoints, speed);
  } else {
    AMOE_REQUIRE(false, "amoe_policy_head_fwd: bad dtype %d", x_dtype);
  }
  AMOE_LAUNCH_OK(ctx);
  return 0;
}
