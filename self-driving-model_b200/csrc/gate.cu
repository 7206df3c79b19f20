// Fused gate kernel: SimpleContextExtractor + expert extractors (MLP+LayerNorm on the
// pooled expert logits) + GatingNetwork (context encoder, expert processors, gate MLP,
// softmax, weighted combine, output projection) in ONE launch, fp32 throughout so the
// top-1 routing matches the reference (SURVEY.md §7 "bit-exact top-1").
//
// One CTA owns FT frames; every weight row is streamed once per CTA with coalesced
// 16-byte loads and reused for the FT frames; reductions are warp shuffles.
//
// Flat parameter layout (fp32; every tensor starts on an 8-float boundary, weights are
// PyTorch [out,in] row-major) — must match models/automoe.py::_pack_gate_params:
//   ctx.enc0.W[32,4] b[32] | ctx.enc3.W[ctx,32] b[ctx] | ctx.ln.g[ctx] b[ctx]
//   per expert e: ext.W1[512,Ce] b1[512] | ext.W2[F,512] b2[F] | ext.ln.g[F] b[F]
//   gate.ctxenc0.W[hid,ctx] b | gate.ctxenc3.W[hid,hid] b
//   per expert e: proc.W0[P,F] b | proc.W3[P,P] b | proc.ln.g[P] b[P]
//   gate.net0.W[hid,hid+P*E] b | gate.net3.W[E,hid] b | out_proj.W[P,P] b
// with F = expert feature dim (256), P = processed dim (256).
#include <cstdlib>

#include "common.cuh"
#include "mlp.cuh"

constexpr int GATE_MAX_E = 4;
// mode bits (amoe_gate_fwd)
constexpr int GATE_MODE_CTX_ONLY = 1;   // get_expert_weights: experts replaced by zeros, weights only
constexpr int GATE_MODE_CTX_IN = 2;     // `state` holds an already-encoded context [B,ctx_dim]
constexpr int GATE_MODE_FEAT_IN = 4;    // `pooled` holds expert features [E][B,F] (extractors skipped)
constexpr int GATE_MODE_STOP_CTX = 8;   // stop after the context extractor
constexpr int GATE_MODE_STOP_FEAT = 16; // stop after the expert extractors
constexpr int GATE_MODE_SIGMOID = 32;   // use_softmax=False: sigmoid(logits) / (sum + 1e-8), gating_network.py:159-160
constexpr int EXT_HID = 512;     // expert_extractors.py:30,64,91

struct GateDims {
  int B, E, ctx_dim, hidden, F, P, sumC;
  int n_ch[GATE_MAX_E];
  float temperature;
  int mode;
  int split;   // tensor-core variant launched as clusters of E+1 CTAs per 16 frames (see gate_fused_kernel)
};

// shared-memory row strides (floats) of the per-CTA activation buffers; the tensor-core variant pads every row to
// 16 mod 32 floats so its 16-byte fragment loads are bank-conflict free
struct GateLds {
  int in, ctx, h, feat, gin, t;
  __host__ __device__ int64_t per_frame() const { return (int64_t)in + ctx + h + feat + gin + t + GATE_MAX_E; }
};
__host__ __device__ inline GateLds gate_lds(int sumC, int ctx_dim, int hidden, int F, int P, int E, bool tc) {
  GateLds l;
  const int gin = hidden + P * E, t = P > hidden ? P : hidden;
  l.in = (int)al8(4 + sumC);
  l.ctx = tc ? ld_tc(ctx_dim) : (int)al8(ctx_dim);
  l.h = tc ? ld_tc(EXT_HID) : EXT_HID;
  l.feat = tc ? ld_tc(F) : (int)al8(F);
  l.gin = tc ? ld_tc(gin) : (int)al8(gin);
  l.t = tc ? ld_tc(t) : (int)al8(t);
  return l;
}

template <int FT, bool CL, bool TC>
__global__ __launch_bounds__(GATE_THREADS) void gate_fused_kernel(
    GateDims d, const float* __restrict__ state, const float* __restrict__ pooled,
    const float* __restrict__ prm, const __nv_bfloat16* __restrict__ prm16, float* __restrict__ context,
    float* __restrict__ features,
    float* __restrict__ processed, float* __restrict__ gate_logits, float* __restrict__ weights,
    float* __restrict__ combined, const float* __restrict__ ext_feat) {
  extern __shared__ __align__(16) float sm[];
  // Tensor-core variant, d.split: the weight stream per CTA bounds this kernel and the E expert chains (extractor +
  // processor, 3/4 of the parameters) are independent, so a cluster of E+1 CTAs owns the 16 frames: rank e < E runs
  // expert e, rank E the context path; each writes its block of the gate input into rank 0's shared memory
  // (distributed shared memory), and after one cluster barrier rank 0 runs the gate MLP, softmax, combine and
  // projection.  Per-CTA weight stream: 1.03 M -> 0.26 M + 0.18 M parameters on the critical path.
  const bool split = TC && d.split;
  const int R = d.E + 1;
  const int rank = split ? (int)cg::this_cluster().block_rank() : 0;
  const int f0 = (CL ? blockIdx.x / CL_RANKS : (split ? blockIdx.x / R : blockIdx.x)) * FT;
  const bool do_ctx = !split || rank == d.E;      // context extractor + gating context encoder
  if (split) cg::this_cluster().sync();           // every CTA of the cluster is running before anyone writes into rank 0
  const int gin = d.hidden + d.P * d.E;  // gate_network input width
  // shared-memory carve-up (all row strides multiples of 4 floats)
  const GateLds L = gate_lds(d.sumC, d.ctx_dim, d.hidden, d.F, d.P, d.E, TC);
  const int ld_in = L.in, ld_ctx = L.ctx, ld_h = L.h, ld_feat = L.feat, ld_gin = L.gin, ld_t = L.t;
  float* s_in = sm;                      // [FT][ld_in]: state(4) | pooled(sumC)
  float* s_ctx = s_in + FT * ld_in;      // [FT][ctx]
  float* s_h = s_ctx + FT * ld_ctx;      // [FT][512] scratch hidden
  float* s_feat = s_h + FT * ld_h;       // [FT][F] current expert feature
  float* s_gin = s_feat + FT * ld_feat;  // [FT][gin]  ctxenc | processed_0..E-1
  float* s_t = s_gin + FT * ld_gin;      // [FT][max(P,hid)] scratch
  float* s_w = s_t + FT * ld_t;          // [FT][E] logits then weights
  // tensor-core variant: the bf16 copy of the parameter buffer has the same element offsets
  auto w16 = [&](const float* W) -> const __nv_bfloat16* { return (TC && prm16) ? prm16 + (W - prm) : nullptr; };

  const bool ctx_in = d.mode & GATE_MODE_CTX_IN, feat_in = d.mode & GATE_MODE_FEAT_IN;
  const bool ctx_only = d.mode & GATE_MODE_CTX_ONLY;
  for (int i = threadIdx.x; i < FT * ld_in; i += blockDim.x) {
    int f = i / ld_in, c = i - f * ld_in;
    float v = 0.f;
    if (f0 + f < d.B) {
      if (c < 4) v = ctx_in ? 0.f : state[(int64_t)(f0 + f) * 4 + c];
      else if (c < 4 + d.sumC && !ctx_only && !feat_in && !(d.mode & GATE_MODE_STOP_CTX)) v = pooled[(int64_t)(f0 + f) * d.sumC + (c - 4)];
    }
    s_in[i] = v;
  }
  if (ctx_in) {
    for (int i = threadIdx.x; i < FT * d.ctx_dim; i += blockDim.x) {
      int f = i / d.ctx_dim, c = i - f * d.ctx_dim;
      s_ctx[f * ld_ctx + c] = (f0 + f < d.B) ? state[(int64_t)(f0 + f) * d.ctx_dim + c] : 0.f;
    }
  }
  __syncthreads();   // CTA-local phase (every rank works on its own copy)

  // ---- SimpleContextExtractor (context_features.py:143-165) ----
  const float* p = prm;
  {
    const float* W0 = p; p += al8(32 * 4);
    const float* b0 = p; p += al8(32);
    const float* W3 = p; p += al8((int64_t)d.ctx_dim * 32);
    const float* b3 = p; p += al8(d.ctx_dim);
    const float* g = p; p += al8(d.ctx_dim);
    const float* bb = p; p += al8(d.ctx_dim);
    if (!ctx_in && do_ctx) {
      linear_ft<FT, CL, TC>(W0, b0, s_in, ld_in, 4, s_h, ld_h, 32, true, w16(W0));
      linear_ft<FT, CL, TC>(W3, b3, s_h, ld_h, 32, s_ctx, ld_ctx, d.ctx_dim, false, w16(W3));
      layernorm_ft<FT, CL>(s_ctx, ld_ctx, d.ctx_dim, g, bb);
    }
    if (context && do_ctx) store_rows<FT, CL>(context, d.ctx_dim, s_ctx, ld_ctx, d.ctx_dim, f0, d.B);
    if (d.mode & GATE_MODE_STOP_CTX) return;
  }

  // ---- expert extractors (expert_extractors.py:27-35) -> features; kept in global, the
  //      processors below re-read them from smem one expert at a time ----
  const float* ext_prm[GATE_MAX_E];
  for (int e = 0; e < d.E; ++e) {
    ext_prm[e] = p;
    p += al8((int64_t)EXT_HID * d.n_ch[e]) + al8(EXT_HID) + al8((int64_t)d.F * EXT_HID) + al8(d.F) + 2 * al8(d.F);
  }
  // gating context encoder (gating_network.py:12-20)
  {
    const float* W0 = p; p += al8((int64_t)d.hidden * d.ctx_dim);
    const float* b0 = p; p += al8(d.hidden);
    const float* W3 = p; p += al8((int64_t)d.hidden * d.hidden);
    const float* b3 = p; p += al8(d.hidden);
    if (!(d.mode & GATE_MODE_STOP_FEAT) && do_ctx) {
      linear_ft<FT, CL, TC>(W0, b0, s_ctx, ld_ctx, d.ctx_dim, s_t, ld_t, d.hidden, true, w16(W0));
      linear_ft<FT, CL, TC>(W3, b3, s_t, ld_t, d.hidden, s_gin, ld_gin, d.hidden, true, w16(W3));
      if (split) {   // encoded context -> rank 0's gate input, columns [0, hidden)
        float* dst = cg::this_cluster().map_shared_rank(s_gin, 0);
        const int n4 = d.hidden >> 2;
        for (int i = threadIdx.x; i < FT * n4; i += blockDim.x) {
          const int f = i / n4, c4 = i - f * n4;
          *reinterpret_cast<float4*>(dst + f * ld_gin + (c4 << 2)) = *reinterpret_cast<const float4*>(s_gin + f * ld_gin + (c4 << 2));
        }
      }
    }
  }
  int ch_off = 4;
  for (int e = 0; e < d.E; ++e) {
    const float* q = ext_prm[e];
    const float* W1 = q; q += al8((int64_t)EXT_HID * d.n_ch[e]);
    const float* b1 = q; q += al8(EXT_HID);
    const float* W2 = q; q += al8((int64_t)d.F * EXT_HID);
    const float* b2 = q; q += al8(d.F);
    const float* g = q; q += al8(d.F);
    const float* bb = q;
    const float* PW0 = p; p += al8((int64_t)d.P * d.F);
    const float* Pb0 = p; p += al8(d.P);
    const float* PW3 = p; p += al8((int64_t)d.P * d.P);
    const float* Pb3 = p; p += al8(d.P);
    const float* Pg = p; p += al8(d.P);
    const float* Pbb = p; p += al8(d.P);
    float* s_proc = s_gin + d.hidden + e * d.P;
    if (split && rank != e) {   // another rank's expert (split implies mode == 0)
      ch_off += d.n_ch[e];
      continue;
    }
    if (!ctx_only) {
      if (feat_in || d.n_ch[e] == 0) {
        // feature of this expert computed outside (all of them: GATE_MODE_FEAT_IN; or this one only: n_ch[e] == 0, an
        // extractor whose input does not fit this kernel's shared-memory rows, e.g. the nuScenes expert's Q*(C+D) vector)
        const float* src = feat_in ? pooled : ext_feat;
        for (int i = threadIdx.x; i < FT * d.F; i += blockDim.x) {
          int f = i / d.F, c = i - f * d.F;
          s_feat[f * ld_feat + c] = (f0 + f < d.B) ? src[((int64_t)e * d.B + f0 + f) * d.F + c] : 0.f;
        }
        __syncthreads();   // CTA-local phase (every rank works on its own copy)
        if (!feat_in && features) store_rows<FT, CL>(features + (int64_t)e * d.B * d.F, d.F, s_feat, ld_feat, d.F, f0, d.B);
      } else {
        linear_ft<FT, CL, TC>(W1, b1, s_in + ch_off, ld_in, d.n_ch[e], s_h, ld_h, EXT_HID, true, w16(W1));
        linear_ft<FT, CL, TC>(W2, b2, s_h, ld_h, EXT_HID, s_feat, ld_feat, d.F, false, w16(W2));
        layernorm_ft<FT, CL>(s_feat, ld_feat, d.F, g, bb);
        if (features) store_rows<FT, CL>(features + (int64_t)e * d.B * d.F, d.F, s_feat, ld_feat, d.F, f0, d.B);
      }
      if (!(d.mode & GATE_MODE_STOP_FEAT)) {
        // ExpertOutputProcessor (gating_network.py:37-43)
        linear_ft<FT, CL, TC>(PW0, Pb0, s_feat, ld_feat, d.F, s_t, ld_t, d.P, true, w16(PW0));
        linear_ft<FT, CL, TC>(PW3, Pb3, s_t, ld_t, d.P, s_proc, ld_gin, d.P, false, w16(PW3));
        layernorm_ft<FT, CL>(s_proc, ld_gin, d.P, Pg, Pbb);
        if (processed) store_rows<FT, CL>(processed + (int64_t)e * d.B * d.P, d.P, s_proc, ld_gin, d.P, f0, d.B);
        if (split && rank != 0) {   // processed_e -> rank 0's gate input, columns [hidden + e*P, hidden + (e+1)*P)
          float* dst = cg::this_cluster().map_shared_rank(s_gin, 0) + d.hidden + e * d.P;
          const int n4 = d.P >> 2;
          for (int i = threadIdx.x; i < FT * n4; i += blockDim.x) {
            const int f = i / n4, c4 = i - f * n4;
            *reinterpret_cast<float4*>(dst + f * ld_gin + (c4 << 2)) = *reinterpret_cast<const float4*>(s_proc + f * ld_gin + (c4 << 2));
          }
        }
      }
    } else {
      // get_expert_weights (gating_network.py:177-199): zeros stand in for the experts
      for (int i = threadIdx.x; i < FT * d.P; i += blockDim.x) s_proc[(i / d.P) * ld_gin + (i % d.P)] = 0.f;
      __syncthreads();   // CTA-local phase (every rank works on its own copy)
    }
    ch_off += d.n_ch[e];
  }

  if (d.mode & GATE_MODE_STOP_FEAT) return;
  if (split) {
    cg::this_cluster().sync();   // every rank's block of the gate input has landed in rank 0's shared memory
    if (rank != 0) return;
  }

  // ---- gate MLP + softmax (gating_network.py:94-99,141-160) ----
  {
    const float* W0 = p; p += al8((int64_t)d.hidden * gin);
    const float* b0 = p; p += al8(d.hidden);
    const float* W3 = p; p += al8((int64_t)d.E * d.hidden);
    const float* b3 = p; p += al8(d.E);
    linear_ft<FT, CL, TC>(W0, b0, s_gin, ld_gin, gin, s_t, ld_t, d.hidden, true, w16(W0));
    linear_ft<FT, CL, TC>(W3, b3, s_t, ld_t, d.hidden, s_w, GATE_MAX_E, d.E, false, w16(W3));
    if (threadIdx.x < FT) {
      int f = threadIdx.x;
      float* lg = s_w + f * GATE_MAX_E;
      float mx = -INFINITY;
      const bool sig = d.mode & GATE_MODE_SIGMOID;
      for (int e = 0; e < d.E; ++e) {
        if (gate_logits && f0 + f < d.B) gate_logits[(int64_t)(f0 + f) * d.E + e] = lg[e];
        if (!sig) lg[e] = lg[e] / d.temperature;
        mx = fmaxf(mx, lg[e]);
      }
      float ssum = 0.f;
      for (int e = 0; e < d.E; ++e) {
        lg[e] = sig ? 1.f / (1.f + expf(-lg[e])) : expf(lg[e] - mx);
        ssum += lg[e];
      }
      if (sig) ssum += 1e-8f;
      for (int e = 0; e < d.E; ++e) {
        lg[e] = lg[e] / ssum;
        if (weights && f0 + f < d.B) weights[(int64_t)(f0 + f) * d.E + e] = lg[e];
      }
    }
    __syncthreads();   // CTA-local phase (every rank works on its own copy)
  }
  if (ctx_only) return;

  // ---- weighted combine + output projection (gating_network.py:162-168) ----
  {
    const float* W = p; p += al8((int64_t)d.P * d.P);
    const float* b = p;
    for (int i = threadIdx.x; i < FT * d.P; i += blockDim.x) {
      int f = i / d.P, c = i - f * d.P;
      float acc = 0.f;  // combined_output starts at zeros and adds w_e * processed_e in order
      for (int e = 0; e < d.E; ++e) acc += s_w[f * GATE_MAX_E + e] * s_gin[f * ld_gin + d.hidden + e * d.P + c];
      s_h[f * ld_h + c] = acc;
    }
    __syncthreads();   // CTA-local phase (every rank works on its own copy)
    linear_ft<FT, CL, TC>(W, b, s_h, ld_h, d.P, s_t, ld_t, d.P, false, w16(W));
    if (combined) store_rows<FT, CL>(combined, d.P, s_t, ld_t, d.P, f0, d.B);
  }
}

static int64_t gate_param_count(const GateDims& d) {
  int64_t n = al8(32 * 4) + al8(32) + al8((int64_t)d.ctx_dim * 32) + 3 * al8(d.ctx_dim);
  for (int e = 0; e < d.E; ++e)
    n += al8((int64_t)EXT_HID * d.n_ch[e]) + al8(EXT_HID) + al8((int64_t)d.F * EXT_HID) + 3 * al8(d.F);
  n += al8((int64_t)d.hidden * d.ctx_dim) + al8(d.hidden) + al8((int64_t)d.hidden * d.hidden) + al8(d.hidden);
  for (int e = 0; e < d.E; ++e)
    n += al8((int64_t)d.P * d.F) + al8(d.P) + al8((int64_t)d.P * d.P) + 3 * al8(d.P);
  int gin = d.hidden + d.P * d.E;
  n += al8((int64_t)d.hidden * gin) + al8(d.hidden) + al8((int64_t)d.E * d.hidden) + al8(d.E);
  n += al8((int64_t)d.P * d.P) + al8(d.P);
  return n;
}

extern "C" int amoe_gate_fwd_ex2(amoe_ctx* ctx, const float* state, const float* pooled,
                                 const float* params, const void* params_bf16, int64_t n_params, int B, int E,
                                 const int* n_ch_host, int ctx_dim, int hidden, float temperature,
                                 int mode, const float* ext_features, float* context, float* features, float* processed,
                                 float* gate_logits, float* weights, float* combined, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && state && params && n_ch_host, "amoe_gate_fwd: NULL argument");
  AMOE_REQUIRE(E >= 1 && E <= GATE_MAX_E, "amoe_gate_fwd: E=%d out of range [1,%d]", E, GATE_MAX_E);
  AMOE_REQUIRE((mode & (GATE_MODE_CTX_ONLY | GATE_MODE_STOP_CTX)) || pooled, "amoe_gate_fwd: pooled is NULL");
  AMOE_REQUIRE(temperature > 0.f, "amoe_gate_fwd: temperature must be > 0");
  GateDims d;
  d.B = B; d.E = E; d.ctx_dim = ctx_dim; d.hidden = hidden; d.F = 256; d.P = 256;
  d.temperature = temperature; d.mode = mode; d.sumC = 0; d.split = 0;
  for (int e = 0; e < GATE_MAX_E; ++e) d.n_ch[e] = 0;
  for (int e = 0; e < E; ++e) {
    AMOE_REQUIRE(n_ch_host[e] >= 1 || (n_ch_host[e] == 0 && ext_features != nullptr),
                 "amoe_gate_fwd: n_ch[%d]=%d (0 = feature supplied in ext_features)", e, n_ch_host[e]);
    d.n_ch[e] = n_ch_host[e];
    d.sumC += n_ch_host[e];
  }
  AMOE_REQUIRE(hidden <= EXT_HID && ctx_dim <= EXT_HID && hidden >= 1 && ctx_dim >= 1,
               "amoe_gate_fwd: hidden/ctx_dim must be in [1,512]");
  int64_t need = gate_param_count(d);
  AMOE_REQUIRE(n_params == need, "amoe_gate_fwd: params has %lld floats, layout needs %lld",
               (long long)n_params, (long long)need);
  if (B == 0) return 0;
  const __nv_bfloat16* p16 = (const __nv_bfloat16*)params_bf16;
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(params_bf16) & 15) == 0, "amoe_gate_fwd: params_bf16 must be 16-byte aligned");
  if (p16 && B >= MMA_FT) {
    // bf16 inference mode: 16 frames per CTA, layers with K % 32 == 0 run on mma.sync TF32 (mlp.cuh)
    const size_t smem = sizeof(float) * (size_t)gate_lds(d.sumC, ctx_dim, hidden, d.F, d.P, E, true).per_frame() * MMA_FT;
    AMOE_REQUIRE(smem <= 226 * 1024, "amoe_gate_fwd: dims too large for shared memory");
    auto kern = gate_fused_kernel<MMA_FT, false, true>;
    AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* e = getenv("AMOE_GATE_SPLIT");
    d.split = ((mode & ~GATE_MODE_SIGMOID) == 0 && E <= 3 && hidden % 4 == 0 && (e == nullptr || atoi(e) != 0)) ? 1 : 0;
    if (d.split) {
      // the forward of AutoMoE: clusters of E+1 CTAs per 16 frames (expert chains and the context path in parallel)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ceil_div(B, MMA_FT) * (E + 1));
      cfg.blockDim = dim3(GATE_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = E + 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      AMOE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, d, state, pooled, params, p16, context, features, processed, gate_logits,
                                         weights, combined, ext_features));
    } else {
      kern<<<ceil_div(B, MMA_FT), GATE_THREADS, smem, (cudaStream_t)stream>>>(
          d, state, pooled, params, p16, context, features, processed, gate_logits, weights, combined, ext_features);
    }
    AMOE_LAUNCH_OK(ctx);
    return 0;
  }
  const size_t per_frame = sizeof(float) * (size_t)gate_lds(d.sumC, ctx_dim, hidden, d.F, d.P, E, false).per_frame();
  if (mlp_use_cluster(B) && per_frame * CL_FT <= 200 * 1024) {
    // large batch: clusters of 8 CTAs split every layer's output rows (each CTA streams 1/8 of the weights)
    const size_t smem = per_frame * CL_FT;
    auto kern = gate_fused_kernel<CL_FT, true, false>;
    AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(B, CL_FT) * CL_RANKS);
    cfg.blockDim = dim3(GATE_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL_RANKS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AMOE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, d, state, pooled, params, (const __nv_bfloat16*)nullptr, context, features, processed, gate_logits,
                                       weights, combined, ext_features));
    AMOE_LAUNCH_OK(ctx);
    return 0;
  }
  const size_t smem = per_frame * GATE_FT;
  auto kern = gate_fused_kernel<GATE_FT, false, false>;
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<ceil_div(B, GATE_FT), GATE_THREADS, smem, (cudaStream_t)stream>>>(
      d, state, pooled, params, (const __nv_bfloat16*)nullptr, context, features, processed, gate_logits, weights, combined, ext_features);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int amoe_gate_fwd_ex(amoe_ctx* ctx, const float* state, const float* pooled,
                                const float* params, const void* params_bf16, int64_t n_params, int B, int E,
                                const int* n_ch_host, int ctx_dim, int hidden, float temperature,
                                int mode, float* context, float* features, float* processed,
                                float* gate_logits, float* weights, float* combined, void* stream) {
  return amoe_gate_fwd_ex2(ctx, state, pooled, params, params_bf16, n_params, B, E, n_ch_host, ctx_dim, hidden, temperature, mode,
                           nullptr, context, features, processed, gate_logits, weights, combined, stream);
}

extern "C" int amoe_gate_fwd(amoe_ctx* ctx, const float* state, const float* pooled,
                             const float* params, int64_t n_params, int B, int E,
                             const int* n_ch_host, int ctx_dim, int hidden, float temperature,
                             int mode, float* context, float* features, float* processed,
                             float* gate_logits, float* weights, float* combined, void* stream) {
  AMOE_ENTER(ctx);
  return amoe_gate_fwd_ex(ctx, state, pooled, params, nullptr, n_params, B, E, n_ch_host, ctx_dim, hidden, temperature, mode,
                          context, features, processed, gate_logits, weights, combined, stream);
}
