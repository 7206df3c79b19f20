// Training-step kernels of the gating / policy MLPs (forward in train mode,
// backward, losses, clip + AdamW).  Everything is fp32 on the CUDA cores: the trainable part of
// AutoMoE is 2.87 M parameters and a per-GPU batch of 32, i.e. launch- and latency-bound, not a
// tensor-core workload; fp32 also keeps gradients within 1e-4 of the reference's autograd.
//
//   amoe_linear_fwd / amoe_linear_bwd      nn.Linear (+ReLU, +Dropout) forward and its three gradients
//   amoe_layernorm_fwd / _bwd              nn.LayerNorm over the last dim
//   amoe_gate_combine_fwd / _bwd           softmax(logits/T) and the weighted sum of processed experts
//   amoe_gating_loss_fwd_bwd               compute_gating_losses (train_gating_network.py:21-74) + d/d(pred)
//   amoe_sq_norm / amoe_fused_clip_adamw   clip_grad_norm_ + AdamW on a flat parameter buffer
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// C[M,N] = opA(A)[M,K] * opB(B)[K,N]   (row-major with leading dimensions, fp32)
//   TA = false: A(m,k) = A[m*lda + k]   TA = true: A(m,k) = A[k*lda + m]
//   TB = false: B(k,n) = B[k*ldb + n]   TB = true: B(k,n) = B[n*ldb + k]
// Epilogue (optional): + bias[n], ReLU, inverted dropout keyed by (seed, element index).
constexpr int GBM = 64, GBN = 64, GBK = 16;

__device__ __forceinline__ float u01_hash(uint64_t seed, uint64_t idx) {
  // splitmix64 finaliser: one independent uniform per (seed, element)
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

template <bool TA, bool TB>
__global__ __launch_bounds__(256) void sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ C, int M, int N, int K, int lda, int ldb,
                                                    int ldc, const float* __restrict__ bias, int relu, float drop_p,
                                                    uint64_t seed, const uint64_t* __restrict__ seed_dev) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GBK) {
    // 64x16 elements of each operand, 4 per thread; the index split keeps the contiguous direction on lanes
    for (int e = tid; e < GBM * GBK; e += 256) {
      int r, k;
      if (TA) { r = e % GBM; k = e / GBM; } else { k = e % GBK; r = e / GBK; }
      const int m = m0 + r, kk = k0 + k;
      float v = 0.f;
      if (m < M && kk < K) v = TA ? A[(int64_t)kk * lda + m] : A[(int64_t)m * lda + kk];
      As[k][r] = v;
    }
    for (int e = tid; e < GBN * GBK; e += 256) {
      int c, k;
      if (TB) { k = e % GBK; c = e / GBK; } else { c = e % GBN; k = e / GBN; }
      const int n = n0 + c, kk = k0 + k;
      float v = 0.f;
      if (n < N && kk < K) v = TB ? B[(int64_t)n * ldb + kk] : B[(int64_t)kk * ldb + n];
      Bs[k][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  if (drop_p > 0.f && seed_dev) seed += *seed_dev;     // per-step part of the key, advanced on the device (amoe_train_tick)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      if (drop_p > 0.f) v = (u01_hash(seed, (uint64_t)m * N + n) >= drop_p) ? v * keep_scale : 0.f;
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same product for the shapes the gating / policy MLPs have in training: M = the per-GPU batch (32), N and K a few
// hundred.  The 64x64 tiles above leave such a launch with N/64 CTAs that each walk all of K behind two barriers per 16
// columns (58 us for 32x768 -> 512).  Here a CTA owns SK_TN output columns of 32 rows, lane = row, and the reduction axis
// is split over its eight warps: every warp stages its slice of B once (coalesced), every lane streams its own row of A
// with 16-byte loads, and the eight partial sums meet in shared memory in a fixed order (deterministic).  N/8 CTAs, one
// round trip to memory each.  TA = false only (the weight gradient has M = out_dim and stays on the tiled kernel).
constexpr int SK_TN = 8, SK_WARPS = 8, SK_KC = 128;

template <bool TB>
__global__ __launch_bounds__(SK_WARPS * 32) void skinny_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                    float* __restrict__ C, int M, int N, int K, int lda,
                                                                    int ldb, int ldc, const float* __restrict__ bias, int relu,
                                                                    float drop_p, uint64_t seed,
                                                                    const uint64_t* __restrict__ seed_dev, int a_vec) {
  __shared__ __align__(16) float Bs[SK_WARPS][SK_KC][SK_TN];
  __shared__ float red[SK_WARPS][32 * SK_TN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * SK_TN, m0 = blockIdx.y * 32;
  const int m_ld = min(m0 + lane, M - 1);                       // rows past M read the last row and store nothing
  const float* __restrict__ arow = A + (int64_t)m_ld * lda;
  const int kslice = (((K + 3) >> 2) + SK_WARPS - 1) / SK_WARPS * 4;   // multiple of 4: 16-byte loads never straddle slices
  const int kb = warp * kslice, ke = min(K, kb + kslice);
  float acc[SK_TN];
#pragma unroll
  for (int j = 0; j < SK_TN; ++j) acc[j] = 0.f;
  for (int kc = kb; kc < ke; kc += SK_KC) {
    const int len = min(SK_KC, ke - kc);
    const int len4 = (len + 3) & ~3;
    // this warp's [len x SK_TN] block of B -> Bs[r][j]; rows past len are zero so the 4-wide steps below may run over
    if (TB) {
#pragma unroll
      for (int j = 0; j < SK_TN; ++j) {
        const bool col = n0 + j < N;
        const float* __restrict__ brow = B + (int64_t)(n0 + j) * ldb + kc;
        for (int r = lane; r < len4; r += 32) Bs[warp][r][j] = (col && r < len) ? __ldg(brow + r) : 0.f;
      }
    } else {
      for (int e = lane; e < len4 * SK_TN; e += 32) {
        const int r = e / SK_TN, j = e % SK_TN;
        (&Bs[warp][0][0])[e] = (r < len && n0 + j < N) ? __ldg(B + (int64_t)(kc + r) * ldb + n0 + j) : 0.f;
      }
    }
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < len4; r += 4) {
      float a[4];
      if (a_vec && kc + r + 3 < K) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(arow + kc + r));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = (kc + r + q < K) ? __ldg(arow + kc + r + q) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[warp][r + q][0]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[warp][r + q][4]);
        acc[0] = fmaf(a[q], b0.x, acc[0]); acc[1] = fmaf(a[q], b0.y, acc[1]);
        acc[2] = fmaf(a[q], b0.z, acc[2]); acc[3] = fmaf(a[q], b0.w, acc[3]);
        acc[4] = fmaf(a[q], b1.x, acc[4]); acc[5] = fmaf(a[q], b1.y, acc[5]);
        acc[6] = fmaf(a[q], b1.z, acc[6]); acc[7] = fmaf(a[q], b1.w, acc[7]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < SK_TN; ++j) red[warp][lane * SK_TN + j] = acc[j];
  __syncthreads();
  // thread t finishes output (row t / 8, column t % 8): a row's eight columns are one 32-byte store segment
  const int t = threadIdx.x;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < SK_WARPS; ++w) v += red[w][t];
  const int m = m0 + t / SK_TN, n = n0 + t % SK_TN;
  if (m >= M || n >= N) return;
  if (bias) v += bias[n];
  if (relu) v = fmaxf(v, 0.f);
  if (drop_p > 0.f) {
    if (seed_dev) seed += *seed_dev;
    v = (u01_hash(seed, (uint64_t)m * N + n) >= drop_p) ? v * (1.f / (1.f - drop_p)) : 0.f;
  }
  C[(int64_t)m * ldc + n] = v;
}

// g = dy * [y > 0] * keep_scale  (gradient through Dropout(ReLU(.)): a kept, active unit has y > 0)
__global__ void mask_grad_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                 float* __restrict__ g, int B, int N, float keep_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, n = i - b * N;
  g[i] = y[(int64_t)b * ldy + n] > 0.f ? dy[(int64_t)b * lddy + n] * keep_scale : 0.f;
}

// out[n] = sum_b g[b*ld + n]   (fixed order: deterministic)
__global__ void colsum_kernel(const float* __restrict__ g, int ld, float* __restrict__ out, int B, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += g[(int64_t)b * ld + n];
  out[n] = s;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension: one warp per row.
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean,
                                     float* __restrict__ rstd, int B, int D, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + (int64_t)row * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += xr[i];
  const float mu = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float d = xr[i] - mu;
    v = fmaf(d, d, v);
  }
  const float rs = rsqrtf(warp_sum(v) / (float)D + eps);
  for (int i = lane; i < D; i += 32) y[(int64_t)row * D + i] = (xr[i] - mu) * rs * gamma[i] + beta[i];
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx = rstd * (dy*gamma - mean_D(dy*gamma) - xhat * mean_D(dy*gamma*xhat)); one warp per row
__global__ void layernorm_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, float* __restrict__ dx, int B, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float mu = mean[row], rs = rstd[row];
  const float* xr = x + (int64_t)row * D;
  const float* dr = dy + (int64_t)row * D;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float gd = dr[i] * gamma[i], xh = (xr[i] - mu) * rs;
    s1 += gd;
    s2 = fmaf(gd, xh, s2);
  }
  s1 = warp_sum(s1) / (float)D;
  s2 = warp_sum(s2) / (float)D;
  for (int i = lane; i < D; i += 32) {
    const float gd = dr[i] * gamma[i], xh = (xr[i] - mu) * rs;
    dx[(int64_t)row * D + i] = rs * (gd - s1 - xh * s2);
  }
}
// dgamma[i] = sum_b dy*xhat, dbeta[i] = sum_b dy   (thread per column, fixed order)
__global__ void layernorm_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  float sg = 0.f, sb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = dy[(int64_t)b * D + i];
    sg = fmaf(d, (x[(int64_t)b * D + i] - mean[b]) * rstd[b], sg);
    sb += d;
  }
  dgamma[i] = sg;
  dbeta[i] = sb;
}

// ------------------------------------------------------------------------------------------------
// Gate: weights = softmax(logits / T); combined = sum_e weights[:,e] * processed_e  (gating_network.py:157-165)
constexpr int MAX_E = 4;
__global__ void gate_combine_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ processed,
                                        int64_t e_stride, int ld_p, float inv_T, float* __restrict__ weights,
                                        float* __restrict__ combined, int B, int E, int P, int sigmoid_gate) {
  const int b = blockIdx.x;
  __shared__ float w[MAX_E];
  if (threadIdx.x == 0) {
    float mx = -INFINITY, z[MAX_E], s = 0.f;
    if (sigmoid_gate) {   // gating_network.py:159-160: sigmoid(logits) / (sum + 1e-8), no temperature
      for (int e = 0; e < E; ++e) {
        z[e] = 1.f / (1.f + expf(-logits[(int64_t)b * E + e]));
        s += z[e];
      }
      s += 1e-8f;
    } else {
      for (int e = 0; e < E; ++e) {
        z[e] = logits[(int64_t)b * E + e] * inv_T;
        mx = fmaxf(mx, z[e]);
      }
      for (int e = 0; e < E; ++e) {
        z[e] = expf(z[e] - mx);
        s += z[e];
      }
    }
    for (int e = 0; e < E; ++e) {
      w[e] = z[e] / s;
      weights[(int64_t)b * E + e] = w[e];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < P; c += blockDim.x) {
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc += w[e] * processed[e * e_stride + (int64_t)b * ld_p + c];
    combined[(int64_t)b * P + c] = acc;
  }
}
// dprocessed_e = w_e * dcombined; dw_e = <dcombined, processed_e> + dweights_e;
// dlogits = inv_T * w * (dw - sum_e w_e dw_e)            (softmax gate)
// dlogits = s(1-s) * (dw - sum_e w_e dw_e) / (sum s + 1e-8)  (sigmoid gate, s = sigmoid(logits))
__global__ void gate_combine_bwd_kernel(const float* __restrict__ dcombined, const float* __restrict__ dweights,
                                        const float* __restrict__ weights, const float* __restrict__ processed,
                                        int64_t e_stride, int ld_p, float inv_T, float* __restrict__ dlogits,
                                        float* __restrict__ dprocessed, int64_t de_stride, int B, int E, int P,
                                        const float* __restrict__ logits) {
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  __shared__ float part[MAX_E][8];
  float dot[MAX_E];
#pragma unroll
  for (int e = 0; e < MAX_E; ++e) dot[e] = 0.f;
  for (int c = threadIdx.x; c < P; c += blockDim.x) {
    const float dc = dcombined ? dcombined[(int64_t)b * P + c] : 0.f;
#pragma unroll
    for (int e = 0; e < MAX_E; ++e) {
      if (e < E) {
        dot[e] = fmaf(dc, processed[e * e_stride + (int64_t)b * ld_p + c], dot[e]);
        dprocessed[e * de_stride + (int64_t)b * P + c] = weights[(int64_t)b * E + e] * dc;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < MAX_E; ++e) {
    dot[e] = warp_sum(dot[e]);
    if (lane == 0) part[e][warp] = dot[e];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float dw[MAX_E], mix = 0.f;
    for (int e = 0; e < E; ++e) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += part[e][w];
      dw[e] = s + (dweights ? dweights[(int64_t)b * E + e] : 0.f);
      mix = fmaf(weights[(int64_t)b * E + e], dw[e], mix);
    }
    if (logits != nullptr) {
      float sg[MAX_E], S = 0.f;
      for (int e = 0; e < E; ++e) {
        sg[e] = 1.f / (1.f + expf(-logits[(int64_t)b * E + e]));
        S += sg[e];
      }
      S += 1e-8f;
      for (int e = 0; e < E; ++e) dlogits[(int64_t)b * E + e] = sg[e] * (1.f - sg[e]) * (dw[e] - mix) / S;
    } else {
      for (int e = 0; e < E; ++e) dlogits[(int64_t)b * E + e] = inv_T * weights[(int64_t)b * E + e] * (dw[e] - mix);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compute_gating_losses (training/train_gating_network.py:21-74), forward value of every term and the
// gradient of total_loss w.r.t. waypoints, speed and expert_weights in one single-CTA kernel.
//   out[0..6] = total, ade, fde, speed, smoothness, load_balancing, entropy_loss
//   coef[0..5] = ade, fde, speed, smoothness, load_balancing, entropy weights
//   speed_mode: 0 = no speed term, 1 = full sequence [B,H] vs [B,H], 2 = last step only
struct LossArgs {
  const float* wp; const float* spd; const float* ew; const float* twp; const float* tspd;
  float* out; float* dwp; float* dspd; float* dew;
  int B, H, E, speed_mode, spd_ld, tspd_ld, use_lb, use_ent;
  float coef[6];
};
__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) - (v < 0.f); }
__device__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];   // fixed order on every thread
  return s;
}
__global__ __launch_bounds__(1024) void gating_loss_kernel(LossArgs a) {
  __shared__ float red[32];
  __shared__ float usage[MAX_E];
  const int B = a.B, H = a.H, E = a.E;
  const int n_wp = B * H * 2;
  // ---- ADE / FDE ----
  float s_ade = 0.f, s_fde = 0.f;
  for (int i = threadIdx.x; i < n_wp; i += blockDim.x) {
    const float d = a.wp[i] - a.twp[i];
    s_ade += fabsf(d);
    if ((i / 2) % H == H - 1) s_fde += fabsf(d);
  }
  const float ade = block_sum_1024(s_ade, red) / (float)n_wp;
  const float fde = block_sum_1024(s_fde, red) / (float)(B * 2);
  // ---- speed ----
  float s_spd = 0.f;
  const int n_spd = a.speed_mode == 1 ? B * H : (a.speed_mode == 2 ? B : 0);
  for (int i = threadIdx.x; i < n_spd; i += blockDim.x) {
    float d;
    if (a.speed_mode == 1) d = a.spd[(i / H) * a.spd_ld + (i % H)] - a.tspd[(i / H) * a.tspd_ld + (i % H)];
    else d = a.spd[i * a.spd_ld + a.spd_ld - 1] - a.tspd[i * a.tspd_ld + a.tspd_ld - 1];
    s_spd += fabsf(d);
  }
  const float spd = n_spd ? block_sum_1024(s_spd, red) / (float)n_spd : 0.f;
  // ---- smoothness: L1 between consecutive deltas = |second difference| over t = 0..H-3 ----
  const int n_sm = B * (H - 2) * 2;
  float s_sm = 0.f;
  for (int i = threadIdx.x; i < n_sm; i += blockDim.x) {
    const int c = i & 1, t = (i >> 1) % (H - 2), b = (i >> 1) / (H - 2);
    const float* w = a.wp + ((int64_t)b * H + t) * 2 + c;
    s_sm += fabsf((w[4] - w[2]) - (w[2] - w[0]));
  }
  const float sm = n_sm > 0 ? block_sum_1024(s_sm, red) / (float)n_sm : 0.f;
  // ---- load balancing: mse(mean_b weights, 1/E) ----
  float lb = 0.f;
  if (a.use_lb) {
    for (int e = 0; e < E; ++e) {
      float s = 0.f;
      for (int b = threadIdx.x; b < B; b += blockDim.x) s += a.ew[(int64_t)b * E + e];
      s = block_sum_1024(s, red) / (float)B;
      if (threadIdx.x == 0) usage[e] = s;
      const float d = s - 1.f / (float)E;
      lb += d * d;
    }
    lb /= (float)E;
  }
  __syncthreads();
  // ---- entropy_loss = -mean_b( -sum_e w log(w + 1e-8) ) ----
  float ent = 0.f;
  if (a.use_ent) {
    float s = 0.f;
    for (int i = threadIdx.x; i < B * E; i += blockDim.x) s += a.ew[i] * logf(a.ew[i] + 1e-8f);
    ent = block_sum_1024(s, red) / (float)B;   // = -entropy
  }
  if (threadIdx.x == 0) {
    a.out[1] = ade; a.out[2] = fde; a.out[3] = spd; a.out[4] = sm; a.out[5] = lb; a.out[6] = ent;
    a.out[0] = a.coef[0] * ade + a.coef[1] * fde + a.coef[2] * spd + a.coef[3] * sm + a.coef[4] * lb + a.coef[5] * ent;
  }
  // ---- gradients of total ----
  if (a.dwp) {
    const float k_ade = a.coef[0] / (float)n_wp, k_fde = a.coef[1] / (float)(B * 2);
    const float k_sm = n_sm > 0 ? a.coef[3] / (float)n_sm : 0.f;
    for (int i = threadIdx.x; i < n_wp; i += blockDim.x) {
      const int c = i & 1, t = (i >> 1) % H, b = (i >> 1) / H;
      const float d = a.wp[i] - a.twp[i];
      float g = k_ade * sgnf(d);
      if (t == H - 1) g += k_fde * sgnf(d);
      // second differences s_u = w[u+2] - 2 w[u+1] + w[u], u = 0..H-3; w[t] appears in u = t-2 (+1), t-1 (-2), t (+1)
      const float* w = a.wp + (int64_t)b * H * 2 + c;
      for (int u = t - 2; u <= t; ++u) {
        if (u < 0 || u > H - 3) continue;
        const float s = w[(u + 2) * 2] - 2.f * w[(u + 1) * 2] + w[u * 2];
        // the reference evaluates (w2 - w1) - (w1 - w0); the sign is the same up to rounding at exact zeros
        const float sg = sgnf((w[(u + 2) * 2] - w[(u + 1) * 2]) - (w[(u + 1) * 2] - w[u * 2]));
        (void)s;
        g += k_sm * sg * (u == t - 1 ? -2.f : 1.f);
      }
      a.dwp[i] = g;
    }
  }
  if (a.dspd) {   // [B, spd_ld], zero-filled by the caller
    const float k = n_spd ? a.coef[2] / (float)n_spd : 0.f;
    if (a.speed_mode == 1) {
      for (int i = threadIdx.x; i < B * H; i += blockDim.x) {
        const int b = i / H, t = i - b * H;
        a.dspd[b * a.spd_ld + t] = k * sgnf(a.spd[b * a.spd_ld + t] - a.tspd[b * a.tspd_ld + t]);
      }
    } else if (a.speed_mode == 2) {
      for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int j = b * a.spd_ld + a.spd_ld - 1;
        a.dspd[j] = k * sgnf(a.spd[j] - a.tspd[b * a.tspd_ld + a.tspd_ld - 1]);
      }
    }
  }
  if (a.dew) {
    for (int i = threadIdx.x; i < B * E; i += blockDim.x) {
      const int e = i % E;
      float g = 0.f;
      if (a.use_lb) g += a.coef[4] * 2.f * (usage[e] - 1.f / (float)E) / (float)E / (float)B;
      if (a.use_ent) {
        const float w = a.ew[i];
        g += a.coef[5] * (logf(w + 1e-8f) + w / (w + 1e-8f)) / (float)B;
      }
      a.dew[i] = g;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ + AdamW on flat fp32 buffers.
__global__ void sq_norm_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void sq_norm_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  // single warp, fixed order: out[0] = sum, out[1] = sqrt(sum)
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    out[0] = s;
    out[1] = sqrtf(s);
  }
}
// One training step further: the optimizer's step count and the per-step part of the dropout key.  They live on the device
// so that a captured step (GraphedTrainStep) advances them on replay.
__global__ void train_tick_kernel(int* __restrict__ step_dev, uint64_t* __restrict__ seed_dev) {
  if (step_dev) *step_dev += 1;
  if (seed_dev) *seed_dev += 0xD1B54A32D192ED03ull;
}

// p, m, v updated in place.  grad is scaled by grad_scale (e.g. 1/world_size after an all-reduce SUM)
// and by the clip coefficient min(1, max_norm / (norm*grad_scale + 1e-6)) with norm read from device.
__global__ void clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, int64_t n, const float* __restrict__ norm, float grad_scale,
                                  float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  float bc1, float bc2, const int* __restrict__ step_dev) {
  if (step_dev) {      // replayed CUDA graph: the step count lives on the device (amoe_train_tick advances it)
    const float t = (float)*step_dev;
    bc1 = 1.f - powf(beta1, t);
    bc2 = 1.f - powf(beta2, t);
  }
  float clip = 1.f;
  if (max_norm > 0.f && norm) {
    const float total = norm[1] * grad_scale;
    clip = fminf(1.f, max_norm / (total + 1e-6f));
  }
  const float gs = grad_scale * clip;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * weight_decay);      // decoupled weight decay (torch.optim.AdamW)
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

template <bool TA, bool TB>
int launch_sgemm(amoe_ctx* ctx, const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                 const float* bias, int relu, float drop_p, uint64_t seed, cudaStream_t st, const uint64_t* seed_dev = nullptr) {
  if constexpr (!TA) {
    if (M <= 64) {      // the training batch: column-tile x split-K kernel (see skinny_gemm_kernel)
      const int a_vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (lda & 3) == 0) ? 1 : 0;
      skinny_gemm_kernel<TB><<<dim3(ceil_div(N, SK_TN), ceil_div(M, 32)), SK_WARPS * 32, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, bias,
                                                                                               relu, drop_p, seed, seed_dev, a_vec);
      AMOE_LAUNCH_OK(ctx);
      return 0;
    }
  }
  dim3 grid(ceil_div(M, GBM), ceil_div(N, GBN));
  sgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, bias, relu, drop_p, seed, seed_dev);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

// y[b][q][:] = max(a[b][:] + c[q][:], 0): first decoder layer of the nuScenes multi-query head - Linear(h0[b] + E[q])
// split into W h0[b] (per frame) + (W E[q] + bias) (per query, constant), models/experts/nuscenes_expert.py:172-180
__global__ void bcast_add_relu_kernel(const float4* __restrict__ a, const float4* __restrict__ c, float4* __restrict__ y, int Q,
                                      int D4, int64_t total4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D4);
    const int64_t r = i / D4;
    const int q = (int)(r % Q);
    const int64_t b = r / Q;
    const float4 u = __ldg(a + b * D4 + d), v = __ldg(c + (int64_t)q * D4 + d);
    y[i] = make_float4(fmaxf(u.x + v.x, 0.f), fmaxf(u.y + v.y, 0.f), fmaxf(u.z + v.z, 0.f), fmaxf(u.w + v.w, 0.f));
  }
}

}  // namespace

extern "C" {

int amoe_bcast_add_relu(amoe_ctx* ctx, const float* a, const float* c, float* y, int B, int Q, int D, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && a && c && y, "amoe_bcast_add_relu: NULL argument");
  AMOE_REQUIRE(D % 4 == 0 && D > 0 && Q > 0 && B >= 0, "amoe_bcast_add_relu: D must be a positive multiple of 4");
  const int64_t total4 = (int64_t)B * Q * (D / 4);
  if (total4 == 0) return 0;
  const int64_t want = (total4 + 255) / 256;
  const unsigned grid = (unsigned)(want < (int64_t)ctx->sm_count * 16 ? want : (int64_t)ctx->sm_count * 16);
  bcast_add_relu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)a, (const float4*)c, (float4*)y, Q, D / 4, total4);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}


int amoe_linear_fwd(amoe_ctx* ctx, const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int B,
                    int in_dim, int out_dim, int relu, float drop_p, uint64_t seed, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && W && y, "amoe_linear_fwd: NULL argument");
  AMOE_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "amoe_linear_fwd: dropout p=%f out of [0,1)", drop_p);
  AMOE_REQUIRE(ldx >= in_dim && ldy >= out_dim, "amoe_linear_fwd: leading dimension too small");
  if (B == 0) return 0;
  // y[B,out] = x[B,in] * W[out,in]^T
  return launch_sgemm<false, true>(ctx, x, W, y, B, out_dim, in_dim, ldx, in_dim, ldy, b, relu, drop_p, seed,
                                   (cudaStream_t)stream);
}

int amoe_linear_fwd_dseed(amoe_ctx* ctx, const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int B,
                          int in_dim, int out_dim, int relu, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && W && y && seed_dev, "amoe_linear_fwd_dseed: NULL argument");
  AMOE_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "amoe_linear_fwd_dseed: dropout p=%f out of [0,1)", drop_p);
  AMOE_REQUIRE(ldx >= in_dim && ldy >= out_dim, "amoe_linear_fwd_dseed: leading dimension too small");
  if (B == 0) return 0;
  return launch_sgemm<false, true>(ctx, x, W, y, B, out_dim, in_dim, ldx, in_dim, ldy, b, relu, drop_p, seed,
                                   (cudaStream_t)stream, seed_dev);
}

int amoe_linear_bwd(amoe_ctx* ctx, const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx,
                    const float* W, float* g_tmp, float* dx, int lddx, float* dW, float* db, int B, int in_dim,
                    int out_dim, int relu, float drop_p, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && W, "amoe_linear_bwd: NULL argument");
  AMOE_REQUIRE(!relu || (y && g_tmp), "amoe_linear_bwd: y and g_tmp are required behind a ReLU");
  AMOE_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "amoe_linear_bwd: dropout p=%f out of [0,1)", drop_p);
  AMOE_REQUIRE(!drop_p || relu, "amoe_linear_bwd: dropout is only defined behind ReLU (the mask is recovered from y > 0)");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* g = dy;
  int ldg = lddy;
  if (relu) {
    mask_grad_kernel<<<ceil_div(B * out_dim, 256), 256, 0, st>>>(dy, lddy, y, ldy, g_tmp, B, out_dim, 1.f / (1.f - drop_p));
    AMOE_LAUNCH_OK(ctx);
    g = g_tmp;
    ldg = out_dim;
  }
  int rc = 0;
  if (dx) {   // dx[B,in] = g[B,out] * W[out,in]
    AMOE_REQUIRE(lddx >= in_dim, "amoe_linear_bwd: lddx too small");
    rc = launch_sgemm<false, false>(ctx, g, W, dx, B, in_dim, out_dim, ldg, in_dim, lddx, nullptr, 0, 0.f, 0, st);
    if (rc) return rc;
  }
  if (dW) {   // dW[out,in] = g[B,out]^T * x[B,in]
    AMOE_REQUIRE(x && ldx >= in_dim, "amoe_linear_bwd: x is required for dW");
    rc = launch_sgemm<true, false>(ctx, g, x, dW, out_dim, in_dim, B, ldg, ldx, in_dim, nullptr, 0, 0.f, 0, st);
    if (rc) return rc;
  }
  if (db) {
    colsum_kernel<<<ceil_div(out_dim, 128), 128, 0, st>>>(g, ldg, db, B, out_dim);
    AMOE_LAUNCH_OK(ctx);
  }
  return 0;
}

int amoe_layernorm_fwd(amoe_ctx* ctx, const float* x, const float* gamma, const float* beta, float* y, float* mean,
                       float* rstd, int B, int D, float eps, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && gamma && beta && y && mean && rstd, "amoe_layernorm_fwd: NULL argument");
  if (B == 0) return 0;
  layernorm_fwd_kernel<<<ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, B, D, eps);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_layernorm_bwd(amoe_ctx* ctx, const float* dy, const float* x, const float* gamma, const float* mean,
                       const float* rstd, float* dx, float* dgamma, float* dbeta, int B, int D, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && x && gamma && mean && rstd, "amoe_layernorm_bwd: NULL argument");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    layernorm_bwd_dx_kernel<<<ceil_div(B, 4), 128, 0, st>>>(dy, x, gamma, mean, rstd, dx, B, D);
    AMOE_LAUNCH_OK(ctx);
  }
  if (dgamma || dbeta) {
    AMOE_REQUIRE(dgamma && dbeta, "amoe_layernorm_bwd: dgamma and dbeta come together");
    layernorm_bwd_param_kernel<<<ceil_div(D, 128), 128, 0, st>>>(dy, x, mean, rstd, dgamma, dbeta, B, D);
    AMOE_LAUNCH_OK(ctx);
  }
  return 0;
}

int amoe_gate_combine_fwd_ex(amoe_ctx* ctx, const float* logits, const float* processed, int64_t expert_stride, int ld_p,
                             float temperature, int use_softmax, float* weights, float* combined, int B, int E, int P,
                             void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && logits && processed && weights && combined, "amoe_gate_combine_fwd: NULL argument");
  AMOE_REQUIRE(E >= 1 && E <= MAX_E && temperature > 0.f, "amoe_gate_combine_fwd: E=%d / temperature out of range", E);
  if (B == 0) return 0;
  gate_combine_fwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(logits, processed, expert_stride, ld_p, 1.f / temperature,
                                                             weights, combined, B, E, P, use_softmax ? 0 : 1);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_gate_combine_fwd(amoe_ctx* ctx, const float* logits, const float* processed, int64_t expert_stride, int ld_p,
                          float temperature, float* weights, float* combined, int B, int E, int P, void* stream) {
  return amoe_gate_combine_fwd_ex(ctx, logits, processed, expert_stride, ld_p, temperature, 1, weights, combined, B, E, P, stream);
}

int amoe_gate_combine_bwd_ex(amoe_ctx* ctx, const float* dcombined, const float* dweights, const float* weights,
                             const float* processed, int64_t expert_stride, int ld_p, float temperature,
                             const float* logits_if_sigmoid, float* dlogits, float* dprocessed, int64_t dexpert_stride, int B,
                             int E, int P, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && weights && processed && dlogits && dprocessed, "amoe_gate_combine_bwd: NULL argument");
  AMOE_REQUIRE(E >= 1 && E <= MAX_E && temperature > 0.f, "amoe_gate_combine_bwd: E=%d / temperature out of range", E);
  if (B == 0) return 0;
  gate_combine_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(dcombined, dweights, weights, processed, expert_stride, ld_p,
                                                             1.f / temperature, dlogits, dprocessed, dexpert_stride, B, E, P,
                                                             logits_if_sigmoid);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_gate_combine_bwd(amoe_ctx* ctx, const float* dcombined, const float* dweights, const float* weights,
                          const float* processed, int64_t expert_stride, int ld_p, float temperature, float* dlogits,
                          float* dprocessed, int64_t dexpert_stride, int B, int E, int P, void* stream) {
  return amoe_gate_combine_bwd_ex(ctx, dcombined, dweights, weights, processed, expert_stride, ld_p, temperature, nullptr,
                                  dlogits, dprocessed, dexpert_stride, B, E, P, stream);
}

int amoe_gating_loss_fwd_bwd(amoe_ctx* ctx, const float* waypoints, const float* speed, int speed_ld,
                             const float* expert_weights, const float* tgt_waypoints, const float* tgt_speed,
                             int tgt_speed_ld, int B, int H, int E, int speed_mode, const float* coef_host, int use_lb,
                             int use_entropy, float* losses, float* d_waypoints, float* d_speed, float* d_weights,
                             void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && waypoints && expert_weights && tgt_waypoints && coef_host && losses,
               "amoe_gating_loss_fwd_bwd: NULL argument");
  AMOE_REQUIRE(speed_mode == 0 || (speed && tgt_speed), "amoe_gating_loss_fwd_bwd: speed tensors missing");
  AMOE_REQUIRE(E >= 1 && E <= MAX_E && H >= 1 && B >= 1, "amoe_gating_loss_fwd_bwd: bad sizes B=%d H=%d E=%d", B, H, E);
  LossArgs a;
  a.wp = waypoints; a.spd = speed; a.ew = expert_weights; a.twp = tgt_waypoints; a.tspd = tgt_speed;
  a.out = losses; a.dwp = d_waypoints; a.dspd = d_speed; a.dew = d_weights;
  a.B = B; a.H = H; a.E = E; a.speed_mode = speed_mode; a.spd_ld = speed_ld; a.tspd_ld = tgt_speed_ld;
  a.use_lb = use_lb; a.use_ent = use_entropy;
  for (int i = 0; i < 6; ++i) a.coef[i] = coef_host[i];
  gating_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_sq_norm(amoe_ctx* ctx, const float* g, int64_t n, float* partial_ws, int ws_floats, float* out2, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && g && partial_ws && out2, "amoe_sq_norm: NULL argument");
  AMOE_REQUIRE(ws_floats >= 1, "amoe_sq_norm: workspace too small");
  const int blocks = (int)std::min<int64_t>(std::min(ws_floats, 4 * ctx->sm_count), std::max<int64_t>(1, (n + 1023) / 1024));
  sq_norm_partial_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, partial_ws);
  AMOE_LAUNCH_OK(ctx);
  sq_norm_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partial_ws, blocks, out2);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_fused_clip_adamw(amoe_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                          const float* norm2, float grad_scale, float max_norm, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && params && grads && exp_avg && exp_avg_sq, "amoe_fused_clip_adamw: NULL argument");
  AMOE_REQUIRE(step >= 1, "amoe_fused_clip_adamw: step counts from 1");
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  const int blocks = (int)std::min<int64_t>(8 * ctx->sm_count, (n + 255) / 256);
  clip_adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, norm2, grad_scale,
                                                             max_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2, nullptr);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_fused_clip_adamw_dstep(amoe_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                const float* norm2, float grad_scale, float max_norm, float lr, float beta1, float beta2,
                                float eps, float weight_decay, const int* step_dev, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && params && grads && exp_avg && exp_avg_sq && step_dev, "amoe_fused_clip_adamw_dstep: NULL argument");
  if (n == 0) return 0;
  const int blocks = (int)std::min<int64_t>(8 * ctx->sm_count, (n + 255) / 256);
  clip_adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, norm2, grad_scale,
                                                             max_norm, lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, step_dev);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_train_tick(amoe_ctx* ctx, int* step_dev, uint64_t* seed_dev, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && (step_dev || seed_dev), "amoe_train_tick: NULL argument");
  train_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, seed_dev);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
