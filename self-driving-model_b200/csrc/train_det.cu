// Training-step kernels of a perception expert (BDDTrainer._train_detection_batch,
// training/train_bdd100k_ddp.py:117-186) that the gating/policy kernels do not already cover:
//   amoe_maxpool3x3s2_bwd        backward of the ResNet stem max-pool (NHWC fp32)
//   amoe_add_relu_fwd / _bwd     BasicBlock tail: out = relu(main + identity)
//   amoe_det_targets             scatter of the Hungarian assignment into per-query targets
//   amoe_det_loss_fwd_bwd        CrossEntropy(ignore_index = num_classes) + w * SmoothL1(matched), with
//                                the gradients w.r.t. the raw logits / box predictions
#include "common.cuh"

namespace {

// dx[n,ih,iw,c] = sum over the (at most 4) pooling windows that contain (ih,iw) of dy[window] if (ih,iw)
// is that window's arg-max.  The arg-max is recomputed with PyTorch's rule (first strictly greater value
// in kh-major, kw-minor scan order), so ties route the gradient exactly as torch's max_pool2d backward.
// One thread = four channels of one input pixel (16-byte loads; the scalar version issued 36 four-byte reads per output).
__global__ void maxpool3x3s2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                        int H, int W, int C, int Ho, int Wo, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  const int c = (int)(i % C4) << 2;
  int64_t r = i / C4;
  const int iw = (int)(r % W);
  r /= W;
  const int ih = (int)(r % H);
  const int64_t n = r / H;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  // windows (oh, ow) with 2*oh-1 <= ih <= 2*oh+1
  for (int oh = (ih) / 2; oh <= (ih + 1) / 2; ++oh) {
    if (oh < 0 || oh >= Ho) continue;
    for (int ow = (iw) / 2; ow <= (iw + 1) / 2; ++ow) {
      if (ow < 0 || ow >= Wo) continue;
      float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int bh[4] = {-1, -1, -1, -1}, bw[4] = {-1, -1, -1, -1};
      for (int kh = 0; kh < 3; ++kh) {
        const int yy = oh * 2 - 1 + kh;
        if (yy < 0 || yy >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int xx = ow * 2 - 1 + kw;
          if (xx < 0 || xx >= W) continue;
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(x + ((n * H + yy) * W + xx) * C + c));
          const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (v[j] > best[j] || bh[j] < 0 || v[j] != v[j]) {   // first element always taken; NaN propagates as in torch
              best[j] = v[j]; bh[j] = yy; bw[j] = xx;
            }
        }
      }
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(dy + ((n * Ho + oh) * Wo + ow) * C + c));
      const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (bh[j] == ih && bw[j] == iw) g[j] += d[j];
    }
  }
  *reinterpret_cast<float4*>(dx + i * 4) = make_float4(g[0], g[1], g[2], g[3]);
}

// The same backward in two passes: (1) the arg-max tap (kh*3 + kw, one byte) of every pooling window, 9 loads per window;
// (2) every input pixel compares its own tap number against the <= 4 windows that contain it and sums their dy.  The one-pass
// kernel above rescans the 3x3 window of each of those windows per input pixel (20 16-byte loads per thread on average:
// 1.85 ms for 8 x 360 x 640 x 64, 0.57 TB/s).  Same arg-max rule, same order of the <= 4 additions: bit-identical dx.
__global__ void maxpool3x3s2_argmax_kernel(const float* __restrict__ x, uchar4* __restrict__ idx, int H, int W, int C, int Ho,
                                           int Wo, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  const int c = (int)(i % C4) << 2;
  int64_t r = i / C4;
  const int ow = (int)(r % Wo);
  r /= Wo;
  const int oh = (int)(r % Ho);
  const int64_t n = r / Ho;
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int bt[4] = {-1, -1, -1, -1};
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int yy = oh * 2 - 1 + kh;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int xx = ow * 2 - 1 + kw;
      if (xx < 0 || xx >= W) continue;
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(x + ((n * H + yy) * W + xx) * C + c));
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (v[j] > best[j] || bt[j] < 0 || v[j] != v[j]) {   // first element always taken; NaN propagates as in torch
          best[j] = v[j]; bt[j] = kh * 3 + kw;
        }
    }
  }
  idx[i] = make_uchar4((unsigned char)bt[0], (unsigned char)bt[1], (unsigned char)bt[2], (unsigned char)bt[3]);
}

__global__ void maxpool3x3s2_bwd_idx_kernel(const uchar4* __restrict__ idx, const float* __restrict__ dy, float* __restrict__ dx,
                                            int H, int W, int C, int Ho, int Wo, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  const int c4 = (int)(i % C4);
  int64_t r = i / C4;
  const int iw = (int)(r % W);
  r /= W;
  const int ih = (int)(r % H);
  const int64_t n = r / H;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  for (int oh = ih / 2; oh <= (ih + 1) / 2; ++oh) {
    if (oh >= Ho) continue;
    const int kh = ih - (oh * 2 - 1);
    for (int ow = iw / 2; ow <= (iw + 1) / 2; ++ow) {
      if (ow >= Wo) continue;
      const unsigned tap = (unsigned)(kh * 3 + (iw - (ow * 2 - 1)));
      const int64_t o = ((n * Ho + oh) * Wo + ow) * C4 + c4;
      const uchar4 id = __ldg(idx + o);
      const float4 d = __ldg(reinterpret_cast<const float4*>(dy) + o);
      if (id.x == tap) g[0] += d.x;
      if (id.y == tap) g[1] += d.y;
      if (id.z == tap) g[2] += d.z;
      if (id.w == tap) g[3] += d.w;
    }
  }
  reinterpret_cast<float4*>(dx)[i] = make_float4(g[0], g[1], g[2], g[3]);
}

__global__ void add_relu_fwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 u = a[i], v = b[i];
  y[i] = make_float4(fmaxf(u.x + v.x, 0.f), fmaxf(u.y + v.y, 0.f), fmaxf(u.z + v.z, 0.f), fmaxf(u.w + v.w, 0.f));
}
__global__ void relu_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, float4* __restrict__ g, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 d = dy[i], v = y[i];
  g[i] = make_float4(v.x > 0.f ? d.x : 0.f, v.y > 0.f ? d.y : 0.f, v.z > 0.f ? d.z : 0.f, v.w > 0.f ? d.w : 0.f);
}

// Adjoint of F.interpolate(low, size=(H,W), mode="bilinear", align_corners=False) + the NHWC->NCHW transpose:
// dlow[b,y,x,c] = sum over output pixels (Y,X) of dy[b,c,Y,X] * wy(Y,y) * wx(X,x), with the forward's index rule
// (aten upsample_bilinear2d: src = max(scale*(dst+0.5)-0.5, 0)).  One warp per (b,c,y,x): lanes walk X of the
// window of output pixels that can touch the cell (coalesced rows of dy), fixed summation order.
__device__ __forceinline__ void up_src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = min((int)s, in_size - 1);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}
__global__ void upsample_bilinear_nchw_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dlow, int h, int w, int C,
                                                  int H, int W, float sh, float sw_, int64_t n_cells) {
  const int64_t cell = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (cell >= n_cells) return;
  const int c = (int)(cell % C);
  int64_t r = cell / C;
  const int x = (int)(r % w);
  r /= w;
  const int y = (int)(r % h);
  const int64_t b = r / h;
  // output rows/cols whose source interval [i0, i1] can contain this cell (generous bounds, exact test inside)
  const float inv_sh = 1.f / sh, inv_sw = 1.f / sw_;
  const int Y0 = max(0, (int)floorf(((float)y - 1.f + 0.5f) * inv_sh - 0.5f) - 1);
  const int Y1 = min(H - 1, (int)ceilf(((float)y + 1.f + 0.5f) * inv_sh - 0.5f) + 1);
  const int X0 = max(0, (int)floorf(((float)x - 1.f + 0.5f) * inv_sw - 0.5f) - 1);
  const int X1 = min(W - 1, (int)ceilf(((float)x + 1.f + 0.5f) * inv_sw - 0.5f) + 1);
  const float* plane = dy + (b * C + c) * (int64_t)H * W;
  float acc = 0.f;
  for (int Y = Y0; Y <= Y1; ++Y) {
    int y0, y1;
    float ly;
    up_src_index(sh, Y, h, y0, y1, ly);
    float wy = 0.f;
    if (y0 == y) wy += 1.f - ly;
    if (y1 == y) wy += ly;
    if (wy == 0.f) continue;
    float row = 0.f;
    for (int X = X0 + lane; X <= X1; X += 32) {
      int xa, xb;
      float lx;
      up_src_index(sw_, X, w, xa, xb, lx);
      float wx = 0.f;
      if (xa == x) wx += 1.f - lx;
      if (xb == x) wx += lx;
      if (wx != 0.f) row = fmaf(plane[(int64_t)Y * W + X], wx, row);
    }
    acc = fmaf(wy, row, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) dlow[((b * h + y) * w + x) * (int64_t)C + c] = acc;
}

// target_classes[b*Q + pred_idx[m]] = labels[m]; target_boxes[...] = boxes[m]  for the matched pairs m of
// image b = batch_of[m] (everything else: class = num_classes (ignored), box = 0; filled by the caller)
__global__ void det_targets_kernel(const int64_t* __restrict__ pred_idx, const int32_t* __restrict__ batch_of,
                                   const int64_t* __restrict__ labels, const float* __restrict__ boxes, int n_match, int Q,
                                   int64_t* __restrict__ target_classes, float* __restrict__ target_boxes) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_match) return;
  const int64_t row = (int64_t)batch_of[m] * Q + pred_idx[m];
  target_classes[row] = labels[m];
#pragma unroll
  for (int j = 0; j < 4; ++j) target_boxes[row * 4 + j] = boxes[(int64_t)m * 4 + j];
}

// one CTA; rows = B*Q.  logits [rows, ld_l] (first C columns), boxes [rows, ld_b] (first 4 columns).
//   class_loss = mean over rows with target != ignore of -log_softmax(logits)[target]
//   bbox_loss  = mean over matched rows x 4 of smooth_l1(pred - target), beta = 1
//   out[0] = class_loss + w_box * bbox_loss, out[1] = class_loss, out[2] = bbox_loss, out[3] = #matched
__global__ __launch_bounds__(1024) void det_loss_kernel(const float* __restrict__ logits, int ld_l, const float* __restrict__ boxes,
                                                        int ld_b, const int64_t* __restrict__ tcls, const float* __restrict__ tbox,
                                                        int64_t rows, int C, int ignore, float w_box, float* __restrict__ out,
                                                        float* __restrict__ dlogits, int ld_dl, float* __restrict__ dboxes,
                                                        int ld_db) {
  __shared__ double red[32];
  __shared__ double tot[3];
  double s_ce = 0.0, s_bx = 0.0, s_n = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) {
    const int64_t t = tcls[r];
    if (t == ignore || t < 0 || t >= C) continue;      // labels outside [0, C) never index the logits (torch raises there)
    const float* l = logits + r * ld_l;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, l[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(l[c] - mx);
    s_ce += (double)(logf(se) + mx - l[t]);
    for (int j = 0; j < 4; ++j) {
      const float d = boxes[r * ld_b + j] - tbox[r * 4 + j], ad = fabsf(d);
      s_bx += (double)(ad < 1.f ? 0.5f * d * d : ad - 0.5f);
    }
    s_n += 1.0;
  }
  double v[3] = {s_ce, s_bx, s_n};
  for (int k = 0; k < 3; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      tot[k] = s;
    }
  }
  __syncthreads();
  const double n = tot[2];
  // CrossEntropyLoss(mean) over zero non-ignored targets is NaN in torch (0/0); SmoothL1 is skipped (0)
  const float ce = n > 0 ? (float)(tot[0] / n) : __int_as_float(0x7fc00000);
  const float bx = n > 0 ? (float)(tot[1] / (4.0 * n)) : 0.f;
  if (threadIdx.x == 0) {
    out[0] = ce + w_box * bx; out[1] = ce; out[2] = bx; out[3] = (float)n;
  }
  const float inv_n = n > 0 ? (float)(1.0 / n) : 0.f, inv_4n = n > 0 ? (float)(w_box / (4.0 * n)) : 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) {
    const int64_t t = tcls[r];
    float* dl = dlogits ? dlogits + r * ld_dl : nullptr;
    float* db = dboxes ? dboxes + r * ld_db : nullptr;
    if (t == ignore || t < 0 || t >= C) {
      if (dl) for (int c = 0; c < C; ++c) dl[c] = 0.f;
      if (db) for (int j = 0; j < 4; ++j) db[j] = 0.f;
      continue;
    }
    const float* l = logits + r * ld_l;
    if (dl) {
      float mx = -INFINITY;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, l[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(l[c] - mx);
      for (int c = 0; c < C; ++c) dl[c] = (expf(l[c] - mx) / se - (c == t ? 1.f : 0.f)) * inv_n;
    }
    if (db) {
      for (int j = 0; j < 4; ++j) {
        const float d = boxes[r * ld_b + j] - tbox[r * 4 + j];
        db[j] = (fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * inv_4n;
      }
    }
  }
}

}  // namespace

extern "C" {

int amoe_maxpool3x3s2_bwd(amoe_ctx* ctx, const float* x, const float* dy, float* dx, int NB, int H, int W, int C, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && dy && dx, "amoe_maxpool3x3s2_bwd: NULL argument");
  if (NB == 0) return 0;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  AMOE_REQUIRE(C % 4 == 0, "amoe_maxpool3x3s2_bwd: C=%d must be a multiple of 4", C);
  const int64_t total4 = (int64_t)NB * H * W * (C / 4);
  maxpool3x3s2_bwd_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, H, W, C, Ho, Wo, total4);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_maxpool3x3s2_bwd_ws(amoe_ctx* ctx, const float* x, const float* dy, float* dx, void* argmax_ws, int NB, int H, int W,
                             int C, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && dy && dx && argmax_ws, "amoe_maxpool3x3s2_bwd_ws: NULL argument");
  if (NB == 0) return 0;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  AMOE_REQUIRE(C % 4 == 0, "amoe_maxpool3x3s2_bwd_ws: C=%d must be a multiple of 4", C);
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(argmax_ws) & 3) == 0, "amoe_maxpool3x3s2_bwd_ws: workspace must be 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t out4 = (int64_t)NB * Ho * Wo * (C / 4), in4 = (int64_t)NB * H * W * (C / 4);
  maxpool3x3s2_argmax_kernel<<<(unsigned)((out4 + 255) / 256), 256, 0, st>>>(x, (uchar4*)argmax_ws, H, W, C, Ho, Wo, out4);
  AMOE_LAUNCH_OK(ctx);
  maxpool3x3s2_bwd_idx_kernel<<<(unsigned)((in4 + 255) / 256), 256, 0, st>>>((const uchar4*)argmax_ws, dy, dx, H, W, C, Ho, Wo, in4);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_add_relu_fwd(amoe_ctx* ctx, const float* a, const float* b, float* y, int64_t n, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && a && b && y, "amoe_add_relu_fwd: NULL argument");
  AMOE_REQUIRE(n % 4 == 0, "amoe_add_relu_fwd: element count must be a multiple of 4");
  if (n == 0) return 0;
  add_relu_fwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)a, (const float4*)b, (float4*)y, n / 4);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_relu_bwd(amoe_ctx* ctx, const float* dy, const float* y, float* g, int64_t n, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && y && g, "amoe_relu_bwd: NULL argument");
  AMOE_REQUIRE(n % 4 == 0, "amoe_relu_bwd: element count must be a multiple of 4");
  if (n == 0) return 0;
  relu_bwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)dy, (const float4*)y,
                                                                                     (float4*)g, n / 4);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_upsample_bilinear_nchw_bwd(amoe_ctx* ctx, const float* dy, float* dlow, int B, int h, int w, int C, int H, int W,
                                    void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy && dlow, "amoe_upsample_bilinear_nchw_bwd: NULL argument");
  const int64_t cells = (int64_t)B * h * w * C;
  if (cells == 0) return 0;
  const float sh = (float)h / (float)H, sw_ = (float)w / (float)W;
  upsample_bilinear_nchw_bwd_kernel<<<(unsigned)((cells * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, dlow, h, w, C, H, W,
                                                                                                         sh, sw_, cells);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_det_targets(amoe_ctx* ctx, const int64_t* pred_idx, const int32_t* batch_of, const int64_t* labels,
                     const float* boxes, int n_match, int Q, int64_t* target_classes, float* target_boxes, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && target_classes && target_boxes, "amoe_det_targets: NULL argument");
  if (n_match == 0) return 0;
  AMOE_REQUIRE(pred_idx && batch_of && labels && boxes, "amoe_det_targets: NULL argument");
  det_targets_kernel<<<ceil_div(n_match, 128), 128, 0, (cudaStream_t)stream>>>(pred_idx, batch_of, labels, boxes, n_match, Q,
                                                                             target_classes, target_boxes);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_det_loss_fwd_bwd(amoe_ctx* ctx, const float* logits, int ld_logits, const float* boxes, int ld_boxes,
                          const int64_t* target_classes, const float* target_boxes, int64_t rows, int C, int ignore_index,
                          float bbox_weight, float* losses4, float* dlogits, int ld_dlogits, float* dboxes, int ld_dboxes,
                          void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && logits && boxes && target_classes && target_boxes && losses4, "amoe_det_loss_fwd_bwd: NULL argument");
  AMOE_REQUIRE(C >= 1 && ld_logits >= C && ld_boxes >= 4, "amoe_det_loss_fwd_bwd: bad sizes");
  det_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, ld_logits, boxes, ld_boxes, target_classes, target_boxes, rows,
                                                       C, ignore_index, bbox_weight, losses4, dlogits, ld_dlogits, dboxes,
                                                       ld_dboxes);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
