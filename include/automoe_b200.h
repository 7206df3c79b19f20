/*
 * automoe_b200.h — C-ABI of the B200-native AutoMoE forward hot path.
 *
 * The reference (immanuel-peter/self-driving-model) has no FFI: its boundary is
 * the Python nn.Module API.  Each entry point below replaces the stock
 * torch/cuDNN/cuBLAS/scipy call sequence of one reference function and is what a
 * maintainer would bind (ctypes stub in INTEGRATION.md) from that function.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  Device pointers unless a name
 *     ends in _host.  The caller (PyTorch) owns every buffer.
 *   - every launch goes to the cudaStream_t passed in (as void*); nothing here
 *     synchronises the device or allocates device memory, except amoe_create
 *     (one 4-byte flag) and the host-side LSAP (host memory only).
 *   - return 0 on success, <0 on error; amoe_last_error() gives the message
 *     (thread-local).
 *   - dtype enum: AMOE_F32 = 0, AMOE_BF16 = 1.
 *   - activations are NHWC ("channels-last"): x[n][h][w][c]; conv weights are
 *     packed [Cout][KH][KW][Cin] (K-major for the implicit GEMM).
 *   - "groups" G = number of experts run in one launch: activations are stacked
 *     on the batch axis ([G*B,H,W,C]) and weights/scale/bias on the Cout axis
 *     ([G*Cout,...]).  G=1 is an ordinary convolution.
 */
#ifndef AUTOMOE_B200_H_
#define AUTOMOE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMOE_F32 0
#define AMOE_BF16 1

typedef struct amoe_ctx amoe_ctx;

/* ---- library / context ------------------------------------------------- */
int amoe_abi_version(void);
const char* amoe_last_error(void);
/* One context per device: resolves cuTensorMapEncodeTiled, caches SM count,
 * sets the max-dynamic-smem attribute of the tcgen05 kernels. */
int amoe_create(int device, amoe_ctx** out);
int amoe_destroy(amoe_ctx* ctx);
int amoe_sm_count(amoe_ctx* ctx);
/* number of kernels launched through this context since creation (bench.py's
 * "gpu_launches" counter) */
int64_t amoe_launch_count(amoe_ctx* ctx);
/* Tile walk direction of the tensor-core convolutions launched through this context from now on (0: front to back,
 * 1: back to front).  Results do not depend on it.  A chain of convolutions that alternates the direction starts every
 * layer on the part of its input (and residual) the previous layer touched last - the part still in L2; the reference has
 * no counterpart (cuDNN schedules its own grids). */
int amoe_set_walk_reverse(amoe_ctx* ctx, int reverse);

/* ---- layout / weight packing ------------------------------------------- */
/* batch['image'] [B,C,H,W] fp32 NCHW (models/automoe.py:212-218 consumers) ->
 * NHWC with C padded to Cp (zeros).  dst dtype f32 or bf16. */
int amoe_image_nchw_to_nhwc(amoe_ctx*, const float* src, void* dst, int B, int C,
                            int H, int W, int Cp, int dst_dtype, void* stream);
/* Same, into a physically padded frame dst: [B,Hpad,Wpad,Cp]: image pixel (h,w) lands at
 * (top+h, left+w), everything else is zero (the layout amoe_stem_fwd / amoe_conv2d_rowwin_fwd read). */
int amoe_image_nchw_to_nhwc_padded(amoe_ctx*, const float* src, void* dst, int B, int C, int H,
                                   int W, int Cp, int left, int Wpad, int top, int Hpad,
                                   int dst_dtype, void* stream);
/* Same with the padding channels (c >= C) set to pad_channel_value at EVERY position of the frame, the
 * zero border included (1.0 lets the stem GEMM add folded biases through that channel). */
int amoe_image_nchw_to_nhwc_padded_v(amoe_ctx*, const float* src, void* dst, int B, int C, int H,
                                     int W, int Cp, int left, int Wpad, int top, int Hpad,
                                     int dst_dtype, float pad_channel_value, void* stream);
/* ---- input staging from camera bytes ------------------- */
/* uint8 HWC RGB frames [B,H,W,3] -> normalised (u/255 - mean)/std, the transform of
 * inference/run_automoe.py:25-31 (ToTensor + Normalize; IEEE fp32 op order, bit-exact), written as the
 * physically padded bf16 NHWC4 frame [B,Hpad,Wpad,4] that amoe_stem_fwd / amoe_stem_pool_fwd read (same
 * geometry arguments as amoe_image_nchw_to_nhwc_padded_v).  mean/std: HOST arrays of 3 floats. */
int amoe_stage_u8_hwc_fwd(amoe_ctx*, const void* src_u8, void* dst, int B, int H, int W, int left,
                          int Wpad, int top, int Hpad, const float* mean3_host,
                          const float* std3_host, float pad_channel_value, void* stream);
/* Same arithmetic into the reference's own tensor: [B,3,H,W] fp32 NCHW (fp32 mode / non-tensor-core stems). */
int amoe_normalize_u8_hwc_to_nchw_fwd(amoe_ctx*, const void* src_u8, float* dst, int B, int H, int W,
                                      const float* mean3_host, const float* std3_host, void* stream);
/* One separable pass of Pillow's 8-bit bilinear/antialias resize (what T.Resize does on the PIL image at
 * run_automoe.py:27; Pillow src/libImaging/Resample.c ImagingResampleHorizontal_8bpc / Vertical_8bpc):
 * src viewed as [outer][in_size][inner] u8 -> dst [outer][out_size][inner] u8;
 * bounds: DEVICE int[2*out_size] (first input index, tap count), coeffs: DEVICE int[out_size*ksize]
 * fixed-point weights (22 fractional bits) prepared as precompute_coeffs + normalize_coeffs_8bpc do. */
int amoe_resample_u8_fwd(amoe_ctx*, const void* src_u8, void* dst_u8, const int* bounds,
                         const int* coeffs, int ksize, int64_t outer, int in_size, int out_size,
                         int inner, void* stream);
/* All first-layer convolutions of the frame (Cin=3, stride 2: the ResNet stems of the experts and
 * the policy's conv1) as one tensor-core GEMM over the raw image rows (csrc/stem_tc.cu).
 *   x_pad: [B,H+6,Wpad,4] bf16 from amoe_image_nchw_to_nhwc_padded(left=4, top=3), Wpad >= W+6
 *   w_img: filters packed as [KH*4][n_total][8] bf16: K index = kh*32 + j*4 + c holds the weight of
 *          padded pixel 2*ow+j, padded row 2*oh+kh, channel c (zero where no tap)
 *   scale,bias: [n_total] folded BN;  output pixel (b,oh,ow), channels [32*i, 32*i+32) go to
 *   dst[i] + ((b*Ho+oh)*Wo+ow)*dst_c[i]  (dst/dst_c: HOST arrays of n_total/32 device pointers /
 *   channel counts).  H, W even, W/2 <= 128. */
int amoe_stem_fwd(amoe_ctx*, const void* x_pad, const void* w_img, const float* scale,
                  const float* bias, int B, int H, int W, int Wpad, int KH, int n_total, int relu,
                  void* const* dst_host, const int* dst_c_host, void* stream);
/* fp32-accurate variant for the training forward (the reference trains in fp32: resnet.py:197 conv1 and
 * models/policy/trajectory_head.py:10 inside training/train_gating_network.py:96): split bf16 operands (three parts each, six
 * product terms, fp32 accumulation).  x3_pad: the three parts of the padded frame stacked on the batch axis
 * [3][B][H+6][Wpad][4] bf16 (4th channel zero); w3_img: three filter images [3][KH*4][n_total][8] bf16 in the layout of
 * amoe_stem_fwd; y: [B][H/2][W/2][n_total] fp32 = conv * scale + bias (+ ReLU).  Wpad = (W + 13) & ~7. */
int amoe_stem_fwd_f32tc_supported(int H, int W, int KH, int n_total);
/* [B][H][W][4] fp32 NHWC frame (amoe_image_nchw_to_nhwc, Cp = 4) -> x3_pad of amoe_stem_fwd_f32tc (borders written as zeros) */
int amoe_stem_split_frame(amoe_ctx*, const float* x_nhwc4, void* x3_pad, int B, int H, int W, int Wpad, void* stream);
int amoe_stem_fwd_f32tc(amoe_ctx*, const void* x3_pad, const void* w3_img, const float* scale, const float* bias,
                        float* y, int B, int H, int W, int Wpad, int KH, int n_total, int relu, void* stream);
/* Same GEMM with nn.MaxPool2d(3, stride 2, pad 1) fused behind the first n_pool_ch channels (the
 * expert stems: resnet conv1+bn1+relu+maxpool in one kernel, the full-resolution stem output never
 * reaches HBM).  pooled: [n_pool_ch/64 * B][H/4+2*out_pad][W/4+2*out_pad][64] bf16; with
 * out_pad = 1 the zero border is written too.  Channels >= n_pool_ch go to dst/dst_c as above
 * (entries below n_pool_ch/32 are ignored).  H, W multiples of 4.
 * scale == bias == NULL selects FOLDED filters: w_img holds bf16(w * scale) and the bias split into two
 * bf16 parts at K slots (kh=0, j=0, c=3) and (kh=0, j=1, c=3); the frame must be staged with
 * pad_channel_value = 1 (amoe_image_nchw_to_nhwc_padded_v).  The epilogue then only packs, pools and
 * applies the ReLU (the scale/bias epilogue costs more issue slots than the N=224 MMAs take). */
int amoe_stem_pool_fwd(amoe_ctx*, const void* x_pad, const void* w_img, const float* scale,
                       const float* bias, int B, int H, int W, int Wpad, int KH, int n_total,
                       int relu, int n_pool_ch, void* pooled, int out_pad, void* const* dst_host,
                       const int* dst_c_host, void* stream);
/* nn.Conv2d weight [Cout,Cin,KH,KW] fp32 -> packed [Cout][KH][KW][Cin_pad]
 * (dtype f32|bf16, zero-padded channels).  dst points at the first row of this
 * conv inside a (possibly grouped) packed buffer. */
int amoe_pack_conv_weight(amoe_ctx*, const float* w_oihw, void* dst, int Cout, int Cin,
                          int KH, int KW, int Cin_pad, int dst_dtype, void* stream);
/* eval-mode BatchNorm2d folded with the optional conv bias into y = scale*conv + bias
 * (torchvision resnet BasicBlock bn1/bn2, models/policy/trajectory_head.py:8-24).
 * gamma==NULL -> no BN: scale=1, bias=conv_bias (or 0). */
int amoe_fold_bn(amoe_ctx*, const float* gamma, const float* beta, const float* mean,
                 const float* var, float eps, const float* conv_bias, int C,
                 float* scale, float* bias, void* stream);

/* ---- convolution (replaces nn.Conv2d + BatchNorm2d + ReLU + residual add;
 *      torchvision BasicBlock.forward, models/experts/bdd_*_expert.py:12-24,
 *      models/policy/trajectory_head.py:27-33) --------------------------- */
/* y = act( scale[c]*conv(x,w) + bias[c] (+ residual) )
 *   x: [G*B,H,W,Cin] (or [B,H,W,Cin] shared by all groups when x_shared!=0)
 *   w: [G*Cout,KH,KW,Cin]   scale,bias: [G*Cout] fp32
 *   y, residual: [G*B,Ho,Wo,Cout]
 *   input pixel of tap (kh,kw) for output (oh,ow): (oh*stride_h - pad_h + kh,
 *   ow*stride_w - pad_w + kw); pad_* are the top/left pads, Ho/Wo are explicit so a
 *   caller can express asymmetric padding (nn.Conv2d: Ho=(H+2*pad-KH)/stride+1).
 * dtype selects activation+weight storage (f32: SIMT fp32 kernel; bf16: tcgen05
 * implicit GEMM when the shape qualifies, else the SIMT kernel with fp32 accum).
 * impl: 0 = auto, 1 = force SIMT, 2 = force tcgen05 (error if unsupported).
 * in_pad / out_pad (tcgen05 path only): the input / the output+residual tensors are stored with a
 * physical border of that many pixels ([N][H+2p][W+2p][C], H and W still name the interior);
 * the input border must hold zeros, the output border is left untouched. */
int amoe_conv2d_fwd(amoe_ctx*, const void* x, const void* w, const float* scale,
                    const float* bias, const void* residual, void* y, int G, int x_shared,
                    int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride_h,
                    int stride_w, int pad_h, int pad_w, int Ho, int Wo, int relu, int dtype,
                    int impl, int in_pad, int out_pad, void* stream);
/* A ResNet stage entry in ONE launch: the KHxKW/stride/pad convolution (conv1 + bn1 + ReLU -> y) and the 1x1/stride/pad-0
 * downsample convolution of the same block (downsample.0 + downsample.1 -> y2) over the same input
 * (torchvision BasicBlock.forward, resnet.py:92-103).  bf16 tcgen05 path only; x, y, y2 layouts as amoe_conv2d_fwd
 * (in_pad / out_pad physical borders); w_1x1: [G*Cout][Cin] bf16.  Every output patch is computed as two consecutive
 * tiles of the persistent kernel (the 1x1 tile re-reads the centre-tap box that conv1 just pulled through L2). */
int amoe_conv2d_dual_fwd(amoe_ctx*, const void* x, const void* w, const float* scale, const float* bias,
                         void* y, const void* w_1x1, const float* scale2, const float* bias2, void* y2,
                         int G, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                         int pad, int Ho, int Wo, int relu, int relu2, int in_pad, int out_pad,
                         void* stream);
/* 1 if amoe_conv2d_fwd(impl=auto, dtype=bf16) would take the tcgen05 path. */
int amoe_conv2d_tc_supported(int H, int W, int Cin, int Cout, int stride_h, int stride_w);
/* "Row-window" convolution for tiny Cin (ResNet stem 7x7/s2 with Cin=3, policy conv1 5x5/s2)
 * on the tensor cores.  x: [B,H,Wpad,Cp] bf16 with zero columns physically stored left/right
 * (amoe_image_nchw_to_nhwc_padded).  For output column ow, filter row kh reads ONE contiguous
 * window of 64/Cp pixels starting at padded column ow*stride_w of input row oh*stride_h-pad_h+kh
 * (rows outside [0,H) are zero); the filter is packed as w: [Cout][KH][64] with zeros at window
 * positions that are not filter taps.  Cout may be the concatenation of several convolutions
 * that share the input (3 expert stems): channel ch goes to output tensor ch/split_c,
 * y: [Cout/split_c][B,Ho,Wo,split_c] bf16. */
int amoe_conv2d_rowwin_fwd(amoe_ctx*, const void* x, const void* w, const float* scale,
                           const float* bias, void* y, int B, int H, int Wpad, int Cp, int Cout,
                           int split_c, int KH, int stride_h, int stride_w, int pad_h, int Ho,
                           int Wo, int relu, void* stream);
/* nn.MaxPool2d(3, stride 2, pad 1) of the ResNet stem, NHWC. */
int amoe_maxpool3x3s2_fwd(amoe_ctx*, const void* x, void* y, int NB, int H, int W, int C,
                          int dtype, int out_pad, void* stream);
/* 3x3 / stride 1 / pad 1 convolution (+folded BN, residual, ReLU) on PHYSICALLY padded bf16
 * activations: x [G*B][H+2][W+2][Cin] with a zero border -> y [G*B][H+2][W+2][Cout] with a zero
 * border (residual: same layout as y).  "Flat-shift" halo reuse: one activation load per 256
 * output positions and 64 channels serves all nine filter taps (csrc/conv_flat.cu).
 * Cin % 64 == 0, Cout in {32..128}, W <= 118.  Same weights layout as amoe_conv2d_fwd. */
int amoe_conv3x3_flat_fwd(amoe_ctx*, const void* x, const void* w, const float* scale,
                          const float* bias, const void* residual, void* y, int G, int B, int H,
                          int W, int Cin, int Cout, int relu, void* stream);
/* Same, for a sub-batch of a larger grouped tensor: expert group g of y (of the residual) starts
 * y_group_images (res_group_images) images after group g-1 instead of B (0 = B).  Lets a caller walk
 * the batch in L2-sized chunks while the last layer of the chunked region writes straight into the
 * full [G*B_total,...] tensor (y points at the chunk's first image inside group 0). */
int amoe_conv3x3_flat_fwd_strided(amoe_ctx*, const void* x, const void* w, const float* scale,
                                  const float* bias, const void* residual, void* y, int G, int B,
                                  int H, int W, int Cin, int Cout, int relu,
                                  int64_t y_group_images, int64_t res_group_images, void* stream);
int amoe_conv3x3_flat_supported(int H, int W, int Cin, int Cout);

/* ---- expert heads ------------------------------------------------------ */
/* 1x1 conv (head[2] / decoder[2]) + global mean over pixels.
 *   x: [B,HW,Cin] (dtype)  w: [N,Cin] fp32  b: [N] fp32
 *   low: [B,HW,N] fp32 (low-res logits, NHWC)   pooled[b*pooled_ld + n] = mean_hw(low)
 *   (pooled_ld >= N lets several experts share one [B,sumC] buffer)
 * pooled feeds the extractors: mean over the x32 bilinear up-sampled map equals
 * the mean over the low-res map (expert_extractors.py:62,89 + interpolate). */
int amoe_head1x1_pool_fwd(amoe_ctx*, const void* x, const float* w, const float* b,
                          float* low, float* pooled, int pooled_ld, int B, int HW, int Cin,
                          int N, int x_dtype, void* stream);
/* F.interpolate(low, size=(H,W), mode="bilinear", align_corners=False) and the
 * NHWC->NCHW transpose in one writer (bdd_segmentation_expert.py:22,
 * bdd_drivable_expert.py:22).  low: [B,h,w,C] fp32 -> out: [B,C,H,W] (dtype). */
int amoe_upsample_bilinear_nchw_fwd(amoe_ctx*, const float* low, void* out, int B, int h,
                                    int w, int C, int H, int W, int out_dtype, void* stream);
/* mean over H*W of an NCHW tensor -> [B,C] fp32 (AdaptiveAvgPool2d(1); only used
 * when the up-sampling factor is not an integer so the identity above fails). */
int amoe_mean_hw_nchw_fwd(amoe_ctx*, const void* x, float* out, int B, int C, int HW,
                          int dtype, void* stream);

/* ---- fused gate (context extractor + expert extractors + GatingNetwork) ---
 * Replaces SimpleContextExtractor.forward (context_features.py:151-165),
 * *ExpertExtractor MLP+LN (expert_extractors.py:27-35), GatingNetwork.forward
 * (gating_network.py:122-175) — eval semantics (dropout = identity), softmax
 * gate, no top-k (unreachable through AutoMoE, automoe.py:83-91).
 *   state:   [B,4] fp32 (speed, steering, throttle, brake)
 *   pooled:  [B,sumC] fp32, experts' pooled logits concatenated (E experts,
 *            channel counts in n_ch[E], E<=4)
 *   params:  flat fp32 buffer, layout documented in gate.cu / models/automoe.py
 *   outputs (all fp32): context [B,ctx_dim], features [E][B,256],
 *            processed [E][B,256], gate_logits [B,E], weights [B,E],
 *            combined [B,256] (after output_projection)
 * mode bits: 32 = sigmoid gate (use_softmax=False, gating_network.py:159-160; temperature unused);
 *            1 = context-only path of get_expert_weights (zeros for experts, weights only);
 *   2 = `state` holds an encoded context [B,ctx_dim] (context extractor skipped);
 *   4 = `pooled` holds expert features [E][B,256] (extractors skipped);
 *   8 = stop after the context extractor; 16 = stop after the expert extractors.
 * Output pointers may be NULL when not wanted. */
int amoe_gate_fwd(amoe_ctx*, const float* state, const float* pooled, const float* params,
                  int64_t n_params, int B, int E, const int* n_ch_host, int ctx_dim,
                  int hidden, float temperature, int mode, float* context, float* features,
                  float* processed, float* gate_logits, float* weights, float* combined,
                  void* stream);

/* Same with an optional bf16 copy of the parameter buffer (same element offsets, 16-byte aligned):
 * params_bf16 != NULL and B >= 16 selects the bf16-inference variant - 16 frames per CTA, every layer with
 * K % 32 == 0 on mma.sync TF32 (bf16 weights, fp32 activations truncated to TF32, fp32 accumulation; the
 * K = 4 / C_e input layers, biases, LayerNorm and softmax stay fp32).  With mode == 0 it launches clusters of
 * E+1 CTAs per 16 frames (one expert chain per CTA, gate input gathered through distributed shared memory). */
int amoe_gate_fwd_ex(amoe_ctx*, const float* state, const float* pooled, const float* params,
                     const void* params_bf16, int64_t n_params, int B, int E, const int* n_ch_host,
                     int ctx_dim, int hidden, float temperature, int mode, float* context,
                     float* features, float* processed, float* gate_logits, float* weights,
                     float* combined, void* stream);
/* Same with features computed OUTSIDE the kernel for the experts whose n_ch_host[e] == 0: ext_features [E][B][256] fp32
 * (only the blocks of those experts are read).  Used for extractors whose input vector does not fit the kernel's
 * shared-memory rows - NuScenesExpertExtractor, Linear(Q*(C+D), 512) on the flattened queries
 * (models/experts/expert_extractors.py:108-137); their slot of the parameter layout holds a zero-width first layer. */
int amoe_gate_fwd_ex2(amoe_ctx*, const float* state, const float* pooled, const float* params,
                      const void* params_bf16, int64_t n_params, int B, int E, const int* n_ch_host,
                      int ctx_dim, int hidden, float temperature, int mode, const float* ext_features,
                      float* context, float* features, float* processed, float* gate_logits,
                      float* weights, float* combined, void* stream);
/* y[b][q][:] = max(a[b][:] + c[q][:], 0), fp32, D % 4 == 0 (first decoder layer of the nuScenes multi-query head,
 * models/experts/nuscenes_expert.py:172-180, split into its per-frame and per-query halves). */
int amoe_bcast_add_relu(amoe_ctx*, const float* a, const float* c, float* y, int B, int Q, int D,
                        void* stream);

/* ---- policy head ------------------------------------------------------- */
/* EasyBackbone pool+fc and both TrajectoryPolicy MLP heads
 * (trajectory_head.py:25-33,44-63) after the 4 convs:
 *   x: [B,HW,Cf] (dtype) conv4 output   ctx: [B,ctx_dim] fp32 (combined) or NULL
 *   params: flat fp32 (fc, head_wp[0,2,4], head_spd[0,2,4] weight+bias in order)
 *   waypoints: [B,2*horizon] fp32   speed: [B,horizon] fp32 */
int amoe_policy_head_fwd(amoe_ctx*, const void* x, const float* ctx, const float* params,
                         int64_t n_params, int B, int HW, int Cf, int backbone_dim,
                         int ctx_dim, int hidden, int horizon, int x_dtype,
                         float* waypoints, float* speed, void* stream);

/* Same with an optional bf16 copy of the parameter buffer (same element offsets as `params`, 16-byte aligned).
 * params_bf16 != NULL and B >= 16 selects the bf16-inference variant: 16 frames per CTA, every layer with
 * K % 32 == 0 on mma.sync TF32 (bf16 weights, fp32 activations truncated to TF32, fp32 accumulation; biases
 * stay fp32).  x may be the pre-pooled feature [B,1,Cf] fp32 (HW = 1). */
int amoe_policy_head_fwd_ex(amoe_ctx*, const void* x, const float* ctx, const float* params,
                            int64_t n_params, int B, int HW, int Cf, int backbone_dim,
                            int ctx_dim, int hidden, int horizon, int x_dtype, const void* params_bf16,
                            float* waypoints, float* speed, void* stream);
/* AdaptiveAvgPool2d(1) of an NHWC bf16 tensor (EasyBackbone.pool, trajectory_head.py:30):
 *   x [B,HW,C] bf16 (C % 8 == 0) -> out [B,C] fp32, fixed summation order */
int amoe_mean_hw_nhwc_fwd(amoe_ctx*, const void* x, float* out, int B, int HW, int C, int dtype,
                          void* stream);

/* ---- Hungarian matcher ------------------------------------------------- */
/* Batched cost matrix of HungarianMatcher.forward (training/hungarian_matcher.py:34-76):
 *   cost[b,q,n] = w_bbox*L1(pb,tb) - w_class*softmax(logits)[q,label_n] - w_giou*GIoU
 * D==4: cxcywh boxes; D==7: BEV GIoU from (x,y,w,l); other D: L1 only.
 *   logits [B,Q,C] fp32, boxes [B,Q,D] fp32, tgt_boxes [B,Nmax,D] fp32 (padded),
 *   tgt_labels [B,Nmax] int64 (padded), n_tgt [B] int32, cost [B,Q,Nmax] fp32
 * Columns n >= n_tgt[b] are written as 0. */
int amoe_hungarian_cost_fwd(amoe_ctx*, const float* logits, const float* boxes,
                            const float* tgt_boxes, const int64_t* tgt_labels,
                            const int32_t* n_tgt, float* cost, int B, int Q, int C, int D,
                            int Nmax, float w_class, float w_bbox, float w_giou,
                            void* stream);
/* Host-side rectangular LSAP (shortest augmenting path, same algorithm and
 * tie-breaking as scipy.optimize.linear_sum_assignment, hungarian_matcher.py:79).
 *   cost_host [B,Q,Nmax] fp32 (row-major, only the first n_tgt[b] columns used)
 *   rows_host/cols_host [B,min(Q,Nmax)] int64, n_match_host[b] = min(Q,n_tgt[b])
 * returns -2 if a cost entry is NaN/-inf (scipy raises ValueError). */
int amoe_lsap_batched_host(const float* cost_host, const int32_t* n_tgt_host, int B, int Q,
                           int Nmax, int64_t* rows_host, int64_t* cols_host,
                           int32_t* n_match_host, int n_threads);

/* ---- training step of the gating / policy part -------------------------------
 * Replaces, for training/train_gating_network.py:76-117 (train_one_epoch), the autograd graph torch
 * builds over the trainable 2.87 M parameters (context extractor, expert extractors, GatingNetwork,
 * TrajectoryPolicy) - experts are frozen and run through the inference entry points above.  All fp32,
 * row-major [B, features]; every function is one or a few launches on `stream`, allocates nothing. */
/* y = Dropout_p(ReLU?(x W^T + b));  x [B,in] (row stride ldx), W [out,in], y [B,out] (row stride ldy).
 * Dropout is the inverted form nn.Dropout uses (kept units scaled by 1/(1-p)), keyed by (seed, b*out+o). */
int amoe_linear_fwd(amoe_ctx*, const float* x, int ldx, const float* W, const float* b, float* y,
                    int ldy, int B, int in_dim, int out_dim, int relu, float drop_p, uint64_t seed,
                    void* stream);
/* The same with the per-step part of the dropout key on the device: the mask is keyed by (seed + *seed_dev, b*out+o), so a
 * captured training step (CUDA graph) draws a new mask on every replay once amoe_train_tick has advanced *seed_dev
 * (train_gating_network.py:85 - model.train() keeps Dropout(0.1) active in every step). */
int amoe_linear_fwd_dseed(amoe_ctx*, const float* x, int ldx, const float* W, const float* b, float* y,
                          int ldy, int B, int in_dim, int out_dim, int relu, float drop_p, uint64_t seed,
                          const uint64_t* seed_dev, void* stream);
/* Gradients of the above.  With relu!=0 the mask of ReLU and Dropout together is y > 0 (g_tmp [B,out]
 * receives dy*[y>0]/(1-p)).  Any of dx/dW/db may be NULL.  dW [out,in], db [out]: overwritten. */
int amoe_linear_bwd(amoe_ctx*, const float* dy, int lddy, const float* y, int ldy, const float* x,
                    int ldx, const float* W, float* g_tmp, float* dx, int lddx, float* dW, float* db,
                    int B, int in_dim, int out_dim, int relu, float drop_p, void* stream);
/* nn.LayerNorm over the last dim (biased variance); mean/rstd [B] are saved for the backward. */
int amoe_layernorm_fwd(amoe_ctx*, const float* x, const float* gamma, const float* beta, float* y,
                       float* mean, float* rstd, int B, int D, float eps, void* stream);
int amoe_layernorm_bwd(amoe_ctx*, const float* dy, const float* x, const float* gamma,
                       const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta,
                       int B, int D, void* stream);
/* weights = softmax(logits/T) [B,E]; combined[b] = sum_e weights[b,e] * processed_e[b]
 * (models/gating/gating_network.py:157-165).  processed_e = processed + e*expert_stride, rows ld_p apart. */
int amoe_gate_combine_fwd(amoe_ctx*, const float* logits, const float* processed,
                          int64_t expert_stride, int ld_p, float temperature, float* weights,
                          float* combined, int B, int E, int P, void* stream);
/* use_softmax = 0: weights = sigmoid(logits) / (sum_e sigmoid(logits) + 1e-8), temperature unused
 * (gating_network.py:159-160). */
int amoe_gate_combine_fwd_ex(amoe_ctx*, const float* logits, const float* processed,
                             int64_t expert_stride, int ld_p, float temperature, int use_softmax,
                             float* weights, float* combined, int B, int E, int P, void* stream);
/* dcombined [B,P] and/or dweights [B,E] (direct gradient on the gate weights: load-balancing and
 * entropy losses) -> dlogits [B,E], dprocessed_e = dprocessed + e*dexpert_stride ([B,P] each). */
int amoe_gate_combine_bwd(amoe_ctx*, const float* dcombined, const float* dweights,
                          const float* weights, const float* processed, int64_t expert_stride,
                          int ld_p, float temperature, float* dlogits, float* dprocessed,
                          int64_t dexpert_stride, int B, int E, int P, void* stream);
/* logits_if_sigmoid != NULL selects the backward of the sigmoid gate (needs the forward's logits). */
int amoe_gate_combine_bwd_ex(amoe_ctx*, const float* dcombined, const float* dweights,
                             const float* weights, const float* processed, int64_t expert_stride,
                             int ld_p, float temperature, const float* logits_if_sigmoid,
                             float* dlogits, float* dprocessed, int64_t dexpert_stride, int B, int E,
                             int P, void* stream);
/* compute_gating_losses (training/train_gating_network.py:21-74): every loss term and the gradient of
 * total_loss w.r.t. the predictions, one launch.
 *   waypoints/tgt_waypoints [B,H,2]; speed [B,*] rows speed_ld apart, tgt_speed rows tgt_speed_ld apart
 *   speed_mode 0: no speed term; 1: L1 over the [B,H] sequences; 2: last step only
 *   coef_host[6] = ade, fde, speed, smoothness, load_balancing, entropy weights (HOST pointer)
 *   losses[7] (device) = total, ade, fde, speed, smoothness, load_balancing, entropy_loss
 *   d_waypoints [B,H,2], d_speed [B,speed_ld] (zero-filled by the caller; mode 2 writes the last column),
 *   d_weights [B,E]: d total / d prediction (may be NULL). */
int amoe_gating_loss_fwd_bwd(amoe_ctx*, const float* waypoints, const float* speed, int speed_ld,
                             const float* expert_weights, const float* tgt_waypoints,
                             const float* tgt_speed, int tgt_speed_ld, int B, int H, int E,
                             int speed_mode, const float* coef_host, int use_lb, int use_entropy,
                             float* losses, float* d_waypoints, float* d_speed, float* d_weights,
                             void* stream);
/* out2[0] = sum g^2, out2[1] = sqrt of it, over a flat fp32 buffer (clip_grad_norm_'s total norm);
 * partial_ws: ws_floats >= 1 scratch floats (more = more CTAs, up to 4 per SM). */
int amoe_sq_norm(amoe_ctx*, const float* g, int64_t n, float* partial_ws, int ws_floats,
                 float* out2, void* stream);
/* torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.AdamW.step() on flat buffers, one launch:
 * g' = g * grad_scale * min(1, max_norm / (norm2[1]*grad_scale + 1e-6)); decoupled weight decay;
 * bias-corrected moments with `step` counting from 1.  norm2 from amoe_sq_norm (device, NULL or
 * max_norm <= 0: no clipping).  grad_scale = 1/world_size after a SUM all-reduce. */
int amoe_fused_clip_adamw(amoe_ctx*, float* params, const float* grads, float* exp_avg,
                          float* exp_avg_sq, int64_t n, const float* norm2, float grad_scale,
                          float max_norm, float lr, float beta1, float beta2, float eps,
                          float weight_decay, int step, void* stream);
/* The same with the step count read from device memory (*step_dev >= 1), for a step replayed as a CUDA graph
 * (train_gating_network.py:103-105: clip_grad_norm_ + optimizer.step() once per iteration). */
int amoe_fused_clip_adamw_dstep(amoe_ctx*, float* params, const float* grads, float* exp_avg,
                                float* exp_avg_sq, int64_t n, const float* norm2, float grad_scale,
                                float max_norm, float lr, float beta1, float beta2, float eps,
                                float weight_decay, const int* step_dev, void* stream);
/* Advances the device-side counters of a training step by one: *step_dev += 1 (AdamW bias correction) and
 * *seed_dev += an odd 64-bit constant (dropout key); either may be NULL.  One single-thread launch. */
int amoe_train_tick(amoe_ctx*, int* step_dev, uint64_t* seed_dev, void* stream);
/* nn.BatchNorm2d in training mode on NHWC fp32 x [M = N*H*W, C]: batch statistics (biased variance for
 * the normalisation, unbiased for running_var), running stats updated in place with `momentum`
 * (NULL: not tracked), y = ReLU?(gamma * xhat + beta).  save_mean/save_rstd [C] feed the backward.
 * workspace: amoe_colreduce_workspace_floats(M, C) floats. */
int64_t amoe_colreduce_workspace_floats(int64_t M, int C);
int amoe_bn_train_fwd(amoe_ctx*, const float* x, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float eps, float* y,
                      float* save_mean, float* save_rstd, float* workspace, int64_t M, int C,
                      int relu, void* stream);
/* G (<= 4) BatchNorm layers of one shape in one launch per pass: x / y [G][M][C], gamma / beta / running_mean / running_var
 * are HOST arrays of G device pointers (the layers' own parameter and buffer tensors; running_* may be NULL),
 * save_mean / save_rstd [G*C], workspace: G * amoe_colreduce_workspace_floats(M, C) floats.  Same arithmetic as G calls of
 * amoe_bn_train_fwd (bit-identical).  The frozen experts of a gating-training step in the reference's train mode
 * (train_gating_network.py:85; torchvision resnet.py BasicBlock BatchNorm2d layers) run layer by layer in lockstep. */
int amoe_bn_train_fwd_grouped(amoe_ctx*, const float* x, const float* const* gamma, const float* const* beta,
                              float* const* running_mean, float* const* running_var, float momentum, float eps,
                              float* y, float* save_mean, float* save_rstd, float* workspace, int G, int64_t M,
                              int C, int relu, void* stream);
/* y = ReLU?(gamma*(x-mean)*rstd + beta) with given statistics (eval-mode BN kept differentiable). */
int amoe_bn_apply_fwd(amoe_ctx*, const float* x, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, float* y, int64_t M, int C, int relu,
                      void* stream);
/* Backward of either: y_relu != NULL masks dy with y > 0 first.  batch_stats=1: statistics were taken
 * from this batch (train mode); 0: constants (eval mode).  dgamma/dbeta [C] always written. */
int amoe_bn_bwd(amoe_ctx*, const float* dy, const float* x, const float* y_relu, const float* gamma,
                const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta,
                float* workspace, int64_t M, int C, int batch_stats, void* stream);
/* out[c] = scale * sum_m x[m,c]  (conv bias gradient; deterministic two-stage reduction). */
int amoe_colsum(amoe_ctx*, const float* x, float* out, float* workspace, int64_t M, int C,
                float scale, void* stream);
/* Convolution backward, NHWC fp32 (EasyBackbone convs): dx from dy and the packed weights
 * [Cout][KH][KW][Cin]; dw in the same packed layout from dy and x (K split over output pixels, summed
 * in a fixed order; workspace size from amoe_conv2d_bwd_weight_workspace_floats). */
int amoe_conv2d_bwd_data(amoe_ctx*, const float* dy, const float* w, float* dx, int B, int H, int W,
                         int Cin, int Cout, int KH, int KW, int stride_h, int stride_w, int pad_h,
                         int pad_w, int Ho, int Wo, void* stream);
int64_t amoe_conv2d_bwd_weight_workspace_floats(amoe_ctx*, int B, int Cin, int Cout, int KH, int KW,
                                                int Ho, int Wo);
int amoe_conv2d_bwd_weight(amoe_ctx*, const float* dy, const float* x, float* dw, float* workspace,
                           int64_t workspace_floats, int B, int H, int W, int Cin, int Cout, int KH,
                           int KW, int stride_h, int stride_w, int pad_h, int pad_w, int Ho, int Wo,
                           void* stream);
/* fp32-accurate training convolutions on the bf16 tensor cores (csrc/conv_tc.cu, "split operands"): every fp32 value is
 * carried as three bf16 parts (24 significant bits), a product keeps its six leading terms, all accumulated in fp32 TMEM.
 *   amoe_split3_bf16:              x [rows][C] fp32 -> [rows][3C] bf16 = (x1 | x2 | x3), C % 8 == 0
 *   amoe_pack_conv_weight_split6:  OIHW fp32 -> [Cout][KH][KW][6][Cin] bf16 (transposed = 1: [Cin][KH][KW][6][Cout] for dgrad)
 *   amoe_conv2d_fwd_f32tc:         y [B,Ho,Wo,Cout] fp32 = relu?(scale * conv(x) + bias)   (nn.Conv2d forward of the
 *                                  ResNet-18 / EasyBackbone layers inside the training step; Cin % 64 == 0, Cout % 32 == 0)
 *   amoe_conv2d_bwd_data_f32tc:    dx [B,H,W,Cin] fp32 from split dy [B,Ho,Wo,3*Cout] and the transposed split weights;
 *                                  stride 2 runs one sub-convolution per input parity class (the caller zero-fills dx when a
 *                                  class has no tap, i.e. 1x1 / stride 2); ones/zeros: Cin floats each. */
int amoe_split3_bf16(amoe_ctx*, const float* x, void* out, int64_t rows, int C, void* stream);
int amoe_pack_conv_weight_split6(amoe_ctx*, const float* w_oihw, void* dst, int Cout, int Cin, int KH,
                                 int KW, int transposed, void* stream);
int amoe_conv2d_f32tc_supported(int H, int W, int Cin, int Cout, int KH, int KW, int stride);
int amoe_conv2d_fwd_f32tc(amoe_ctx*, const void* x_split, const void* w_split, const float* scale,
                          const float* bias, float* y, int B, int H, int W, int Cin, int Cout, int KH,
                          int KW, int stride, int pad, int Ho, int Wo, int relu, void* stream);
/* Same forward for G stacked convolutions (activations [G*B,...] or shared [B,...], weights [G*Cout][KH][KW][6][Cin]) with an
 * optional fp32 residual added before the ReLU: the fp32 ("parity") inference mode of the ResNet trunks and heads. */
int amoe_conv2d_fwd_f32tc_grouped(amoe_ctx*, const void* x_split, const void* w_split, const float* scale,
                                  const float* bias, const float* residual, float* y, int G, int x_shared,
                                  int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                  int pad, int Ho, int Wo, int relu, void* stream);
int amoe_conv2d_bwd_data_f32tc(amoe_ctx*, const void* dy_split, const void* wT_split, const float* ones,
                               const float* zeros, float* dx, int B, int H, int W, int Cin, int Cout,
                               int KH, int KW, int stride, int pad, int Ho, int Wo, void* stream);
/* Weight gradient of a 3x3 / stride-1 / pad-1 convolution on the tensor cores, fp32-accurate (csrc/wgrad_tc.cu): a GEMM per
 * filter tap with the padded position grid as K; the operands are the padded NHWC tensors themselves (MN-major UMMA
 * operands), split in three bf16 parts per value:
 *   amoe_split3_padded:        x [NB,H,W,C] fp32 -> [NB,H+2,W+2,3C] bf16 = (x1 | x2 | x3), zero border (memset + kernel)
 *   amoe_conv3x3_wgrad_f32tc:  dy3 [NB,H+2,W+2,3*Cout], x3 [NB,H+2,W+2,3*Cin]; positions = NB*(H+2)*(W+2)
 *                              -> dw [Cout][3][3][Cin] fp32; K split over CTAs, partials summed in a fixed order
 *                              (workspace size from amoe_conv3x3_wgrad_f32tc_workspace_floats).
 *   Cin, Cout multiples of 64, <= 512. */
int amoe_split3_padded(amoe_ctx*, const float* x, void* out, int NB, int H, int W, int C, void* stream);
int amoe_conv3x3_wgrad_f32tc_supported(int Cin, int Cout);
int64_t amoe_conv3x3_wgrad_f32tc_workspace_floats(amoe_ctx*, int Cin, int Cout, int64_t positions);
int amoe_conv3x3_wgrad_f32tc(amoe_ctx*, const void* dy3, const void* x3, float* dw, float* workspace,
                             int64_t workspace_floats, int W, int Cin, int Cout, int64_t positions,
                             void* stream);
/* nn.AdaptiveAvgPool2d(1) on NHWC fp32: x [B,HW,C] -> out [B,C]; backward broadcasts dy/HW. */
int amoe_gap_fwd(amoe_ctx*, const float* x, float* out, int B, int HW, int C, void* stream);
int amoe_gap_bwd(amoe_ctx*, const float* dy, float* dx, int B, int HW, int C, void* stream);

/* ---- detection-expert training step (training/train_bdd100k_ddp.py:117-186) -----
 * The trunk/head forward+backward reuse the fp32 conv / BatchNorm entry points above; these add the
 * pieces only a full expert needs. */
/* backward of nn.MaxPool2d(3, stride 2, pad 1), NHWC fp32: x [NB,H,W,C], dy [NB,Ho,Wo,C] -> dx like x.
 * Arg-max rule of torch (first strictly greater element in kh-major scan order). */
int amoe_maxpool3x3s2_bwd(amoe_ctx*, const float* x, const float* dy, float* dx, int NB, int H,
                          int W, int C, void* stream);
/* The same gradient in two passes through a caller-owned workspace of NB*Ho*Wo*C bytes (Ho = (H-1)/2+1): the arg-max tap of
 * every window, then one comparison per (input pixel, containing window).  Bit-identical to amoe_maxpool3x3s2_bwd, 5x fewer
 * loads (torchvision resnet.py:199 maxpool inside training/train_bdd100k_ddp.py:117-186). */
int amoe_maxpool3x3s2_bwd_ws(amoe_ctx*, const float* x, const float* dy, float* dx, void* argmax_ws, int NB,
                             int H, int W, int C, void* stream);
/* BasicBlock tail: y = relu(a + b) (n % 4 == 0); g = dy * [y > 0] (the gradient of both addends). */
int amoe_add_relu_fwd(amoe_ctx*, const float* a, const float* b, float* y, int64_t n, void* stream);
int amoe_relu_bwd(amoe_ctx*, const float* dy, const float* y, float* g, int64_t n, void* stream);
/* Backward of amoe_upsample_bilinear_nchw_fwd (segmentation / drivable expert training,
 * train_bdd100k_ddp.py:188-194): dy [B,C,H,W] fp32 -> dlow [B,h,w,C] fp32. */
int amoe_upsample_bilinear_nchw_bwd(amoe_ctx*, const float* dy, float* dlow, int B, int h, int w,
                                    int C, int H, int W, void* stream);
/* Scatter of the Hungarian assignment (train_bdd100k_ddp.py:167-170) for the whole batch in one launch:
 * for matched pair m of image batch_of[m]: target_classes[batch_of[m]*Q + pred_idx[m]] = labels[m],
 * target_boxes[...] = boxes[m] (cxcywh).  The caller pre-fills classes with num_classes, boxes with 0. */
int amoe_det_targets(amoe_ctx*, const int64_t* pred_idx, const int32_t* batch_of,
                     const int64_t* labels, const float* boxes, int n_match, int Q,
                     int64_t* target_classes, float* target_boxes, void* stream);
/* nn.CrossEntropyLoss(ignore_index) + bbox_weight * nn.SmoothL1Loss(mean) over the matched rows
 * (train_bdd100k_ddp.py:172-185) and d(total)/d(logits), d(total)/d(boxes), one launch.
 *   logits [rows, ld_logits] (C classes), boxes [rows, ld_boxes] (4), target_classes [rows] int64,
 *   target_boxes [rows,4]; losses4 = total, class_loss, bbox_loss, #matched rows. */
int amoe_det_loss_fwd_bwd(amoe_ctx*, const float* logits, int ld_logits, const float* boxes,
                          int ld_boxes, const int64_t* target_classes, const float* target_boxes,
                          int64_t rows, int C, int ignore_index, float bbox_weight, float* losses4,
                          float* dlogits, int ld_dlogits, float* dboxes, int ld_dboxes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUTOMOE_B200_H_ */
