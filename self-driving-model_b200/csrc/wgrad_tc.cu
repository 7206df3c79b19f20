// Convolution weight gradient on the tensor cores, fp32-accurate (training/train_bdd100k_ddp.py:97 `loss.backward()`:
// the dW of every 3x3 / stride-1 / pad-1 ResNet convolution; 62 % of the detection-expert step on the CUDA cores).
//
//     dW[co][kh][kw][ci] = sum over output positions P of  dy[P][co] * x[P + (kh-1)*Wp + (kw-1)][ci]
//
// on the PHYSICALLY padded position grid (n, y, x) -> P = (n*Hp + y)*Wp + x with Hp = H+2, Wp = W+2 and zeros on the
// border, so a filter tap is a pure shift along P.  That makes dW a plain GEMM per tap with the positions as the K
// dimension, D[co][ci] = sum_P A[P][co] * B[P + shift][ci], whose operands are the padded NHWC tensors THEMSELVES, read
// as MN-major UMMA operands (M / N = channels contiguous, K = positions = rows): a TMA box of 64 positions x 64 channels
// is one swizzled 8 KB block, a 128-wide M tile is two such blocks (descriptor LBO = 8 KB, SBO = 1 KB between 8-row
// groups), and the tap shift is the box's ROW coordinate - no transposed copies.  (A K-major formulation needs the
// shift on the contiguous axis; TMA rejects start coordinates that are not 16-byte aligned there.)
// Both tensors are stored split in three bf16 parts per value, [P][x1 | x2 | x3] (amoe_split3_padded); the six product
// terms of the split (see conv_tc.cu) are six passes over the same K range, smallest terms first, fp32 TMEM accumulation.
//
// K is huge (0.5 M positions for a 720x1280 batch of 8 at layer1) and dW tiny, so the work is split along K: a tile is
// (k-split, tap, 128 output channels, <= 256 input channels); its partial D goes to a workspace
// [ksplit][9][Cout][Cin] and wgrad_reduce_kernel sums the splits in a fixed order (bit-reproducible, no atomics).
// Warp roles as conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue; two TMEM accumulators.
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace wg {

using namespace tc;

constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int TERMS = 6;

constexpr int BLK_BYTES = BLOCK_K * 128;       // one 64-position x 64-channel operand block

struct Params {
  int Cout, Cin, ntaps;
  int m_tiles, n_tiles, block_n;
  int ksplit, chunks_per_split, total_chunks;   // K chunks of 64 positions
  int stages;
  int total_tiles;
  int shift[9];                                 // position shift of every tap
  float* partial;                               // [ksplit][ntaps][Cout][Cin]
};

// MN-major, 128B-swizzled operand: rows = K (positions), 128 bytes = 64 channels per row; 8-row groups 1024 B apart (SBO),
// 64-channel blocks BLK_BYTES apart (LBO)
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(BLK_BYTES >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor of make_idesc with both operands MN-major (bits 15 / 16)
__device__ __forceinline__ uint32_t make_idesc_mn(int n) { return make_idesc(n) | (1u << 15) | (1u << 16); }

__device__ __constant__ int kXPart[TERMS] = {0, 0, 1, 0, 1, 2};   // dy part of term t (A operand)
__device__ __constant__ int kWPart[TERMS] = {0, 1, 0, 2, 1, 0};   // x part of term t (B operand)

struct Tile { int ks, tap, mt, nt; };
__device__ __forceinline__ Tile decode(const Params& p, int t) {
  Tile c;
  c.nt = t % p.n_tiles; t /= p.n_tiles;
  c.mt = t % p.m_tiles; t /= p.m_tiles;
  c.tap = t % p.ntaps;
  c.ks = t / p.ntaps;
  return c;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 4];
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t smem_a = smem_base, smem_b = smem_base + (uint32_t)p.stages * A_STAGE_BYTES;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[MAX_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * MAX_STAGES]), bar_tempty = smem_u32(&bars[2 * MAX_STAGES + 2]);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const Tile tl = decode(p, t);
        const int kc0 = tl.ks * p.chunks_per_split, kc1 = min(kc0 + p.chunks_per_split, p.total_chunks);
        const int shift = p.shift[tl.tap];
        for (int term = TERMS - 1; term >= 0; --term) {       // smallest terms first (see conv_tc.cu)
          const int acol = kXPart[term] * p.Cout + tl.mt * BLOCK_M;     // channel (column) offsets of this term's parts
          const int bcol = kWPart[term] * p.Cin + tl.nt * p.block_n;
          for (int kc = kc0; kc < kc1; ++kc) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * stage, A_STAGE_BYTES + b_stage_bytes);
            const int row = kc * BLOCK_K;
            tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmA, bar_full + 8 * stage, acol, row);
            tma_load_2d(smem_a + stage * A_STAGE_BYTES + BLK_BYTES, &tmA, bar_full + 8 * stage, acol + 64, row);
            for (int j = 0; j < p.block_n; j += 64)
              tma_load_2d(smem_b + stage * b_stage_bytes + (uint32_t)(j >> 6) * BLK_BYTES, &tmB, bar_full + 8 * stage, bcol + j,
                          row + shift);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_mn(p.block_n);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const Tile tl = decode(p, t);
      const int kc0 = tl.ks * p.chunks_per_split, kc1 = min(kc0 + p.chunks_per_split, p.total_chunks);
      const int k_iters = TERMS * max(kc1 - kc0, 0);
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(bar_tempty + 8 * as, aphase ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * ACC_STRIDE);
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(bar_full + 8 * stage, phase);
        tcgen05_fence_after();
        const uint64_t a_desc = make_mn_desc(smem_a + stage * A_STAGE_BYTES);
        const uint64_t b_desc = make_mn_desc(smem_b + stage * b_stage_bytes);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)     // 16 positions = 16 rows of 128 bytes further: +2048 B = +128 in the >>4 field
          umma_bf16(d_tmem, a_desc + (uint64_t)(kk * 128), b_desc + (uint64_t)(kk * 128), idesc, (uint32_t)((k | kk) != 0));
        umma_commit(bar_empty + 8 * stage);
        if (k == k_iters - 1) umma_commit(bar_tfull + 8 * as);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (k_iters <= 0) {   // empty K range (cannot happen with the host's split, kept for safety): nothing to accumulate
        umma_commit(bar_tfull + 8 * as);
      }
    }
  } else {
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const Tile tl = decode(p, t);
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int co = tl.mt * BLOCK_M + row;
      const bool valid = co < p.Cout;
      const int kc0 = tl.ks * p.chunks_per_split, kc1 = min(kc0 + p.chunks_per_split, p.total_chunks);
      float* dst = p.partial + ((((int64_t)tl.ks * p.ntaps + tl.tap) * p.Cout + co) * p.Cin + tl.nt * p.block_n);
      mbar_wait(bar_tfull + 8 * as, aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * ACC_STRIDE);
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            float4 o = make_float4(__uint_as_float(acc[v * 4]), __uint_as_float(acc[v * 4 + 1]), __uint_as_float(acc[v * 4 + 2]),
                                   __uint_as_float(acc[v * 4 + 3]));
            if (kc1 <= kc0) o = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dst + c0 + v * 4) = o;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dW[co][tap][ci] (the packed [Cout][KH][KW][Cin] fp32 layout of the training kernels) = sum over splits, fixed order
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ksplit,
                                                           int ntaps, int Cout, int Cin, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t r = i / Cin;
    const int tap = (int)(r % ntaps);
    const int co = (int)(r / ntaps);
    float s = 0.f;
    const int64_t per_split = (int64_t)ntaps * Cout * Cin;
    const int64_t off = ((int64_t)tap * Cout + co) * Cin + ci;
    for (int k = 0; k < ksplit; ++k) s += partial[(int64_t)k * per_split + off];
    dw[i] = s;
  }
}

// x [NB][H][W][C] fp32 -> out [NB][H+2][W+2][3C] bf16 = (x1 | x2 | x3) per pixel, physical zero border (the caller
// zero-fills `out`); one thread = 8 channels of one pixel
__global__ void __launch_bounds__(256) split3_padded_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int H, int W,
                                                            int C, int64_t total8) {
  const int c8n = C >> 3;
  const int Hp = H + 2, Wp = W + 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / c8n;
    const int c0 = (int)(i - pix * c8n) << 3;
    const int w = (int)(pix % W);
    const int64_t r = pix / W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(x + pix * C + c0));
    const float4 u1 = __ldg(reinterpret_cast<const float4*>(x + pix * C + c0 + 4));
    const float v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
    __align__(16) __nv_bfloat16 a[8], b[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tc::split3(v[j], a[j], b[j], c[j]);
    __nv_bfloat16* o = out + ((n * Hp + y + 1) * Wp + w + 1) * (3 * (int64_t)C) + c0;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(a);
    *reinterpret_cast<uint4*>(o + C) = *reinterpret_cast<const uint4*>(b);
    *reinterpret_cast<uint4*>(o + 2 * C) = *reinterpret_cast<const uint4*>(c);
  }
}

}  // namespace wg

int amoe_wgrad_tc_init(amoe_ctx* ctx) {
  AMOE_ENTER(ctx);
  (void)ctx;
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(wg::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BUDGET + 1024));
  return 0;
}

extern "C" {

int amoe_split3_padded(amoe_ctx* ctx, const float* x, void* out, int NB, int H, int W, int C, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && out, "amoe_split3_padded: NULL argument");
  AMOE_REQUIRE(NB > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "amoe_split3_padded: bad geometry (C %% 8 == 0)");
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "amoe_split3_padded: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  AMOE_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)NB * (H + 2) * (W + 2) * 3 * C * 2, st));
  const int64_t total8 = (int64_t)NB * H * W * (C / 8);
  const int64_t want = (total8 + 255) / 256;
  wg::split3_padded_kernel<<<(unsigned)std::min<int64_t>(want, (int64_t)ctx->sm_count * 16), 256, 0, st>>>(x, (__nv_bfloat16*)out, H, W, C, total8);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_conv3x3_wgrad_f32tc_supported(int Cin, int Cout) {
  return (Cin % 64 == 0 && Cout % 64 == 0 && Cin <= 512 && Cout <= 512) ? 1 : 0;
}

// K split: about two tiles per SM; every split non-empty
static void wgrad_split(const amoe_ctx* ctx, int Cin, int Cout, int64_t Ppad, int& ksplit, int& chunks_per_split) {
  const int m_tiles = (Cout + 127) / 128, block_n = Cin >= 256 ? 256 : Cin, n_tiles = Cin / block_n;
  const int base = 9 * m_tiles * n_tiles;
  const int64_t chunks = Ppad / 64;
  const int64_t want = std::max<int64_t>(1, std::min<int64_t>(chunks, (2 * ctx->sm_count + base - 1) / base));
  chunks_per_split = (int)((chunks + want - 1) / want);
  ksplit = (int)((chunks + chunks_per_split - 1) / chunks_per_split);
}

int64_t amoe_conv3x3_wgrad_f32tc_workspace_floats(amoe_ctx* ctx, int Cin, int Cout, int64_t positions) {
  if (!ctx) return -1;
  int ksplit, cps;
  wgrad_split(ctx, Cin, Cout, (positions + 63) / 64 * 64, ksplit, cps);
  return (int64_t)ksplit * 9 * Cout * Cin;
}

// dy3 [NB][H+2][W+2][3*Cout], x3 [NB][H+2][W+2][3*Cin] bf16 (amoe_split3_padded of dy / x of a 3x3, stride 1, pad 1
// convolution); positions = NB*(H+2)*(W+2) -> dw [Cout][3][3][Cin] fp32
int amoe_conv3x3_wgrad_f32tc(amoe_ctx* ctx, const void* dy3, const void* x3, float* dw, float* workspace, int64_t workspace_floats,
                             int W, int Cin, int Cout, int64_t positions, void* stream) {
  AMOE_ENTER(ctx);
  using namespace wg;
  AMOE_REQUIRE(ctx && dy3 && x3 && dw && workspace, "amoe_conv3x3_wgrad_f32tc: NULL argument");
  AMOE_REQUIRE(amoe_conv3x3_wgrad_f32tc_supported(Cin, Cout) && positions > 0, "amoe_conv3x3_wgrad_f32tc: unsupported shape");
  AMOE_REQUIRE(workspace_floats >= amoe_conv3x3_wgrad_f32tc_workspace_floats(ctx, Cin, Cout, positions), "amoe_conv3x3_wgrad_f32tc: workspace too small");
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(dy3) & 15) == 0 && (reinterpret_cast<uintptr_t>(x3) & 15) == 0, "amoe_conv3x3_wgrad_f32tc: unaligned operand");
  const int64_t Ppad = (positions + 63) / 64 * 64;
  Params p;
  p.Cout = Cout; p.Cin = Cin; p.ntaps = 9;
  p.m_tiles = (Cout + 127) / 128;
  p.block_n = Cin >= 256 ? 256 : Cin;
  p.n_tiles = Cin / p.block_n;
  p.total_chunks = (int)(Ppad / 64);
  wgrad_split(ctx, Cin, Cout, Ppad, p.ksplit, p.chunks_per_split);
  p.total_tiles = p.ksplit * 9 * p.m_tiles * p.n_tiles;
  const int Wp = W + 2;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) p.shift[kh * 3 + kw] = (kh - 1) * Wp + (kw - 1);
  p.partial = workspace;
  const int stage_bytes = A_STAGE_BYTES + p.block_n * 128;
  p.stages = std::min(MAX_STAGES, SMEM_BUDGET / stage_bytes);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)3 * Cout, (cuuint64_t)positions};
    cuuint64_t strides[1] = {(cuuint64_t)3 * Cout * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)BLOCK_K};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy3), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "amoe_conv3x3_wgrad_f32tc: cuTensorMapEncodeTiled(dy) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)3 * Cin, (cuuint64_t)positions};
    cuuint64_t strides[1] = {(cuuint64_t)3 * Cin * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)BLOCK_K};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x3), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "amoe_conv3x3_wgrad_f32tc: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = std::min(p.total_tiles, ctx->sm_count);
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  wgrad_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, p);
  AMOE_LAUNCH_OK(ctx);
  const int64_t total = (int64_t)Cout * 9 * Cin;
  const int64_t want = (total + 255) / 256;
  wgrad_reduce_kernel<<<(unsigned)std::min<int64_t>(want, (int64_t)ctx->sm_count * 16), 256, 0, st>>>(workspace, dw, p.ksplit, 9, Cout, Cin, total);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
