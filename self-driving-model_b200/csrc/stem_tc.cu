// First-layer convolutions (Cin = 3: ResNet stem 7x7/s2/p3 of every expert + policy conv1
// 5x5/s2/p2) as ONE tensor-core GEMM that reads the raw image rows - no im2col copy anywhere.
//
// The frame is staged as [B][H+6][Wpad][4] bf16 (3 zero rows on top/bottom, 4 zero pixels left,
// zeros right).  For output pixel (oh, ow) and filter row kh the needed input is the window of 8
// pixels x 4 channels = 32 bf16 starting at padded pixel 2*ow of padded row 2*oh+kh, i.e.
//     A[ow][k] = row[8*ow + k]        (byte address = row_base + 16*ow + 2*k)
// - consecutive output pixels read OVERLAPPING windows.  A K-major SWIZZLE_NONE UMMA operand is
// made of 8x16B core matrices with a programmable stride between 8-row groups (SBO) and between
// the two 16-byte K chunks of one MMA (LBO); SBO = 128 B and LBO = 16 B make the tensor core read
// exactly this Toeplitz matrix straight from the image row in shared memory (verified on B200 by
// tools/probe_umma_nosw.cu).  So a tile (one output row, 128 pixels) needs 7 image rows = 15 KB of
// shared memory instead of 7 x 16 KB of expanded windows, and K is 7*32 = 224 (147 useful).
//
// B operand: all filters of all convolutions sharing the frame, concatenated on N (3 experts x 64
// + policy 32 = 224), resident in shared memory for the whole kernel in the no-swizzle layout
// [K/8][N][8] (LBO = N*16 B, SBO = 128 B); filter taps sit at window position kw+1 (stem) /
// kw+2, row kh+1 (policy), zeros elsewhere.
//
// Warp roles as in conv_tc.cu: warp 0 producer (bulk copies of 7 contiguous image rows), warp 1
// MMA issuer (14 tcgen05.mma of N=n_total per tile), warps 2-5 epilogue (BN scale/bias + ReLU,
// scatter of 32-channel chunks to their destination tensors).
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace stem {

using namespace tc;

constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;
constexpr int A_STAGES = 4;
constexpr int MAX_CHUNKS = 8;
constexpr int WIN_K = 32;  // bf16 per filter row window (8 pixels x 4 channels)

struct Params {
  int B, Ho, Wo, Hpad, KH;
  int row_bytes;       // Wpad * 8
  int a_stage_bytes;   // KH * row_bytes + slack for the garbage rows ow >= Wo, multiple of 128
  int n_total;         // multiple of 32, <= 256
  int w_bytes;         // KH*4 * n_total * 16
  int total_tiles;     // B * Ho
  int relu;
  const uint8_t* x;
  const uint8_t* w;
  const float* scale;
  const float* bias;
  __nv_bfloat16* dst[MAX_CHUNKS];  // per 32-channel chunk: destination (sub-tensor + channel offset applied)
  int dst_c[MAX_CHUNKS];           // channels per pixel of that destination
};

__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100); layout type 0 = SWIZZLE_NONE
  return d;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1) stem_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * A_STAGES + 5];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_scale[256];
  __shared__ __align__(16) float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t smem_w = smem_base;
  const uint32_t smem_a = smem_base + (uint32_t)((p.w_bytes + 127) & ~127);
  const uint32_t bar_afull = smem_u32(&bars[0]);
  const uint32_t bar_aempty = smem_u32(&bars[A_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * A_STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * A_STAGES + 2]);
  const uint32_t bar_w = smem_u32(&bars[2 * A_STAGES + 4]);

  for (int i = threadIdx.x; i < p.n_total; i += NUM_THREADS) {
    s_scale[i] = __ldg(p.scale + i);
    s_bias[i] = __ldg(p.bias + i);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const uint32_t tile_bytes = (uint32_t)(p.KH * p.row_bytes);

  if (warp == 0) {
    // ============================ producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, (uint32_t)p.w_bytes);
      // weights: a few bulk copies (each <= 32 KB keeps the request size modest)
      for (int o = 0; o < p.w_bytes; o += 32768)
        bulk_g2s(smem_w + (uint32_t)o, p.w + o, (uint32_t)std::min(32768, p.w_bytes - o), bar_w);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int b = t / p.Ho, oh = t - b * p.Ho;
        mbar_wait(bar_aempty + 8 * stage, phase ^ 1u);
        mbar_arrive_expect_tx(bar_afull + 8 * stage, tile_bytes);
        // KH contiguous padded rows starting at row 2*oh of image b
        bulk_g2s(smem_a + (uint32_t)stage * p.a_stage_bytes, p.x + ((int64_t)b * p.Hpad + 2 * oh) * p.row_bytes,
                 tile_bytes, bar_afull + 8 * stage);
        if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==========================
    const uint32_t idesc = make_idesc(p.n_total);
    const uint32_t lbo_b = (uint32_t)p.n_total * 16u;
    mbar_wait(bar_w, 0);
    tcgen05_fence_after();
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(bar_tempty + 8 * acc, tphase ^ 1u);
      mbar_wait(bar_afull + 8 * stage, phase);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
      const uint32_t a_base = smem_a + (uint32_t)stage * p.a_stage_bytes;
      for (int kh = 0; kh < p.KH; ++kh) {
#pragma unroll
        for (int ks = 0; ks < WIN_K / UMMA_K; ++ks) {
          const uint64_t a_desc = make_nosw_desc(a_base + (uint32_t)(kh * p.row_bytes + ks * 32), 16u, 128u);
          const uint64_t b_desc = make_nosw_desc(smem_w + (uint32_t)(kh * 4 + ks * 2) * lbo_b, lbo_b, 128u);
          umma_bf16(d_tmem, a_desc, b_desc, idesc, (uint32_t)((kh | ks) != 0));
        }
      }
      umma_commit(bar_aempty + 8 * stage);
      umma_commit(bar_tfull + 8 * acc);
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ============================ epilogue ============================
    const int lg = warp & 3;
    const int ow = lg * 32 + lane;
    const bool valid = ow < p.Wo;
    const int n_chunks = p.n_total >> 5;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      const int64_t pix = (int64_t)t * p.Wo + ow;  // (b*Ho + oh)*Wo + ow
      mbar_wait(bar_tfull + 8 * acc, tphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t a[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), a);
        tmem_ld_wait();
        if (valid) {
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c * 32);
          const float4* bs4 = reinterpret_cast<const float4*>(s_bias + c * 32);
          __nv_bfloat16* o = p.dst[c] + pix * p.dst_c[c];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 s0 = sc4[2 * v], s1 = sc4[2 * v + 1], b0 = bs4[2 * v], b1 = bs4[2 * v + 1];
            float f[8];
            f[0] = fmaf(__uint_as_float(a[v * 8 + 0]), s0.x, b0.x);
            f[1] = fmaf(__uint_as_float(a[v * 8 + 1]), s0.y, b0.y);
            f[2] = fmaf(__uint_as_float(a[v * 8 + 2]), s0.z, b0.z);
            f[3] = fmaf(__uint_as_float(a[v * 8 + 3]), s0.w, b0.w);
            f[4] = fmaf(__uint_as_float(a[v * 8 + 4]), s1.x, b1.x);
            f[5] = fmaf(__uint_as_float(a[v * 8 + 5]), s1.y, b1.y);
            f[6] = fmaf(__uint_as_float(a[v * 8 + 6]), s1.z, b1.z);
            f[7] = fmaf(__uint_as_float(a[v * 8 + 7]), s1.w, b1.w);
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            *reinterpret_cast<uint4*>(o + v * 8) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                               pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// fp32-accurate variant for the training forward (the reference trains in fp32): the same Toeplitz GEMM with every operand
// split into three bf16 parts (v = v1 + v2 + v3, tc_common.cuh split3) and the six leading product terms accumulated in the
// fp32 TMEM accumulator, smallest terms first - the scheme of the split-operand convolutions in conv_tc.cu.  The frame parts
// are three padded bf16 frames stacked on the batch axis, the filter parts three [K/8][N][8] images; a tile (one output row)
// stages 3 x KH image rows and runs 6 x KH x 2 MMAs.  fp32 output [B][Ho][Wo][N], scale/bias applied in fp32.
constexpr int F32_TERMS = 6;
__device__ __constant__ int kStemXPart[F32_TERMS] = {0, 0, 1, 0, 1, 2};
__device__ __constant__ int kStemWPart[F32_TERMS] = {0, 1, 0, 2, 1, 0};

struct F32Params {
  Params base;             // dst[] unused
  int stages;              // activation stages (each holds the three parts)
  int64_t part_bytes;      // distance between the frame parts: B * Hpad * row_bytes
  float* y;                // [B][Ho][Wo][n_total] fp32
};

__global__ void __launch_bounds__(NUM_THREADS, 1) stem_tc_f32_kernel(const __grid_constant__ F32Params fp) {
  const Params& p = fp.base;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * A_STAGES + 5];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_scale[256];
  __shared__ __align__(16) float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t w_part = (uint32_t)((p.w_bytes + 127) & ~127);
  const uint32_t smem_w = smem_base;
  const uint32_t smem_a = smem_base + 3u * w_part;
  const uint32_t stage_bytes = 3u * (uint32_t)p.a_stage_bytes;
  const uint32_t bar_afull = smem_u32(&bars[0]);
  const uint32_t bar_aempty = smem_u32(&bars[A_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * A_STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * A_STAGES + 2]);
  const uint32_t bar_w = smem_u32(&bars[2 * A_STAGES + 4]);

  for (int i = threadIdx.x; i < p.n_total; i += NUM_THREADS) {
    s_scale[i] = __ldg(p.scale + i);
    s_bias[i] = __ldg(p.bias + i);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const uint32_t tile_bytes = (uint32_t)(p.KH * p.row_bytes);

  if (warp == 0) {
    // ============================ producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, 3u * (uint32_t)p.w_bytes);
      for (int part = 0; part < 3; ++part)
        for (int o = 0; o < p.w_bytes; o += 32768)
          bulk_g2s(smem_w + (uint32_t)part * w_part + (uint32_t)o, p.w + (size_t)part * p.w_bytes + o,
                   (uint32_t)std::min(32768, p.w_bytes - o), bar_w);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int b = t / p.Ho, oh = t - b * p.Ho;
        mbar_wait(bar_aempty + 8 * stage, phase ^ 1u);
        mbar_arrive_expect_tx(bar_afull + 8 * stage, 3u * tile_bytes);
        for (int part = 0; part < 3; ++part)
          bulk_g2s(smem_a + (uint32_t)stage * stage_bytes + (uint32_t)part * p.a_stage_bytes,
                   p.x + (int64_t)part * fp.part_bytes + ((int64_t)b * p.Hpad + 2 * oh) * p.row_bytes, tile_bytes, bar_afull + 8 * stage);
        if (++stage == fp.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==========================
    const uint32_t idesc = make_idesc(p.n_total);
    const uint32_t lbo_b = (uint32_t)p.n_total * 16u;
    mbar_wait(bar_w, 0);
    tcgen05_fence_after();
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(bar_tempty + 8 * acc, tphase ^ 1u);
      mbar_wait(bar_afull + 8 * stage, phase);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
      const uint32_t a_stage = smem_a + (uint32_t)stage * stage_bytes;
      // the tensor core truncates when it adds into the accumulator: the 2^-16 and 2^-8 terms go in while the sum is small
      for (int term = F32_TERMS - 1; term >= 0; --term) {
        const uint32_t a_base = a_stage + (uint32_t)kStemXPart[term] * (uint32_t)p.a_stage_bytes;
        const uint32_t w_base = smem_w + (uint32_t)kStemWPart[term] * w_part;
        for (int kh = 0; kh < p.KH; ++kh) {
#pragma unroll
          for (int ks = 0; ks < WIN_K / UMMA_K; ++ks) {
            const uint64_t a_desc = make_nosw_desc(a_base + (uint32_t)(kh * p.row_bytes + ks * 32), 16u, 128u);
            const uint64_t b_desc = make_nosw_desc(w_base + (uint32_t)(kh * 4 + ks * 2) * lbo_b, lbo_b, 128u);
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (uint32_t)((term != F32_TERMS - 1) | kh | ks));
          }
        }
      }
      umma_commit(bar_aempty + 8 * stage);
      umma_commit(bar_tfull + 8 * acc);
      if (++stage == fp.stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ============================ epilogue ============================
    const int lg = warp & 3;
    const int ow = lg * 32 + lane;
    const bool valid = ow < p.Wo;
    const int n_chunks = p.n_total >> 5;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      const int64_t pix = (int64_t)t * p.Wo + ow;  // (b*Ho + oh)*Wo + ow
      mbar_wait(bar_tfull + 8 * acc, tphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t a[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), a);
        tmem_ld_wait();
        if (valid) {
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c * 32);
          const float4* bs4 = reinterpret_cast<const float4*>(s_bias + c * 32);
          float* o = fp.y + pix * p.n_total + c * 32;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 s0 = sc4[2 * v], s1 = sc4[2 * v + 1], b0 = bs4[2 * v], b1 = bs4[2 * v + 1];
            float f[8];
            ffma2(f[0], f[1], a[v * 8 + 0], a[v * 8 + 1], s0.x, s0.y, b0.x, b0.y);
            ffma2(f[2], f[3], a[v * 8 + 2], a[v * 8 + 3], s0.z, s0.w, b0.z, b0.w);
            ffma2(f[4], f[5], a[v * 8 + 4], a[v * 8 + 5], s1.x, s1.y, b1.x, b1.y);
            ffma2(f[6], f[7], a[v * 8 + 6], a[v * 8 + 7], s1.z, s1.w, b1.z, b1.w);
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            stg256(o + v * 8, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])),
                   make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Variant with the ResNet max-pool (3x3, stride 2, pad 1) fused behind the stem: the first
// n_pool_ch channels (expert stems) never reach HBM at full resolution.  A CTA owns a contiguous
// range of POOLED rows and walks the conv rows they need in order (2py-1 once as "carry", then
// 2py and 2py+1 for every pooled row); the pooling itself is described at the kernel.  Remaining
// channels (policy conv1) are stored at full resolution as in the plain kernel.
constexpr int POOL_A_STAGES = 4;
constexpr int POOL_THREADS = 320;   // warp 0 producer, warp 1 MMA issuer, warps 2-9 epilogue + pooling
constexpr int POOL_EPI_THREADS = 256;
constexpr int R_PITCH = 400;  // bytes per pixel row of the staging buffer (192 ch * 2 B + pad: conflict-free 16-byte stores)

struct PoolParams {
  Params base;
  int n_pool_ch;          // leading channels that are max-pooled (multiple of 64)
  int Hp, Wp, out_pad;    // pooled size, border of the pooled output tensor
  __nv_bfloat16* pooled;  // [n_pool_ch/64 * B][Hp+2*out_pad][Wp+2*out_pad][64]
  int dbg;                // AMOE_STEM_DBG experiment bits (results wrong on purpose): 1 no horizontal pass, 2 no full-resolution
                          // stores, 4 no staging stores, 8 no border zeroing, 16 no CTA-wide barriers
};

enum { ROLE_CARRY = 0, ROLE_EVEN = 1, ROLE_ODD = 2 };

struct Seq {
  int p_lo, p_hi, has_carry, n_tiles;
};
__device__ __forceinline__ Seq make_seq(int p_total, int Hp) {
  Seq s;
  s.p_lo = (int)((int64_t)blockIdx.x * p_total / gridDim.x);
  s.p_hi = (int)((int64_t)(blockIdx.x + 1) * p_total / gridDim.x);
  s.has_carry = (s.p_hi > s.p_lo && (s.p_lo % Hp) != 0) ? 1 : 0;
  s.n_tiles = s.has_carry + 2 * (s.p_hi - s.p_lo);
  return s;
}
__device__ __forceinline__ void tile_at(const Seq& s, int k, int Hp, int& b, int& py, int& oh, int& role) {
  if (s.has_carry && k == 0) {
    b = s.p_lo / Hp; py = s.p_lo - b * Hp; oh = 2 * py - 1; role = ROLE_CARRY;
  } else {
    const int kk = k - s.has_carry;
    const int gp = s.p_lo + (kk >> 1);
    b = gp / Hp; py = gp - b * Hp; oh = 2 * py + (kk & 1); role = (kk & 1) ? ROLE_ODD : ROLE_EVEN;
  }
}
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}
// folded BN (+ optional ReLU) of 32 accumulator columns -> 4 x 8 packed bf16
template <bool RELU>
__device__ __forceinline__ void bn_pack32(const uint32_t (&a)[32], const float* s_scale, const float* s_bias, uint4 (&q)[4]) {
  const float4* sc4 = reinterpret_cast<const float4*>(s_scale);
  const float4* bs4 = reinterpret_cast<const float4*>(s_bias);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 s0 = sc4[2 * v], s1 = sc4[2 * v + 1], b0 = bs4[2 * v], b1 = bs4[2 * v + 1];
    float f[8];
    ffma2(f[0], f[1], a[v * 8 + 0], a[v * 8 + 1], s0.x, s0.y, b0.x, b0.y);
    ffma2(f[2], f[3], a[v * 8 + 2], a[v * 8 + 3], s0.z, s0.w, b0.z, b0.w);
    ffma2(f[4], f[5], a[v * 8 + 4], a[v * 8 + 5], s1.x, s1.y, b1.x, b1.y);
    ffma2(f[6], f[7], a[v * 8 + 6], a[v * 8 + 7], s1.z, s1.w, b1.z, b1.w);
    if (RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    q[v] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

// 32 accumulator columns -> 4 x 8 packed bf16 (folded mode: nothing left to apply)
__device__ __forceinline__ void pack32(const uint32_t (&a)[32], uint4 (&q)[4]) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
    q[v] = make_uint4(pack_bf16x2(__uint_as_float(a[v * 8 + 0]), __uint_as_float(a[v * 8 + 1])),
                      pack_bf16x2(__uint_as_float(a[v * 8 + 2]), __uint_as_float(a[v * 8 + 3])),
                      pack_bf16x2(__uint_as_float(a[v * 8 + 4]), __uint_as_float(a[v * 8 + 5])),
                      pack_bf16x2(__uint_as_float(a[v * 8 + 6]), __uint_as_float(a[v * 8 + 7])));
}

// Horizontal 3-max at even columns of the staged vertical max (sR: [128 conv columns][R_PITCH]), ReLU, store of pooled row
// (b, py); then the zero border of the padded pooled tensor around that row.  Executed by `nthr` threads (index te) of one role.
template <bool RUNS>
__device__ __forceinline__ void pool_row_out(const PoolParams& pp, const uint8_t* sR, int b, int py, int te, int nthr, int dbg) {
  const Params& p = pp.base;
  const int NV = pp.n_pool_ch >> 3;           // 16-byte channel vectors per pooled pixel
  const int Hq = pp.Hp + 2 * pp.out_pad, Wq = pp.Wp + 2 * pp.out_pad;
  if (!(dbg & 1)) {
    // mapping: thread -> (pixel lane, channel vector v).  RUNS: a pixel lane owns a contiguous run of pooled pixels and
    // carries the odd column 2px+1 over to pixel px+1 - two staged vectors read per output instead of three (the staging
    // traffic competes with the MMAs' operand reads for shared-memory bandwidth).
    const int px_lanes = nthr / NV;                  // 128 threads: 16 / 8 / 5 for 1 / 2 / 3 experts
    const int h_px0 = te / NV, h_v = te - h_px0 * NV;
    if (h_px0 < px_lanes) {
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      if (RUNS) {
        const int run = (pp.Wp + px_lanes - 1) / px_lanes;
        const int px_lo = h_px0 * run, px_hi = min(pp.Wp, px_lo + run);
        if (px_lo < px_hi) {
          __nv_bfloat16* g = pp.pooled + (((int64_t)(h_v >> 3) * p.B + b) * Hq + py + pp.out_pad) * (int64_t)Wq * 64 +
                             (px_lo + pp.out_pad) * 64 + (h_v & 7) * 8;
          const uint8_t* r0 = sR + (2 * px_lo) * R_PITCH + h_v * 16;
          uint4 prev = px_lo > 0 ? *reinterpret_cast<const uint4*>(r0 - R_PITCH) : zero;   // packed bf16: 0 is only used when ...
          const bool has_prev0 = px_lo > 0;
#pragma unroll 4
          for (int px = px_lo; px < px_hi; ++px, r0 += 2 * R_PITCH, g += 64) {
            const uint4 e = *reinterpret_cast<const uint4*>(r0), o = *reinterpret_cast<const uint4*>(r0 + R_PITCH);
            uint4 m = hmax8(e, o);
            if (px > px_lo || has_prev0) m = hmax8(m, prev);   // ... a left neighbour exists
            prev = o;
            if (p.relu) m = hmax8(m, zero);
            *reinterpret_cast<uint4*>(g) = m;
          }
        }
      } else {
        __nv_bfloat16* g = pp.pooled + (((int64_t)(h_v >> 3) * p.B + b) * Hq + py + pp.out_pad) * (int64_t)Wq * 64 +
                           (h_px0 + pp.out_pad) * 64 + (h_v & 7) * 8;
        const uint8_t* r0 = sR + (2 * h_px0) * R_PITCH + h_v * 16;
        const int h_rstep = 2 * px_lanes * R_PITCH, h_gstep = px_lanes * 64;
#pragma unroll 2
        for (int px = h_px0; px < pp.Wp; px += px_lanes, r0 += h_rstep, g += h_gstep) {
          uint4 m = hmax8(*reinterpret_cast<const uint4*>(r0), *reinterpret_cast<const uint4*>(r0 + R_PITCH));
          if (px > 0) m = hmax8(m, *reinterpret_cast<const uint4*>(r0 - R_PITCH));
          if (p.relu) m = hmax8(m, zero);
          *reinterpret_cast<uint4*>(g) = m;
        }
      }
    }
  }
}
__device__ __forceinline__ void pool_row_border(const PoolParams& pp, int b, int py, int te, int nthr) {
  const Params& p = pp.base;
  const int NV = pp.n_pool_ch >> 3;
  const int Hq = pp.Hp + 2 * pp.out_pad, Wq = pp.Wp + 2 * pp.out_pad;
  // zero border of the padded pooled tensor: left/right pixel of this row, plus the rows above/below the image
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int item = te; item < 2 * NV; item += nthr) {
    const int side = item / NV, v = item - side * NV, e = v >> 3, cv = v & 7;
    __nv_bfloat16* d = pp.pooled + ((((int64_t)e * p.B + b) * Hq + py + 1) * Wq + (side ? Wq - 1 : 0)) * 64 + cv * 8;
    *reinterpret_cast<uint4*>(d) = z;
  }
  if (py == 0 || py == pp.Hp - 1) {
    const int row = (py == 0) ? 0 : Hq - 1;
    for (int item = te; item < Wq * NV; item += nthr) {
      const int x = item / NV, v = item - x * NV, e = v >> 3, cv = v & 7;
      __nv_bfloat16* d = pp.pooled + ((((int64_t)e * p.B + b) * Hq + row) * Wq + x) * 64 + cv * 8;
      *reinterpret_cast<uint4*>(d) = z;
    }
    if (pp.Hp == 1) {  // single pooled row: both borders
      for (int item = te; item < Wq * NV; item += nthr) {
        const int x = item / NV, v = item - x * NV, e = v >> 3, cv = v & 7;
        __nv_bfloat16* d = pp.pooled + ((((int64_t)e * p.B + b) * Hq + Hq - 1) * Wq + x) * 64 + cv * 8;
        *reinterpret_cast<uint4*>(d) = z;
      }
    }
  }
}

// Channels that are not pooled (policy conv1): full-resolution store of this warp's 32 output pixels of conv row (b, oh)
template <bool FOLDED>
__device__ __forceinline__ void store_rest_chunks(const Params& p, uint32_t taddr, int pool_chunks, int n_chunks, int b, int oh, int ow,
                                                  bool store, const float* s_scale, const float* s_bias) {
  const int64_t pix = ((int64_t)b * p.Ho + oh) * p.Wo + ow;
  uint32_t a[32];
  for (int c = pool_chunks; c < n_chunks; ++c) {
    tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), a);
    tmem_ld_wait();
    if (store) {
      uint4 q[4];
      if (FOLDED) {
        pack32(a, q);
        if (p.relu) {
          const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int v = 0; v < 4; ++v) q[v] = hmax8(q[v], zero);
        }
      } else if (p.relu) {
        bn_pack32<true>(a, s_scale + c * 32, s_bias + c * 32, q);
      } else {
        bn_pack32<false>(a, s_scale + c * 32, s_bias + c * 32, q);
      }
      // 32-byte stores: with 16-byte ones every instruction writes half sectors at a 64-byte stride (32-channel pixels)
      __nv_bfloat16* d = p.dst[c] + pix * p.dst_c[c];
      if ((reinterpret_cast<uintptr_t>(d) & 31) == 0) {
        stg256(d, q[0], q[1]);
        stg256(d + 16, q[2], q[3]);
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) *reinterpret_cast<uint4*>(d + v * 8) = q[v];
      }
    }
  }
}

// Epilogue organisation: the eight epilogue warps form two groups; warps g*4+lq of both groups read
// TMEM lane quarter lq (output pixels 32*lq..+31) and split the 32-channel chunks between them
// (group 0: first half of the pooled chunks; group 1: second half + the full-resolution chunks).
// The vertical 3-max runs first and entirely in registers: V (one packed bf16 row segment per owned
// chunk) carries conv row 2py-1 into pooled row py, conv row 2py folds in, and conv row 2py+1 both
// completes the pooled row and becomes the next carry.  Only the completed vertical max goes through
// shared memory, once per POOLED row, for the horizontal 3-max at even columns.
// FOLDED: the BatchNorm scale is folded into the packed filters and the bias rides on the frame's padding
// channel (staged as 1.0; filter slots (kh=0, j=0/1, c=3) hold the bias split into two bf16 parts), so the
// accumulator already is scale*conv+bias and the epilogue only packs, pools and applies the ReLU.
// NGRP: epilogue warp groups of four warps (one per TMEM lane quarter) that split the 32-channel chunks between them.
// NGRP = 2 is the 8-warp epilogue of round 1 (3-4 chunks per warp, 162 registers); NGRP = 4 gives every warp at most two
// chunks so that sixteen epilogue warps fit the register file (<= 112 registers per thread at 576 threads per CTA).
// POOLW: warps (0 or 4) that take the horizontal pass + stores of the pooled rows off the draining warps: the drain warps
// stage the vertical max of pooled row py (bar.arrive "full"), the pooling warps turn it into the output row and hand the
// staging buffer back (bar.arrive "empty") while the next two conv rows are drained.  Measured on B200 (tools/stem_bench.py):
// the in-line horizontal pass costs 108 of the kernel's 450 us.
// (Measured and removed: the horizontal 3-max in registers - two warp shuffles per packed register, even lanes storing pooled
// pixels straight to global memory, only the lane-quarter boundary columns through shared memory.  Bit-identical, but 654 us
// against 337 us: ~100 more instructions per half-chunk on the two draining warps of a sub-partition, whose issue latency
// is what the TMEM drain hides behind, and 16-byte stores at a 128-byte stride.)
template <bool FOLDED, int NGRP, int POOLW>
__global__ void __launch_bounds__(64 + NGRP * 128 + POOLW * 32, 1) stem_pool_kernel(const __grid_constant__ PoolParams pp) {
  constexpr int EPI_THREADS = NGRP * 128;
  constexpr int SYNC_THREADS = EPI_THREADS + POOLW * 32;   // participants of the staging-buffer barriers
  constexpr int MAXC = NGRP == 2 ? 3 : 2;      // pooled chunks owned by one warp
  const Params& p = pp.base;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * POOL_A_STAGES + 5];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_scale[256];
  __shared__ __align__(16) float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_bytes_al = (uint32_t)((p.w_bytes + 127) & ~127);
  const uint32_t smem_w = smem_base;
  const uint32_t smem_a = smem_base + w_bytes_al;
  uint8_t* sR = gen_base + w_bytes_al + (size_t)POOL_A_STAGES * p.a_stage_bytes;  // [128][R_PITCH]
  const uint32_t bar_afull = smem_u32(&bars[0]);
  const uint32_t bar_aempty = smem_u32(&bars[POOL_A_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * POOL_A_STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * POOL_A_STAGES + 2]);
  const uint32_t bar_w = smem_u32(&bars[2 * POOL_A_STAGES + 4]);

  if (!FOLDED) {
    for (int i = threadIdx.x; i < p.n_total; i += 64 + SYNC_THREADS) {
      s_scale[i] = __ldg(p.scale + i);
      s_bias[i] = __ldg(p.bias + i);
    }
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < POOL_A_STAGES; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, NGRP * 4 + (POOLW ? 4 : 0));
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const uint32_t tile_bytes = (uint32_t)(p.KH * p.row_bytes);
  const Seq seq = make_seq(p.B * pp.Hp, pp.Hp);
  griddep_launch_dependents();          // PDL: the next kernel of the chain may be scheduled (see tc_common.cuh)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, (uint32_t)p.w_bytes);
      for (int o = 0; o < p.w_bytes; o += 32768)      // constant filters: loaded while the staging kernel may still run
        bulk_g2s(smem_w + (uint32_t)o, p.w + o, (uint32_t)std::min(32768, p.w_bytes - o), bar_w);
      griddep_wait();                   // the staged frame is the previous kernel's output
      int stage = 0;
      uint32_t phase = 0;
      for (int k = 0; k < seq.n_tiles; ++k) {
        int b, py, oh, role;
        tile_at(seq, k, pp.Hp, b, py, oh, role);
        mbar_wait(bar_aempty + 8 * stage, phase ^ 1u);
        mbar_arrive_expect_tx(bar_afull + 8 * stage, tile_bytes);
        bulk_g2s(smem_a + (uint32_t)stage * p.a_stage_bytes, p.x + ((int64_t)b * p.Hpad + 2 * oh) * p.row_bytes,
                 tile_bytes, bar_afull + 8 * stage);
        if (++stage == POOL_A_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(p.n_total);
    const uint32_t lbo_b = (uint32_t)p.n_total * 16u;
    mbar_wait(bar_w, 0);
    tcgen05_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    for (int k = 0; k < seq.n_tiles; ++k) {
      const int acc = k & 1;
      const uint32_t tphase = (uint32_t)(k >> 1) & 1u;
      mbar_wait(bar_tempty + 8 * acc, tphase ^ 1u);
      mbar_wait(bar_afull + 8 * stage, phase);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
      const uint32_t a_base = smem_a + (uint32_t)stage * p.a_stage_bytes;
      for (int kh = 0; kh < p.KH; ++kh) {
#pragma unroll
        for (int ks = 0; ks < WIN_K / UMMA_K; ++ks) {
          const uint64_t a_desc = make_nosw_desc(a_base + (uint32_t)(kh * p.row_bytes + ks * 32), 16u, 128u);
          const uint64_t b_desc = make_nosw_desc(smem_w + (uint32_t)(kh * 4 + ks * 2) * lbo_b, lbo_b, 128u);
          umma_bf16(d_tmem, a_desc, b_desc, idesc, (uint32_t)((kh | ks) != 0));
        }
      }
      umma_commit(bar_aempty + 8 * stage);
      umma_commit(bar_tfull + 8 * acc);
      if (++stage == POOL_A_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= 2 + NGRP * 4) {
    // ============================ pooling warps (POOLW > 0) ======================
    griddep_wait();
    // They also drain and store the channels that are not pooled (policy conv1) of every conv row - warp w may read TMEM
    // lane quarter w % 4 - so that every draining warp owns the same number of chunks.
    const int te = threadIdx.x - (64 + EPI_THREADS);
    const int lq = warp & 3, ow = lq * 32 + lane;
    asm volatile("bar.arrive 2, %0;" ::"n"(SYNC_THREADS) : "memory");          // staging buffer starts empty
    for (int k = 0; k < seq.n_tiles; ++k) {
      const int acc = k & 1;
      const uint32_t tphase = (uint32_t)(k >> 1) & 1u;
      int b, py, oh, role;
      tile_at(seq, k, pp.Hp, b, py, oh, role);
      if (warp < 2 + NGRP * 4 + 4) {   // the first four pooling warps (one per TMEM lane quarter)
        mbar_wait(bar_tfull + 8 * acc, tphase);
        tcgen05_fence_after();
        store_rest_chunks<FOLDED>(p, tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * ACC_STRIDE), pp.n_pool_ch >> 5, p.n_total >> 5,
                                  b, oh, ow, ow < p.Wo && role != ROLE_CARRY && !(pp.dbg & 2), s_scale, s_bias);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      }
      if (role != ROLE_ODD) continue;
      asm volatile("bar.sync 1, %0;" ::"n"(SYNC_THREADS) : "memory");          // vertical max of pooled row py staged
      if (pp.dbg & 128) pool_row_out<false>(pp, sR, b, py, te, POOLW * 32, pp.dbg);
      else pool_row_out<true>(pp, sR, b, py, te, POOLW * 32, pp.dbg);
      if (k != seq.n_tiles - 1) asm volatile("bar.arrive 2, %0;" ::"n"(SYNC_THREADS) : "memory");   // staging buffer free again
      if (pp.out_pad && !(pp.dbg & 8)) pool_row_border(pp, b, py, te, POOLW * 32);
    }
  } else {
    // ============================ epilogue + max-pool ============================
    griddep_wait();                             // output buffers may still be read by the previous kernels of the stream
    const int lq = warp & 3;                    // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 2;            // chunk group
    const int ow = lq * 32 + lane;
    const bool valid = ow < p.Wo;
    const int te = threadIdx.x - 64;            // index among the epilogue threads
    const int n_chunks = p.n_total >> 5;
    const int pool_chunks = pp.n_pool_ch >> 5;  // <= 6
    int c_begin, c_cnt, rest_grp;
    if (NGRP == 2) {
      const int h0 = (pool_chunks + 1) >> 1;
      c_begin = grp ? h0 : 0;
      c_cnt = grp ? pool_chunks - h0 : h0;      // <= 3 pooled chunks owned by this warp
      rest_grp = 1;
    } else {
      c_begin = (grp * pool_chunks) / NGRP;     // 6 chunks over 4 groups: 1, 2, 1, 2
      c_cnt = ((grp + 1) * pool_chunks) / NGRP - c_begin;
      rest_grp = 0;                             // the group with the fewest pooled chunks also stores the full-resolution ones
    }
    uint4 V[MAXC][4];  // running vertical max of the owned chunks (registers)
#pragma unroll
    for (int ci = 0; ci < MAXC; ++ci)
#pragma unroll
      for (int v = 0; v < 4; ++v) V[ci][v] = make_uint4(0u, 0u, 0u, 0u);

    for (int k = 0; k < seq.n_tiles; ++k) {
      const int acc = k & 1;
      const uint32_t tphase = (uint32_t)(k >> 1) & 1u;
      int b, py, oh, role;
      tile_at(seq, k, pp.Hp, b, py, oh, role);
      const bool fresh = role == ROLE_CARRY || (role == ROLE_EVEN && py == 0);  // nothing above: top padding / range start
      mbar_wait(bar_tfull + 8 * acc, tphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
      if (role == ROLE_ODD && !(pp.dbg & 16)) {   // previous horizontal pass has left sR
        if (POOLW) asm volatile("bar.sync 2, %0;" ::"n"(SYNC_THREADS) : "memory");
        else asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      }
      uint32_t a[32];
      if (FOLDED) {
        // 16-column TMEM loads, one always in flight behind the half-chunk being packed (same 32 registers as
        // one 32-column load; a second 32-column buffer would spill)
        uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&a[0]);
        uint32_t (&hi)[16] = *reinterpret_cast<uint32_t (*)[16]>(&a[16]);
        if (c_cnt > 0) tmem_ld_32x32b_x16(taddr + (uint32_t)(c_begin * 32), lo);
#pragma unroll
        for (int hc = 0; hc < 2 * MAXC; ++hc) {
          if (hc < 2 * c_cnt) {
            const int ci = hc >> 1, c = c_begin + ci, half = hc & 1;
            tmem_ld_wait();
            if (hc + 1 < 2 * c_cnt) tmem_ld_32x32b_x16(taddr + (uint32_t)(c_begin * 32 + (hc + 1) * 16), half ? lo : hi);
            const uint32_t (&cur)[16] = half ? hi : lo;
#pragma unroll
            for (int vv = 0; vv < 2; ++vv) {
              const int v = half * 2 + vv;
              const uint4 q = make_uint4(pack_bf16x2(__uint_as_float(cur[vv * 8 + 0]), __uint_as_float(cur[vv * 8 + 1])),
                                         pack_bf16x2(__uint_as_float(cur[vv * 8 + 2]), __uint_as_float(cur[vv * 8 + 3])),
                                         pack_bf16x2(__uint_as_float(cur[vv * 8 + 4]), __uint_as_float(cur[vv * 8 + 5])),
                                         pack_bf16x2(__uint_as_float(cur[vv * 8 + 6]), __uint_as_float(cur[vv * 8 + 7])));
              if (fresh) {
                V[ci][v] = q;
              } else if (role == ROLE_EVEN) {
                V[ci][v] = hmax8(V[ci][v], q);
              } else {
                if (!(pp.dbg & 4)) *reinterpret_cast<uint4*>(sR + ow * R_PITCH + c * 64 + v * 16) = hmax8(V[ci][v], q);
                V[ci][v] = q;   // conv row 2py+1 is row 2(py+1)-1 of the next pooled row
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < MAXC; ++ci) {
          if (ci < c_cnt) {
            const int c = c_begin + ci;
            tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), a);
            tmem_ld_wait();
            uint4 q[4];
            bn_pack32<false>(a, s_scale + c * 32, s_bias + c * 32, q);   // ReLU commutes with max: applied once after pooling
            if (fresh) {
#pragma unroll
              for (int v = 0; v < 4; ++v) V[ci][v] = q[v];
            } else if (role == ROLE_EVEN) {
#pragma unroll
              for (int v = 0; v < 4; ++v) V[ci][v] = hmax8(V[ci][v], q[v]);
            } else {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                *reinterpret_cast<uint4*>(sR + ow * R_PITCH + c * 64 + v * 16) = hmax8(V[ci][v], q[v]);
                V[ci][v] = q[v];   // conv row 2py+1 is row 2(py+1)-1 of the next pooled row
              }
            }
          }
        }
      }
      if (POOLW == 0 && grp == rest_grp) {   // not pooled      } else if (POOLW == 0 && grp == rest_grp) {   // not pooled (policy conv1); a carry row belongs to another CTA's range
        store_rest_chunks<FOLDED>(p, taddr, pool_chunks, n_chunks, b, oh, ow, valid && role != ROLE_CARRY && !(pp.dbg & 2), s_scale, s_bias);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);   // accumulator drained: the MMA warp may reuse it
      if (role != ROLE_ODD) continue;
      if (POOLW) {
        asm volatile("bar.arrive 1, %0;" ::"n"(SYNC_THREADS) : "memory");   // vertical max of pooled row py staged: over to the pooling warps
        continue;
      }
      if (!(pp.dbg & 16)) asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");       // vertical max of pooled row py staged
      // The index math is redone per pooled row on purpose: hoisting it out of the tile loop made ptxas spill it to local
      // memory, and those reloads (L1 misses under the streaming stores) cost more than the divisions.
      pool_row_out<false>(pp, sR, b, py, te, EPI_THREADS, pp.dbg);
      if (pp.out_pad && !(pp.dbg & 8)) pool_row_border(pp, b, py, te, EPI_THREADS);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace stem

int amoe_stem_init(amoe_ctx* ctx) {
  AMOE_ENTER(ctx);
  (void)ctx;
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(stem::stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(stem::stem_tc_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(stem::stem_pool_kernel<false, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(stem::stem_pool_kernel<true, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(stem::stem_pool_kernel<true, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  return 0;
}

static int stem_fill(stem::Params& p, const void* x_pad, const void* w_img, const float* scale, const float* bias,
                     int B, int H, int W, int Wpad, int KH, int n_total, int relu, void* const* dst_host,
                     const int* dst_c_host, int first_dst_chunk);

extern "C" int amoe_stem_pool_fwd(amoe_ctx* ctx, const void* x_pad, const void* w_img, const float* scale,
                                  const float* bias, int B, int H, int W, int Wpad, int KH, int n_total, int relu,
                                  int n_pool_ch, void* pooled, int out_pad, void* const* dst_host,
                                  const int* dst_c_host, void* stream) {
  AMOE_ENTER(ctx);
  using namespace stem;
  AMOE_REQUIRE(ctx && pooled, "amoe_stem_pool_fwd: NULL argument");
  AMOE_REQUIRE(n_pool_ch % 64 == 0 && n_pool_ch >= 64 && n_pool_ch <= 192 && n_pool_ch <= n_total,
               "amoe_stem_pool_fwd: n_pool_ch=%d must be a multiple of 64 in [64,192]", n_pool_ch);
  AMOE_REQUIRE(H % 4 == 0 && W % 4 == 0, "amoe_stem_pool_fwd: H and W must be multiples of 4 (got %dx%d)", H, W);
  AMOE_REQUIRE(out_pad == 0 || out_pad == 1, "amoe_stem_pool_fwd: out_pad must be 0 or 1");
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(pooled) & 15) == 0, "amoe_stem_pool_fwd: pooled must be 16-byte aligned");
  PoolParams pp;
  int rc = stem_fill(pp.base, x_pad, w_img, scale, bias, B, H, W, Wpad, KH, n_total, relu, dst_host, dst_c_host, n_pool_ch / 32);
  if (rc) return rc;
  pp.n_pool_ch = n_pool_ch;
  pp.Hp = H / 4; pp.Wp = W / 4; pp.out_pad = out_pad;
  pp.pooled = (__nv_bfloat16*)pooled;
  { const char* e = getenv("AMOE_STEM_DBG"); pp.dbg = e ? atoi(e) : 0; }
  const int p_total = B * pp.Hp;
  if (p_total == 0) return 0;
  const size_t smem = ((size_t)pp.base.w_bytes + 127) / 128 * 128 + (size_t)POOL_A_STAGES * pp.base.a_stage_bytes +
                      (size_t)128 * R_PITCH + 256;
  AMOE_REQUIRE(smem <= 224 * 1024, "amoe_stem_pool_fwd: shared memory budget exceeded (%zu bytes)", smem);
  const int grid = std::min(p_total, ctx->sm_count);
  // AMOE_STEM_POOLW=0: the round-1 epilogue (draining warps also run the horizontal pass)
  const char* pw = getenv("AMOE_STEM_POOLW");
  const int poolw = pw ? atoi(pw) : 4;
  if (scale == nullptr && poolw == 4)
    AMOE_CHECK_CUDA(amoe_launch_pdl(stem_pool_kernel<true, 2, 4>, dim3(grid), dim3(POOL_THREADS + 128), smem, (cudaStream_t)stream, pp));
  else if (scale == nullptr)
    AMOE_CHECK_CUDA(amoe_launch_pdl(stem_pool_kernel<true, 2, 0>, dim3(grid), dim3(POOL_THREADS), smem, (cudaStream_t)stream, pp));
  else
    AMOE_CHECK_CUDA(amoe_launch_pdl(stem_pool_kernel<false, 2, 0>, dim3(grid), dim3(POOL_THREADS), smem, (cudaStream_t)stream, pp));
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int amoe_stem_fwd(amoe_ctx* ctx, const void* x_pad, const void* w_img, const float* scale,
                             const float* bias, int B, int H, int W, int Wpad, int KH, int n_total, int relu,
                             void* const* dst_host, const int* dst_c_host, void* stream) {
  AMOE_ENTER(ctx);
  using namespace stem;
  AMOE_REQUIRE(ctx != nullptr, "amoe_stem_fwd: NULL ctx");
  AMOE_REQUIRE(scale != nullptr, "amoe_stem_fwd: folded filters (scale == NULL) are only taken by amoe_stem_pool_fwd");
  Params p;
  int rc = stem_fill(p, x_pad, w_img, scale, bias, B, H, W, Wpad, KH, n_total, relu, dst_host, dst_c_host, 0);
  if (rc) return rc;
  if (p.total_tiles == 0) return 0;
  const size_t smem = ((size_t)p.w_bytes + 127) / 128 * 128 + (size_t)A_STAGES * p.a_stage_bytes + 256;
  AMOE_REQUIRE(smem <= 220 * 1024, "amoe_stem_fwd: shared memory budget exceeded (%zu bytes)", smem);
  const int grid = std::min(p.total_tiles, ctx->sm_count);
  stem_tc_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(p);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

// [B][H][W][4] fp32 NHWC (4th channel zero) -> the padded frame of the stem kernels, [3][B][H+6][Wpad][4] bf16: 3 zero rows above
// and below, 4 zero pixels left, zeros right, as the three bf16 parts of every value.  One thread per padded pixel.
__global__ void __launch_bounds__(256) stem_split_frame_kernel(const float4* __restrict__ x, uint2* __restrict__ out, int B, int H, int W,
                                                               int Wpad, int64_t total) {
  const int Hpad = H + 6;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xp = (int)(i % Wpad);
    const int64_t r = i / Wpad;
    const int yp = (int)(r % Hpad);
    const int b = (int)(r / Hpad);
    const int xx = xp - 4, yy = yp - 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xx >= 0 && xx < W && yy >= 0 && yy < H) v = __ldg(x + ((int64_t)b * H + yy) * W + xx);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 p0[4], p1[4], p2[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) tc::split3(f[c], p0[c], p1[c], p2[c]);
    out[i] = *reinterpret_cast<const uint2*>(p0);
    out[i + total] = *reinterpret_cast<const uint2*>(p1);
    out[i + 2 * total] = *reinterpret_cast<const uint2*>(p2);
  }
}

extern "C" int amoe_stem_split_frame(amoe_ctx* ctx, const float* x_nhwc4, void* out3, int B, int H, int W, int Wpad, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x_nhwc4 && out3, "amoe_stem_split_frame: NULL argument");
  AMOE_REQUIRE(B >= 0 && H > 0 && W > 0 && Wpad >= W + 6, "amoe_stem_split_frame: bad shape");
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(x_nhwc4) & 15) == 0 && (reinterpret_cast<uintptr_t>(out3) & 7) == 0,
               "amoe_stem_split_frame: pointers must be 16-byte (input) / 8-byte (output) aligned");
  const int64_t total = (int64_t)B * (H + 6) * Wpad;
  if (total == 0) return 0;
  const int64_t want = (total + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
  stem_split_frame_kernel<<<(unsigned)std::min(want, cap), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x_nhwc4), reinterpret_cast<uint2*>(out3), B, H, W, Wpad, total);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int amoe_stem_fwd_f32tc_supported(int H, int W, int KH, int n_total) {
  if (H <= 0 || W <= 0 || (H & 1) || (W & 1) || W / 2 > 128 || KH < 1 || KH > 7) return 0;
  if (n_total % 32 != 0 || n_total < 32 || n_total > 256) return 0;
  const int Wpad = (W + 6 + 7) & ~7;
  const size_t a_stage = (size_t)((KH * Wpad * 8 + 16 * 128 + 64 + 127) & ~127);
  const size_t w_part = (size_t)((KH * 4 * n_total * 16 + 127) & ~127);
  return (3 * w_part + 2 * 3 * a_stage + 256 <= 224u * 1024u) ? 1 : 0;   // weights + two activation stages
}

extern "C" int amoe_stem_fwd_f32tc(amoe_ctx* ctx, const void* x3_pad, const void* w3_img, const float* scale, const float* bias,
                                   float* y, int B, int H, int W, int Wpad, int KH, int n_total, int relu, void* stream) {
  AMOE_ENTER(ctx);
  using namespace stem;
  AMOE_REQUIRE(ctx && x3_pad && w3_img && scale && bias && y, "amoe_stem_fwd_f32tc: NULL argument");
  AMOE_REQUIRE(amoe_stem_fwd_f32tc_supported(H, W, KH, n_total), "amoe_stem_fwd_f32tc: unsupported shape H=%d W=%d KH=%d N=%d", H, W, KH, n_total);
  AMOE_REQUIRE(Wpad == ((W + 6 + 7) & ~7), "amoe_stem_fwd_f32tc: Wpad=%d, expected %d", Wpad, (W + 6 + 7) & ~7);
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(y) & 31) == 0, "amoe_stem_fwd_f32tc: y must be 32-byte aligned");
  F32Params fp;
  void* dummy_dst[MAX_CHUNKS];
  int dummy_c[MAX_CHUNKS];
  for (int c = 0; c < MAX_CHUNKS; ++c) { dummy_dst[c] = y; dummy_c[c] = 8; }
  int rc = stem_fill(fp.base, x3_pad, w3_img, scale, bias, B, H, W, Wpad, KH, n_total, relu, dummy_dst, dummy_c, 0);
  if (rc) return rc;
  if (fp.base.total_tiles == 0) return 0;
  fp.y = y;
  fp.part_bytes = (int64_t)B * fp.base.Hpad * fp.base.row_bytes;
  const size_t w_part = ((size_t)fp.base.w_bytes + 127) / 128 * 128;
  const size_t stage = 3 * (size_t)fp.base.a_stage_bytes;
  fp.stages = (int)std::min<size_t>(A_STAGES, (224 * 1024 - 256 - 3 * w_part) / stage);
  AMOE_REQUIRE(fp.stages >= 2, "amoe_stem_fwd_f32tc: shared memory budget exceeded");
  const size_t smem = 3 * w_part + (size_t)fp.stages * stage + 256;
  const int grid = std::min(fp.base.total_tiles, ctx->sm_count);
  stem_tc_f32_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(fp);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

static int stem_fill(stem::Params& p, const void* x_pad, const void* w_img, const float* scale, const float* bias,
                     int B, int H, int W, int Wpad, int KH, int n_total, int relu, void* const* dst_host,
                     const int* dst_c_host, int first_dst_chunk) {
  using namespace stem;
  AMOE_REQUIRE(x_pad && w_img && dst_host && dst_c_host, "amoe_stem_fwd: NULL argument");
  AMOE_REQUIRE((scale == nullptr) == (bias == nullptr), "amoe_stem_fwd: scale and bias come together");
  AMOE_REQUIRE(H % 2 == 0 && W % 2 == 0, "amoe_stem_fwd: H and W must be even (got %dx%d)", H, W);
  AMOE_REQUIRE(n_total % 32 == 0 && n_total >= 32 && n_total <= 256, "amoe_stem_fwd: n_total=%d must be a multiple of 32 in [32,256]", n_total);
  AMOE_REQUIRE(KH >= 1 && KH <= 7, "amoe_stem_fwd: KH=%d out of range", KH);
  const int Ho = H / 2, Wo = W / 2;
  AMOE_REQUIRE(Wo <= 128, "amoe_stem_fwd: output rows wider than 128 pixels are not tiled yet (Wo=%d)", Wo);
  AMOE_REQUIRE(Wpad >= W + 6 && Wpad % 2 == 0, "amoe_stem_fwd: Wpad=%d too small/odd for W=%d", Wpad, W);
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(x_pad) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_img) & 15) == 0,
               "amoe_stem_fwd: pointers must be 16-byte aligned");
  p.B = B; p.Ho = Ho; p.Wo = Wo; p.Hpad = H + 6; p.KH = KH;
  p.row_bytes = Wpad * 8;
  // garbage rows (ow >= Wo) read up to byte 16*127 + 64 past the start of the last filter row
  p.a_stage_bytes = (KH * p.row_bytes + 16 * 128 + 64 + 127) & ~127;
  p.n_total = n_total;
  p.w_bytes = KH * 4 * n_total * 16;
  p.total_tiles = B * Ho;
  p.relu = relu;
  p.x = (const uint8_t*)x_pad;
  p.w = (const uint8_t*)w_img;
  p.scale = scale; p.bias = bias;
  for (int c = 0; c < MAX_CHUNKS; ++c) { p.dst[c] = nullptr; p.dst_c[c] = 0; }
  for (int c = first_dst_chunk; c < n_total / 32; ++c) {
    AMOE_REQUIRE(dst_host[c] != nullptr && dst_c_host[c] % 8 == 0 && (reinterpret_cast<uintptr_t>(dst_host[c]) & 15) == 0,
                 "amoe_stem_fwd: bad destination for chunk %d", c);
    p.dst[c] = (__nv_bfloat16*)dst_host[c];
    p.dst_c[c] = dst_c_host[c];
  }
  return 0;
}
