// Implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a):
//   tcgen05.mma (bf16 x bf16 -> fp32, accumulators in TMEM), operands staged in shared
//   memory by TMA (cp.async.bulk.tensor, 128B swizzle), mbarrier producer/consumer
//   pipeline, persistent CTAs, folded-BN scale/bias + residual + ReLU fused in the epilogue.
//
// GEMM view:  D[M, N] = sum over taps (kh,kw) and 64-channel chunks of  A_tap[M, 64] * W_tap[N, 64]^T
//   M tile = 128 output pixels = a (nb x th x tw) patch (images x rows x cols), so the A
//            operand of one tap is ONE TMA box of the NHWC input shifted by the tap offset;
//            TMA zero-fills out-of-bounds coordinates = the convolution's zero padding.
//   N tile = BLOCK_N output channels (64/128/256) = one TMA box of the packed weights
//            [G*Cout][KH*KW*Cin] (K contiguous).
//   Stride-2 convolutions read a "parity view" of the same memory: [N, H/2, 2, W/2, 2*C]
//   so that every tap is again a dense box (no element strides needed).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (one elected lane), warps 2..5 = epilogue (TMEM -> registers -> global).  Two TMEM
// accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace tc {

constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int MAX_STAGES = 8;
constexpr int MAX_TAPS = 64;
constexpr int NUM_THREADS = 192;       // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int NUM_THREADS_EPI8 = 320;  // ... plus warps 6-9: a second set of epilogue warps taking the upper half of a tile's columns
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // TMEM columns between the two accumulator buffers
constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int SMEM_RESIDENT = 212 * 1024;   // operand bytes of the resident-weight mode (scale/bias staging comes on top)

struct Tap {
  int c_off, dw, hp, dh;  // TMA start-coordinate offsets of this filter tap
  int w_k;                // first K column of this tap's weights in the packed [rows][K] weight matrix
};

struct Params {
  int tw, th, nb;                   // M-tile patch: cols, rows, images (tw*th*nb == 128)
  int tiles_w, tiles_h, tiles_b;    // patches per group
  int n_tiles_n;                    // Cout / block_n
  int G, B, Ho, Wo, Cout;
  int block_n;
  int num_taps, k_chunks;           // K iterations per tile = num_taps * k_chunks
  int stages;
  int relu, x_shared;
  int out_pad;                      // output (and residual) tensors carry a physical border of out_pad pixels
  int out_f32;                      // output tensor is fp32 (split-operand training convolutions), no residual
  int o_hm, o_ha, o_wm, o_wa;       // output pixel (oh, ow) lands at row oh*o_hm + o_ha, column ow*o_wm + o_wa of an
  int o_H, o_W;                     //   [N][o_H][o_W][C] tensor (dense: 1, out_pad, 1, out_pad, Ho+2*out_pad, Wo+2*out_pad)
  int split_c;                      // output sub-tensor width: channel ch of group g goes to tensor
                                    // (g*Cout/split_c + ch/split_c), channel ch%split_c (== Cout normally)
  int total_tiles;
  // Tail splitting (nprob == 1): the persistent grid walks `total_units` work units.  The first `full_units` (a multiple of
  // the grid size) are whole tiles; each of the remaining tiles - the partial last wave, where most SMs would idle for a
  // whole tile time (layer4: 768 tiles on 148 SMs = 5.19 waves) - is cut into `split` units of sub_n = block_n / split
  // channels (their weight boxes come through tmW2), so the last wave costs 1/split of a tile time.
  int total_units, full_units, split, sub_n;
  // Resident weights (short-K stage entries: Cin = 64, one N tile): every tap's weight box of the CTA's expert group - and
  // the second problem's - is loaded ONCE into shared memory, the ring then carries activation boxes only (the launch was
  // bound by L2->SMEM traffic: 16 KB of weights per 16 KB activation box).  The grid is (CTAs per group, G): a CTA stays
  // inside one expert group, tiles_per_group tiles each.
  int w_resident, tiles_per_group;
  // (Measured and removed: cp.async.bulk.prefetch.tensor of the next unit's activation boxes into L2 when a unit starts -
  // every launch got slower, layer2 entry 150 -> 193 us, layer4 147 -> 178 us: the prefetches queue in front of the loads.)
  // Neither did plain prefetch.global.L2 of the unit after the next one's input region by the idle lanes of the producer warp
  // (layer2 entry 159 -> 259 us).  The stage entries are HBM-bound, not latency-bound: the layer2 entry reads 403 MB and
  // writes 168 MB (two outputs in the dual launch: 374 MB) of DRAM, 4.8 TB/s when its epilogue only drains.
  // (Measured and removed, twice - rounds 1 and 2: a write-out staged through swizzled shared memory, 4 complete 128-byte row
  // pieces per store instruction instead of 32 16-byte pieces.  Bit-identical, slower everywhere it was on: layer2 entry
  // 296 -> 415 us, layer3 entry 156 -> 198 us, policy conv3 57 -> 72 us - N <= 128 MMA steps already read their operands at
  // the shared-memory bandwidth limit, the staging traffic comes out of the same budget.)
  // (Measured and removed, round 2 session 3: only the FOUR plane-opening boxes of the next unit prefetched a unit ahead
  // (stride-2 resident-weight launches, 3-box ring): layer2 entry dual 224 -> 247 us, single 157 -> 199 us, policy conv3
  // 56 -> 68 us, layer3/4 entries unchanged.  Any cp.async.bulk.prefetch.tensor traffic slows these launches down.)
  // (Measured and removed, same session: four accumulator buffers of 128 columns for N <= 128 tiles, so that the short 1x1 tile of
  // a dual launch never waits for the 3x3 tile's epilogue - layer2 entry 207.0 -> 206.1 us, everything else unchanged: the
  // stage entries are bound by the activation boxes coming through TMA from L2/HBM (~5.5 TB/s of A operand traffic, the same
  // rate the layer3 convolutions see), not by accumulator hand-over.)
  int reverse;   // walk the whole tiles back to front (amoe_set_walk_reverse); tail-split units stay last
  int wide32;   // lean epilogue: 32-byte stores / residual loads (y and residual 32-byte aligned; AMOE_TC_W32=0 -> 16-byte)
  int dbg;   // AMOE_TC_DBG experiment bits (results wrong on purpose): 1 = the epilogue only drains the accumulator
  int n_ch_total;                   // G*Cout: scale/bias entries staged in shared memory
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  // Optional SECOND convolution over the same input, output geometry and Cout, run as a second tile behind every tile of
  // the first one (nprob == 2): the 1x1/stride-2 downsample of a ResNet stage entry next to its 3x3/stride-2 conv1.
  // Its single tap is the centre tap of conv1, so the activation box it loads was loaded a moment ago (L2 hit), and the
  // short-K tile rides in the same persistent pipeline instead of paying a launch, a prologue and an L2-cold sweep.
  int nprob;
  int relu2;
  Tap tap2;
  const float* scale2;
  const float* bias2;
  __nv_bfloat16* y2;
  Tap taps[MAX_TAPS];
};

struct TileCoord {
  int g, bt, ht, wt, nt;
};
// CTAS = 2 (CTA pairs): `t` numbers PAIRS of M tiles - p.tiles_b counts pairs of image tiles, the CTA of cluster rank r
// takes image tile 2*bt + r (an odd image-tile count leaves the last tile of rank 1 out of bounds: zeros in, nothing out)
template <int CTAS>
__device__ __forceinline__ TileCoord decode_tile(const Params& p, int t, int rank) {
  TileCoord c;
  c.nt = t % p.n_tiles_n;
  t /= p.n_tiles_n;
  c.wt = t % p.tiles_w;
  t /= p.tiles_w;
  c.ht = t % p.tiles_h;
  t /= p.tiles_h;
  c.bt = (t % p.tiles_b) * CTAS + rank;
  c.g = t / p.tiles_b;
  return c;
}

struct Unit {
  int tile, n_off, bn;
};
__device__ __forceinline__ Unit decode_unit(const Params& p, int u) {
  Unit r;
  if (u < p.full_units) {
    r.n_off = 0; r.bn = p.block_n;
    if (!p.reverse) r.tile = u;
    else if (p.w_resident) { const int g = u / p.tiles_per_group; r.tile = (2 * g + 1) * p.tiles_per_group - 1 - u; }   // inside its group
    else r.tile = p.full_units - 1 - u;
  } else {
    const int k = u - p.full_units;
    r.tile = p.full_units + k / p.split;
    r.n_off = (k % p.split) * p.sub_n;
    r.bn = p.sub_n;
  }
  return r;
}

// ---- lean bf16 epilogue (LEAN = true): the common inference case - bf16 output, one output tensor per group (split_c == Cout) ----
// The generic epilogue below spends ~250 instructions per 32-column chunk (an integer division for the sub-tensor split,
// the fp32 / bf16 output and residual variants, scalar FFMA / FMNMX); short-K launches (stage entries, 1x1 convolutions, the
// policy backbone) are bound by it.  Here a chunk is: 16 LDS.128 (scale / bias), 16 two-lane FMAs (fma.rn.f32x2, per-element
// IEEE: the bits of fmaf), 16 packs, the ReLU on the packed pairs (max(bf16(x), 0) == bf16(max(x, 0))), 4 16-byte stores.
// one 32-column chunk: acc -> scale/bias (+ residual) -> bf16 (-> ReLU) -> two 32-byte stores at dst (32-byte aligned:
// channel offsets are multiples of 32, rows of Cout % 32 == 0 channels, tensors from 32-byte aligned allocations)
template <bool RES>
__device__ __forceinline__ void lean_chunk(const uint32_t (&acc)[32], const float* sc, const float* bs, const __nv_bfloat16* res,
                                           __nv_bfloat16* dst, bool relu, bool wide) {
  const float4* sc4 = reinterpret_cast<const float4*>(sc);
  const float4* bs4 = reinterpret_cast<const float4*>(bs);
  uint4 rr[4];
  if (RES) {
    if (wide) {
      ldg256_nc(res, rr[0], rr[1]);
      ldg256_nc(res + 16, rr[2], rr[3]);
    } else {
#pragma unroll
      for (int v = 0; v < 4; ++v) rr[v] = __ldg(reinterpret_cast<const uint4*>(res + v * 8));
    }
  }
  uint4 o[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 s0 = sc4[2 * v], s1 = sc4[2 * v + 1], b0 = bs4[2 * v], b1 = bs4[2 * v + 1];
    float f[8];
    ffma2(f[0], f[1], acc[v * 8 + 0], acc[v * 8 + 1], s0.x, s0.y, b0.x, b0.y);
    ffma2(f[2], f[3], acc[v * 8 + 2], acc[v * 8 + 3], s0.z, s0.w, b0.z, b0.w);
    ffma2(f[4], f[5], acc[v * 8 + 4], acc[v * 8 + 5], s1.x, s1.y, b1.x, b1.y);
    ffma2(f[6], f[7], acc[v * 8 + 6], acc[v * 8 + 7], s1.z, s1.w, b1.z, b1.w);
    if (RES) {
      add_bf16x2(f[0], f[1], rr[v].x);
      add_bf16x2(f[2], f[3], rr[v].y);
      add_bf16x2(f[4], f[5], rr[v].z);
      add_bf16x2(f[6], f[7], rr[v].w);
    }
    o[v] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    if (relu) o[v] = make_uint4(relu_bf16x2(o[v].x), relu_bf16x2(o[v].y), relu_bf16x2(o[v].z), relu_bf16x2(o[v].w));
  }
  if (wide) {
    stg256(dst, o[0], o[1]);
    stg256(dst + 16, o[2], o[3]);
  } else {
#pragma unroll
    for (int v = 0; v < 4; ++v) *reinterpret_cast<uint4*>(dst + v * 8) = o[v];
  }
}

// CTAS = 2: clusters of two CTAs work on two M tiles (adjacent image tiles) of the same N tile; the leader issues M = 256
// tcgen05.mma.cta_group::2 whose B operand is split over the pair - each CTA loads and holds only HALF of every weight box
// (16 instead of 32 KB per K chunk at N = 256: a third less L2->SMEM traffic per CTA, half the B reads per MMA step).
template <int CTAS, bool LEAN>
__global__ void __launch_bounds__(NUM_THREADS_EPI8, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 5];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B needs 1024 B alignment
  const uint32_t b_stage_bytes = (uint32_t)(p.block_n / CTAS) * 128u;   // this CTA's rows of a weight box
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int u_first = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // work units of this CTA (pair)
  const int u_step = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // work units of this CTA: u = u_lo + u_first, + u_step, ... < u_hi
  const int u_lo = p.w_resident ? (int)blockIdx.y * p.tiles_per_group : 0;
  const int u_hi = p.w_resident ? u_lo + p.tiles_per_group : p.total_units;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + (uint32_t)p.stages * A_STAGE_BYTES;   // weight stages, or the resident weight boxes
  const uint32_t w_boxes = (uint32_t)(p.num_taps * p.k_chunks + (p.nprob == 2 ? p.k_chunks : 0));   // resident boxes
  // folded-BN scale/bias of every group, staged once per CTA (the epilogue reads them with LDS.128)
  float* s_scale = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                            (p.w_resident ? (size_t)p.stages * A_STAGE_BYTES + (size_t)w_boxes * b_stage_bytes
                                                          : (size_t)p.stages * (A_STAGE_BYTES + b_stage_bytes)));
  float* s_bias = s_scale + p.n_ch_total;
  for (int i = threadIdx.x; i < p.n_ch_total; i += (int)blockDim.x) {
    s_scale[i] = __ldg(p.scale + i);
    s_bias[i] = __ldg(p.bias + i);
  }
  float* s_scale2 = s_bias + p.n_ch_total;      // second problem (nprob == 2)
  float* s_bias2 = s_scale2 + p.n_ch_total;
  if (p.nprob == 2) {
    for (int i = threadIdx.x; i < p.n_ch_total; i += (int)blockDim.x) {
      s_scale2[i] = __ldg(p.scale2 + i);
      s_bias2[i] = __ldg(p.bias2 + i);
    }
  }
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[MAX_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * MAX_STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * MAX_STAGES + 2]);
  const uint32_t bar_w = smem_u32(&bars[2 * MAX_STAGES + 4]);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    if (p.nprob == 2 || p.split > 1) prefetch_tmap(&tmW2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, (blockDim.x - 64u) / 32u * CTAS);  // one arrive per epilogue warp (of both CTAs, on the leader's barrier)
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair(smem_u32(&tmem_holder), TMEM_COLS);
    else tmem_alloc(smem_u32(&tmem_holder), TMEM_COLS);
  }
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  // PDL: everything above touched only kernel parameters and constant weights (folded BN); the next kernel of the chain
  // may be scheduled now, and nothing below runs before the previous kernel's results are complete
  griddep_launch_dependents();
  if (p.w_resident && warp == 0 && lane == 0) {
    // constant weights: loaded while the previous kernel of the stream may still be running
    const int wrow = (int)blockIdx.y * p.Cout + (int)rank * (p.block_n / CTAS);
    const uint32_t wf = CTAS == 2 ? mapa_rank(bar_w, 0u) : bar_w;
    if (rank == 0) mbar_arrive_expect_tx(bar_w, (uint32_t)CTAS * w_boxes * b_stage_bytes);
    for (int j = 0; j < p.num_taps * p.k_chunks; ++j) {
      const int tap = j / p.k_chunks, kc = j - tap * p.k_chunks;
      if (CTAS == 2) tma_load_2d_pair(smem_b + (uint32_t)j * b_stage_bytes, &tmW, wf, p.taps[tap].w_k + kc * BLOCK_K, wrow);
      else tma_load_2d(smem_b + (uint32_t)j * b_stage_bytes, &tmW, wf, p.taps[tap].w_k + kc * BLOCK_K, wrow);
    }
    if (p.nprob == 2)
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        const uint32_t dst = smem_b + (uint32_t)(p.num_taps * p.k_chunks + kc) * b_stage_bytes;
        if (CTAS == 2) tma_load_2d_pair(dst, &tmW2, wf, kc * BLOCK_K, wrow);
        else tma_load_2d(dst, &tmW2, wf, kc * BLOCK_K, wrow);
      }
  }
  griddep_wait();

  const int k_iters1 = p.num_taps * p.k_chunks;

  if (warp == 0) {
    // ============================ TMA producer ============================
    int stage = 0;
    uint32_t phase = 0;
    auto full_bar = [&](uint32_t bar) { return CTAS == 2 ? mapa_rank(bar, 0u) : bar; };
    auto load_a = [&](uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
      if (CTAS == 2) tma_load_5d_pair(dst, &tmA, bar, c0, c1, c2, c3, c4);
      else tma_load_5d(dst, &tmA, bar, c0, c1, c2, c3, c4);
    };
    auto load_w = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
      if (CTAS == 2) tma_load_2d_pair(dst, m, bar, c0, c1);
      else tma_load_2d(dst, m, bar, c0, c1);
    };
    for (int u = u_lo + u_first; u < u_hi; u += u_step) {
      if (lane == 0) {
        const Unit un = decode_unit(p, u);
        const TileCoord tc_ = decode_tile<CTAS>(p, un.tile, (int)rank);
        // an image tile past the batch (rank 1 of the last pair) starts at an image index TMA treats as out of bounds: zeros
        const int n0 = (p.x_shared ? 0 : tc_.g * p.B) + tc_.bt * p.nb + ((CTAS == 2 && tc_.bt * p.nb >= p.B) ? (1 << 28) : 0);
        const int oh0 = tc_.ht * p.th, ow0 = tc_.wt * p.tw;
        const int wrow0 = tc_.g * p.Cout + tc_.nt * p.block_n + un.n_off + (int)rank * (un.bn / CTAS);
        const CUtensorMap* wmap = un.bn == p.block_n ? &tmW : &tmW2;      // sub-tile units: the narrow weight box
        const uint32_t unit_b_bytes = (uint32_t)(un.bn / CTAS) * 128u;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const Tap tp = p.taps[tap];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)CTAS * (A_STAGE_BYTES + (p.w_resident ? 0u : unit_b_bytes)));
            const uint32_t fb = full_bar(bar_full + 8 * stage);
            load_a(smem_a + stage * A_STAGE_BYTES, fb, tp.c_off + kc * BLOCK_K, ow0 + tp.dw, tp.hp, oh0 + tp.dh, n0);
            if (!p.w_resident) load_w(smem_b + stage * b_stage_bytes, wmap, fb, tp.w_k + kc * BLOCK_K, wrow0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        if (p.nprob == 2) {   // the second convolution's tile for the same output patch: one tap, k_chunks chunks
          const Tap tp = p.tap2;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)CTAS * (A_STAGE_BYTES + (p.w_resident ? 0u : b_stage_bytes)));
            const uint32_t fb = full_bar(bar_full + 8 * stage);
            load_a(smem_a + stage * A_STAGE_BYTES, fb, tp.c_off + kc * BLOCK_K, ow0 + tp.dw, tp.hp, oh0 + tp.dh, n0);
            if (!p.w_resident) load_w(smem_b + stage * b_stage_bytes, &tmW2, fb, kc * BLOCK_K, wrow0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (pairs: the leader CTA only) =
    if (rank == 0) {
    const uint32_t idesc_full = CTAS == 2 ? make_idesc_pair(p.block_n) : make_idesc(p.block_n);
    const uint32_t idesc_sub = CTAS == 2 ? make_idesc_pair(p.sub_n) : make_idesc(p.sub_n);
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t accum) {
      if (CTAS == 2) umma_bf16_pair(d, a, b, id, accum);
      else umma_bf16(d, a, b, id, accum);
    };
    auto commit = [&](uint32_t bar) {
      if (CTAS == 2) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (p.w_resident) {
      mbar_wait(bar_w, 0);
      tcgen05_fence_after();
    }
    for (int u = u_lo + u_first; u < u_hi; u += u_step)
    for (int prob = 0; prob < p.nprob; ++prob, ++it) {
      const uint32_t idesc = u < p.full_units ? idesc_full : idesc_sub;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      if (CTAS == 2) mbar_wait_cluster(bar_tempty + 8 * as, aphase ^ 1u);
      else mbar_wait(bar_tempty + 8 * as, aphase ^ 1u);  // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * ACC_STRIDE);
      const int k_iters = prob ? p.k_chunks : k_iters1;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(bar_full + 8 * stage, phase);  // TMA bytes have landed
        tcgen05_fence_after();
        {
          const uint64_t a_desc = make_sw128_desc(smem_a + stage * A_STAGE_BYTES);
          const uint64_t b_desc = make_sw128_desc(smem_b + (p.w_resident ? (uint32_t)(prob ? k_iters1 + k : k) : (uint32_t)stage) * b_stage_bytes);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            // advance 32 bytes (16 bf16) inside the 128B swizzle row: +2 in the >>4 address field
            mma(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, (uint32_t)((k | kk) != 0));
          }
          commit(bar_empty + 8 * stage);                     // frees the smem slot
          if (k == k_iters - 1) commit(bar_tfull + 8 * as);  // accumulator ready
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    }
  } else {
    // ============================ epilogue ================================
    // Short-K launches (stage entries, 1x1 convolutions) are bound by this epilogue, one warp per SM sub-partition working
    // through ~250 dependent instructions per 32-column chunk (measured with the body removed: layer2 entry 246 -> 165 us,
    // layer3 entry 150 -> 98 us); they run with a second set of four warps that takes the upper half of the columns.
    const int lg = warp & 3;  // TMEM lane group this warp may access: lanes [32*lg, 32*lg+32)
    const int n_sets = ((int)blockDim.x - 64) >> 7, eset = (warp - 2) >> 2;
    const int row = lg * 32 + lane;
    const int wi = row % p.tw, hi = (row / p.tw) % p.th, bi = row / (p.tw * p.th);
    int it = 0;
    const uint32_t tempty_leader = CTAS == 2 ? mapa_rank(bar_tempty, 0u) : bar_tempty;
    for (int u = u_lo + u_first; u < u_hi; u += u_step)
    for (int prob = 0; prob < p.nprob; ++prob, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const Unit un = decode_unit(p, u);
      const TileCoord tc_ = decode_tile<CTAS>(p, un.tile, (int)rank);
      const float* e_scale = prob ? s_scale2 : s_scale;
      const float* e_bias = prob ? s_bias2 : s_bias;
      __nv_bfloat16* e_y = prob ? p.y2 : p.y;
      const int e_relu = prob ? p.relu2 : p.relu;
      const int nl = tc_.bt * p.nb + bi, oh = tc_.ht * p.th + hi, ow = tc_.wt * p.tw + wi;
      const bool valid = nl < p.B && oh < p.Ho && ow < p.Wo;
      const int ch0 = tc_.g * p.Cout + tc_.nt * p.block_n + un.n_off;  // index into scale/bias
      const int chn = tc_.nt * p.block_n + un.n_off;  // first channel of this unit inside its group
      const int Hop = p.o_H, Wop = p.o_W;
      const int64_t pix = ((int64_t)nl * Hop + oh * p.o_hm + p.o_ha) * Wop + ow * p.o_wm + p.o_wa;
      const int64_t sub_stride = (int64_t)p.B * Hop * Wop * p.split_c;
      const int nsplit = p.Cout / p.split_c;
      const bool use_res = p.residual != nullptr && valid && prob == 0;
      if (use_res && !p.out_f32) {
        // pull this row of the residual towards L2 while the MMAs of the tile are still running
        // (a residual implies split_c == Cout: the row's block_n channels are contiguous)
        const __nv_bfloat16* rrow = p.residual + (int64_t)tc_.g * sub_stride + pix * p.split_c + chn;
        for (int l = 0; l < un.bn * 2; l += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + l / 2));
      }
      mbar_wait(bar_tfull + 8 * as, aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * ACC_STRIDE);
      // columns of this warp: a multiple of 32 per set
      const int c_per = n_sets == 1 ? un.bn : (((un.bn >> 5) + 1) >> 1) << 5;
      const int c_lo = eset * c_per, c_hi = min(un.bn, c_lo + c_per);
      bool handed_back = false;   // LEAN: the accumulator is released as soon as its last chunk is in registers
      if constexpr (LEAN) {
        // two TMEM loads in flight: the next chunk's load is issued before the current chunk's arithmetic
        const float* sc = e_scale + ch0;
        const float* bs = e_bias + ch0;
        __nv_bfloat16* yrow = e_y + (int64_t)tc_.g * sub_stride + pix * p.Cout + chn;
        const __nv_bfloat16* rrow = p.residual + (int64_t)tc_.g * sub_stride + pix * p.Cout + chn;
        const bool wide = p.wide32 != 0;
        auto hand_back = [&]() {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) mbar_arrive_cluster(tempty_leader + 8 * as);
            else mbar_arrive(bar_tempty + 8 * as);
          }
          handed_back = true;
        };
        uint32_t accA[32], accB[32];
        if (c_lo < c_hi) tmem_ld_32x32b_x32(taddr + (uint32_t)c_lo, accA);
        for (int c0 = c_lo; c0 < c_hi; c0 += 64) {
          const bool has_b = c0 + 32 < c_hi, has_next = c0 + 64 < c_hi;
          tmem_ld_wait();
          if (has_b) tmem_ld_32x32b_x32(taddr + (uint32_t)(c0 + 32), accB);
          else hand_back();
          if (valid) {
            if (use_res) lean_chunk<true>(accA, sc + c0, bs + c0, rrow + c0, yrow + c0, e_relu != 0, wide);
            else lean_chunk<false>(accA, sc + c0, bs + c0, nullptr, yrow + c0, e_relu != 0, wide);
          }
          if (has_b) {
            tmem_ld_wait();
            if (has_next) tmem_ld_32x32b_x32(taddr + (uint32_t)(c0 + 64), accA);
            else hand_back();
            if (valid) {
              if (use_res) lean_chunk<true>(accB, sc + c0 + 32, bs + c0 + 32, rrow + c0 + 32, yrow + c0 + 32, e_relu != 0, wide);
              else lean_chunk<false>(accB, sc + c0 + 32, bs + c0 + 32, nullptr, yrow + c0 + 32, e_relu != 0, wide);
            }
          }
        }
      } else
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        // a 32-channel chunk never straddles output sub-tensors (split_c % 32 == 0)
        const int ch = chn + c0;
        const int64_t off = (int64_t)(tc_.g * nsplit + ch / p.split_c) * sub_stride + pix * p.split_c + ch % p.split_c - c0;
        uint4 rr[4];
        if (use_res && !p.out_f32) {  // issue the residual loads before the TMEM load so the latencies overlap
#pragma unroll
          for (int v = 0; v < 4; ++v) rr[v] = __ldg(reinterpret_cast<const uint4*>(p.residual + off + c0 + v * 8));
        }
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (valid && !(p.dbg & 1)) {
          const float4* sc4 = reinterpret_cast<const float4*>(e_scale + ch0 + c0);
          const float4* bs4 = reinterpret_cast<const float4*>(e_bias + ch0 + c0);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 s0 = sc4[2 * v], s1 = sc4[2 * v + 1], b0 = bs4[2 * v], b1 = bs4[2 * v + 1];
            float f[8];
            f[0] = fmaf(__uint_as_float(acc[v * 8 + 0]), s0.x, b0.x);
            f[1] = fmaf(__uint_as_float(acc[v * 8 + 1]), s0.y, b0.y);
            f[2] = fmaf(__uint_as_float(acc[v * 8 + 2]), s0.z, b0.z);
            f[3] = fmaf(__uint_as_float(acc[v * 8 + 3]), s0.w, b0.w);
            f[4] = fmaf(__uint_as_float(acc[v * 8 + 4]), s1.x, b1.x);
            f[5] = fmaf(__uint_as_float(acc[v * 8 + 5]), s1.y, b1.y);
            f[6] = fmaf(__uint_as_float(acc[v * 8 + 6]), s1.z, b1.z);
            f[7] = fmaf(__uint_as_float(acc[v * 8 + 7]), s1.w, b1.w);
            if (use_res && p.out_f32) {     // fp32 tensors (split-operand path): the residual is fp32 too
              const float* rf = reinterpret_cast<const float*>(p.residual) + off + c0 + v * 8;
              const float4 r0 = __ldg(reinterpret_cast<const float4*>(rf)), r1 = __ldg(reinterpret_cast<const float4*>(rf + 4));
              f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
              f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
            } else if (use_res) {
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr[v]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 rf = __bfloat1622float2(r2[j]);
                f[2 * j] += rf.x;
                f[2 * j + 1] += rf.y;
              }
            }
            if (e_relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.out_f32) {
              float* yf = reinterpret_cast<float*>(e_y) + off + c0 + v * 8;
              *reinterpret_cast<float4*>(yf) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(yf + 4) = make_float4(f[4], f[5], f[6], f[7]);
            } else {
              uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                   pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
              *reinterpret_cast<uint4*>(e_y + off + c0 + v * 8) = o;
            }
          }
        }
      }
      if (!handed_back) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_cluster(tempty_leader + 8 * as);
          else mbar_arrive(bar_tempty + 8 * as);
        }
      }
      if (p.out_pad == 1 && prob == 0 && valid) {
        // Physical zero border of the padded output (the consumer's 3x3 taps read it as their padding): written by the
        // threads that own the neighbouring interior pixels, for this tile's channels - no memset of the whole tensor.
        const int dh0 = (oh == 0) ? -1 : 0, dh1 = (oh == p.Ho - 1) ? 1 : 0;
        const int dw0 = (ow == 0) ? -1 : 0, dw1 = (ow == p.Wo - 1) ? 1 : 0;
        if ((dh0 | dh1 | dw0 | dw1) != 0) {
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          for (int dh = dh0; dh <= dh1; ++dh)
            for (int dw = dw0; dw <= dw1; ++dw) {
              if (dh == 0 && dw == 0) continue;
              const int64_t pixb = pix + (int64_t)dh * Wop + dw;
              for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                const int ch = chn + c0;
                __nv_bfloat16* d = e_y + (int64_t)(tc_.g * nsplit + ch / p.split_c) * sub_stride + pixb * p.split_c + ch % p.split_c;
#pragma unroll
                for (int v = 0; v < 4; ++v) *reinterpret_cast<uint4*>(d + v * 8) = z;
              }
            }
        }
      }
    }
  }

  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}
static int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

static bool supported(int H, int W, int Cin, int Cout, int sh, int sw) {
  if (Cin % 64 != 0 || Cout % 32 != 0) return false;
  if (Cout > 256 && Cout % 256 != 0) return false;
  if (sh != 1 && sh != 2) return false;
  if (sw != 1 && sw != 2) return false;
  if (sh == 2 && (H & 1)) return false;
  if (sw == 2 && (W & 1)) return false;
  return true;
}

}  // namespace tc

int amoe_conv_tc_init(amoe_ctx* ctx) {
  AMOE_ENTER(ctx);
  (void)ctx;
  const int smem_max = tc::SMEM_BUDGET + 1024 + 24 * 1024;
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(tc::conv_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(tc::conv_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(tc::conv_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(tc::conv_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  return 0;
}

int amoe_conv2d_simt(amoe_ctx* ctx, const void* x, const void* w, const float* scale,
                     const float* bias, const void* residual, void* y, int G, int x_shared, int B,
                     int H, int W, int Cin, int Cout, int KH, int KW, int sh, int sw, int ph, int pw,
                     int Ho, int Wo, int relu, int dtype, cudaStream_t st);

// How the kernel sees the input: a 5-D tensor (inner -> outer) {Cv, Wv, P, Hv, N} with byte
// strides for dims 1..4 (dim 0 is contiguous bf16).
struct AView {
  uint64_t dims[5];
  uint64_t strides[4];
};

// where / how the output tile is stored when it is not the dense bf16 tensor (training convolutions: fp32, parity-strided)
struct OutMap {
  int f32 = 0;
  int o_hm = 1, o_ha = 0, o_wm = 1, o_wa = 0, o_H = 0, o_W = 0;   // o_H == 0: dense addressing from out_pad
};

// second convolution of a dual launch (see Params::nprob): 1x1 weights [G*Cout][Cin], its folded BatchNorm and output
struct Second {
  const void* w = nullptr;
  const float* scale = nullptr;
  const float* bias = nullptr;
  void* y = nullptr;
  int relu = 0;
  tc::Tap tap;
};

static int launch_generic(amoe_ctx* ctx, const void* x, const AView& av, const void* w, int Ktot,
                          const float* scale, const float* bias, const void* residual, void* y, int G,
                          int x_shared, int B, int Ho, int Wo, int Cout, int split_c, int num_taps,
                          const tc::Tap* taps, int k_chunks, int relu, int out_pad, cudaStream_t st,
                          const Second* second = nullptr, const OutMap* omap = nullptr) {
  using namespace tc;
  AMOE_REQUIRE(num_taps <= MAX_TAPS, "conv_tc: %d taps exceed the limit of %d", num_taps, MAX_TAPS);
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
               "conv_tc: pointers must be 16-byte aligned");
  AMOE_REQUIRE(Cout % 32 == 0 && split_c % 32 == 0 && Cout % split_c == 0, "conv_tc: bad Cout/split_c %d/%d", Cout, split_c);
  Params p;
  p.tw = std::min(128, pow2_ceil(Wo));
  p.th = std::min(128 / p.tw, pow2_ceil(Ho));
  p.nb = 128 / (p.tw * p.th);
  p.tiles_w = ceil_div(Wo, p.tw);
  p.tiles_h = ceil_div(Ho, p.th);
  p.tiles_b = ceil_div(B, p.nb);
  p.block_n = Cout >= 256 ? 256 : Cout;
  {
    // experiment switch: N tile for wide layers (AMOE_TC_BLOCKN=128 trades operand re-reads for finer wave granularity)
    static const int forced = [] { const char* e = getenv("AMOE_TC_BLOCKN"); return e ? atoi(e) : 0; }();
    if (forced >= 32 && forced <= 256 && Cout % forced == 0 && Cout > forced) p.block_n = forced;
  }
  AMOE_REQUIRE(Cout % p.block_n == 0 && p.block_n % 32 == 0, "conv_tc: unsupported channel tiling Cout=%d", Cout);
  p.n_tiles_n = Cout / p.block_n;
  // CTA pairs (see the kernel): AMOE_TC_PAIR = 0 off, 1 wherever possible, unset: the N = 256 tiles
  bool pair;
  {
    const char* e = getenv("AMOE_TC_PAIR");
    const int mode = e ? atoi(e) : -1;
    // (dual stage-entry launches too, since the lean epilogue: layer3 entry 144 -> 138 us, layer4 entry 118 -> 110 us; N = 128
    // pairs stay off - layer2 entry 222 -> 322 us, an N = 128 pair MMA step takes ~98 instead of 64 cycles)
    pair = mode == 1 ? (p.block_n >= 64) : (mode == -1 ? (p.block_n == 256) : false);
    if (p.tiles_b < 2 || ctx->sm_count < 2) pair = false;
  }
  // resident weights (see Params): short-K launches with one N tile whose weight boxes fit beside >= 4 activation stages;
  // CTA pairs halve what each CTA holds.  AMOE_TC_WRES=0 switches it off.
  p.w_resident = 0;
  {
    const char* e = getenv("AMOE_TC_WRES");
    const char* ep = getenv("AMOE_TC_PAIR");
    const bool pair_allowed = ep == nullptr || atoi(ep) != 0;
    const int boxes = num_taps * k_chunks + (second != nullptr ? k_chunks : 0);
    const bool wanted = e != nullptr ? atoi(e) != 0 : num_taps * k_chunks <= 12;   // long K: the weight stream is amortised
    if (wanted && p.n_tiles_n == 1 && p.block_n >= 64 && omap == nullptr && p.tiles_b >= 2 && ctx->sm_count / G >= 2) {
      // (CTA pairs would halve the resident bytes, but an N = 128 pair MMA step takes ~98 instead of 64 cycles: measured
      // slower - layer2 entry 156 -> 184 us, policy conv3 53 -> 61 us - so the boxes must fit one CTA, beside >= 3 stages)
      (void)pair_allowed;
      if (boxes * (p.block_n / (pair ? 2 : 1)) * 128 + 3 * A_STAGE_BYTES <= SMEM_RESIDENT) p.w_resident = 1;
    }
  }
  const int ctas = pair ? 2 : 1;
  if (pair) p.tiles_b = ceil_div(p.tiles_b, 2);     // pairs of image tiles
  p.G = G; p.B = B; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.split_c = split_c; p.out_pad = out_pad;
  { const char* e = getenv("AMOE_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  {
    const char* e = getenv("AMOE_TC_W32");
    const bool al32 = (reinterpret_cast<uintptr_t>(y) & 31) == 0 && (reinterpret_cast<uintptr_t>(residual) & 31) == 0 &&
                      (second == nullptr || (reinterpret_cast<uintptr_t>(second->y) & 31) == 0);
    p.wide32 = ((e == nullptr || atoi(e) != 0) && al32) ? 1 : 0;
  }
  p.reverse = ctx->walk_reverse;
  p.num_taps = num_taps;
  p.k_chunks = k_chunks;
  const int stage_bytes = p.w_resident ? A_STAGE_BYTES : A_STAGE_BYTES + (p.block_n / ctas) * 128;
  const int w_res_bytes = p.w_resident ? (num_taps * k_chunks + (second != nullptr ? k_chunks : 0)) * (p.block_n / ctas) * 128 : 0;
  // eight epilogue warps (AMOE_TC_EPI8 = 1; measured neutral, off by default)
  int threads = NUM_THREADS;
  { const char* e = getenv("AMOE_TC_EPI8"); if (e != nullptr && atoi(e) != 0 && p.block_n >= 64) threads = NUM_THREADS_EPI8; }
  const int stg_bytes = 0;
  const int sb_bytes = (second != nullptr ? 4 : 2) * G * Cout * (int)sizeof(float);
  {
    const int limit = SMEM_BUDGET + 24 * 1024;          // dynamic shared memory the kernel may use (+ 1 KB alignment slack)
    const int normal = (p.w_resident ? SMEM_RESIDENT : SMEM_BUDGET) - w_res_bytes;
    p.stages = std::min(MAX_STAGES, std::min(normal, limit - w_res_bytes - sb_bytes - stg_bytes) / stage_bytes);
    AMOE_REQUIRE(p.stages >= 2, "conv_tc: shared memory budget exceeded");
  }
  p.relu = relu; p.x_shared = x_shared;
  int64_t total = (int64_t)G * p.tiles_b * p.tiles_h * p.tiles_w * p.n_tiles_n;
  AMOE_REQUIRE(total < (1ll << 31), "conv_tc: too many tiles");
  p.total_tiles = (int)total;
  p.scale = scale; p.bias = bias;
  p.n_ch_total = G * Cout;
  p.residual = (const __nv_bfloat16*)residual;
  p.y = (__nv_bfloat16*)y;
  for (int t = 0; t < num_taps; ++t) p.taps[t] = taps[t];
  p.out_f32 = 0;
  p.o_hm = 1; p.o_ha = out_pad; p.o_wm = 1; p.o_wa = out_pad; p.o_H = Ho + 2 * out_pad; p.o_W = Wo + 2 * out_pad;
  if (omap != nullptr) {
    p.out_f32 = omap->f32;
    if (omap->o_H > 0) { p.o_hm = omap->o_hm; p.o_ha = omap->o_ha; p.o_wm = omap->o_wm; p.o_wa = omap->o_wa; p.o_H = omap->o_H; p.o_W = omap->o_W; }
  }
  p.nprob = 1; p.relu2 = 0; p.tap2 = taps[0]; p.scale2 = nullptr; p.bias2 = nullptr; p.y2 = nullptr;
  if (second != nullptr) {
    AMOE_REQUIRE(second->w && second->scale && second->bias && second->y && split_c == Cout,
                 "conv_tc: incomplete second convolution of a dual launch");
    AMOE_REQUIRE((reinterpret_cast<uintptr_t>(second->w) & 15) == 0 && (reinterpret_cast<uintptr_t>(second->y) & 15) == 0,
                 "conv_tc: pointers must be 16-byte aligned");
    p.nprob = 2; p.relu2 = second->relu; p.tap2 = second->tap; p.scale2 = second->scale; p.bias2 = second->bias;
    p.y2 = (__nv_bfloat16*)second->y;
  }
  if (total == 0) return 0;
  int grid = std::min(p.total_tiles, ctx->sm_count / ctas);    // CTAs, or CTA pairs
  p.tiles_per_group = p.total_tiles / G;
  if (p.w_resident) grid = std::max(1, std::min(p.tiles_per_group, ctx->sm_count / ctas / G));   // per expert group (grid.y = G)
  p.total_units = p.total_tiles; p.full_units = p.total_tiles; p.split = 1; p.sub_n = p.block_n;
  {
    // tail splitting (see Params): only for long-K single-problem launches whose last wave is partial
    const char* e_ts = getenv("AMOE_TC_TAIL_SPLIT");
    const int tail_on = (e_ts == nullptr || atoi(e_ts) != 0) ? 1 : 0;
    const int rem = p.total_tiles % grid;
    if (tail_on && !p.w_resident && second == nullptr && p.total_tiles > grid && rem != 0 && num_taps * k_chunks >= 16) {
      int best = 1;
      double best_cost = 1.0;                       // time of the tail in tile times
      for (int sp = 2; sp <= 4; sp *= 2) {
        if (p.block_n % (32 * sp) != 0) continue;
        // narrow units pay for their share of the tile more than once: every unit re-reads the activation boxes, and an
        // N = 64 MMA step is shared-memory-bound (measured, tools/tc_bench.py: the 512->256 head, 88 tiles in the tail,
        // got slower with four N = 64 units per tile: 86 -> 101 us; layer4, 28 tiles in the tail: 168 -> 163 us)
        const double penalty = p.block_n / sp >= 128 ? 1.25 : 1.9;
        const double cost = (double)ceil_div(rem * sp, grid) / sp * penalty;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
      }
      if (best > 1) {
        p.split = best; p.sub_n = p.block_n / best;
        p.full_units = p.total_tiles - rem;
        p.total_units = p.full_units + rem * best;
      }
    }
  }

  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[5], strides[4];
    for (int i = 0; i < 5; ++i) dims[i] = av.dims[i];
    for (int i = 0; i < 4; ++i) strides[i] = av.strides[i];
    cuuint32_t box[5] = {(cuuint32_t)BLOCK_K, (cuuint32_t)p.tw, 1u, (cuuint32_t)p.th, (cuuint32_t)p.nb};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = ctx->encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(activations) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)G * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(p.block_n / ctas)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  }
  CUtensorMap tmW2 = tmW;
  if (p.split > 1) {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)G * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(p.sub_n / ctas)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(sub-tile weights) failed with %d", (int)r);
  }
  if (second != nullptr) {
    const int K2 = k_chunks * BLOCK_K;          // 1x1 filter: K = Cin
    cuuint64_t dims[2] = {(cuuint64_t)K2, (cuuint64_t)G * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K2 * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(p.block_n / ctas)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(second->w), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled(second weights) failed with %d", (int)r);
  }
  const size_t smem = (size_t)p.stages * stage_bytes + w_res_bytes + 1024 + (size_t)sb_bytes + stg_bytes;
  AMOE_REQUIRE(smem <= (size_t)SMEM_BUDGET + 1024 + 24 * 1024, "conv_tc: too many channels for the shared-memory scale/bias stage (%zu bytes)", smem);
  const int gy = p.w_resident ? G : 1;
  // lean bf16 epilogue (see lean_chunk) wherever it applies; AMOE_TC_LEAN=0 keeps the generic one (A/B switch)
  const bool lean_on = [] { const char* e = getenv("AMOE_TC_LEAN"); return e == nullptr || atoi(e) != 0; }();
  const bool lean = lean_on && !p.out_f32 && split_c == Cout && p.dbg == 0;
  if (pair) {
    if (lean) AMOE_CHECK_CUDA(amoe_launch_pdl_cluster(conv_tc_kernel<2, true>, 2, dim3(2 * grid, gy), dim3(threads), smem, st, tmA, tmW, tmW2, p));
    else AMOE_CHECK_CUDA(amoe_launch_pdl_cluster(conv_tc_kernel<2, false>, 2, dim3(2 * grid, gy), dim3(threads), smem, st, tmA, tmW, tmW2, p));
  } else {
    if (lean) AMOE_CHECK_CUDA(amoe_launch_pdl(conv_tc_kernel<1, true>, dim3(grid, gy), dim3(threads), smem, st, tmA, tmW, tmW2, p));
    else AMOE_CHECK_CUDA(amoe_launch_pdl(conv_tc_kernel<1, false>, dim3(grid, gy), dim3(threads), smem, st, tmA, tmW, tmW2, p));
  }
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

static int conv_tc_launch(amoe_ctx* ctx, const void* x, const void* w, const float* scale,
                          const float* bias, const void* residual, void* y, int G, int x_shared,
                          int B, int H, int W, int Cin, int Cout, int KH, int KW, int sh, int sw,
                          int ph, int pw, int Ho, int Wo, int relu, int in_pad, int out_pad, cudaStream_t st,
                          Second* second = nullptr) {
  using namespace tc;
  // a physically padded input [N][H+2*in_pad][W+2*in_pad][C] is just a bigger image whose
  // filter taps start in_pad pixels further right/down
  H += 2 * in_pad; W += 2 * in_pad; ph -= in_pad; pw -= in_pad;
  AMOE_REQUIRE(KH * KW <= MAX_TAPS, "conv_tc: %dx%d filter has more than %d taps", KH, KW, MAX_TAPS);
  Tap taps[MAX_TAPS];
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      Tap& t = taps[kh * KW + kw];
      int ho = kh - ph, wo = kw - pw;
      if (sh == 1) { t.dh = ho; t.hp = 0; } else { t.dh = floordiv2(ho); t.hp = ho - 2 * t.dh; }
      if (sw == 1) { t.dw = wo; t.c_off = 0; } else { t.dw = floordiv2(wo); t.c_off = (wo - 2 * t.dw) * Cin; }
      t.w_k = (kh * KW + kw) * Cin;
    }
  // parity view (inner -> outer): Cv = sw*Cin, Wv = W/sw, P = sh, Hv = H/sh, N
  const int NB = x_shared ? B : G * B;
  AView av;
  av.dims[0] = (uint64_t)sw * Cin; av.dims[1] = (uint64_t)(W / sw); av.dims[2] = (uint64_t)sh;
  av.dims[3] = (uint64_t)(H / sh); av.dims[4] = (uint64_t)NB;
  av.strides[0] = (uint64_t)sw * Cin * 2; av.strides[1] = (uint64_t)W * Cin * 2;
  av.strides[2] = (uint64_t)sh * W * Cin * 2; av.strides[3] = (uint64_t)H * W * Cin * 2;
  if (second != nullptr) {
    // 1x1 / same stride / no padding over the (physically padded) input: the tap at offset (in_pad, in_pad)
    Tap& t = second->tap;
    const int ho = in_pad, wo = in_pad;      // ph, pw were shifted by in_pad above; a 1x1/p0 filter reads pixel (s*oh, s*ow)
    if (sh == 1) { t.dh = ho; t.hp = 0; } else { t.dh = floordiv2(ho); t.hp = ho - 2 * t.dh; }
    if (sw == 1) { t.dw = wo; t.c_off = 0; } else { t.dw = floordiv2(wo); t.c_off = (wo - 2 * t.dw) * Cin; }
    t.w_k = 0;
  }
  return launch_generic(ctx, x, av, w, KH * KW * Cin, scale, bias, residual, y, G, x_shared, B, Ho, Wo, Cout,
                        Cout, KH * KW, taps, Cin / BLOCK_K, relu, out_pad, st, second);
}


// ---------------------------------------------------------------------------------------------------------
// fp32-accurate convolutions on the bf16 tensor cores (training path: the reference trains in fp32).
//
// Every fp32 value is split into three bf16 parts, v = v1 + v2 + v3 (each the bf16 rounding of what is left), which
// carries 24 significant bits.  A product x*w keeps the six terms down to 2^-16 of its magnitude,
//     x1 w1 + (x1 w2 + x2 w1) + (x1 w3 + x2 w2 + x3 w1),
// all accumulated in the fp32 TMEM accumulator, so a dot product is as accurate as an fp32 FMA chain (what is dropped,
// x2 w3 + x3 w2 + x3 w3, is below 2^-24).  The six terms ride through the SAME implicit-GEMM kernel as six K blocks per
// filter tap: activations are stored once as [pixel][x1 | x2 | x3] (3C channels) and the taps address a part through
// their channel offset; the weights are packed per tap as [w1 | w2 | w1 | w3 | w2 | w1].  Six bf16 MMAs per fp32 MAC is
// ~230 TFLOP/s of fp32-accurate throughput against ~60 TFLOP/s of the FP32 pipe.
namespace tc {

constexpr int SPLIT_TERMS = 6;
__device__ __constant__ int kTermXPart[SPLIT_TERMS] = {0, 0, 1, 0, 1, 2};   // activation part of each term
__device__ __constant__ int kTermWPart[SPLIT_TERMS] = {0, 1, 0, 2, 1, 0};   // weight part of each term
static const int hTermXPart[SPLIT_TERMS] = {0, 0, 1, 0, 1, 2};

// x [rows][C] fp32 -> out [rows][3C] bf16 = (x1 | x2 | x3); C % 8 == 0; one thread = 8 channels
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C,
                                                     int64_t total8) {
  const int c8n = C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / c8n;
    const int c0 = (int)(i - row * c8n) << 3;
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(x + row * C + c0));
    const float4 u1 = __ldg(reinterpret_cast<const float4*>(x + row * C + c0 + 4));
    const float v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
    __align__(16) __nv_bfloat16 a[8], b[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split3(v[j], a[j], b[j], c[j]);
    __nv_bfloat16* o = out + row * (3 * (int64_t)C) + c0;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(a);
    *reinterpret_cast<uint4*>(o + C) = *reinterpret_cast<const uint4*>(b);
    *reinterpret_cast<uint4*>(o + 2 * C) = *reinterpret_cast<const uint4*>(c);
  }
}

// nn.Conv2d weight [Cout][Cin][KH][KW] fp32 -> six split terms per tap, K-major:
//   transposed == 0 (forward):  dst[co][kh][kw][t][ci]            rows = Cout, K = KH*KW*6*Cin
//   transposed == 1 (dgrad):    dst[ci][kh][kw][t][co]            rows = Cin,  K = KH*KW*6*Cout
__global__ void __launch_bounds__(256) pack_split6_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout,
                                                          int Cin, int KH, int KW, int transposed, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates (row, kh, kw, inner) without the term axis
    const int inner_n = transposed ? Cout : Cin;
    const int inner = (int)(i % inner_n);
    int64_t r = i / inner_n;
    const int kw = (int)(r % KW); r /= KW;
    const int kh = (int)(r % KH); r /= KH;
    const int row = (int)r;
    const int co = transposed ? inner : row, ci = transposed ? row : inner;
    const float v = w[(((int64_t)co * Cin + ci) * KH + kh) * KW + kw];
    __nv_bfloat16 part[3];
    split3(v, part[0], part[1], part[2]);
    __nv_bfloat16* d = dst + ((((int64_t)row * KH + kh) * KW + kw) * SPLIT_TERMS) * inner_n + inner;
#pragma unroll
    for (int t = 0; t < SPLIT_TERMS; ++t) d[(int64_t)t * inner_n] = part[kTermWPart[t]];
  }
}

}  // namespace tc

static unsigned tc_grid(const amoe_ctx* ctx, int64_t items) {
  const int64_t want = (items + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

// Common launcher of the split-operand convolutions.  xs: [NB][Hin][Win][3*C] bf16 (split activations or split dy),
// ws: packed split weights with `rows` output channels and K = ntap_spatial*6*C.  Spatial taps are given as
// (dh, dw, weight tap index); stride applies to the INPUT view (1 for dgrad sub-convolutions).
struct SpatialTap { int dh, dw, widx; };
static int launch_split(amoe_ctx* ctx, const void* xs, const void* ws, const float* scale, const float* bias, float* y,
                        int NB, int Hin, int Win, int C, int rows, int n_wtaps, const SpatialTap* st, int nst, int sh, int sw,
                        int Ho, int Wo, int relu, const OutMap& om, cudaStream_t stream, int G = 1, int x_shared = 0,
                        const float* residual = nullptr) {
  using namespace tc;
  AMOE_REQUIRE(C % BLOCK_K == 0 && rows % 32 == 0 && (rows <= 256 || rows % 256 == 0),
               "split conv: needs C %% 64 == 0 and Cout %% 32 == 0 (C=%d Cout=%d)", C, rows);
  AMOE_REQUIRE(nst * SPLIT_TERMS <= MAX_TAPS, "split conv: %d spatial taps x 6 terms exceed %d", nst, MAX_TAPS);
  AMOE_REQUIRE((sh == 1 || (sh == 2 && Hin % 2 == 0)) && (sw == 1 || (sw == 2 && Win % 2 == 0)), "split conv: stride 2 needs even H, W");
  Tap taps[MAX_TAPS];
  int n = 0;
  // Term-major, smallest terms first: the tensor core truncates (rounds toward zero) when it adds into the fp32
  // accumulator, an error relative to the running sum - so the 2^-16 and 2^-8 terms are accumulated while the sum is
  // still small, and only the nst*k_chunks*4 leading-term MMAs (what any bf16 GEMM pays) add into the full-size sum.
  for (int t = SPLIT_TERMS - 1; t >= 0; --t)
    for (int i = 0; i < nst; ++i) {
      Tap& tp = taps[n++];
      const int ho = st[i].dh, wo = st[i].dw;
      if (sh == 1) { tp.dh = ho; tp.hp = 0; } else { tp.dh = floordiv2(ho); tp.hp = ho - 2 * tp.dh; }
      if (sw == 1) { tp.dw = wo; tp.c_off = 0; } else { tp.dw = floordiv2(wo); tp.c_off = (wo - 2 * tp.dw) * 3 * C; }
      tp.c_off += hTermXPart[t] * C;
      tp.w_k = (st[i].widx * SPLIT_TERMS + t) * C;
    }
  const int C3 = 3 * C;
  AView av;
  av.dims[0] = (uint64_t)sw * C3; av.dims[1] = (uint64_t)(Win / sw); av.dims[2] = (uint64_t)sh;
  av.dims[3] = (uint64_t)(Hin / sh); av.dims[4] = (uint64_t)(x_shared ? NB : G * NB);
  av.strides[0] = (uint64_t)sw * C3 * 2; av.strides[1] = (uint64_t)Win * C3 * 2;
  av.strides[2] = (uint64_t)sh * Win * C3 * 2; av.strides[3] = (uint64_t)Hin * Win * C3 * 2;
  return launch_generic(ctx, xs, av, ws, n_wtaps * SPLIT_TERMS * C, scale, bias, residual, y, G, x_shared, NB, Ho, Wo, rows, rows, n,
                        taps, C / BLOCK_K, relu, 0, stream, nullptr, &om);
}

extern "C" {

int amoe_conv2d_tc_supported(int H, int W, int Cin, int Cout, int stride_h, int stride_w) {
  return tc::supported(H, W, Cin, Cout, stride_h, stride_w) ? 1 : 0;
}

int amoe_conv2d_rowwin_fwd(amoe_ctx* ctx, const void* x, const void* w, const float* scale,
                           const float* bias, void* y, int B, int H, int Wpad, int Cp, int Cout,
                           int split_c, int KH, int stride_h, int stride_w, int pad_h, int Ho, int Wo,
                           int relu, void* stream) {
  AMOE_ENTER(ctx);
  using namespace tc;
  AMOE_REQUIRE(ctx && x && w && scale && bias && y, "amoe_conv2d_rowwin_fwd: NULL argument");
  AMOE_REQUIRE(Cp > 0 && BLOCK_K % Cp == 0, "amoe_conv2d_rowwin_fwd: Cp=%d must divide %d", Cp, BLOCK_K);
  const int win = BLOCK_K / Cp;  // pixels per window
  AMOE_REQUIRE((stride_w * Cp * 2) % 16 == 0, "amoe_conv2d_rowwin_fwd: window stride must be a multiple of 16 bytes");
  AMOE_REQUIRE((Wo - 1) * stride_w + win <= Wpad, "amoe_conv2d_rowwin_fwd: windows run past the padded row (Wpad=%d)", Wpad);
  AMOE_REQUIRE((Wpad * Cp * 2) % 16 == 0, "amoe_conv2d_rowwin_fwd: padded row must be a multiple of 16 bytes");
  AMOE_REQUIRE(stride_h == 1 || (stride_h == 2 && H % 2 == 0), "amoe_conv2d_rowwin_fwd: stride_h must be 1, or 2 with even H");
  AMOE_REQUIRE(KH <= MAX_TAPS, "amoe_conv2d_rowwin_fwd: KH too large");
  Tap taps[MAX_TAPS];
  for (int kh = 0; kh < KH; ++kh) {
    int ho = kh - pad_h;
    taps[kh].c_off = 0;
    taps[kh].dw = 0;
    taps[kh].w_k = kh * BLOCK_K;
    if (stride_h == 1) { taps[kh].dh = ho; taps[kh].hp = 0; } else { taps[kh].dh = floordiv2(ho); taps[kh].hp = ho - 2 * taps[kh].dh; }
  }
  // overlapping windows: dim0 = 64 contiguous elements (win pixels), dim1 = output column (stride_w pixels apart)
  AView av;
  const uint64_t row = (uint64_t)Wpad * Cp * 2;
  av.dims[0] = BLOCK_K; av.dims[1] = (uint64_t)Wo; av.dims[2] = (uint64_t)stride_h;
  av.dims[3] = (uint64_t)(H / stride_h); av.dims[4] = (uint64_t)B;
  av.strides[0] = (uint64_t)stride_w * Cp * 2; av.strides[1] = row; av.strides[2] = (uint64_t)stride_h * row;
  av.strides[3] = (uint64_t)H * row;
  return launch_generic(ctx, x, av, w, KH * BLOCK_K, scale, bias, nullptr, y, 1, 0, B, Ho, Wo, Cout, split_c, KH,
                        taps, 1, relu, 0, (cudaStream_t)stream);
}

int amoe_conv2d_dual_fwd(amoe_ctx* ctx, const void* x, const void* w, const float* scale, const float* bias, void* y,
                         const void* w_1x1, const float* scale2, const float* bias2, void* y2, int G, int B, int H, int W,
                         int Cin, int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, int relu, int relu2,
                         int in_pad, int out_pad, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && w && scale && bias && y && w_1x1 && scale2 && bias2 && y2, "amoe_conv2d_dual_fwd: NULL argument");
  AMOE_REQUIRE(G >= 1 && B >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && Ho > 0 && Wo > 0 &&
                   in_pad >= 0 && out_pad >= 0, "amoe_conv2d_dual_fwd: bad shape");
  AMOE_REQUIRE((Ho - 1) * stride - pad < H && (Wo - 1) * stride - pad < W && (Ho - 1) * stride < H && (Wo - 1) * stride < W,
               "amoe_conv2d_dual_fwd: Ho/Wo inconsistent with input size");
  AMOE_REQUIRE(tc::supported(H + 2 * in_pad, W + 2 * in_pad, Cin, Cout, stride, stride),
               "amoe_conv2d_dual_fwd: the tcgen05 path does not take this shape");
  Second second;
  second.w = w_1x1; second.scale = scale2; second.bias = bias2; second.y = y2; second.relu = relu2;
  return conv_tc_launch(ctx, x, w, scale, bias, nullptr, y, G, 0, B, H, W, Cin, Cout, KH, KW, stride, stride, pad, pad, Ho, Wo,
                        relu, in_pad, out_pad, (cudaStream_t)stream, &second);
}


int amoe_split3_bf16(amoe_ctx* ctx, const float* x, void* out, int64_t rows, int C, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && (rows == 0 || (x && out)), "amoe_split3_bf16: NULL argument");
  AMOE_REQUIRE(C % 8 == 0 && C > 0 && rows >= 0, "amoe_split3_bf16: C must be a positive multiple of 8");
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "amoe_split3_bf16: pointers must be 16-byte aligned");
  const int64_t total8 = rows * (C / 8);
  if (total8 == 0) return 0;
  tc::split3_kernel<<<tc_grid(ctx, total8), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, C, total8);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_pack_conv_weight_split6(amoe_ctx* ctx, const float* w_oihw, void* dst, int Cout, int Cin, int KH, int KW,
                                 int transposed, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && w_oihw && dst, "amoe_pack_conv_weight_split6: NULL argument");
  const int64_t total = (int64_t)Cout * Cin * KH * KW;
  if (total == 0) return 0;
  tc::pack_split6_kernel<<<tc_grid(ctx, total), 256, 0, (cudaStream_t)stream>>>(w_oihw, (__nv_bfloat16*)dst, Cout, Cin, KH, KW,
                                                                                transposed ? 1 : 0, total);
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

int amoe_conv2d_f32tc_supported(int H, int W, int Cin, int Cout, int KH, int KW, int stride) {
  return (Cin % 64 == 0 && Cout % 32 == 0 && (Cout <= 256 || Cout % 256 == 0) && KH * KW * tc::SPLIT_TERMS <= tc::MAX_TAPS &&
          (stride == 1 || (stride == 2 && H % 2 == 0 && W % 2 == 0))) ? 1 : 0;
}

int amoe_conv2d_fwd_f32tc(amoe_ctx* ctx, const void* x_split, const void* w_split, const float* scale, const float* bias, float* y,
                          int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo, int relu,
                          void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x_split && w_split && scale && bias && y, "amoe_conv2d_fwd_f32tc: NULL argument");
  AMOE_REQUIRE(amoe_conv2d_f32tc_supported(H, W, Cin, Cout, KH, KW, stride), "amoe_conv2d_fwd_f32tc: unsupported shape");
  AMOE_REQUIRE((Ho - 1) * stride - pad < H && (Wo - 1) * stride - pad < W && Ho > 0 && Wo > 0, "amoe_conv2d_fwd_f32tc: bad Ho/Wo");
  SpatialTap st[tc::MAX_TAPS];
  int n = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) st[n++] = SpatialTap{kh - pad, kw - pad, kh * KW + kw};
  OutMap om;
  om.f32 = 1;
  return launch_split(ctx, x_split, w_split, scale, bias, y, B, H, W, Cin, Cout, KH * KW, st, n, stride, stride, Ho, Wo, relu, om,
                      (cudaStream_t)stream);
}

int amoe_conv2d_fwd_f32tc_grouped(amoe_ctx* ctx, const void* x_split, const void* w_split, const float* scale, const float* bias,
                                  const float* residual, float* y, int G, int x_shared, int B, int H, int W, int Cin, int Cout, int KH,
                                  int KW, int stride, int pad, int Ho, int Wo, int relu, void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x_split && w_split && scale && bias && y && G >= 1, "amoe_conv2d_fwd_f32tc_grouped: NULL argument");
  AMOE_REQUIRE(amoe_conv2d_f32tc_supported(H, W, Cin, Cout, KH, KW, stride), "amoe_conv2d_fwd_f32tc_grouped: unsupported shape");
  AMOE_REQUIRE((Ho - 1) * stride - pad < H && (Wo - 1) * stride - pad < W && Ho > 0 && Wo > 0, "amoe_conv2d_fwd_f32tc_grouped: bad Ho/Wo");
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(residual) & 15) == 0, "amoe_conv2d_fwd_f32tc_grouped: residual must be 16-byte aligned");
  SpatialTap st[tc::MAX_TAPS];
  int n = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) st[n++] = SpatialTap{kh - pad, kw - pad, kh * KW + kw};
  OutMap om;
  om.f32 = 1;
  return launch_split(ctx, x_split, w_split, scale, bias, y, B, H, W, Cin, Cout, KH * KW, st, n, stride, stride, Ho, Wo, relu, om,
                      (cudaStream_t)stream, G, x_shared, residual);
}

int amoe_conv2d_bwd_data_f32tc(amoe_ctx* ctx, const void* dy_split, const void* wT_split, const float* ones, const float* zeros,
                               float* dx, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo,
                               void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && dy_split && wT_split && ones && zeros && dx, "amoe_conv2d_bwd_data_f32tc: NULL argument");
  // roles swap: the "input" is dy [B,Ho,Wo,3*Cout], the "output channels" are Cin
  AMOE_REQUIRE(Cout % 64 == 0 && Cin % 32 == 0 && (Cin <= 256 || Cin % 256 == 0) && (stride == 1 || stride == 2),
               "amoe_conv2d_bwd_data_f32tc: unsupported shape");
  AMOE_REQUIRE(stride == 1 || (H % 2 == 0 && W % 2 == 0), "amoe_conv2d_bwd_data_f32tc: stride 2 needs even H, W");
  cudaStream_t st_ = (cudaStream_t)stream;
  // dx[ih] = sum_kh dy[(ih + pad - kh) / stride] w[kh]  over the kh with (ih + pad - kh) % stride == 0
  for (int a = 0; a < stride; ++a)
    for (int b = 0; b < stride; ++b) {
      SpatialTap st[tc::MAX_TAPS];
      int n = 0;
      for (int kh = 0; kh < KH; ++kh) {
        if ((a + pad - kh) % stride != 0) continue;
        for (int kw = 0; kw < KW; ++kw) {
          if ((b + pad - kw) % stride != 0) continue;
          // C++ division of possibly negative numerators: (a + pad - kh) is a multiple of stride here
          st[n++] = SpatialTap{(a + pad - kh) / stride, (b + pad - kw) / stride, kh * KW + kw};
        }
      }
      const int Hs = (H - a + stride - 1) / stride, Ws = (W - b + stride - 1) / stride;   // rows / columns of this parity class
      OutMap om;
      om.f32 = 1; om.o_hm = stride; om.o_ha = a; om.o_wm = stride; om.o_wa = b; om.o_H = H; om.o_W = W;
      if (n == 0) continue;     // no filter tap reaches this parity class: the caller zero-fills dx (1x1 / stride 2)
      AMOE_REQUIRE(n * tc::SPLIT_TERMS <= tc::MAX_TAPS, "amoe_conv2d_bwd_data_f32tc: too many taps");
      int rc = launch_split(ctx, dy_split, wT_split, ones, zeros, dx, B, Ho, Wo, Cout, Cin, KH * KW, st, n, 1, 1, Hs, Ws, 0, om, st_);
      if (rc) return rc;
    }
  return 0;
}

int amoe_conv2d_fwd(amoe_ctx* ctx, const void* x, const void* w, const float* scale,
                    const float* bias, const void* residual, void* y, int G, int x_shared, int B,
                    int H, int W, int Cin, int Cout, int KH, int KW, int stride_h, int stride_w,
                    int pad_h, int pad_w, int Ho, int Wo, int relu, int dtype, int impl, int in_pad, int out_pad,
                    void* stream) {
  AMOE_ENTER(ctx);
  AMOE_REQUIRE(ctx && x && w && scale && bias && y, "amoe_conv2d_fwd: NULL argument");
  AMOE_REQUIRE(dtype == AMOE_F32 || dtype == AMOE_BF16, "amoe_conv2d_fwd: bad dtype %d", dtype);
  AMOE_REQUIRE(G >= 1 && B >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 &&
                   stride_h > 0 && stride_w > 0 && Ho > 0 && Wo > 0,
               "amoe_conv2d_fwd: bad shape");
  // every tap of every output must stay within [-(K), H+K): padding is implicit zero
  AMOE_REQUIRE((Ho - 1) * stride_h - pad_h < H && (Wo - 1) * stride_w - pad_w < W,
               "amoe_conv2d_fwd: Ho/Wo inconsistent with input size");
  cudaStream_t st = (cudaStream_t)stream;
  AMOE_REQUIRE(in_pad >= 0 && out_pad >= 0, "amoe_conv2d_fwd: negative in_pad/out_pad");
  bool tc_ok = dtype == AMOE_BF16 && tc::supported(H + 2 * in_pad, W + 2 * in_pad, Cin, Cout, stride_h, stride_w);
  if (in_pad || out_pad)
    AMOE_REQUIRE(tc_ok && impl != 1, "amoe_conv2d_fwd: padded layouts are only implemented by the tcgen05 path");
  if (impl == 2) AMOE_REQUIRE(tc_ok, "amoe_conv2d_fwd: tcgen05 path does not take this shape/dtype");
  if (tc_ok && impl != 1)
    return conv_tc_launch(ctx, x, w, scale, bias, residual, y, G, x_shared, B, H, W, Cin, Cout, KH, KW,
                          stride_h, stride_w, pad_h, pad_w, Ho, Wo, relu, in_pad, out_pad, st);
  return amoe_conv2d_simt(ctx, x, w, scale, bias, residual, y, G, x_shared, B, H, W, Cin, Cout, KH, KW,
                          stride_h, stride_w, pad_h, pad_w, Ho, Wo, relu, dtype, st);
}

}  // extern "C"
