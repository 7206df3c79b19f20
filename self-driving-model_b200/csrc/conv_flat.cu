// "Flat-shift" 3x3 / stride-1 / pad-1 convolution on the tensor cores with halo reuse.
//
// Activations live in a PHYSICALLY padded NHWC layout [N][H+2][W+2][C] with a zero border.  With
// P = (n*(H+2) + y)*(W+2) + x the flattened padded position, the input of filter tap (kh,kw) for
// output position P is simply position P + (kh-1)*(W+2) + (kw-1): a pure ROW OFFSET in the
// flattened [P][C] matrix.  So one M tile = 256 consecutive positions needs ONE load of
// 256 + 2*(W+2) + 2 pixel rows per 64-channel chunk, and the nine taps are nine UMMA operand
// descriptors that start at different rows of the same shared-memory buffer (a 128B-swizzled
// K-major descriptor may start at any row: the swizzle is a function of the absolute smem
// address; verified on B200 by tools/probe_umma.cu).  Compared with one TMA box per tap
// (conv_tc.cu) the L2->SMEM activation traffic drops ~5x, which is what bounds the 64- and
// 128-channel layers.  Border positions are computed too (2/(W+2) waste) and stored as zeros,
// which keeps the output's zero border intact for the next convolution.
//
// Per CTA (192 threads, persistent over the tiles of one expert group):
//   warp 0: TMA producer (activation halo buffers; weight tiles unless they are smem-resident)
//   warps 1-2: tcgen05.mma issuers, one per M-half (warp 1 also owns the TMEM allocation): per tile
//           9 taps x C/64 chunks x 4 K-steps each, two 128xN fp32 accumulators per half, double
//           buffered (4N <= 512 cols)
//   warps 3-6: epilogue: tcgen05.ld -> scale/bias (+residual) -> ReLU -> zero at borders -> bf16,
//           transposed through swizzled staging rows to coalesced 16-byte stores
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace flat {

using namespace tc;

constexpr int TILE_P = 256;
constexpr int NUM_THREADS = 352;  // warp 0 producer, warps 1-2 MMA issuers (one per M-half), warps 3-10 epilogue (four per M-half)
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;
constexpr int SMEM_BUDGET = 192 * 1024;  // activation + weight stages of the N=64 kernels (their staging rows, 32 KB, come on top)
constexpr int MAX_A_STAGES = 4;
constexpr int MAX_W_STAGES = 8;
constexpr int W_RESIDENT_MAX = 80 * 1024;

struct Params {
  int G, B, H, W, Hp, Wp, C, N;
  int chunks;            // C / 64
  int rows_pad;          // rows per activation buffer (multiple of 16, loaded as two TMA boxes)
  int a_stages, w_stages, w_resident;
  uint32_t stg_off;      // byte offset of the epilogue staging rows behind the aligned smem base
  int tiles_per_group;   // ceil(B*Hp*Wp / TILE_P)
  int group_positions;   // B*Hp*Wp
  int relu;
  int reverse;           // walk the tiles back to front (amoe_set_walk_reverse)
  int res_prefetch;      // epilogue prefetches the next tile's residual rows into L2 (AMOE_FLAT_RES_PREFETCH=0 disables)
  int64_t y_group_elems, res_group_elems;  // element distance between the expert groups of y / residual
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
};

constexpr int STG_BYTES = 8 * 32 * 64 * 2;   // N=64: 8 epilogue warps x 32 rows x 128 B
constexpr int STG_BYTES_128 = 8 * 32 * 128 * 2;   // N=128: 8 epilogue warps x 32 rows x 256 B
constexpr int SMEM_BUDGET_128 = 160 * 1024;       // operand stages of the N=128 kernel (224 KB with its staging rows)

// CTAS = 2: the CTA-pair variant.  Clusters of two CTAs take two consecutive 256-position tiles; the leader (rank 0) issues
// M = 256 MMAs (tcgen05.mma.cta_group::2) whose B operand is split over the pair - each CTA holds only N/2 rows of every
// weight tile - so an MMA step reads 4 KB + N*16 B instead of 4 KB + N*32 B of shared memory per CTA (N = 64: 5 instead of
// 6 KB per 32 tensor cycles), and the resident weights of layer1 shrink from 72 to 36 KB (one more activation stage).
template <int N, int CTAS, int ISS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_flat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_A_STAGES + 2 * MAX_W_STAGES + 4];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_stage_bytes = (uint32_t)p.rows_pad * 128u;
  constexpr uint32_t w_tile_bytes = (uint32_t)(N / CTAS) * 128u;   // this CTA's rows of a weight tile
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w = smem_base + (uint32_t)p.a_stages * a_stage_bytes;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  // tiles of this CTA: t = t_first + k * t_step for k < n_iter (the two CTAs of a pair run the same number of iterations;
  // an odd tile count leaves the last iteration of rank 1 without positions: it loads zeros and stores nothing)
  // (reverse walk: the same set of iterations, mirrored - pair p of the forward walk takes the tiles of pair n_pairs-1-p)
  const int tiles_even = CTAS == 2 ? 2 * ((p.tiles_per_group + 1) / 2) : p.tiles_per_group;
  const int t_first = p.reverse ? (CTAS == 2 ? tiles_even - 2 - (int)(blockIdx.x & ~1u) + (int)rank : tiles_even - 1 - (int)blockIdx.x)
                                : (CTAS == 2 ? (int)(blockIdx.x & ~1u) + (int)rank : (int)blockIdx.x);
  const int t_step = p.reverse ? -(int)gridDim.x : (int)gridDim.x;
  const int n_iter = CTAS == 2 ? ((p.tiles_per_group + 1) / 2 - (int)(blockIdx.x >> 1) + (int)(gridDim.x >> 1) - 1) / (int)(gridDim.x >> 1)
                               : (p.tiles_per_group - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // folded-BN scale/bias of this expert group, staged once per CTA
  __shared__ __align__(16) float s_scale[128];
  __shared__ __align__(16) float s_bias[128];
  for (int i = threadIdx.x; i < N; i += NUM_THREADS) {
    s_scale[i] = __ldg(p.scale + blockIdx.y * N + i);
    s_bias[i] = __ldg(p.bias + blockIdx.y * N + i);
  }
  const uint32_t bar_afull = smem_u32(&bars[0]);
  const uint32_t bar_aempty = smem_u32(&bars[MAX_A_STAGES]);
  const uint32_t bar_wfull = smem_u32(&bars[2 * MAX_A_STAGES]);
  const uint32_t bar_wempty = smem_u32(&bars[2 * MAX_A_STAGES + MAX_W_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * MAX_A_STAGES + 2 * MAX_W_STAGES]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * MAX_A_STAGES + 2 * MAX_W_STAGES + 2]);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int s = 0; s < MAX_A_STAGES; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, ISS);   // every MMA issuer commits
    }
    for (int s = 0; s < MAX_W_STAGES; ++s) {
      mbar_init(bar_wfull + 8 * s, 1);
      mbar_init(bar_wempty + 8 * s, ISS);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, ISS);
      mbar_init(bar_tempty + 8 * a, 8 * CTAS);   // one arrive per epilogue warp (of both CTAs: the leader's barrier is used)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair(smem_u32(&tmem_holder), TMEM_COLS);
    else tmem_alloc(smem_u32(&tmem_holder), TMEM_COLS);
  }
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();              // the peer's barriers are initialised before anything remote touches them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const int group_row0 = g * p.group_positions;  // first flattened position of this expert group
  const int wrow0 = g * N + (int)rank * (N / CTAS);   // first weight row of this expert group (of this CTA's half)
  griddep_launch_dependents();                    // PDL: the next kernel of the chain may be scheduled (see tc_common.cuh)
  // "full" barriers live in the leader: both CTAs' loads complete there (shared::cluster address of rank 0's copy)
  auto full_bar = [&](uint32_t bar) { return CTAS == 2 ? mapa_rank(bar, 0u) : bar; };
  auto load2d = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    if (CTAS == 2) tma_load_2d_pair(dst, m, bar, c0, c1);
    else tma_load_2d(dst, m, bar, c0, c1);
  };

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      if (p.w_resident) {
        // all chunks x taps weight tiles, once: tile j = chunk*9 + tap holds K columns tap*C + chunk*64 ...
        // (constant weights: loaded while the previous kernel may still be running)
        if (rank == 0) mbar_arrive_expect_tx(bar_wfull, (uint32_t)(p.chunks * 9 * CTAS) * w_tile_bytes);
        const uint32_t wf = full_bar(bar_wfull);
        for (int j = 0; j < p.chunks * 9; ++j) {
          const int chunk = j / 9, tap = j - chunk * 9;
          load2d(smem_w + (uint32_t)j * w_tile_bytes, &tmW, wf, tap * p.C + chunk * BLOCK_K, wrow0);
        }
      }
      griddep_wait();                             // activations below are the previous kernel's output
      int as = 0, ws = 0;
      uint32_t aphase = 0, wphase = 0;
      const int half_rows = p.rows_pad >> 1;
      for (int k = 0, t = t_first; k < n_iter; ++k, t += t_step) {
        const int row_start = group_row0 + t * TILE_P - p.Wp - 1;  // may be negative / past the end: TMA zero-fills
        for (int chunk = 0; chunk < p.chunks; ++chunk) {
          mbar_wait(bar_aempty + 8 * as, aphase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(bar_afull + 8 * as, (uint32_t)CTAS * a_stage_bytes);
          const uint32_t af = full_bar(bar_afull + 8 * as);
          const uint32_t dst = smem_a + (uint32_t)as * a_stage_bytes;
          load2d(dst, &tmA, af, chunk * BLOCK_K, row_start);
          load2d(dst + (uint32_t)half_rows * 128u, &tmA, af, chunk * BLOCK_K, row_start + half_rows);
          if (++as == p.a_stages) { as = 0; aphase ^= 1u; }
          if (!p.w_resident) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(bar_wempty + 8 * ws, wphase ^ 1u);
              if (rank == 0) mbar_arrive_expect_tx(bar_wfull + 8 * ws, (uint32_t)CTAS * w_tile_bytes);
              load2d(smem_w + (uint32_t)ws * w_tile_bytes, &tmW, full_bar(bar_wfull + 8 * ws), tap * p.C + chunk * BLOCK_K, wrow0);
              if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp <= 2) {
    if (rank == 0 && warp <= ISS) {
    // ============================ MMA issuers (leader CTA only) ===========
    // An N=64 MMA occupies the tensor pipe for 32 cycles, but computing its descriptors and moving
    // them to uniform registers costs one warp ~19 issue slots, so the two M-halves of a tile are
    // issued by two warps on different SM sub-partitions; each commits to the shared barriers itself.
    // (Measured on B200: what then bounds the N=64 layers is shared-memory bandwidth - every
    // 128x64x16 MMA reads 4 KB of A and 2 KB of B from smem, 432 KB per 256-position tile, ~4300
    // cycles at 128 B/clk against 2304 cycles of tensor time; row-unaligned tap shifts cost nothing.)
    // ISS = 1: one warp issues both halves (MMAs of a CTA pair arrive at the peer SM as one ordered stream)
    constexpr int NH = 3 - ISS;          // M-halves issued by this warp
    const int half0 = ISS == 2 ? warp - 1 : 0;
    const uint32_t idesc = CTAS == 2 ? make_idesc_pair(N) : make_idesc(N);
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accum) {
      if (CTAS == 2) umma_bf16_pair(d, a, b, idesc, accum);
      else umma_bf16(d, a, b, idesc, accum);
    };
    auto commit = [&](uint32_t bar) {
      if (CTAS == 2) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    int as = 0, ws = 0;
    uint32_t aphase = 0, wphase = 0;
    if (p.w_resident) {
      mbar_wait(bar_wfull, 0);
      tcgen05_fence_after();
    }
    const uint32_t wp128 = (uint32_t)p.Wp * 128u;
    for (int it = 0; it < n_iter; ++it) {
      const int acc = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      if (CTAS == 2) mbar_wait_cluster(bar_tempty + 8 * acc, tphase ^ 1u);
      else mbar_wait(bar_tempty + 8 * acc, tphase ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE + half0 * N);
      for (int chunk = 0; chunk < p.chunks; ++chunk) {
        mbar_wait(bar_afull + 8 * as, aphase);
        tcgen05_fence_after();
        const uint32_t a_half = smem_a + (uint32_t)as * a_stage_bytes + (uint32_t)half0 * (BLOCK_M * 128u);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          uint32_t w_addr;
          if (p.w_resident) {
            w_addr = smem_w + (uint32_t)(chunk * 9 + tap) * w_tile_bytes;
          } else {
            mbar_wait(bar_wfull + 8 * ws, wphase);
            tcgen05_fence_after();
            w_addr = smem_w + (uint32_t)ws * w_tile_bytes;
          }
          const uint64_t a_desc = make_sw128_desc(a_half + (uint32_t)(tap / 3) * wp128 + (uint32_t)(tap % 3) * 128u);  // row shift of this tap
          const uint64_t b_desc = make_sw128_desc(w_addr);
#pragma unroll
          for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
              mma(d_tmem + (uint32_t)(h * N), a_desc + (uint64_t)(h * (BLOCK_M * 128 / 16) + kk * 2), b_desc + (uint64_t)(kk * 2),
                  (uint32_t)((chunk | tap | kk) != 0));
          if (!p.w_resident) {
            commit(bar_wempty + 8 * ws);
            if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
          }
        }
        commit(bar_aempty + 8 * as);
        if (chunk == p.chunks - 1) commit(bar_tfull + 8 * acc);
        if (++as == p.a_stages) { as = 0; aphase ^= 1u; }
      }
    }
    }
  } else {
    // ============================ epilogue ================================
    // TMEM hands every thread one output ROW (32 rows per warp); written straight to global that is
    // 32 different 128-byte lines per store instruction, which costs more L1 cycles than the MMAs of
    // an N=64 tile take.  So each warp transposes through its own swizzled staging rows in shared
    // memory: residual rows arrive there by cp.async (issued before the accumulator is ready),
    // every thread updates its row in place, and the warp then streams the 32 x N block - which is
    // contiguous in the flattened [P][N] output - with fully coalesced 16-byte stores.
    // (Measured and removed: every thread reading its residual row straight from global memory into registers before the
    // accumulator is ready, no staging - layer1 275 -> 300 us.  The residual convolutions of layer1 move 1.25 GB of DRAM
    // in 275 us = 4.5 TB/s: they are HBM-bound, the staging is not what they wait for.)
    // Eight warps: warp (lg, hs) owns TMEM lane quarter lg of M-half hs.  (Measured with the epilogue body removed:
    // the N=128 kernel runs at 90 % tensor-busy, 156-166 us per layer2 convolution, against 70-76 % / 189-214 us with
    // four warps draining both halves one after the other - it was bound by this epilogue, not by its MMAs.)
    // (Measured and removed: lane 0 alone polling the mbarriers + __syncwarp.  ncu --set full counts ~29 M shared-memory "bank
    // conflicts" per layer1 launch - 32 lanes of a try_wait hitting one word, the LSU shared pipe 77 % busy beside the tensor
    // core's operand pipe at 82 % - but they cost nothing: epilogue warps polling with one lane is neutral (222/273 us either
    // way), and in the MMA issuer warps it is slower (layer2 pair kernel 186 -> 463 us: a lane-0 branch next to the elected
    // MMA issue takes the descriptors off the uniform datapath).)
    griddep_wait();                      // residual reads / output writes only after the previous kernel is complete
    const int lg = warp & 3, hs = (warp - 3) >> 2;
    const int img = p.Hp * p.Wp;
    constexpr int NV = N / 8;            // 16-byte pieces per output row
    constexpr int ROWB = N * 2;          // bytes per staged row
    const uint32_t stg_off_w = p.stg_off + (uint32_t)(lg * 2 + hs) * (32u * ROWB);
    uint8_t* stg_warp = smem_raw + (smem_base - smem_u32(smem_raw)) + stg_off_w;
    const uint32_t stg_warp_s = smem_base + stg_off_w;
    const bool has_res = p.residual != nullptr;
    uint8_t* my_row = stg_warp + (size_t)lane * ROWB;
    const uint32_t tempty_leader = CTAS == 2 ? mapa_rank(bar_tempty, 0u) : bar_tempty;
    for (int it = 0, t = t_first; it < n_iter; ++it, t += t_step) {
      const int acc = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      const int q0 = t * TILE_P + hs * BLOCK_M + lg * 32;  // first position of this warp inside the group
      const int q = q0 + lane;
      const int rows_valid = min(32, max(0, p.group_positions - q0));
      const int rem = q % img;
      const int yy = rem / p.Wp, xx = rem - yy * p.Wp;
      const bool interior = q < p.group_positions && yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W;
      const int64_t off0 = (int64_t)g * p.y_group_elems + (int64_t)q0 * N;      // element offset of the warp's first row in y
      const int64_t roff0 = (int64_t)g * p.res_group_elems + (int64_t)q0 * N;   // ... and in the residual
      if (has_res) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.residual + roff0);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int m = i * 32 + lane, row = m / NV, j = m - row * NV;
          if (row < rows_valid) {
            const uint32_t dst = stg_warp_s + (uint32_t)row * ROWB + (uint32_t)(((j & ~7) | ((j ^ row) & 7)) * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)m * 16) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // The residual was last touched a whole convolution ago (long evicted from L2), so the cp.async above pays an
        // HBM round trip that the ~1.4 us of MMAs per tile only partly cover: ask L2 for the NEXT tile's rows now.
        if (p.res_prefetch) {
          const int tn = t + t_step;
          if (tn >= 0 && tn < p.tiles_per_group) {
            const int q0n = tn * TILE_P + hs * BLOCK_M + lg * 32;
            const int bytes = min(32, max(0, p.group_positions - q0n)) * ROWB;
            const uint8_t* srcn = reinterpret_cast<const uint8_t*>(p.residual + (int64_t)g * p.res_group_elems + (int64_t)q0n * N);
            for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(srcn + o));
          }
        }
      }
      mbar_wait(bar_tfull + 8 * acc, tphase);
      tcgen05_fence_after();
      if (has_res) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * ACC_STRIDE + hs * N);
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 64) {
        uint32_t a[2][32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c0, a[0]);
        tmem_ld_32x32b_x32(taddr + (uint32_t)(c0 + 32), a[1]);
        tmem_ld_wait();
        if (c0 + 64 >= N) {
          // this warp's part of the accumulator is in registers now: hand it back before the math and the stores
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) mbar_arrive_cluster(tempty_leader + 8 * acc);
            else mbar_arrive(bar_tempty + 8 * acc);
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0 + h * 32);
          const float4* bs4 = reinterpret_cast<const float4*>(s_bias + c0 + h * 32);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int j = (c0 >> 3) + h * 4 + v;   // logical 16-byte piece of the row
            uint4* slot = reinterpret_cast<uint4*>(my_row + (((j & ~7) | ((j ^ lane) & 7)) * 16));
            uint4 o = make_uint4(0u, 0u, 0u, 0u);  // border positions stay zero
            if (interior) {
              // two-lane FMAs / adds and the ReLU on the packed pairs (tc_common.cuh): the bits of the scalar sequence
              const float4 s0 = sc4[2 * v], s1 = sc4[2 * v + 1], b0 = bs4[2 * v], b1 = bs4[2 * v + 1];
              float f[8];
              ffma2(f[0], f[1], a[h][v * 8 + 0], a[h][v * 8 + 1], s0.x, s0.y, b0.x, b0.y);
              ffma2(f[2], f[3], a[h][v * 8 + 2], a[h][v * 8 + 3], s0.z, s0.w, b0.z, b0.w);
              ffma2(f[4], f[5], a[h][v * 8 + 4], a[h][v * 8 + 5], s1.x, s1.y, b1.x, b1.y);
              ffma2(f[6], f[7], a[h][v * 8 + 6], a[h][v * 8 + 7], s1.z, s1.w, b1.z, b1.w);
              if (has_res) {
                const uint4 rr = *slot;
                add_bf16x2(f[0], f[1], rr.x);
                add_bf16x2(f[2], f[3], rr.y);
                add_bf16x2(f[4], f[5], rr.z);
                add_bf16x2(f[6], f[7], rr.w);
              }
              o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                             pack_bf16x2(f[6], f[7]));
              if (p.relu) o = make_uint4(relu_bf16x2(o.x), relu_bf16x2(o.y), relu_bf16x2(o.z), relu_bf16x2(o.w));
            }
            *slot = o;
          }
        }
      }
      __syncwarp();
      // coalesced write-out of the warp's 32 x N block
      uint8_t* dst = reinterpret_cast<uint8_t*>(p.y + off0);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int m = i * 32 + lane, row = m / NV, j = m - row * NV;
        if (row < rows_valid)
          *reinterpret_cast<uint4*>(dst + (size_t)m * 16) =
              *reinterpret_cast<const uint4*>(stg_warp + (size_t)row * ROWB + (((j & ~7) | ((j ^ row) & 7)) * 16));
      }
      __syncwarp();   // staging rows are overwritten by the next tile's residual
    }
  }

  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();   // nothing of the peer (multicast commits, remote arrives, its TMEM half) is in flight
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace flat

// ---------------------------------------------------------------------------------------------------------
// 64 -> 64 channels (ResNet layer1): two horizontal taps of a filter row in ONE MMA.
//
// With N = 64 every 128x64x16 MMA reads 4 KB of activations and 2 KB of weights from shared memory for 32 cycles
// of tensor time - 192 B/clk against the ~128 B/clk the operand path delivers, which is what holds the kernel above
// at 41-52 % tensor-busy.  The A operands of the taps (kh,0) and (kh,1) are the same activation window shifted by
// one position, so here the window of tap (kh,1) is multiplied by both weight tiles at once
// (B = [W(kh,0); W(kh,1)], N = 128: adjacent in the resident weight buffer) into two 64-column accumulators
//     D0[P] = sum_kh X[P + (kh-1)*Wp] * W(kh,0)^T         D1[P] = sum_kh X[P + (kh-1)*Wp] * W(kh,1)^T
// and tap (kh,2) stays an N = 64 MMA on the window shifted by +1 that accumulates straight into D1.  Then
//     out[Q] = D0[Q-1] + D1[Q]
// Operand traffic per filter row and K step: 8 KB / 64 cycles + 6 KB / 32 cycles = 146 B/clk instead of 192.
// (All three taps in one N = 192 MMA - 107 B/clk - was measured first: it needs a second shuffle and add per
// output in the epilogue, and ~500 instructions per epilogue warp and 128-row sub-tile do not fit into the
// 1152 cycles its MMAs take; it ran at 30 % tensor-busy.)  Accumulator rows are TMEM lanes, so the one-row shift
// of D0 is a warp shuffle; lane 0 of every warp takes it from the previous lane quarter's warp through a 128-byte
// shared-memory exchange, and row 0 of every 128-row MMA is halo: a sub-tile yields 127 outputs, a tile = two
// sub-tiles = 254 consecutive positions sharing one activation buffer.
//   warp 0: TMA producer   warp 1: MMA issuer (+TMEM owner)   warp 2: idle   warps 3-18: epilogue
namespace flat {

constexpr int KW3_STATIC = 3072;         // its static shared memory (row exchange) comes out of the dynamic budget
constexpr int KW3_THREADS = 11 * 32;     // warp 0 producer, warps 1-2 MMA issuers (one per sub-tile), warps 3-10 epilogue
constexpr int KW3_SUB = 127;             // outputs per 128-row sub-tile
constexpr int KW3_TILE = 2 * KW3_SUB;    // positions per tile

// n / d for 0 <= n < 2^31 with a precomputed multiplier (host: fastdiv_init)
struct FastDiv { uint32_t mul, shift, d; };
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return (__umulhi(n, f.mul) + n) >> f.shift; }
static FastDiv fastdiv_init(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = (uint32_t)((((1ull << s) - d) << 32) / d + 1);
  return f;
}

struct Kw3Params {
  Params p;
  FastDiv div_img, div_wp;
};

// (a0 + a1) * s + b on two fp32 lanes per instruction (Blackwell add/fma.rn.f32x2: per-element IEEE)
__device__ __forceinline__ void add_scale2(float& h0, float& h1, uint32_t u0, uint32_t u1, uint32_t c0, uint32_t c1,
                                           float s0, float s1, float b0, float b1) {
  asm("{\n"
      ".reg .b64 ru, rc, rs, rb, rd;\n"
      "mov.b64 ru, {%2, %3};\n"
      "mov.b64 rc, {%4, %5};\n"
      "mov.b64 rs, {%6, %7};\n"
      "mov.b64 rb, {%8, %9};\n"
      "add.rn.f32x2 rd, ru, rc;\n"
      "fma.rn.f32x2 rd, rd, rs, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(h0), "=f"(h1)
      : "r"(u0), "r"(u1), "r"(c0), "r"(c1), "f"(s0), "f"(s1), "f"(b0), "f"(b1));
}

__global__ void __launch_bounds__(KW3_THREADS, 1)
conv3x3_flat_kw3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                        const __grid_constant__ Kw3Params kp) {
  constexpr int N = 64;
  const Params& p = kp.p;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_A_STAGES + 1 + 4];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_scale[N];
  __shared__ __align__(16) float s_bias[N];
  __shared__ __align__(16) float xch[2][2][4][32];   // [toggle][channel half][lane quarter][channel]: D0 row of lane 31

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_stage_bytes = (uint32_t)p.rows_pad * 128u;
  constexpr uint32_t w_tile_bytes = (uint32_t)N * 128u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w = smem_base + (uint32_t)p.a_stages * a_stage_bytes;
  for (int i = threadIdx.x; i < N; i += KW3_THREADS) {
    s_scale[i] = __ldg(p.scale + blockIdx.y * N + i);
    s_bias[i] = __ldg(p.bias + blockIdx.y * N + i);
  }
  const uint32_t bar_afull = smem_u32(&bars[0]);
  const uint32_t bar_aempty = smem_u32(&bars[MAX_A_STAGES]);
  const uint32_t bar_wfull = smem_u32(&bars[2 * MAX_A_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * MAX_A_STAGES + 1]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * MAX_A_STAGES + 3]);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int s = 0; s < MAX_A_STAGES; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 2);    // both MMA issuers commit
    }
    mbar_init(bar_wfull, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 8);    // one arrive per epilogue warp of the sub-tile
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const int group_row0 = g * p.group_positions;
  const int wrow0 = g * N;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_wfull, 9u * w_tile_bytes);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(smem_w + (uint32_t)tap * w_tile_bytes, &tmW, bar_wfull, tap * p.C, wrow0);
      int as = 0;
      uint32_t aphase = 0;
      const int half_rows = p.rows_pad >> 1;
      for (int t = blockIdx.x; t < p.tiles_per_group; t += gridDim.x) {
        const int row_start = group_row0 + t * KW3_TILE - 1 - p.Wp;   // may be negative: TMA zero-fills
        mbar_wait(bar_aempty + 8 * as, aphase ^ 1u);
        mbar_arrive_expect_tx(bar_afull + 8 * as, a_stage_bytes);
        const uint32_t dst = smem_a + (uint32_t)as * a_stage_bytes;
        tma_load_2d(dst, &tmA, bar_afull + 8 * as, 0, row_start);
        tma_load_2d(dst + (uint32_t)half_rows * 128u, &tmA, bar_afull + 8 * as, 0, row_start + half_rows);
        if (++as == p.a_stages) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp <= 2) {
    // ============================ MMA issuers =============================
    // One warp per sub-tile (disjoint accumulators): computing the descriptors of an MMA and moving them to uniform
    // registers costs a warp ~20 dependent issue slots, more than an N=64 MMA occupies the tensor pipe.
    const int sub = warp - 1;
    const uint32_t idesc2 = make_idesc(2 * N), idesc1 = make_idesc(N);
    int as = 0, it = 0;
    uint32_t aphase = 0;
    mbar_wait(bar_wfull, 0);
    tcgen05_fence_after();
    const uint32_t wp128 = (uint32_t)p.Wp * 128u;
    const uint32_t d_tmem = tmem_base + (uint32_t)(sub * ACC_STRIDE);
    for (int t = blockIdx.x; t < p.tiles_per_group; t += gridDim.x, ++it) {
      const uint32_t tphase = (uint32_t)it & 1u;
      mbar_wait(bar_afull + 8 * as, aphase);
      mbar_wait(bar_tempty + 8 * sub, tphase ^ 1u);
      tcgen05_fence_after();
      // buffer row 0 is position (tile base - 1 - Wp); MMA row r of this sub-tile is position base + sub*127 - 1 + r
      const uint32_t a_base = smem_a + (uint32_t)as * a_stage_bytes + (uint32_t)(sub * KW3_SUB) * 128u;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const uint64_t a_desc = make_sw128_desc(a_base + (uint32_t)kh * wp128);
        const uint64_t a_desc1 = make_sw128_desc(a_base + 128u + (uint32_t)kh * wp128);
        const uint64_t b_desc01 = make_sw128_desc(smem_w + (uint32_t)(kh * 3) * w_tile_bytes);
        const uint64_t b_desc2 = make_sw128_desc(smem_w + (uint32_t)(kh * 3 + 2) * w_tile_bytes);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
          umma_bf16(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc01 + (uint64_t)(kk * 2), idesc2, (uint32_t)((kh | kk) != 0));
          umma_bf16(d_tmem + (uint32_t)N, a_desc1 + (uint64_t)(kk * 2), b_desc2 + (uint64_t)(kk * 2), idesc1, 1u);
        }
      }
      umma_commit(bar_tfull + 8 * sub);
      umma_commit(bar_aempty + 8 * as);
      if (++as == p.a_stages) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 3) {
    // ============================ epilogue ================================
    // Eight warps: warp (c, lg) owns TMEM lane quarter lg (MMA rows [32*lg, 32*lg+32)) and output channels
    // [32c, 32c+32) of both sub-tiles; two warps per SM sub-partition hide each other's latencies.  (Sixteen warps,
    // one set per sub-tile, were measured slower.)  This epilogue is what bounds the kernel, so everything that does
    // not depend on the sub-tile is a per-thread constant: a thread moves the same four 16-byte pieces (row
    // 8i + lane/4, piece lane%4) of every staged 32-row x 64-byte block, between fixed staging addresses and
    // base + i*1024 bytes in global memory.
    const int lg = warp & 3, c = (warp - 3) >> 2;
    constexpr int HROWB = 64;             // bytes of a staged half row (32 channels)
    const uint32_t stg_warp_s = smem_base + p.stg_off + (uint32_t)(c * 4 + lg) * 2 * (32u * HROWB);
    const bool has_res = p.residual != nullptr;
    const int r = lg * 32 + lane;         // MMA row of this thread
    const int rl = lane >> 2;             // row (mod 8) of the pieces this thread moves
    // staging address / global byte offset of piece i = 0 (pieces of odd row pairs are swizzled: (j ^ row/2) & 3)
    const uint32_t piece_s = stg_warp_s + (uint32_t)rl * HROWB + (uint32_t)((((lane & 3) ^ (lane >> 3)) & 3) * 16);
    const int piece_g = rl * (N * 2) + (lane & 3) * 16;
    const int i0_lo = (lg == 0 && rl == 0) ? 1 : 0;     // MMA row 0 is halo: lanes 0-3 of lane quarter 0 skip piece 0
    const uint32_t my_row_s = stg_warp_s + (uint32_t)lane * HROWB;
    const uint32_t my_sw = (uint32_t)(lane >> 1);
    const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c * 32);   // folded BatchNorm of this warp's channels
    const float4* bs4 = reinterpret_cast<const float4*>(s_bias + c * 32);
    const int up = (lane + 31) & 31;
    const uint8_t* res_base = reinterpret_cast<const uint8_t*>(p.residual) + ((int64_t)g * p.res_group_elems + c * 32) * 2 + piece_g;
    uint8_t* y_base = reinterpret_cast<uint8_t*>(p.y) + ((int64_t)g * p.y_group_elems + c * 32) * 2 + piece_g;
    // Residual half rows of the warp's block of sub-tile (tt, ss) -> staging buffer ss.  Issued one sub-tile AHEAD:
    // the epilogue, not the MMA, sets the pace of this kernel.  One cp.async group per call.
    auto fetch_residual = [&](int tt, int ss) {
      if (tt < p.tiles_per_group) {
        const int q0 = tt * KW3_TILE + ss * KW3_SUB - 1 + lg * 32;
        const int rows = p.group_positions - q0 - rl;     // piece i is inside the group iff 8i < rows
        const uint8_t* src = res_base + (int64_t)q0 * (N * 2);
        const uint32_t dst = piece_s + (uint32_t)ss * (32u * HROWB);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i >= i0_lo && 8 * i < rows)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 8 * HROWB), "l"(src + i * 8 * (N * 2)) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (has_res) fetch_residual(blockIdx.x, 0);
    int it = 0, xi = 0;
    for (int t = blockIdx.x; t < p.tiles_per_group; t += gridDim.x, ++it) {
      const uint32_t tphase = (uint32_t)it & 1u;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const uint32_t stg_sub = (uint32_t)sub * (32u * HROWB);   // staging buffer of this sub-tile
        // output position of the warp's row 0 / of this thread's row, inside the group
        const int q0 = t * KW3_TILE + sub * KW3_SUB - 1 + lg * 32;
        const int q = q0 + lane;
        const uint32_t qq = (uint32_t)max(q, 0);
        const uint32_t rem = qq - fdiv(qq, kp.div_img) * kp.div_img.d;
        const uint32_t yy = fdiv(rem, kp.div_wp), xx = rem - yy * kp.div_wp.d;
        const bool interior = r >= 1 && q < p.group_positions && yy - 1u < (uint32_t)p.H && xx - 1u < (uint32_t)p.W;
        // the other staging buffer was written out one sub-tile ago: fill it with the next sub-tile's residual
        if (has_res) fetch_residual(sub == 0 ? t : t + (int)gridDim.x, sub ^ 1);
        mbar_wait(bar_tfull + 8 * sub, tphase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(sub * ACC_STRIDE + c * 32);
        uint32_t a0[32], a1[32];
        tmem_ld_32x32b_x32(taddr, a0);
        tmem_ld_32x32b_x32(taddr + (uint32_t)N, a1);
        tmem_ld_wait();
        // hand the accumulator back as soon as it is in registers
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * sub);
        // One-row shift of D0 (out[Q] = D0[Q-1] + D1[Q]): a warp rotate, plus the row at the lower edge of the lane quarter, which
        // comes from lane 31 of the PREVIOUS lane quarter's warp through shared memory.  Lane 31 publishes its row, then every lane
        // rotates - before the exchange barrier, so the 32 shuffles are off the path behind it - and lane 0 replaces what the
        // rotate gave it (its own warp's lane 31) by the published row.  MMA row 0 is halo.
        // (A clock64 trace of this epilogue showed ~580-1100 cycles for the arithmetic of one sub-tile: the shared-memory
        // accesses were asm volatile with a memory clobber, which kept the four 8-channel groups strictly one after the other -
        // LDS -> add/FMA -> pack -> STS chains of ~140 cycles each.  Now: all loads, then the arithmetic, then all stores; the
        // shared-memory asm statements are volatile (ordered among themselves and against the barriers) without the clobber.)
        if (lane == 31) {
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<uint4*>(&xch[xi][c][lg][v * 4]) = make_uint4(a0[v * 4], a0[v * 4 + 1], a0[v * 4 + 2], a0[v * 4 + 3]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) a0[i] = __shfl_sync(0xffffffffu, a0[i], up);
        // the four lane-quarter warps of this channel half
        if (c == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const uint4 t4 = *reinterpret_cast<const uint4*>(&xch[xi][c][(lg + 3) & 3][v * 4]);
            a0[v * 4] = t4.x; a0[v * 4 + 1] = t4.y; a0[v * 4 + 2] = t4.z; a0[v * 4 + 3] = t4.w;
          }
        }
        xi ^= 1;
        uint4 rr[4];
        if (has_res) {
          asm volatile("cp.async.wait_group 1;" ::: "memory");   // this sub-tile's residual landed (the next one may be in flight)
          __syncwarp();
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t slot = my_row_s + stg_sub + (((uint32_t)v ^ my_sw) & 3u) * 16u;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rr[v].x), "=r"(rr[v].y), "=r"(rr[v].z), "=r"(rr[v].w) : "r"(slot));
          }
        }
        uint4 o[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float h[8];
          const float4 s0 = sc4[2 * v], s1 = sc4[2 * v + 1], b0 = bs4[2 * v], b1 = bs4[2 * v + 1];
          const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float bsv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int jj = 0; jj < 8; jj += 2) {
            const int i = v * 8 + jj;
            add_scale2(h[jj], h[jj + 1], a0[i], a0[i + 1], a1[i], a1[i + 1], scv[jj], scv[jj + 1], bsv[jj], bsv[jj + 1]);
          }
          if (has_res) {
            add_bf16x2(h[0], h[1], rr[v].x);
            add_bf16x2(h[2], h[3], rr[v].y);
            add_bf16x2(h[4], h[5], rr[v].z);
            add_bf16x2(h[6], h[7], rr[v].w);
          }
          o[v] = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
          if (p.relu) o[v] = make_uint4(relu_bf16x2(o[v].x), relu_bf16x2(o[v].y), relu_bf16x2(o[v].z), relu_bf16x2(o[v].w));   // max(bf16(x), 0) == bf16(max(x, 0))
          if (!interior) o[v] = make_uint4(0u, 0u, 0u, 0u);   // border / halo positions stay zero
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint32_t slot = my_row_s + stg_sub + (((uint32_t)v ^ my_sw) & 3u) * 16u;
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(slot), "r"(o[v].x), "r"(o[v].y), "r"(o[v].z), "r"(o[v].w));
        }
        __syncwarp();
        // write-out: piece i of this thread = row 8i + lane/4 of the block, 16 bytes at byte 16*(lane%4) of its 64
        {
          uint8_t* dst = y_base + (int64_t)q0 * (N * 2);
          const int rows = p.group_positions - q0 - rl;
          const uint32_t src = piece_s + stg_sub;
          uint4 v4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v4[i].x), "=r"(v4[i].y), "=r"(v4[i].z), "=r"(v4[i].w) : "r"(src + i * 8 * HROWB));
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i >= i0_lo && 8 * i < rows) *reinterpret_cast<uint4*>(dst + i * 8 * (N * 2)) = v4[i];
        }
        __syncwarp();   // staging rows of this sub-tile are rewritten by the next tile's residual fetch
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Opt-in (AMOE_FLAT_KW3=1).  Measured on B200 at batch 256 (layer1, 3 experts, four convolutions per forward): correct,
// but the whole forward is 0.08-0.19 ms slower with it (4.56-4.69 ms against 4.48-4.50 ms on the same box).  First
// version: 308 / 369 us per convolution (without / with residual) against 249 / 302 us of conv3x3_flat_kernel<64>.
// Timing the parts alone: loads + MMAs + TMEM drain 171 us - the operand-bandwidth argument above holds - but the
// epilogue alone 270 us: ~400 instructions per warp and 128-row sub-tile (8 warps -> ~3200 issue slots per sub-tile
// against ~1300 cycles of MMA time).  More epilogue warps (16, one set per sub-tile), a second MMA issuer, the
// residual prefetched one sub-tile ahead, dropping the row-exchange barrier or the shuffles: each within noise - the
// epilogue is bound by instruction issue as a whole.  Making everything sub-tile-independent a per-thread constant
// (this version, ~250 instructions) recovered two thirds of the gap.  What is left is inherent to the shift
// (32 shuffles, the exchange and its barrier, eight warps' fixed overhead); the 171 us bound stays the target.
// Round 2, session 3: a clock64 trace of one epilogue warp (phases per 127-output sub-tile, cycles): index arithmetic ~150,
// try_wait on a ready barrier ~155, TMEM load 30, exchange (lane-31 row, barrier, 32 shuffles) ~400, arithmetic ~400 when the
// warp has its scheduler to itself and ~1000 when its partner warp is in the same phase, write-out ~170-250: ~1350-2300 per
// sub-tile against 1152 cycles of MMAs.  Batching loads / arithmetic / stores (no memory clobbers between the 8-channel
// groups) took it from 267 / 315 us to 260 / 304 us; the ~380 instructions per warp and sub-tile (many of them half-rate:
// f32x2, F2FP, SHFL, LDS) are ~1400 scheduler cycles per sub-tile for the two warps of a scheduler - more than the MMAs
// take, so even a perfect schedule stays epilogue-bound at ~210 us against 228 us of the default kernel.  Not pursued further.
// Sixteen epilogue warps of 16 channels each (88 registers, two pieces per thread, bit-identical to
// the eight-warp version): 272 / 320 us against 271 / 313 us - the epilogue is not bound by latency hiding either.
// tcgen05.shift.down (tools/probe_shift.cu: moves 8 columns of ALL 128 lanes by one lane towards lane 0 inside each
// 32-lane block, lane 31 of a block keeps its value; ~70 cycles per instruction) cannot replace the shuffles: 8 shifts
// per 64-column accumulator are ~560 tensor-pipe cycles per sub-tile and the block boundary still needs the exchange.
static bool pair_enabled(int cout) {
  // default: pairs for the N = 128 kernel only.  Measured on B200 (tools/flat_bench.py, 3 x 256 frames): layer2 195 -> 180 us
  // (215 -> 198 us with residual), but layer1 (N = 64) 232 -> 303 us - a pair MMA reads its operands at ~64 B/clk per SM
  // (88 cycles per N=64 step, 98 per N=128 step), so it only pays where an MMA step is long enough to cover 4 KB of A.
  const char* e = getenv("AMOE_FLAT_PAIR");
  if (e == nullptr) return cout == 128;
  return atoi(e) != 0;
}
static bool kw3_enabled() {
  const char* e = getenv("AMOE_FLAT_KW3");
  return e != nullptr && atoi(e) != 0;
}

}  // namespace flat

namespace flat {
static bool supported(int H, int W, int C, int N) {
  return C % 64 == 0 && (N == 64 || N == 128) && W + 2 <= 120 && H >= 1;
}

}  // namespace flat

int amoe_conv_flat_init(amoe_ctx* ctx) {
  AMOE_ENTER(ctx);
  (void)ctx;
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(flat::conv3x3_flat_kernel<64, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       flat::SMEM_BUDGET + flat::STG_BYTES + 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(flat::conv3x3_flat_kernel<128, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       flat::SMEM_BUDGET_128 + flat::STG_BYTES_128 + 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(flat::conv3x3_flat_kernel<64, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       flat::SMEM_BUDGET + flat::STG_BYTES + 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(flat::conv3x3_flat_kernel<128, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       flat::SMEM_BUDGET_128 + flat::STG_BYTES_128 + 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(flat::conv3x3_flat_kw3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       flat::SMEM_BUDGET - flat::KW3_STATIC + flat::STG_BYTES + 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(flat::conv3x3_flat_kernel<64, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       flat::SMEM_BUDGET + flat::STG_BYTES + 1024));
  AMOE_CHECK_CUDA(cudaFuncSetAttribute(flat::conv3x3_flat_kernel<128, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       flat::SMEM_BUDGET_128 + flat::STG_BYTES_128 + 1024));
  return 0;
}

extern "C" {

int amoe_conv3x3_flat_supported(int H, int W, int Cin, int Cout) { return flat::supported(H, W, Cin, Cout) ? 1 : 0; }

int amoe_conv3x3_flat_fwd(amoe_ctx* ctx, const void* x, const void* w, const float* scale, const float* bias,
                          const void* residual, void* y, int G, int B, int H, int W, int Cin, int Cout, int relu,
                          void* stream) {
  AMOE_ENTER(ctx);
  return amoe_conv3x3_flat_fwd_strided(ctx, x, w, scale, bias, residual, y, G, B, H, W, Cin, Cout, relu, 0, 0, stream);
}

int amoe_conv3x3_flat_fwd_strided(amoe_ctx* ctx, const void* x, const void* w, const float* scale, const float* bias,
                                  const void* residual, void* y, int G, int B, int H, int W, int Cin, int Cout, int relu,
                                  int64_t y_group_images, int64_t res_group_images, void* stream) {
  AMOE_ENTER(ctx);
  using namespace flat;
  AMOE_REQUIRE(ctx && x && w && scale && bias && y, "amoe_conv3x3_flat_fwd: NULL argument");
  AMOE_REQUIRE(flat::supported(H, W, Cin, Cout), "amoe_conv3x3_flat_fwd: unsupported shape H=%d W=%d Cin=%d Cout=%d", H, W, Cin, Cout);
  AMOE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
               "amoe_conv3x3_flat_fwd: pointers must be 16-byte aligned");
  Params p;
  p.G = G; p.B = B; p.H = H; p.W = W; p.Hp = H + 2; p.Wp = W + 2; p.C = Cin; p.N = Cout;
  p.chunks = Cin / BLOCK_K;
  // 64 -> 64 channels: the kw-fused kernel (three horizontal taps per MMA, 252-position tiles with a one-row halo)
  const bool kw3 = Cin == 64 && Cout == 64 && kw3_enabled();
  const int tile_p = kw3 ? KW3_TILE : TILE_P;
  p.rows_pad = ((kw3 ? KW3_TILE + 2 : TILE_P + 2) + 2 * p.Wp + 15) & ~15;   // kw3: + halo row before, + 1 for the shifted tap
  const int64_t gp = (int64_t)B * p.Hp * p.Wp;
  AMOE_REQUIRE(gp * G < (1ll << 31) - 4096, "amoe_conv3x3_flat_fwd: too many positions");
  p.group_positions = (int)gp;
  p.tiles_per_group = (int)((gp + tile_p - 1) / tile_p);
  p.relu = relu;
  p.reverse = ctx->walk_reverse;
  { const char* e = getenv("AMOE_FLAT_RES_PREFETCH"); p.res_prefetch = (e == nullptr || atoi(e) != 0) ? 1 : 0; }
  p.y_group_elems = (y_group_images > 0 ? y_group_images : B) * (int64_t)p.Hp * p.Wp * Cout;
  p.res_group_elems = (res_group_images > 0 ? res_group_images : B) * (int64_t)p.Hp * p.Wp * Cout;
  p.scale = scale; p.bias = bias;
  p.residual = (const __nv_bfloat16*)residual;
  p.y = (__nv_bfloat16*)y;
  // CTA pairs (cta_group::2 MMAs, weights split over the pair): AMOE_FLAT_PAIR=0 restores one CTA per tile
  const bool pair = !kw3 && pair_enabled(Cout) && p.tiles_per_group >= 2 && ctx->sm_count / G >= 2;
  const int a_stage = p.rows_pad * 128;
  const int w_tile = (pair ? Cout / 2 : Cout) * 128;     // rows of a weight tile held by one CTA
  const int w_all = p.chunks * 9 * w_tile;
  const int budget = Cout == 128 ? SMEM_BUDGET_128 : SMEM_BUDGET;     // operand stages; the staging rows come on top
  const int stg_bytes = Cout == 128 ? STG_BYTES_128 : STG_BYTES;
  p.w_resident = (w_all <= W_RESIDENT_MAX && budget - w_all >= 2 * a_stage) ? 1 : 0;
  int w_bytes;
  if (p.w_resident) {
    p.w_stages = 0;
    w_bytes = w_all;
  } else {
    p.w_stages = std::min(MAX_W_STAGES, std::max(2, (budget - 2 * a_stage) / w_tile));
    // keep at least 2 (preferably 3) activation stages
    while (p.w_stages > 3 && budget - p.w_stages * w_tile < 3 * a_stage) --p.w_stages;
    w_bytes = p.w_stages * w_tile;
  }
  p.a_stages = std::min(MAX_A_STAGES, (budget - (kw3 ? KW3_STATIC : 0) - w_bytes) / a_stage);
  AMOE_REQUIRE(p.a_stages >= 1, "amoe_conv3x3_flat_fwd: shared memory budget exceeded");
  if (gp == 0) return 0;

  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)(gp * G)};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(p.rows_pad / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "amoe_conv3x3_flat_fwd: cuTensorMapEncodeTiled(activations) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)G * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(pair ? Cout / 2 : Cout)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AMOE_REQUIRE(r == CUDA_SUCCESS, "amoe_conv3x3_flat_fwd: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  }
  int ctas = std::max(1, std::min(p.tiles_per_group, ctx->sm_count / G));
  if (pair) ctas = std::min((ctx->sm_count / G) & ~1, 2 * ((p.tiles_per_group + 1) / 2));
  p.stg_off = (uint32_t)(p.a_stages * a_stage + w_bytes);
  const size_t smem = (size_t)p.a_stages * a_stage + w_bytes + stg_bytes + 1024;
  if (kw3) {
    AMOE_REQUIRE(p.w_resident && p.a_stages >= 1, "amoe_conv3x3_flat_fwd: kw-fused kernel needs resident weights");
    Kw3Params kp;
    kp.p = p;
    kp.div_img = fastdiv_init((uint32_t)(p.Hp * p.Wp));
    kp.div_wp = fastdiv_init((uint32_t)p.Wp);
    conv3x3_flat_kw3_kernel<<<dim3(ctas, G), KW3_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmW, kp);
  } else if (pair) {
    const char* e = getenv("AMOE_FLAT_PAIR_ISS");
    const int iss = e ? atoi(e) : 2;
    if (Cout == 64 && iss == 1)
      AMOE_CHECK_CUDA(amoe_launch_pdl_cluster(conv3x3_flat_kernel<64, 2, 1>, 2, dim3(ctas, G), dim3(NUM_THREADS), smem, (cudaStream_t)stream, tmA, tmW, p));
    else if (Cout == 64)
      AMOE_CHECK_CUDA(amoe_launch_pdl_cluster(conv3x3_flat_kernel<64, 2, 2>, 2, dim3(ctas, G), dim3(NUM_THREADS), smem, (cudaStream_t)stream, tmA, tmW, p));
    else if (iss == 1)
      AMOE_CHECK_CUDA(amoe_launch_pdl_cluster(conv3x3_flat_kernel<128, 2, 1>, 2, dim3(ctas, G), dim3(NUM_THREADS), smem, (cudaStream_t)stream, tmA, tmW, p));
    else
      AMOE_CHECK_CUDA(amoe_launch_pdl_cluster(conv3x3_flat_kernel<128, 2, 2>, 2, dim3(ctas, G), dim3(NUM_THREADS), smem, (cudaStream_t)stream, tmA, tmW, p));
  } else if (Cout == 64)
    AMOE_CHECK_CUDA(amoe_launch_pdl(conv3x3_flat_kernel<64, 1, 2>, dim3(ctas, G), dim3(NUM_THREADS), smem, (cudaStream_t)stream, tmA, tmW, p));
  else
    AMOE_CHECK_CUDA(amoe_launch_pdl(conv3x3_flat_kernel<128, 1, 2>, dim3(ctas, G), dim3(NUM_THREADS), smem, (cudaStream_t)stream, tmA, tmW, p));
  AMOE_LAUNCH_OK(ctx);
  return 0;
}

}  // extern "C"
