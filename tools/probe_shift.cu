// Probe: what does tcgen05.shift.cta_group::1.down do to a TMEM accumulator on B200?
// (Which lanes / columns move, where the first and last row go, how long a batch of shifts takes, and whether a shift
// issued behind tcgen05.mma by the same thread sees the MMA's result.)  Needed for the kw-fused layer1 kernel
// (csrc/conv_flat.cu): out[P] = D0[P-1] + D1[P] wants D0 moved down by one accumulator row.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_shift tools/probe_shift.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int COLS = 64;      // columns written / read back
constexpr int NVAR = 8;

struct Variant { int lane, col, count, col_step, lane_step, lane_count; };
__constant__ Variant c_var[NVAR];

// out[v][lane][col]; cyc[v] = cycles from the first shift to the commit's arrival
__global__ void __launch_bounds__(128, 1) probe(uint32_t* out, long long* cyc) {
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = holder;
  const uint32_t my = tb + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;
  for (int v = 0; v < NVAR; ++v) {
    // fill: value = lane * 1024 + col + 1
    for (int c0 = 0; c0 < COLS; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = (uint32_t)tid * 1024u + (uint32_t)(c0 + j) + 1u;
      st16(my + (uint32_t)c0, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const Variant va = c_var[v];
    if (tid == 0) {
      const long long t0 = clock64();
      for (int l = 0; l < va.lane_count; ++l)
        for (int i = 0; i < va.count; ++i)
          shift_down(tb + ((uint32_t)(va.lane + l * va.lane_step) << 16) + (uint32_t)(va.col + i * va.col_step));
      commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      cyc[v] = clock64() - t0;
    }
    if (tid != 0) mbar_wait(smem_u32(&bar), phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < COLS; c0 += 16) {
      uint32_t r[16];
      ld16(my + (uint32_t)c0, r);
#pragma unroll
      for (int j = 0; j < 16; ++j) out[((size_t)v * 128 + tid) * COLS + c0 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

int main() {
  const Variant hv[NVAR] = {
      {0, 0, 1, 0, 0, 1},     // one shift at lane 0, column 0
      {32, 0, 1, 0, 0, 1},    // one shift at lane 32
      {0, 8, 1, 0, 0, 1},     // column 8
      {0, 4, 1, 0, 0, 1},     // column 4 (not a multiple of 8)
      {0, 0, 8, 8, 0, 1},     // 8 shifts over column groups 0, 8, .. 56 at lane 0
      {0, 0, 8, 8, 32, 4},    // ... at lanes 0, 32, 64, 96
      {0, 0, 2, 0, 0, 1},     // the same address twice
      {96, 0, 1, 0, 0, 1},    // the last lane quarter
  };
  cudaMemcpyToSymbol(c_var, hv, sizeof(hv));
  uint32_t* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, sizeof(uint32_t) * NVAR * 128 * COLS);
  cudaMalloc(&d_cyc, sizeof(long long) * NVAR);
  cudaMemset(d_out, 0, sizeof(uint32_t) * NVAR * 128 * COLS);
  probe<<<1, 128>>>(d_out, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  uint32_t* h = (uint32_t*)malloc(sizeof(uint32_t) * NVAR * 128 * COLS);
  long long hc[NVAR];
  cudaMemcpy(h, d_out, sizeof(uint32_t) * NVAR * 128 * COLS, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  for (int v = 0; v < NVAR; ++v) {
    printf("variant %d: lane %d col %d count %d (col step %d) x %d lane blocks (step %d): %lld cycles\n", v, hv[v].lane, hv[v].col,
           hv[v].count, hv[v].col_step, hv[v].lane_count, hv[v].lane_step, hc[v]);
    // classify every (lane, col): '.' unchanged, 'd' = value of lane-1 (moved down), 'D' = lane-2, 'u' = lane+1, '?' other
    int changed = 0;
    for (int l = 0; l < 128; ++l) {
      char line[COLS + 1];
      bool any = false;
      for (int c = 0; c < COLS; ++c) {
        const uint32_t val = h[((size_t)v * 128 + l) * COLS + c];
        const uint32_t col = (val - 1) & 1023u, src = (val - 1) >> 10;
        char ch = '?';
        if (col == (uint32_t)c) {
          if (src == (uint32_t)l) ch = '.';
          else if ((int)src == l - 1) ch = 'd';
          else if ((int)src == l - 2) ch = 'D';
          else if ((int)src == l + 1) ch = 'u';
        }
        if (ch != '.') { any = true; ++changed; }
        line[c] = ch;
      }
      line[COLS] = 0;
      if (any && (l % 32 < 3 || l % 32 > 29 || l % 8 == 0)) printf("  lane %3d: %s\n", l, line);
    }
    printf("  changed cells: %d\n", changed);
    // first '?' cells in detail
    int shown = 0;
    for (int l = 0; l < 128 && shown < 6; ++l)
      for (int c = 0; c < COLS && shown < 6; ++c) {
        const uint32_t val = h[((size_t)v * 128 + l) * COLS + c];
        const uint32_t col = (val - 1) & 1023u, src = (val - 1) >> 10;
        if (!(col == (uint32_t)c && ((int)src == l || (int)src == l - 1 || (int)src == l - 2 || (int)src == l + 1))) {
          printf("  other: lane %d col %d holds (lane %u, col %u) raw %u\n", l, c, src, col, val);
          ++shown;
        }
      }
  }
  return 0;
}
