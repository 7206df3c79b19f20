// Probe: can a 128B-swizzled K-major UMMA operand descriptor start at an arbitrary ROW of a
// larger smem buffer (start address = base + s*128 B), and which base_offset encoding does
// that need?  Needed for the "flat-shift" halo reuse of the 3x3 convolutions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_umma tools/probe_umma.cu
#include <cuda_bf16.h>
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
__device__ __forceinline__ void umma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// variant 0: base_offset = 0 ; variant 1: base_offset = (addr >> 7) & 7
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int variant) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (variant == 1) d |= (uint64_t)((addr >> 7) & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int ROWS = 256, NSHIFT = 20;

__global__ void __launch_bounds__(128, 1) probe(float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(gen);               // ROWS x 64, SW128
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(gen + ROWS * 128);  // 64 x 64 identity, SW128
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < ROWS * 64; i += 128) {
    int r = i / 64, c = i % 64;
    float v = (c & 1) ? (float)c : (float)r;
    int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16(v);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    int r = i / 64, c = i % 64;
    int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(gen + ROWS * 128 + off) = __float2bfloat16(r == c ? 1.f : 0.f);
  }
  (void)A; (void)Bm;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = holder;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int s = 0; s < NSHIFT; ++s) {
      if (tid == 0) {
        uint32_t a_addr = base + s * 128;
        uint64_t bd = make_desc(base + ROWS * 128, 0);
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t ad = make_desc(a_addr + kk * 32, variant);
          umma_bf16(tmem, ad, bd + (uint64_t)(kk * 2), idesc, kk != 0);
        }
        umma_commit(smem_u32(&bar));
      }
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* o = out + (((size_t)variant * NSHIFT + s) * 128 + warp * 32 + lane) * 64 + c0;
        for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  size_t n = (size_t)2 * NSHIFT * 128 * 64;
  float* d;
  cudaMalloc(&d, n * sizeof(float));
  cudaMemset(d, 0xff, n * sizeof(float));
  size_t smem = ROWS * 128 + 64 * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  float* h = (float*)malloc(n * sizeof(float));
  cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost);
  for (int v = 0; v < 2; ++v)
    for (int s = 0; s < NSHIFT; ++s) {
      int bad_row = 0, bad_col = 0, first_bad = -1;
      for (int m = 0; m < 128; ++m)
        for (int c = 0; c < 64; ++c) {
          float got = h[(((size_t)v * NSHIFT + s) * 128 + m) * 64 + c];
          float want = (c & 1) ? (float)c : (float)(s + m);
          if (got != want) { if (c & 1) bad_col++; else bad_row++; if (first_bad < 0) first_bad = m * 64 + c; }
        }
      float g0 = h[(((size_t)v * NSHIFT + s) * 128 + 0) * 64 + 0], g1 = h[(((size_t)v * NSHIFT + s) * 128 + 1) * 64 + 0];
      float g8 = h[(((size_t)v * NSHIFT + s) * 128 + 8) * 64 + 0], c1 = h[(((size_t)v * NSHIFT + s) * 128 + 0) * 64 + 1];
      printf("variant %d shift %2d: %s  bad_row=%d bad_col=%d  D[0][0]=%g D[1][0]=%g D[8][0]=%g D[0][1]=%g\n", v, s,
             (bad_row + bad_col) ? "MISMATCH" : "ok", bad_row, bad_col, g0, g1, g8, c1);
    }
  return 0;
}
