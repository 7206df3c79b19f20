// Probe 2: K-major SWIZZLE_NONE ("interleaved") UMMA operand descriptors with OVERLAPPING rows:
//   A[m][k] = data[8*m + k]  (bf16 elements)  <=>  byte address = base + 16*m + 2*k
// expressed as 8x16B core matrices with SBO (8-row group stride) = 128 B and LBO (second 16-byte
// K chunk) = 16 B.  This is the Toeplitz/"row window" structure of a small-Cin stride-2 convolution
// read straight from the raw image row in shared memory (no im2col copy).  B: [k/8][n][8] with
// LBO = N*16 B, SBO = 128 B.  Two variants: (0) fields as above, (1) LBO/SBO swapped.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_umma_nosw tools/probe_umma_nosw.cu
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
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version 1, layout type 0 = SWIZZLE_NONE
  return d;
}

constexpr int NB = 64;        // N
constexpr int NDATA = 2048;   // flat A data elements (Toeplitz variants)
constexpr int ABYTES = 128 * 64 * 2;  // A region (canonical variants need 16 KB)

__global__ void __launch_bounds__(128, 1) probe(float* out, int variant) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const uint32_t base = (smem_u32(raw) + 127u) & ~127u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(gen);                  // flat data
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(gen + ABYTES);        // [k/8][n][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (variant < 2) {
    // canonical no-swizzle A: [k/8][m][8]  with A[m][k] = (8*m + k) % 61 (same logical matrix as the Toeplitz view)
    for (int i = tid; i < 128 * 64; i += 128) {
      int m = i / 64, k = i % 64;
      A[((k >> 3) * 128 + m) * 8 + (k & 7)] = __float2bfloat16((float)((8 * m + k) % 61));
    }
  } else {
    for (int i = tid; i < NDATA; i += 128) A[i] = __float2bfloat16((float)(i % 61));
  }
  for (int i = tid; i < 64 * NB; i += 128) {
    int k = i / NB, n = i % NB;  // identity: B[n][k] = (n == k)
    Bm[((k >> 3) * NB + n) * 8 + (k & 7)] = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t phase = 0;
  {
    if (tid == 0) {
      const uint32_t a_addr = base, b_addr = base + ABYTES;
      for (int ks = 0; ks < 4; ++ks) {  // K = 64 = 4 steps of 16
        uint64_t ad, bd;
        if (variant == 0) {         // canonical A, roles: LBO = K-chunk stride, SBO = 8-row-group stride
          ad = make_desc(a_addr + ks * 2 * 128 * 16, 128 * 16, 128);
          bd = make_desc(b_addr + ks * 2 * NB * 16, NB * 16, 128);
        } else if (variant == 1) {  // canonical A, roles swapped
          ad = make_desc(a_addr + ks * 2 * 128 * 16, 128, 128 * 16);
          bd = make_desc(b_addr + ks * 2 * NB * 16, 128, NB * 16);
        } else if (variant == 2) {  // Toeplitz A (overlapping rows), roles as variant 0
          ad = make_desc(a_addr + ks * 32, 16, 128);
          bd = make_desc(b_addr + ks * 2 * NB * 16, NB * 16, 128);
        } else {                    // Toeplitz A, roles as variant 1
          ad = make_desc(a_addr + ks * 32, 128, 16);
          bd = make_desc(b_addr + ks * 2 * NB * 16, 128, NB * 16);
        }
        umma_bf16(tmem, ad, bd, idesc, ks != 0);
      }
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    for (int c0 = 0; c0 < NB; c0 += 32) {
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
      float* o = out + ((size_t)(warp * 32 + lane)) * NB + c0;
      for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  size_t n = (size_t)128 * NB;
  float* d;
  cudaMalloc(&d, n * sizeof(float));
  cudaMemset(d, 0xff, n * sizeof(float));
  size_t smem = ABYTES + 64 * NB * 2 + 256;
  probe<<<1, 128, smem>>>(d, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d: CUDA error: %s\n", variant, cudaGetErrorString(e)); return 1; }
  float* h = (float*)malloc(n * sizeof(float));
  cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int c = 0; c < NB; ++c)
      if (h[(size_t)m * NB + c] != (float)((8 * m + c) % 61)) bad++;
  const char* names[4] = {"canonical A, LBO=K-chunk stride / SBO=8-row stride", "canonical A, roles swapped",
                          "Toeplitz A (LBO=16,SBO=128)", "Toeplitz A, roles swapped"};
  printf("variant %d (%s): %s bad=%d  D[0][0..3]=%g %g %g %g  D[1][0..1]=%g %g  D[9][0]=%g D[0][8]=%g D[0][16]=%g\n", variant,
         names[variant & 3], bad ? "MISMATCH" : "ok", bad, h[0], h[1], h[2], h[3], h[NB], h[NB + 1], h[9 * NB], h[8], h[16]);
  printf("expected D[m][c] = (8*m + c) %% 61: 0 1 2 3 | 8 9 | 11 | 8 | 16\n");
  return 0;
}
