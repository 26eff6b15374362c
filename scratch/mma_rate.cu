// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operand source (SS vs TS + tcgen05.cp)
// and layout.  One CTA per SM (grid = 148) to include chip-level effects.  Data content is irrelevant (zeros).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }

// mode 0: SS, SW32 rows (32 B).  mode 1: SS, no-swizzle chunked (LBO/SBO).  mode 2: TS (A in TMEM, copied by tcgen05.cp 128x256b per MMA)
// mode 3: TS without the copy (A resident)   mode 4: SS SW128 rows (128 B, K advance inside the row)
__global__ void __launch_bounds__(128, 1) k(int N, int mode, int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tbase;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t sa = smem_u32(smem), sb = sa + 48 * 1024;
    uint32_t a_lo, a_hi, b_lo, b_hi, a_step;
    if (mode == 1) {   // no swizzle: LBO (K chunk) = 2048+, SBO = 128
      a_lo = ((sa & 0x3FFFF) >> 4) | ((4096u >> 4) << 16); a_hi = (128u >> 4) | (1u << 14);
      b_lo = ((sb & 0x3FFFF) >> 4) | ((4096u >> 4) << 16); b_hi = a_hi; a_step = 1;
    } else if (mode == 4) {
      a_lo = ((sa & 0x3FFFF) >> 4) | (1u << 16); a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      b_lo = ((sb & 0x3FFFF) >> 4) | (1u << 16); b_hi = a_hi; a_step = 8;
    } else {
      a_lo = ((sa & 0x3FFFF) >> 4) | (1u << 16); a_hi = (256u >> 4) | (1u << 14) | (6u << 29);
      b_lo = ((sb & 0x3FFFF) >> 4) | (1u << 16); b_hi = a_hi; a_step = 2;
    }
    uint64_t bdesc = pack64(b_lo, b_hi);
    uint32_t a_tmem = tm + 384;     // columns 384.. hold the A operand in TS modes (8 columns per K=16 tile)
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t d = tm + (uint32_t)((i % nacc) * N);
      uint64_t adesc = pack64(a_lo + (uint32_t)(i & 7) * a_step, a_hi);
      if (mode == 2 || mode == 3) {
        if (mode == 2) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_tmem), "l"(adesc) : "memory");
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(1) : "memory");
      } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    long long t1 = clock64();
    while (!mbar_try(&done, 0)) { if (clock64() - t1 > 2000000000ll) __trap(); }
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
  long long* dout; cudaMalloc(&dout, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[5] = {"SS sw32", "SS noswz", "TS + cp ", "TS resid", "SS sw128"};
  const int iters = 4000;
  for (int mode = 0; mode < 5; ++mode)
    for (int N : {32, 64, 128, 256})
      for (int nacc : {1, 4}) {
        if (nacc * N > 384) continue;
        k<<<148, 128, 100 * 1024>>>(N, mode, iters, nacc, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
        long long h[2]; cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
        printf("%s N=%3d nacc=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA  (floor N/2 = %d)\n", names[mode], N, nacc, (double)h[0] / iters, (double)h[1] / iters, N / 2);
      }
  return 0;
}
