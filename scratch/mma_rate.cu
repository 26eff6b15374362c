// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) vs N, operand major-ness / swizzle, A source (SMEM vs TMEM).
// Warp-uniform issue loop (uniform datapath), 4 accumulators round-robin, grid = 148 CTAs.  Data content irrelevant.
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(1) : "memory");
}

// mode 0: SS K-major SW32   1: SS K-major no-swizzle   2: SS MN-major no-swizzle (both operands)   3: SS MN-major SW32 (both)
// mode 4: TS (A resident in TMEM), B K-major SW32      5: SS K-major SW128
__global__ void __launch_bounds__(128, 1) k(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tbase;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tbase;
  if (warp == 0) {
    const bool mn = (mode == 2 || mode == 3);
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24) | (mn ? (3u << 15) : 0u);
    const uint32_t sa = smem_u32(smem), sb = sa + 96 * 1024;
    uint32_t a_lo, a_hi, b_lo, b_hi;
    if (mode == 1) {          // K-major no swizzle: LBO = K-chunk stride, SBO = 128
      a_lo = ((sa & 0x3FFFF) >> 4) | ((4096u >> 4) << 16); a_hi = (128u >> 4) | (1u << 14); b_lo = ((sb & 0x3FFFF) >> 4) | ((4096u >> 4) << 16); b_hi = a_hi;
    } else if (mode == 2) {   // MN-major no swizzle: LBO = 128 (k groups), SBO = chunk stride 4096
      a_lo = ((sa & 0x3FFFF) >> 4) | ((128u >> 4) << 16); a_hi = (4096u >> 4) | (1u << 14); b_lo = ((sb & 0x3FFFF) >> 4) | ((128u >> 4) << 16); b_hi = a_hi;
    } else if (mode == 3) {   // MN-major SW32: LBO = 16-channel block stride 8192, SBO = 256
      a_lo = ((sa & 0x3FFFF) >> 4) | ((8192u >> 4) << 16); a_hi = (256u >> 4) | (1u << 14) | (6u << 29); b_lo = ((sb & 0x3FFFF) >> 4) | ((8192u >> 4) << 16); b_hi = a_hi;
    } else if (mode == 5) {
      a_lo = ((sa & 0x3FFFF) >> 4) | (1u << 16); a_hi = (1024u >> 4) | (1u << 14) | (2u << 29); b_lo = ((sb & 0x3FFFF) >> 4) | (1u << 16); b_hi = a_hi;
    } else {
      a_lo = ((sa & 0x3FFFF) >> 4) | (1u << 16); a_hi = (256u >> 4) | (1u << 14) | (6u << 29); b_lo = ((sb & 0x3FFFF) >> 4) | (1u << 16); b_hi = a_hi;
    }
    const uint64_t bdesc = pack64(b_lo, b_hi);
    const uint32_t a_tmem = tm + 448;
    const int nacc = (4 * N <= 448) ? 4 : (2 * N <= 448 ? 2 : 1);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t d = tm + (uint32_t)((j % nacc) * N);
        if (elect_one()) {
          if (mode == 4) mma_ts(d, a_tmem, bdesc, idesc);
          else mma_ss(d, pack64(a_lo + 2u * (uint32_t)j, a_hi), bdesc, idesc);
        }
      }
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    __syncwarp();
    long long t1 = clock64();
    while (!mbar_try(&done, 0)) { if (clock64() - t1 > 2000000000ll) __trap(); }
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
  long long* dout; cudaMalloc(&dout, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[6] = {"SS K-major SW32 ", "SS K-major noswz", "SS MN-major nosw", "SS MN-major SW32", "TS A-in-TMEM    ", "SS K-major SW128"};
  const int iters = 8000;
  for (int mode = 0; mode < 6; ++mode)
    for (int N : {16, 32, 64, 128, 256}) {
      k<<<148, 128, 200 * 1024>>>(N, mode, iters, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
      printf("%s N=%3d : issue %.1f cyc/MMA, complete %.1f cyc/MMA  (tensor floor N/2 = %d)\n", names[mode], N, (double)h[0] / iters, (double)h[1] / iters, N / 2);
    }
  return 0;
}
