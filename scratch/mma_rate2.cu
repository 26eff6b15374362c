// Micro-benchmark 2: cycles per tcgen05.mma (K=16, bf16) vs M (64/128), N, and the major-ness of EACH operand separately.
// a/b layout: 0 = K-major SW32, 1 = MN-major SW32, 2 = MN-major SW128.  Data content irrelevant.
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

__global__ void __launch_bounds__(128, 1) k(int M, int N, int amode, int bmode, int iters, long long* out) {
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
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) | (amode ? (1u << 15) : 0u) | (bmode ? (1u << 16) : 0u);
    const uint32_t sa = smem_u32(smem), sb = sa + 96 * 1024;
    auto mk = [](uint32_t addr, int mode, uint32_t& lo, uint32_t& hi) {
      if (mode == 0) { lo = ((addr & 0x3FFFF) >> 4) | (1u << 16); hi = (256u >> 4) | (1u << 14) | (6u << 29); }
      else if (mode == 1) { lo = ((addr & 0x3FFFF) >> 4) | ((2048u >> 4) << 16); hi = (256u >> 4) | (1u << 14) | (6u << 29); }
      else { lo = ((addr & 0x3FFFF) >> 4) | ((8192u >> 4) << 16); hi = (1024u >> 4) | (1u << 14) | (2u << 29); }
    };
    uint32_t a_lo, a_hi, b_lo, b_hi;
    mk(sa, amode, a_lo, a_hi); mk(sb, bmode, b_lo, b_hi);
    const uint64_t bdesc = pack64(b_lo, b_hi);
    const int nacc = (4 * N <= 448) ? 4 : (2 * N <= 448 ? 2 : 1);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t d = tm + (uint32_t)((j % nacc) * N);
        if (elect_one()) mma_ss(d, pack64(a_lo + 2u * (uint32_t)j, a_hi), bdesc, idesc);
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

// M = 64 accumulator layout probe: A = K-major no-swizzle [64 rows][16 k], row r has A[r][0] = r + 1, B = [N=16][16 k] with B[n][0] = 1
// -> D[r][n] = r + 1.  Every TMEM lane dumps column 0: shows which lanes hold rows 0..63.
__global__ void __launch_bounds__(128, 1) probe64(float* out) {
  __shared__ __align__(1024) uint8_t sA[128 * 32];
  __shared__ __align__(1024) uint8_t sB[16 * 32];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 32 / 4; i += 128) ((uint32_t*)sA)[i] = 0;
  for (int i = threadIdx.x; i < 16 * 32 / 4; i += 128) ((uint32_t*)sB)[i] = 0;
  __syncthreads();
  // K-major no swizzle: core matrix = 8 rows x 16 B contiguous (128 B); k-chunk 1 at LBO; row groups at SBO = 256 (two chunks per group)
  auto bf = [](float f) { uint32_t u = __float_as_uint(f); return (uint16_t)(u >> 16); };
  if (threadIdx.x < 64) { const int r = threadIdx.x; *(uint16_t*)(sA + (r / 8) * 256 + (r % 8) * 16) = bf((float)(r + 1)); }
  if (threadIdx.x < 16) { const int n = threadIdx.x; *(uint16_t*)(sB + (n / 8) * 256 + (n % 8) * 16) = bf(1.0f); }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tbase;
  // zero the accumulator columns first through a M=128 MMA with zero accumulate? simpler: write via tcgen05.st
  {
    uint32_t z = __float_as_uint(-7.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + ((uint32_t)(warp * 32) << 16)), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((64u >> 4) << 24);
    const uint64_t ad = (uint64_t)((smem_u32(sA) & 0x3FFFF) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
    const uint64_t bd = (uint64_t)((smem_u32(sB) & 0x3FFFF) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  __syncwarp();
  { long long t1 = clock64(); while (!mbar_try(&done, 0)) { if (clock64() - t1 > 2000000000ll) __trap(); } }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tm + ((uint32_t)(warp * 32) << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[threadIdx.x] = __uint_as_float(v);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

int main() {
  long long* dout; cudaMalloc(&dout, 16);
  float* fout; cudaMalloc(&fout, 512);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* ln[3] = {"K-SW32 ", "MN-SW32", "MN-SW128"};
  const int iters = 8000;
  for (int M : {128, 64})
    for (int am = 0; am < 3; ++am)
      for (int bm = 0; bm < 3; ++bm) {
        if ((am == 2) != (bm == 2) && am + bm != 2) continue;
        for (int N : {16, 32, 64, 96, 128, 256}) {
          if (M == 64 && N % 8) continue;
          if ((am || bm) && N > 128) continue;
          k<<<148, 128, 200 * 1024>>>(M, N, am, bm, iters, dout);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("M %d a %d b %d N %d: %s\n", M, am, bm, N, cudaGetErrorString(e)); return 1; }
          long long h[2]; cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
          printf("M=%3d A %s B %s N=%3d : issue %.1f, complete %.1f cyc/MMA\n", M, ln[am], ln[bm], N, (double)h[0] / iters, (double)h[1] / iters);
        }
      }
  probe64<<<1, 128>>>(fout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe64: %s\n", cudaGetErrorString(e)); return 1; }
  float h[128]; cudaMemcpy(h, fout, 512, cudaMemcpyDeviceToHost);
  printf("M=64 accumulator rows by TMEM lane (col 0; -7 = untouched):\n");
  for (int i = 0; i < 128; ++i) printf("%g%s", h[i], (i % 32 == 31) ? "\n" : " ");
  return 0;
}
