// Experiment: does a K-major SWIZZLED UMMA A-operand tolerate a start address shifted by s rows (not a multiple of 8)?
// For each swizzle mode (32/64/128 B rows), shift s and base_offset policy, compare D = A[s:s+128] * B^T with the CPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) if (clock64() - t0 > 2000000000ll) __trap();
}

struct P { int rowbytes; int layout_type; int shift; int base_off_mode; int ksteps; float* out; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, P p) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full, done;
  __shared__ uint32_t tbase;
  uint8_t* sa = smem;            // 160 rows * rowbytes <= 20 KB
  uint8_t* sb = smem + 32768;    // 32 rows * rowbytes
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tbase;
  if (threadIdx.x == 0) {
    uint32_t bytes = 160 * p.rowbytes + 32 * p.rowbytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sa)), "l"(&ma), "r"(smem_u32(&full)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sb)), "l"(&mb), "r"(smem_u32(&full)), "r"(0), "r"(0) : "memory");
    mbar_wait(&full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < p.ksteps; ++ks) {
      uint32_t a_addr = smem_u32(sa) + p.shift * p.rowbytes + ks * 32;
      uint32_t b_addr = smem_u32(sb) + ks * 32;
      uint32_t sbo = 8 * p.rowbytes;
      uint64_t bo = 0;
      if (p.base_off_mode == 1) bo = (a_addr >> 7) & 7;
      uint64_t ad = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (bo << 49) | ((uint64_t)p.layout_type << 61);
      uint64_t bd = (uint64_t)((b_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)p.layout_type << 61);
      uint32_t acc = ks > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  mbar_wait(&done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  int warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) p.out[threadIdx.x * 32 + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  Enc enc; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const int Pn = 256, C = 64, N = 32;
  std::vector<__nv_bfloat16> hx(Pn * C), hb(N * C);
  std::vector<float> fx(Pn * C), fb(N * C);
  srand(1);
  for (int i = 0; i < Pn * C; ++i) { float v = (rand() % 17 - 8) / 8.f; hx[i] = __float2bfloat16(v); fx[i] = __bfloat162float(hx[i]); }
  for (int i = 0; i < N * C; ++i) { float v = (rand() % 13 - 6) / 8.f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *dx, *db; float* dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct Mode { int rowbytes; int lt; CUtensorMapSwizzle sw; } modes[3] = {{32, 6, CU_TENSOR_MAP_SWIZZLE_32B}, {64, 4, CU_TENSOR_MAP_SWIZZLE_64B}, {128, 2, CU_TENSOR_MAP_SWIZZLE_128B}};
  for (auto& m : modes) {
    int kc = m.rowbytes / 2;   // channels per row
    CUtensorMap ma, mb;
    cuuint64_t da[2] = {(cuuint64_t)C, (cuuint64_t)Pn}, sa[1] = {(cuuint64_t)C * 2};
    cuuint32_t ba[2] = {(cuuint32_t)kc, 160}, es[2] = {1, 1};
    CUresult r1 = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, da, sa, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, m.sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dbb[2] = {(cuuint64_t)C, (cuuint64_t)N};
    cuuint32_t bb[2] = {(cuuint32_t)kc, 32};
    CUresult r2 = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dbb, sa, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, m.sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2) { printf("encode failed %d %d\n", r1, r2); return 1; }
    for (int bom = 0; bom < 2; ++bom) {
      printf("rowbytes %3d base_offset_mode %d :", m.rowbytes, bom);
      for (int s = 0; s <= 17; ++s) {
        P p{m.rowbytes, m.lt, s, bom, kc / 16, dout};
        cudaMemset(dout, 0, 128 * 32 * 4);
        k<<<1, 128, 64 * 1024>>>(ma, mb, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" ERR %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<float> ho(128 * 32);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int mm = 0; mm < 128; ++mm) for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int c = 0; c < kc; ++c) ref += (double)fx[(s + mm) * C + c] * fb[n * C + c];
          maxerr = fmax(maxerr, fabs(ref - ho[mm * 32 + n]));
        }
        printf(" s%d:%s", s, maxerr < 1e-3 ? "ok" : "BAD");
      }
      printf("\n");
    }
  }
  return 0;
}
