// First layer of the U-Net (Cin = 1: the CT intensity; reference models/networks/UNet.py:153 with in_channels = 1) on tcgen05.
// The layer is HBM-bound (AI ~ 25 FLOP/B: 2 B read, 2*Cout B written per voxel), so the job of these kernels is to keep the
// arithmetic off the FFMA / LSU pipes: the im2col row of a voxel (27 taps, zero-padded to K = 32) is 64 bytes, a tile of 128
// consecutive voxels of one image row is an 8 KB SWIZZLE_64B matrix that the threads build in shared memory from the 9 (or 3)
// input rows it touches, and ONE matrix serves both directions:
//   forward : A = tile as a K-major operand   (M = 128 voxels, K = 32 taps), B = weights [Cout][32 taps]      -> D[voxel][co]
//   wgrad   : A = tile as an MN-major operand (M = 32 taps (+ 96 ignored rows), K = 128 voxels), B = dy rows  -> D[tap][co]
// The forward epilogue (TMEM lane = voxel) adds the optional bias / ReLU, stores bf16 rows (32 contiguous bytes per voxel, a warp
// writes 1 KB) and accumulates the BatchNorm batch statistics of the stored values in registers for the CTA's lifetime.
// The weight-gradient kernel accumulates over all tiles of a CTA in TMEM (double-buffered shared-memory tiles) and adds its
// partial [taps][Cout] block to dW with fp32 atomics once.
#include "tc_common.cuh"

namespace {

struct C1Params {
  int N, D, H, W, Cout, relu;
  int x_ld, y_ld;
  int y32;                // forward: output rows are 32-byte aligned (256-bit stores)
  int nseg;               // 128-voxel segments per image row
  long long n_tiles;      // N * D * H * nseg
};

constexpr int C1_THREADS = 128;
constexpr int XS_PITCH = 144;     // xs[rr][8 + (w - w0)], w - w0 in [-8, 136): 18 aligned 16-byte chunks per input row

// SWIZZLE_64B, K-major: rows of 64 B (32 bf16), 8-row groups SBO = 512 B apart; LBO unused (1).
__device__ __forceinline__ uint64_t desc_sw64_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// MN-major swizzled operand: a K index (voxel) is one row of ROWB bytes (ROWB = 64: SWIZZLE_64B, 32: SWIZZLE_32B), groups of 8
// K rows are SBO = 8*ROWB apart, MN blocks of ROWB/2 elements are LBO apart.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t rowb, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)((8u * rowb) >> 4) << 32) | (1ull << 46) |
         ((rowb == 64 ? 4ull : 6ull) << 61);
}

// These kernels are instruction-issue bound (ncu: the first version spent 625 instructions per voxel-thread per tile, mostly
// 64-bit index arithmetic and scalar loads), so a CTA walks a CONTIGUOUS range of tiles and advances the coordinates with carries.
struct TileCoord { uint32_t row, nd; int d, h, w0; };     // row = (n*D + d)*H + h, nd = n*D + d
__device__ __forceinline__ TileCoord tile_coords(const C1Params& p, uint32_t tile) {
  TileCoord c;
  const uint32_t nseg = (uint32_t)p.nseg, H = (uint32_t)p.H, D = (uint32_t)p.D;
  c.row = tile / nseg;
  c.w0 = (int)(tile - c.row * nseg) * 128;
  c.nd = c.row / H;
  c.h = (int)(c.row - c.nd * H);
  c.d = (int)(c.nd % D);
  return c;
}
__device__ __forceinline__ void tile_next(const C1Params& p, TileCoord& c) {
  c.w0 += 128;
  if (c.w0 >= p.W) {
    c.w0 = 0; ++c.row;
    if (++c.h == p.H) { c.h = 0; ++c.nd; if (++c.d == p.D) c.d = 0; }
  }
}

// The KD*3 input rows of a tile (zero padded): xs[rr][8 + i] = x[n, d + kd - KD/2, h + kh - 1, w0 + i], i in [-1, 128], as bf16
// bit patterns.  Split in two so that the global loads of a LATER tile are in flight while the current one is built / multiplied:
// load_rows (global -> registers), store_rows (registers -> shared memory).
//   VEC  (x_ld == 1, W % 8 == 0, 16-byte aligned base): the rows are fetched as aligned 16-byte chunks, <= 2 per thread;
//   !VEC (anything else): one element per row per thread + one halo element.
template <int KD, bool VEC>
struct RowRegs;
template <int KD>
struct RowRegs<KD, true> { uint4 c[2]; };
template <int KD>
struct RowRegs<KD, false> { uint16_t v[KD * 3]; uint16_t halo; };

template <int KD>
__device__ __forceinline__ void load_rows(RowRegs<KD, true>& r, const bf16* __restrict__ x, const C1Params& p, const TileCoord& tc) {
  constexpr int CH = KD * 3 * 18;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (128 * k < CH) {
      const int cidx = (int)threadIdx.x + 128 * k;
      const int rr = cidx / 18, j = cidx - rr * 18;
      const int kd = rr / 3, kh = rr - kd * 3;
      const int dd = tc.d + kd - KD / 2, hh = tc.h + kh - 1, ww = tc.w0 - 8 + 8 * j;
      const bool ok = cidx < CH && (unsigned)dd < (unsigned)p.D && (unsigned)hh < (unsigned)p.H && (unsigned)ww < (unsigned)p.W;
      const long long off = ((long long)(tc.nd + kd - KD / 2) * p.H + hh) * p.W + ww;
      r.c[k] = ok ? *reinterpret_cast<const uint4*>(x + off) : make_uint4(0, 0, 0, 0);
    }
  }
}
template <int KD>
__device__ __forceinline__ void store_rows(uint16_t (*xs)[XS_PITCH], const RowRegs<KD, true>& r) {
  constexpr int CH = KD * 3 * 18;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (128 * k < CH) {
      const int cidx = (int)threadIdx.x + 128 * k;
      const int rr = cidx / 18, j = cidx - rr * 18;
      if (cidx < CH) *reinterpret_cast<uint4*>(&xs[rr][8 * j]) = r.c[k];
    }
  }
}

template <int KD>
__device__ __forceinline__ void load_rows(RowRegs<KD, false>& r, const bf16* __restrict__ x, const C1Params& p, const TileCoord& tc) {
  constexpr int ROWS = KD * 3;
  const int tid = threadIdx.x;
  const int hr = tid >> 1, hside = tid & 1;              // halo element of row hr: left (w0 - 1) or right (w0 + 128)
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int kd = rr / 3, kh = rr - kd * 3;
    const int dd = tc.d + kd - KD / 2, hh = tc.h + kh - 1, ww = tc.w0 + tid;
    const bool row_ok = (unsigned)dd < (unsigned)p.D && (unsigned)hh < (unsigned)p.H;
    const bf16* src = x + (((long long)(tc.nd + kd - KD / 2) * p.H + hh) * (long long)p.W) * p.x_ld;
    r.v[rr] = (row_ok && ww < p.W) ? __bfloat16_as_ushort(src[(long long)ww * p.x_ld]) : (uint16_t)0;
    if (rr == hr) {
      const int wh = hside ? tc.w0 + 128 : tc.w0 - 1;
      r.halo = (row_ok && (unsigned)wh < (unsigned)p.W) ? __bfloat16_as_ushort(src[(long long)wh * p.x_ld]) : (uint16_t)0;
    }
  }
}
template <int KD>
__device__ __forceinline__ void store_rows(uint16_t (*xs)[XS_PITCH], const RowRegs<KD, false>& r) {
  constexpr int ROWS = KD * 3;
  const int tid = threadIdx.x;
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) xs[rr][8 + tid] = r.v[rr];
  if (tid < 2 * ROWS) xs[tid >> 1][(tid & 1) ? 136 : 7] = r.halo;
}

// Thread `pos` writes row `pos` of the im2col tile: 32 bf16 = taps (rr, kw) in order, zero padded; 16-byte chunk c of row r
// lives at r*64 + ((c ^ ((r >> 1) & 3)) << 4)  (Swizzle<2,4,3> on a 1024-byte aligned tile).
template <int KD>
__device__ __forceinline__ void build_row(uint8_t* tile, const uint16_t (*xs)[XS_PITCH], int pos) {
  constexpr int ROWS = KD * 3, TAPS = ROWS * 3;
  uint32_t e[TAPS + 1];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) e[rr * 3 + kw] = xs[rr][7 + pos + kw];
  e[TAPS] = 0u;
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = (2 * i < TAPS) ? __byte_perm(e[2 * i], e[2 * i + 1], 0x5410) : 0u;
  const uint32_t sw = (uint32_t)(pos >> 1) & 3u;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + pos * 64 + (((uint32_t)c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------- forward
// Per CTA: tile i+1 is built (shared-memory buffer (i+1)&1) while the MMA of tile i runs; the epilogue of tile i (TMEM buffer i&1)
// runs while the MMA of tile i+1 does.  Two __syncthreads per tile; only warp 0 polls the MMA barrier.
template <int KD, int NPAD, bool STATS, bool VEC>
__global__ void __launch_bounds__(C1_THREADS, NPAD == 16 ? 6 : 4)
conv_cin1_tc_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ wp /*[taps][Cout] fp32*/, const float* __restrict__ bias,
                        bf16* __restrict__ y, double* __restrict__ stat_sum, double* __restrict__ stat_sumsq, const C1Params p) {
  constexpr int ROWS = KD * 3, TAPS = ROWS * 3;
  __shared__ __align__(1024) uint8_t sA[2][128 * 64];
  __shared__ __align__(1024) uint8_t sB[32 * 64];
  __shared__ __align__(16) uint16_t xs[ROWS][XS_PITCH];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red[STATS ? 4 * 2 * NPAD : 1];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_smem, 2 * NPAD);
  // weights: row co = 32 taps (K-major, SWIZZLE_64B), zero beyond Cout / TAPS
  for (int i = tid; i < 32 * 32; i += C1_THREADS) {
    const int co = i >> 5, t = i & 31;
    const float v = (co < p.Cout && t < TAPS) ? wp[t * p.Cout + co] : 0.f;
    const uint32_t off = (uint32_t)co * 64u + ((((uint32_t)t >> 3) ^ (((uint32_t)co >> 1) & 3u)) << 4) + ((uint32_t)t & 7u) * 2u;
    *reinterpret_cast<bf16*>(sB + off) = __float2bfloat16_rn(v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t bdesc = desc_sw64_kmajor(smem_u32(sB));

  float csum[STATS ? NPAD : 1], csq[STATS ? NPAD : 1];
  if (STATS) {
#pragma unroll
    for (int k = 0; k < NPAD; ++k) csum[k] = csq[k] = 0.f;
  }
  const uint32_t n_tiles = (uint32_t)p.n_tiles;
  const uint32_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t t_begin = blockIdx.x * per, t_end = t_begin + per < n_tiles ? t_begin + per : n_tiles;
  const int count = t_begin < t_end ? (int)(t_end - t_begin) : 0;

  auto issue = [&](int i) {     // MMA of tile i: shared-memory buffer i&1 -> TMEM buffer i&1 (warp 0, converged)
    tc_fence_after();
    if (elect_one()) {
      const uint64_t adesc = desc_sw64_kmajor(smem_u32(sA[i & 1]));
      const uint32_t dcol = tmem_base + (uint32_t)((i & 1) * NPAD);
      umma_bf16(dcol, adesc, bdesc, idesc, 0u);
      umma_bf16(dcol, adesc + 2, bdesc + 2, idesc, 1u);      // taps 16..31: +32 bytes inside the swizzled rows
      umma_commit(&bar[i & 1]);
    }
    __syncwarp();
  };

  RowRegs<KD, VEC> regs;
  TileCoord tl{}, te{};      // tl: the tile whose rows are loaded next; te: the tile whose epilogue runs next
  if (count > 0) {
    tl = tile_coords(p, t_begin);
    te = tl;
    load_rows<KD>(regs, x, p, tl);
    store_rows<KD>(xs, regs);
    __syncthreads();
    if (count > 1) { tile_next(p, tl); load_rows<KD>(regs, x, p, tl); }
    build_row<KD>(sA[0], xs, tid);
    fence_proxy_async();        // generic-proxy writes of the tile -> visible to the tensor core (async proxy)
    __syncthreads();
    if (warp == 0) issue(0);
  }
  for (int i = 0; i < count; ++i) {
    const bool has_next = i + 1 < count;
    if (has_next) {
      store_rows<KD>(xs, regs);                       // rows of tile i+1 (build(i) finished reading xs before the last barrier)
      __syncthreads();
      if (i + 2 < count) { tile_next(p, tl); load_rows<KD>(regs, x, p, tl); }     // in flight during build + epilogue
      build_row<KD>(sA[(i + 1) & 1], xs, tid);        // MMA(i-1), the last reader of this buffer, retired an iteration ago
      fence_proxy_async();
    }
    if (warp == 0) mbar_wait(&bar[i & 1], (uint32_t)(i >> 1) & 1u);     // MMA(i) retired
    tc_fence_before();          // epilogue(i-1)'s tcgen05.ld are ordered before MMA(i+1), which overwrites that accumulator
    __syncthreads();
    if (has_next && warp == 0) issue(i + 1);
    tc_fence_after();
    // ---- epilogue of tile i (TMEM lane = voxel)
    const int w = te.w0 + tid;
    const bool valid = w < p.W;
    bf16* yrow = y + ((long long)te.row * p.W + w) * (long long)p.y_ld;
#pragma unroll
    for (int c0 = 0; c0 < NPAD; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((i & 1) * NPAD + c0), v);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a = __uint_as_float(v[2 * k]), b = __uint_as_float(v[2 * k + 1]);
        if (!STATS) {
          if (bias) { a += (c0 + 2 * k < p.Cout) ? bias[c0 + 2 * k] : 0.f; b += (c0 + 2 * k + 1 < p.Cout) ? bias[c0 + 2 * k + 1] : 0.f; }
          if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
        }
        pk[k] = pack_bf16x2(a, b);
      }
      if (valid) {
        if (STATS) {            // statistics of the values as stored (bf16-rounded), like the other conv epilogues
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float lo = __uint_as_float(pk[k] << 16), hi = __uint_as_float(pk[k] & 0xffff0000u);
            csum[c0 + 2 * k] += lo; csq[c0 + 2 * k] = fmaf(lo, lo, csq[c0 + 2 * k]);
            csum[c0 + 2 * k + 1] += hi; csq[c0 + 2 * k + 1] = fmaf(hi, hi, csq[c0 + 2 * k + 1]);
          }
        }
        if (c0 + 16 <= p.Cout) st_global_32B(yrow + c0, pk, p.y32 != 0);
        else if (c0 + 8 <= p.Cout) *reinterpret_cast<uint4*>(yrow + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    tile_next(p, te);
  }
  if (STATS) {   // once per CTA: warp tree, cross-warp sum in shared memory, one fp64 atomic per channel
#pragma unroll
    for (int k = 0; k < NPAD; ++k) {
      const float a = warp_sum(csum[k]), b = warp_sum(csq[k]);
      if (lane == 0) { red[(warp * 2) * NPAD + k] = a; red[(warp * 2 + 1) * NPAD + k] = b; }
    }
    __syncthreads();
    if (tid < 2 * NPAD) {
      const int which = tid / NPAD, k = tid % NPAD;
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < 4; ++wq) t += red[(wq * 2 + which) * NPAD + k];
      if (k < p.Cout) atomicAdd(which ? &stat_sumsq[k] : &stat_sum[k], (double)t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * NPAD); }
}

// ---------------------------------------------------------------------------------------------------------- weight gradient
template <int KD, int NPAD, bool VEC>
__global__ void __launch_bounds__(C1_THREADS, NPAD == 16 ? 6 : 4)
conv_cin1_tc_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw /*[Cout][taps]*/, const C1Params p) {
  constexpr int ROWS = KD * 3, TAPS = ROWS * 3;
  constexpr int ROWB = NPAD * 2;                       // bytes per dy row in shared memory (32: SWIZZLE_32B, 64: SWIZZLE_64B)
  __shared__ __align__(1024) uint8_t sT[2][128 * 64];  // im2col tiles (A, MN-major: M = taps)
  __shared__ __align__(1024) uint8_t sG[2][128 * ROWB];// dy tiles (B, MN-major: N = co); also the slack the 96 ignored A rows read into
  __shared__ __align__(16) uint16_t xs[ROWS][XS_PITCH];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_smem, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // both operands MN-major (bits 15, 16); M = 128 (rows 32.. are the next 8-voxel groups re-read as "taps": ignored), N = NPAD
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);

  // this thread's dy row of a tile (16 / 32 channels)
  auto load_dy = [&](uint4* g, const TileCoord& tc) {
    const int w = tc.w0 + tid;
    const bf16* src = dy + ((long long)tc.row * p.W + w) * (long long)p.y_ld;
#pragma unroll
    for (int c = 0; c < NPAD / 8; ++c) g[c] = (w < p.W && c * 8 < p.Cout) ? *reinterpret_cast<const uint4*>(src + c * 8) : make_uint4(0, 0, 0, 0);
  };
  const uint32_t n_tiles = (uint32_t)p.n_tiles;
  const uint32_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t t_begin = blockIdx.x * per, t_end = t_begin + per < n_tiles ? t_begin + per : n_tiles;
  const int count = t_begin < t_end ? (int)(t_end - t_begin) : 0;

  RowRegs<KD, VEC> regs;
  uint4 g[NPAD / 8];
  TileCoord tl{};
  if (count > 0) {
    tl = tile_coords(p, t_begin);
    load_dy(g, tl);
    load_rows<KD>(regs, x, p, tl);
  }
  for (int it = 0; it < count; ++it) {
    const int b = it & 1;
    store_rows<KD>(xs, regs);
    if (warp == 0 && it >= 2) mbar_wait(&bar[b], (uint32_t)((it >> 1) - 1) & 1u);    // the MMAs that read buffer b two tiles ago have retired
    __syncthreads();
    {
      const uint32_t sw = ROWB == 64 ? ((uint32_t)(tid >> 1) & 3u) : ((uint32_t)(tid >> 2) & 1u);
#pragma unroll
      for (int c = 0; c < NPAD / 8; ++c) *reinterpret_cast<uint4*>(sG[b] + tid * ROWB + (((uint32_t)c ^ sw) << 4)) = g[c];
    }
    if (it + 1 < count) {     // next tile: loads in flight while this one is built and multiplied
      tile_next(p, tl);
      load_dy(g, tl);
      load_rows<KD>(regs, x, p, tl);
    }
    build_row<KD>(sT[b], xs, tid);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = desc_mnmajor(smem_u32(sT[b]), 64, 512), bdesc = desc_mnmajor(smem_u32(sG[b]), ROWB, 512);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)      // 16 voxels per MMA = two 8-row groups
          umma_bf16(tmem_base, adesc + (uint64_t)((ks * 16 * 64) >> 4), bdesc + (uint64_t)((ks * 16 * ROWB) >> 4), idesc, (it | ks) ? 1u : 0u);
        umma_commit(&bar[b]);
      }
      __syncwarp();
    }
  }
  // drain: the last commit covers every MMA issued before it
  if (count >= 1) { const int last = count - 1; mbar_wait(&bar[last & 1], (uint32_t)(last >> 1) & 1u); }
  tc_fence_after();
  if (warp == 0 && count > 0) {       // TMEM lanes 0..31 = taps
#pragma unroll
    for (int c0 = 0; c0 < NPAD; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + (uint32_t)c0, v);
      tmem_ld_wait();
      if (lane < TAPS) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (c0 + k < p.Cout) atomicAdd(&dw[(c0 + k) * TAPS + lane], __uint_as_float(v[k]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

bool c1_shape_ok(int N, int D, int H, int W, int Cout, int KD) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return false;
  if ((long long)N * D * H * ((W + 127) / 128) >= (1ll << 31)) return false;     // 32-bit tile coordinates in the kernels
  return (KD == 1 || KD == 3) && Cout >= 8 && Cout <= 32 && Cout % 8 == 0;
}

C1Params c1_params(int N, int D, int H, int W, int Cout, int x_ld, int y_ld, int relu) {
  C1Params p{};
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu; p.x_ld = x_ld; p.y_ld = y_ld;
  p.nseg = (W + 127) / 128;
  p.n_tiles = (long long)N * D * H * p.nseg;
  return p;
}

bool c1_vec_ok(const void* x, int x_ld, int W) { return x_ld == 1 && W % 8 == 0 && ((uintptr_t)x & 15) == 0; }

}  // namespace

extern "C" {

int ich_conv_cin1_tc_supported(int N, int D, int H, int W, int Cout, int KD) { return c1_shape_ok(N, D, H, W, Cout, KD) ? 1 : 0; }

int ich_conv_cin1_tc_fwd(const void* x, int x_ld, const float* wpack, const float* bias, void* y, int y_ld, double* sum, double* sumsq, int N,
                         int D, int H, int W, int Cout, int KD, int relu, void* stream) {
  ICH_REQUIRE(c1_shape_ok(N, D, H, W, Cout, KD), "ich_conv_cin1_tc_fwd: unsupported shape N%d D%d H%d W%d Cout%d KD%d", N, D, H, W, Cout, KD);
  ICH_REQUIRE(y_ld % 8 == 0 && ((uintptr_t)y & 15) == 0, "ich_conv_cin1_tc_fwd: output rows must be 16-byte aligned (y_ld %d)", y_ld);
  ICH_REQUIRE((sum == nullptr) == (sumsq == nullptr), "ich_conv_cin1_tc_fwd: fused statistics need both buffers");
  ICH_REQUIRE(!sum || (!bias && !relu), "ich_conv_cin1_tc_fwd: the fused-statistics form takes no bias / ReLU (BatchNorm follows)");
  cudaStream_t s = (cudaStream_t)stream;
  C1Params p = c1_params(N, D, H, W, Cout, x_ld, y_ld, relu);
  p.y32 = (((uintptr_t)y & 31) == 0 && y_ld % 16 == 0) ? 1 : 0;
  if (sum) {
    cudaMemsetAsync(sum, 0, sizeof(double) * Cout, s);
    cudaMemsetAsync(sumsq, 0, sizeof(double) * Cout, s);
  }
  const int npad = Cout <= 16 ? 16 : 32;
  long long grid = (long long)ich_num_sms() * (npad == 16 ? 6 : 4);
  if (grid > p.n_tiles) grid = p.n_tiles;
  const bf16* xp = (const bf16*)x;
  bf16* yp = (bf16*)y;
  const bool vec = c1_vec_ok(x, x_ld, W);
#define ICH_C1F2(K, NP, V)                                                                                                            \
  do {                                                                                                                                \
    if (sum) conv_cin1_tc_fwd_kernel<K, NP, true, V><<<(unsigned)grid, C1_THREADS, 0, s>>>(xp, wpack, bias, yp, sum, sumsq, p);        \
    else conv_cin1_tc_fwd_kernel<K, NP, false, V><<<(unsigned)grid, C1_THREADS, 0, s>>>(xp, wpack, bias, yp, nullptr, nullptr, p);     \
  } while (0)
#define ICH_C1F(K, NP) do { if (vec) ICH_C1F2(K, NP, true); else ICH_C1F2(K, NP, false); } while (0)
  if (KD == 3) { if (npad == 16) ICH_C1F(3, 16); else ICH_C1F(3, 32); }
  else { if (npad == 16) ICH_C1F(1, 16); else ICH_C1F(1, 32); }
#undef ICH_C1F
#undef ICH_C1F2
  return ich_check_launch("ich_conv_cin1_tc_fwd");
}

int ich_conv_cin1_tc_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, float* dw, int N, int D, int H, int W, int Cout, int KD,
                           void* stream) {
  ICH_REQUIRE(c1_shape_ok(N, D, H, W, Cout, KD), "ich_conv_cin1_tc_wgrad: unsupported shape N%d D%d H%d W%d Cout%d KD%d", N, D, H, W, Cout, KD);
  ICH_REQUIRE(dy_ld % 8 == 0 && ((uintptr_t)dy & 15) == 0, "ich_conv_cin1_tc_wgrad: gradient rows must be 16-byte aligned (dy_ld %d)", dy_ld);
  cudaStream_t s = (cudaStream_t)stream;
  const C1Params p = c1_params(N, D, H, W, Cout, x_ld, dy_ld, 0);
  if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * KD * 9, s) != cudaSuccess) return ich_check_launch("ich_conv_cin1_tc_wgrad memset");
  const int npad = Cout <= 16 ? 16 : 32;
  long long grid = (long long)ich_num_sms() * (npad == 16 ? 6 : 4);
  if (grid > p.n_tiles) grid = p.n_tiles;
  const bf16* xp = (const bf16*)x;
  const bf16* gp = (const bf16*)dy;
  const bool vec = c1_vec_ok(x, x_ld, W);
#define ICH_C1W(K, NP) do { if (vec) conv_cin1_tc_wgrad_kernel<K, NP, true><<<(unsigned)grid, C1_THREADS, 0, s>>>(xp, gp, dw, p); \
                            else conv_cin1_tc_wgrad_kernel<K, NP, false><<<(unsigned)grid, C1_THREADS, 0, s>>>(xp, gp, dw, p); } while (0)
  if (KD == 3) { if (npad == 16) ICH_C1W(3, 16); else ICH_C1W(3, 32); }
  else { if (npad == 16) ICH_C1W(1, 16); else ICH_C1W(1, 32); }
#undef ICH_C1W
  return ich_check_launch("ich_conv_cin1_tc_wgrad");
}

}  // extern "C"
