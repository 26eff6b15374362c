// Plane-streaming tcgen05 convolution (3x3x3, forward and data-gradient): the depth taps are folded into the MMA N dimension.
//
// conv_tc.cu loads, for every output plane d, the three input planes d-1, d, d+1 and issues 27 MMAs of N = Cout-block per
// channel chunk and tile.  Measured on B200 (scratch/mma_rate.cu, profiles/r01_mma_issue_notes.md) an SS-mode MMA costs
// ~76 cycles for ANY N <= 64 (operand fetch), so small-Cout layers are bound by the NUMBER of MMAs and by the 3x re-read
// of every input plane.  Here a CTA walks the planes of a column (n, row-block) once:
//   * input plane p is loaded ONCE (one TMA box, 32-byte swizzled rows as in conv_tc.cu);
//   * it contributes to the output planes p-1, p, p+1 with the taps kd = 2, 1, 0.  The weights are packed
//     [kh][kw][kd' = 2-kd][Cout][Cin], so for a fixed (kh, kw) the three depth taps are 3*NB CONSECUTIVE B rows and ONE MMA of
//     N = 3*NB accumulates into three adjacent accumulator slots (output planes p-1, p, p+1) -- 9 MMAs per plane, tile and
//     chunk instead of 27, each doing 3x the work for the same A fetch;
//   * the accumulators are a ring of 4 slots in TMEM ([tile][slot][NB] columns): when plane p has been consumed, output
//     plane p-1 is complete; the epilogue drains it (bias / ReLU / bf16 / BatchNorm partial sums), zero-fills the slot and
//     hands it back while the MMAs of the next planes run.  Every MMA accumulates (slots are always zero when acquired).
//     When the three live slots wrap around the ring the MMA is split in two.
// Same warp roles / pipeline primitives as conv_tc.cu.  Work item = (n, depth segment of DS planes, row block, w block).
#include "tc_common.cuh"

namespace {

struct SParams {
  int N, D, H, W, Cin, Cout;
  int WB, PW, R, RB, T, row_mode, NB;
  int n_wb, n_rb, n_nb, KC;
  int DS, n_ds;
  int slots, slot_shift;     // accumulator ring size (4 or 8) and log2 of it
  int stages;
  int b_resident;            // 1: the weights of ALL channel chunks stay in shared memory for the CTA's lifetime (stage = slab only)
  uint32_t a_bytes, a_tx_bytes, b_bytes, stage_bytes, w_offset, tmem_cols;
  long long n_items;
  bf16* y;
  int y_ld;
  const float* bias;
  int relu;
  double* stat_sum;
  double* stat_sumsq;
  int issuers;   // MMA issuer warps: 2, or 3 (warp 3 joins) when an item has >= 3 tiles
  int cs;        // thread-block cluster size (1 = none).  > 1: the CTAs of a cluster work on adjacent row blocks of the SAME (sample, depth
                 // segment, cout block) in lockstep, and the weight tile of every (plane, channel chunk) stage -- identical for all of them --
                 // is fetched from L2 ONCE and TMA-multicast into the stage of every CTA (weights that do not fit in shared memory made
                 // the kd-fold L2-bound on the mid-resolution layers: 55 KB per stage for 2 tiles of work)
  int dbg;   // profiling ablations (ICH_TC_DBG): 1 = no MMA issue, 2 = no TMA slab loads, 4 = no epilogue math / stores
  int y32;   // output rows are 32-byte aligned: 256-bit stores
  int cring; // 1: ONE accumulator ring of slots * T cells shared by the tiles (tile t, output plane g -> cell (slots * t + g) mod (slots * T)):
             // the three live planes of a tile then wrap -- and its MMA splits in two -- only where the ring passes the end of the
             // allocation, 2 of every slots * T planes instead of 2 of every `slots` (N = 96: 56 cycles unsplit, 48 + 40 split)
};

constexpr int S_THREADS = 384;      // warp 0: TMA, warps 1 and 2: MMA issuers (even / odd tiles), warp 3: idle, warps 4..11: epilogue (two per TMEM lane quadrant)
constexpr int S_EPI_WARPS = 8;
constexpr int S_ISSUERS = 2;
constexpr int S_MAX_STAGES = 6;
constexpr int MAX_SLOTS = 8;          // accumulator ring: p.slots = 4 or 8 output planes (power of two)
constexpr int S_MAX_TI = 2;           // tiles per issuer warp (T <= 4 tiles per item, >= 2 issuers)

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t s_pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool STATS, int SNB>
__global__ void __launch_bounds__(S_THREADS, 1)
conv_tc_stream_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const SParams p,
                      const __grid_constant__ CUtensorMap map_w_lo, const __grid_constant__ CUtensorMap map_w_hi) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[S_MAX_STAGES], empty_bar[S_MAX_STAGES], tfull_bar[MAX_SLOTS], tempty_bar[MAX_SLOTS], w_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_stat[2][64];          // per-CTA BatchNorm partial sums (one global atomic per channel per CTA)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (STATS && threadIdx.x < 128) s_stat[threadIdx.x >> 6][threadIdx.x & 63] = 0.f;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    // cluster mode: a stage is free when the issuers of EVERY CTA of the cluster have released it (the leader's multicast writes all of them)
    for (int s = 0; s < S_MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], p.issuers * p.cs); }
    for (int a = 0; a < MAX_SLOTS; ++a) { mbar_init(&tfull_bar[a], p.issuers); mbar_init(&tempty_bar[a], S_EPI_WARPS); }
    mbar_init(&w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t crank = p.cs > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
  if (p.cs > 1) cluster_sync_all();          // every CTA's barriers are initialised before any peer signals them
  const uint32_t tmem_base = tmem_base_smem;
  if (warp >= 4 && warp <= 7) {   // all accumulator slots start out zero: every MMA accumulates
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < (uint32_t)(p.T * p.slots * p.NB); c += 16) tmem_st16_zero(lane_base + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int nb_fixed = (int)(blockIdx.x % p.n_nb);
  const long long s_begin = blockIdx.x / p.n_nb, s_step = gridDim.x / p.n_nb, n_spatial = p.n_items / p.n_nb;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (p.b_resident && s_begin < n_spatial) {   // all channel chunks of this CTA's weight block, once (they are re-read 100s of times otherwise,
      if (elect_one()) {                         // by every CTA at the same moment: hot L2 lines, address-dependent slowdowns were measured)
        mbar_expect_tx(&w_bar, p.b_bytes * (uint32_t)p.KC);
        for (int kc = 0; kc < p.KC; ++kc)
          tma_load_3d(smem + p.w_offset + (size_t)kc * p.b_bytes, &map_w, &w_bar, kc * 16, nb_fixed * p.NB, 0);
      }
      __syncwarp();
    }
    int stage = 0; uint32_t phase = 0;
    for (long long sp = s_begin; sp < n_spatial; sp += s_step) {
      long long t = sp;
      const int wb = (int)(t % p.n_wb); t /= p.n_wb;
      const int rb = (int)(t % p.n_rb); t /= p.n_rb;
      const int ds = (int)(t % p.n_ds); const int n = (int)(t / p.n_ds);
      const int w0 = wb * p.WB, h0 = rb * p.R;
      const int d0 = ds * p.DS, dend = min(p.D, d0 + p.DS);
      for (int pl = d0 - 1; pl <= dend; ++pl) {
        if (pl < 0 || pl >= p.D) continue;
        for (int kc = 0; kc < p.KC; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sb = sa + p.a_bytes;
          if (elect_one()) {
            if (ICH_DBG(p) & 2) mbar_arrive(&full_bar[stage]);
            else {
              mbar_expect_tx(&full_bar[stage], p.a_tx_bytes + (p.b_resident ? 0u : p.b_bytes));
              tma_load_4d(sa, &map_x, &full_bar[stage], kc * 16, w0 - 1, h0 - 1, n * p.D + pl);
              if (!p.b_resident) {
                if (p.cs == 1) tma_load_3d(sb, &map_w, &full_bar[stage], kc * 16, nb_fixed * p.NB, 0);
                // cluster: each CTA fetches ITS SHARE of the 27 taps and multicasts it into the stage of all -- 1/cs of the L2 reads and, per
                // CTA, 1/cs of the TMA row requests (the 32-byte rows of the 55 KB weight tile are what bounds the producer)
                else {
                  const int wt = (27 + p.cs - 1) / p.cs, t0 = (int)crank * wt;      // 14 / 13 taps (pair), 7 / 7 / 7 / 6 (cluster of 4)
                  tma_load_3d_multicast(sb + (size_t)t0 * p.NB * 32, (int)crank == p.cs - 1 ? &map_w_hi : &map_w_lo, &full_bar[stage], kc * 16,
                                        nb_fixed * p.NB, t0, cmask);
                }
              }
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 1 && warp <= p.issuers) {
    // ===================================================== MMA issuers.  The issue loop is bound by the latency of its own (uniform
    // datapath) instruction stream -- ~68 cycles per MMA measured with the MMAs themselves ablated, against 56 cycles of tensor
    // work for N = 96 -- so TWO warps on different SM sub-partitions issue the even and the odd tiles of every (plane, chunk, tap).
    const int ii = warp - 1, NI = p.issuers;
    const uint32_t NB = (uint32_t)p.NB;
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t desc_hi = (256u >> 4) | (1u << 14) | (6u << 29);   // SBO = 256 B, version 1, SWIZZLE_32B
    const uint32_t row16 = ((uint32_t)p.PW * 32u) >> 4;
    const uint32_t tile16 = (p.row_mode ? (uint32_t)p.PW : 128u) * 2u;
    const int T = p.T;
    int stage = 0; uint32_t phase = 0;
    long long g_base = 0;                                              // output planes completed by this CTA so far
    if (p.b_resident && s_begin < n_spatial) { mbar_wait(&w_bar, 0); tc_fence_after(); }
    const uint32_t w_base = smem_u32(smem + p.w_offset);
    for (long long sp = s_begin; sp < n_spatial; sp += s_step) {
      long long t = sp / p.n_wb / p.n_rb;
      const int ds = (int)(t % p.n_ds);
      const int d0 = ds * p.DS, dend = min(p.D, d0 + p.DS);
      int acquired = d0 - 1;
      for (int pl = d0 - 1; pl <= dend; ++pl) {
        if (pl >= 0 && pl < p.D) {
          const int lo = max(d0, pl - 1), hi = min(dend - 1, pl + 1);
          for (int d = acquired + 1; d <= hi; ++d) {                   // first touch of an output plane: its slot must be drained + zeroed
            const long long g = g_base + (d - d0);
            mbar_wait(&tempty_bar[g & (p.slots - 1)], (uint32_t)(((g >> p.slot_shift) & 1) ^ 1));
            acquired = d;
          }
          tc_fence_after();
          // The live output planes lo..hi of a tile occupy <= 2 runs of consecutive ring cells (the ring wraps): resolve the runs ONCE per
          // input plane and tile -- accumulator column, instruction descriptor (N = run length * NB) and the depth-tap row offset inside a
          // (kh, kw) weight group -- so that the per-tap work below is two adds and a pack per MMA.
          uint32_t r_dcol0[S_MAX_TI], r_idesc0[S_MAX_TI], r_dcol1[S_MAX_TI], r_idesc1[S_MAX_TI], r_brow1[S_MAX_TI];
          bool r_two[S_MAX_TI];
          const int cnt = hi - lo + 1;
          const uint32_t r_brow0 = (uint32_t)(lo - (pl - 1)) * NB * 2u;     // kd' = d - pl + 1, rows of 32 B = 2 units
          {
            const long long g0 = g_base + (lo - d0);
            const int ring = p.cring ? p.slots * T : p.slots;               // cells per ring: shared by the tiles, or one ring per tile
            const int o_plane = p.cring ? (int)(g0 % ring) : (int)(g0 & (p.slots - 1));
#pragma unroll
            for (int j = 0; j < S_MAX_TI; ++j) {
              const int tt = ii + j * NI;
              int base = 0, o = o_plane;
              if (p.cring) { o += tt * p.slots; if (o >= ring) o -= ring; if (o >= ring) o %= ring; }
              else base = tt * p.slots;
              const int m0 = min(cnt, ring - o);
              r_dcol0[j] = tmem_base + (uint32_t)(base + o) * NB;
              r_idesc0[j] = idesc_base | ((((uint32_t)m0 * NB) >> 3) << 17);
              r_two[j] = m0 < cnt;
              r_dcol1[j] = tmem_base + (uint32_t)base * NB;                 // wrapped: first cell of the ring
              r_idesc1[j] = idesc_base | ((((uint32_t)(cnt - m0) * NB) >> 3) << 17);
              r_brow1[j] = (uint32_t)(lo + m0 - (pl - 1)) * NB * 2u;
            }
          }
          for (int kc = 0; kc < p.KC; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
            const uint32_t a_lo0 = (((sa & 0x3FFFFu) >> 4) | (1u << 16)) + (uint32_t)ii * tile16;
            const uint32_t b_addr = p.b_resident ? (w_base + (uint32_t)kc * p.b_bytes) : (sa + p.a_bytes);
            uint32_t b_tap = ((b_addr & 0x3FFFFu) >> 4) | (1u << 16);
            uint32_t a_kh = a_lo0;
            if (!(ICH_DBG(p) & 1)) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh, a_kh += row16) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw, b_tap += 6u * NB) {
                  const uint64_t bdesc0 = pack64(b_tap + r_brow0, desc_hi);
                  uint32_t a_lo = a_kh + 2u * (uint32_t)kw;
#pragma unroll
                  for (int j = 0; j < S_MAX_TI; ++j) {
                    if (ii + j * NI < T) {
                      const uint64_t adesc = pack64(a_lo, desc_hi);
                      if (elect_one()) umma_bf16(r_dcol0[j], adesc, bdesc0, r_idesc0[j], 1u);
                      if (r_two[j]) {
                        if (elect_one()) umma_bf16(r_dcol1[j], adesc, pack64(b_tap + r_brow1[j], desc_hi), r_idesc1[j], 1u);
                      }
                      a_lo += (uint32_t)NI * tile16;
                    }
                  }
                }
              }
            }
            __syncwarp();
            if (elect_one()) { if (p.cs > 1) umma_commit_multicast(&empty_bar[stage], cmask); else umma_commit(&empty_bar[stage]); }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        const int dc = pl - 1;                                          // output plane completed by this step
        if (dc >= d0 && dc < dend) {
          const long long g = g_base + (dc - d0);
          if (elect_one()) umma_commit(&tfull_bar[g & (p.slots - 1)]);
          __syncwarp();
        }
      }
      g_base += dend - d0;
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue warps (4..11): TMEM lane quadrant = warp % 4, two warps per quadrant
    // taking the even / odd tiles of every output plane.  One warp needs ~600 cycles per 16-column chunk (tcgen05.ld latency + a
    // dependent instruction stream): with four warps the Cin <= 32 layers were bound by the epilogue, not by the MMAs.
    const int q = warp & 3, eset = (warp - 4) >> 2;
    const int l = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float csum[STATS ? SNB : 1], csq[STATS ? SNB : 1];
    if (STATS) {
#pragma unroll
      for (int k = 0; k < SNB; ++k) csum[k] = csq[k] = 0.f;
    }
    const int n0 = nb_fixed * p.NB;
    const bool plain = !(ICH_DBG(p) & 4);
    // 16 accumulator columns -> bias / ReLU -> bf16 (-> statistics of the stored values) -> two 16-byte stores
    auto emit16 = [&](const uint32_t* v, const int c0, bf16* yrow) __attribute__((always_inline)) {
      uint32_t pk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a = __uint_as_float(v[2 * k]), b = __uint_as_float(v[2 * k + 1]);
        if (!STATS) {
          if (p.bias) { a += p.bias[n0 + c0 + 2 * k]; b += p.bias[n0 + c0 + 2 * k + 1]; }
          if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
        }
        pk[k] = s_pack_bf16x2(a, b);
      }
      if (STATS) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float lo = __uint_as_float(pk[k] << 16), hi = __uint_as_float(pk[k] & 0xffff0000u);
          const int i0 = (c0 + 2 * k) % SNB, i1 = (c0 + 2 * k + 1) % SNB;
          csum[i0] += lo; csq[i0] = fmaf(lo, lo, csq[i0]);
          csum[i1] += hi; csq[i1] = fmaf(hi, hi, csq[i1]);
        }
      }
      st_global_32B(yrow + c0, pk, p.y32 != 0);
    };
    long long g_base = 0;
    for (long long sp = s_begin; sp < n_spatial; sp += s_step) {
      long long t = sp;
      const int wb = (int)(t % p.n_wb); t /= p.n_wb;
      const int rb = (int)(t % p.n_rb); t /= p.n_rb;
      const int ds = (int)(t % p.n_ds); const int n = (int)(t / p.n_ds);
      const int w0 = wb * p.WB, h0 = rb * p.R;
      const int d0 = ds * p.DS, dend = min(p.D, d0 + p.DS);
      for (int d = d0; d < dend; ++d) {
        const long long g = g_base + (d - d0);
        const int slot = (int)(g & (p.slots - 1));
        const int ring = p.slots * p.T, gm = (int)(g % ring);
        mbar_wait(&tfull_bar[slot], (uint32_t)((g >> p.slot_shift) & 1));
        tc_fence_after();
        for (int tt = eset; tt < p.T; tt += 2) {
          const int f = (p.row_mode ? tt * p.PW : tt * 128) + l;
          const int r = f / p.PW, pos = f - r * p.PW;
          const bool valid = (pos < p.WB) && (r < p.R) && (h0 + r < p.H) && (w0 + pos < p.W) && plain;
          const long long vox = (((long long)n * p.D + d) * p.H + (h0 + r)) * p.W + (w0 + pos);
          bf16* yrow = p.y + vox * p.y_ld + n0;
          const uint32_t taddr = lane_base + (uint32_t)((p.cring ? (tt * p.slots + gm) % ring : tt * p.slots + slot) * p.NB);
          if ((p.NB & 31) == 0) {
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {       // compile-time column offsets: the statistics stay in registers
              if (c0 < p.NB) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
                tmem_st16_zero(taddr + (uint32_t)c0);                   // hand the slot back zeroed
                tmem_st16_zero(taddr + (uint32_t)c0 + 16u);
                if (valid) { emit16(v, c0, yrow); emit16(v + 16, c0 + 16, yrow); }
              }
            }
          } else {
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
              if (c0 < p.NB) {
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
                tmem_st16_zero(taddr + (uint32_t)c0);
                if (valid) emit16(v, c0, yrow);
              }
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[slot]);
      }
      g_base += dend - d0;
    }
    if (STATS) {   // warp tree -> shared-memory sums of the 8 epilogue warps -> ONE fp64 atomic per channel per CTA
#pragma unroll
      for (int k = 0; k < SNB; ++k) {
        const float a = warp_sum(csum[k]), b = warp_sum(csq[k]);
        if (lane == 0) { atomicAdd(&s_stat[0][k], a); atomicAdd(&s_stat[1][k], b); }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(S_EPI_WARPS * 32) : "memory");
      if (warp == 4) {
        for (int k = lane; k < p.NB; k += 32) {
          atomicAdd(&p.stat_sum[n0 + k], (double)s_stat[0][k]);
          atomicAdd(&p.stat_sumsq[n0 + k], (double)s_stat[1][k]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();          // no CTA leaves while a peer may still signal its barriers / write its stages
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

struct SPlan {
  bool ok = false;
  SParams p{};
  size_t smem_bytes = 0;
};

SPlan make_splan(int N, int D, int H, int W, int Cin, int Cout) {
  SPlan pl;
  if (Cin % 16 || Cout % 16 || Cin <= 0 || Cout <= 0) return pl;
  if (N <= 0 || D < 4 || H <= 0 || W < 4) return pl;
  int WB;
  if (W <= 128) WB = W;
  else if (W % 128 == 0) WB = 128;
  else return pl;
  const int PW = WB + 2;
  int NB = 0;
  static int nb_cap = -1;      // experiment: ICH_TC_STREAM_NB=32 caps the cout block (weights of a 64 -> 64 layer then stay resident, T = 4 tiles)
  if (nb_cap < 0) { const char* e = getenv("ICH_TC_STREAM_NB"); nb_cap = e ? atoi(e) : 64; if (nb_cap < 16 || nb_cap > 64) nb_cap = 64; }
  for (int c = nb_cap; c >= 16; c -= 16)
    if (Cout % c == 0) { NB = c; break; }
  if (!NB) return pl;
  const bool row_mode = (WB == 128);
  // Ring of 8 output planes for the narrowest cout block: the three live planes of an input plane wrap around the ring (and the MMA
  // of N = 3*NB splits into two) for 2 of every `slots` planes -- a quarter instead of half of them.  Measured: a win for NB = 16
  // (d0.c2 data-gradient 0.328 -> 0.293 ms, the tile count per item stays 4), a loss for NB = 32 (only 2 tiles per item fit the 512
  // TMEM columns: twice the halo rows and half the MMAs per pipeline stage; u2.c1 forward 0.744 -> 0.849 ms).  ICH_TC_STREAM_SLOTS=32
  // forces it for NB = 32 (experiments).
  static int slots_env = -1;
  if (slots_env < 0) { const char* e = getenv("ICH_TC_STREAM_SLOTS"); slots_env = e ? atoi(e) : 16; }
  const int SLOTS = (slots_env >= 8 && NB <= slots_env && NB <= 32) ? 8 : 4;
  const int Tmax = 512 / (SLOTS * NB) < 2 * S_MAX_TI ? 512 / (SLOTS * NB) : 2 * S_MAX_TI;
  const uint32_t b_bytes = 27u * NB * 32u;
  int bestR = 0, bestT = 0, bestStages = 0, bestRes = 0;
  size_t best_smem = 0;
  uint32_t best_a = 0;
  long long best_cost = -1;
  for (int R = 1; R <= H && R <= 64; ++R) {
    const int RB = R + 2;
    const int T = row_mode ? R : (((R - 1) * PW + WB) + 127) / 128;
    if (T > Tmax) break;
    uint32_t a_bytes = (uint32_t)RB * PW * 32u;
    a_bytes = (a_bytes + 127u) & ~127u;
    long long over = row_mode ? 0 : ((long long)(128 * T + 2 * PW + 2) - (long long)RB * PW) * 32;
    if (over < 0) over = 0;
    // resident weights (all Cin/16 chunks) when they leave room for >= 3 slab stages, else weights travel with every stage
    const size_t w_all = (size_t)b_bytes * (Cin / 16);
    size_t stage = ((size_t)a_bytes + 1023) & ~(size_t)1023;
    int resident = 1;
    int stages = (SMEM_LIMIT > w_all + (size_t)over + 2048) ? (int)((SMEM_LIMIT - w_all - (size_t)over - 2048) / stage) : 0;
    if (stages < 3) {
      resident = 0;
      stage = ((size_t)a_bytes + b_bytes + 1023) & ~(size_t)1023;
      stages = (int)((SMEM_LIMIT - (size_t)over - 1024) / stage);
    }
    if (stages > S_MAX_STAGES) stages = S_MAX_STAGES;
    if (stages < 2) continue;
    size_t total = (size_t)stages * stage + (resident ? w_all + 1024 : 0) + (size_t)over + 1024;
    long long blocks = (H + R - 1) / R;
    long long cost = blocks * T * 1000 + blocks * RB * 30;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost; bestR = R; bestT = T; bestStages = stages; best_smem = total; best_a = a_bytes; bestRes = resident;
    }
  }
  if (best_cost < 0) return pl;
  SParams& p = pl.p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.WB = WB; p.PW = PW; p.R = bestR; p.RB = bestR + 2; p.T = bestT; p.row_mode = row_mode; p.NB = NB;
  p.n_wb = (W + WB - 1) / WB; p.n_rb = (H + bestR - 1) / bestR; p.n_nb = Cout / NB; p.KC = Cin / 16;
  p.DS = D >= 16 ? 16 : D;
  p.n_ds = (D + p.DS - 1) / p.DS;
  p.stages = bestStages;
  p.a_bytes = best_a;
  p.a_tx_bytes = (uint32_t)p.RB * PW * 32u;     // bytes the TMA box really delivers (a_bytes is rounded up for alignment)
  p.b_bytes = b_bytes;
  p.b_resident = bestRes;
  p.stage_bytes = (uint32_t)(((size_t)best_a + (bestRes ? 0 : b_bytes) + 1023) & ~(size_t)1023);
  p.w_offset = (uint32_t)((size_t)bestStages * p.stage_bytes);   // resident weights sit after the slab stages (1024-byte aligned)
  uint32_t cols = 32;
  while (cols < (uint32_t)(SLOTS * bestT * NB)) cols <<= 1;
  p.slots = SLOTS; p.slot_shift = SLOTS == 8 ? 3 : 2;
  p.tmem_cols = cols;
  p.n_items = (long long)N * p.n_ds * p.n_rb * p.n_wb * p.n_nb;
  // thread-block clusters with the weight stage multicast (see SParams::cs): only where the weights travel with every stage, and where
  // consecutive CTAs are adjacent row blocks of one (sample, depth segment): one cout block, one w block, row blocks a multiple of cs
  static int cs_env = -1;
  if (cs_env < 0) { const char* e = getenv("ICH_TC_STREAM_CLUSTER"); cs_env = e ? atoi(e) : 0; }
  p.cs = 1;
  // lockstep needs the same number of processed planes in every item: one depth segment, or two full ones (each skips one halo plane)
  const bool same_planes = p.n_ds == 1 || (p.n_ds == 2 && D % p.DS == 0);
  if ((cs_env == 2 || cs_env == 4) && !bestRes && p.n_nb == 1 && same_planes && p.n_items % cs_env == 0 && p.n_items >= 2 * cs_env) p.cs = cs_env;
  pl.smem_bytes = best_smem;
  pl.ok = true;
  return pl;
}

int stream_shared_ring(const SParams& p) {
  static int cr = -1;
  if (cr < 0) { const char* e = getenv("ICH_TC_STREAM_CRING"); cr = e ? atoi(e) : 1; if (cr < 0 || cr > 2) cr = 1; }
  return cr == 2 || (cr == 1 && p.slots == 4 && p.NB <= 32) ? 1 : 0;
}

}  // namespace

extern "C" int ich_conv_tc_stream_plan_info(int N, int D, int H, int W, int Cin, int Cout, long long* out) {
  SPlan pl = make_splan(N, D, H, W, Cin, Cout);
  if (!pl.ok) return 1;
  const SParams& p = pl.p;
  out[0] = p.NB; out[1] = p.R; out[2] = p.T; out[3] = p.slots; out[4] = p.stages; out[5] = p.b_resident; out[6] = (long long)pl.smem_bytes;
  out[7] = p.tmem_cols; out[8] = stream_shared_ring(p); out[9] = p.n_items;
  return 0;
}

// Called by conv_tc.cu's entry points (ich_conv_tc_variant decides which kernel a shape uses).
bool ich_stream_eligible(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW) {
  const char* e = getenv("ICH_TC_STREAM");
  if (e && atoi(e) == 0) return false;
  if (KD != 3 || KH != 3 || KW != 3 || !get_encode()) return false;
  // Measured (profiles/r01_conv_layers.txt): the streaming kernel wins on the full-resolution layers (row-exact 128-wide
  // tiles, big planes: 1.3-1.6x) and loses on the small planes of the deeper levels, where the per-plane weight reload and the
  // TMEM limit on tiles per item (4 slots) dominate.  ICH_TC_STREAM=2 forces it everywhere (tests).
  // ... except narrow-Cin / wide-Cout layers on 64-wide rows (d1.c2, 32 -> 64: 0.168 -> 0.139 ms), where N = 192 per MMA pays.
  const bool narrow_to_wide = (W % 64 == 0 && Cin <= 32 && Cout >= 64 && Cout % 64 == 0);
  if (!(e && atoi(e) == 2) && W % 128 != 0 && !narrow_to_wide) return false;
  return make_splan(N, D, H, W, Cin, Cout).ok;
}

int ich_stream_launch(const void* x, int x_ld, const void* wpack_bf16, const float* bias, void* y, int y_ld, int N, int D, int H, int W, int Cin,
                      int Cout, int relu, double* stat_sum, double* stat_sumsq, cudaStream_t stream, const char* what) {
  SPlan pl = make_splan(N, D, H, W, Cin, Cout);
  ICH_REQUIRE(pl.ok, "%s: unsupported shape for the plane-streaming kernel", what);
  ICH_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)wpack_bf16 & 15) == 0,
              "%s: pointers / pitches must be 16-byte aligned (x_ld %d, y_ld %d)", what, x_ld, y_ld);
  EncodeTiledFn enc = get_encode();
  SParams& p = pl.p;
  p.y = (bf16*)y; p.y_ld = y_ld; p.bias = bias; p.relu = relu;
  p.y32 = (((uintptr_t)y & 31) == 0 && y_ld % 16 == 0) ? 1 : 0;
  p.stat_sum = stat_sum; p.stat_sumsq = stat_sumsq;
  { const char* e = getenv("ICH_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  { static int ni = -1; if (ni < 0) { const char* e = getenv("ICH_TC_STREAM_ISSUERS"); ni = (e && atoi(e) == 3) ? 3 : 2; } p.issuers = (ni == 3 && p.T >= 3) ? 3 : 2; }
  // Shared ring where it measured faster (profiles/r02x_stream_cring.txt): 4-slot plans with cout blocks of 32 (u2.c1 / u2.c2 / d0.c2 forward
  // 4-5 %, d1.c2 data-gradient 10 %); the 8-slot ring of the 16-wide blocks and the 2-tile items of the 64-wide blocks were 2-5 % slower with
  // it.  ICH_TC_STREAM_CRING=0: never, =2: always (tests).
  p.cring = stream_shared_ring(p);
  ICH_REQUIRE((p.T + p.issuers - 1) / p.issuers <= S_MAX_TI, "%s: %d tiles per item exceed the issuer's tile table", what, p.T);
  if (stat_sum) {
    ICH_REQUIRE(stat_sumsq != nullptr, "%s: fused statistics need both buffers", what);
    cudaMemsetAsync(stat_sum, 0, sizeof(double) * Cout, stream);
    cudaMemsetAsync(stat_sumsq, 0, sizeof(double) * Cout, stream);
  }
  CUtensorMap map_x, map_w;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * D};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2};
    cuuint32_t box[4] = {16, (cuuint32_t)p.PW, (cuuint32_t)p.RB, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(x) failed with %d", what, (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 27};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t box[3] = {16, (cuuint32_t)p.NB, 27};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpack_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(w) failed with %d", what, (int)r);
  }
  CUtensorMap map_w_lo = map_w, map_w_hi = map_w;      // cluster mode: boxes of ceil(27 / cs) taps, and of the remainder for the last CTA
  if (p.cs > 1) {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 27};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t estr[3] = {1, 1, 1};
    const int wt = (27 + p.cs - 1) / p.cs;
    for (int half = 0; half < 2; ++half) {
      cuuint32_t box[3] = {16, (cuuint32_t)p.NB, (cuuint32_t)(half == 0 ? wt : 27 - (p.cs - 1) * wt)};
      CUresult r = enc(half == 0 ? &map_w_lo : &map_w_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpack_bf16), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(w half %d) failed with %d", what, half, (int)r);
    }
  }
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is a per-DEVICE setting
  int attr_dev = 0; cudaGetDevice(&attr_dev);
  bool& attr_set = attr_done[attr_dev & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_stream_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_stream_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_stream_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e != cudaSuccess) cudaGetLastError();
    ICH_REQUIRE(e == cudaSuccess, "%s: cannot raise dynamic shared memory: %s", what, cudaGetErrorString(e));
    attr_set = true;
  }
  long long grid = p.n_items < ich_num_sms() ? p.n_items : ich_num_sms();
  grid = grid / p.n_nb * p.n_nb;
  if (grid < p.n_nb) grid = p.n_nb;
  if (p.cs > 1) grid = grid / p.cs * p.cs;       // whole clusters; n_items is a multiple of cs, so every CTA of a cluster gets the same number of items
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(S_THREADS);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.cs > 1 ? 1 : 0;
  cudaError_t le;
  if (stat_sum && p.NB <= 32) le = cudaLaunchKernelEx(&cfg, conv_tc_stream_kernel<true, 32>, map_x, map_w, p, map_w_lo, map_w_hi);
  else if (stat_sum) le = cudaLaunchKernelEx(&cfg, conv_tc_stream_kernel<true, 64>, map_x, map_w, p, map_w_lo, map_w_hi);
  else le = cudaLaunchKernelEx(&cfg, conv_tc_stream_kernel<false, 1>, map_x, map_w, p, map_w_lo, map_w_hi);
  (void)le;
  return ich_check_launch(what);
}
