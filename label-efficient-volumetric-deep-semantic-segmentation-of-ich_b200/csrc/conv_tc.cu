// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a: 3x3x3 (or 1x3x3) "same" conv on channel-last bf16, fp32
// accumulate in tensor memory.  Used for the forward pass and, with the flipped/transposed weight pack, the data gradient.
// Replaces the cuDNN kernels behind nn.Conv3d/Conv2d at reference models/networks/UNet.py:153,155,158,160.
//
// Shift-GEMM on a zero-padded slab (no im2col, no per-tap reloads):
//   * one TMA box load brings a slab of the input -- KD planes x (R+2) rows x (WB+2) positions x 16 channels -- into shared
//     memory; out-of-bounds rows / columns are zero-filled by TMA, which IS the convolution padding.  The tensor map
//     splits C into (8, C/8) so the slab lands as [plane][chan8-chunk][row][pos][8 ch]: the canonical NO-SWIZZLE K-major
//     UMMA layout (8 rows x 16 B core matrices, 128 B apart) for ANY starting position.
//   * a tap (kd,kh,kw) is therefore just a different descriptor start address: +kd*plane + (kh*PW + kw)*16 B.  Each M
//     tile (128 consecutive flat positions of the padded slab) issues KD*9 MMAs of K=16 per channel chunk; positions that
//     fall on the halo columns produce garbage accumulator rows that the epilogue discards.
//   * weights [tap][Cout][Cin] arrive the same way ([tap][chunk][cout][8 ch]) as the B operand.
//   * warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps2-5 = epilogue (tcgen05.ld -> bias/ReLU ->
//     bf16 -> 16 B global stores).  Two smem stages, two TMEM accumulator sets; persistent CTAs, static round-robin.
#include "tc_common.cuh"

namespace {

struct Params {
  int N, D, H, W, Cin, Cout, KD, KS;   // KS = in-plane kernel size (3, or 1 for 1x1 / transposed-conv GEMMs)
  int WB, PW, R, RB, T, row_mode, NB;
  int up_fd, up_cout;                  // transposed-conv scatter epilogue: depth factor (0 = off) and its Cout
  int n_wb, n_rb, n_nb, KC, taps;
  int nacc, stages;
  uint32_t a_bytes, a_tx_bytes, b_bytes, stage_bytes, tmem_cols;
  long long n_items;
  bf16* y;
  int y_ld;
  const float* bias;
  int relu;
  double* stat_sum;     // optional fused BatchNorm statistics (per output channel sum / sum of squares of the stored bf16 values)
  double* stat_sumsq;
  int dbg;   // ablation switches for profiling only (ICH_TC_DBG): 1 = no MMA issue, 2 = no TMA loads, 4 = no epilogue stores
  int y32;   // output rows are 32-byte aligned: 256-bit stores
  int cw;    // channels per K chunk: 16 (32-byte swizzled rows) for the 3x3 kernels; 16 / 32 / 64 (32 / 64 / 128-byte rows) for the 1x1
             // GEMMs (transposed conv, heads), whose stages are small and whose TMA loads were bound by the number of 32-byte requests
  int tap_cout;   // > 0: 1x1 GEMM whose K chunks come from the per-tap maps (transposed-conv data gradient), = Cout of the transposed conv
  int issuers;    // MMA issuer warps (2 or 3): ncu shows the MMA unit only 41-62 % busy on the N = 32 / 64 layers with two -- the issue loop, not the MMA, sets the pace
  int kds;   // kd-split (3-D layers with Cout % 128 == 0): a pipeline stage holds ONE input plane and the 9 taps of ONE depth tap, so that
             // a 128-wide cout block fits (9 x 128 x 32 B = 36 KB of weights per stage instead of 27 x 64 x 32 B = 55 KB): N = 128 MMAs
             // run at the full tensor rate (64 cycles) where N = 64 ones are operand-fetch bound (48 cycles for half the work)
};

// Transposed-conv backward without the space-to-depth re-pack: the up-sampled gradient dup[n][FD*d + i][2h + j][2w + l][co] is read in
// place.  For a fixed tap (i, j, l) it is a strided view of the fine grid on the COARSE grid (element steps 2 / 2 / FD along w / h / plane),
// i.e. an ordinary tensor map with doubled strides and a per-tap base pointer; the K chunks (data-gradient GEMM) / N chunks (weight
// gradient GEMM) of the packed layout [voxel][tap * Cout + co] become "chunk c of tap t" = one box of map t.
struct TapMaps { CUtensorMap m[8]; };

constexpr int STAGES = 2;        // weight-gradient kernel
constexpr int MAX_STAGES = 8;    // forward kernel: runtime depth (2 for the big 3x3 slabs, up to 8 for the small 1x1 GEMM stages)
// slab kernel: warp 0 TMA, warps 1 and 2 MMA issuers (even / odd tiles), warp 3 idle, warps 4.. epilogue.  Four epilogue warps when the
// BatchNorm statistics are fused (128 statistic registers per thread), eight otherwise (two per TMEM lane quadrant, each taking half of
// the accumulator columns): the transposed-conv scatter epilogue (16 column chunks per tile) was bound by its four warps.
constexpr int F_ISSUERS = 2;      // default; Params::issuers (2 or 3: warps 1..3, warp 3 is otherwise idle) is what the kernel uses
// WIDE = cout blocks of 128 with fused statistics: eight epilogue warps there too, each set keeping the statistics of its 64 columns
template <bool STATS, bool WIDE = false> __host__ __device__ constexpr int f_epi_warps() { return (STATS && !WIDE) ? 4 : 8; }
template <bool STATS, bool WIDE = false> __host__ __device__ constexpr int f_threads() { return 32 * (4 + f_epi_warps<STATS, WIDE>()); }
constexpr int W_THREADS = 256;   // weight-gradient kernel: warp 0 TMA, warps 1 / 6 / 7 MMA issuers (one per accumulator group), warps 2..5 epilogue
constexpr int W_ISSUERS = 3;

template <int KS, bool STATS, bool WIDE = false>
__global__ void __launch_bounds__(f_threads<STATS, WIDE>(), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const Params p,
               const __grid_constant__ TapMaps tmaps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stage0 A|B][stage1 A|B] ... then barriers
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_stat[2][128];         // per-CTA BatchNorm partial sums (one global atomic per channel per CTA)

  // broadcast from lane 0 so the compiler KNOWS the role index is warp-uniform (uniform branches + uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (STATS && threadIdx.x < 256) s_stat[threadIdx.x >> 7][threadIdx.x & 127] = 0.f;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], p.issuers); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], p.issuers); mbar_init(&tempty_bar[a], f_epi_warps<STATS, WIDE>()); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int planes_lo = (p.KD == 3) ? 1 : 0;     // slab plane 0 corresponds to d - planes_lo
  // a CTA keeps ONE cout block for its whole life (per-thread BatchNorm partial sums stay valid); CTAs that share a
  // spatial item (same slab, different cout block) are neighbours -> the slab is served from L2 for the second one
  const int nb_fixed = (int)(blockIdx.x % p.n_nb);
  const long long s_begin = blockIdx.x / p.n_nb, s_step = gridDim.x / p.n_nb, n_spatial = p.n_items / p.n_nb;

  if (warp == 0) {
    // ===================================================== TMA producer (warp-uniform loop, elected lane issues)
    {
      int stage = 0; uint32_t phase = 0;
      for (long long sp = s_begin; sp < n_spatial; sp += s_step) {
        long long t = sp;
        const int nb = nb_fixed;
        const int wb = (int)(t % p.n_wb); t /= p.n_wb;
        const int rb = (int)(t % p.n_rb); t /= p.n_rb;
        const int d = (int)(t % p.D); const int n = (int)(t / p.D);
        const int w0 = wb * p.WB, h0 = rb * p.R;
        // kd-split: one stage per (channel chunk, depth tap) -- the slab box is ONE plane, the weight box the 9 taps of that depth tap;
        // planes outside the volume are skipped (the issuers skip the same stages)
        const int nsub = p.kds ? 3 : 1;
        for (int kc = 0; kc < p.KC; ++kc) {
          for (int sub = 0; sub < nsub; ++sub) {
            const int dd = p.kds ? d + sub - 1 : d - planes_lo;
            if (p.kds && (dd < 0 || dd >= p.D)) continue;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
            uint8_t* sb = sa + p.a_bytes;
            if (elect_one()) {
              if (ICH_DBG(p) & 2) { mbar_arrive(&full_bar[stage]); }
              else {
                mbar_expect_tx(&full_bar[stage], p.a_tx_bytes + p.b_bytes);
                if (KS == 1 && p.tap_cout) {       // K chunk kc = channels [cc, cc + cw) of tap t, read in place from the fine grid
                  const int col = kc * p.cw, t = col / p.tap_cout, cc = col - t * p.tap_cout;
                  tma_load_4d(sa, &tmaps.m[t], &full_bar[stage], cc, w0, h0, n * p.D + dd);
                } else {
                  tma_load_4d(sa, &map_x, &full_bar[stage], kc * p.cw, w0 - KS / 2, h0 - KS / 2, n * p.D + dd);
                }
                tma_load_3d(sb, &map_w, &full_bar[stage], kc * p.cw, nb * p.NB, p.kds ? sub * 9 : 0);
              }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 1 && warp <= p.issuers) {
    // ===================================================== MMA issuers (warp-uniform loops, elected lane issues).  One warp's issue
    // loop costs more cycles per MMA (~68) than an N <= 64 MMA itself (40-48, profiles/r01_mma_rate2.txt): two warps on different SM
    // sub-partitions issue the even and the odd M tiles of every tap.
    {
      const int ii = warp - 1, NI = p.issuers;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NB >> 3) << 17) | ((128u >> 4) << 24);
      // The single issuing thread is the critical resource (measured: ~130 cycles per MMA when descriptors were rebuilt
      // from scratch): keep the per-MMA work to two adds + one register pack.  Descriptor = {lo: start>>4 | LBO, hi: const}.
      const uint32_t desc_hi = (256u >> 4) | (1u << 14) | (6u << 29);   // SBO = 256 B, version 1, SWIZZLE_32B
      const uint32_t plane16 = ((uint32_t)p.RB * p.PW * 32u) >> 4;       // plane stride in 16-byte units
      const uint32_t row16 = ((uint32_t)p.PW * 32u) >> 4;                // one slab row
      const uint32_t btap16 = ((uint32_t)p.NB * 32u) >> 4;               // one tap of the weight tile
      const uint32_t tile16 = (p.row_mode ? (uint32_t)p.PW : 128u) * ((uint32_t)p.cw >> 3); // M-tile step (cw * 2 B per position)
      // 1x1 GEMMs with wide chunks: rows of 64 / 128 bytes (SWIZZLE_64B / 128B), cw / 16 K-steps of 32 bytes inside every row
      const uint32_t wide_hi = p.cw == 64 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : p.cw == 32 ? ((512u >> 4) | (1u << 14) | (4u << 29)) : desc_hi;
      const uint32_t NB = (uint32_t)p.NB;
      const int T = p.T;
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      for (long long sp = s_begin; sp < n_spatial; sp += s_step, ++it) {
        long long t = sp / p.n_wb / p.n_rb;
        const int d = (int)(t % p.D);
        // kd-split: the depth taps are separate pipeline stages of ONE plane each (sub-stages below), not planes of one slab
        const int kd_lo = (p.KD == 3 && !p.kds && d == 0) ? 1 : 0;
        const int kd_hi = (p.KD == 3 && !p.kds) ? ((d == p.D - 1) ? 1 : 2) : 0;
        const int nsub = p.kds ? 3 : 1;
        bool first_stage = true;
        const int acc = (p.nacc == 2) ? (it & 1) : 0;
        const uint32_t acc_phase = (p.nacc == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T) * NB;
        for (int kc = 0; kc < p.KC; ++kc)
        for (int sub = 0; sub < nsub; ++sub) {
          if (p.kds && (d + sub - 1 < 0 || d + sub - 1 >= p.D)) continue;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t a_lo0 = (((sa & 0x3FFFFu) >> 4) | (1u << 16)) + (uint32_t)kd_lo * plane16;
          uint32_t b_lo = ((((sa + p.a_bytes) & 0x3FFFFu) >> 4) | (1u << 16)) + (uint32_t)(kd_lo * KS * KS) * btap16;
          uint32_t accum = first_stage ? 0u : 1u;
          first_stage = false;
          uint32_t a_kd = a_lo0;
          if (KS == 1) {
            if (!(ICH_DBG(p) & 1)) {
              for (uint32_t ks = 0; ks < ((uint32_t)p.cw >> 4); ++ks) {
                const uint64_t bdesc = pack64(b_lo + 2u * ks, wide_hi);
                uint32_t a_lo = a_lo0 + 2u * ks + (uint32_t)ii * tile16;
                uint32_t dcol = d_tmem + (uint32_t)ii * NB;
#pragma unroll 2
                for (int tt = ii; tt < T; tt += NI) {
                  if (elect_one()) umma_bf16(dcol, pack64(a_lo, wide_hi), bdesc, idesc, accum);
                  a_lo += (uint32_t)NI * tile16;
                  dcol += (uint32_t)NI * NB;
                }
                accum = 1u;
              }
            }
          } else
          for (int kd = kd_lo; kd <= kd_hi && !(ICH_DBG(p) & 1); ++kd, a_kd += plane16) {
            uint32_t a_kh = a_kd;
#pragma unroll
            for (int kh = 0; kh < KS; ++kh, a_kh += row16) {
#pragma unroll
              for (int kw = 0; kw < KS; ++kw) {
                const uint64_t bdesc = pack64(b_lo, desc_hi);
                b_lo += btap16;
                uint32_t a_lo = a_kh + 2u * (uint32_t)kw + (uint32_t)ii * tile16;
                uint32_t dcol = d_tmem + (uint32_t)ii * NB;
#pragma unroll 2
                for (int tt = ii; tt < T; tt += NI) {
                  if (elect_one()) umma_bf16(dcol, pack64(a_lo, desc_hi), bdesc, idesc, accum);
                  a_lo += (uint32_t)NI * tile16;
                  dcol += (uint32_t)NI * NB;
                }
                accum = 1u;
              }
            }
          }
          __syncwarp();
          if (elect_one()) umma_commit(&empty_bar[stage]);            // frees the smem stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull_bar[acc]);                // accumulators complete
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue warps (4..): TMEM lane quadrant = warp % 4; with eight warps the
    // second set (warps 8..11) takes the upper half of the accumulator columns of every tile
    const int q = warp & 3;
    const int l = q * 32 + lane;
    const int eset = (warp - 4) >> 2;
    const int c_lo = (f_epi_warps<STATS, WIDE>() == 8 && p.NB >= 32) ? eset * (p.NB / 32) * 16 : 0;
    const int c_hi = (f_epi_warps<STATS, WIDE>() == 8 && p.NB >= 32) ? (eset == 0 ? (p.NB / 32) * 16 : p.NB) : (eset == 0 ? p.NB : 0);
    uint32_t it = 0;
    float csum[STATS ? 64 : 1], csq[STATS ? 64 : 1];
    if (STATS) {
#pragma unroll
      for (int k = 0; k < 64; ++k) csum[k] = csq[k] = 0.f;
    }
    for (long long sp = s_begin; sp < n_spatial; sp += s_step, ++it) {
      long long t = sp;
      const int nb = nb_fixed;
      const int wb = (int)(t % p.n_wb); t /= p.n_wb;
      const int rb = (int)(t % p.n_rb); t /= p.n_rb;
      const int d = (int)(t % p.D); const int n = (int)(t / p.D);
      const int w0 = wb * p.WB, h0 = rb * p.R, n0 = nb * p.NB;
      const int acc = (p.nacc == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.nacc == 2) ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      for (int tt = 0; tt < p.T; ++tt) {
        const int f = (p.row_mode ? tt * p.PW : tt * 128) + l;
        const int r = f / p.PW, pos = f - r * p.PW;
        const bool valid = (pos < p.WB) && (r < p.R) && (h0 + r < p.H) && (w0 + pos < p.W);
        const long long vox = (((long long)n * p.D + d) * p.H + (h0 + r)) * p.W + (w0 + pos);
        bf16* yrow = p.y + vox * p.y_ld + n0;
        int bias0 = n0;
        // transposed conv: fine-grid index of tap (0,0,0) of this coarse voxel; a tap (i,j,l) adds i*4HW + j*2W + l
        const long long fine0 = p.up_fd ? ((((long long)n * p.D + d) * p.up_fd) * (2 * p.H) + 2 * (h0 + r)) * (2LL * p.W) + 2 * (w0 + pos) : 0;
        const int fine_i = 4 * p.H * p.W, fine_j = 2 * p.W;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * p.T + tt) * p.NB);
#pragma unroll
        for (int c0 = 0; c0 < (KS == 1 ? 256 : 128); c0 += 16) {
          if (c0 >= c_lo && c0 < c_hi) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (p.up_fd) {   // transposed conv k2 s2: this 16-column chunk belongs to one tap (i, j, l) -> scatter to the fine grid
              const uint32_t col = (uint32_t)(n0 + c0);
              const uint32_t tap = col / (uint32_t)p.up_cout;
              bias0 = (int)(col - tap * (uint32_t)p.up_cout) - c0;
              const int ti = (int)(tap >> 2), tj = (int)(tap >> 1) & 1, tl = (int)tap & 1;
              yrow = p.y + (fine0 + ti * fine_i + tj * fine_j + tl) * p.y_ld + bias0;
            }
            if (valid && !(ICH_DBG(p) & 4)) {
              float f32[16];
              float bv[16];
              if (p.bias) {       // 16 consecutive, 64-byte aligned bias values
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(p.bias + bias0 + c0 + 4 * k4);
                  bv[4 * k4] = t4.x; bv[4 * k4 + 1] = t4.y; bv[4 * k4 + 2] = t4.z; bv[4 * k4 + 3] = t4.w;
                }
              }
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                float a = __uint_as_float(v[k]);
                if (p.bias) a += bv[k];
                if (p.relu) a = fmaxf(a, 0.f);
                f32[k] = a;
                if (STATS && (WIDE || c0 < 64)) {   // statistics of the value as stored (bf16-rounded); WIDE: this warp set's 64 columns
                  const float rv = __bfloat162float(__float2bfloat16_rn(a));
                  csum[(c0 + k) & 63] += rv;
                  csq[(c0 + k) & 63] = fmaf(rv, rv, csq[(c0 + k) & 63]);
                }
              }
              uint32_t pk[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) pk[k] = pack_bf16x2_rn(f32[2 * k], f32[2 * k + 1]);
              st_global_32B(yrow + c0, pk, p.y32 != 0);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (STATS) {   // once per CTA: warp tree over the 32 voxel lanes, shared-memory sums of the 4 warps, ONE fp64 atomic per channel
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const float a = warp_sum(csum[k]), b = warp_sum(csq[k]);
        if (lane == 0) { atomicAdd(&s_stat[0][(WIDE ? c_lo : 0) + k], a); atomicAdd(&s_stat[1][(WIDE ? c_lo : 0) + k], b); }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(32 * f_epi_warps<STATS, WIDE>()) : "memory");
      if (warp == 4) {
        for (int k = lane; k < p.NB; k += 32) {
          atomicAdd(&p.stat_sum[nb_fixed * p.NB + k], (double)s_stat[0][k]);
          atomicAdd(&p.stat_sumsq[nb_fixed * p.NB + k], (double)s_stat[1][k]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------- host side
struct Plan {
  bool ok = false;
  Params p{};
  size_t smem_bytes = 0;
};


Plan make_plan(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW, int nb_must_divide = 0, int cw_must_divide = 0) {
  Plan pl;
  static int cw_env = -1;
  if (cw_env < 0) { const char* e = getenv("ICH_TC_WIDE_K"); cw_env = e ? atoi(e) : 1; }
  // channels per K chunk (cw_must_divide: a chunk must not straddle two taps of the per-tap maps)
  const int cdiv = cw_must_divide ? cw_must_divide : Cin;
  const int cw = (KH == 1 && cw_env) ? (cdiv % 64 == 0 ? 64 : cdiv % 32 == 0 ? 32 : 16) : 16;
  const uint32_t rowb = (uint32_t)cw * 2u;                                                   // bytes per position in a stage
  if (KH != KW || (KH != 3 && KH != 1) || (KD != 1 && KD != 3) || (KH == 1 && KD != 1)) return pl;
  const int KS = KH, hw = KS / 2;
  if (Cin % 16 || Cout % 16 || Cin <= 0 || Cout <= 0) return pl;
  if (N <= 0 || D <= 0 || H <= 0 || W < 4) return pl;
  int WB;
  if (W <= 128) WB = W;
  else if (W % 128 == 0) WB = 128;
  else return pl;
  const int PW = WB + 2 * hw;
  const int taps = KD * KS * KS;
  const bool row_mode = (WB == 128);
  long long best_cost = -1;
  int bestR = 0, bestT = 0, bestAcc = 0, bestNB = 0, bestStages = 0;
  size_t best_smem = 0;
  uint32_t best_a = 0;
  // cout block: <= 64 for the 3x3x3 kernels whose stage holds all 27 taps, 128 when a stage holds 9 taps (2-D layers, and 3-D layers in
  // kd-split mode: one plane + one depth tap per stage) -- an N = 128 MMA runs at the full tensor rate (64 cycles), an N = 64 one is
  // operand-fetch bound (48 cycles for half the work, profiles/r01_mma_rate2.txt); up to 256 for 1x1 GEMMs
  static int wide_env = -1;
  if (wide_env < 0) { const char* e = getenv("ICH_TC_NB128"); wide_env = e ? atoi(e) : 1; }
  // ... but not on the smallest planes (16 x 16 at the bottleneck of cfg-3): an item is then a single M tile and the 36 KB weight stage
  // is re-fetched for 9 MMAs of work (measured: bt.c2 forward 0.051 -> 0.069 ms with the wide block)
  const bool wide = KS == 3 && wide_env && Cout % 128 == 0 && (long long)H * W >= 1024;
  // experiment (ICH_TC_KDS64=1, off by default): per-plane staging also for the 64-wide cout blocks -- smaller stages let an item take
  // more rows, i.e. fewer weight bytes per output on the L2-bound mid-resolution layers
  static int kds64_env = -1;
  if (kds64_env < 0) { const char* e = getenv("ICH_TC_KDS64"); kds64_env = e ? atoi(e) : 0; }
  const int kds = (KD == 3 && KS == 3 && (wide || (kds64_env && Cout % 64 == 0 && (long long)H * W >= 1024))) ? 1 : 0;
  const int stage_taps = kds ? 9 : taps, stage_planes = kds ? 1 : KD;
  for (int NB = (KS == 1 ? 256 : (wide ? 128 : 64)); NB >= 16; NB -= 16) {
    if (Cout % NB || (nb_must_divide && nb_must_divide % NB)) continue;
    if (KS == 3 && NB > 64 && NB != 128) continue;              // 3x3 epilogues: column halves of 64 (statistics) -> 128 or <= 64
    const uint32_t b_bytes = (uint32_t)stage_taps * NB * rowb;
    for (int R = 1; R <= H && R <= 64; ++R) {
      const int RB = R + 2 * hw;
      const int T = row_mode ? R : (((R - 1) * PW + WB) + 127) / 128;
      int nacc = 0;
      if (2 * T * NB <= 512) nacc = 2;
      else if (T * NB <= 512) nacc = 1;
      else continue;
      uint32_t a_bytes = (uint32_t)stage_planes * RB * PW * rowb;
      a_bytes = cw > 16 ? ((a_bytes + 1023u) & ~1023u) : ((a_bytes + 127u) & ~127u);     // the weight tile follows: keep wide-swizzle tiles 1024-aligned
      long long over = row_mode ? 0 : ((long long)(128 * T + 2 * hw * PW + 2 * hw) - (long long)RB * PW) * (long long)rowb;
      if (over < 0) over = 0;
      size_t stage = ((size_t)a_bytes + b_bytes + 1023) & ~(size_t)1023;
      int stages = (int)((SMEM_LIMIT - (size_t)over - 1024) / stage);
      if (stages > MAX_STAGES) stages = MAX_STAGES;
      if (KS == 3 && stages > (kds ? 4 : 3)) stages = kds ? 4 : 3;
      if (stages < 2) continue;
      size_t total = (size_t)stages * stage + (size_t)over + 1024;   // +1024: manual alignment of the dynamic base
      long long blocks = (H + R - 1) / R;
      // MMA instructions (each ~max(76, 40 + NB/2) cycles) + halo rows + penalty for un-overlapped epilogue / shallow pipeline
      long long mma = (NB <= 64 ? 76 : 40 + NB / 2);
      long long cost = blocks * T * (Cout / NB) * mma * 13 + blocks * RB * 20 * (Cout / NB) + (nacc == 1 ? blocks * T * 150 : 0) +
                       (KS == 1 && stages < 4 ? blocks * T * 500 : 0);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost; bestR = R; bestT = T; bestAcc = nacc; best_smem = total; best_a = a_bytes; bestNB = NB; bestStages = stages;
      }
    }
    if (KS == 3 && best_cost >= 0) break;    // 3x3: take the largest feasible cout block (fewest slab re-reads)
  }
  const int NB = bestNB;
  const uint32_t b_bytes = (uint32_t)stage_taps * NB * rowb;
  if (best_cost < 0) return pl;
  Params& p = pl.p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KD = KD; p.KS = KS;
  p.up_fd = 0; p.up_cout = 0; p.tap_cout = 0;
  p.WB = WB; p.PW = PW; p.R = bestR; p.RB = bestR + 2 * hw; p.T = bestT; p.row_mode = row_mode; p.NB = NB;
  p.n_wb = (W + WB - 1) / WB; p.n_rb = (H + bestR - 1) / bestR; p.n_nb = Cout / NB; p.KC = Cin / cw; p.taps = taps; p.cw = cw;
  p.nacc = bestAcc; p.stages = bestStages;
  p.a_bytes = best_a;
  p.a_tx_bytes = (uint32_t)stage_planes * p.RB * PW * rowb;
  p.kds = kds;
  p.b_bytes = b_bytes;
  p.stage_bytes = (uint32_t)(((size_t)best_a + b_bytes + 1023) & ~(size_t)1023);
  uint32_t cols = 32;
  while (cols < (uint32_t)(bestAcc * bestT * NB)) cols <<= 1;
  p.tmem_cols = cols;
  p.n_items = (long long)N * D * p.n_rb * p.n_wb * p.n_nb;
  pl.smem_bytes = best_smem;
  pl.ok = true;
  return pl;
}

}  // namespace

// plane-streaming variant (conv_tc_stream.cu)
bool ich_stream_eligible(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW);
int ich_stream_launch(const void* x, int x_ld, const void* wpack_bf16, const float* bias, void* y, int y_ld, int N, int D, int H, int W, int Cin,
                      int Cout, int relu, double* stat_sum, double* stat_sumsq, cudaStream_t stream, const char* what);

extern "C" {

// Which tensor-core kernel a forward / data-gradient shape uses -- it decides the weight pack layout:
//   0 = none (CUDA-core path), 1 = slab kernel, pack [kd][kh][kw][Cout][Cin], 2 = plane-streaming kernel, pack [kh][kw][2-kd][Cout][Cin].
int ich_conv_tc_variant(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW) {
  if (!get_encode()) return 0;
  if (ich_stream_eligible(N, D, H, W, Cin, Cout, KD, KH, KW)) return 2;
  return make_plan(N, D, H, W, Cin, Cout, KD, KH, KW).ok ? 1 : 0;
}

int ich_conv_tc_supported(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW) {
  if (!get_encode()) return 0;
  return make_plan(N, D, H, W, Cin, Cout, KD, KH, KW).ok ? 1 : 0;
}

// Host-only: the tiling the slab kernel would use for a shape (no driver needed, so the planner is testable without a GPU).
// out[0..9] = NB, R, T, accumulator sets, pipeline stages, kd-split, dynamic shared memory bytes, TMEM columns, K chunk width, work items.
int ich_conv_tc_plan_info(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW, long long* out) {
  Plan pl = make_plan(N, D, H, W, Cin, Cout, KD, KH, KW);
  if (!pl.ok) return 1;
  const Params& p = pl.p;
  out[0] = p.NB; out[1] = p.R; out[2] = p.T; out[3] = p.nacc; out[4] = p.stages; out[5] = p.kds; out[6] = (long long)pl.smem_bytes;
  out[7] = p.tmem_cols; out[8] = p.cw; out[9] = p.n_items;
  return 0;
}

// Per-tap maps of the up-sampled gradient `dup` (channel slab of the fine grid, pitch ld): map t = tap (i, j, l), dims (Cout, W, H, N*D) on
// the COARSE grid, box (cbox, wbox, rbox, 1).
static int make_tap_maps(TapMaps& tm, const void* dup, int ld, int N, int D, int H, int W, int Cout, int FD, int cbox, int wbox, int rbox,
                         const char* what) {
  EncodeTiledFn enc = get_encode();
  ICH_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled not available", what);
  ICH_REQUIRE((FD == 1 || FD == 2) && ld % 8 == 0 && ((uintptr_t)dup & 15) == 0 && Cout % cbox == 0, "%s: bad transposed-conv gradient layout", what);
  const long long Wf = 2LL * W, Hf = 2LL * H;
  const CUtensorMapSwizzle sw = cbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  for (int t = 0; t < 4 * FD; ++t) {
    const int i = FD == 2 ? (t >> 2) : 0, j = (t >> 1) & 1, l = t & 1;
    const char* base = (const char*)dup + (((long long)i * Hf + j) * Wf + l) * ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * D};
    cuuint64_t strides[3] = {(cuuint64_t)2 * ld * 2, (cuuint64_t)2 * Wf * ld * 2, (cuuint64_t)FD * Hf * Wf * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)cbox, (cuuint32_t)wbox, (cuuint32_t)rbox, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm.m[t], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(tap %d) failed with %d", what, t, (int)r);
  }
  for (int t = 4 * FD; t < 8; ++t) tm.m[t] = tm.m[0];
  return 0;
}

static int launch_conv_tc(Plan& pl, const void* x, int x_ld, const void* wpack_bf16, const float* bias, void* y, int y_ld, int relu,
                          cudaStream_t stream, const char* what, double* stat_sum = nullptr, double* stat_sumsq = nullptr,
                          const TapMaps* tap_maps = nullptr) {
  Params& p = pl.p;
  const int N = p.N, D = p.D, H = p.H, W = p.W, Cin = p.Cin, Cout = p.Cout, KD = p.KD;
  ICH_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)wpack_bf16 & 15) == 0,
              "%s: pointers / pitches must be 16-byte aligned (x_ld %d, y_ld %d)", what, x_ld, y_ld);
  EncodeTiledFn enc = get_encode();
  ICH_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled not available", what);
  ICH_REQUIRE(((uintptr_t)bias & 15) == 0, "%s: the bias vector must be 16-byte aligned", what);
  p.y32 = (((uintptr_t)y & 31) == 0 && y_ld % 16 == 0) ? 1 : 0;
  p.y = (bf16*)y; p.y_ld = y_ld; p.bias = bias; p.relu = relu;
  p.stat_sum = stat_sum; p.stat_sumsq = stat_sumsq;
  { const char* e = getenv("ICH_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  { static int ni = -1; if (ni < 0) { const char* e = getenv("ICH_TC_ISSUERS"); ni = (e && atoi(e) == 2) ? 2 : 3; } p.issuers = (ni == 3 && p.T >= 3) ? 3 : 2; }
  if (stat_sum) {
    ICH_REQUIRE(stat_sumsq != nullptr && p.KS == 3 && !p.up_fd, "%s: fused statistics need both buffers and a 3x3 conv", what);
    cudaMemsetAsync(stat_sum, 0, sizeof(double) * Cout, stream);
    cudaMemsetAsync(stat_sumsq, 0, sizeof(double) * Cout, stream);
  }

  CUtensorMap map_x, map_w;
  if (p.tap_cout) map_x = tap_maps->m[0];      // the A operand comes from the per-tap maps
  else {
    // x as (C, W, H, N*D): the box lands as [plane][row][pos][16 ch] = 32-byte rows, 32B-swizzled (full 32 B L2 sectors)
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * D};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.cw, (cuuint32_t)p.PW, (cuuint32_t)p.RB, (cuuint32_t)(p.kds ? 1 : KD)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     p.cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : p.cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(x) failed with %d", what, (int)r);
  }
  {
    // w [taps][Cout][Cin] as (Cin, Cout, taps): box = [tap][NB rows][16 ch], same 32-byte swizzled rows
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)p.taps};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)p.cw, (cuuint32_t)p.NB, (cuuint32_t)(p.kds ? 9 : p.taps)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpack_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     p.cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : p.cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(w) failed with %d", what, (int)r);
  }
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is a per-DEVICE setting
  int attr_dev = 0; cudaGetDevice(&attr_dev);
  bool& attr_set = attr_done[attr_dev & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e != cudaSuccess) cudaGetLastError();
    ICH_REQUIRE(e == cudaSuccess, "%s: cannot raise dynamic shared memory: %s", what, cudaGetErrorString(e));
    attr_set = true;
  }
  long long grid = p.n_items < ich_num_sms() ? p.n_items : ich_num_sms();
  grid = grid / p.n_nb * p.n_nb;             // every CTA owns one cout block; n_items is a multiple of n_nb
  if (grid < p.n_nb) grid = p.n_nb;
  static TapMaps no_taps;      // unused unless p.tap_cout (zero-initialised; passed by value as a kernel parameter)
  const TapMaps& tm = tap_maps ? *tap_maps : no_taps;
  if (p.KS == 3 && stat_sum && p.NB == 128) conv_tc_kernel<3, true, true><<<(unsigned)grid, f_threads<true, true>(), pl.smem_bytes, stream>>>(map_x, map_w, p, tm);
  else if (p.KS == 3 && stat_sum) conv_tc_kernel<3, true><<<(unsigned)grid, f_threads<true>(), pl.smem_bytes, stream>>>(map_x, map_w, p, tm);
  else if (p.KS == 3) conv_tc_kernel<3, false><<<(unsigned)grid, f_threads<false>(), pl.smem_bytes, stream>>>(map_x, map_w, p, tm);
  else conv_tc_kernel<1, false><<<(unsigned)grid, f_threads<false>(), pl.smem_bytes, stream>>>(map_x, map_w, p, tm);
  return ich_check_launch(what);
}

int ich_conv_tc_fwd(const void* x, int x_ld, const void* wpack_bf16, const float* bias, void* y, int y_ld, int N, int D, int H, int W, int Cin,
                    int Cout, int KD, int KH, int KW, int relu, void* stream) {
  if (ich_stream_eligible(N, D, H, W, Cin, Cout, KD, KH, KW))
    return ich_stream_launch(x, x_ld, wpack_bf16, bias, y, y_ld, N, D, H, W, Cin, Cout, relu, nullptr, nullptr, (cudaStream_t)stream, "ich_conv_tc_fwd");
  Plan pl = make_plan(N, D, H, W, Cin, Cout, KD, KH, KW);
  ICH_REQUIRE(pl.ok, "ich_conv_tc_fwd: unsupported shape N%d D%d H%d W%d Cin%d Cout%d k%dx%dx%d", N, D, H, W, Cin, Cout, KD, KH, KW);
  return launch_conv_tc(pl, x, x_ld, wpack_bf16, bias, y, y_ld, relu, (cudaStream_t)stream, "ich_conv_tc_fwd");
}

// Forward conv with the BatchNorm batch statistics fused into the epilogue: sum[c], sumsq[c] (fp64, zeroed here) of the stored
// bf16 outputs -- replaces the separate ich_colstats pass over y.
int ich_conv_tc_fwd_stats(const void* x, int x_ld, const void* wpack_bf16, void* y, int y_ld, double* sum, double* sumsq, int N, int D, int H,
                          int W, int Cin, int Cout, int KD, int KH, int KW, void* stream) {
  if (ich_stream_eligible(N, D, H, W, Cin, Cout, KD, KH, KW))
    return ich_stream_launch(x, x_ld, wpack_bf16, nullptr, y, y_ld, N, D, H, W, Cin, Cout, 0, sum, sumsq, (cudaStream_t)stream, "ich_conv_tc_fwd_stats");
  Plan pl = make_plan(N, D, H, W, Cin, Cout, KD, KH, KW);
  ICH_REQUIRE(pl.ok && KH == 3, "ich_conv_tc_fwd_stats: unsupported shape N%d D%d H%d W%d Cin%d Cout%d k%dx%dx%d", N, D, H, W, Cin, Cout, KD, KH, KW);
  return launch_conv_tc(pl, x, x_ld, wpack_bf16, nullptr, y, y_ld, 0, (cudaStream_t)stream, "ich_conv_tc_fwd_stats", sum, sumsq);
}

// Transposed conv k2 s2 on tensor cores: a 1x1 GEMM [voxels x Cin] x [Cin x taps*Cout] whose epilogue scatters every
// (tap, cout-block) to the fine grid (depth-to-space) -- straight into the channel slab of the concat buffer.
// wpack_bf16 = [taps*Cout][Cin] bf16 (row n = tap*Cout + co), taps = 4*FD.  Grid args = the COARSE grid.
int ich_convT2_tc_supported(int N, int D, int H, int W, int Cin, int Cout, int FD) {
  if (!get_encode() || (FD != 1 && FD != 2) || Cout % 16) return 0;
  return make_plan(N, D, H, W, Cin, 4 * FD * Cout, 1, 1, 1).ok ? 1 : 0;
}

int ich_convT2_tc_fwd(const void* x, int x_ld, const void* wpack_bf16, const float* bias, void* y, int y_ld, int N, int D, int H, int W, int Cin,
                      int Cout, int FD, void* stream) {
  ICH_REQUIRE((FD == 1 || FD == 2) && Cout % 16 == 0, "ich_convT2_tc_fwd: unsupported FD %d / Cout %d", FD, Cout);
  Plan pl = make_plan(N, D, H, W, Cin, 4 * FD * Cout, 1, 1, 1);
  ICH_REQUIRE(pl.ok, "ich_convT2_tc_fwd: unsupported shape N%d D%d H%d W%d Cin%d Cout%d", N, D, H, W, Cin, Cout);
  pl.p.up_fd = FD;
  pl.p.up_cout = Cout;
  return launch_conv_tc(pl, x, x_ld, wpack_bf16, bias, y, y_ld, 0, (cudaStream_t)stream, "ich_convT2_tc_fwd");
}

// Transposed conv k2 s2 data gradient straight from the up-sampled gradient (no space-to-depth re-pack): dx[v][ci] = sum_{tap, co}
// dup[fine(v, tap)][co] * W[ci][co][tap] = a 1x1 GEMM over K = taps * Cout whose K chunks are boxes of the per-tap maps.
// wpack_bf16 = [Cin][taps * Cout] bf16 (ops pack 'convT_dgrad_tc').  Grid args = the COARSE grid; dup = channel slab of the fine grid.
int ich_convT2_tc_dgrad(const void* dup, int dup_ld, const void* wpack_bf16, void* dx, int dx_ld, int N, int D, int H, int W, int Cin, int Cout,
                        int FD, void* stream) {
  ICH_REQUIRE((FD == 1 || FD == 2) && Cout % 16 == 0, "ich_convT2_tc_dgrad: unsupported FD %d / Cout %d", FD, Cout);
  Plan pl = make_plan(N, D, H, W, 4 * FD * Cout, Cin, 1, 1, 1, 0, Cout);
  ICH_REQUIRE(pl.ok, "ich_convT2_tc_dgrad: unsupported shape N%d D%d H%d W%d Cin%d Cout%d", N, D, H, W, Cin, Cout);
  pl.p.tap_cout = Cout;
  TapMaps tm;
  if (int rc = make_tap_maps(tm, dup, dup_ld, N, D, H, W, Cout, FD, pl.p.cw, pl.p.PW, pl.p.RB, "ich_convT2_tc_dgrad")) return rc;
  return launch_conv_tc(pl, dup, dup_ld, wpack_bf16, nullptr, dx, dx_ld, 0, (cudaStream_t)stream, "ich_convT2_tc_dgrad", nullptr, nullptr, &tm);
}

int ich_convT2_tc_dgrad_supported(int N, int D, int H, int W, int Cin, int Cout, int FD) {
  if (!get_encode() || (FD != 1 && FD != 2) || Cout % 16) return 0;
  return make_plan(N, D, H, W, 4 * FD * Cout, Cin, 1, 1, 1, 0, Cout).ok ? 1 : 0;
}

}  // extern "C"

// =====================================================================================================================
// Weight gradient on tensor cores.  dW[tap][ci][co] = sum_pos x[pos + shift(tap)][ci] * dy[pos][co].
//   * K = voxel positions: both operands are MN-major (channels contiguous per position), no swizzle.
//   * The x slab (KD planes x (R+2) rows x (WB+2) positions x CU channels) lands as [plane][chunk][row][pos][8]: 8-channel
//     chunks are a uniform stride apart ACROSS planes, so the 128 MMA rows are (kd, ci) -- 3 planes x 32 channels = 96 live
//     rows for 3-D layers; (kh, kw) are start-address shifts again -> 9 accumulators [128 x NB] in TMEM.
//   * dy rows land without halo as [chunk][row][pos][8]; K runs over whole 16-position steps of each row, so no halo
//     column is ever multiplied (no garbage in K).
//   * persistent CTAs accumulate over all their positions in TMEM, then ONE epilogue adds the partial into dW with fp32
//     atomics (split-K across CTAs).  Planes outside the volume are requested at an out-of-range coordinate -> TMA zero fill.
// =====================================================================================================================
namespace {

struct WParams {
  int N, D, H, W, Cin, Cout, KD, KS;
  int out_mode, up_cout, taps_out;   // out_mode 1: transposed-conv layout dW[ci][co][tap], columns n = tap*up_cout + co;
                                     // out_mode 2: operands swapped (U = dy, V = x): rows = co, columns = ci, taps flipped
  int khs;                           // kh-split: a CTA handles ONE kernel row kh (3 accumulators -> N up to 128); grid.y = pairs * 3
  int kwf;                           // kw-fold: the three kw taps are the N dimension (N = 3 * 32): dy lands as 64-byte rows WITH a column halo and
                                     // MN block j of the B operand is the same tile shifted by j positions (LBO = one row) -> 3 MMAs per K step
  uint32_t a_off, b_off;             // operand offsets inside a stage
  int bu;                            // dy tile rows (non-kw-fold modes): 16-byte units per position = 2 (16 ch, SWIZZLE_32B), 4 (32 ch, 64B), 8 (64 ch, 128B)
  int a64;                           // x slab as 64-byte SWIZZLE_64B rows (32-channel blocks) instead of 32-byte rows: half the TMA requests
  int WB, PW, R, RB, CU, NB, chunks_u, chunks_v;
  int n_wb, n_rb, n_cb, n_nb;
  uint32_t plane_bytes, a_bytes, b_bytes, stage_bytes, tmem_cols;
  long long n_pos_items;   // N * D * n_rb * n_wb
  int splits;
  float* dw;
  int dbg;   // profiling ablations (ICH_TC_DBG): 1 = no MMA issue, 2 = no TMA loads
  int tap_cout;   // > 0: transposed-conv weight gradient with dy read in place from the fine grid through the per-tap maps (= its Cout)
};

template <int KS>
__global__ void __launch_bounds__(W_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const WParams p,
                     const __grid_constant__ TapMaps tmaps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], done_bar;
  __shared__ uint32_t tmem_base_smem;
  // broadcast from lane 0 so the compiler KNOWS the role index is warp-uniform (uniform branches + uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], W_ISSUERS); }
    mbar_init(&done_bar, W_ISSUERS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int n_pairs = p.n_cb * p.n_nb;
  const int pair = blockIdx.y % n_pairs;
  const int khf = p.khs ? (int)(blockIdx.y / n_pairs) : 0;      // the kernel row of this CTA in kh-split mode
  const int khn = p.khs ? 1 : KS;                               // kernel rows accumulated by this CTA
  const int cb = pair / p.n_nb, nb = pair % p.n_nb;
  const long long per = (p.n_pos_items + p.splits - 1) / p.splits;
  const long long it_begin = (long long)blockIdx.x * per;
  const long long it_end = it_begin + per < p.n_pos_items ? it_begin + per : p.n_pos_items;

  if (warp == 0) {
    {
      int stage = 0; uint32_t phase = 0;
      for (long long item = it_begin; item < it_end; ++item) {
        // depth is the FASTEST item index: consecutive items of a CTA are the same (row block, w block) at d, d + 1, ... so two of the
        // three x planes of an item were loaded by the previous one and are still in L2 (with depth slowest the re-reads went to DRAM:
        // the full-resolution layers were DRAM-bound at 3x the algorithmic bytes)
        long long t = item;
        const int d = (int)(t % p.D); t /= p.D;
        const int wb = (int)(t % p.n_wb); t /= p.n_wb;
        const int rb = (int)(t % p.n_rb); const int n = (int)(t / p.n_rb);
        const int w0 = wb * p.WB, h0 = rb * p.R;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * p.stage_bytes + p.a_off;
        uint8_t* sb = smem + (size_t)stage * p.stage_bytes + p.b_off;
        if (ICH_DBG(p) & 2) { if (elect_one()) mbar_arrive(&full_bar[stage]); }
        else if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
          for (int pl = 0; pl < p.KD; ++pl) {
            const int dd = d + pl - (p.KD == 3 ? 1 : 0);
            const int coord = (dd < 0 || dd >= p.D) ? -1 : n * p.D + dd;    // -1: out of range -> the whole plane is zero-filled
            tma_load_5d(sa + (size_t)pl * p.plane_bytes, &map_x, &full_bar[stage], 0, w0 - KS / 2, h0 - KS / 2 + khf, cb * (p.CU / (p.a64 ? 32 : 16)), coord);
          }
          if (p.kwf) tma_load_4d(sb, &map_dy, &full_bar[stage], nb * p.NB, w0 - 1, h0, n * p.D + d);
          else if (KS == 1 && p.tap_cout) {
            // the N block = NB / cbk channel chunks; chunk q = channels [cc, cc + cbk) of tap t, one box of map t each, laid out exactly
            // as the single 5-D box of the packed layout would land them ([chunk][row][pos][cbk])
            const int cbk = 8 * p.bu;
            const uint32_t chunk_bytes = (uint32_t)p.R * p.WB * cbk * 2u;
            for (int q = 0; q < p.NB / cbk; ++q) {
              const int col = nb * p.NB + q * cbk, t = col / p.tap_cout, cc = col - t * p.tap_cout;
              tma_load_4d(sb + (size_t)q * chunk_bytes, &tmaps.m[t], &full_bar[stage], cc, w0, h0, n * p.D + d);
            }
          } else tma_load_5d(sb, &map_dy, &full_bar[stage], 0, w0, h0, nb * (p.NB / (8 * p.bu)), n * p.D + d);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 || warp >= 6) {
    {
      // Three issuer warps (different SM sub-partitions), each owning its own accumulators: the issue loop of ONE warp costs more
      // cycles per MMA (~70-90, uniform-datapath instruction latency) than the MMA itself (40-56).  Issuer ii takes kernel row
      // kh = ii (kernel column kw = ii in kh-split mode); the 1x1 GEMMs have a single accumulator (issuer 0).
      const int ii = warp == 1 ? 0 : warp - 5;
      // both operands MN-major (bits 15, 16), bf16 x bf16 -> fp32, M = 128, N = NB
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((p.kwf ? 3 * p.NB : p.NB) >> 3) << 17) | ((128u >> 4) << 24);
      // MN-major SWIZZLE_32B operands: a position is a 32-byte row of 16 channels; 16-channel blocks are LBO apart (uniform
      // across the planes of the x slab), 8-position groups SBO = 256 B apart.  descriptor = {lo: start>>4 | LBO>>4 << 16,
      // hi: SBO>>4 | version | layout}; per MMA only the start changes (one add + one pack).  All offsets below are in 16-byte units.
      const uint32_t desc_hi = (256u >> 4) | (1u << 14) | (6u << 29);
      const uint32_t AU = p.a64 ? 4u : 2u;                                         // 16-byte units per x position
      const uint32_t a_lbo = (((uint32_t)p.RB * p.PW * 16u * AU) >> 4) << 16;      // channel-block stride (16 or 32 channels per block)
      const uint32_t BU = (uint32_t)p.bu;                                          // 16-byte units per dy position
      const uint32_t b_lbo = (((uint32_t)p.R * p.WB * 16u * BU) >> 4) << 16;       // channel-block stride (16 / 32 / 64 channels per block)
      const uint32_t a_hi = p.a64 ? ((512u >> 4) | (1u << 14) | (4u << 29)) : desc_hi;
      const uint32_t b_hi = BU == 8 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : BU == 4 ? ((512u >> 4) | (1u << 14) | (4u << 29)) : desc_hi;
      const uint32_t PW = AU * (uint32_t)p.PW, WB = BU * (uint32_t)p.WB, NB = (uint32_t)p.NB;   // row pitches in 16-byte units
      int stage = 0; uint32_t phase = 0;
      uint32_t accum = 0u;
      for (long long item = it_begin; item < it_end; ++item) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t sa = st + p.a_off;
        uint32_t a_row = ((sa & 0x3FFFFu) >> 4) | a_lbo;                           // one position = 32 B = 2 units
        uint32_t b_row = (((st + p.b_off) & 0x3FFFFu) >> 4) | b_lbo;
        if (p.kwf) {
          // B = dy slab [row][WB + 2 positions][32 co] as 64-byte SWIZZLE_64B rows; MN block j (32 columns) = the tile shifted by j
          // positions (LBO = 64 B), i.e. dy[q + j - 1] for x position q: column block j is kernel column kw = 2 - j.
          // NB = 64: the same with 128-byte SWIZZLE_128B rows (LBO = 128 B), N = 192, one kernel row per CTA (kh-split).
          const uint32_t KU = NB >> 3;                                             // 16-byte units per dy position: 4 (NB = 32) or 8 (NB = 64)
          const uint32_t kb_hi = KU == 8 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : KU == 4 ? ((512u >> 4) | (1u << 14) | (4u << 29)) : desc_hi;
          uint32_t bk_row = (((st + p.b_off) & 0x3FFFFu) >> 4) | (KU << 16);       // LBO = one position
          const uint32_t BW = KU * (uint32_t)(p.WB + 2);                           // dy row pitch in 16-byte units
          const bool mine = ii < khn;                                              // kh-split: a single accumulator (issuer 0)
          for (int r = 0; r < p.R && !(ICH_DBG(p) & 1) && mine; ++r, a_row += PW, bk_row += BW) {
            for (uint32_t s = 0; s < (uint32_t)p.WB; s += 16) {                    // 16 positions per MMA
              const uint64_t bdesc = pack64(bk_row + KU * s, kb_hi);
              const uint32_t a_kh = a_row + AU * s + AU + (uint32_t)ii * PW;       // x position q = s (skip the halo column), kernel row ii
              if (elect_one()) umma_bf16(tmem_base + (uint32_t)ii * 3u * NB, pack64(a_kh, a_hi), bdesc, idesc, accum);
              accum = 1u;
            }
          }
        } else
        for (int r = 0; r < p.R && !(ICH_DBG(p) & 1); ++r, a_row += PW, b_row += WB) {
          for (uint32_t s = 0; s < (uint32_t)p.WB; s += 16) {                      // 16 positions per MMA
            const uint64_t bdesc = pack64(b_row + BU * s, b_hi);
            const uint32_t a_s = a_row + AU * s;
            if (KS == 1) {                 // 1x1 GEMM: one accumulator
              if (ii == 0 && elect_one()) umma_bf16(tmem_base, pack64(a_s, a_hi), bdesc, idesc, accum);
            } else if (khn == 1) {         // kh-split: accumulators = kw, this issuer's is kw = ii
              if (elect_one()) umma_bf16(tmem_base + (uint32_t)ii * NB, pack64(a_s + AU * (uint32_t)ii, a_hi), bdesc, idesc, accum);
            } else {                       // 9 accumulators (kh, kw): this issuer's kernel row is kh = ii
              const uint32_t a_kh = a_s + (uint32_t)ii * PW;
              uint32_t dcol = tmem_base + (uint32_t)ii * (uint32_t)KS * NB;
#pragma unroll
              for (int kw = 0; kw < KS; ++kw) {
                if (elect_one()) umma_bf16(dcol, pack64(a_kh + AU * (uint32_t)kw, a_hi), bdesc, idesc, accum);
                dcol += NB;
              }
            }
            accum = 1u;
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&done_bar);
      __syncwarp();
    }
  } else if (warp >= 2 && warp <= 5 && it_begin < it_end) {
    // epilogue: TMEM lane = MMA row m = (kd, ci_local); 9 accumulators of NB columns
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int kd = m / p.CU, ci = cb * p.CU + (m % p.CU);
    const bool valid = m < p.KD * p.CU;
    const int taps = p.KD * KS * KS;
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    for (int j = 0; j < khn * KS; ++j) {
      const int tap = kd * KS * KS + khf * KS + (p.kwf ? (j / 3) * 3 + 2 - (j % 3) : j);
      for (int c0 = 0; c0 < p.NB; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.NB + c0), v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int col = nb * p.NB + c0 + k;
            size_t idx;
            if (p.out_mode == 0) idx = ((size_t)col * p.Cin + ci) * taps + tap;                    // conv: dW[co][ci][tap]
            else if (p.out_mode == 2) idx = ((size_t)ci * p.Cout + col) * taps + (taps - 1 - tap);  // swapped: rows = co ("ci" here), cols = ci
            else { const int t2 = col / p.up_cout; idx = ((size_t)ci * p.up_cout + (col - t2 * p.up_cout)) * p.taps_out + t2; }   // convT: dW[ci][co][tap]
            atomicAdd(&p.dw[idx], __uint_as_float(v[k]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

struct WPlan {
  bool ok = false;
  WParams p{};
  size_t smem_bytes = 0;
};

WPlan make_wplan(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW, bool khs = false, bool kwf = false) {
  WPlan pl;
  if (khs && KH != 3) return pl;
  if (kwf && (KH != 3 || Cout % (khs ? 64 : 16))) return pl;      // kw-fold: Cout blocks of 32 / 16 (all kernel rows) or 64 (with kh-split)
  if (KH != KW || (KH != 3 && KH != 1) || (KD != 1 && KD != 3) || (KH == 1 && KD != 1)) return pl;
  const int KS = KH, hw = KS / 2;
  if (Cin % 16 || Cout % 16 || Cin <= 0 || Cout <= 0) return pl;
  if (N <= 0 || D <= 0 || H <= 0) return pl;
  if (W > 128 && W % 128) return pl;
  if (W % 16) return pl;
  int CU = 0;
  const int cu_max = KD == 3 ? 32 : 128;
  for (int c = cu_max; c >= 16; c -= 16)
    if (Cin % c == 0) { CU = c; break; }
  int NB = 0;
  const int hwr = khs ? 0 : hw;                           // kh-split: the slab starts at the CTA's kernel row, no row halo
  for (int c = (KS == 1 ? 256 : (khs ? 128 : 48)); c >= 16; c -= 16)   // accumulators (KS*KS, or KS with kh-split) x NB columns must fit 512 TMEM columns
    if (Cout % c == 0) { NB = c; break; }
  if (kwf) NB = khs ? 64 : (Cout % 32 == 0 ? 32 : 16);
  if (!CU || !NB) return pl;
  const int chunks_u = CU / 8, chunks_v = NB / 8;
  // Search the slab shape (w-block WB x R rows): the full-resolution layers are bound by L2->SMEM traffic, so minimise the halo
  // amplification (RB/R)*(PW/WB) of the x slab under the shared-memory budget (narrower w-blocks allow more rows per slab).
  int bestR = 0, WB = 0, PW = 0;
  size_t best_smem = 0;
  double best_amp = 1e30;
  for (int wb = (W <= 128 ? W : 128); wb >= 16; wb >>= 1) {
    if (wb % 16 || W % wb) continue;
    const int pw = wb + 2 * hw;
    for (int R = 1; R <= H + 1 && R <= 32; ++R) {
      const int RB = R + 2 * hwr;
      if (((CU / 16) * RB * pw) % 4) continue;   // every plane of the slab is its own TMA destination: keep it 128-byte aligned
      size_t a = (size_t)KD * chunks_u * RB * pw * 16;
      size_t b = kwf ? (size_t)R * (wb + 2) * NB * 2 : (size_t)chunks_v * R * wb * 16;
      size_t stage = ((a + 1023) & ~(size_t)1023) + ((b + 1023) & ~(size_t)1023);   // both operand tiles start 1024-byte aligned
      // the MMA always reads 16 chunks (128 rows): chunks beyond KD*chunks_u are garbage rows, but must stay inside the allocation
      size_t over = (size_t)(16 - KD * chunks_u > 0 ? 16 - KD * chunks_u : 0) * RB * pw * 16 + (size_t)(2 * pw + 32) * 16;
      size_t total = STAGES * stage + over + 1024;
      if (total > SMEM_LIMIT) break;
      const double amp = ((double)RB / R) * ((double)pw / wb) * (1.0 + 0.02 * (128 / wb)) * (1.0 + 0.05 / R);   // mild preference for wide blocks / more rows
      if (amp < best_amp) { best_amp = amp; bestR = R; WB = wb; PW = pw; best_smem = total; }
    }
  }
  if (!bestR) return pl;
  WParams& p = pl.p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KD = KD; p.KS = KS;
  p.out_mode = 0; p.up_cout = 0; p.taps_out = 0; p.khs = khs ? 1 : 0; p.tap_cout = 0;
  p.WB = WB; p.PW = PW; p.R = bestR; p.RB = bestR + 2 * hwr; p.CU = CU; p.NB = NB; p.chunks_u = chunks_u; p.chunks_v = chunks_v;
  p.n_wb = (W + WB - 1) / WB; p.n_rb = (H + bestR - 1) / bestR; p.n_cb = Cin / CU; p.n_nb = Cout / NB;
  p.plane_bytes = (uint32_t)chunks_u * p.RB * PW * 16u;
  p.a_bytes = (uint32_t)KD * p.plane_bytes;
  p.b_bytes = kwf ? (uint32_t)bestR * (WB + 2) * (uint32_t)NB * 2u : (uint32_t)chunks_v * bestR * WB * 16u;
  p.stage_bytes = (uint32_t)((((size_t)p.a_bytes + 1023) & ~(size_t)1023) + (((size_t)p.b_bytes + 1023) & ~(size_t)1023));
  p.kwf = kwf ? 1 : 0;
  { static int b64_env = -1; if (b64_env < 0) { const char* e = getenv("ICH_TC_WGRAD_B64"); b64_env = e ? atoi(e) : 1; }
    p.bu = (b64_env && NB % 64 == 0) ? 8 : (b64_env && NB % 32 == 0) ? 4 : 2; }
  { static int a64_env = -1; if (a64_env < 0) { const char* e = getenv("ICH_TC_WGRAD_A64"); a64_env = e ? atoi(e) : 1; }
    p.a64 = (a64_env && CU % 32 == 0) ? 1 : 0; }
  // kw-fold: the 64-byte-swizzled dy tile sits at the (1024-byte aligned) stage base, the x planes follow it
  p.a_off = kwf ? (uint32_t)(((size_t)p.b_bytes + 1023) & ~(size_t)1023) : 0u;
  p.b_off = kwf ? 0u : (uint32_t)(((size_t)p.a_bytes + 1023) & ~(size_t)1023);   // dy tile (32 / 64 / 128-byte swizzled rows) 1024-byte aligned
  uint32_t cols = 32;
  while (cols < (uint32_t)((khs ? KS : KS * KS) * NB)) cols <<= 1;
  p.tmem_cols = cols;
  p.n_pos_items = (long long)N * D * p.n_rb * p.n_wb;
  const int pairs = p.n_cb * p.n_nb * (khs ? 3 : 1);
  long long splits = ich_num_sms() / pairs;      // ONE wave: splits * pairs <= #SMs (a partial second wave would double the time)
  if (splits > p.n_pos_items) splits = p.n_pos_items;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  pl.smem_bytes = best_smem;
  pl.ok = pairs <= 65535;
  return pl;
}

}  // namespace

extern "C" {

int ich_conv_tc_wgrad_supported(int N, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW) {
  if (!get_encode()) return 0;
  return make_wplan(N, D, H, W, Cin, Cout, KD, KH, KW).ok ? 1 : 0;
}

static int launch_conv_tc_wgrad(WPlan& pl, const void* x, int x_ld, const void* dy, int dy_ld, float* dw, size_t dw_elems, cudaStream_t s,
                                const char* what, const TapMaps* tap_maps = nullptr) {
  WParams& p = pl.p;
  const int N = p.N, D = p.D, H = p.H, W = p.W, Cin = p.Cin, Cout = p.Cout;
  ICH_REQUIRE(x_ld % 8 == 0 && dy_ld % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0,
              "%s: pointers / pitches must be 16-byte aligned (x_ld %d, dy_ld %d)", what, x_ld, dy_ld);
  EncodeTiledFn enc = get_encode();
  ICH_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled not available", what);
  p.dw = dw;
  { const char* e = getenv("ICH_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  if (cudaMemsetAsync(dw, 0, sizeof(float) * dw_elems, s) != cudaSuccess) return ich_check_launch(what);

  CUtensorMap map_x, map_dy;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  {
    // (c16, W, H, C/16, N*D): the box lands as [16-channel block][row][pos][16 ch] = 32-byte swizzled rows (full L2 sectors)
    const cuuint32_t cbk = p.a64 ? 32 : 16;     // channels per block = one swizzled row
    cuuint64_t dims[5] = {cbk, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(Cin / cbk), (cuuint64_t)N * D};
    cuuint64_t strides[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, 2ull * cbk, (cuuint64_t)H * W * x_ld * 2};
    cuuint32_t box[5] = {cbk, (cuuint32_t)p.PW, (cuuint32_t)p.RB, (cuuint32_t)(p.CU / cbk), 1};
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     p.a64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(x) failed with %d", what, (int)r);
  }
  if (p.kwf) {
    // (C, W, H, N*D): the box lands as [row][WB + 2 positions][32 channels] = 64-byte swizzled rows; the column halo is zero-filled
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * D};
    cuuint64_t strides[3] = {(cuuint64_t)dy_ld * 2, (cuuint64_t)W * dy_ld * 2, (cuuint64_t)H * W * dy_ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.NB, (cuuint32_t)(p.WB + 2), (cuuint32_t)p.R, 1};
    CUresult r = enc(&map_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     p.NB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : p.NB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(dy, kw-fold) failed with %d", what, (int)r);
  } else if (p.tap_cout) {
    map_dy = tap_maps->m[0];                          // dy comes from the per-tap maps
  } else {
    const cuuint32_t cbk = 8u * (cuuint32_t)p.bu;     // channels per block = one swizzled row (16 / 32 / 64)
    cuuint64_t dims[5] = {cbk, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(Cout / cbk), (cuuint64_t)N * D};
    cuuint64_t strides[4] = {(cuuint64_t)dy_ld * 2, (cuuint64_t)W * dy_ld * 2, 2ull * cbk, (cuuint64_t)H * W * dy_ld * 2};
    cuuint32_t box[5] = {cbk, (cuuint32_t)p.WB, (cuuint32_t)p.R, (cuuint32_t)(p.NB / cbk), 1};
    CUresult r = enc(&map_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     p.bu == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : p.bu == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICH_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(dy) failed with %d", what, (int)r);
  }
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is a per-DEVICE setting
  int attr_dev = 0; cudaGetDevice(&attr_dev);
  bool& attr_set = attr_done[attr_dev & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_LIMIT));
    if (e != cudaSuccess) cudaGetLastError();
    ICH_REQUIRE(e == cudaSuccess, "%s: cannot raise dynamic shared memory: %s", what, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)p.splits, (unsigned)(p.n_cb * p.n_nb * (p.khs ? 3 : 1)));
  static TapMaps no_taps;
  const TapMaps& tm = tap_maps ? *tap_maps : no_taps;
  if (p.KS == 3) conv_tc_wgrad_kernel<3><<<grid, W_THREADS, pl.smem_bytes, s>>>(map_x, map_dy, p, tm);
  else conv_tc_wgrad_kernel<1><<<grid, W_THREADS, pl.smem_bytes, s>>>(map_x, map_dy, p, tm);
  return ich_check_launch(what);
}

int ich_conv_tc_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, float* dw, int N, int D, int H, int W, int Cin, int Cout, int KD, int KH,
                      int KW, void* stream) {
  // An MN-major MMA costs ~110 cycles whatever N <= 128 is, and 9 accumulators cap N at 48.  Two ways to put more work in an MMA:
  //   kh-split : each CTA accumulates ONE kernel row (3 accumulators) -> N up to 128 (used when Cout is a multiple of 64);
  //   swap     : when only Cin is wide (64 -> 32 layers) the gradient of the transposed problem is computed (U = dy, V = x)
  //              and written back transposed with flipped taps.
  static int mode_env = -1, kwf_env = -1;
  if (mode_env < 0) { const char* e = getenv("ICH_TC_WGRAD_KHS"); mode_env = e ? atoi(e) : 1; }
  if (kwf_env < 0) { const char* e = getenv("ICH_TC_WGRAD_KWF"); kwf_env = e ? atoi(e) : 3; }
  const size_t n_dw = (size_t)Cout * Cin * KD * KH * KW;
  // narrow-Cin layers (16 -> 32): with rows = (kd, ci) only 48 of the 128 MMA rows are live; the transposed problem has
  // rows = (kd, co) = 96 live rows and N = Cin = 16 columns per accumulator (cheaper MMAs, same count)
  static int swap16_env = -1;
  if (swap16_env < 0) { const char* e = getenv("ICH_TC_WGRAD_SWAP16"); swap16_env = e ? atoi(e) : 3; }
  if ((swap16_env & 2) && KH == 3 && KD == 3 && Cin == 16 && Cout % 32 == 0) {
    // transposed problem WITH the kw-fold on the 16-channel operand: rows = (kd, co) (96 live), N = 3 * 16 = 48 (44-cycle MMAs)
    WPlan ps = make_wplan(N, D, H, W, Cout, Cin, KD, KH, KW, false, (swap16_env & 2) != 0);
    if (ps.ok) {
      ps.p.out_mode = 2;
      return launch_conv_tc_wgrad(ps, dy, dy_ld, x, x_ld, dw, n_dw, (cudaStream_t)stream, "ich_conv_tc_wgrad<swap>");
    }
  }
  //   kw-fold  : Cout blocks of 32 with the three kw taps folded into N = 96 (an MMA costs 32 + N/4 cycles of operand fetch for
  //              N <= 128, scratch/mma_rate2.cu: 3 MMAs of 56 cycles replace 9 of 40) -- the narrow-Cout (32 / 64) layers.
  if (kwf_env && KH == 3 && Cout == 32) {
    WPlan pk = make_wplan(N, D, H, W, Cin, Cout, KD, KH, KW, false, true);
    if (pk.ok) return launch_conv_tc_wgrad(pk, x, x_ld, dy, dy_ld, dw, n_dw, (cudaStream_t)stream, "ich_conv_tc_wgrad<kwf>");
  }
  //   kw-fold + kh-split : Cout = 64 layers, N = 192 (96 cycles for three taps instead of 3 x 48), one kernel row per CTA
  if ((kwf_env & 2) && KH == 3 && Cout == 64) {
    WPlan pk = make_wplan(N, D, H, W, Cin, Cout, KD, KH, KW, true, true);
    if (pk.ok) return launch_conv_tc_wgrad(pk, x, x_ld, dy, dy_ld, dw, n_dw, (cudaStream_t)stream, "ich_conv_tc_wgrad<khs,kwf>");
  }
  if (mode_env && KH == 3) {
    if (Cout % 64 == 0) {
      WPlan pk = make_wplan(N, D, H, W, Cin, Cout, KD, KH, KW, true);
      if (pk.ok) return launch_conv_tc_wgrad(pk, x, x_ld, dy, dy_ld, dw, n_dw, (cudaStream_t)stream, "ich_conv_tc_wgrad<khs>");
    } else if (Cin % 64 == 0) {
      WPlan pk = make_wplan(N, D, H, W, Cout, Cin, KD, KH, KW, true);
      if (pk.ok) {
        pk.p.out_mode = 2;
        return launch_conv_tc_wgrad(pk, dy, dy_ld, x, x_ld, dw, n_dw, (cudaStream_t)stream, "ich_conv_tc_wgrad<khs,swap>");
      }
    }
  }
  WPlan pl = make_wplan(N, D, H, W, Cin, Cout, KD, KH, KW);
  ICH_REQUIRE(pl.ok, "ich_conv_tc_wgrad: unsupported shape N%d D%d H%d W%d Cin%d Cout%d k%dx%dx%d", N, D, H, W, Cin, Cout, KD, KH, KW);
  return launch_conv_tc_wgrad(pl, x, x_ld, dy, dy_ld, dw, n_dw, (cudaStream_t)stream, "ich_conv_tc_wgrad");
}

// Transposed conv k2 s2 weight gradient: g = the up-sampled gradient re-packed to the coarse grid by ich_space_to_depth2
// ([voxel][tap*Cout + co]); dW[ci][co][tap] = sum_v x[v][ci] * g[v][tap*Cout + co]  (a 1x1 weight-gradient GEMM).
int ich_convT2_tc_wgrad_supported(int N, int D, int H, int W, int Cin, int Cout, int FD) {
  if (!get_encode() || (FD != 1 && FD != 2)) return 0;
  return make_wplan(N, D, H, W, Cin, 4 * FD * Cout, 1, 1, 1).ok ? 1 : 0;
}

int ich_convT2_tc_wgrad(const void* x, int x_ld, const void* g, int g_ld, float* dw, int N, int D, int H, int W, int Cin, int Cout, int FD,
                        void* stream) {
  WPlan pl = make_wplan(N, D, H, W, Cin, 4 * FD * Cout, 1, 1, 1);
  ICH_REQUIRE(pl.ok && (FD == 1 || FD == 2), "ich_convT2_tc_wgrad: unsupported shape N%d D%d H%d W%d Cin%d Cout%d FD%d", N, D, H, W, Cin, Cout, FD);
  pl.p.out_mode = 1;
  pl.p.up_cout = Cout;
  pl.p.taps_out = 4 * FD;
  return launch_conv_tc_wgrad(pl, x, x_ld, g, g_ld, dw, (size_t)Cin * Cout * 4 * FD, (cudaStream_t)stream, "ich_convT2_tc_wgrad");
}

// The same weight gradient with the up-sampled gradient `dup` (channel slab of the FINE grid, pitch dup_ld) read in place through the per-tap
// maps -- no re-packed copy.  Needs the dy channel chunk (16 / 32 / 64) to divide Cout.
static bool convT2_wgrad_direct_plan(WPlan& pl, int N, int D, int H, int W, int Cin, int Cout, int FD) {
  pl = make_wplan(N, D, H, W, Cin, 4 * FD * Cout, 1, 1, 1);
  if (!pl.ok) return false;
  int cbk = 8 * pl.p.bu;
  while (cbk > 16 && Cout % cbk) cbk >>= 1;          // a channel chunk must not straddle two taps
  if (Cout % cbk || pl.p.NB % cbk) return false;
  pl.p.bu = cbk / 8;
  return true;
}

int ich_convT2_tc_wgrad_direct_supported(int N, int D, int H, int W, int Cin, int Cout, int FD) {
  if (!get_encode() || (FD != 1 && FD != 2) || Cout % 16) return 0;
  WPlan pl;
  return convT2_wgrad_direct_plan(pl, N, D, H, W, Cin, Cout, FD) ? 1 : 0;
}

int ich_convT2_tc_wgrad_direct(const void* x, int x_ld, const void* dup, int dup_ld, float* dw, int N, int D, int H, int W, int Cin, int Cout, int FD,
                               void* stream) {
  WPlan pl;
  ICH_REQUIRE((FD == 1 || FD == 2) && Cout % 16 == 0 && convT2_wgrad_direct_plan(pl, N, D, H, W, Cin, Cout, FD),
              "ich_convT2_tc_wgrad_direct: unsupported shape N%d D%d H%d W%d Cin%d Cout%d FD%d", N, D, H, W, Cin, Cout, FD);
  pl.p.out_mode = 1;
  pl.p.up_cout = Cout;
  pl.p.taps_out = 4 * FD;
  pl.p.tap_cout = Cout;
  TapMaps tm;
  if (int rc = make_tap_maps(tm, dup, dup_ld, N, D, H, W, Cout, FD, 8 * pl.p.bu, pl.p.WB, pl.p.R, "ich_convT2_tc_wgrad_direct")) return rc;
  return launch_conv_tc_wgrad(pl, x, x_ld, dup, dup_ld, dw, (size_t)Cin * Cout * 4 * FD, (cudaStream_t)stream, "ich_convT2_tc_wgrad_direct", &tm);
}

}  // extern "C"
