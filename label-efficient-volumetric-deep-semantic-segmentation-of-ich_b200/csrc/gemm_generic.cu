// Generic (CUDA-core FFMA, fp32-accumulate) implicit-GEMM kernels.
//
// These are the fp32 verification mode required by the north star ("1e-4 (fp32 mode)", bit-exact masks)
// and the catch-all for shapes the tcgen05 path does not take (Cin = 1 first layer, channel counts that are not
// multiples of 16).  Two tile engines, both templated on an operand functor:
//   fwd-type   C[m][n]  = sum_k A(m,k) * B[k][n]      (conv fwd, conv dgrad, convT fwd/dgrad, 1x1 conv)
//   wgrad-type D[k][n] += sum_m A(m,k) * G(m,n)       (conv wgrad, convT wgrad, 1x1 wgrad), split over m + fp32 atomics
// Replaces the cuDNN calls behind nn.Conv3d / nn.ConvTranspose3d at reference models/networks/UNet.py:75-76,84,153,155.
#include "common.cuh"

namespace {

struct Grid4 {  // voxel grid of the GEMM's M dimension
  int N, D, H, W;
};

struct MVox {  // decoded row index
  int m, r /* n*D+d */, d, h, w;
};

__device__ __forceinline__ MVox decode_m(int m, const Grid4& g) {
  MVox v;
  v.m = m;
  int q = m / g.W;
  v.w = m - q * g.W;
  v.r = q / g.H;
  v.h = q - v.r * g.H;
  v.d = v.r % g.D;
  return v;
}

// ---- A-operand functors ------------------------------------------------------------------------------------
// Implicit im2col of a stride-1 "same" convolution: k = tap*Cin + ci.
template <typename T>
struct AConv {
  const T* x;
  int ld, Cin, KD, KH, KW;
  Grid4 g;
  struct KInfo { int ci, od, oh, ow, off; bool ok; };
  __device__ KInfo kinfo(int k, int Ktot) const {
    KInfo i;
    i.ok = k < Ktot;
    int tap = k / Cin;
    i.ci = k - tap * Cin;
    int kd = tap / (KH * KW);
    int r = tap - kd * KH * KW;
    int kh = r / KW;
    i.od = kd - KD / 2; i.oh = kh - KH / 2; i.ow = (r - kh * KW) - KW / 2;
    i.off = (i.od * g.H + i.oh) * g.W + i.ow;
    return i;
  }
  __device__ float load(const MVox& v, const KInfo& i) const {
    if (!i.ok) return 0.f;
    int d = v.d + i.od, h = v.h + i.oh, w = v.w + i.ow;
    if ((unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W) return 0.f;
    return to_f32(x[(size_t)(v.m + i.off) * ld + i.ci]);
  }
};

// Plain row-major A[m][k].
template <typename T>
struct APlain {
  const T* x;
  int ld;
  struct KInfo { int k; bool ok; };
  __device__ KInfo kinfo(int k, int Ktot) const { return KInfo{k, k < Ktot}; }
  __device__ float load(const MVox& v, const KInfo& i) const { return i.ok ? to_f32(x[(size_t)v.m * ld + i.k]) : 0.f; }
};

// Rows of the fine grid gathered per coarse voxel: k = tap*C + c, tap = (i,j,l) of the 2x2x2 (or 1x2x2) stencil.
// Used for convT dgrad (A = dy) and, as the G operand, for convT wgrad.
template <typename T>
struct AUp {
  const T* y;
  int ld, C, FD;  // FD = depth factor (2 for 3-D, 1 for 2-D)
  Grid4 g;        // coarse grid
  struct KInfo { int c, i, j, l; bool ok; };
  __device__ KInfo kinfo(int k, int Ktot) const {
    KInfo q;
    q.ok = k < Ktot;
    int tap = k / C;
    q.c = k - tap * C;
    q.i = tap >> 2; q.j = (tap >> 1) & 1; q.l = tap & 1;
    return q;
  }
  __device__ size_t fine_row(const MVox& v, int i, int j, int l) const {
    return ((size_t)(v.r * FD + i) * (2 * g.H) + (2 * v.h + j)) * (2 * g.W) + (2 * v.w + l);
  }
  __device__ float load(const MVox& v, const KInfo& q) const {
    return q.ok ? to_f32(y[fine_row(v, q.i, q.j, q.l) * ld + q.c]) : 0.f;
  }
};

// ---- epilogues for the fwd-type engine -------------------------------------------------------------------------
template <typename T>
struct EpiRow {  // y[m][n] = acc + bias[n]
  T* y;
  int ld;
  const float* bias;
  int relu;
  __device__ void store(const MVox& v, int n, float acc) const {
    if (bias) acc += bias[n];
    if (relu) acc = fmaxf(acc, 0.f);
    y[(size_t)v.m * ld + n] = from_f32<T>(acc);
  }
};
template <typename T>
struct EpiUp {  // transposed-conv scatter: n = tap*Cout + co
  T* y;
  int ld, Cout, FD;
  Grid4 g;
  const float* bias;
  __device__ void store(const MVox& v, int n, float acc) const {
    int tap = n / Cout, co = n - tap * Cout;
    int i = tap >> 2, j = (tap >> 1) & 1, l = tap & 1;
    size_t row = ((size_t)(v.r * FD + i) * (2 * g.H) + (2 * v.h + j)) * (2 * g.W) + (2 * v.w + l);
    if (bias) acc += bias[co];
    y[row * ld + co] = from_f32<T>(acc);
  }
};

// ---- fwd-type engine: 128 x BN tile, BK = 16, 256 threads, 8 x BN/16 outputs per thread ----------------------
constexpr int FBM = 128, FBK = 16;

template <int BN, typename AOp, typename Epi>
__global__ void __launch_bounds__(256) gemm_fwd_kernel(AOp A, const float* __restrict__ B, Epi epi, Grid4 g, int M, int Nn,
                                                       int Ktot) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[FBK][FBM + 4];
  __shared__ __align__(16) float Bs[FBK][BN + 4];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.x * FBM, n0 = blockIdx.y * BN;
  const int kl = t & 15, mb = t >> 4;
  MVox vox[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + mb + 16 * i;
    vox[i] = decode_m(m < M ? m : M - 1, g);
    if (m >= M) vox[i].m = -1;
  }
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += FBK) {
    typename AOp::KInfo ki = A.kinfo(k0 + kl, Ktot);
#pragma unroll
    for (int i = 0; i < 8; ++i) As[kl][mb + 16 * i] = vox[i].m >= 0 ? A.load(vox[i], ki) : 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx + 16 * j, k = k0 + ty;
      Bs[ty][tx + 16 * j] = (n < Nn && k < Ktot) ? B[(size_t)k * Nn + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < FBK; ++kk) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue: rows owned in compute layout are ty*8+i, which differ from the load layout -> decode again
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    MVox v = decode_m(m, g);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < Nn) epi.store(v, n, acc[i][j]);
    }
  }
}

template <typename AOp, typename Epi>
int launch_fwd(const AOp& A, const float* B, const Epi& epi, Grid4 g, int Nn, int Ktot, cudaStream_t s, const char* what) {
  long long M = (long long)g.N * g.D * g.H * g.W;
  if (M <= 0 || Nn <= 0) return 0;
  ICH_REQUIRE(M < (1ll << 31), "%s: voxel count %lld exceeds 2^31", what, M);
  dim3 block(256);
  if (Nn <= 16) {
    dim3 grid((unsigned)((M + FBM - 1) / FBM), (Nn + 15) / 16);
    gemm_fwd_kernel<16, AOp, Epi><<<grid, block, 0, s>>>(A, B, epi, g, (int)M, Nn, Ktot);
  } else if (Nn <= 32) {
    dim3 grid((unsigned)((M + FBM - 1) / FBM), (Nn + 31) / 32);
    gemm_fwd_kernel<32, AOp, Epi><<<grid, block, 0, s>>>(A, B, epi, g, (int)M, Nn, Ktot);
  } else {
    dim3 grid((unsigned)((M + FBM - 1) / FBM), (Nn + 63) / 64);
    gemm_fwd_kernel<64, AOp, Epi><<<grid, block, 0, s>>>(A, B, epi, g, (int)M, Nn, Ktot);
  }
  return ich_check_launch(what);
}

// ---- wgrad-type engine: 64(k) x BN(n) tile, 16 rows of m per step, split over m -----------------------------
constexpr int WBK = 64, WBM = 16;

struct OutConvW {  // torch layout [Cout][Cin][taps], k = tap*Cin + ci
  float* dw;
  int Cin, taps;
  __device__ void add(int k, int n, float v) const {
    int tap = k / Cin, ci = k - tap * Cin;
    atomicAdd(&dw[((size_t)n * Cin + ci) * taps + tap], v);
  }
};
struct OutConvTW {  // torch layout [Cin][Cout][taps]; rows k = ci, cols n = tap*Cout + co
  float* dw;
  int Cout, taps;
  __device__ void add(int k, int n, float v) const {
    int tap = n / Cout, co = n - tap * Cout;
    atomicAdd(&dw[((size_t)k * Cout + co) * taps + tap], v);
  }
};

template <int BN, typename AOp, typename GOp, typename Out>
__global__ void __launch_bounds__(256) gemm_wgrad_kernel(AOp A, GOp G, Out out, Grid4 g, int M, int Nn, int Ktot, int m_per_split) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[WBM][WBK + 4];
  __shared__ __align__(16) float Gs[WBM][BN + 4];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int k0 = blockIdx.x * WBK, n0 = blockIdx.y * BN;
  const int m_begin = blockIdx.z * m_per_split;
  const int m_end = min(M, m_begin + m_per_split);
  const int ak = t & 63, am = t >> 6;  // A loads: k fixed per thread, 4 rows
  typename AOp::KInfo ki = A.kinfo(k0 + ak, Ktot);
  typename GOp::KInfo gi[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) gi[j] = G.kinfo(n0 + tx + 16 * j, Nn);
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int mc = m_begin; mc < m_end; mc += WBM) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = mc + am + 4 * i;
      float v = 0.f;
      if (m < m_end) v = A.load(decode_m(m, g), ki);
      As[am + 4 * i][ak] = v;
    }
    {
      int m = mc + ty;
      bool ok = m < m_end;
      MVox vx = decode_m(ok ? m : m_begin, g);
#pragma unroll
      for (int j = 0; j < TN; ++j) Gs[ty][tx + 16 * j] = ok ? G.load(vx, gi[j]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WBM; ++mm) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Gs[mm][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + ty * 4 + i;
    if (k >= Ktot) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < Nn) out.add(k, n, acc[i][j]);
    }
  }
}

template <typename AOp, typename GOp, typename Out>
int launch_wgrad(const AOp& A, const GOp& G, const Out& out, Grid4 g, int Nn, int Ktot, cudaStream_t s, const char* what) {
  long long M = (long long)g.N * g.D * g.H * g.W;
  if (M <= 0 || Nn <= 0 || Ktot <= 0) return 0;
  ICH_REQUIRE(M < (1ll << 31), "%s: voxel count %lld exceeds 2^31", what, M);
  int BN = Nn <= 16 ? 16 : Nn <= 32 ? 32 : 64;
  int tiles = ((Ktot + WBK - 1) / WBK) * ((Nn + BN - 1) / BN);
  int want = 4 * ich_num_sms();
  long long splits = (want + tiles - 1) / tiles;
  long long max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long per = (M + splits - 1) / splits;
  per = (per + WBM - 1) / WBM * WBM;
  splits = (M + per - 1) / per;
  dim3 grid((Ktot + WBK - 1) / WBK, (Nn + BN - 1) / BN, (unsigned)splits), block(256);
  if (BN == 16) gemm_wgrad_kernel<16, AOp, GOp, Out><<<grid, block, 0, s>>>(A, G, out, g, (int)M, Nn, Ktot, (int)per);
  else if (BN == 32) gemm_wgrad_kernel<32, AOp, GOp, Out><<<grid, block, 0, s>>>(A, G, out, g, (int)M, Nn, Ktot, (int)per);
  else gemm_wgrad_kernel<64, AOp, GOp, Out><<<grid, block, 0, s>>>(A, G, out, g, (int)M, Nn, Ktot, (int)per);
  return ich_check_launch(what);
}


// ---- first layer (Cin == 1): bandwidth-bound direct kernels --------------------------------------------------------------
// The network input has one channel (CT intensity): AI ~ 25 FLOP/B, so these are HBM / LSU bound, not tensor bound.
// One block = one row segment (n, d, h, w0 .. w0+127); the KD*3 input rows it touches are staged (zero padded) in shared memory.
constexpr int C1_SEG = 128;

template <typename T, int COUT, int KD>
__global__ void __launch_bounds__(C1_SEG) conv_cin1_fwd_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ wp /*[taps][COUT]*/,
                                                               const float* __restrict__ bias, T* __restrict__ y, int y_ld, Grid4 g, int relu) {
  constexpr int TAPS = KD * 9;
  __shared__ float xs[KD * 3][C1_SEG + 2];
  __shared__ __align__(16) float ws[TAPS][COUT];
  const int nseg = (g.W + C1_SEG - 1) / C1_SEG;
  const int seg = blockIdx.x % nseg;
  long long row = blockIdx.x / nseg;               // (n*D + d)*H + h
  const int h = (int)(row % g.H);
  const long long r = row / g.H;                   // n*D + d
  const int d = (int)(r % g.D);
  const int w0 = seg * C1_SEG;
  for (int i = threadIdx.x; i < TAPS * COUT; i += C1_SEG) ws[i / COUT][i % COUT] = wp[i];
  for (int i = threadIdx.x; i < KD * 3 * (C1_SEG + 2); i += C1_SEG) {
    const int rr = i / (C1_SEG + 2), pw = i - rr * (C1_SEG + 2);
    const int kd = rr / 3, kh = rr - kd * 3;
    const int dd = d + kd - KD / 2, hh = h + kh - 1, ww = w0 + pw - 1;
    float v = 0.f;
    if ((unsigned)dd < (unsigned)g.D && (unsigned)hh < (unsigned)g.H && (unsigned)ww < (unsigned)g.W)
      v = to_f32(x[(((r - d + dd) * g.H + hh) * (long long)g.W + ww) * x_ld]);
    xs[rr][pw] = v;
  }
  __syncthreads();
  const int w = w0 + threadIdx.x;
  if (w >= g.W) return;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = bias ? bias[c] : 0.f;
#pragma unroll
  for (int rr = 0; rr < KD * 3; ++rr)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float xv = xs[rr][threadIdx.x + kw];
#pragma unroll
      for (int c = 0; c < COUT; c += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[rr * 3 + kw][c]);
        acc[c] = fmaf(xv, wv.x, acc[c]); acc[c + 1] = fmaf(xv, wv.y, acc[c + 1]);
        acc[c + 2] = fmaf(xv, wv.z, acc[c + 2]); acc[c + 3] = fmaf(xv, wv.w, acc[c + 3]);
      }
    }
  if (relu)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = fmaxf(acc[c], 0.f);
  T* out = y + (row * g.W + w) * y_ld;
#pragma unroll
  for (int c = 0; c < COUT; c += Vec<T>::N) Vec<T>::store(out + c, acc + c);
}

// dW[co][tap] = sum_vox dy[vox][co] * x[vox + tap].  Block = (co, w-slice) threads; per row the x rows and the dy row are staged in
// shared memory; each thread slides a KD*3 x 3 register window along its w-slice.  Per-block partials -> fp32 atomics.
template <typename T, int COUT, int KD>
__global__ void __launch_bounds__(256) conv_cin1_wgrad_kernel(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld,
                                                              float* __restrict__ dw /*[COUT][taps]*/, Grid4 g, long long rows_total, int rows_per_block) {
  constexpr int TAPS = KD * 9, ROWS = KD * 3, SLICES = 256 / COUT, WPS = C1_SEG / SLICES;   // w positions per slice
  __shared__ float xs[ROWS][C1_SEG + 2];
  __shared__ float ds[C1_SEG][COUT + 1];
  const int co = threadIdx.x % COUT, sl = threadIdx.x / COUT;
  const int nseg = (g.W + C1_SEG - 1) / C1_SEG;
  float acc[TAPS];
#pragma unroll
  for (int k = 0; k < TAPS; ++k) acc[k] = 0.f;
  const long long it0 = (long long)blockIdx.x * rows_per_block;
  const long long it1 = min(rows_total * nseg, it0 + rows_per_block);
  for (long long it = it0; it < it1; ++it) {
    const int seg = (int)(it % nseg);
    const long long row = it / nseg;
    const int h = (int)(row % g.H);
    const long long r = row / g.H;
    const int d = (int)(r % g.D);
    const int w0 = seg * C1_SEG;
    __syncthreads();
    for (int i = threadIdx.x; i < ROWS * (C1_SEG + 2); i += 256) {
      const int rr = i / (C1_SEG + 2), pw = i - rr * (C1_SEG + 2);
      const int kd = rr / 3, kh = rr - kd * 3;
      const int dd = d + kd - KD / 2, hh = h + kh - 1, ww = w0 + pw - 1;
      float v = 0.f;
      if ((unsigned)dd < (unsigned)g.D && (unsigned)hh < (unsigned)g.H && (unsigned)ww < (unsigned)g.W)
        v = to_f32(x[(((r - d + dd) * g.H + hh) * (long long)g.W + ww) * x_ld]);
      xs[rr][pw] = v;
    }
    for (int i = threadIdx.x; i < C1_SEG * COUT; i += 256) {
      const int pw = i / COUT, c = i - pw * COUT;
      const int ww = w0 + pw;
      ds[pw][c] = ww < g.W ? to_f32(dy[(row * g.W + ww) * dy_ld + c]) : 0.f;
    }
    __syncthreads();
    const int p0 = sl * WPS;
    float win[ROWS][3];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) { win[rr][1] = xs[rr][p0]; win[rr][2] = xs[rr][p0 + 1]; }
#pragma unroll 4
    for (int i = 0; i < WPS; ++i) {
      const float a = ds[p0 + i][co];
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        win[rr][0] = win[rr][1]; win[rr][1] = win[rr][2]; win[rr][2] = xs[rr][p0 + i + 2];
        acc[rr * 3 + 0] = fmaf(a, win[rr][0], acc[rr * 3 + 0]);
        acc[rr * 3 + 1] = fmaf(a, win[rr][1], acc[rr * 3 + 1]);
        acc[rr * 3 + 2] = fmaf(a, win[rr][2], acc[rr * 3 + 2]);
      }
    }
  }
  // reduce the SLICES partials of each (co, tap) through shared memory, then one atomic per (co, tap) per block
  __syncthreads();
  float* red = &ds[0][0];     // >= COUT * TAPS floats: 128 * (COUT + 1) >= 27 * COUT
  for (int i = threadIdx.x; i < COUT * TAPS; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < TAPS; ++k) atomicAdd(&red[co * TAPS + k], acc[k]);
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * TAPS; i += 256) atomicAdd(&dw[i], red[i]);
}

template <typename T>
bool cin1_fwd_launch(const T* x, int x_ld, const float* wp, const float* bias, T* y, int y_ld, Grid4 g, int Cout, int KD, int relu, cudaStream_t s) {
  if (y_ld % Vec<T>::N || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
  const int nseg = (g.W + C1_SEG - 1) / C1_SEG;
  long long blocks = (long long)g.N * g.D * g.H * nseg;
  if (blocks <= 0 || blocks > 0x7fffffffLL) return false;
#define ICH_C1F(CO, K) conv_cin1_fwd_kernel<T, CO, K><<<(unsigned)blocks, C1_SEG, 0, s>>>(x, x_ld, wp, bias, y, y_ld, g, relu)
  if (KD == 3) { if (Cout == 8) ICH_C1F(8, 3); else if (Cout == 16) ICH_C1F(16, 3); else if (Cout == 32) ICH_C1F(32, 3); else return false; }
  else { if (Cout == 8) ICH_C1F(8, 1); else if (Cout == 16) ICH_C1F(16, 1); else if (Cout == 32) ICH_C1F(32, 1); else return false; }
#undef ICH_C1F
  return true;
}

template <typename T>
bool cin1_wgrad_launch(const T* x, int x_ld, const T* dy, int dy_ld, float* dw, Grid4 g, int Cout, int KD, cudaStream_t s) {
  const int nseg = (g.W + C1_SEG - 1) / C1_SEG;
  long long rows = (long long)g.N * g.D * g.H, items = rows * nseg;
  if (items <= 0) return false;
  int rows_per_block = (int)((items + 8LL * ich_num_sms() - 1) / (8LL * ich_num_sms()));
  if (rows_per_block < 1) rows_per_block = 1;
  unsigned blocks = (unsigned)((items + rows_per_block - 1) / rows_per_block);
#define ICH_C1W(CO, K) conv_cin1_wgrad_kernel<T, CO, K><<<blocks, 256, 0, s>>>(x, x_ld, dy, dy_ld, dw, g, rows, rows_per_block)
  if (KD == 3) { if (Cout == 8) ICH_C1W(8, 3); else if (Cout == 16) ICH_C1W(16, 3); else if (Cout == 32) ICH_C1W(32, 3); else return false; }
  else { if (Cout == 8) ICH_C1W(8, 1); else if (Cout == 16) ICH_C1W(16, 1); else if (Cout == 32) ICH_C1W(32, 1); else return false; }
#undef ICH_C1W
  return true;
}

}  // namespace

// ---- C-ABI (declared in include/ich_b200.h) ---------------------------------------------------------------------
extern "C" {

int ich_conv_fwd(const void* x, int x_ld, const float* wpack, const float* bias, void* y, int y_ld, int dtype, int N, int D,
                 int H, int W, int Cin, int Cout, int KD, int KH, int KW, int relu, void* stream) {
  Grid4 g{N, D, H, W};
  cudaStream_t s = (cudaStream_t)stream;
  int Ktot = KD * KH * KW * Cin;
  ICH_REQUIRE((KD & 1) && (KH & 1) && (KW & 1), "ich_conv_fwd: odd kernel sizes only (got %dx%dx%d)", KD, KH, KW);
  if (Cin == 1 && KH == 3 && KW == 3 && (KD == 1 || KD == 3)) {   // first layer: direct bandwidth-bound kernel
    bool done = dtype == ICH_F32 ? cin1_fwd_launch<float>((const float*)x, x_ld, wpack, bias, (float*)y, y_ld, g, Cout, KD, relu, s)
              : dtype == ICH_BF16 ? cin1_fwd_launch<bf16>((const bf16*)x, x_ld, wpack, bias, (bf16*)y, y_ld, g, Cout, KD, relu, s) : false;
    if (done) return ich_check_launch("ich_conv_fwd<cin1>");
  }
  if (dtype == ICH_F32) {
    AConv<float> A{(const float*)x, x_ld, Cin, KD, KH, KW, g};
    EpiRow<float> E{(float*)y, y_ld, bias, relu};
    return launch_fwd(A, wpack, E, g, Cout, Ktot, s, "ich_conv_fwd<f32>");
  } else if (dtype == ICH_BF16) {
    AConv<bf16> A{(const bf16*)x, x_ld, Cin, KD, KH, KW, g};
    EpiRow<bf16> E{(bf16*)y, y_ld, bias, relu};
    return launch_fwd(A, wpack, E, g, Cout, Ktot, s, "ich_conv_fwd<bf16>");
  }
  ICH_REQUIRE(false, "ich_conv_fwd: bad dtype %d", dtype);
}

int ich_conv_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, int dtype, float* dw, int N, int D, int H, int W, int Cin,
                   int Cout, int KD, int KH, int KW, void* stream) {
  Grid4 g{N, D, H, W};
  cudaStream_t s = (cudaStream_t)stream;
  int taps = KD * KH * KW, Ktot = taps * Cin;
  if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Ktot, s) != cudaSuccess) return ich_check_launch("ich_conv_wgrad memset");
  if (Cin == 1 && KH == 3 && KW == 3 && (KD == 1 || KD == 3)) {   // first layer: dW[co][0][tap] == [Cout][taps]
    bool done = dtype == ICH_F32 ? cin1_wgrad_launch<float>((const float*)x, x_ld, (const float*)dy, dy_ld, dw, g, Cout, KD, s)
              : dtype == ICH_BF16 ? cin1_wgrad_launch<bf16>((const bf16*)x, x_ld, (const bf16*)dy, dy_ld, dw, g, Cout, KD, s) : false;
    if (done) return ich_check_launch("ich_conv_wgrad<cin1>");
  }
  OutConvW O{dw, Cin, taps};
  if (dtype == ICH_F32) {
    AConv<float> A{(const float*)x, x_ld, Cin, KD, KH, KW, g};
    APlain<float> G{(const float*)dy, dy_ld};
    return launch_wgrad(A, G, O, g, Cout, Ktot, s, "ich_conv_wgrad<f32>");
  } else if (dtype == ICH_BF16) {
    AConv<bf16> A{(const bf16*)x, x_ld, Cin, KD, KH, KW, g};
    APlain<bf16> G{(const bf16*)dy, dy_ld};
    return launch_wgrad(A, G, O, g, Cout, Ktot, s, "ich_conv_wgrad<bf16>");
  }
  ICH_REQUIRE(false, "ich_conv_wgrad: bad dtype %d", dtype);
}

// Transposed conv, kernel 2 stride 2 (depth factor FD = 2, or 1 for the 2-D nets). Grid args = the COARSE (input) grid.
int ich_convT2_fwd(const void* x, int x_ld, const float* wpack /*[Cin][taps*Cout]*/, const float* bias, void* y, int y_ld,
                   int dtype, int N, int D, int H, int W, int Cin, int Cout, int FD, void* stream) {
  Grid4 g{N, D, H, W};
  cudaStream_t s = (cudaStream_t)stream;
  int taps = FD * 4;
  ICH_REQUIRE(FD == 1 || FD == 2, "ich_convT2_fwd: FD must be 1 or 2");
  if (dtype == ICH_F32) {
    APlain<float> A{(const float*)x, x_ld};
    EpiUp<float> E{(float*)y, y_ld, Cout, FD, g, bias};
    return launch_fwd(A, wpack, E, g, taps * Cout, Cin, s, "ich_convT2_fwd<f32>");
  } else if (dtype == ICH_BF16) {
    APlain<bf16> A{(const bf16*)x, x_ld};
    EpiUp<bf16> E{(bf16*)y, y_ld, Cout, FD, g, bias};
    return launch_fwd(A, wpack, E, g, taps * Cout, Cin, s, "ich_convT2_fwd<bf16>");
  }
  ICH_REQUIRE(false, "ich_convT2_fwd: bad dtype %d", dtype);
}

int ich_convT2_dgrad(const void* dy, int dy_ld, const float* wpack_d /*[taps*Cout][Cin]*/, void* dx, int dx_ld, int dtype, int N,
                     int D, int H, int W, int Cin, int Cout, int FD, void* stream) {
  Grid4 g{N, D, H, W};
  cudaStream_t s = (cudaStream_t)stream;
  int taps = FD * 4;
  if (dtype == ICH_F32) {
    AUp<float> A{(const float*)dy, dy_ld, Cout, FD, g};
    EpiRow<float> E{(float*)dx, dx_ld, nullptr, 0};
    return launch_fwd(A, wpack_d, E, g, Cin, taps * Cout, s, "ich_convT2_dgrad<f32>");
  } else if (dtype == ICH_BF16) {
    AUp<bf16> A{(const bf16*)dy, dy_ld, Cout, FD, g};
    EpiRow<bf16> E{(bf16*)dx, dx_ld, nullptr, 0};
    return launch_fwd(A, wpack_d, E, g, Cin, taps * Cout, s, "ich_convT2_dgrad<bf16>");
  }
  ICH_REQUIRE(false, "ich_convT2_dgrad: bad dtype %d", dtype);
}

int ich_convT2_wgrad(const void* x, int x_ld, const void* dy, int dy_ld, int dtype, float* dw /*[Cin][Cout][taps]*/, int N, int D,
                     int H, int W, int Cin, int Cout, int FD, void* stream) {
  Grid4 g{N, D, H, W};
  cudaStream_t s = (cudaStream_t)stream;
  int taps = FD * 4;
  if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cin * Cout * taps, s) != cudaSuccess)
    return ich_check_launch("ich_convT2_wgrad memset");
  OutConvTW O{dw, Cout, taps};
  if (dtype == ICH_F32) {
    APlain<float> A{(const float*)x, x_ld};
    AUp<float> G{(const float*)dy, dy_ld, Cout, FD, g};
    return launch_wgrad(A, G, O, g, taps * Cout, Cin, s, "ich_convT2_wgrad<f32>");
  } else if (dtype == ICH_BF16) {
    APlain<bf16> A{(const bf16*)x, x_ld};
    AUp<bf16> G{(const bf16*)dy, dy_ld, Cout, FD, g};
    return launch_wgrad(A, G, O, g, taps * Cout, Cin, s, "ich_convT2_wgrad<bf16>");
  }
  ICH_REQUIRE(false, "ich_convT2_wgrad: bad dtype %d", dtype);
}

}  // extern "C"
