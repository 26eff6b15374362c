// Bandwidth-bound kernels of the U-Net block: layout conversion, BatchNorm statistics / finalize / apply(+ReLU)
// forward and backward, 2x max-pool forward / backward, channel-slab copies (skip concat), global average pool.
// All activations are channel-last rows [M][C] with an explicit channel pitch `ld` so a tensor can be a channel slab
// of a wider concat buffer.  Vector path: 16-byte accesses (8 bf16 / 4 fp32) when C, ld and the base are aligned.
// Replaces ATen batch_norm / relu / max_pool3d / cat behind reference models/networks/UNet.py:82,119,149,154-161.
#include "common.cuh"
#include <stdlib.h>

namespace {

template <typename T>
__host__ __device__ inline bool vec_ok(const void* p, int ld, int C) {
  return (C % Vec<T>::N == 0) && (ld % Vec<T>::N == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

// The BN-backward partial sums are accumulated with fp64 atomics into BN_COPIES replicas of the [2][C] vector (replica = block %
// BN_COPIES): ~1200 blocks finishing together on the same 2*C addresses cost 50-100 us per launch in serialised atomics.
constexpr int BN_COPIES = 16;
__device__ __forceinline__ double bn_total(const double* __restrict__ sums, int i, int C) {
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < BN_COPIES; ++k) t += sums[(size_t)k * 2 * C + i];
  return t;
}

// ---- fused dropout (nn.Dropout after the second ReLU of a ConvBlock, reference models/networks/UNet.py:150,175-176) -------------
// Counter-based: the keep decision of element (row r, channel c) is a pure function of (seed, r, c), so the forward kernel and the
// two backward kernels regenerate the same mask and no mask tensor is ever stored.  One Philox4x32-10 block yields eight 16-bit
// uniforms = the eight channels c & ~7 .. (c & ~7) + 7 of a row; keep iff u16 >= thr, thr = round(p * 65536), kept values are
// scaled by 65536 / (65536 - thr).  thr == 0 disables the whole thing (uniform branch).
struct Drop {
  uint32_t thr;
  float scale;
  uint32_t seed_lo, seed_hi;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return ctr;
}

// keep multipliers (0 or scale) of the V channels c .. c + V - 1 of row r  (V in {1, 4, 8}, c % V == 0)
template <int V>
__device__ __forceinline__ void drop_mult(const Drop& d, long long r, int c, float* m) {
  const uint4 u = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)((unsigned long long)r >> 32), (uint32_t)(c >> 3), 0x1C4B200u), d.seed_lo, d.seed_hi);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int h = (c & 7) + k;                       // half-word index 0..7
    const uint32_t v16 = (w[h >> 1] >> ((h & 1) * 16)) & 0xffffu;
    m[k] = v16 >= d.thr ? d.scale : 0.f;
  }
}

inline Drop make_drop(float p, unsigned long long seed) {
  Drop d{0u, 1.f, (uint32_t)seed, (uint32_t)(seed >> 32)};
  if (p > 0.f) {
    double t = (double)p * 65536.0 + 0.5;
    d.thr = t >= 65536.0 ? 65536u : (uint32_t)t;
    if (d.thr == 0) d.thr = 1;                       // p > 0 below the 16-bit resolution: smallest representable rate
    d.scale = d.thr >= 65536u ? 0.f : (float)(65536.0 / (65536.0 - (double)d.thr));
  }
  return d;
}

inline int grid_for(long long work, int threads, int per_sm = 8) {
  long long blocks = (work + threads - 1) / threads;
  long long cap = (long long)ich_num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- layout: NC(S) fp32 <-> N(S)C T --------------------------------------------------------------------------
template <typename T>
__global__ void nc_to_nl_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, long long S, int ld) {
  __shared__ float tile[32][33];
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const long long n = blockIdx.z;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && s < S) ? src[(n * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long s = s0 + i;
    int c = c0 + threadIdx.x;
    if (c < C && s < S) dst[(n * S + s) * ld + c] = from_f32<T>(tile[threadIdx.x][i]);
  }
}
template <typename T>
__global__ void nl_to_nc_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, long long S, int ld) {
  __shared__ float tile[32][33];
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const long long n = blockIdx.z;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long s = s0 + i;
    int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && s < S) ? to_f32(src[(n * S + s) * ld + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long s = s0 + threadIdx.x;
    if (c < C && s < S) dst[(n * C + c) * S + s] = tile[threadIdx.x][i];
  }
}
// C == 1 fast path: a pure cast, fully coalesced.
template <typename T>
__global__ void cast_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f32<T>(src[i]);
}
template <typename T>
__global__ void cast_to_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = to_f32(src[i]);
}

// ---- per-channel statistics over rows: sum and sum of squares -------------------------------------------------
// Block = 256 threads; thread owns one 16-byte channel group and strides over rows; fp32 partials per thread (<= a few
// thousand adds), block tree in shared memory, one fp64 atomic per channel per block.
template <typename T, bool SQ>
__global__ void __launch_bounds__(256) colstats_vec_kernel(const T* __restrict__ x, int ld, long long M, int C, double* __restrict__ sum,
                                                           double* __restrict__ sumsq, int rows_per_block) {
  constexpr int V = Vec<T>::N;
  const int groups = C / V;                    // channel groups (<= 256 guaranteed by the launcher)
  const int lanes = 256 / groups;              // row lanes
  const int gidx = threadIdx.x % groups, lane = threadIdx.x / groups;
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = q[i] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  if (lane < lanes) {
    for (long long r = r0 + lane; r < r1; r += lanes) {
      float v[V];
      Vec<T>::load(x + r * ld + gidx * V, v);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s[i] += v[i];
        if (SQ) q[i] = fmaf(v[i], v[i], q[i]);
      }
    }
  }
  __shared__ float sh[2][256][V + 1];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sh[0][threadIdx.x][i] = s[i];
    if (SQ) sh[1][threadIdx.x][i] = q[i];
  }
  __syncthreads();
  // thread t < C reduces channel t over the row lanes
  for (int c = threadIdx.x; c < C; c += 256) {
    int g = c / V, i = c % V;
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a += sh[0][l * groups + g][i];
      if (SQ) b += sh[1][l * groups + g][i];
    }
    atomicAdd(&sum[c], (double)a);
    if (SQ) atomicAdd(&sumsq[c], (double)b);
  }
}
template <typename T, bool SQ>
__global__ void __launch_bounds__(256) colstats_scalar_kernel(const T* __restrict__ x, int ld, long long M, int C, double* __restrict__ sum,
                                                              double* __restrict__ sumsq, int rows_per_block) {
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  for (int c = threadIdx.x; c < C; c += 256) {   // uncoalesced but correct for any C / alignment
    float a = 0.f, b = 0.f;
    for (long long r = r0; r < r1; ++r) {
      float v = to_f32(x[r * ld + c]);
      a += v;
      if (SQ) b = fmaf(v, v, b);
    }
    atomicAdd(&sum[c], (double)a);
    if (SQ) atomicAdd(&sumsq[c], (double)b);
  }
}

template <typename T>
int colstats_launch(const T* x, int ld, long long M, int C, double* sum, double* sumsq, cudaStream_t s) {
  if (M <= 0 || C <= 0) return 0;
  int rows_per_block = 2048;
  long long blocks = (M + rows_per_block - 1) / rows_per_block;
  bool vec = vec_ok<T>(x, ld, C) && (C / Vec<T>::N <= 256) && (256 % (C / Vec<T>::N) == 0);
  if (vec) {
    if (sumsq) colstats_vec_kernel<T, true><<<(unsigned)blocks, 256, 0, s>>>(x, ld, M, C, sum, sumsq, rows_per_block);
    else colstats_vec_kernel<T, false><<<(unsigned)blocks, 256, 0, s>>>(x, ld, M, C, sum, sumsq, rows_per_block);
  } else {
    if (sumsq) colstats_scalar_kernel<T, true><<<(unsigned)blocks, 256, 0, s>>>(x, ld, M, C, sum, sumsq, rows_per_block);
    else colstats_scalar_kernel<T, false><<<(unsigned)blocks, 256, 0, s>>>(x, ld, M, C, sum, sumsq, rows_per_block);
  }
  return ich_check_launch("ich_colstats");
}

// ---- BatchNorm finalize ------------------------------------------------------------------------------------------
// training: batch mean / biased var from the fp64 sums -> scale, shift, saved mean / invstd; running stats updated with
// momentum and the UNBIASED variance (torch semantics).  The conv bias is not added by the conv kernel in training mode
// (BatchNorm cancels it), so it is folded into the running-mean update here.
// eval: scale/shift from the running stats, conv bias folded into the shift.
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd, int training) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  float cb = conv_bias ? conv_bias[c] : 0.f;
  if (training) {
    double mean = sum[c] / count;
    double var = sumsq[c] / count - mean * mean;
    if (var < 0) var = 0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    save_mean[c] = (float)mean;
    save_invstd[c] = invstd;
    if (running_mean) {
      double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)mean + cb);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    float invstd = 1.f / sqrtf(running_var[c] + eps);
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b + (cb - running_mean[c]) * sc;
    save_mean[c] = running_mean[c] - cb;
    save_invstd[c] = invstd;
  }
}

// ---- z = [relu](y * scale[c] + shift[c]) --------------------------------------------------------------------------
template <typename T, bool VEC, bool DROP>
__global__ void __launch_bounds__(256) affine_act_kernel(const T* __restrict__ y, int y_ld, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, T* __restrict__ z, int z_ld, long long M, int C,
                                                         int relu, const Drop drop) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / groups;
    int c = (int)(i - r * groups) * V;
    float v[V];
    if (VEC) Vec<T>::load(y + r * y_ld + c, v); else v[0] = to_f32(y[r * y_ld + c]);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float t = fmaf(v[k], scale[c + k], shift[c + k]);
      v[k] = relu ? fmaxf(t, 0.f) : t;
    }
    if (DROP) {
      float m[V];
      drop_mult<V>(drop, r, c, m);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] *= m[k];
    }
    if (VEC) Vec<T>::store(z + r * z_ld + c, v); else z[r * z_ld + c] = from_f32<T>(v[0]);
  }
}

// ---- BatchNorm(+ReLU) backward ------------------------------------------------------------------------------------
// pass 1: dbeta[c] = sum g, dgamma[c] = sum g * xhat, g = dz * [y*scale+shift > 0]
template <typename T, bool VEC, bool DROP>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* __restrict__ dz, int dz_ld, const T* __restrict__ y, int y_ld,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd, long long M,
                                                            int C, int relu, double* __restrict__ sums /*[2][C]*/, int rows_per_block, const Drop drop) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V;
  const int lanes = max(1, 256 / groups);
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  extern __shared__ float sh[];  // [2][C]
  for (int c = threadIdx.x; c < 2 * C; c += 256) sh[c] = 0.f;
  __syncthreads();
  for (int gi = threadIdx.x % 256; gi < groups * lanes; gi += 256) {
    const int g = gi % groups, lane = gi / groups, c = g * V;
    float sb[V], sg[V], sc[V], sf[V], mu[V], is[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      sb[k] = sg[k] = 0.f;
      sc[k] = scale[c + k]; sf[k] = shift[c + k]; mu[k] = mean[c + k]; is[k] = invstd[c + k];
    }
    for (long long r = r0 + lane; r < r1; r += lanes) {
      float a[V], b[V];
      if (VEC) { Vec<T>::load(dz + r * dz_ld + c, a); Vec<T>::load(y + r * y_ld + c, b); }
      else { a[0] = to_f32(dz[r * dz_ld + c]); b[0] = to_f32(y[r * y_ld + c]); }
      if (DROP) {
        float m[V];
        drop_mult<V>(drop, r, c, m);
#pragma unroll
        for (int k = 0; k < V; ++k) a[k] *= m[k];
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float gk = (!relu || fmaf(b[k], sc[k], sf[k]) > 0.f) ? a[k] : 0.f;
        sb[k] += gk;
        sg[k] = fmaf(gk, (b[k] - mu[k]) * is[k], sg[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      atomicAdd(&sh[c + k], sb[k]);
      atomicAdd(&sh[C + c + k], sg[k]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) atomicAdd(&sums[(size_t)(blockIdx.x % BN_COPIES) * 2 * C + c], (double)sh[c]);
}

// pass 2: dy = scale * (g - dbeta/M - xhat * dgamma/M)   (training) ;  dy = scale * g  (eval)
template <typename T, bool VEC, bool DROP>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ dz, int dz_ld, const T* __restrict__ y, int y_ld,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const double* __restrict__ sums, T* __restrict__ dy, int dy_ld, long long M, int C,
                                                           int relu, int training, float* __restrict__ dgamma, float* __restrict__ dbeta, const Drop drop,
                                                           long long count) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V;
  const long long total = M * groups;
  const float invM = training ? (float)(1.0 / (double)count) : 0.f;   // count = rows behind the statistics (all ranks with SyncBN)
  extern __shared__ float tot[];   // [2][C]: the replicas summed once per block
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) tot[c] = (float)bn_total(sums, c, C);
  __syncthreads();
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = tot[c];
      if (dgamma) dgamma[c] = tot[C + c];
    }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / groups;
    int c = (int)(i - r * groups) * V;
    float a[V], b[V], o[V];
    if (VEC) { Vec<T>::load(dz + r * dz_ld + c, a); Vec<T>::load(y + r * y_ld + c, b); }
    else { a[0] = to_f32(dz[r * dz_ld + c]); b[0] = to_f32(y[r * y_ld + c]); }
    if (DROP) {
      float m[V];
      drop_mult<V>(drop, r, c, m);
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] *= m[k];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float sc = scale[c + k];
      float gk = (!relu || fmaf(b[k], sc, shift[c + k]) > 0.f) ? a[k] : 0.f;
      float xhat = (b[k] - mean[c + k]) * invstd[c + k];
      o[k] = sc * (gk - tot[c + k] * invM - xhat * tot[C + c + k] * invM);
    }
    if (VEC) Vec<T>::store(dy + r * dy_ld + c, o); else dy[r * dy_ld + c] = from_f32<T>(o[0]);
  }
}


// ---- fast paths: C/V divides 256, so a thread owns ONE 16-byte channel group for the whole kernel (per-channel constants live
//      in registers, no index division in the loop) and strides over rows with several independent loads in flight. ---------
template <typename T, bool DROP>
__global__ void __launch_bounds__(256) affine_act_rows_kernel(const T* __restrict__ y, int y_ld, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, T* __restrict__ z, int z_ld, long long M, int C, int relu,
                                                              const Drop drop) {
  constexpr int V = Vec<T>::N;
  const int groups = C / V, rpb = 256 / groups;
  const int c = (threadIdx.x % groups) * V;
  float sc[V], sh[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { sc[k] = scale[c + k]; sh[k] = shift[c + k]; }
  const long long step = (long long)gridDim.x * rpb;
  long long r = (long long)blockIdx.x * rpb + threadIdx.x / groups;
  for (; r + step < M; r += 2 * step) {
    float a[V], b[V];
    Vec<T>::load(y + r * y_ld + c, a);
    Vec<T>::load(y + (r + step) * y_ld + c, b);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      a[k] = fmaf(a[k], sc[k], sh[k]); b[k] = fmaf(b[k], sc[k], sh[k]);
      if (relu) { a[k] = fmaxf(a[k], 0.f); b[k] = fmaxf(b[k], 0.f); }
    }
    if (DROP) {
      float m0[V], m1[V];
      drop_mult<V>(drop, r, c, m0); drop_mult<V>(drop, r + step, c, m1);
#pragma unroll
      for (int k = 0; k < V; ++k) { a[k] *= m0[k]; b[k] *= m1[k]; }
    }
    Vec<T>::store(z + r * z_ld + c, a);
    Vec<T>::store(z + (r + step) * z_ld + c, b);
  }
  if (r < M) {
    float a[V];
    Vec<T>::load(y + r * y_ld + c, a);
#pragma unroll
    for (int k = 0; k < V; ++k) { a[k] = fmaf(a[k], sc[k], sh[k]); if (relu) a[k] = fmaxf(a[k], 0.f); }
    if (DROP) {
      float m0[V];
      drop_mult<V>(drop, r, c, m0);
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] *= m0[k];
    }
    Vec<T>::store(z + r * z_ld + c, a);
  }
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(256) bn_bwd_reduce_rows_kernel(const T* __restrict__ dz, int dz_ld, const T* __restrict__ y, int y_ld,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd, long long M, int C,
                                                                 int relu, double* __restrict__ sums, const Drop drop) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 4;                       // rows in flight per thread: 8 independent 16-byte loads
  const int groups = C / V, rpb = 256 / groups;
  const int c = (threadIdx.x % groups) * V;
  float sc[V], sf[V], mu[V], is[V], sb[V], sg[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { sc[k] = scale[c + k]; sf[k] = shift[c + k]; mu[k] = mean[c + k]; is[k] = invstd[c + k]; sb[k] = sg[k] = 0.f; }
  const long long step = (long long)gridDim.x * rpb;
  long long r = (long long)blockIdx.x * rpb + threadIdx.x / groups;
  for (; r + (U - 1) * step < M; r += U * step) {
    float a[U][V], b[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) { Vec<T>::load(dz + (r + u * step) * dz_ld + c, a[u]); Vec<T>::load(y + (r + u * step) * y_ld + c, b[u]); }
    if (DROP) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float m[V];
        drop_mult<V>(drop, r + u * step, c, m);
#pragma unroll
        for (int k = 0; k < V; ++k) a[u][k] *= m[k];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float g0 = (!relu || fmaf(b[u][k], sc[k], sf[k]) > 0.f) ? a[u][k] : 0.f;
        sb[k] += g0;
        sg[k] = fmaf(g0, (b[u][k] - mu[k]) * is[k], sg[k]);
      }
  }
  for (; r < M; r += step) {
    float a0[V], b0[V];
    Vec<T>::load(dz + r * dz_ld + c, a0); Vec<T>::load(y + r * y_ld + c, b0);
    if (DROP) {
      float m[V];
      drop_mult<V>(drop, r, c, m);
#pragma unroll
      for (int k = 0; k < V; ++k) a0[k] *= m[k];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float g0 = (!relu || fmaf(b0[k], sc[k], sf[k]) > 0.f) ? a0[k] : 0.f;
      sb[k] += g0;
      sg[k] = fmaf(g0, (b0[k] - mu[k]) * is[k], sg[k]);
    }
  }
  // block reduction without shared-memory atomics (64-way contended float atomics cost ~25 us per launch): lanes that own the same
  // channel group are `groups` apart -> butterfly over those lanes, one row of partials per warp, then a sum over the 8 warps
  extern __shared__ float sh[];  // [8 warps][2][C]
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    for (int o = groups; o < 32; o <<= 1) {
      sb[k] += __shfl_xor_sync(0xffffffffu, sb[k], o);
      sg[k] += __shfl_xor_sync(0xffffffffu, sg[k], o);
    }
  }
  if (groups >= 32 || lane < groups) {
    // groups >= 32: every lane owns a distinct channel group and a warp covers only 32 of the groups; the final sum below reads
    // exactly the warps that own a channel (the other slots are never written nor read)
#pragma unroll
    for (int k = 0; k < V; ++k) { sh[(wrp * 2) * C + c + k] = sb[k]; sh[(wrp * 2 + 1) * C + c + k] = sg[k]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float t = 0.f;
    if (groups >= 32) {
      // channel i % C belongs to group (i % C) / V, owned by the warps w with (w * 32 + lane) % groups == group
      const int g = (i % C) / V;
      for (int w = 0; w < 8; ++w)
        if (((w * 32) % groups) <= g && g < ((w * 32) % groups) + 32) t += sh[(w * 2 + i / C) * C + (i % C)];
    } else {
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sh[(w * 2 + i / C) * C + (i % C)];
    }
    atomicAdd(&sums[(size_t)(blockIdx.x % BN_COPIES) * 2 * C + i], (double)t);
  }
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(256) bn_bwd_apply_rows_kernel(const T* __restrict__ dz, int dz_ld, const T* __restrict__ y, int y_ld,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const double* __restrict__ sums, T* __restrict__ dy, int dy_ld, long long M, int C,
                                                                int relu, int training, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                const Drop drop, long long count) {
  constexpr int V = Vec<T>::N;
  const int groups = C / V, rpb = 256 / groups;
  const int c = (threadIdx.x % groups) * V;
  const float invM = training ? (float)(1.0 / (double)count) : 0.f;   // count = rows behind the statistics (all ranks with SyncBN)
  extern __shared__ float tot[];   // [2][C]: the replicas summed once per block
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) tot[i] = (float)bn_total(sums, i, C);
  __syncthreads();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      if (dbeta) dbeta[i] = tot[i];
      if (dgamma) dgamma[i] = tot[C + i];
    }
  // dy = sc * g - k0 - y * k1  with  k1 = sc * is * dgamma/M,  k0 = sc * dbeta/M - mu * k1
  float sc[V], sf[V], k0[V], k1[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    sc[k] = scale[c + k]; sf[k] = shift[c + k];
    k1[k] = sc[k] * invstd[c + k] * tot[C + c + k] * invM;
    k0[k] = sc[k] * tot[c + k] * invM - mean[c + k] * k1[k];
  }
  const long long step = (long long)gridDim.x * rpb;
  long long r = (long long)blockIdx.x * rpb + threadIdx.x / groups;
  for (; r + step < M; r += 2 * step) {
    float a0[V], b0[V], a1[V], b1[V];
    Vec<T>::load(dz + r * dz_ld + c, a0); Vec<T>::load(y + r * y_ld + c, b0);
    Vec<T>::load(dz + (r + step) * dz_ld + c, a1); Vec<T>::load(y + (r + step) * y_ld + c, b1);
    if (DROP) {
      float m0[V], m1[V];
      drop_mult<V>(drop, r, c, m0); drop_mult<V>(drop, r + step, c, m1);
#pragma unroll
      for (int k = 0; k < V; ++k) { a0[k] *= m0[k]; a1[k] *= m1[k]; }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float g0 = (!relu || fmaf(b0[k], sc[k], sf[k]) > 0.f) ? a0[k] : 0.f;
      float g1 = (!relu || fmaf(b1[k], sc[k], sf[k]) > 0.f) ? a1[k] : 0.f;
      a0[k] = fmaf(sc[k], g0, -fmaf(b0[k], k1[k], k0[k]));
      a1[k] = fmaf(sc[k], g1, -fmaf(b1[k], k1[k], k0[k]));
    }
    Vec<T>::store(dy + r * dy_ld + c, a0);
    Vec<T>::store(dy + (r + step) * dy_ld + c, a1);
  }
  if (r < M) {
    float a0[V], b0[V];
    Vec<T>::load(dz + r * dz_ld + c, a0); Vec<T>::load(y + r * y_ld + c, b0);
    if (DROP) {
      float m0[V];
      drop_mult<V>(drop, r, c, m0);
#pragma unroll
      for (int k = 0; k < V; ++k) a0[k] *= m0[k];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float g0 = (!relu || fmaf(b0[k], sc[k], sf[k]) > 0.f) ? a0[k] : 0.f;
      a0[k] = fmaf(sc[k], g0, -fmaf(b0[k], k1[k], k0[k]));
    }
    Vec<T>::store(dy + r * dy_ld + c, a0);
  }
}

inline bool rows_fast_ok(int C, int V) { return C % V == 0 && (C / V) <= 256 && 256 % (C / V) == 0; }
inline int rows_grid(long long M, int C, int V) {
  const int rpb = 256 / (C / V);
  long long blocks = (M + rpb - 1) / rpb;
  long long cap = (long long)ich_num_sms() * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ---- 2x max-pool (kernel 2 stride 2; depth factor FD = 2 or 1) ------------------------------------------------------
// Tie rule = ATen max_pool3d_with_indices: scan (d,h,w) in order, update on strict '>' (NaN propagates) -> first max wins.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int N, int D,
                                                          int H, int W, int C, int FD, T* __restrict__ skip, int skip_ld) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V, Do = D / FD, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Do * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long o = i / groups;
    int c = (int)(i - o * groups) * V;
    int wo = (int)(o % Wo); long long t = o / Wo;
    int ho = (int)(t % Ho); t /= Ho;
    int dd = (int)(t % Do); int n = (int)(t / Do);
    float best[V];
#pragma unroll
    for (int k = 0; k < V; ++k) best[k] = -INFINITY;
    for (int a = 0; a < FD; ++a)
      for (int b = 0; b < 2; ++b)
        for (int e = 0; e < 2; ++e) {
          long long row = (((long long)n * D + dd * FD + a) * H + 2 * ho + b) * W + 2 * wo + e;
          float v[V];
          if (VEC) Vec<T>::load(x + row * x_ld + c, v); else v[0] = to_f32(x[row * x_ld + c]);
          if (skip) {   // the same tensor is the skip connection: lay it out as a channel slab of the decoder's concat buffer on the way
            if (VEC) *reinterpret_cast<uint4*>(skip + row * skip_ld + c) = *reinterpret_cast<const uint4*>(x + row * x_ld + c);
            else skip[row * skip_ld + c] = x[row * x_ld + c];
          }
#pragma unroll
          for (int k = 0; k < V; ++k) if (v[k] > best[k] || v[k] != v[k]) best[k] = v[k];
        }
    if (VEC) Vec<T>::store(y + o * y_ld + c, best); else y[o * y_ld + c] = from_f32<T>(best[0]);
  }
}
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld,
                                                          T* __restrict__ dx, int dx_ld, int N, int D, int H, int W, int C, int FD,
                                                          const T* __restrict__ dskip, int ds_ld) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V, Do = D / FD, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Do * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long o = i / groups;
    int c = (int)(i - o * groups) * V;
    int wo = (int)(o % Wo); long long t = o / Wo;
    int ho = (int)(t % Ho); t /= Ho;
    int dd = (int)(t % Do); int n = (int)(t / Do);
    float best[V], g[V];
    int arg[V];
    if (VEC) Vec<T>::load(dy + o * dy_ld + c, g); else g[0] = to_f32(dy[o * dy_ld + c]);
#pragma unroll
    for (int k = 0; k < V; ++k) { best[k] = -INFINITY; arg[k] = 0; }
    for (int a = 0; a < FD; ++a)
      for (int b = 0; b < 2; ++b)
        for (int e = 0; e < 2; ++e) {
          long long row = (((long long)n * D + dd * FD + a) * H + 2 * ho + b) * W + 2 * wo + e;
          float v[V];
          if (VEC) Vec<T>::load(x + row * x_ld + c, v); else v[0] = to_f32(x[row * x_ld + c]);
#pragma unroll
          for (int k = 0; k < V; ++k) if (v[k] > best[k] || v[k] != v[k]) { best[k] = v[k]; arg[k] = (a << 2) | (b << 1) | e; }
        }
    for (int a = 0; a < FD; ++a)
      for (int b = 0; b < 2; ++b)
        for (int e = 0; e < 2; ++e) {
          long long row = (((long long)n * D + dd * FD + a) * H + 2 * ho + b) * W + 2 * wo + e;
          float v[V];
#pragma unroll
          for (int k = 0; k < V; ++k) v[k] = (arg[k] == ((a << 2) | (b << 1) | e)) ? g[k] : 0.f;
          if (dskip) {   // the pooled tensor is also a skip connection: add the gradient arriving through the decoder (fp32 add, one rounding)
            float sk[V];
            if (VEC) Vec<T>::load(dskip + row * ds_ld + c, sk); else sk[0] = to_f32(dskip[row * ds_ld + c]);
#pragma unroll
            for (int k = 0; k < V; ++k) v[k] += sk[k];
          }
          if (VEC) Vec<T>::store(dx + row * dx_ld + c, v); else dx[row * dx_ld + c] = from_f32<T>(v[0]);
        }
  }
}

// ---- channel-slab copy: dst[r][0:C] = src[r][0:C] with independent pitches (skip-connection concat / split) --------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) slab_copy_kernel(const T* __restrict__ src, int s_ld, T* __restrict__ dst, int d_ld, long long M, int C) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / groups;
    int c = (int)(i - r * groups) * V;
    if (VEC) *reinterpret_cast<uint4*>(dst + r * d_ld + c) = *reinterpret_cast<const uint4*>(src + r * s_ld + c);
    else dst[r * d_ld + c] = src[r * s_ld + c];
  }
}


// ---- space-to-depth (2x; depth factor FD): fine [N][FD*D][2H][2W][C] (pitch ld) -> coarse [N][D][H][W][taps*C] --------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) space_to_depth_kernel(const T* __restrict__ src, int s_ld, T* __restrict__ dst, int N, int D, int H, int W, int C,
                                                             int FD, double* __restrict__ colsum) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V, taps = 4 * FD;
  const long long total = (long long)N * D * H * W * taps * groups;
  // optional per-channel sum of everything copied (= the ConvTranspose bias gradient, reference UNet.py:75-76): the host only asks
  // for it when 256 % groups == 0, so a thread keeps ONE channel group over its whole grid-stride loop
  float cs[V];
#pragma unroll
  for (int k = 0; k < V; ++k) cs[k] = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % groups); long long t = i / groups;
    int tap = (int)(t % taps); t /= taps;            // t = coarse voxel
    int w = (int)(t % W); long long u = t / W;
    int h = (int)(u % H); long long r = u / H;       // r = n*D + d
    int ti = tap >> 2, tj = (tap >> 1) & 1, tl = tap & 1;
    long long fine = ((r * FD + ti) * (2 * H) + (2 * h + tj)) * (2LL * W) + (2 * w + tl);
    const T* sp = src + fine * s_ld + g * V;
    T* dp = dst + (t * taps + tap) * (long long)C + g * V;
    if (VEC) {
      const uint4 raw = *reinterpret_cast<const uint4*>(sp);
      *reinterpret_cast<uint4*>(dp) = raw;
      if (colsum) {
        float v[V];
        Vec<T>::load(reinterpret_cast<const T*>(&raw), v);
#pragma unroll
        for (int k = 0; k < V; ++k) cs[k] += v[k];
      }
    } else {
      *dp = *sp;
      if (colsum) cs[0] += to_f32(*sp);
    }
  }
  if (colsum) {
    // lanes `groups` apart own the same channels: butterfly, per-warp partials, block sum, one fp64 atomic per channel per block
    __shared__ float part[8][256];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, g = threadIdx.x % groups;
#pragma unroll
    for (int k = 0; k < V; ++k)
      for (int o = groups; o < 32; o <<= 1) cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], o);
    for (int i = threadIdx.x; i < 8 * 256; i += 256) part[i >> 8][i & 255] = 0.f;
    __syncthreads();
    if (groups >= 32 || lane < groups) {
#pragma unroll
      for (int k = 0; k < V; ++k) part[wrp][g * V + k] = cs[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][c];
      atomicAdd(&colsum[c], (double)t);
    }
  }
}


// ---- nn.Upsample(scale_factor=2, mode='trilinear'/'bilinear', align_corners=True) of the bilinear=True decoder
//      (reference models/networks/UNet.py:69-72,117).  ATen's index arithmetic in fp32: ratio = (I-1)/(O-1), src = ratio*o,
//      i0 = (int)src, i1 = i0 + (i0 < I-1), w1 = src - i0.  Grid args = the INPUT grid; the output (FD*D, 2H, 2W) may be a channel
//      slab of the concat buffer.  Backward is the gather form (deterministic): every input voxel sums the output voxels whose
//      two taps per dimension include it.
struct UpAxis { int I, O; float ratio; };
__device__ __forceinline__ void up_src(const UpAxis& a, int o, int& i0, int& i1, float& w1) {
  const float src = a.ratio * (float)o;
  i0 = (int)src;
  if (i0 > a.I - 1) i0 = a.I - 1;
  i1 = i0 + (i0 < a.I - 1 ? 1 : 0);
  w1 = src - (float)i0;
}
// weight with which output index o reads input index i along one axis
__device__ __forceinline__ float up_w(const UpAxis& a, int o, int i) {
  int i0, i1; float w1;
  up_src(a, o, i0, i1, w1);
  return (i0 == i ? 1.f - w1 : 0.f) + (i1 == i ? w1 : 0.f);
}
// candidate output range [lo, hi] for input index i: src(o) in (i - 1, i + 1)
__device__ __forceinline__ void up_range(const UpAxis& a, int i, int& lo, int& hi) {
  if (a.ratio <= 0.f) { lo = 0; hi = a.O - 1; return; }
  lo = (int)floorf((float)(i - 1) / a.ratio); hi = (int)ceilf((float)(i + 1) / a.ratio);
  if (lo < 0) lo = 0;
  if (hi > a.O - 1) hi = a.O - 1;
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) upsample2_fwd_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int N, UpAxis ad,
                                                            UpAxis ah, UpAxis aw, int C) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V;
  const long long total = (long long)N * ad.O * ah.O * aw.O * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int c = (int)(t % groups) * V; t /= groups;
    const int ow = (int)(t % aw.O); t /= aw.O;
    const int oh = (int)(t % ah.O); t /= ah.O;
    const int od = (int)(t % ad.O); const long long n = t / ad.O;
    int d0, d1, h0, h1, w0, w1; float fd, fh, fw;
    up_src(ad, od, d0, d1, fd); up_src(ah, oh, h0, h1, fh); up_src(aw, ow, w0, w1, fw);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      const int dd = (corner & 4) ? d1 : d0, hh = (corner & 2) ? h1 : h0, ww = (corner & 1) ? w1 : w0;
      const float wt = ((corner & 4) ? fd : 1.f - fd) * ((corner & 2) ? fh : 1.f - fh) * ((corner & 1) ? fw : 1.f - fw);
      const long long src = (((n * ad.I + dd) * ah.I + hh) * aw.I + ww) * (long long)x_ld + c;
      float v[V];
      if (VEC) Vec<T>::load(x + src, v); else v[0] = to_f32(x[src]);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = fmaf(wt, v[k], acc[k]);
    }
    const long long dst = (((n * ad.O + od) * ah.O + oh) * aw.O + ow) * (long long)y_ld + c;
    if (VEC) Vec<T>::store(y + dst, acc); else y[dst] = from_f32<T>(acc[0]);
  }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const T* __restrict__ dy, int dy_ld, T* __restrict__ dx, int dx_ld, int N, UpAxis ad,
                                                            UpAxis ah, UpAxis aw, int C) {
  constexpr int V = VEC ? Vec<T>::N : 1;
  const int groups = C / V;
  const long long total = (long long)N * ad.I * ah.I * aw.I * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int c = (int)(t % groups) * V; t /= groups;
    const int iw = (int)(t % aw.I); t /= aw.I;
    const int ih = (int)(t % ah.I); t /= ah.I;
    const int id = (int)(t % ad.I); const long long n = t / ad.I;
    int dlo, dhi, hlo, hhi, wlo, whi;
    up_range(ad, id, dlo, dhi); up_range(ah, ih, hlo, hhi); up_range(aw, iw, wlo, whi);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int od = dlo; od <= dhi; ++od) {
      const float wd = up_w(ad, od, id);
      if (wd == 0.f) continue;
      for (int oh = hlo; oh <= hhi; ++oh) {
        const float wh = wd * up_w(ah, oh, ih);
        if (wh == 0.f) continue;
        for (int ow = wlo; ow <= whi; ++ow) {
          const float wt = wh * up_w(aw, ow, iw);
          if (wt == 0.f) continue;
          const long long src = (((n * ad.O + od) * ah.O + oh) * aw.O + ow) * (long long)dy_ld + c;
          float v[V];
          if (VEC) Vec<T>::load(dy + src, v); else v[0] = to_f32(dy[src]);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] = fmaf(wt, v[k], acc[k]);
        }
      }
    }
    const long long dst = (((n * ad.I + id) * ah.I + ih) * aw.I + iw) * (long long)dx_ld + c;
    if (VEC) Vec<T>::store(dx + dst, acc); else dx[dst] = from_f32<T>(acc[0]);
  }
}

inline UpAxis make_axis(int I, int factor) {
  UpAxis a;
  a.I = I; a.O = I * factor;
  a.ratio = a.O > 1 ? (float)(I - 1) / (float)(a.O - 1) : 0.f;
  return a;
}

// ---- weight packing: dst = permute(flip(src)) of a 5-D fp32 tensor, cast to the destination type, one launch ---------------------
struct Perm5 { int dims[5]; int perm[5]; int flip; };
template <typename T>
__global__ void __launch_bounds__(256) permute5_kernel(const float* __restrict__ src, T* __restrict__ dst, Perm5 q, long long total) {
  long long sstride[5];
  sstride[4] = 1;
  for (int i = 3; i >= 0; --i) sstride[i] = sstride[i + 1] * q.dims[i + 1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i, off = 0;
#pragma unroll
    for (int k = 4; k >= 0; --k) {             // destination dim k iterates source dim perm[k]
      const int sd = q.perm[k], n = q.dims[sd];
      int idx = (int)(t % n); t /= n;
      if ((q.flip >> sd) & 1) idx = n - 1 - idx;
      off += idx * sstride[sd];
    }
    dst[i] = from_f32<T>(src[off]);
  }
}

// Batched form: every weight pack of a network that went stale with the optimizer step, in ONE launch (blockIdx.y = job).
constexpr int PERM_BATCH = 48;
struct Perm5Job { const float* src; void* dst; int dims[5]; int perm[5]; int flip; int bf16; int total; };
struct Perm5Batch { Perm5Job job[PERM_BATCH]; };
__global__ void __launch_bounds__(256) permute5_batch_kernel(const __grid_constant__ Perm5Batch b) {
  const Perm5Job& q = b.job[blockIdx.y];
  int sstride[5];
  sstride[4] = 1;
  for (int i = 3; i >= 0; --i) sstride[i] = sstride[i + 1] * q.dims[i + 1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q.total; i += gridDim.x * blockDim.x) {
    int t = i, off = 0;
#pragma unroll
    for (int k = 4; k >= 0; --k) {
      const int sd = q.perm[k], n = q.dims[sd];
      int idx = t % n; t /= n;
      if ((q.flip >> sd) & 1) idx = n - 1 - idx;
      off += idx * sstride[sd];
    }
    const float v = q.src[off];
    if (q.bf16) reinterpret_cast<bf16*>(q.dst)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(q.dst)[i] = v;
  }
}

// ---- global average pool over the voxels of each sample: x [N][S][C] -> out fp32 [N][C]; and its backward -----------
template <typename T>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const T* __restrict__ x, int ld, float* __restrict__ out, long long S, int C) {
  __shared__ float sh[8][33];
  const int n = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);   // gridDim.x == ceil(C / 32)
  float a = 0.f;
  if (c < C)
    for (long long s = threadIdx.x >> 5; s < S; s += 8) a += to_f32(x[((long long)n * S + s) * ld + c]);
  sh[threadIdx.x >> 5][threadIdx.x & 31] = a;
  __syncthreads();
  if (threadIdx.x < 32 && c < C) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    out[(long long)n * C + c] = t / (float)S;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dout, T* __restrict__ dx, int ld, long long S, int C, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    long long n = r / S;
    dx[r * ld + c] = from_f32<T>(dout[n * C + c] / (float)S);
  }
}

// ---- last ConvBlock unit fused with the single-class 1x1 head (reference models/networks/UNet.py:122 after :173-174) -----------------
// z = ReLU(BN(y)) of the last decoder unit is consumed by final_conv only, so it is never materialised: the head reads the conv
// output y (BN + ReLU applied in registers), and the backward pass derives dz = d(logit) * w_head on the fly inside the two
// BatchNorm-backward passes (which read y anyway).  Saves the z write + read, the dz write + two dz reads and one kernel per
// direction.  Row layout as in the *_rows_kernel family: a thread owns one 16-byte channel group, the `groups` = C / V lanes of a
// voxel are adjacent lanes of one warp (groups is a power of two <= 32).
// 16-byte row chunks are kept as raw registers (4 per chunk) while in flight and widened to fp32 only when consumed, so that 8 rows
// per thread can be outstanding (HBM needs ~64-96 KB in flight per SM; measured 3.4 TB/s with 4 widened rows, see profiles/).
template <typename T> struct Raw16;
template <> struct Raw16<float> {
  typedef float4 R;
  __device__ static R zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ static R load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static void unpack(const R& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};
template <> struct Raw16<bf16> {
  typedef uint4 R;
  __device__ static R zero() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static R load(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static void unpack(const R& t, float* v) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(256, 3) bn_head_fwd_rows_kernel(const T* __restrict__ y, int y_ld, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, const float* __restrict__ w,
                                                               const float* __restrict__ b, float* __restrict__ out, long long M, int C, int relu,
                                                               int act) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 8;                       // rows in flight per thread
  const int groups = C / V, rpb = 256 / groups;
  const int gi = threadIdx.x % groups, c = gi * V;
  float sc[V], sf[V], wk[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { sc[k] = scale[c + k]; sf[k] = shift[c + k]; wk[k] = w[c + k]; }
  const float bias = b ? b[0] : 0.f;
  const long long step = (long long)gridDim.x * rpb;
  // ncu (round 2): issue-bound (71 % issue-active, 104 instructions per thread-row), not HBM-bound -> the row index and the row pointer
  // are carried instead of recomputed (integer division + 64-bit multiply per row), and the bf16 engine uses the fast exponential
  const int trow = threadIdx.x / groups;
  const long long ystep = step * y_ld;
  const T* yp = y + ((long long)blockIdx.x * rpb + trow) * y_ld + c;
  // the trip count is uniform over the block (the shuffles below need whole warps); rows past M are masked
  for (long long base = (long long)blockIdx.x * rpb; base < M; base += U * step, yp += U * ystep) {
    typename Raw16<T>::R raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = base + u * step + trow;
      raw[u] = r < M ? Raw16<T>::load(yp + u * ystep) : Raw16<T>::zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = base + u * step + trow;
      float a[V];
      Raw16<T>::unpack(raw[u], a);
      float p = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float z = fmaf(a[k], sc[k], sf[k]);
        if (relu) z = fmaxf(z, 0.f);
        p = fmaf(z, wk[k], p);
      }
      for (int o = 1; o < groups; o <<= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (r < M && gi == 0) {
        p += bias;
        if (act != 1) out[r] = p;
        else if (sizeof(T) == 2) out[r] = __fdividef(1.f, 1.f + __expf(-p));     // bf16 engine: ~2 ulp of fp32, far below its 1e-2 tolerance
        else out[r] = 1.f / (1.f + expf(-p));                                  // fp32 verification mode: as torch.sigmoid
      }
    }
  }
}

// Reduction pass: d(beta) = sum dz, d(gamma) = sum dz * yhat (into the replicated fp64 `sums` of the BN-backward kernels) and the head's
// d(w) = sum z * dl, d(b) = sum dl (into `hsums`, [BN_COPIES][C + 1] fp64), with dl = d(logit) and dz = dl * w_head * [z > 0].
template <typename T, int U>
__global__ void __launch_bounds__(256, U == 8 ? 2 : 3) bn_head_bwd_reduce_rows_kernel(const T* __restrict__ y, int y_ld, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, const float* __restrict__ mean,
                                                                      const float* __restrict__ invstd, const float* __restrict__ w,
                                                                      const float* __restrict__ out, const float* __restrict__ dout, long long M,
                                                                      int C, int relu, int act, double* __restrict__ sums,
                                                                      double* __restrict__ hsums) {
  constexpr int V = Vec<T>::N;
  const int groups = C / V, rpb = 256 / groups;
  const int gi = threadIdx.x % groups, c = gi * V;
  float sc[V], sf[V], mu[V], sb[V], sg[V], gw[V];      // w_head and invstd are applied after the loop: not live across it
#pragma unroll
  for (int k = 0; k < V; ++k) {
    sc[k] = scale[c + k]; sf[k] = shift[c + k]; mu[k] = mean[c + k];
    sb[k] = sg[k] = gw[k] = 0.f;
  }
  float gb = 0.f;
  const long long step = (long long)gridDim.x * rpb;
  for (long long r0 = (long long)blockIdx.x * rpb + threadIdx.x / groups; r0 < M; r0 += U * step) {
    // every load of the iteration is issued before the first use (ncu on the first version: 12 long-scoreboard stalls per issue --
    // d(logit) was computed inside the load loop, so the loads of row u + 1 waited for the scalar loads of row u)
    typename Raw16<T>::R raw[U];
    float dl[U], pl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * step;
      const bool ok = r < M;               // masked row: dout = 0 -> dl = 0 contributes nothing to any sum
      raw[u] = ok ? Raw16<T>::load(y + r * y_ld + c) : Raw16<T>::zero();
      pl[u] = ok ? out[r] : 0.f;
      dl[u] = ok ? dout[r] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) dl[u] = act == 1 ? dl[u] * pl[u] * (1.f - pl[u]) : dl[u];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float a[V];
      Raw16<T>::unpack(raw[u], a);
      if (gi == 0) gb += dl[u];
      // The kernel is issue-bound, not HBM-bound (ncu: 2.0 TB/s with 12 instructions per element): the per-channel constants w_head and
      // invstd are factored out of the sums (applied once after the loop), which leaves 7 instructions per element.
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float z = fmaf(a[k], sc[k], sf[k]);
        const float t = (!relu || z > 0.f) ? dl[u] : 0.f;           // d(logit) where the unit is active
        gw[k] = fmaf(t, z, gw[k]);                                   // head d(w) = sum dl * relu(z)
        sb[k] += t;                                                  // sum dl * [z > 0]             (x w_head      = d(beta))
        sg[k] = fmaf(t, a[k] - mu[k], sg[k]);                        // sum dl * [z > 0] * (y - mean) (x w_head*invstd = d(gamma))
      }
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { const float wk = w[c + k]; sb[k] *= wk; sg[k] *= wk * invstd[c + k]; }
  // lanes that own the same channel group are `groups` apart: butterfly over those, one row of partials per warp, sum over the warps
  extern __shared__ float sh[];  // [8 warps][3][C] + [8]
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < V; ++k)
    for (int o = groups; o < 32; o <<= 1) {
      sb[k] += __shfl_xor_sync(0xffffffffu, sb[k], o);
      sg[k] += __shfl_xor_sync(0xffffffffu, sg[k], o);
      gw[k] += __shfl_xor_sync(0xffffffffu, gw[k], o);
    }
  gb = warp_sum(gb);
  if (lane < groups) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      sh[(wrp * 3 + 0) * C + c + k] = sb[k];
      sh[(wrp * 3 + 1) * C + c + k] = sg[k];
      sh[(wrp * 3 + 2) * C + c + k] = gw[k];
    }
  }
  if (lane == 0) sh[8 * 3 * C + wrp] = gb;
  __syncthreads();
  const int copy = blockIdx.x % BN_COPIES;
  for (int i = threadIdx.x; i < 3 * C; i += 256) {
    const int j = i / C, ch = i - j * C;
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += sh[(wq * 3 + j) * C + ch];
    if (j < 2) atomicAdd(&sums[(size_t)copy * 2 * C + i], (double)t);
    else atomicAdd(&hsums[(size_t)copy * (C + 1) + ch], (double)t);
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += sh[8 * 3 * C + wq];
    atomicAdd(&hsums[(size_t)copy * (C + 1) + C], (double)t);
  }
}

// Apply pass: dy = scale * (dz - d(beta)/M - yhat * d(gamma)/M) with dz rebuilt from d(logit); block 0 also publishes d(gamma),
// d(beta) and the head's d(w), d(b).
template <typename T, int U>
__global__ void __launch_bounds__(256, U == 8 ? 2 : U == 4 ? 3 : 4) bn_head_bwd_apply_rows_kernel(const T* __restrict__ y, int y_ld, const float* __restrict__ scale,
                                                                     const float* __restrict__ shift, const float* __restrict__ mean,
                                                                     const float* __restrict__ invstd, const float* __restrict__ w,
                                                                     const float* __restrict__ out, const float* __restrict__ dout,
                                                                     const double* __restrict__ sums, const double* __restrict__ hsums,
                                                                     T* __restrict__ dy, int dy_ld, long long M, int C, int relu, int training, int act,
                                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw,
                                                                     float* __restrict__ db) {
  constexpr int V = Vec<T>::N;
  const int groups = C / V, rpb = 256 / groups;
  const int c = (threadIdx.x % groups) * V;
  const float invM = training ? (float)(1.0 / (double)M) : 0.f;
  extern __shared__ float tot[];   // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) tot[i] = (float)bn_total(sums, i, C);
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      if (dbeta) dbeta[i] = tot[i];
      if (dgamma) dgamma[i] = tot[C + i];
    }
    for (int i = threadIdx.x; i <= C; i += blockDim.x) {
      double t = 0.0;
      for (int k = 0; k < BN_COPIES; ++k) t += hsums[(size_t)k * (C + 1) + i];
      if (i < C) { if (dw) dw[i] = (float)t; }
      else if (db) db[0] = (float)t;
    }
  }
  float sc[V], sf[V], k0[V], k1[V], wk[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    sc[k] = scale[c + k]; sf[k] = shift[c + k];
    k1[k] = sc[k] * invstd[c + k] * tot[C + c + k] * invM;
    k0[k] = sc[k] * tot[c + k] * invM - mean[c + k] * k1[k];
    wk[k] = sc[k] * w[c + k];                                  // scale * w_head: dy = [z > 0] * dl * (scale * w_head) - (y * k1 + k0)
  }
  const long long step = (long long)gridDim.x * rpb;
  for (long long r0 = (long long)blockIdx.x * rpb + threadIdx.x / groups; r0 < M; r0 += U * step) {
    typename Raw16<T>::R raw[U];       // all loads of the iteration first, d(logit) afterwards (see the reduction kernel)
    float dl[U], pl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * step;
      const bool ok = r < M;
      raw[u] = ok ? Raw16<T>::load(y + r * y_ld + c) : Raw16<T>::zero();
      pl[u] = ok ? out[r] : 0.f;
      dl[u] = ok ? dout[r] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) dl[u] = act == 1 ? dl[u] * pl[u] * (1.f - pl[u]) : dl[u];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * step;
      float a[V];
      Raw16<T>::unpack(raw[u], a);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float t = (!relu || fmaf(a[k], sc[k], sf[k]) > 0.f) ? dl[u] : 0.f;
        a[k] = fmaf(t, wk[k], -fmaf(a[k], k1[k], k0[k]));
      }
      if (r < M) Vec<T>::store(dy + r * dy_ld + c, a);
    }
  }
}

// one wave of resident blocks: grid = #SMs x blocks that fit per SM (the kernels walk the rows with a grid-sized stride)
template <typename K>
inline int one_wave_grid(K kernel, size_t smem, long long M, int rpb) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 2;
  long long blocks = (M + rpb - 1) / rpb, cap = (long long)ich_num_sms() * per_sm;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template <typename T>
inline bool bn_head_ok(int C) {
  const int V = Vec<T>::N;
  if (C % V) return false;
  const int groups = C / V;
  return groups >= 1 && groups <= 32 && (groups & (groups - 1)) == 0;
}

}  // namespace

template <typename T, bool DROPV>
static void affine_act_launch(const void* y, int y_ld, const float* scale, const float* shift, void* z, int z_ld, long long M, int C, int relu,
                              const Drop drop, cudaStream_t s) {
  if (vec_ok<T>(y, y_ld, C) && vec_ok<T>(z, z_ld, C) && rows_fast_ok(C, Vec<T>::N))
    affine_act_rows_kernel<T, DROPV><<<rows_grid(M, C, Vec<T>::N), 256, 0, s>>>((const T*)y, y_ld, scale, shift, (T*)z, z_ld, M, C, relu, drop);
  else if (vec_ok<T>(y, y_ld, C) && vec_ok<T>(z, z_ld, C))
    affine_act_kernel<T, true, DROPV><<<grid_for(M * (C / Vec<T>::N), 256), 256, 0, s>>>((const T*)y, y_ld, scale, shift, (T*)z, z_ld, M, C, relu, drop);
  else
    affine_act_kernel<T, false, DROPV><<<grid_for(M * C, 256), 256, 0, s>>>((const T*)y, y_ld, scale, shift, (T*)z, z_ld, M, C, relu, drop);
}

template <typename T, bool DROPV>
static void bn_act_bwd_launch(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                              const float* invstd, double* sums, void* dy, int dy_ld, float* dgamma, float* dbeta, long long M, int C, int relu,
                              int training, const Drop drop, cudaStream_t s, int phases, long long count) {
  // phases: bit 0 = reduction pass, bit 1 = apply pass (SyncBN all-reduces `sums` between the two)
  const int rows_per_block = 2048;
  unsigned blocks = (unsigned)((M + rows_per_block - 1) / rows_per_block);
  size_t shbytes = sizeof(float) * 2 * C;
  bool vec = vec_ok<T>(dz, dz_ld, C) && vec_ok<T>(y, y_ld, C) && vec_ok<T>(dy, dy_ld, C);
  // the rows kernel keeps 8 per-warp copies of the [2][C] partial sums in shared memory: stay under the 48 KB default limit
  // (C = 1024 -- the bottleneck of the default depth-5 / 64-filter constructor -- takes the generic path below)
  if (vec && rows_fast_ok(C, Vec<T>::N) && 8 * shbytes <= 48 * 1024) {
    const int grid = rows_grid(M, C, Vec<T>::N);
    if (phases & 1) bn_bwd_reduce_rows_kernel<T, DROPV><<<grid, 256, 8 * shbytes, s>>>((const T*)dz, dz_ld, (const T*)y, y_ld, scale, shift, mean, invstd, M, C, relu, sums, drop);
    if (phases & 2) bn_bwd_apply_rows_kernel<T, DROPV><<<grid, 256, shbytes, s>>>((const T*)dz, dz_ld, (const T*)y, y_ld, scale, shift, mean, invstd, sums, (T*)dy, dy_ld, M, C, relu, training, dgamma, dbeta, drop, count);
  } else if (vec) {
    if (phases & 1) bn_bwd_reduce_kernel<T, true, DROPV><<<blocks, 256, shbytes, s>>>((const T*)dz, dz_ld, (const T*)y, y_ld, scale, shift, mean, invstd, M, C, relu, sums, rows_per_block, drop);
    if (phases & 2) bn_bwd_apply_kernel<T, true, DROPV><<<grid_for(M * (C / Vec<T>::N), 256), 256, shbytes, s>>>((const T*)dz, dz_ld, (const T*)y, y_ld, scale, shift, mean, invstd, sums, (T*)dy, dy_ld, M, C, relu, training, dgamma, dbeta, drop, count);
  } else {
    if (phases & 1) bn_bwd_reduce_kernel<T, false, DROPV><<<blocks, 256, shbytes, s>>>((const T*)dz, dz_ld, (const T*)y, y_ld, scale, shift, mean, invstd, M, C, relu, sums, rows_per_block, drop);
    if (phases & 2) bn_bwd_apply_kernel<T, false, DROPV><<<grid_for(M * C, 256), 256, shbytes, s>>>((const T*)dz, dz_ld, (const T*)y, y_ld, scale, shift, mean, invstd, sums, (T*)dy, dy_ld, M, C, relu, training, dgamma, dbeta, drop, count);
  }
}

#define DISPATCH_T(dtype, what, ...)                                   \
  if (dtype == ICH_F32) { typedef float T; __VA_ARGS__ }               \
  else if (dtype == ICH_BF16) { typedef bf16 T; __VA_ARGS__ }          \
  else { ich_set_error("%s: bad dtype %d", what, dtype); return 1; }

extern "C" {

int ich_layout_nc_to_nl(const float* src, void* dst, int dtype, int N, int C, long long S, int dst_ld, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if ((long long)N * C * S == 0) return 0;
  DISPATCH_T(dtype, "ich_layout_nc_to_nl", {
    if (C == 1 && dst_ld == 1) {
      cast_from_f32_kernel<T><<<grid_for((long long)N * S, 256), 256, 0, s>>>(src, (T*)dst, (long long)N * S);
    } else {
      dim3 grid((unsigned)((S + 31) / 32), (C + 31) / 32, N), block(32, 8);
      nc_to_nl_kernel<T><<<grid, block, 0, s>>>(src, (T*)dst, C, S, dst_ld);
    }
  })
  return ich_check_launch("ich_layout_nc_to_nl");
}

int ich_layout_nl_to_nc(const void* src, int dtype, int src_ld, float* dst, int N, int C, long long S, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if ((long long)N * C * S == 0) return 0;
  DISPATCH_T(dtype, "ich_layout_nl_to_nc", {
    if (C == 1 && src_ld == 1) {
      cast_to_f32_kernel<T><<<grid_for((long long)N * S, 256), 256, 0, s>>>((const T*)src, dst, (long long)N * S);
    } else {
      dim3 grid((unsigned)((S + 31) / 32), (C + 31) / 32, N), block(32, 8);
      nl_to_nc_kernel<T><<<grid, block, 0, s>>>((const T*)src, dst, C, S, src_ld);
    }
  })
  return ich_check_launch("ich_layout_nl_to_nc");
}

int ich_colstats(const void* x, int ld, int dtype, long long M, int C, double* sum, double* sumsq, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(sum, 0, sizeof(double) * C, s);
  if (sumsq) cudaMemsetAsync(sumsq, 0, sizeof(double) * C, s);
  DISPATCH_T(dtype, "ich_colstats", { return colstats_launch<T>((const T*)x, ld, M, C, sum, sumsq, s); })
}

int ich_bn_finalize(const double* sum, const double* sumsq, long long count, int C, const float* gamma, const float* beta,
                    const float* conv_bias, float* running_mean, float* running_var, float momentum, float eps, float* scale,
                    float* shift, float* save_mean, float* save_invstd, int training, void* stream) {
  ICH_REQUIRE(training || (running_mean && running_var), "ich_bn_finalize: eval mode needs running statistics");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, (double)count, C, gamma, beta, conv_bias, running_mean,
                                                                        running_var, momentum, eps, scale, shift, save_mean, save_invstd,
                                                                        training);
  return ich_check_launch("ich_bn_finalize");
}

static int affine_act_impl(const void* y, int y_ld, const float* scale, const float* shift, void* z, int z_ld, int dtype, long long M, int C,
                           int relu, const Drop drop, cudaStream_t s, const char* what) {
  if (M * C == 0) return 0;
  // the dropout variants are separate instantiations: the Philox code costs ~50 registers, which halves the occupancy of the
  // (bandwidth-bound) plain kernels if it is merely branched around
  DISPATCH_T(dtype, what, {
    if (drop.thr) affine_act_launch<T, true>(y, y_ld, scale, shift, z, z_ld, M, C, relu, drop, s);
    else affine_act_launch<T, false>(y, y_ld, scale, shift, z, z_ld, M, C, relu, drop, s);
  })
  return ich_check_launch(what);
}

int ich_affine_act(const void* y, int y_ld, const float* scale, const float* shift, void* z, int z_ld, int dtype, long long M, int C,
                   int relu, void* stream) {
  return affine_act_impl(y, y_ld, scale, shift, z, z_ld, dtype, M, C, relu, make_drop(0.f, 0), (cudaStream_t)stream, "ich_affine_act");
}

int ich_affine_act_drop(const void* y, int y_ld, const float* scale, const float* shift, void* z, int z_ld, int dtype, long long M, int C,
                        int relu, float drop_p, long long seed, void* stream) {
  ICH_REQUIRE(drop_p >= 0.f && drop_p <= 1.f, "ich_affine_act_drop: dropout probability %g outside [0, 1]", (double)drop_p);
  return affine_act_impl(y, y_ld, scale, shift, z, z_ld, dtype, M, C, relu, make_drop(drop_p, (unsigned long long)seed), (cudaStream_t)stream,
                         "ich_affine_act_drop");
}

static int bn_act_bwd_impl(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                           const float* invstd, double* sums /*[ICH_BN_SUM_COPIES*2*C] workspace*/, void* dy, int dy_ld, float* dgamma, float* dbeta, int dtype,
                           long long M, int C, int relu, int training, const Drop drop, cudaStream_t s, const char* what, int phases = 3,
                           long long count = 0) {
  if (M * C == 0) return 0;
  if (count <= 0) count = M;
  if (phases & 1) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C * BN_COPIES, s);
  DISPATCH_T(dtype, what, {
    if (drop.thr) bn_act_bwd_launch<T, true>(dz, dz_ld, y, y_ld, scale, shift, mean, invstd, sums, dy, dy_ld, dgamma, dbeta, M, C, relu, training, drop, s, phases, count);
    else bn_act_bwd_launch<T, false>(dz, dz_ld, y, y_ld, scale, shift, mean, invstd, sums, dy, dy_ld, dgamma, dbeta, M, C, relu, training, drop, s, phases, count);
  })
  return ich_check_launch(what);
}

int ich_bn_act_bwd(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                   const float* invstd, double* sums /*[ICH_BN_SUM_COPIES*2*C] workspace*/, void* dy, int dy_ld, float* dgamma, float* dbeta, int dtype,
                   long long M, int C, int relu, int training, void* stream) {
  return bn_act_bwd_impl(dz, dz_ld, y, y_ld, scale, shift, mean, invstd, sums, dy, dy_ld, dgamma, dbeta, dtype, M, C, relu, training,
                         make_drop(0.f, 0), (cudaStream_t)stream, "ich_bn_act_bwd");
}

int ich_bn_act_bwd_drop(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                        const float* invstd, double* sums /*[ICH_BN_SUM_COPIES*2*C] workspace*/, void* dy, int dy_ld, float* dgamma, float* dbeta, int dtype,
                        long long M, int C, int relu, int training, float drop_p, long long seed, void* stream) {
  ICH_REQUIRE(drop_p >= 0.f && drop_p <= 1.f, "ich_bn_act_bwd_drop: dropout probability %g outside [0, 1]", (double)drop_p);
  return bn_act_bwd_impl(dz, dz_ld, y, y_ld, scale, shift, mean, invstd, sums, dy, dy_ld, dgamma, dbeta, dtype, M, C, relu, training,
                         make_drop(drop_p, (unsigned long long)seed), (cudaStream_t)stream, "ich_bn_act_bwd_drop");
}

int ich_upsample2_fwd(const void* x, int x_ld, void* y, int y_ld, int dtype, int N, int D, int H, int W, int C, int FD, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(FD == 1 || FD == 2, "ich_upsample2_fwd: depth factor %d (1 = 2-D, 2 = 3-D)", FD);
  if ((long long)N * D * H * W * C == 0) return 0;
  const UpAxis ad = make_axis(D, FD), ah = make_axis(H, 2), aw = make_axis(W, 2);
  const long long out_vox = (long long)N * ad.O * ah.O * aw.O;
  DISPATCH_T(dtype, "ich_upsample2_fwd", {
    if (vec_ok<T>(x, x_ld, C) && vec_ok<T>(y, y_ld, C))
      upsample2_fwd_kernel<T, true><<<grid_for(out_vox * (C / Vec<T>::N), 256), 256, 0, s>>>((const T*)x, x_ld, (T*)y, y_ld, N, ad, ah, aw, C);
    else
      upsample2_fwd_kernel<T, false><<<grid_for(out_vox * C, 256), 256, 0, s>>>((const T*)x, x_ld, (T*)y, y_ld, N, ad, ah, aw, C);
  })
  return ich_check_launch("ich_upsample2_fwd");
}

int ich_upsample2_bwd(const void* dy, int dy_ld, void* dx, int dx_ld, int dtype, int N, int D, int H, int W, int C, int FD, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(FD == 1 || FD == 2, "ich_upsample2_bwd: depth factor %d (1 = 2-D, 2 = 3-D)", FD);
  if ((long long)N * D * H * W * C == 0) return 0;
  const UpAxis ad = make_axis(D, FD), ah = make_axis(H, 2), aw = make_axis(W, 2);
  const long long in_vox = (long long)N * D * H * W;
  DISPATCH_T(dtype, "ich_upsample2_bwd", {
    if (vec_ok<T>(dy, dy_ld, C) && vec_ok<T>(dx, dx_ld, C))
      upsample2_bwd_kernel<T, true><<<grid_for(in_vox * (C / Vec<T>::N), 256), 256, 0, s>>>((const T*)dy, dy_ld, (T*)dx, dx_ld, N, ad, ah, aw, C);
    else
      upsample2_bwd_kernel<T, false><<<grid_for(in_vox * C, 256), 256, 0, s>>>((const T*)dy, dy_ld, (T*)dx, dx_ld, N, ad, ah, aw, C);
  })
  return ich_check_launch("ich_upsample2_bwd");
}

int ich_bn_act_bwd_sync(const void* dz, int dz_ld, const void* y, int y_ld, const float* scale, const float* shift, const float* mean,
                        const float* invstd, double* sums /*[ICH_BN_SUM_COPIES*2*C] workspace*/, void* dy, int dy_ld, float* dgamma, float* dbeta, int dtype,
                        long long M, int C, int relu, int training, float drop_p, long long seed, int phase, long long global_count, void* stream) {
  ICH_REQUIRE(phase == 1 || phase == 2, "ich_bn_act_bwd_sync: phase must be 1 (reduce) or 2 (apply), got %d", phase);
  ICH_REQUIRE(drop_p >= 0.f && drop_p <= 1.f, "ich_bn_act_bwd_sync: dropout probability %g outside [0, 1]", (double)drop_p);
  return bn_act_bwd_impl(dz, dz_ld, y, y_ld, scale, shift, mean, invstd, sums, dy, dy_ld, dgamma, dbeta, dtype, M, C, relu, training,
                         make_drop(drop_p, (unsigned long long)seed), (cudaStream_t)stream, "ich_bn_act_bwd_sync", phase, global_count);
}

static int maxpool2_fwd_impl(const void* x, int x_ld, void* y, int y_ld, void* skip, int skip_ld, int dtype, int N, int D, int H, int W, int C,
                             int FD, cudaStream_t s, const char* what) {
  // odd sizes: nn.MaxPool floors (models/networks/UNet.py:82), the last plane / row / column is simply not pooled -- except with a skip
  // destination, which must receive EVERY voxel (the kernel walks the pooled windows only)
  ICH_REQUIRE(FD == 1 || FD == 2, "%s: depth factor %d", what, FD);
  ICH_REQUIRE(!skip || (D % FD == 0 && H % 2 == 0 && W % 2 == 0), "%s: grid %dx%dx%d not divisible by the pool (skip copy fused)", what, D, H, W);
  long long outv = (long long)N * (D / FD) * (H / 2) * (W / 2);
  if (outv * C == 0) return 0;
  DISPATCH_T(dtype, what, {
    if (vec_ok<T>(x, x_ld, C) && vec_ok<T>(y, y_ld, C) && (!skip || vec_ok<T>(skip, skip_ld, C)))
      maxpool_fwd_kernel<T, true><<<grid_for(outv * (C / Vec<T>::N), 256), 256, 0, s>>>((const T*)x, x_ld, (T*)y, y_ld, N, D, H, W, C, FD, (T*)skip, skip_ld);
    else
      maxpool_fwd_kernel<T, false><<<grid_for(outv * C, 256), 256, 0, s>>>((const T*)x, x_ld, (T*)y, y_ld, N, D, H, W, C, FD, (T*)skip, skip_ld);
  })
  return ich_check_launch(what);
}

int ich_maxpool2_fwd(const void* x, int x_ld, void* y, int y_ld, int dtype, int N, int D, int H, int W, int C, int FD, void* stream) {
  return maxpool2_fwd_impl(x, x_ld, y, y_ld, nullptr, 0, dtype, N, D, H, W, C, FD, (cudaStream_t)stream, "ich_maxpool2_fwd");
}

int ich_maxpool2_fwd_skip(const void* x, int x_ld, void* y, int y_ld, void* skip, int skip_ld, int dtype, int N, int D, int H, int W, int C,
                          int FD, void* stream) {
  ICH_REQUIRE(skip != nullptr, "ich_maxpool2_fwd_skip: the skip destination is required");
  return maxpool2_fwd_impl(x, x_ld, y, y_ld, skip, skip_ld, dtype, N, D, H, W, C, FD, (cudaStream_t)stream, "ich_maxpool2_fwd_skip");
}

int ich_maxpool2_bwd(const void* x, int x_ld, const void* dy, int dy_ld, void* dx, int dx_ld, int dtype, int N, int D, int H, int W,
                     int C, int FD, const void* dskip, int dskip_ld, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  // odd sizes (floor pooling): voxels outside every window get no pooled gradient -- the CALLER zero-fills dx first; a skip gradient
  // would have to reach them too, so that combination is refused
  ICH_REQUIRE(FD == 1 || FD == 2, "ich_maxpool2_bwd: depth factor %d", FD);
  ICH_REQUIRE(!dskip || (D % FD == 0 && H % 2 == 0 && W % 2 == 0), "ich_maxpool2_bwd: grid %dx%dx%d not divisible by the pool (skip gradient fused)", D, H, W);
  long long outv = (long long)N * (D / FD) * (H / 2) * (W / 2);
  if (outv * C == 0) return 0;
  DISPATCH_T(dtype, "ich_maxpool2_bwd", {
    if (vec_ok<T>(x, x_ld, C) && vec_ok<T>(dy, dy_ld, C) && vec_ok<T>(dx, dx_ld, C) && (!dskip || vec_ok<T>(dskip, dskip_ld, C)))
      maxpool_bwd_kernel<T, true><<<grid_for(outv * (C / Vec<T>::N), 256), 256, 0, s>>>((const T*)x, x_ld, (const T*)dy, dy_ld, (T*)dx, dx_ld, N, D, H, W, C, FD, (const T*)dskip, dskip_ld);
    else
      maxpool_bwd_kernel<T, false><<<grid_for(outv * C, 256), 256, 0, s>>>((const T*)x, x_ld, (const T*)dy, dy_ld, (T*)dx, dx_ld, N, D, H, W, C, FD, (const T*)dskip, dskip_ld);
  })
  return ich_check_launch("ich_maxpool2_bwd");
}

int ich_slab_copy(const void* src, int src_ld, void* dst, int dst_ld, int dtype, long long M, int C, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (M * C == 0) return 0;
  DISPATCH_T(dtype, "ich_slab_copy", {
    if (vec_ok<T>(src, src_ld, C) && vec_ok<T>(dst, dst_ld, C))
      slab_copy_kernel<T, true><<<grid_for(M * (C / Vec<T>::N), 256), 256, 0, s>>>((const T*)src, src_ld, (T*)dst, dst_ld, M, C);
    else
      slab_copy_kernel<T, false><<<grid_for(M * C, 256), 256, 0, s>>>((const T*)src, src_ld, (T*)dst, dst_ld, M, C);
  })
  return ich_check_launch("ich_slab_copy");
}

static int space_to_depth_impl(const void* src, int src_ld, void* dst, int dtype, int N, int D, int H, int W, int C, int FD, double* colsum,
                               cudaStream_t s, const char* what) {
  ICH_REQUIRE(FD == 1 || FD == 2, "%s: FD must be 1 or 2", what);
  long long total = (long long)N * D * H * W * 4 * FD * C;
  if (colsum) cudaMemsetAsync(colsum, 0, sizeof(double) * C, s);
  if (total == 0) return 0;
  DISPATCH_T(dtype, what, {
    const bool vec = vec_ok<T>(src, src_ld, C) && vec_ok<T>(dst, C, C);
    const int groups = vec ? C / Vec<T>::N : C;
    ICH_REQUIRE(!colsum || (C <= 256 && groups <= 256 && 256 % groups == 0), "%s: fused channel sums need C / vector width to divide 256 (C = %d)", what, C);
    if (vec)
      space_to_depth_kernel<T, true><<<grid_for(total / Vec<T>::N, 256), 256, 0, s>>>((const T*)src, src_ld, (T*)dst, N, D, H, W, C, FD, colsum);
    else
      space_to_depth_kernel<T, false><<<grid_for(total, 256), 256, 0, s>>>((const T*)src, src_ld, (T*)dst, N, D, H, W, C, FD, colsum);
  })
  return ich_check_launch(what);
}

int ich_space_to_depth2(const void* src, int src_ld, void* dst, int dtype, int N, int D, int H, int W, int C, int FD, void* stream) {
  return space_to_depth_impl(src, src_ld, dst, dtype, N, D, H, W, C, FD, nullptr, (cudaStream_t)stream, "ich_space_to_depth2");
}

int ich_space_to_depth2_sum(const void* src, int src_ld, void* dst, int dtype, int N, int D, int H, int W, int C, int FD, double* colsum,
                            void* stream) {
  ICH_REQUIRE(colsum != nullptr, "ich_space_to_depth2_sum: colsum is required");
  return space_to_depth_impl(src, src_ld, dst, dtype, N, D, H, W, C, FD, colsum, (cudaStream_t)stream, "ich_space_to_depth2_sum");
}

int ich_permute5(const float* src, void* dst, int dtype, int d0, int d1, int d2, int d3, int d4, int p0, int p1, int p2, int p3, int p4,
                 int flipmask, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  Perm5 q{{d0, d1, d2, d3, d4}, {p0, p1, p2, p3, p4}, flipmask};
  int seen = 0;
  for (int k = 0; k < 5; ++k) { ICH_REQUIRE(q.perm[k] >= 0 && q.perm[k] < 5, "ich_permute5: bad permutation"); seen |= 1 << q.perm[k]; }
  ICH_REQUIRE(seen == 31, "ich_permute5: bad permutation");
  long long total = (long long)d0 * d1 * d2 * d3 * d4;
  if (total == 0) return 0;
  DISPATCH_T(dtype, "ich_permute5", { permute5_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(src, (T*)dst, q, total); })
  return ich_check_launch("ich_permute5");
}

int ich_permute5_batch(int n_jobs, const void* const* src, void* const* dst, const int* dtype, const int* dims /*[n][5]*/,
                       const int* perm /*[n][5]*/, const int* flipmask, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(n_jobs >= 0, "ich_permute5_batch: negative job count");
  for (int j0 = 0; j0 < n_jobs; j0 += PERM_BATCH) {
    Perm5Batch b;
    const int nb = n_jobs - j0 < PERM_BATCH ? n_jobs - j0 : PERM_BATCH;
    int max_total = 0;
    for (int j = 0; j < nb; ++j) {
      Perm5Job& q = b.job[j];
      q.src = (const float*)src[j0 + j]; q.dst = dst[j0 + j]; q.flip = flipmask[j0 + j];
      ICH_REQUIRE(dtype[j0 + j] == ICH_F32 || dtype[j0 + j] == ICH_BF16, "ich_permute5_batch: bad dtype %d", dtype[j0 + j]);
      q.bf16 = dtype[j0 + j] == ICH_BF16;
      long long total = 1;
      int seen = 0;
      for (int k = 0; k < 5; ++k) {
        q.dims[k] = dims[(j0 + j) * 5 + k]; q.perm[k] = perm[(j0 + j) * 5 + k];
        ICH_REQUIRE(q.perm[k] >= 0 && q.perm[k] < 5, "ich_permute5_batch: bad permutation");
        seen |= 1 << q.perm[k];
        total *= q.dims[k];
      }
      ICH_REQUIRE(seen == 31 && total < (1ll << 31), "ich_permute5_batch: bad permutation / tensor too large");
      q.total = (int)total;
      if (q.total > max_total) max_total = q.total;
    }
    if (max_total == 0) continue;
    int gx = (max_total + 256 * 8 - 1) / (256 * 8);      // ~8 elements per thread for the largest job
    if (gx > 64) gx = 64;
    permute5_batch_kernel<<<dim3((unsigned)gx, (unsigned)nb), 256, 0, s>>>(b);
  }
  return ich_check_launch("ich_permute5_batch");
}

int ich_avgpool_fwd(const void* x, int ld, int dtype, float* out, int N, long long S, int C, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if ((long long)N * S * C == 0) return 0;
  dim3 grid((C + 31) / 32, N);
  DISPATCH_T(dtype, "ich_avgpool_fwd", { avgpool_fwd_kernel<T><<<grid, 256, 0, s>>>((const T*)x, ld, out, S, C); })
  return ich_check_launch("ich_avgpool_fwd");
}

int ich_avgpool_bwd(const float* dout, void* dx, int ld, int dtype, int N, long long S, int C, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  long long total = (long long)N * S * C;
  if (total == 0) return 0;
  DISPATCH_T(dtype, "ich_avgpool_bwd", { avgpool_bwd_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(dout, (T*)dx, ld, S, C, total); })
  return ich_check_launch("ich_avgpool_bwd");
}

// ---- last ConvBlock unit + single-class head, fused (see bn_head_*_rows_kernel) ---------------------------------------------------
int ich_bn_head_supported(int dtype, int C) {
  if (dtype == ICH_F32) return bn_head_ok<float>(C) ? 1 : 0;
  if (dtype == ICH_BF16) return bn_head_ok<bf16>(C) ? 1 : 0;
  return 0;
}

int ich_bn_head_fwd(const void* y, int y_ld, int dtype, const float* scale, const float* shift, const float* w, const float* b, float* out,
                    long long M, int C, int relu, int act, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(act == 0 || act == 1, "ich_bn_head_fwd: activation %d (0 none, 1 sigmoid)", act);
  ICH_REQUIRE(ich_bn_head_supported(dtype, C), "ich_bn_head_fwd: unsupported dtype %d / channel count %d", dtype, C);
  if (M == 0) return 0;
  DISPATCH_T(dtype, "ich_bn_head_fwd", {
    ICH_REQUIRE(vec_ok<T>(y, y_ld, C), "ich_bn_head_fwd: rows must be 16-byte aligned (ld %d)", y_ld);
    const int grid = one_wave_grid(bn_head_fwd_rows_kernel<T>, 0, M, 256 / (C / Vec<T>::N));
    bn_head_fwd_rows_kernel<T><<<grid, 256, 0, s>>>((const T*)y, y_ld, scale, shift, w, b, out, M, C, relu, act);
  })
  return ich_check_launch("ich_bn_head_fwd");
}

int ich_bn_head_bwd(const void* y, int y_ld, int dtype, const float* scale, const float* shift, const float* mean, const float* invstd,
                    const float* w, const float* out, const float* dout, double* sums /*[ICH_BN_SUM_COPIES*(3*C+1)] workspace*/, void* dy, int dy_ld,
                    float* dgamma, float* dbeta, float* dw, float* db, long long M, int C, int relu, int training, int act, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(act == 0 || act == 1, "ich_bn_head_bwd: activation %d (0 none, 1 sigmoid)", act);
  ICH_REQUIRE(ich_bn_head_supported(dtype, C), "ich_bn_head_bwd: unsupported dtype %d / channel count %d", dtype, C);
  cudaMemsetAsync(sums, 0, sizeof(double) * BN_COPIES * (3 * C + 1), s);
  double* hsums = sums + (size_t)BN_COPIES * 2 * C;
  if (M == 0) return 0;
  DISPATCH_T(dtype, "ich_bn_head_bwd", {
    ICH_REQUIRE(vec_ok<T>(y, y_ld, C) && vec_ok<T>(dy, dy_ld, C), "ich_bn_head_bwd: rows must be 16-byte aligned (ld %d / %d)", y_ld, dy_ld);
    const int rpb = 256 / (C / Vec<T>::N);
    const size_t sh_reduce = sizeof(float) * (8 * 3 * C + 8), sh_apply = sizeof(float) * 2 * C;
    // rows in flight per thread (A/B switch ICH_HEAD_BWD_U): 8 = one wave of 2 resident blocks per SM with 8 chunks in flight per thread,
    // 4 = up to 8 blocks per SM queued, 4 resident (more warps, fewer bytes in flight per warp -- the shape of the generic BatchNorm kernels)
    static int bwd_u = -1;
    if (bwd_u < 0) { const char* e = getenv("ICH_HEAD_BWD_U"); bwd_u = e ? atoi(e) : 2; if (bwd_u != 8 && bwd_u != 4) bwd_u = 2; }   // measured: head family 0.59 (4) -> 0.50 ms (2) per cfg-3 step
    if (bwd_u == 8) {
      const int grid_r = one_wave_grid(bn_head_bwd_reduce_rows_kernel<T, 8>, sh_reduce, M, rpb);
      const int grid_a = one_wave_grid(bn_head_bwd_apply_rows_kernel<T, 8>, sh_apply, M, rpb);
      bn_head_bwd_reduce_rows_kernel<T, 8><<<grid_r, 256, sh_reduce, s>>>((const T*)y, y_ld, scale, shift, mean, invstd, w, out, dout, M, C, relu, act, sums, hsums);
      bn_head_bwd_apply_rows_kernel<T, 8><<<grid_a, 256, sh_apply, s>>>((const T*)y, y_ld, scale, shift, mean, invstd, w, out, dout, sums, hsums, (T*)dy,
                                                                        dy_ld, M, C, relu, training, act, dgamma, dbeta, dw, db);
    } else {
      const int grid = rows_grid(M, C, Vec<T>::N);
      bn_head_bwd_reduce_rows_kernel<T, 4><<<grid, 256, sh_reduce, s>>>((const T*)y, y_ld, scale, shift, mean, invstd, w, out, dout, M, C, relu, act, sums, hsums);
      if (bwd_u == 2)      // A/B: the shape of the generic BatchNorm apply kernel (2 rows per thread, 4 resident blocks per SM)
        bn_head_bwd_apply_rows_kernel<T, 2><<<grid, 256, sh_apply, s>>>((const T*)y, y_ld, scale, shift, mean, invstd, w, out, dout, sums, hsums, (T*)dy,
                                                                        dy_ld, M, C, relu, training, act, dgamma, dbeta, dw, db);
      else
        bn_head_bwd_apply_rows_kernel<T, 4><<<grid, 256, sh_apply, s>>>((const T*)y, y_ld, scale, shift, mean, invstd, w, out, dout, sums, hsums, (T*)dy,
                                                                        dy_ld, M, C, relu, training, act, dgamma, dbeta, dw, db);
    }
  })
  return ich_check_launch("ich_bn_head_bwd");
}

}  // extern "C"
