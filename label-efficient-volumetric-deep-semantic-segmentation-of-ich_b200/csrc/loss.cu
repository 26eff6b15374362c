// Segmentation head and loss kernels (bandwidth / latency bound):
//   * final 1x1 conv + sigmoid / softmax written straight to fp32 NC(S)          (reference models/networks/UNet.py:84-91,122)
//   * soft-Dice and Dice+BCE ("ComboLoss") one-pass reductions and one-pass backward (reference models/optim/LossFunctions.py:39-63,143-166)
//   * InfoNCE / local InfoNCE: row normalisation, masked log-sum-exp over the cosine-similarity rows, analytic backward,
//     region gather / scatter                                                  (reference models/optim/LossFunctions.py:208-230,308-341)
//   * batch binary confusion matrix                                            (reference utils/tensor_utils.py:12-36)
#include "common.cuh"

namespace {

inline int grid_for(long long work, int threads, int per_sm = 8) {
  long long blocks = (work + threads - 1) / threads;
  long long cap = (long long)ich_num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

constexpr int HEAD_MAX_COUT = 8;
constexpr int HEAD_MAX_CIN = 64;

// ---- final 1x1 conv + activation ----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ x, int ld, const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ out, long long M, long long S, int Cin, int Cout, int act) {
  __shared__ float ws[HEAD_MAX_COUT * HEAD_MAX_CIN + HEAD_MAX_COUT];
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[HEAD_MAX_COUT * HEAD_MAX_CIN + i] = b ? b[i] : 0.f;
  __syncthreads();
  const bool vec = (Cin % Vec<T>::N == 0) && (ld % Vec<T>::N == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    float acc[HEAD_MAX_COUT];
#pragma unroll
    for (int co = 0; co < HEAD_MAX_COUT; ++co) acc[co] = co < Cout ? ws[HEAD_MAX_COUT * HEAD_MAX_CIN + co] : 0.f;
    const T* row = x + m * ld;
    if (vec) {
      for (int c = 0; c < Cin; c += Vec<T>::N) {
        float v[Vec<T>::N];
        Vec<T>::load(row + c, v);
#pragma unroll
        for (int co = 0; co < HEAD_MAX_COUT; ++co)
          if (co < Cout)
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) acc[co] = fmaf(v[k], ws[co * Cin + c + k], acc[co]);
      }
    } else {
      for (int c = 0; c < Cin; ++c) {
        float v = to_f32(row[c]);
#pragma unroll
        for (int co = 0; co < HEAD_MAX_COUT; ++co)
          if (co < Cout) acc[co] = fmaf(v, ws[co * Cin + c], acc[co]);
      }
    }
    if (act == 1) {
#pragma unroll
      for (int co = 0; co < HEAD_MAX_COUT; ++co) acc[co] = 1.f / (1.f + expf(-acc[co]));
    } else if (act == 2) {
      float mx = -INFINITY, sum = 0.f;
#pragma unroll
      for (int co = 0; co < HEAD_MAX_COUT; ++co) if (co < Cout) mx = fmaxf(mx, acc[co]);
#pragma unroll
      for (int co = 0; co < HEAD_MAX_COUT; ++co) if (co < Cout) { acc[co] = expf(acc[co] - mx); sum += acc[co]; }
#pragma unroll
      for (int co = 0; co < HEAD_MAX_COUT; ++co) acc[co] = acc[co] / sum;
    }
    const long long n = m / S, s = m - n * S;
#pragma unroll
    for (int co = 0; co < HEAD_MAX_COUT; ++co)
      if (co < Cout) out[(n * Cout + co) * S + s] = acc[co];
  }
}

// d(logit) from d(out) for the head activation, in place of ATen sigmoid/softmax backward. out/dout are fp32 NC(S);
// dl is written channel-last [M][Cout] in the activation dtype so the generic dgrad / wgrad kernels can consume it.
template <typename T>
__global__ void __launch_bounds__(256) head_dlogit_kernel(const float* __restrict__ out, const float* __restrict__ dout, T* __restrict__ dl,
                                                          long long M, long long S, int Cout, int act) {
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const long long n = m / S, s = m - n * S;
    float dot = 0.f;
    if (act == 2)
      for (int co = 0; co < Cout; ++co) dot += out[(n * Cout + co) * S + s] * dout[(n * Cout + co) * S + s];
    for (int co = 0; co < Cout; ++co) {
      float p = out[(n * Cout + co) * S + s], g = dout[(n * Cout + co) * S + s];
      float v = act == 1 ? g * p * (1.f - p) : act == 2 ? p * (g - dot) : g;
      dl[m * Cout + co] = from_f32<T>(v);
    }
  }
}

// Fused backward for the shipped single-class head (Cout == 1): dx, dw, db in one pass over x.
template <typename T>
__global__ void __launch_bounds__(256) head1_bwd_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ w, const float* __restrict__ out,
                                                        const float* __restrict__ dout, T* __restrict__ dx, int dx_ld, float* __restrict__ dw,
                                                        float* __restrict__ db, long long M, int Cin, int act, int need_dx) {
  __shared__ float ws[HEAD_MAX_CIN];
  __shared__ float red[HEAD_MAX_CIN + 1];
  for (int i = threadIdx.x; i < Cin; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i <= HEAD_MAX_CIN; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float gw[HEAD_MAX_CIN];
#pragma unroll
  for (int c = 0; c < HEAD_MAX_CIN; ++c) gw[c] = 0.f;
  float gb = 0.f;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    float p = out[m], g = dout[m];
    float dl = act == 1 ? g * p * (1.f - p) : g;
    gb += dl;
    const T* row = x + m * x_ld;
#pragma unroll
    for (int c = 0; c < HEAD_MAX_CIN; ++c)
      if (c < Cin) {
        gw[c] = fmaf(dl, to_f32(row[c]), gw[c]);
        if (need_dx) dx[m * dx_ld + c] = from_f32<T>(dl * ws[c]);
      }
  }
#pragma unroll
  for (int c = 0; c < HEAD_MAX_CIN; ++c)
    if (c < Cin) {
      float v = warp_sum(gw[c]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&red[c], v);
    }
  gb = warp_sum(gb);
  if ((threadIdx.x & 31) == 0) atomicAdd(&red[HEAD_MAX_CIN], gb);
  __syncthreads();
  for (int c = threadIdx.x; c < Cin; c += blockDim.x) atomicAdd(&dw[c], red[c]);
  if (threadIdx.x == 0 && db) atomicAdd(db, red[HEAD_MAX_CIN]);
}


// Vectorised variant for the shipped head (Cout == 1, Cin a multiple of the 16-byte vector, aligned rows): one voxel per
// thread per iteration, 16-byte loads of x and stores of dx, Cin register accumulators reduced once per block.
template <typename T, int CIN>
__global__ void __launch_bounds__(256) head1_bwd_vec_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ w, const float* __restrict__ out,
                                                            const float* __restrict__ dout, T* __restrict__ dx, int dx_ld, float* __restrict__ dw,
                                                            float* __restrict__ db, long long M, int act, int need_dx) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[CIN + 1];
  float wreg[CIN], gw[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) { wreg[c] = w[c]; gw[c] = 0.f; }
  for (int i = threadIdx.x; i <= CIN; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float gb = 0.f;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    float p = out[m], g = dout[m];
    float dl = act == 1 ? g * p * (1.f - p) : g;
    gb += dl;
    const T* row = x + m * x_ld;
#pragma unroll
    for (int c = 0; c < CIN; c += V) {
      float v[V], o[V];
      Vec<T>::load(row + c, v);
#pragma unroll
      for (int k = 0; k < V; ++k) { gw[c + k] = fmaf(dl, v[k], gw[c + k]); o[k] = dl * wreg[c + k]; }
      if (need_dx) Vec<T>::store(dx + m * dx_ld + c, o);
    }
  }
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    float v = warp_sum(gw[c]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[c], v);
  }
  gb = warp_sum(gb);
  if ((threadIdx.x & 31) == 0) atomicAdd(&red[CIN], gb);
  __syncthreads();
  for (int c = threadIdx.x; c < CIN; c += blockDim.x) atomicAdd(&dw[c], red[c]);
  if (threadIdx.x == 0 && db) atomicAdd(db, red[CIN]);
}

template <typename T>
bool head1_bwd_vec_launch(const T* x, int x_ld, const float* w, const float* out, const float* dout, T* dx, int dx_ld, float* dw, float* db,
                          long long M, int Cin, int act, cudaStream_t s) {
  const bool aligned = (x_ld % Vec<T>::N == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                       (!dx || ((dx_ld % Vec<T>::N == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0)));
  if (!aligned) return false;
  int grid = grid_for(M, 256, 4);
#define ICH_H1(C) head1_bwd_vec_kernel<T, C><<<grid, 256, 0, s>>>(x, x_ld, w, out, dout, dx, dx_ld, dw, db, M, act, dx != nullptr)
  switch (Cin) {
    case 8: if (Vec<T>::N <= 8) { ICH_H1(8); return true; } return false;
    case 16: ICH_H1(16); return true;
    case 32: ICH_H1(32); return true;
    case 64: ICH_H1(64); return true;
    default: return false;
  }
#undef ICH_H1
}

// ---- soft-Dice / Dice+BCE ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float powp(float v, float P) { return P == 1.f ? v : P == 2.f ? v * v : powf(v, P); }
__device__ __forceinline__ float dpowp(float v, float P) { return P == 1.f ? 1.f : P == 2.f ? 2.f * v : P * powf(v, P - 1.f); }

// acc[b] = { sum p*m, sum p^P, sum m^P, sum m, sum bce_terms }
__global__ void __launch_bounds__(256) seg_loss_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ mask, long long S, float P,
                                                              float beta, int use_bce, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const float* p = pred + (long long)b * S;
  const float* m = mask + (long long)b * S;
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i];
    v[0] = fmaf(pi, mi, v[0]);
    v[1] += powp(pi, P);
    v[2] += powp(mi, P);
    v[3] += mi;
    if (use_bce) v[4] += beta * mi * logf(pi + 1e-14f) + (1.f - beta) * (1.f - mi) * logf(1.f - pi + 1e-14f);
  }
  __shared__ float sh[32 * 5];
  block_sum<5>(v, sh);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 5; ++k) atomicAdd(&acc[b * 5 + k], (double)v[k]);
}

__global__ void seg_loss_finalize_kernel(const double* __restrict__ acc, int B, float eps, float alpha_empty, float w_bce, float w_dice,
                                         int reduction, float* __restrict__ per_sample, float* __restrict__ loss) {
  // single block; B is small
  __shared__ float sh[32];
  float local = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    double inter = acc[b * 5 + 0], uni = acc[b * 5 + 1] + acc[b * 5 + 2];
    float dl = 1.f - (float)((2.0 * inter + eps) / (uni + eps));
    if (!(acc[b * 5 + 3] > 0.0)) dl *= alpha_empty;
    float v = w_dice * dl + w_bce * (float)(-acc[b * 5 + 4]);
    per_sample[b] = v;
    local += v;
  }
  float v1[1] = {local};
  block_sum<1>(v1, sh);
  if (threadIdx.x == 0) loss[0] = reduction == 1 ? v1[0] / (float)B : v1[0];
}

// dpred = gscale[b] * ( w_bce * dBCE/dp + w_dice * dDL/dp )
__global__ void __launch_bounds__(256) seg_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ mask, const double* __restrict__ acc,
                                                           const float* __restrict__ gscale, long long S, float P, float eps, float alpha_empty,
                                                           float w_bce, float w_dice, float beta, int use_bce, float* __restrict__ dpred) {
  const int b = blockIdx.y;
  const float* p = pred + (long long)b * S;
  const float* m = mask + (long long)b * S;
  float* d = dpred + (long long)b * S;
  const float gs = gscale[b];
  const float num = (float)(2.0 * acc[b * 5 + 0] + eps);
  const float den = (float)(acc[b * 5 + 1] + acc[b * 5 + 2] + eps);
  const float a = (acc[b * 5 + 3] > 0.0) ? 1.f : alpha_empty;
  const float c1 = gs * w_dice * a / den, c2 = num / den;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i];
    // DL = 1 - num/den  ->  dDL/dp = -(2 m den - num * P p^(P-1)) / den^2
    float g = -c1 * (2.f * mi - c2 * dpowp(pi, P));
    if (use_bce) g -= gs * w_bce * (beta * mi / (pi + 1e-14f) - (1.f - beta) * (1.f - mi) / (1.f - pi + 1e-14f));
    d[i] = g;
  }
}

// ---- Tversky loss (models/optim/LossFunctions.py:65-114) ------------------------------------------------------------------
// Uses seg_loss_reduce_kernel with P = 1: acc[b] = { TP = sum p*m, sum p, sum m, sum m, - }  ->  FP = sum p - TP, FN = sum m - TP.
// TL = 1 - (TP + eps) / (TP + beta*FN + gamma*FP + eps), times alpha_empty when the mask has no positive voxel.
__global__ void tversky_finalize_kernel(const double* __restrict__ acc, int B, float eps, float alpha_empty, float beta, float gamma, int reduction,
                                        float* __restrict__ per_sample, float* __restrict__ loss) {
  __shared__ float sh[32];
  float local = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double tp = acc[b * 5 + 0], fp = acc[b * 5 + 1] - tp, fn = acc[b * 5 + 2] - tp;
    float tl = 1.f - (float)((tp + eps) / (tp + (double)beta * fn + (double)gamma * fp + eps));
    if (!(acc[b * 5 + 3] > 0.0)) tl *= alpha_empty;
    per_sample[b] = tl;
    local += tl;
  }
  float v1[1] = {local};
  block_sum<1>(v1, sh);
  if (threadIdx.x == 0) loss[0] = reduction == 1 ? v1[0] / (float)B : v1[0];
}

// dTL/dp = -( m*den - num*(m - beta*m + gamma*(1 - m)) ) / den^2   (dTP/dp = m, dFN/dp = -m, dFP/dp = 1 - m)
__global__ void __launch_bounds__(256) tversky_bwd_kernel(const float* __restrict__ mask, const double* __restrict__ acc, const float* __restrict__ gscale,
                                                          long long S, float eps, float alpha_empty, float beta, float gamma,
                                                          float* __restrict__ dpred) {
  const int b = blockIdx.y;
  const float* m = mask + (long long)b * S;
  float* d = dpred + (long long)b * S;
  const double tp = acc[b * 5 + 0], fp = acc[b * 5 + 1] - tp, fn = acc[b * 5 + 2] - tp;
  const double num = tp + eps, den = tp + (double)beta * fn + (double)gamma * fp + eps;
  const float a = (acc[b * 5 + 3] > 0.0) ? 1.f : alpha_empty;
  const float c = gscale[b] * a;
  // the gradient is affine in m: g = k0 + k1 * m
  const float k0 = c * (float)(num * gamma / (den * den));
  const float k1 = c * (float)((num * (1.0 - (double)beta - (double)gamma) - den) / (den * den));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) d[i] = fmaf(k1, m[i], k0);
}

// ---- InfoNCE family ------------------------------------------------------------------------------------------------------
// rows P[rows][E] -> Pn = P / max(||P||, eps), invn = 1 / max(||P||, eps). One warp per row.
__global__ void __launch_bounds__(256) rownorm_kernel(const float* __restrict__ P, float* __restrict__ Pn, float* __restrict__ invn, int rows, int E,
                                                      float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int e = lane; e < E; e += 32) { float v = P[(long long)row * E + e]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  float inv = 1.f / fmaxf(sqrtf(s), eps);
  if (lane == 0) invn[row] = inv;
  for (int e = lane; e < E; e += 32) Pn[(long long)row * E + e] = P[(long long)row * E + e] * inv;
}

// Block per (set b, row i): s_ij = <Pn_i, Pn_j> / tau, lse_i = logsumexp_{j != i} s_ij, rowloss = lse_i - s_{i,(i+A) mod R}.
// The last block to finish sums the row losses in index order (deterministic) into loss[0] = mean.
__global__ void __launch_bounds__(256) infonce_fwd_kernel(const float* __restrict__ Pn, int R, int E, float inv_tau, float* __restrict__ lse,
                                                          float* __restrict__ rowloss, float* __restrict__ loss, unsigned int* __restrict__ counter,
                                                          int total_rows) {
  extern __shared__ float sh[];  // Pn_i [E] + reduction [64]
  float* pi = sh;
  float* red = sh + E;
  const int b = blockIdx.y, i = blockIdx.x, A = R / 2;
  const float* base = Pn + (long long)b * R * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) pi[e] = base[(long long)i * E + e];
  __syncthreads();
  float mx = -INFINITY, spos = 0.f;
  // pass 1: max (recompute in pass 2; R*E is tiny)
  const int jpos = (i + A) % R;
  for (int j = threadIdx.x; j < R; j += blockDim.x) {
    if (j == i) continue;
    const float* pj = base + (long long)j * E;
    float d = 0.f;
    for (int e = 0; e < E; ++e) d = fmaf(pi[e], pj[e], d);
    d *= inv_tau;
    mx = fmaxf(mx, d);
    if (j == jpos) spos = d;
  }
  // block max
  for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); spos += __shfl_xor_sync(0xffffffffu, spos, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mx; red[32 + (threadIdx.x >> 5)] = spos; }
  __syncthreads();
  float bmx = -INFINITY, bpos = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) { bmx = fmaxf(bmx, red[w]); bpos += red[32 + w]; }
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < R; j += blockDim.x) {
    if (j == i) continue;
    const float* pj = base + (long long)j * E;
    float d = 0.f;
    for (int e = 0; e < E; ++e) d = fmaf(pi[e], pj[e], d);
    sum += expf(d * inv_tau - bmx);
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    float l = bmx + logf(t);
    lse[b * R + i] = l;
    rowloss[b * R + i] = l - bpos;
    __threadfence();
    last = atomicAdd(counter, 1u) == (unsigned)(total_rows - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float t = 0.f;
    for (int r = threadIdx.x; r < total_rows; r += blockDim.x) t += ((volatile float*)rowloss)[r];
    float v1[1] = {t};
    block_sum<1>(v1, red);
    if (threadIdx.x == 0) { loss[0] = v1[0] / (float)total_rows; *counter = 0; }
  }
}

// Block per (b, i): dPn_i = sum_j (G_ij + G_ji) Pn_j with G_ij = gs * (softmax_ij [j != i] - [j == pos(i)]) / tau, then through the
// normalisation: dP_i = invn_i * (dPn_i - Pn_i <Pn_i, dPn_i>).
__global__ void __launch_bounds__(256) infonce_bwd_kernel(const float* __restrict__ Pn, const float* __restrict__ invn, const float* __restrict__ lse,
                                                          int R, int E, float inv_tau, const float* __restrict__ gout, float gmul,
                                                          float* __restrict__ dP) {
  extern __shared__ float sh[];  // Pn_i [E] + coef [R] + red [32]
  float* pi = sh;
  float* coef = sh + E;
  float* red = coef + R;
  const int b = blockIdx.y, i = blockIdx.x, A = R / 2;
  const float* base = Pn + (long long)b * R * E;
  const float gs = gout[0] * gmul * inv_tau;
  for (int e = threadIdx.x; e < E; e += blockDim.x) pi[e] = base[(long long)i * E + e];
  __syncthreads();
  const float lse_i = lse[b * R + i];
  const int jpos = (i + A) % R;
  for (int j = threadIdx.x; j < R; j += blockDim.x) {
    float c = 0.f;
    if (j != i) {
      const float* pj = base + (long long)j * E;
      float d = 0.f;
      for (int e = 0; e < E; ++e) d = fmaf(pi[e], pj[e], d);
      d *= inv_tau;
      c = expf(d - lse_i) + expf(d - lse[b * R + j]);
      if (j == jpos) c -= 2.f;   // pos(i) == j  <=>  pos(j) == i
      c *= gs;
    }
    coef[j] = c;
  }
  __syncthreads();
  float dot = 0.f;
  // each thread owns channels e, e + blockDim, ... ; keep dPn in registers for up to 8 strides, else recompute
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < R; ++j) a = fmaf(coef[j], base[(long long)j * E + e], a);
    dP[((long long)b * R + i) * E + e] = a;   // stash dPn
    dot = fmaf(a, pi[e], dot);
  }
  float v1[1] = {dot};
  block_sum<1>(v1, red);
  __shared__ float sdot;
  if (threadIdx.x == 0) sdot = v1[0];
  __syncthreads();
  const float inv = invn[b * R + i];
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    long long o = ((long long)b * R + i) * E + e;
    dP[o] = inv * (dP[o] - pi[e] * sdot);
  }
}

// Gather K x K x C regions (row-major h, w, c) of f [bs][H][W][C] into P[bs][2A][K*K*C] rows view*A .. view*A + A - 1.
__global__ void region_gather_kernel(const float* __restrict__ f, const int* __restrict__ corners, float* __restrict__ P, int bs, int H, int W, int C,
                                     int A, int K, int view) {
  const int E = K * K * C;
  const long long total = (long long)bs * A * E;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int e = (int)(i % E);
    long long t = i / E;
    int a = (int)(t % A), b = (int)(t / A);
    int c = e % C, kw = (e / C) % K, kh = e / (C * K);
    int h0 = corners[(b * A + a) * 2], w0 = corners[(b * A + a) * 2 + 1];
    P[((long long)b * 2 * A + view * A + a) * E + e] = f[(((long long)b * H + h0 + kh) * W + w0 + kw) * C + c];
  }
}
__global__ void region_scatter_kernel(const float* __restrict__ dP, const int* __restrict__ corners, float* __restrict__ df, int bs, int H, int W, int C,
                                      int A, int K, int view) {
  const int E = K * K * C;
  const long long total = (long long)bs * A * E;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int e = (int)(i % E);
    long long t = i / E;
    int a = (int)(t % A), b = (int)(t / A);
    int c = e % C, kw = (e / C) % K, kh = e / (C * K);
    int h0 = corners[(b * A + a) * 2], w0 = corners[(b * A + a) * 2 + 1];
    df[(((long long)b * H + h0 + kh) * W + w0 + kw) * C + c] = dP[((long long)b * 2 * A + view * A + a) * E + e];
  }
}

// ---- confusion matrix: out[b] = {tn, fp, fn, tp} with p = pred >= thr (thr < 0: use pred as given) -------------------------
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long S, float thr,
                                                        double* __restrict__ out) {
  const int b = blockIdx.y;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    float p = pred[(long long)b * S + i], t = target[(long long)b * S + i];
    if (thr >= 0.f) p = p >= thr ? 1.f : 0.f;
    v[0] += (1.f - p) * (1.f - t);
    v[1] += p * (1.f - t);
    v[2] += (1.f - p) * t;
    v[3] += p * t;
  }
  __shared__ float sh[32 * 4];
  block_sum<4>(v, sh);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(&out[b * 4 + k], (double)v[k]);
}

}  // namespace

extern "C" {

int ich_head_fwd(const void* x, int x_ld, int dtype, const float* w, const float* b, float* out, int N, long long S, int Cin, int Cout,
                 int act, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(Cout >= 1 && Cout <= HEAD_MAX_COUT && Cin <= HEAD_MAX_CIN, "ich_head_fwd: supports Cin <= %d, Cout <= %d (got %d, %d)",
              HEAD_MAX_CIN, HEAD_MAX_COUT, Cin, Cout);
  long long M = (long long)N * S;
  if (M == 0) return 0;
  if (dtype == ICH_F32) head_fwd_kernel<float><<<grid_for(M, 256), 256, 0, s>>>((const float*)x, x_ld, w, b, out, M, S, Cin, Cout, act);
  else if (dtype == ICH_BF16) head_fwd_kernel<bf16><<<grid_for(M, 256), 256, 0, s>>>((const bf16*)x, x_ld, w, b, out, M, S, Cin, Cout, act);
  else ICH_REQUIRE(false, "ich_head_fwd: bad dtype %d", dtype);
  return ich_check_launch("ich_head_fwd");
}

int ich_head_dlogit(const float* out, const float* dout, void* dl, int dtype, int N, long long S, int Cout, int act, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  long long M = (long long)N * S;
  if (M == 0) return 0;
  if (dtype == ICH_F32) head_dlogit_kernel<float><<<grid_for(M, 256), 256, 0, s>>>(out, dout, (float*)dl, M, S, Cout, act);
  else if (dtype == ICH_BF16) head_dlogit_kernel<bf16><<<grid_for(M, 256), 256, 0, s>>>(out, dout, (bf16*)dl, M, S, Cout, act);
  else ICH_REQUIRE(false, "ich_head_dlogit: bad dtype %d", dtype);
  return ich_check_launch("ich_head_dlogit");
}

int ich_head1_bwd(const void* x, int x_ld, int dtype, const float* w, const float* out, const float* dout, void* dx, int dx_ld, float* dw,
                  float* db, long long M, int Cin, int act, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(Cin <= HEAD_MAX_CIN, "ich_head1_bwd: Cin <= %d (got %d)", HEAD_MAX_CIN, Cin);
  cudaMemsetAsync(dw, 0, sizeof(float) * Cin, s);
  if (db) cudaMemsetAsync(db, 0, sizeof(float), s);
  if (M == 0) return 0;
  if (dtype == ICH_F32 && head1_bwd_vec_launch<float>((const float*)x, x_ld, w, out, dout, (float*)dx, dx_ld, dw, db, M, Cin, act, s))
    return ich_check_launch("ich_head1_bwd");
  if (dtype == ICH_BF16 && head1_bwd_vec_launch<bf16>((const bf16*)x, x_ld, w, out, dout, (bf16*)dx, dx_ld, dw, db, M, Cin, act, s))
    return ich_check_launch("ich_head1_bwd");
  int grid = grid_for(M, 256, 4);
  if (dtype == ICH_F32) head1_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)x, x_ld, w, out, dout, (float*)dx, dx_ld, dw, db, M, Cin, act, dx != nullptr);
  else if (dtype == ICH_BF16) head1_bwd_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, x_ld, w, out, dout, (bf16*)dx, dx_ld, dw, db, M, Cin, act, dx != nullptr);
  else ICH_REQUIRE(false, "ich_head1_bwd: bad dtype %d", dtype);
  return ich_check_launch("ich_head1_bwd");
}

int ich_seg_loss_fwd(const float* pred, const float* mask, int B, long long S, float P, float eps, float alpha_empty, float w_bce, float w_dice,
                     float beta, int reduction, double* acc /*[B*5]*/, float* per_sample /*[B]*/, float* loss /*[1]*/, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(B > 0 && B <= 65535, "ich_seg_loss_fwd: batch %d out of range", B);
  cudaMemsetAsync(acc, 0, sizeof(double) * 5 * B, s);
  int bx = (int)((S + 256 * 8 - 1) / (256 * 8));
  int cap = (ich_num_sms() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  seg_loss_reduce_kernel<<<dim3(bx, B), 256, 0, s>>>(pred, mask, S, P, beta, w_bce != 0.f, acc);
  seg_loss_finalize_kernel<<<1, 256, 0, s>>>(acc, B, eps, alpha_empty, w_bce, w_dice, reduction, per_sample, loss);
  return ich_check_launch("ich_seg_loss_fwd");
}

int ich_seg_loss_bwd(const float* pred, const float* mask, const double* acc, const float* gscale /*[B]*/, int B, long long S, float P, float eps,
                     float alpha_empty, float w_bce, float w_dice, float beta, float* dpred, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int bx = (int)((S + 256 * 4 - 1) / (256 * 4));
  int cap = (ich_num_sms() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  seg_loss_bwd_kernel<<<dim3(bx, B), 256, 0, s>>>(pred, mask, acc, gscale, S, P, eps, alpha_empty, w_bce, w_dice, beta, w_bce != 0.f, dpred);
  return ich_check_launch("ich_seg_loss_bwd");
}

int ich_tversky_loss_fwd(const float* pred, const float* mask, int B, long long S, float eps, float alpha_empty, float beta, float gamma,
                         int reduction, double* acc /*[B*5]*/, float* per_sample /*[B]*/, float* loss /*[1]*/, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(B > 0 && B <= 65535, "ich_tversky_loss_fwd: batch %d out of range", B);
  cudaMemsetAsync(acc, 0, sizeof(double) * 5 * B, s);
  int bx = (int)((S + 256 * 8 - 1) / (256 * 8));
  int cap = (ich_num_sms() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  seg_loss_reduce_kernel<<<dim3(bx, B), 256, 0, s>>>(pred, mask, S, 1.f, 0.f, 0, acc);
  tversky_finalize_kernel<<<1, 256, 0, s>>>(acc, B, eps, alpha_empty, beta, gamma, reduction, per_sample, loss);
  return ich_check_launch("ich_tversky_loss_fwd");
}

int ich_tversky_loss_bwd(const float* mask, const double* acc, const float* gscale /*[B]*/, int B, long long S, float eps, float alpha_empty,
                         float beta, float gamma, float* dpred, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(B > 0 && B <= 65535, "ich_tversky_loss_bwd: batch %d out of range", B);
  int bx = (int)((S + 256 * 4 - 1) / (256 * 4));
  int cap = (ich_num_sms() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  tversky_bwd_kernel<<<dim3(bx, B), 256, 0, s>>>(mask, acc, gscale, S, eps, alpha_empty, beta, gamma, dpred);
  return ich_check_launch("ich_tversky_loss_bwd");
}

// P: [B][R][E] fp32 (R = 2 * set size). Outputs: Pn [B][R][E], invn [B*R], lse [B*R], rowloss [B*R], loss [1]; counter: zeroed u32.
int ich_infonce_fwd(const float* P, int B, int R, int E, float tau, float* Pn, float* invn, float* lse, float* rowloss, float* loss,
                    unsigned int* counter, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ICH_REQUIRE(R >= 2 && (R % 2) == 0, "ich_infonce_fwd: R must be even (2 * set size), got %d", R);
  ICH_REQUIRE((size_t)(E + 64) * 4 <= 200 * 1024, "ich_infonce_fwd: embedding dim %d too large", E);
  int rows = B * R;
  rownorm_kernel<<<(rows + 7) / 8, 256, 0, s>>>(P, Pn, invn, rows, E, 1e-8f);
  size_t sh = sizeof(float) * (E + 64);
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is a per-DEVICE setting
  int attr_dev = 0; cudaGetDevice(&attr_dev);
  bool& attr_set = attr_done[attr_dev & 63];
  if (!attr_set) { cudaFuncSetAttribute(infonce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  infonce_fwd_kernel<<<dim3(R, B), 256, sh, s>>>(Pn, R, E, 1.f / tau, lse, rowloss, loss, counter, rows);
  return ich_check_launch("ich_infonce_fwd");
}

int ich_infonce_bwd(const float* Pn, const float* invn, const float* lse, int B, int R, int E, float tau, const float* gout, float* dP,
                    void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  size_t sh = sizeof(float) * (E + R + 32);
  ICH_REQUIRE(sh <= 200 * 1024, "ich_infonce_bwd: E + R too large (%d, %d)", E, R);
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is a per-DEVICE setting
  int attr_dev = 0; cudaGetDevice(&attr_dev);
  bool& attr_set = attr_done[attr_dev & 63];
  if (!attr_set) { cudaFuncSetAttribute(infonce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  infonce_bwd_kernel<<<dim3(R, B), 256, sh, s>>>(Pn, invn, lse, R, E, 1.f / tau, gout, 1.f / (float)(B * R), dP);
  return ich_check_launch("ich_infonce_bwd");
}

int ich_region_gather(const float* f, const int* corners, float* P, int bs, int H, int W, int C, int A, int K, int view, void* stream) {
  long long total = (long long)bs * A * K * K * C;
  if (total == 0) return 0;
  region_gather_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(f, corners, P, bs, H, W, C, A, K, view);
  return ich_check_launch("ich_region_gather");
}

int ich_region_scatter(const float* dP, const int* corners, float* df, int bs, int H, int W, int C, int A, int K, int view, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(df, 0, sizeof(float) * (size_t)bs * H * W * C, s);
  long long total = (long long)bs * A * K * K * C;
  if (total == 0) return 0;
  region_scatter_kernel<<<grid_for(total, 256), 256, 0, s>>>(dP, corners, df, bs, H, W, C, A, K, view);
  return ich_check_launch("ich_region_scatter");
}

int ich_confusion(const float* pred, const float* target, int B, long long S, float thr, double* out /*[B*4]*/, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double) * 4 * B, s);
  int bx = (int)((S + 256 * 8 - 1) / (256 * 8));
  int cap = (ich_num_sms() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  confusion_kernel<<<dim3(bx, B), 256, 0, s>>>(pred, target, S, thr, out);
  return ich_check_launch("ich_confusion");
}

}  // extern "C"
