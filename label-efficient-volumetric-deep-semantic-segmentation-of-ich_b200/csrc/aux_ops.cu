// Kernels around the U-Net block, all HBM- or latency-bound:
//   * ich_stage_ct        -- input staging: raw CT (int16 / uint16 / uint8 / fp32 Hounsfield units) -> window -> clip -> engine dtype, one pass
//                            (replaces the host-side window_ct of utils/ct_utils.py:13-36 + `input.to(device).float()` of
//                            models/optim/UNet2D.py:137-138 + the layout hop; a 1-channel volume is already channel-last)
//   * ich_window_gather / ich_window_scatter / ich_blend_threshold -- sliding-window inference (models/optim/UNet2D.py:272-314 turned
//                            into a 3-D window driver): window extraction, stitching (+ mean blending of
//                            overlaps) and the `pred >= 0.5` mask of UNet2D.py:220,303 on the device
//   * ich_linear_fwd / _bwd -- the MLPHead's Linear(+ReLU) layers on a [B, K] matrix (models/networks/UNet.py:179-209)
//   * ich_gate_mul_fwd / _bwd -- out = feat * sigmoid(gate) of GatedConv (models/networks/GatedUNet.py:303-322)
#include "common.cuh"

namespace {

inline int grid_1d(long long work, int threads, int per_sm = 8) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)ich_num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------------------------- staging
template <typename S> __device__ __forceinline__ float hu_to_f32(S v) { return (float)v; }

// 8 elements per thread per iteration: one 16-byte load of int16 (two for fp32, half of one for uint8), one 16-byte bf16 store
template <typename S, typename T>
__global__ void __launch_bounds__(256) stage_ct_kernel(const S* __restrict__ src, T* __restrict__ dst, long long M, float a, float b, float lo,
                                                       float hi) {
  const long long n8 = M / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    if (sizeof(S) == 2) {
      const uint4 t = reinterpret_cast<const uint4*>(src)[i];
      const S* h = reinterpret_cast<const S*>(&t);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = hu_to_f32(h[k]);
    } else if (sizeof(S) == 4) {
      const float4 t0 = reinterpret_cast<const float4*>(src)[2 * i], t1 = reinterpret_cast<const float4*>(src)[2 * i + 1];
      v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    } else {
      const uint2 t = reinterpret_cast<const uint2*>(src)[i];
      const S* h = reinterpret_cast<const S*>(&t);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = hu_to_f32(h[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fminf(fmaxf(fmaf(v[k], a, b), lo), hi);
    if (sizeof(T) == 2) {
      Vec<bf16>::store(reinterpret_cast<bf16*>(dst) + 8 * i, v);
    } else {
      reinterpret_cast<float4*>(dst)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  // tail (M % 8 elements)
  if (blockIdx.x == 0 && threadIdx.x < (int)(M - n8 * 8)) {
    const long long i = n8 * 8 + threadIdx.x;
    dst[i] = from_f32<T>(fminf(fmaxf(fmaf(hu_to_f32(src[i]), a, b), lo), hi));
  }
}

template <typename S>
int stage_launch(const void* src, void* dst, int dtype, long long M, float a, float b, float lo, float hi, cudaStream_t s) {
  const int grid = grid_1d(M / 8 + 1, 256);
  if (dtype == ICH_BF16) stage_ct_kernel<S, bf16><<<grid, 256, 0, s>>>((const S*)src, (bf16*)dst, M, a, b, lo, hi);
  else stage_ct_kernel<S, float><<<grid, 256, 0, s>>>((const S*)src, (float*)dst, M, a, b, lo, hi);
  return ich_check_launch("ich_stage_ct");
}

// ---------------------------------------------------------------------------------------------------------- window gather / scatter
// Volume [D][H][W] (one channel), windows [nw][wd][wh][ww]; starts[nw][3] = (d0, h0, w0).  Voxels of a window that fall outside the
// volume (window larger than the volume along an axis, or a corner outside it) read as 0 / are not written.
template <typename T>
__global__ void __launch_bounds__(256) window_gather_kernel(const T* __restrict__ vol, int D, int H, int W, const int* __restrict__ starts, int wd,
                                                            int wh, int ww, T* __restrict__ out, long long total) {
  const long long per = (long long)wd * wh * ww;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    long long r = i - (long long)n * per;
    const int x = (int)(r % ww); r /= ww;
    const int y = (int)(r % wh); const int z = (int)(r / wh);
    const int d = starts[3 * n] + z, h = starts[3 * n + 1] + y, w = starts[3 * n + 2] + x;
    out[i] = ((unsigned)d < (unsigned)D && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W) ? vol[((long long)d * H + h) * W + w] : from_f32<T>(0.f);
  }
}

// mode 0 (windows do not overlap): pred_out[v] = p, mask_out[v] = p >= thr written directly.
// mode 1 (overlap): acc[v] += p, cnt[v] += 1 (fp32 atomics; ich_blend_threshold finishes).
__global__ void __launch_bounds__(256) window_scatter_kernel(const float* __restrict__ p, int D, int H, int W, const int* __restrict__ starts, int wd,
                                                             int wh, int ww, long long total, int mode, float thr, float* __restrict__ acc,
                                                             float* __restrict__ cnt, unsigned char* __restrict__ mask) {
  const long long per = (long long)wd * wh * ww;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    long long r = i - (long long)n * per;
    const int x = (int)(r % ww); r /= ww;
    const int y = (int)(r % wh); const int z = (int)(r / wh);
    const int d = starts[3 * n] + z, h = starts[3 * n + 1] + y, w = starts[3 * n + 2] + x;
    if ((unsigned)d >= (unsigned)D || (unsigned)h >= (unsigned)H || (unsigned)w >= (unsigned)W) continue;      // also rejects negative corners
    const long long v = ((long long)d * H + h) * W + w;
    const float val = p[i];
    if (mode == 0) {
      if (acc) acc[v] = val;
      if (mask) mask[v] = val >= thr ? 1 : 0;
    } else {
      atomicAdd(&acc[v], val);
      atomicAdd(&cnt[v], 1.f);
    }
  }
}

__global__ void __launch_bounds__(256) blend_threshold_kernel(float* __restrict__ acc, const float* __restrict__ cnt, long long M, float thr,
                                                              unsigned char* __restrict__ mask) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float c = cnt[i];
    const float v = c > 0.f ? acc[i] / c : 0.f;
    acc[i] = v;
    if (mask) mask[i] = v >= thr ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------ linear
// out[b][n] = act(bias[n] + sum_k x[b][k] * w[n][k]); one warp per output element (K-strided lanes + shuffle tree)
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, int B, int K, int N, int relu) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long o = warp; o < (long long)B * N; o += nwarps) {
    const int b = (int)(o / N), n = (int)(o - (long long)b * N);
    const float* xr = x + (long long)b * K;
    const float* wr = w + (long long)n * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(xr[k], wr[k], s);
    s = warp_sum(s);
    if (lane == 0) {
      s += bias ? bias[n] : 0.f;
      out[o] = relu ? fmaxf(s, 0.f) : s;
    }
  }
}

// g[b][n] = dout[b][n] * (relu ? out[b][n] > 0 : 1);  dx[b][k] = sum_n g[b][n] w[n][k]   (thread per (b, k): coalesced over k)
__global__ void __launch_bounds__(256) linear_dx_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ w,
                                                        float* __restrict__ dx, int B, int K, int N, int relu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * K; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / K), k = (int)(i - (long long)b * K);
    float s = 0.f;
    for (int n = 0; n < N; ++n) {
      const float g = (relu && !(out[(long long)b * N + n] > 0.f)) ? 0.f : dout[(long long)b * N + n];
      s = fmaf(g, w[(long long)n * K + k], s);
    }
    dx[i] = s;
  }
}

// dw[n][k] = sum_b g[b][n] x[b][k] (thread per (n, k)); db[n] = sum_b g[b][n] (threads with k == 0)
__global__ void __launch_bounds__(256) linear_dw_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ x,
                                                        float* __restrict__ dw, float* __restrict__ db, int B, int K, int N, int relu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * K; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / K), k = (int)(i - (long long)n * K);
    float s = 0.f, sb = 0.f;
    for (int b = 0; b < B; ++b) {
      const float g = (relu && !(out[(long long)b * N + n] > 0.f)) ? 0.f : dout[(long long)b * N + n];
      s = fmaf(g, x[(long long)b * K + k], s);
      sb += g;
    }
    dw[i] = s;
    if (k == 0 && db) db[n] = sb;
  }
}

// --------------------------------------------------------------------------------------------------------------- gated conv
template <typename T>
__global__ void __launch_bounds__(256) gate_mul_fwd_kernel(const T* __restrict__ feat, const T* __restrict__ gate, T* __restrict__ out, long long M) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float s = 1.f / (1.f + __expf(-to_f32(gate[i])));
    out[i] = from_f32<T>(to_f32(feat[i]) * s);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) gate_mul_bwd_kernel(const T* __restrict__ feat, const T* __restrict__ gate, const T* __restrict__ dout,
                                                           T* __restrict__ dfeat, T* __restrict__ dgate, long long M) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float s = 1.f / (1.f + __expf(-to_f32(gate[i])));
    const float g = to_f32(dout[i]);
    dfeat[i] = from_f32<T>(g * s);
    dgate[i] = from_f32<T>(g * to_f32(feat[i]) * s * (1.f - s));
  }
}

}  // namespace

extern "C" {

// src_dtype: 0 fp32, 1 int16, 2 uint16, 3 uint8.  dst[i] = clip((src[i] - win_min) * (out_hi - out_lo) / (win_max - win_min) + out_lo, out_lo, out_hi)
int ich_stage_ct(const void* src, int src_dtype, void* dst, int dtype, long long M, float win_min, float win_max, float out_lo, float out_hi,
                 void* stream) {
  ICH_REQUIRE(M >= 0 && win_max > win_min && out_hi >= out_lo, "ich_stage_ct: bad window [%g, %g] -> [%g, %g]", win_min, win_max, out_lo, out_hi);
  ICH_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "ich_stage_ct: buffers must be 16-byte aligned");
  ICH_REQUIRE(dtype == ICH_F32 || dtype == ICH_BF16, "ich_stage_ct: bad dtype %d", dtype);
  if (M == 0) return 0;
  const float a = (out_hi - out_lo) / (win_max - win_min), b = out_lo - win_min * a;
  cudaStream_t s = (cudaStream_t)stream;
  switch (src_dtype) {
    case 0: return stage_launch<float>(src, dst, dtype, M, a, b, out_lo, out_hi, s);
    case 1: return stage_launch<short>(src, dst, dtype, M, a, b, out_lo, out_hi, s);
    case 2: return stage_launch<unsigned short>(src, dst, dtype, M, a, b, out_lo, out_hi, s);
    case 3: return stage_launch<unsigned char>(src, dst, dtype, M, a, b, out_lo, out_hi, s);
  }
  ich_set_error("ich_stage_ct: bad src_dtype %d", src_dtype);
  return 1;
}

int ich_window_gather(const void* vol, int dtype, int D, int H, int W, const int* starts, int n_win, int wd, int wh, int ww, void* out, void* stream) {
  ICH_REQUIRE(n_win >= 0 && wd > 0 && wh > 0 && ww > 0, "ich_window_gather: bad window %dx%dx%d x %d", wd, wh, ww, n_win);
  const long long total = (long long)n_win * wd * wh * ww;
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ICH_BF16) window_gather_kernel<bf16><<<grid_1d(total, 256), 256, 0, s>>>((const bf16*)vol, D, H, W, starts, wd, wh, ww, (bf16*)out, total);
  else if (dtype == ICH_F32) window_gather_kernel<float><<<grid_1d(total, 256), 256, 0, s>>>((const float*)vol, D, H, W, starts, wd, wh, ww, (float*)out, total);
  else { ich_set_error("ich_window_gather: bad dtype %d", dtype); return 1; }
  return ich_check_launch("ich_window_gather");
}

// mode 0: non-overlapping windows, acc (optional) receives the prediction and mask (optional, uint8) the thresholded prediction;
// mode 1: overlapping windows, acc / cnt (both required, zeroed by the caller) accumulate; finish with ich_blend_threshold.
int ich_window_scatter(const float* pred, int D, int H, int W, const int* starts, int n_win, int wd, int wh, int ww, int mode, float thr,
                       float* acc, float* cnt, unsigned char* mask, void* stream) {
  ICH_REQUIRE(mode == 0 || (acc && cnt), "ich_window_scatter: overlap mode needs acc and cnt");
  const long long total = (long long)n_win * wd * wh * ww;
  if (total == 0) return 0;
  window_scatter_kernel<<<grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(pred, D, H, W, starts, wd, wh, ww, total, mode, thr, acc, cnt, mask);
  return ich_check_launch("ich_window_scatter");
}

// acc[i] /= cnt[i] (0 where no window covered the voxel); mask[i] = acc[i] >= thr (optional)
int ich_blend_threshold(float* acc, const float* cnt, long long M, float thr, unsigned char* mask, void* stream) {
  if (M == 0) return 0;
  blend_threshold_kernel<<<grid_1d(M, 256), 256, 0, (cudaStream_t)stream>>>(acc, cnt, M, thr, mask);
  return ich_check_launch("ich_blend_threshold");
}

int ich_linear_fwd(const float* x, const float* w, const float* bias, float* out, int B, int K, int N, int relu, void* stream) {
  ICH_REQUIRE(B > 0 && K > 0 && N > 0, "ich_linear_fwd: bad shape %d x %d -> %d", B, K, N);
  linear_fwd_kernel<<<grid_1d((long long)B * N * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, w, bias, out, B, K, N, relu);
  return ich_check_launch("ich_linear_fwd");
}

// out = the forward output (needed only when relu != 0); any of dx / dw / db may be NULL (db needs dw)
int ich_linear_bwd(const float* dout, const float* out, const float* x, const float* w, float* dx, float* dw, float* db, int B, int K, int N,
                   int relu, void* stream) {
  ICH_REQUIRE(B > 0 && K > 0 && N > 0 && (!relu || out), "ich_linear_bwd: bad arguments");
  ICH_REQUIRE(!db || dw, "ich_linear_bwd: the bias gradient is produced together with the weight gradient");
  cudaStream_t s = (cudaStream_t)stream;
  if (dx) linear_dx_kernel<<<grid_1d((long long)B * K, 256), 256, 0, s>>>(dout, out, w, dx, B, K, N, relu);
  if (dw) linear_dw_kernel<<<grid_1d((long long)N * K, 256), 256, 0, s>>>(dout, out, x, dw, db, B, K, N, relu);
  return ich_check_launch("ich_linear_bwd");
}

int ich_gate_mul_fwd(const void* feat, const void* gate, void* out, int dtype, long long M, void* stream) {
  if (M == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ICH_BF16) gate_mul_fwd_kernel<bf16><<<grid_1d(M, 256), 256, 0, s>>>((const bf16*)feat, (const bf16*)gate, (bf16*)out, M);
  else if (dtype == ICH_F32) gate_mul_fwd_kernel<float><<<grid_1d(M, 256), 256, 0, s>>>((const float*)feat, (const float*)gate, (float*)out, M);
  else { ich_set_error("ich_gate_mul_fwd: bad dtype %d", dtype); return 1; }
  return ich_check_launch("ich_gate_mul_fwd");
}

int ich_gate_mul_bwd(const void* feat, const void* gate, const void* dout, void* dfeat, void* dgate, int dtype, long long M, void* stream) {
  if (M == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ICH_BF16)
    gate_mul_bwd_kernel<bf16><<<grid_1d(M, 256), 256, 0, s>>>((const bf16*)feat, (const bf16*)gate, (const bf16*)dout, (bf16*)dfeat, (bf16*)dgate, M);
  else if (dtype == ICH_F32)
    gate_mul_bwd_kernel<float><<<grid_1d(M, 256), 256, 0, s>>>((const float*)feat, (const float*)gate, (const float*)dout, (float*)dfeat, (float*)dgate, M);
  else { ich_set_error("ich_gate_mul_bwd: bad dtype %d", dtype); return 1; }
  return ich_check_launch("ich_gate_mul_bwd");
}

}  // extern "C"
