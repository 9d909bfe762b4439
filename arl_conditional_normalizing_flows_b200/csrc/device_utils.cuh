// Small device helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>

#include "cnf_internal.h"

namespace cnf {

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : CNF_LRELU_SLOPE * x; }

// Address of element (i, j, k) of mask(view, m, compress=True) inside the strided active view.
//   m = 0/1 checkerboard (M:726-748): channels [0,D) <- pixel (2i, 2j+a), [D,2D) <- (2i+1, 2j+1-a), a = m
//   m = 2/3 channel parity (M:753-759): channel k <- 2k + (m-2)
//   m = 4   dense (input already compressed, A_wrapper/b_wrapper M:452-472)
__device__ __forceinline__ long long comp_off(const FlowView& v, int m, int b, int i, int j, int k) {
  int y, x, c;
  if (m < 2) {
    const int half = k >= v.D ? 1 : 0;
    c = k - half * v.D;
    y = 2 * i + half;
    x = 2 * j + (half ? 1 - m : m);
  } else if (m < 4) {
    y = i;
    x = j;
    c = 2 * k + (m - 2);
  } else {
    y = i;
    x = j;
    c = k;
  }
  return (long long)b * v.sb + (long long)y * v.sy + (long long)x * v.sx + c;
}

// LayerNorm coefficients of one sample from the (sum, sum of squares) pair accumulated by the
// producing kernel.  Biased variance, eps 1e-3 (F:350-360 -> keras LayerNormalization defaults).
__device__ __forceinline__ void ln_coeffs(const double* __restrict__ stats, long long idx, double n,
                                          float& mean, float& rstd) {
  // The sums are fp64; mean and the CENTRED variance are formed in fp64 too (E[x^2] - mean^2 cancels catastrophically in
  // fp32 once |mean| is tens of standard deviations: large biases / a growing residual stream).  Three DP multiplies and one
  // DP subtraction per call; only the reciprocal square root runs in fp32.
  const double inv_n = 1.0 / n;
  const double m = stats[2 * idx] * inv_n;
  double var = stats[2 * idx + 1] * inv_n - m * m;
  var = var > 0.0 ? var : 0.0;
  mean = (float)m;
  rstd = 1.0f / sqrtf((float)var + (float)CNF_LN_EPS);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of two floats; result valid on thread 0 (as doubles).  `red` is >= 2*32 floats.
__device__ __forceinline__ void block_sum2(float a, float b, float* red, double& ra, double& rb) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) {
    red[wid] = a;
    red[32 + wid] = b;
  }
  __syncthreads();
  ra = 0.0;
  rb = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nw; ++i) {
      ra += (double)red[i];
      rb += (double)red[32 + i];
    }
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace cnf
