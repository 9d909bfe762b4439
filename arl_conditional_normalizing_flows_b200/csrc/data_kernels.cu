// HBM-bound element-wise kernels on the INPUT and OUTPUT side of the flow: the data
// helpers of conv_cINN_base_functions.py that the reference runs as tf.data maps on the host CPU.
//   down / up                     F:74-164   2x2 average pool / 2x2 pixel repeat (NHWC)
//   preprocess_dataset_SR         F:233-279  x = down^lx(hires), y = up^(ly-lx)(down^ly(hires)), x -= y, concat  -> ONE kernel
//   preprocess_dataset_class      F:174-231  x -> (logit(a + (1-a) b x) - logit(a)) / (logit(1-a) - logit(a))
//   de_logitify                   F:287-318  the inverse map, applied to samples
//   instance_noise / renew_noise  F:635-676  alpha x + (1 - alpha) N(0,1)  (Philox4x32-10 + Box-Muller, counter based:
//                                            the result depends on (seed, offset, element index) only, not on the grid)
#include <cuda_runtime.h>

#include <cstdint>

#include "cnf_internal.h"

namespace cnf {

// mean over the 2^L x 2^L block whose top-left corner is (y0, x0), evaluated as the reference nests it:
// down(down(..)) = mean of means, each level ((a + b) + (c + d)) / 4 in fp32 (F:113-127).
template <int L>
__device__ __forceinline__ float block_mean(const float* __restrict__ img, int W, int D, int y0, int x0, int c) {
  if constexpr (L == 0) {
    return __ldg(img + ((long long)y0 * W + x0) * D + c);
  } else {
    constexpr int h = 1 << (L - 1);
    const float a = block_mean<L - 1>(img, W, D, y0, x0, c);
    const float b = block_mean<L - 1>(img, W, D, y0, x0 + h, c);
    const float e = block_mean<L - 1>(img, W, D, y0 + h, x0, c);
    const float f = block_mean<L - 1>(img, W, D, y0 + h, x0 + h, c);
    return __fmul_rn(__fadd_rn(__fadd_rn(a, b), __fadd_rn(e, f)), 0.25f);
  }
}

__device__ __forceinline__ float block_mean_rt(const float* img, int W, int D, int y, int x, int c, int L) {
  // (y, x) index the 2^L-downsampled grid
  switch (L) {
    case 0: return block_mean<0>(img, W, D, y, x, c);
    case 1: return block_mean<1>(img, W, D, y << 1, x << 1, c);
    case 2: return block_mean<2>(img, W, D, y << 2, x << 2, c);
    case 3: return block_mean<3>(img, W, D, y << 3, x << 3, c);
    default: return block_mean<4>(img, W, D, y << 4, x << 4, c);
  }
}

// out [B, H>>L, W>>L, D] (rows / columns beyond the last full block are cropped, F:107-110)
__global__ void __launch_bounds__(256) down_kernel(const float* __restrict__ in, float* __restrict__ out, long long n,
                                                   int H, int W, int D, int L) {
  const int h = H >> L, w = W >> L;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % D);
    const long long px = e / D;
    const int x = (int)(px % w), y = (int)((px / w) % h);
    const long long b = px / ((long long)w * h);
    out[e] = block_mean_rt(in + b * (long long)H * W * D, W, D, y, x, c, L);
  }
}

// out [B, H<<L, W<<L, D] = in repeated 2^L times along H and W (tf.repeat, F:152-157)
__global__ void __launch_bounds__(256) up_kernel(const float* __restrict__ in, float* __restrict__ out, long long n,
                                                 int H, int W, int D, int L) {
  const int Ho = H << L, Wo = W << L;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % D);
    const long long px = e / D;
    const int x = (int)(px % Wo), y = (int)((px / Wo) % Ho);
    const long long b = px / ((long long)Wo * Ho);
    out[e] = __ldg(in + ((b * H + (y >> L)) * W + (x >> L)) * D + c);
  }
}

// out [B, H>>lx, W>>lx, 2D]: channels [0,D) = x (minus y when residual), channels [D,2D) = y
__global__ void __launch_bounds__(256) sr_preprocess_kernel(const float* __restrict__ hires, float* __restrict__ out,
                                                            long long n, int H, int W, int D, int lx, int ly,
                                                            int residual) {
  const int h = H >> lx, w = W >> lx, D2 = 2 * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % D2);
    const long long px = e / D2;
    const int x = (int)(px % w), y = (int)((px / w) % h);
    const long long b = px / ((long long)w * h);
    const float* img = hires + b * (long long)H * W * D;
    const int c = k < D ? k : k - D;
    const float yv = block_mean_rt(img, W, D, y >> (ly - lx), x >> (ly - lx), c, ly);
    float v = yv;
    if (k < D) {
      v = block_mean_rt(img, W, D, y, x, c, lx);
      if (residual) v = __fsub_rn(v, yv);
    }
    out[e] = v;
  }
}

// F:199-231 forward, F:287-318 inverse; the constants are evaluated on the host in double and rounded once
struct LogitConsts { float a1b, a, minv, range, b1a; };

__global__ void __launch_bounds__(256) logit_kernel(const float* __restrict__ x, float* __restrict__ out, long long n,
                                                    LogitConsts k, int inverse) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = x[e];
    float r;
    if (!inverse) {
      const float t = __fadd_rn(k.a, __fmul_rn(k.a1b, v));               // a + (1-a) b x
      r = __fdiv_rn(__fsub_rn(logf(__fdiv_rn(t, __fsub_rn(1.0f, t))), k.minv), k.range);
    } else {
      const float t = __fadd_rn(__fmul_rn(v, k.range), k.minv);          // x (max - min) + min
      const float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-t)));        // logistic
      r = __fdiv_rn(__fsub_rn(s, k.a), k.b1a);                          // (s - a) / (b (1 - a))
    }
    out[e] = r;
  }
}

// ---- Philox4x32-10 (Salmon et al. 2011), hand-written ------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ void box_muller(uint32_t u0, uint32_t u1, float& n0, float& n1) {
  // u in (0, 1]: (k + 1) * 2^-32 rounded towards zero stays <= 1 and never reaches 0
  const float a = __fmaf_rz((float)u0, 2.3283064365386963e-10f, 2.3283064365386963e-10f);
  const float b = (float)u1 * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(a));
  float s, c;
  sincospif(2.0f * b, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// out[e] = alpha * x[e] + (1 - alpha) * N(0,1); x == nullptr -> out[e] = N(0,1) (renew_noise).  Elements 4q..4q+3 use
// Philox counter (q + offset) under key `seed`.
__global__ void __launch_bounds__(256) instance_noise_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                             long long n, float alpha, float om,
                                                             unsigned long long seed, unsigned long long offset) {
  const long long nq = (n + 3) >> 2;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    const unsigned long long cq = (unsigned long long)q + offset;
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)cq, (uint32_t)(cq >> 32), 0u, 0u), key);
    float z[4];
    box_muller(rnd.x, rnd.y, z[0], z[1]);
    box_muller(rnd.z, rnd.w, z[2], z[3]);
    const long long e = q << 2;
    if (e + 3 < n) {
      float4 o;
      if (x) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + q);
        o = make_float4(__fadd_rn(__fmul_rn(alpha, v.x), __fmul_rn(om, z[0])), __fadd_rn(__fmul_rn(alpha, v.y), __fmul_rn(om, z[1])),
                        __fadd_rn(__fmul_rn(alpha, v.z), __fmul_rn(om, z[2])), __fadd_rn(__fmul_rn(alpha, v.w), __fmul_rn(om, z[3])));
      } else {
        o = make_float4(z[0], z[1], z[2], z[3]);
      }
      __stcs(reinterpret_cast<float4*>(out) + q, o);
    } else {
      for (int j = 0; e + j < n; ++j) out[e + j] = x ? __fadd_rn(__fmul_rn(alpha, x[e + j]), __fmul_rn(om, z[j])) : z[j];
    }
  }
}

static unsigned grid_for(long long n) {
  // grid-stride: at most 148 SMs x 8 resident CTAs of 256 threads
  const long long want = (n + 255) / 256;
  return (unsigned)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
}

int launch_down(const float* in, float* out, int B, int H, int W, int D, int L, void* stream) {
  const long long n = (long long)B * (H >> L) * (W >> L) * D;
  if (n) down_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, H, W, D, L);
  return (int)cudaGetLastError();
}

int launch_up(const float* in, float* out, int B, int H, int W, int D, int L, void* stream) {
  const long long n = ((long long)B * H * W * D) << (2 * L);
  if (n) up_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, H, W, D, L);
  return (int)cudaGetLastError();
}

int launch_sr_preprocess(const float* hires, float* out, int B, int H, int W, int D, int lx, int ly, int residual,
                         void* stream) {
  const long long n = (long long)B * (H >> lx) * (W >> lx) * 2 * D;
  if (n) sr_preprocess_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(hires, out, n, H, W, D, lx, ly, residual);
  return (int)cudaGetLastError();
}

int launch_logit(const float* x, float* out, long long n, double a, int inverse, void* stream) {
  const double b = (1.0 - 2.0 * a) / (1.0 - a);
  const double minv = log(a / (1.0 - a)), maxv = log((1.0 - a) / a);
  LogitConsts k;
  k.a1b = (float)((1.0 - a) * b);
  k.a = (float)a;
  k.minv = (float)minv;
  k.range = (float)(maxv - minv);
  k.b1a = (float)(b * (1.0 - a));
  if (n) logit_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, out, n, k, inverse);
  return (int)cudaGetLastError();
}

int launch_instance_noise(const float* x, float* out, long long n, double alpha, unsigned long long seed,
                          unsigned long long offset, void* stream) {
  // alpha and (1 - alpha) are Python floats in the reference (C:596-603): each is rounded to fp32 on its own
  if (n)
    instance_noise_kernel<<<grid_for((n + 3) >> 2), 256, 0, (cudaStream_t)stream>>>(x, out, n, (float)alpha,
                                                                                   (float)(1.0 - alpha), seed, offset);
  return (int)cudaGetLastError();
}

}  // namespace cnf
