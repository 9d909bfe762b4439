// HBM-bound kernels around the s/t networks: the standalone fused affine-coupling law with mask
// addressing and per-sample log-det, the Gaussian-prior / L1 log-likelihood reduction, and the
// pure index permutations the reference materialises (mask, decompress_mask, space_to_depth).
// Reference: conv_cINN_make_model.py M:155-217, M:500-761, M:763-1073, M:1215-1253, M:1307-1326,
// M:1800-1848.
#include <cuda_runtime.h>

#include <cstdint>

#include "cnf_internal.h"
#include "device_utils.cuh"

namespace cnf {

#define CU_TRY(x)                          \
  do {                                     \
    cudaError_t e_ = (x);                  \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Standalone coupling law.  One pass: read u (4N B), s and t (2N + 2N B, compressed halves), write
// v (4N B) = 12 bytes per element of u; the pass-through half is copied, the complement half is
// transformed in registers; sum(s) per sample is reduced warp-shuffle -> block -> one atomic.
// ---------------------------------------------------------------------------------------------
struct LawGeom {
  int H, W, D, mask, inverse;
  int h, w, c2;  // compressed shape of the complement half
};

__device__ __forceinline__ bool law_locate(const LawGeom& g, int e, int& sidx) {
  // e = flat index inside one sample; returns true if the element is in the complement (transformed) half
  const int c = e % g.D;
  const int px = e / g.D;
  const int x = px % g.W, y = px / g.W;
  if (g.mask < 2) {
    const int par = (y + x) & 1;
    if (par == g.mask) return false;  // mask 0 keeps (y+x) even, mask 1 keeps odd (M:632-660)
    sidx = (((y >> 1) * g.w) + (x >> 1)) * g.c2 + (y & 1) * g.D + c;  // M:726-748 of the complement
    return true;
  }
  const int par = c & 1;
  if (par == g.mask - 2) return false;
  sidx = ((y * g.w) + x) * g.c2 + (c >> 1);
  return true;
}

__device__ __forceinline__ float law_apply(float u, float s, float t, int inverse) {
  if (!inverse) return __fadd_rn(__fmul_rn(expf(s), u), t);       // M:1307, M:1230-1231
  return __fmul_rn(__frcp_rn(expf(s)), __fsub_rn(u, t));           // M:1379, M:1250-1251
}

template <bool VEC>
__global__ void __launch_bounds__(256) coupling_law_kernel(const float* __restrict__ u, const float* __restrict__ s,
                                                           const float* __restrict__ t, float* __restrict__ v,
                                                           float* __restrict__ logdet, LawGeom g, int n_per,
                                                           int ns_per) {
  __shared__ float red[64];
  const int b = blockIdx.y;
  const float* ub = u + (long long)b * n_per;
  float* vb = v + (long long)b * n_per;
  const float* sb = s + (long long)b * ns_per;
  const float* tb = t + (long long)b * ns_per;
  float ld = 0.f;
  if (VEC) {
    const int n4 = n_per >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      float4 x = __ldcs(reinterpret_cast<const float4*>(ub) + i);
      float o[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int si;
        if (law_locate(g, 4 * i + j, si)) {
          const float sv = __ldg(sb + si), tv = __ldg(tb + si);
          o[j] = law_apply(o[j], sv, tv, g.inverse);
          ld += sv;
        }
      }
      __stcs(reinterpret_cast<float4*>(vb) + i, make_float4(o[0], o[1], o[2], o[3]));
    }
  } else {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_per; e += gridDim.x * blockDim.x) {
      float x = ub[e];
      int si;
      if (law_locate(g, e, si)) {
        const float sv = sb[si], tv = tb[si];
        x = law_apply(x, sv, tv, g.inverse);
        ld += sv;
      }
      vb[e] = x;
    }
  }
  if (logdet) {
    double d1, d2;
    block_sum2(ld, 0.f, red, d1, d2);
    if (threadIdx.x == 0) atomicAdd(logdet + b, (float)d1);
  }
}

// Fast paths (the shapes the flow produces): 4 consecutive channels of one pixel per thread, no integer
// division by runtime values on the channel masks (even D: the compressed index of element e is e >> 1),
// 128-bit loads of u and 64/128-bit loads of s and t, LAW_U independent items per thread in flight.
constexpr int LAW_U = 4;

template <int MASK>   // 2 or 3, D % 4 == 0
__global__ void __launch_bounds__(256) coupling_law_chan_kernel(const float4* __restrict__ u, const float2* __restrict__ s,
                                                                const float2* __restrict__ t, float4* __restrict__ v,
                                                                float* __restrict__ logdet, int n4, int inverse) {
  __shared__ float red[64];
  const int b = blockIdx.y;
  const float4* ub = u + (long long)b * n4;
  const float2* sb = s + (long long)b * n4;
  const float2* tb = t + (long long)b * n4;
  float4* vb = v + (long long)b * n4;
  float4 x[LAW_U];
  float2 sv[LAW_U], tv[LAW_U];
  int idx[LAW_U];
#pragma unroll
  for (int k = 0; k < LAW_U; ++k) {
    idx[k] = (blockIdx.x * LAW_U + k) * 256 + threadIdx.x;
    if (idx[k] < n4) {
      x[k] = __ldcs(ub + idx[k]);
      sv[k] = __ldcs(sb + idx[k]);
      tv[k] = __ldcs(tb + idx[k]);
    }
  }
  float ld = 0.f;
#pragma unroll
  for (int k = 0; k < LAW_U; ++k) {
    if (idx[k] < n4) {
      float4 o = x[k];
      if (MASK == 2) {   // even channels pass through, odd channels are transformed
        o.y = law_apply(o.y, sv[k].x, tv[k].x, inverse);
        o.w = law_apply(o.w, sv[k].y, tv[k].y, inverse);
      } else {
        o.x = law_apply(o.x, sv[k].x, tv[k].x, inverse);
        o.z = law_apply(o.z, sv[k].y, tv[k].y, inverse);
      }
      ld += sv[k].x + sv[k].y;
      __stcs(vb + idx[k], o);
    }
  }
  if (logdet) {
    double d1, d2;
    block_sum2(ld, 0.f, red, d1, d2);
    if (threadIdx.x == 0) atomicAdd(logdet + b, (float)d1);
  }
}

// checkerboard masks, D % 4 == 0: a float4 lies inside one pixel, so it is either copied or transformed whole
__global__ void __launch_bounds__(256) coupling_law_cb_kernel(const float4* __restrict__ u, const float* __restrict__ s,
                                                              const float* __restrict__ t, float4* __restrict__ v,
                                                              float* __restrict__ logdet, LawGeom g, int n4, int ns_per) {
  __shared__ float red[64];
  const int b = blockIdx.y;
  const float4* ub = u + (long long)b * n4;
  const float* sb = s + (long long)b * ns_per;
  const float* tb = t + (long long)b * ns_per;
  float4* vb = v + (long long)b * n4;
  const int q = g.D >> 2;
  float ld = 0.f;
#pragma unroll
  for (int k = 0; k < LAW_U; ++k) {
    const int i = (blockIdx.x * LAW_U + k) * 256 + threadIdx.x;
    if (i >= n4) continue;
    float4 o = __ldcs(ub + i);
    const int px = i / q, cq = i - px * q;
    const int y = px / g.W, x = px - y * g.W;
    if (((y + x) & 1) != g.mask) {
      const int si = (((y >> 1) * g.w) + (x >> 1)) * g.c2 + (y & 1) * g.D + 4 * cq;
      const float4 sv = __ldcs(reinterpret_cast<const float4*>(sb + si));
      const float4 tv = __ldcs(reinterpret_cast<const float4*>(tb + si));
      o.x = law_apply(o.x, sv.x, tv.x, g.inverse);
      o.y = law_apply(o.y, sv.y, tv.y, g.inverse);
      o.z = law_apply(o.z, sv.z, tv.z, g.inverse);
      o.w = law_apply(o.w, sv.w, tv.w, g.inverse);
      ld += (sv.x + sv.y) + (sv.z + sv.w);
    }
    __stcs(vb + i, o);
  }
  if (logdet) {
    double d1, d2;
    block_sum2(ld, 0.f, red, d1, d2);
    if (threadIdx.x == 0) atomicAdd(logdet + b, (float)d1);
  }
}

int launch_coupling_law(const float* u, const float* s, const float* t, float* v, float* logdet, int B, int H,
                        int W, int D, int mask, int inverse, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LawGeom g;
  g.H = H; g.W = W; g.D = D; g.mask = mask; g.inverse = inverse;
  const int mc = mask ^ 1;  // complement: 0<->1, 2<->3 (M:426-433)
  if (mc < 2) { g.h = H / 2; g.w = W / 2; g.c2 = 2 * D; }
  else { g.h = H; g.w = W; g.c2 = mc == 2 ? (D + 1) / 2 : D / 2; }
  const int n_per = H * W * D, ns_per = g.h * g.w * g.c2;
  if (logdet) CU_TRY(cudaMemsetAsync(logdet, 0, sizeof(float) * B, st));
  if (D % 4 == 0 && B <= 65535 && ((uintptr_t)s & 15) == 0 && ((uintptr_t)t & 15) == 0) {
    const int n4 = n_per / 4;
    dim3 grid((n4 + 256 * LAW_U - 1) / (256 * LAW_U), B);
    const float4* u4 = reinterpret_cast<const float4*>(u);
    float4* v4 = reinterpret_cast<float4*>(v);
    if (mask == 2)
      coupling_law_chan_kernel<2><<<grid, 256, 0, st>>>(u4, reinterpret_cast<const float2*>(s), reinterpret_cast<const float2*>(t), v4, logdet, n4, inverse);
    else if (mask == 3)
      coupling_law_chan_kernel<3><<<grid, 256, 0, st>>>(u4, reinterpret_cast<const float2*>(s), reinterpret_cast<const float2*>(t), v4, logdet, n4, inverse);
    else
      coupling_law_cb_kernel<<<grid, 256, 0, st>>>(u4, s, t, v4, logdet, g, n4, ns_per);
    return (int)cudaGetLastError();
  }
  const bool vec = (n_per % 4) == 0;
  const int work = vec ? n_per / 4 : n_per;
  int bx = (work + 256 * 4 - 1) / (256 * 4);  // ~4 items per thread
  if (bx < 1) bx = 1;
  dim3 grid(bx, B);
  if (vec) coupling_law_kernel<true><<<grid, 256, 0, st>>>(u, s, t, v, logdet, g, n_per, ns_per);
  else coupling_law_kernel<false><<<grid, 256, 0, st>>>(u, s, t, v, logdet, g, n_per, ns_per);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Prior / L1 reduction (M:1826-1848): per sample
//   ll_z = sum_{h,w} ( -1/2 sum_{c<x_d} z^2 - 1/2 x_d ln 2pi ),  ll_y = -lambda_y sum |y - y'|
// then the four batch-mean scalars the reference returns.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prior_kernel(const float* __restrict__ zy, const float* __restrict__ xy,
                                                    long long n_per, int D, int x_d, double lambda_y,
                                                    double hw, float* __restrict__ ll_z, float* __restrict__ ll_y) {
  __shared__ float red[64];
  const int b = blockIdx.x;
  const float* z = zy + (long long)b * n_per;
  const float* x = xy + (long long)b * n_per;
  float sz = 0.f, sy = 0.f;
  for (long long e = threadIdx.x; e < n_per; e += blockDim.x) {
    const int c = (int)(e % D);
    const float zv = z[e];
    if (c < x_d) sz = fmaf(zv, zv, sz);
    else sy += fabsf(zv - x[e]);
  }
  double d1, d2;
  block_sum2(sz, sy, red, d1, d2);
  if (threadIdx.x == 0) {
    ll_z[b] = (float)(-0.5 * d1 - 0.5 * (double)x_d * 1.8378770664093453 * hw);  // ln(2*pi)
    ll_y[b] = (float)(-lambda_y * d2);
  }
}

__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ ll_z, const float* __restrict__ ll_y,
                                                            const float* __restrict__ logdet, int B,
                                                            float* __restrict__ loss4) {
  __shared__ double red[3][256];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < B; i += 256) {
    a += ll_z[i];
    b += ll_y[i];
    c += logdet[i];
  }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = b; red[2][threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mz = red[0][0] / B, my = red[1][0] / B, md = red[2][0] / B;
    loss4[0] = (float)(-(mz + my + md));  // M:1840-1846
    loss4[1] = (float)(-mz);
    loss4[2] = (float)(-my);
    loss4[3] = (float)(-md);
  }
}

int launch_prior_loss(const float* zy, const float* xy, const float* logdet, int B, int64_t HW, int D, int x_d,
                      double lambda_y, float* ll_z, float* ll_y, float* loss4, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  prior_kernel<<<B, 256, 0, st>>>(zy, xy, (long long)HW * D, D, x_d, lambda_y, (double)HW, ll_z, ll_y);
  CU_TRY(cudaGetLastError());
  if (loss4) {
    loss_finalize_kernel<<<1, 256, 0, st>>>(ll_z, ll_y, logdet, B, loss4);
    CU_TRY(cudaGetLastError());
  }
  return 0;
}

// out[b] = per-sample log-det; with_mean: out[B] = their batch mean (the reference's scalar, M:1325-1326 / Q1), summed
// in fp64 in a fixed order by block 0
__global__ void __launch_bounds__(256) logdet_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out, int B,
                                                              int with_mean) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) out[i] = (float)acc[i];
  if (with_mean && blockIdx.x == 0) {
    __shared__ double part[256];
    double s = 0.0;
    for (int b = threadIdx.x; b < B; b += 256) s += (double)(float)acc[b];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[B] = (float)(part[0] / (double)B);
  }
}

int launch_logdet_finalize(const double* acc, float* out, int B, void* stream, int with_mean) {
  logdet_finalize_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(acc, out, B, with_mean);
  return (int)cudaGetLastError();
}

int launch_copy(const float* src, float* dst, int64_t n, void* stream) {
  return (int)cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Index permutations as stand-alone ops (bit-exact; the flow itself never materialises them).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_mask(int m, int y, int x, int c) {
  return m < 2 ? (((y + x) & 1) == m) : ((c & 1) == m - 2);
}

__global__ void mask_uncompressed_kernel(const float* __restrict__ uv, float* __restrict__ out, long long n, int H,
                                         int W, int D, int m) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int c = (int)(e % D);
  const long long px = e / D;
  const int x = (int)(px % W), y = (int)((px / W) % H);
  // literal 0/1 multiply accumulated into a zero output (einsum at M:715-717): masked-out elements are
  // +0 (NaN for non-finite input), kept elements are u*1+0, i.e. a copy except -0.0 -> +0.0 (Q8)
  out[e] = __fadd_rn(__fmul_rn(in_mask(m, y, x, c) ? 1.0f : 0.0f, uv[e]), 0.0f);
}

__global__ void mask_compressed_kernel(const float* __restrict__ uv, float* __restrict__ out, long long n, FlowView v,
                                       int m, int h, int w, int dc) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int k = (int)(e % dc);
  const long long px = e / dc;
  const int j = (int)(px % w), i = (int)((px / w) % h), b = (int)(px / ((long long)w * h));
  out[e] = uv[comp_off(v, m, b, i, j, k)];
}

__global__ void decompress_kernel(const float* __restrict__ uvc, float* __restrict__ out, long long n, int H, int W,
                                  int D, int m, int h, int w, int dc) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int c = (int)(e % D);
  const long long px = e / D;
  const int x = (int)(px % W), y = (int)((px / W) % H), b = (int)(px / ((long long)W * H));
  float val = 0.f;
  if (in_mask(m, y, x, c)) {
    long long si;
    if (m < 2) si = (((long long)b * h + (y >> 1)) * w + (x >> 1)) * dc + (y & 1) * D + c;
    else si = (((long long)b * h + y) * w + x) * dc + (c >> 1);
    val = uvc[si];
  }
  out[e] = val;
}

int launch_mask(const float* uv, float* out, int B, int H, int W, int D, int m, int compress, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!compress) {
    const long long n = (long long)B * H * W * D;
    if (n) mask_uncompressed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(uv, out, n, H, W, D, m);
  } else {
    int h, w, dc;
    if (m < 2) { h = H / 2; w = W / 2; dc = 2 * D; }
    else { h = H; w = W; dc = m == 2 ? (D + 1) / 2 : D / 2; }
    const long long n = (long long)B * h * w * dc;
    FlowView v = make_view(const_cast<float*>(uv), H, W, D, 0);
    if (n) mask_compressed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(uv, out, n, v, m, h, w, dc);
  }
  return (int)cudaGetLastError();
}

int launch_decompress(const float* uvc, float* out, int B, int H, int W, int D, int m, void* stream) {
  int h, w, dc;
  if (m < 2) { h = H / 2; w = W / 2; dc = 2 * D; }
  else { h = H; w = W; dc = m == 2 ? (D + 1) / 2 : D / 2; }
  const long long n = (long long)B * H * W * D;
  if (n) decompress_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(uvc, out, n, H, W, D, m, h, w, dc);
  return (int)cudaGetLastError();
}

// space_to_depth (block 2, NHWC): out[b,i,j,(dy*2+dx)*C+c] = in[b,2i+dy,2j+dx,c]; inverse swaps roles.
__global__ void s2d_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int H, int W, int C,
                           int inverse) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  // e indexes the SQUEEZED tensor [B, H/2, W/2, 4C]
  const int C4 = 4 * C, h = H / 2, w = W / 2;
  const int k = (int)(e % C4);
  const long long px = e / C4;
  const int j = (int)(px % w), i = (int)((px / w) % h);
  const long long b = px / ((long long)w * h);
  const int c = k % C, dx = (k / C) & 1, dy = k / (2 * C);
  const long long big = ((b * H + 2 * i + dy) * W + 2 * j + dx) * C + c;
  if (!inverse) out[e] = in[big];
  else out[big] = in[e];
}

int launch_space_to_depth(const float* in, float* out, int B, int H, int W, int C, int inverse, void* stream) {
  // H, W, C always describe the UNSQUEEZED tensor
  const long long n = (long long)B * H * W * C;
  if (n) s2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, H, W, C, inverse);
  return (int)cudaGetLastError();
}

}  // namespace cnf
