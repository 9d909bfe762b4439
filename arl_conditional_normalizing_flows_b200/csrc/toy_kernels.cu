// Toy dense cINN (BASELINE config 1): the whole flow of TOYcINN_make_model.py in ONE kernel.
//   coupling_layer MLPs  T:29-97   Dense+LeakyReLU(0.3) x (num_layers+1), Dense; A ends in tanh
//   masks                T:154-166 u1 index sets {0},{1},{2},{0,1},{0,2},{1,2}; u2 = complement
//   cINN_affine.call     T:248-402 direction -1: layers n-1..0, v2 = exp(A)*u2 + b,
//                                  log_detJ[b] += log(prod exp(A)); direction +1: u2 = (v2-b)/exp(A)
//   log_loss             T:404-451
// One thread owns one sample (3 floats of state, the MLP activations in registers); the CTA stages
// each coupling layer's weights (both nets, ~50 KB at width 32) in shared memory once and every
// thread reads them as broadcast LDS.128.
#include <cuda_runtime.h>

#include "cnf_internal.h"
#include "device_utils.cuh"

namespace cnf {

#define CU_TRY(x)                          \
  do {                                     \
    cudaError_t e_ = (x);                  \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

// Per-net parameter slots (floats), identical for every layer so that offsets are uniform:
//   W0 [2][I] (rows beyond dim(u1) unused), b0 [I], num_layers x { W [I][I], b [I] }, WL [I][2]
//   (columns beyond dim(u2) unused), bL [2] (+2 pad)
__host__ __device__ inline long long toy_net_size(int I, int num_layers) {
  return 2LL * I + I + (long long)num_layers * ((long long)I * I + I) + 2LL * I + 4;
}

struct ToyArgs {
  const float* u;
  const float* params;
  float* v;
  float* logdet;
  int B, n_layers_c, num_layers, direction;
  unsigned char order[256];  // mask_indices (T:206-217)
};

__constant__ int c_m1[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {0, 2}, {1, 2}};  // T:154-159 (dim1: 1,1,1,2,2,2)
__constant__ int c_m2[6][2] = {{1, 2}, {0, 2}, {0, 1}, {2, 2}, {1, 1}, {0, 0}};  // T:160-165

template <int I>
__device__ __forceinline__ void toy_mlp(const float* __restrict__ P, int num_layers, int d1, int d2,
                                        const float* in, float* out) {
  float h[I], g[I];
  const float* W0 = P;
  const float* b0 = P + 2 * I;
#pragma unroll
  for (int o = 0; o < I; ++o) {
    float a = b0[o];
    a = fmaf(in[0], W0[o], a);
    if (d1 > 1) a = fmaf(in[1], W0[I + o], a);
    h[o] = lrelu(a);
  }
  const float* Pl = P + 3 * I;
  for (int l = 0; l < num_layers; ++l) {
    const float* W = Pl;
    const float* bb = Pl + I * I;
#pragma unroll
    for (int o = 0; o < I; ++o) g[o] = bb[o];
#pragma unroll
    for (int i = 0; i < I; ++i) {
#pragma unroll
      for (int o4 = 0; o4 < I; o4 += 4) {
        const float4 w = ld4(W + i * I + o4);
        g[o4] = fmaf(h[i], w.x, g[o4]);
        g[o4 + 1] = fmaf(h[i], w.y, g[o4 + 1]);
        g[o4 + 2] = fmaf(h[i], w.z, g[o4 + 2]);
        g[o4 + 3] = fmaf(h[i], w.w, g[o4 + 3]);
      }
    }
#pragma unroll
    for (int o = 0; o < I; ++o) h[o] = lrelu(g[o]);
    Pl += I * I + I;
  }
  const float* WL = Pl;
  const float* bL = Pl + 2 * I;
  float o0 = bL[0], o1 = bL[1];
#pragma unroll
  for (int i = 0; i < I; ++i) {
    o0 = fmaf(h[i], WL[2 * i], o0);
    o1 = fmaf(h[i], WL[2 * i + 1], o1);
  }
  out[0] = o0;
  out[1] = d2 > 1 ? o1 : 0.f;
}

template <int I>
__global__ void __launch_bounds__(128) toy_flow_kernel(const ToyArgs a) {
  extern __shared__ __align__(16) float sp[];
  const long long net_sz = toy_net_size(I, a.num_layers);
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = s < a.B;
  float x[3] = {0.f, 0.f, 0.f};
  if (valid) {
    x[0] = a.u[3 * s];
    x[1] = a.u[3 * s + 1];
    x[2] = a.u[3 * s + 2];
  }
  float ld = 0.f;
  for (int step = 0; step < a.n_layers_c; ++step) {
    const int i = a.direction == 1 ? step : a.n_layers_c - 1 - step;   // T:295
    const int j = a.order[i];
    const int mk = j % 6;
    const int d1 = mk < 3 ? 1 : 2, d2 = 3 - d1;
    __syncthreads();
    {
      const float* src = a.params + 2 * net_sz * j;
      for (long long t = threadIdx.x; t < 2 * net_sz / 4; t += blockDim.x)
        st4(sp + 4 * t, ld4(src + 4 * t));
    }
    __syncthreads();
    float u1[2], u2[2], A[2], bb[2];
    u1[0] = x[c_m1[mk][0]];
    u1[1] = x[c_m1[mk][1]];
    u2[0] = x[c_m2[mk][0]];
    u2[1] = x[c_m2[mk][1]];
    toy_mlp<I>(sp, a.num_layers, d1, d2, u1, A);
    toy_mlp<I>(sp + net_sz, a.num_layers, d1, d2, u1, bb);
    A[0] = tanhf(A[0]);
    A[1] = tanhf(A[1]);
    const float e0 = expf(A[0]), e1 = d2 > 1 ? expf(A[1]) : 1.f;
    float t0, t1;
    if (a.direction == 1) {                       // T:369-375
      t0 = __fmul_rn(__frcp_rn(e0), __fsub_rn(u2[0], bb[0]));
      t1 = __fmul_rn(__frcp_rn(e1), __fsub_rn(u2[1], bb[1]));
    } else {                                      // T:379-387
      t0 = __fadd_rn(__fmul_rn(e0, u2[0]), bb[0]);
      t1 = __fadd_rn(__fmul_rn(e1, u2[1]), bb[1]);
      ld += logf(d2 > 1 ? __fmul_rn(e0, e1) : e0);
    }
    x[c_m2[mk][0]] = t0;
    if (d2 > 1) x[c_m2[mk][1]] = t1;
  }
  if (valid) {
    a.v[3 * s] = x[0];
    a.v[3 * s + 1] = x[1];
    a.v[3 * s + 2] = x[2];
    if (a.logdet) a.logdet[s] = a.direction == -1 ? ld : 0.f;
  }
}

int launch_toy(const float* u, const float* params, const int* mask_idx_host, int n_layers_c, int width,
               int num_layers, int direction, float* v, float* logdet, int B, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_layers_c > 256) return (int)cudaErrorInvalidValue;
  ToyArgs a;
  a.u = u; a.params = params; a.v = v; a.logdet = logdet;
  a.B = B; a.n_layers_c = n_layers_c; a.num_layers = num_layers; a.direction = direction;
  for (int i = 0; i < n_layers_c; ++i) a.order[i] = (unsigned char)mask_idx_host[i];
  const size_t smem = 2 * (size_t)toy_net_size(width, num_layers) * sizeof(float);
  if (smem > 227 * 1024) return (int)cudaErrorInvalidConfiguration;
  const int grid = (B + 127) / 128;
  if (grid == 0) return 0;
#define TOY_CASE(I)                                                                                        \
  case I: {                                                                                                \
    CU_TRY(cudaFuncSetAttribute(toy_flow_kernel<I>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                (int)(smem > 48 * 1024 ? smem : 48 * 1024)));                              \
    toy_flow_kernel<I><<<grid, 128, smem, st>>>(a);                                                        \
    break;                                                                                                 \
  }
  switch (width) {
    TOY_CASE(8)
    TOY_CASE(16)
    TOY_CASE(32)
    TOY_CASE(64)
    default: return (int)cudaErrorInvalidConfiguration;
  }
#undef TOY_CASE
  return (int)cudaGetLastError();
}

// T:436-451: ll_z = log N(z; 0, I_{x_d}) per sample, ll_y = -lambda_y sum |y - y'|,
// loss = -mean(ll_z + ll_y + log_detJ), plus the three component means.
__global__ void __launch_bounds__(256) toy_loss_kernel(const float* __restrict__ zy, const float* __restrict__ xy,
                                                       const float* __restrict__ logdet, int B, int x_d,
                                                       double lambda_y, float* __restrict__ ll_z,
                                                       float* __restrict__ ll_y, float* __restrict__ loss4) {
  __shared__ double red[3][256];
  double sa = 0, sb = 0, sc = 0;
  for (int s = threadIdx.x; s < B; s += 256) {
    float sz = 0.f, sy = 0.f;
    for (int c = 0; c < 3; ++c) {
      const float z = zy[3 * s + c];
      if (c < x_d) sz = fmaf(z, z, sz);
      else sy += fabsf(z - xy[3 * s + c]);
    }
    const float lz = (float)(-0.5 * (double)sz - 0.5 * (double)x_d * 1.8378770664093453);
    const float ly = (float)(-lambda_y * (double)sy);
    ll_z[s] = lz;
    ll_y[s] = ly;
    sa += lz; sb += ly; sc += logdet[s];
  }
  red[0][threadIdx.x] = sa; red[1][threadIdx.x] = sb; red[2][threadIdx.x] = sc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mz = red[0][0] / B, my = red[1][0] / B, md = red[2][0] / B;
    loss4[0] = (float)(-(mz + my + md));
    loss4[1] = (float)(-mz);
    loss4[2] = (float)(-my);
    loss4[3] = (float)(-md);
  }
}

int launch_toy_loss(const float* zy, const float* xy, const float* logdet, int B, int x_d, double lambda_y,
                    float* ll_z, float* ll_y, float* loss4, void* stream) {
  toy_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(zy, xy, logdet, B, x_d, lambda_y, ll_z, ll_y, loss4);
  return (int)cudaGetLastError();
}

}  // namespace cnf
