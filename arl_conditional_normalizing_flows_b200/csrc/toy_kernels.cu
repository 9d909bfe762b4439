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

// ------------------------------------------------------------------------------------------------------------------
// Training: gradient of the toy log_loss w.r.t. every Dense kernel / bias (the tf.GradientTape block of
// cINN_affine.train_step, T:453-482).  One thread owns one sample, one CTA NS samples.  The backward pass walks the
// coupling layers in the reverse of the forward order and needs NO stored activations: a coupling layer leaves u1
// untouched, so its MLPs are re-evaluated from v1 and the layer input is recovered with the inverse law
// u2 = (v2 - b) / exp(A) (T:369-375) - activation-free backward via invertibility.  Per net: forward MLP with the
// hidden activations parked in shared memory, then delta back-propagation; every weight gradient dW_l = H_{l-1}^T D_l is
// a CTA-wide [I x NS] x [NS x I] product from those shared tiles, added to the global buffer with one atomic per entry
// and CTA.
// ------------------------------------------------------------------------------------------------------------------
struct ToyGradArgs {
  const float* xy;
  const float* zy;
  const float* params;
  float* grads;      // layout of params, zero-initialised by the caller
  int B, n_layers_c, num_layers, x_d;
  float lambda_y, inv_B;
  unsigned char order[256];
};

__device__ __forceinline__ float lrelu_grad(float h) { return h > 0.f ? 1.f : CNF_LRELU_SLOPE; }   // sign(h) == sign(pre)

// forward MLP of one net for this thread's sample; hidden activations h_0..h_L -> Hs[l][tid][.] (row stride I + 4)
template <int I>
__device__ __forceinline__ void toy_mlp_store(const float* __restrict__ P, int num_layers, int d1, int d2, const float* in,
                                              float* __restrict__ Hs, int ns, float* out) {
  constexpr int LD = I + 4;
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
  float* row = Hs + (long long)threadIdx.x * LD;
#pragma unroll
  for (int o = 0; o < I; o += 4) st4(row + o, make_float4(h[o], h[o + 1], h[o + 2], h[o + 3]));
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
    row = Hs + ((long long)(l + 1) * ns + threadIdx.x) * LD;
#pragma unroll
    for (int o = 0; o < I; o += 4) st4(row + o, make_float4(h[o], h[o + 1], h[o + 2], h[o + 3]));
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

// backward of one net: dout[2] = dL/d(output) of this thread's sample (0 for absent outputs / padded samples).
// Adds the CTA's weight-gradient contributions to G (global, layout of P) and returns din[2] = dL/d(input).
template <int I>
__device__ __forceinline__ void toy_mlp_backward(const float* __restrict__ P, float* __restrict__ G, int num_layers, int d1,
                                                 const float* in, const float* dout, const float* __restrict__ Hs,
                                                 float* __restrict__ Ds, float* __restrict__ Sm, int ns, float* din) {
  constexpr int LD = I + 4;
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long offL = 3LL * I + (long long)num_layers * (I * I + I);   // WL, bL
  float d[I];
  // ---- output layer: WL [I][2], bL [2]
  {
    const float* WL = P + offL;
    const float* hrow = Hs + ((long long)num_layers * ns + tid) * LD;
#pragma unroll
    for (int i = 0; i < I; ++i) d[i] = (WL[2 * i] * dout[0] + WL[2 * i + 1] * dout[1]) * lrelu_grad(hrow[i]);
    Sm[2 * tid] = dout[0];
    Sm[2 * tid + 1] = dout[1];
    __syncthreads();
    for (int e = tid; e < 2 * I + 2; e += nt) {          // dWL[i][c] = sum_s h_L[s][i] dout[s][c];  dbL[c] = sum_s dout[s][c]
      float acc = 0.f;
      if (e < 2 * I) {
        const int i = e >> 1, c = e & 1;
        const float* hcol = Hs + (long long)num_layers * ns * LD + i;
        for (int s = 0; s < ns; ++s) acc = fmaf(hcol[(long long)s * LD], Sm[2 * s + c], acc);
      } else {
        const int c = e - 2 * I;
        for (int s = 0; s < ns; ++s) acc += Sm[2 * s + c];
      }
      atomicAdd(G + offL + e, acc);
    }
    __syncthreads();
  }
  // ---- hidden layers l = L..1: W_l [I][I] maps h_{l-1} -> pre_l
  for (int l = num_layers; l >= 1; --l) {
    const long long off = 3LL * I + (long long)(l - 1) * (I * I + I);
    float* drow = Ds + (long long)tid * LD;
#pragma unroll
    for (int o = 0; o < I; o += 4) st4(drow + o, make_float4(d[o], d[o + 1], d[o + 2], d[o + 3]));
    // delta_{l-1}[i] = (sum_o W_l[i][o] delta_l[o]) * lrelu'(h_{l-1}[i])
    const float* W = P + off;
    const float* hrow = Hs + ((long long)(l - 1) * ns + tid) * LD;
    float dn[I];
#pragma unroll
    for (int i = 0; i < I; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int o4 = 0; o4 < I; o4 += 4) {
        const float4 w = ld4(W + i * I + o4);
        acc = fmaf(w.x, d[o4], fmaf(w.y, d[o4 + 1], fmaf(w.z, d[o4 + 2], fmaf(w.w, d[o4 + 3], acc))));
      }
      dn[i] = acc * lrelu_grad(hrow[i]);
    }
    __syncthreads();
    // dW_l[i][o..o+3] = sum_s h_{l-1}[s][i] delta_l[s][o..o+3];  db_l[o] = sum_s delta_l[s][o]
    const float* Hl = Hs + (long long)(l - 1) * ns * LD;
    for (int e = tid; e < I * I / 4 + I; e += nt) {
      if (e < I * I / 4) {
        const int i = e / (I / 4), o4 = (e - i * (I / 4)) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < ns; ++s) {
          const float hv = Hl[(long long)s * LD + i];
          const float4 dv = ld4(Ds + (long long)s * LD + o4);
          acc.x = fmaf(hv, dv.x, acc.x); acc.y = fmaf(hv, dv.y, acc.y);
          acc.z = fmaf(hv, dv.z, acc.z); acc.w = fmaf(hv, dv.w, acc.w);
        }
        float* gp = G + off + i * I + o4;
        atomicAdd(gp, acc.x); atomicAdd(gp + 1, acc.y); atomicAdd(gp + 2, acc.z); atomicAdd(gp + 3, acc.w);
      } else {
        const int o = e - I * I / 4;
        float acc = 0.f;
        for (int s = 0; s < ns; ++s) acc += Ds[(long long)s * LD + o];
        atomicAdd(G + off + I * I + o, acc);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < I; ++i) d[i] = dn[i];
  }
  // ---- first layer: W0 [2][I] (rows >= d1 unused), b0 [I]
  {
    float* drow = Ds + (long long)tid * LD;
#pragma unroll
    for (int o = 0; o < I; o += 4) st4(drow + o, make_float4(d[o], d[o + 1], d[o + 2], d[o + 3]));
    Sm[2 * tid] = in[0];
    Sm[2 * tid + 1] = d1 > 1 ? in[1] : 0.f;
    const float* W0 = P;
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int o = 0; o < I; ++o) {
      g0 = fmaf(W0[o], d[o], g0);
      g1 = fmaf(W0[I + o], d[o], g1);
    }
    din[0] = g0;
    din[1] = d1 > 1 ? g1 : 0.f;
    __syncthreads();
    for (int e = tid; e < 3 * I; e += nt) {              // dW0[r][o] = sum_s in[s][r] delta_0[s][o];  db0[o] = sum_s delta_0[s][o]
      const int r = e / I, o = e - r * I;
      float acc = 0.f;
      if (r < 2) {
        if (r < d1)
          for (int s = 0; s < ns; ++s) acc = fmaf(Sm[2 * s + r], Ds[(long long)s * LD + o], acc);
      } else {
        for (int s = 0; s < ns; ++s) acc += Ds[(long long)s * LD + o];
      }
      if (r >= 2 || r < d1) atomicAdd(G + (r < 2 ? r * I + o : 2 * I + o), acc);
    }
    __syncthreads();
  }
}

template <int I>
__global__ void __launch_bounds__(128) toy_grad_kernel(const ToyGradArgs a) {
  extern __shared__ __align__(16) float sp[];
  constexpr int LD = I + 4;
  const int ns = blockDim.x;
  const long long net_sz = toy_net_size(I, a.num_layers);
  float* Wn = sp;                                              // one net's parameters
  float* Hs = Wn + ((net_sz + 3) & ~3LL);                      // [L + 1][ns][LD]
  float* Ds = Hs + (long long)(a.num_layers + 1) * ns * LD;    // [ns][LD]
  float* Sm = Ds + (long long)ns * LD;                         // [ns][2]
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = s < a.B;
  float x[3] = {0.f, 0.f, 0.f}, g[3] = {0.f, 0.f, 0.f};
  if (valid) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      x[c] = a.zy[3 * s + c];
      if (c < a.x_d) {
        g[c] = x[c] * a.inv_B;                                 // d/dz of -mean(-z^2 / 2)
      } else {
        const float df = x[c] - a.xy[3 * s + c];
        g[c] = a.lambda_y * a.inv_B * (df > 0.f ? 1.f : df < 0.f ? -1.f : 0.f);   // d/dy of mean(lambda |y - y'|)
      }
    }
  }
  const float g_ld = valid ? -a.inv_B : 0.f;                   // d/d(log_detJ) of -mean(log_detJ)
  for (int i = 0; i < a.n_layers_c; ++i) {                     // reverse of the forward order n-1..0 (T:295)
    const int j = a.order[i];
    const int mk = j % 6;
    const int d1 = mk < 3 ? 1 : 2, d2 = 3 - d1;
    const float* Pj = a.params + 2 * net_sz * j;
    float* Gj = a.grads + 2 * net_sz * j;
    float v1[2] = {x[c_m1[mk][0]], x[c_m1[mk][1]]};
    float v2[2] = {x[c_m2[mk][0]], x[c_m2[mk][1]]};
    float g2[2] = {g[c_m2[mk][0]], d2 > 1 ? g[c_m2[mk][1]] : 0.f};
    float bb[2], raw[2], din_b[2], din_A[2];
    // ---- net b: t = b(u1); dL/db = dL/dv2
    __syncthreads();
    for (long long t = threadIdx.x; t < net_sz / 4; t += blockDim.x) st4(Wn + 4 * t, ld4(Pj + net_sz + 4 * t));
    __syncthreads();
    toy_mlp_store<I>(Wn, a.num_layers, d1, d2, v1, Hs, ns, bb);
    __syncthreads();
    toy_mlp_backward<I>(Wn, Gj + net_sz, a.num_layers, d1, v1, g2, Hs, Ds, Sm, ns, din_b);
    // ---- net A: A = tanh(raw); u2 = (v2 - b) / exp(A); dL/dA = dL/dv2 exp(A) u2 + dL/dlogdet
    for (long long t = threadIdx.x; t < net_sz / 4; t += blockDim.x) st4(Wn + 4 * t, ld4(Pj + 4 * t));
    __syncthreads();
    toy_mlp_store<I>(Wn, a.num_layers, d1, d2, v1, Hs, ns, raw);
    __syncthreads();
    float dA[2], u2[2], e[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float A = tanhf(raw[c]);
      e[c] = (c < d2) ? expf(A) : 1.f;
      u2[c] = __fmul_rn(__frcp_rn(e[c]), __fsub_rn(v2[c], bb[c]));
      dA[c] = (c < d2) ? (g2[c] * e[c] * u2[c] + g_ld) * (1.f - A * A) : 0.f;
    }
    toy_mlp_backward<I>(Wn, Gj, a.num_layers, d1, v1, dA, Hs, Ds, Sm, ns, din_A);
    // ---- state and gradient in front of this layer
    x[c_m2[mk][0]] = u2[0];
    g[c_m2[mk][0]] = g2[0] * e[0];
    if (d2 > 1) {
      x[c_m2[mk][1]] = u2[1];
      g[c_m2[mk][1]] = g2[1] * e[1];
    }
    g[c_m1[mk][0]] += din_b[0] + din_A[0];
    if (d1 > 1) g[c_m1[mk][1]] += din_b[1] + din_A[1];
  }
}

int launch_toy_grad(const float* xy, const float* zy, const float* params, float* grads, const int* mask_idx_host,
                    int n_layers_c, int width, int num_layers, int x_d, double lambda_y, int B, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_layers_c > 256) return (int)cudaErrorInvalidValue;
  if (B <= 0) return 0;
  ToyGradArgs a;
  a.xy = xy; a.zy = zy; a.params = params; a.grads = grads;
  a.B = B; a.n_layers_c = n_layers_c; a.num_layers = num_layers; a.x_d = x_d;
  a.lambda_y = (float)lambda_y; a.inv_B = 1.0f / (float)B;
  for (int i = 0; i < n_layers_c; ++i) a.order[i] = (unsigned char)mask_idx_host[i];
  const long long net_sz = toy_net_size(width, num_layers);
  // samples per CTA: as many (128, 64, 32) as keep one net's weights + the activation tiles inside shared memory
  int ns = 128;
  size_t smem = 0;
  for (; ns >= 32; ns >>= 1) {
    smem = (size_t)(((net_sz + 3) & ~3LL) + (long long)(num_layers + 2) * ns * (width + 4) + 2LL * ns) * sizeof(float);
    if (smem <= 227 * 1024) break;
  }
  if (ns < 32) return (int)cudaErrorInvalidConfiguration;
  const int grid = (B + ns - 1) / ns;
#define TOY_GCASE(I)                                                                                       \
  case I: {                                                                                                \
    CU_TRY(cudaFuncSetAttribute(toy_grad_kernel<I>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                (int)(smem > 48 * 1024 ? smem : 48 * 1024)));                              \
    toy_grad_kernel<I><<<grid, ns, smem, st>>>(a);                                                         \
    break;                                                                                                 \
  }
  switch (width) {
    TOY_GCASE(8)
    TOY_GCASE(16)
    TOY_GCASE(32)
    TOY_GCASE(64)
    default: return (int)cudaErrorInvalidConfiguration;
  }
#undef TOY_GCASE
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
