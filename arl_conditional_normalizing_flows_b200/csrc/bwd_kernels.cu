// Backward pass of one coupling layer (training step, conv_cINN_make_model.py M:1850-1880: the reference
// differentiates cFlow.log_loss with tf.GradientTape; here every gradient is a hand-written kernel).
//
// Forward (per net, see stnet_kernels.cu):  x0 = stem(u1c);  R x { y1 = pw1(LN1(lrelu(x_r))),
//   y2 = gconv(LN2(lrelu(y1))), x_{r+1} = x_r + pw2(LN3(lrelu(y2))) };  raw = head(LNf(lrelu(x_R)));
//   A = w*tanh(raw_A), t = raw_b;  v2 = exp(A)*u2 + t;  logdet_b += sum A.
// Backward, given G = dL/dv in the flow-buffer layout (updated in place to dL/du):
//   head_bwd_kernel      dA, dt -> draw (both nets), dL/du2 = dv2*exp(A) in place, d(tanh scale)
//   wgrad3_small_kernel  3x3 weight/bias gradients of head and stem (LN-on-load / mask gather on load)
//   conv3t_small_kernel  3x3 data gradients of head (c2 -> nk) and stem (nk -> c1, scatter-add into G)
//   ln_bwd_stats_kernel  d(gamma), d(beta) (sum over the batch) + the two per-sample sums LayerNorm's
//                        backward needs;  ln_bwd_apply_kernel turns dL/d(LN out) into dL/d(pre-activation)
//   wgrad_pw_kernel      1x1 weight/bias gradients (register-tiled A^T * dY, LN-on-load)
//   wgrad_gconv_kernel   grouped dilated 3x3 weight/bias gradients
//   data gradients of the 1x1 and grouped convs reuse the forward kernels with transposed / flipped
//   weights (dgrad_pw, dgrad_gconv in stnet_kernels.cu).
// All weight gradients are ACCUMULATED (fp32 atomics) into `grads`, which has the layout of `params`.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>

#include "cnf_internal.h"
#include "device_utils.cuh"
#include "wgrad_tc.cuh"

namespace cnf {

#define CU_TRY(x)                          \
  do {                                     \
    cudaError_t e_ = (cudaError_t)(x);     \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

__device__ __forceinline__ float lrelu_slope(float x) { return x > 0.f ? 1.f : CNF_LRELU_SLOPE; }

// ------------------------------------------------------------------------------------------
// dL/d(zy) of cFlow.log_loss (M:1826-1848): L = -(mean_b(ll_z + ll_y) + logdet), so
//   x channels: z / B;   y channels: lambda_y * sign(zy - xy) / B.
// ------------------------------------------------------------------------------------------
__global__ void loss_grad_kernel(const float* __restrict__ zy, const float* __restrict__ xy, float* __restrict__ G,
                                 long long n, int D, int x_d, float lam, float invB) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % D);
    const float z = zy[i];
    float g;
    if (c < x_d) {
      g = z * invB;
    } else {
      const float d = z - xy[i];
      g = lam * invB * (d > 0.f ? 1.f : d < 0.f ? -1.f : 0.f);
    }
    G[i] = g;
  }
}

int launch_loss_grad(const float* zy, const float* xy, float* G, int64_t n, int D, int x_d, float lambda_y,
                     float inv_batch, void* stream) {
  if (n <= 0) return 0;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  loss_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(zy, xy, G, n, D, x_d, lambda_y, inv_batch);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Coupling law backward (M:1307-1326): v2 = exp(A) u2 + t, logdet += sum A, A = w tanh(raw_A).
// ------------------------------------------------------------------------------------------
struct HeadBwdArgs {
  FlowView g, s;       // gradient buffer and saved layer input, same geometry
  int mask_c, B, h, w, c2;
  const float* TH;     // tanh(raw_A) [B][hw][c2]
  const float* tanh_w; // scalar w of net A
  float* dtanh_w;
  float* DR;           // [2][B][hw][c2]: d raw_A, d raw_b
  long long dr_net_stride;
  float invB;
};

__global__ void __launch_bounds__(256) head_bwd_kernel(const HeadBwdArgs a) {
  __shared__ float red[64];
  const long long total = (long long)a.B * a.h * a.w * a.c2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float dw = 0.f;
  if (idx < total) {
    const int co = (int)(idx % a.c2);
    const long long pix = idx / a.c2;
    const int x = (int)(pix % a.w), y = (int)((pix / a.w) % a.h), b = (int)(pix / ((long long)a.w * a.h));
    const long long off = comp_off(a.g, a.mask_c, b, y, x, co);
    const float dv2 = a.g.base[off];
    const float u2 = a.s.base[off];
    const float th = a.TH[idx];
    const float w = *a.tanh_w;
    const float eA = expf(w * th);
    const float dA = dv2 * eA * u2 - a.invB;          // d(-mean_b logdet)/dA = -1/B
    a.g.base[off] = dv2 * eA;                         // dL/du2
    a.DR[idx] = dA * w * (1.f - th * th);             // d raw_A
    a.DR[a.dr_net_stride + idx] = dv2;                // d raw_b = dt
    dw = dA * th;
  }
  double d1, d2;
  block_sum2(dw, 0.f, red, d1, d2);
  if (threadIdx.x == 0 && d1 != 0.0) atomicAdd(a.dtanh_w, (float)d1);
}

// ------------------------------------------------------------------------------------------
// Transposed 3x3 conv with a small channel count on one side (head: c2 -> nk, stem: nk -> c1).
//   da[q, ci] = sum_tap sum_co dy[q - off(tap), co] W[tap][ci][co]
// mode 0: dense output [2][B][hw][CI] per net.  mode 1: both nets summed and ADDED into the flow
// gradient buffer at the positions of mask(., m, compress=True) (the u1 half read by the stem).
// ------------------------------------------------------------------------------------------
struct Conv3tArgs {
  const float* in;
  long long in_net_stride;
  const float* params;
  long long net_stride, w_off;
  int B, h, w, CI, CO, ks, mode;
  float* out;
  long long out_net_stride;
  FlowView view;
  int mask;
};

__global__ void __launch_bounds__(256) conv3t_small_kernel(const Conv3tArgs a) {
  const long long per_net = (long long)a.B * a.h * a.w * a.CI;
  const long long total = a.mode == 0 ? 2 * per_net : per_net;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int net0 = a.mode == 0 ? (int)(idx / per_net) : 0;
  const long long r = idx % per_net;
  const int ci = (int)(r % a.CI);
  const long long pix = r / a.CI;
  const int x = (int)(pix % a.w), y = (int)((pix / a.w) % a.h), b = (int)(pix / ((long long)a.w * a.h));
  const int pad = (a.ks - 1) / 2;
  const bool vec = (a.CO & 3) == 0;
  float acc = 0.f;
  const int n_nets = a.mode == 0 ? 1 : 2;
  for (int nn = 0; nn < n_nets; ++nn) {
    const int net = net0 + nn;
    const float* W = a.params + (long long)net * a.net_stride + a.w_off;
    const float* src = a.in + (long long)net * a.in_net_stride + (long long)b * a.h * a.w * a.CO;
    for (int ky = 0; ky < a.ks; ++ky) {
      const int py = y - (ky - pad);
      if (py < 0 || py >= a.h) continue;
      for (int kx = 0; kx < a.ks; ++kx) {
        const int px = x - (kx - pad);
        if (px < 0 || px >= a.w) continue;
        const float* d = src + ((long long)py * a.w + px) * a.CO;
        const float* wt = W + ((long long)(ky * a.ks + kx) * a.CI + ci) * a.CO;
        if (vec) {
          for (int co = 0; co < a.CO; co += 4) {
            const float4 dv = ld4(d + co), wv = ld4(wt + co);
            acc = fmaf(dv.x, wv.x, acc); acc = fmaf(dv.y, wv.y, acc);
            acc = fmaf(dv.z, wv.z, acc); acc = fmaf(dv.w, wv.w, acc);
          }
        } else {
          for (int co = 0; co < a.CO; ++co) acc = fmaf(d[co], wt[co], acc);
        }
      }
    }
  }
  if (a.mode == 0) {
    a.out[(long long)net0 * a.out_net_stride + r] = acc;
  } else {
    float* ptr = a.view.base + comp_off(a.view, a.mask, b, y, x, ci);
    *ptr += acc;   // every (b, y, x, ci) is owned by exactly one thread
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward.  a = LN(l) = (l - mean) rstd gamma + beta with l = lrelu(x) over one whole
// sample (F:350-360).  With g = dL/da, xh = (l - mean) rstd, n = elements per sample:
//   d gamma[e] = sum_b g xh,  d beta[e] = sum_b g,
//   dL/dl = rstd (g gamma - S1/n - xh S2/n),  S1 = sum_e g gamma,  S2 = sum_e g gamma xh,
//   dL/dx = lrelu'(x) dL/dl.
// ------------------------------------------------------------------------------------------
struct LnBwdArgs {
  const float* dy;
  const float* x;
  float* dx;
  const float* params;
  float* grads;
  long long net_stride, g_off, be_off;
  const double* stats;  // forward sums [2][B][2]
  double* bst;          // backward sums [2][B][2]
  int B, ln, accumulate, SB;
  long long n;
};

constexpr int LN_SB_MAX = 64;

template <int V>
__global__ void __launch_bounds__(256) ln_bwd_stats_kernel(const LnBwdArgs a) {
  __shared__ float sacc[LN_SB_MAX][2];
  __shared__ float mr[LN_SB_MAX][2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int net = blockIdx.z, b0 = blockIdx.y * a.SB;
  const int ns = min(a.SB, a.B - b0);
  const long long e0 = ((long long)blockIdx.x * 256 + tid) * V;
  const bool active = e0 < a.n;
  if (tid < LN_SB_MAX) {
    sacc[tid][0] = 0.f;
    sacc[tid][1] = 0.f;
    float m_ = 0.f, r_ = 1.f;
    if (tid < ns) ln_coeffs(a.stats, (long long)net * a.B + b0 + tid, (double)a.n, m_, r_);
    mr[tid][0] = m_;
    mr[tid][1] = r_;
  }
  __syncthreads();
  float gam[V], dg[V], db[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { gam[j] = 0.f; dg[j] = 0.f; db[j] = 0.f; }
  if (active) {
    const float* gp = a.params + (long long)net * a.net_stride + a.g_off + e0;
#pragma unroll
    for (int j = 0; j < V; ++j) gam[j] = gp[j];
  }
  const float* dyp = a.dy + ((long long)net * a.B + b0) * a.n + e0;
  const float* xp = a.x + ((long long)net * a.B + b0) * a.n + e0;
  for (int bi = 0; bi < ns; ++bi) {
    float s1 = 0.f, s2 = 0.f;
    if (active) {
      float g[V], xv[V];
      if (V == 4) {
        const float4 t0 = ld4(dyp + (long long)bi * a.n), t1 = ld4(xp + (long long)bi * a.n);
        g[0] = t0.x; g[V > 1 ? 1 : 0] = t0.y; g[V > 2 ? 2 : 0] = t0.z; g[V > 3 ? 3 : 0] = t0.w;
        xv[0] = t1.x; xv[V > 1 ? 1 : 0] = t1.y; xv[V > 2 ? 2 : 0] = t1.z; xv[V > 3 ? 3 : 0] = t1.w;
      } else {
        g[0] = dyp[(long long)bi * a.n];
        xv[0] = xp[(long long)bi * a.n];
      }
      const float mean = mr[bi][0], rstd = mr[bi][1];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float xh = (lrelu(xv[j]) - mean) * rstd;
        dg[j] = fmaf(g[j], xh, dg[j]);
        db[j] += g[j];
        const float t = g[j] * gam[j];
        s1 += t;
        s2 = fmaf(t, xh, s2);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      atomicAdd(&sacc[bi][0], s1);
      atomicAdd(&sacc[bi][1], s2);
    }
  }
  if (active) {
    float* gg = a.grads + (long long)net * a.net_stride + a.g_off + e0;
    float* gb = a.grads + (long long)net * a.net_stride + a.be_off + e0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      atomicAdd(gg + j, dg[j]);
      atomicAdd(gb + j, db[j]);
    }
  }
  __syncthreads();
  if (tid < 2 * ns) atomicAdd(a.bst + 2 * ((long long)net * a.B + b0 + (tid >> 1)) + (tid & 1), (double)sacc[tid >> 1][tid & 1]);
}

template <int V>
__global__ void __launch_bounds__(256) ln_bwd_apply_kernel(const LnBwdArgs a) {
  const int net = blockIdx.z, b = blockIdx.y;
  const long long e0 = ((long long)blockIdx.x * 256 + threadIdx.x) * V;
  if (e0 >= a.n) return;
  const long long sidx = (long long)net * a.B + b;
  const long long base = sidx * a.n + e0;
  float g[V], xv[V], o[V];
  if (V == 4) {
    const float4 t0 = ld4(a.dy + base), t1 = ld4(a.x + base);
    g[0] = t0.x; g[V > 1 ? 1 : 0] = t0.y; g[V > 2 ? 2 : 0] = t0.z; g[V > 3 ? 3 : 0] = t0.w;
    xv[0] = t1.x; xv[V > 1 ? 1 : 0] = t1.y; xv[V > 2 ? 2 : 0] = t1.z; xv[V > 3 ? 3 : 0] = t1.w;
  } else {
    g[0] = a.dy[base];
    xv[0] = a.x[base];
  }
  if (a.ln) {
    float mean, rstd;
    ln_coeffs(a.stats, sidx, (double)a.n, mean, rstd);
    const float inv_n = 1.0f / (float)a.n;
    const float c1 = (float)a.bst[2 * sidx] * inv_n, c2 = (float)a.bst[2 * sidx + 1] * inv_n;
    const float* gp = a.params + (long long)net * a.net_stride + a.g_off + e0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xh = (lrelu(xv[j]) - mean) * rstd;
      o[j] = lrelu_slope(xv[j]) * rstd * (g[j] * gp[j] - c1 - xh * c2);
    }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = lrelu_slope(xv[j]) * g[j];
  }
  if (a.accumulate) {
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] += a.dx[base + j];
  }
  if (V == 4) {
    st4(a.dx + base, make_float4(o[0], o[V > 1 ? 1 : 0], o[V > 2 ? 2 : 0], o[V > 3 ? 3 : 0]));
  } else {
    a.dx[base] = o[0];
  }
}

// dy -> dx through LN(lrelu(.)); dx may alias dy.  bst: scratch of 2*B*2 doubles.
static int ln_backward(const float* dy, const float* x, float* dx, const float* params, float* grads,
                       long long net_stride, long long g_off, long long be_off, const double* stats, double* bst, int B,
                       long long n, int ln, int accumulate, cudaStream_t st) {
  LnBwdArgs a = {};
  a.dy = dy; a.x = x; a.dx = dx; a.params = params; a.grads = grads; a.net_stride = net_stride;
  a.g_off = g_off; a.be_off = be_off; a.stats = stats; a.bst = bst; a.B = B; a.ln = ln; a.accumulate = accumulate; a.n = n;
  const bool v4 = (n % 4) == 0;
  const int per = v4 ? 1024 : 256;
  const int tiles = (int)((n + per - 1) / per);
  if (B > 65535) return (int)cudaErrorInvalidConfiguration;
  if (ln) {
    CU_TRY(cudaMemsetAsync(bst, 0, sizeof(double) * 4 * B, st));
    // enough CTAs for ~4 waves, at most LN_SB_MAX samples per CTA
    int chunks = std::max(1, std::min(B, (148 * 8 + 2 * tiles - 1) / (2 * tiles)));
    a.SB = std::min(LN_SB_MAX, (B + chunks - 1) / chunks);
    chunks = (B + a.SB - 1) / a.SB;
    dim3 grid(tiles, chunks, 2);
    if (v4) ln_bwd_stats_kernel<4><<<grid, 256, 0, st>>>(a);
    else ln_bwd_stats_kernel<1><<<grid, 256, 0, st>>>(a);
    CU_TRY(cudaGetLastError());
  }
  dim3 grid(tiles, B, 2);
  if (v4) ln_bwd_apply_kernel<4><<<grid, 256, 0, st>>>(a);
  else ln_bwd_apply_kernel<1><<<grid, 256, 0, st>>>(a);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// 1x1 conv weight gradient: dW[k][n] += sum_{b,p} a[b,p,k] dy[b,p,n],  db[n] += sum dy,
// a = LN(lrelu(x)) applied while staging.  CTA = (pixel chunk, (k tile, n tile), net); threads form a
// rows x cols grid of RK x 4 register tiles; spare threads form extra teams that split the staged pixels.
// ------------------------------------------------------------------------------------------
struct WgradPwArgs {
  const float* x;
  long long x_net_stride;
  const float* dy;
  long long dy_net_stride;
  const float* params;
  float* grads;
  long long net_stride, w_off, b_off, g_off, be_off;
  const double* stats;
  int B, hw, K, N, ln;
  int KT, NTl, k_tiles, n_tiles, rows, cols, teams;
  int tiles_per_sample, n_items, items_per_cta;
};

constexpr int WG_PT = 32;  // pixels per staged tile

template <int RK>
__global__ void __launch_bounds__(256) wgrad_pw_kernel(const WgradPwArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, NT = blockDim.x;
  const int net = blockIdx.z;
  const int kt = blockIdx.y / a.n_tiles, nt = blockIdx.y % a.n_tiles;
  const int k0 = kt * a.KT, n0 = nt * a.NTl;
  const int kt_len = min(a.KT, a.K - k0), nt_len = min(a.NTl, a.N - n0);
  const int KS = a.rows * RK + 4;          // padded row of As (covers every thread's RK slice)
  const int NS = a.cols * 4;
  float* As = smem;                        // [WG_PT][KS]
  float* Ds = smem + WG_PT * KS;           // [WG_PT][NS]
  const int tpt = a.rows * a.cols;
  const int team = tid / tpt, t = tid % tpt;
  const bool worker = team < a.teams;
  const int tk = t / a.cols, tn = t % a.cols;
  const float* P = a.params + (long long)net * a.net_stride;
  const float* gam = P + a.g_off;
  const float* bet = P + a.be_off;
  const bool kvec = (a.K % 4) == 0 && (k0 % 4) == 0, nvec = (a.N % 4) == 0;

  float acc[RK][4];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < RK; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int it0 = blockIdx.x * a.items_per_cta, it1 = min(a.n_items, it0 + a.items_per_cta);
  int cur_b = -1;
  float mean = 0.f, rstd = 1.f;
  for (int it = it0; it < it1; ++it) {
    const int b = it / a.tiles_per_sample, p0 = (it % a.tiles_per_sample) * WG_PT;
    const int np = min(WG_PT, a.hw - p0);
    if (b != cur_b) {
      cur_b = b;
      if (a.ln) ln_coeffs(a.stats, (long long)net * a.B + b, (double)a.hw * (double)a.K, mean, rstd);
    }
    const float* xs = a.x + (long long)net * a.x_net_stride + (long long)b * a.hw * a.K;
    const float* ds = a.dy + (long long)net * a.dy_net_stride + (long long)b * a.hw * a.N;
    __syncthreads();
    // ---- stage a = LN(lrelu(x)) for columns [k0, k0 + KS) (zero beyond kt_len)
    if (kvec) {
      const int kq_n = KS / 4;
      for (int idx = tid; idx < WG_PT * kq_n; idx += NT) {
        const int p = idx / kq_n, k = (idx % kq_n) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < np && k < kt_len) {   // kt_len % 4 == 0 here
          const long long e = (long long)(p0 + p) * a.K + k0 + k;
          v = ld4(xs + e);
          v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w);
          if (a.ln) {
            const float4 g = ld4(gam + e), be = ld4(bet + e);
            v.x = (v.x - mean) * rstd * g.x + be.x;
            v.y = (v.y - mean) * rstd * g.y + be.y;
            v.z = (v.z - mean) * rstd * g.z + be.z;
            v.w = (v.w - mean) * rstd * g.w + be.w;
          }
        }
        st4(&As[p * KS + k], v);
      }
    } else {
      for (int idx = tid; idx < WG_PT * KS; idx += NT) {
        const int p = idx / KS, k = idx % KS;
        float v = 0.f;
        if (p < np && k < kt_len) {
          const long long e = (long long)(p0 + p) * a.K + k0 + k;
          v = lrelu(xs[e]);
          if (a.ln) v = (v - mean) * rstd * gam[e] + bet[e];
        }
        As[p * KS + k] = v;
      }
    }
    // ---- stage dy columns [n0, n0 + NS)
    if (nvec) {
      const int nq_n = NS / 4;
      for (int idx = tid; idx < WG_PT * nq_n; idx += NT) {
        const int p = idx / nq_n, n = (idx % nq_n) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < np && n < nt_len) v = ld4(ds + (long long)(p0 + p) * a.N + n0 + n);
        st4(&Ds[p * NS + n], v);
      }
    } else {
      for (int idx = tid; idx < WG_PT * NS; idx += NT) {
        const int p = idx / NS, n = idx % NS;
        float v = 0.f;
        if (p < np && n < nt_len) v = ds[(long long)(p0 + p) * a.N + n0 + n];
        Ds[p * NS + n] = v;
      }
    }
    __syncthreads();
    if (worker) {
      for (int p = team; p < np; p += a.teams) {
        const float4 d = ld4(&Ds[p * NS + tn * 4]);
        float av[RK];
#pragma unroll
        for (int i = 0; i < RK; i += 4) {
          const float4 t4 = ld4(&As[p * KS + tk * RK + i]);
          av[i] = t4.x; av[i + 1] = t4.y; av[i + 2] = t4.z; av[i + 3] = t4.w;
        }
#pragma unroll
        for (int i = 0; i < RK; ++i) {
          acc[i][0] = fmaf(av[i], d.x, acc[i][0]);
          acc[i][1] = fmaf(av[i], d.y, acc[i][1]);
          acc[i][2] = fmaf(av[i], d.z, acc[i][2]);
          acc[i][3] = fmaf(av[i], d.w, acc[i][3]);
        }
        if (tk == 0) { bacc[0] += d.x; bacc[1] += d.y; bacc[2] += d.z; bacc[3] += d.w; }
      }
    }
  }
  float* gW = a.grads + (long long)net * a.net_stride + a.w_off;
  float* gB = a.grads + (long long)net * a.net_stride + a.b_off;
  if (a.teams > 1) {
    // several teams hold partial sums of the same outputs: combine them in shared memory first
    __syncthreads();
    const int KP = a.rows * RK;
    float* red_w = smem;               // [KP][NS]
    float* red_b = smem + KP * NS;     // [NS]
    for (int i = tid; i < KP * NS + NS; i += NT) smem[i] = 0.f;
    __syncthreads();
    if (worker) {
#pragma unroll
      for (int i = 0; i < RK; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(red_w + (tk * RK + i) * NS + tn * 4 + j, acc[i][j]);
      if (tk == 0)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(red_b + tn * 4 + j, bacc[j]);
    }
    __syncthreads();
    for (int i = tid; i < KP * NS; i += NT) {
      const int k = i / NS, n = i % NS;
      if (k < kt_len && n < nt_len) atomicAdd(gW + (long long)(k0 + k) * a.N + n0 + n, red_w[i]);
    }
    if (kt == 0)
      for (int n = tid; n < nt_len; n += NT) atomicAdd(gB + n0 + n, red_b[n]);
    return;
  }
  if (!worker) return;
#pragma unroll
  for (int i = 0; i < RK; ++i) {
    const int k = tk * RK + i;
    if (k >= kt_len) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tn * 4 + j;
      if (n < nt_len) atomicAdd(gW + (long long)(k0 + k) * a.N + n0 + n, acc[i][j]);
    }
  }
  if (tk == 0 && kt == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tn * 4 + j;
      if (n < nt_len) atomicAdd(gB + n0 + n, bacc[j]);
    }
  }
}

// tensor-core version (wgrad_tc.cuh): K <= 128 input channels, N <= 64 output channels; CNF_NOT_ELIGIBLE otherwise
static int launch_wgrad_pw_tc(const WgradPwArgs& w, float* partial, cudaStream_t st) {
  if (w.K % 4 || w.K > 128 || w.K < 4 || w.N % 4 || w.N > 64 || w.N < 4 || !partial) return CNF_NOT_ELIGIBLE;
  if ((((uintptr_t)w.x) & 15) || (((uintptr_t)w.dy) & 15) || ((w.x_net_stride * 4) % 16) || ((w.dy_net_stride * 4) % 16)) return CNF_NOT_ELIGIBLE;
  WgtArgs a{};
  a.x = w.x; a.dy = w.dy; a.x_net_stride = w.x_net_stride; a.dy_net_stride = w.dy_net_stride;
  a.params = w.params; a.net_stride = w.net_stride; a.g_off = w.g_off; a.be_off = w.be_off; a.stats = w.stats;
  a.partial = partial;
  a.B = w.B; a.hw = w.hw; a.K = w.K; a.N = w.N; a.ln = w.ln;
  a.sps = (w.hw + WGT_PX - 1) / WGT_PX;
  a.n_stages = w.B * a.sps;
  int n_sm = 0;
  CU_TRY((cudaError_t)device_sm_count(&n_sm));
  a.nc = std::max(1, std::min(std::min(a.n_stages, n_sm), WGT_MAX_NC));   // two CTAs per SM, half of them per net
  const int NB = w.N <= 16 ? 16 : w.N <= 32 ? 32 : 64;
  const size_t smem = (size_t)2 * (2 * 128 * WGT_PX + 2 * NB * WGT_PX) * sizeof(float);
  dim3 grid(a.nc, 2);
  if (NB == 16) {
    static SmemAttrCache cache;
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)wgrad_tc_kernel<16>, smem, cache));
    wgrad_tc_kernel<16><<<grid, WGT_NT, smem, st>>>(a);
  } else if (NB == 32) {
    static SmemAttrCache cache;
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)wgrad_tc_kernel<32>, smem, cache));
    wgrad_tc_kernel<32><<<grid, WGT_NT, smem, st>>>(a);
  } else {
    static SmemAttrCache cache;
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)wgrad_tc_kernel<64>, smem, cache));
    wgrad_tc_kernel<64><<<grid, WGT_NT, smem, st>>>(a);
  }
  CU_TRY(cudaGetLastError());
  const int total = w.K * w.N + w.N;
  wgrad_tc_reduce_kernel<<<dim3((total + 255) / 256, 2), 256, 0, st>>>(partial, a.nc, w.K * w.N, w.N, w.grads, w.net_stride, w.w_off, w.b_off);
  return (int)cudaGetLastError();
}

static int launch_wgrad_pw(WgradPwArgs a, cudaStream_t st) {
  a.KT = std::min(a.K, 128);
  a.NTl = std::min((a.N + 3) & ~3, 64);
  a.k_tiles = (a.K + a.KT - 1) / a.KT;
  a.n_tiles = (a.N + a.NTl - 1) / a.NTl;
  const int RK = a.KT > 64 ? 8 : 4;
  a.rows = (a.KT + RK - 1) / RK;
  a.cols = (a.NTl + 3) / 4;
  const int tpt = a.rows * a.cols;       // <= 16 * 16
  a.teams = std::max(1, std::min(WG_PT, 256 / tpt));
  const int NT = ((tpt * a.teams + 31) / 32) * 32;
  a.tiles_per_sample = (a.hw + WG_PT - 1) / WG_PT;
  a.n_items = a.B * a.tiles_per_sample;
  const int want_ctas = std::max(1, 148 * 4 / (2 * a.k_tiles * a.n_tiles));
  a.items_per_cta = std::max(4, (a.n_items + want_ctas - 1) / want_ctas);
  const int chunks = (a.n_items + a.items_per_cta - 1) / a.items_per_cta;
  size_t smem = (size_t)WG_PT * ((a.rows * RK + 4) + a.cols * 4) * sizeof(float);
  if (a.teams > 1) smem = std::max(smem, (size_t)(a.rows * RK + 1) * a.cols * 4 * sizeof(float));
  dim3 grid(chunks, a.k_tiles * a.n_tiles, 2);
  if (RK == 8) wgrad_pw_kernel<8><<<grid, NT, smem, st>>>(a);
  else wgrad_pw_kernel<4><<<grid, NT, smem, st>>>(a);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// 3x3 weight gradient with a small channel count on one side (stem: c1 -> nk, head: nk -> c2):
//   dW[tap][ci][co] += sum_{b,q} a[b, q + off(tap), ci] dy[b, q, co],  db[co] += sum dy.
// mode 0: a = LN(lrelu(x)) of a dense x [2][B][hw][CI];  mode 1: a = masked gather from the saved flow
// state (the stem's input, identical for both nets).  One thread per output (up to WG3_ACC each).
// ------------------------------------------------------------------------------------------
struct Wgrad3Args {
  const float* x;
  long long x_net_stride;
  FlowView view;
  int mask, mode;
  const float* dy;
  long long dy_net_stride;
  const float* params;
  float* grads;
  long long net_stride, w_off, b_off, g_off, be_off;
  const double* stats;
  int B, h, w, CI, CO, ks, ln;
  int TH, TW, tiles_y, tiles_x, SB, CIC, ci_chunks;
};

constexpr int WG3_ACC = 12;

__global__ void __launch_bounds__(256) wgrad3_small_kernel(const Wgrad3Args a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, NT = blockDim.x;
  const int net = blockIdx.z;
  const int chunk = blockIdx.y;
  const int c0 = chunk * a.CIC, cic = min(a.CIC, a.CI - c0);
  const int tiles = a.tiles_y * a.tiles_x;
  const int tile = blockIdx.x % tiles, sb = blockIdx.x / tiles;
  const int y0 = (tile / a.tiles_x) * a.TH, x0 = (tile % a.tiles_x) * a.TW;
  const int th = min(a.TH, a.h - y0), tw = min(a.TW, a.w - x0);
  const int pad = (a.ks - 1) / 2;
  const int SH = a.TH + 2 * pad, SW = a.TW + 2 * pad;
  float* a_s = smem;                               // [SH][SW][CIC]
  float* d_s = smem + SH * SW * a.CIC;             // [TH][TW][CO]
  const float* P = a.params + (long long)net * a.net_stride;
  const float* gam = P + a.g_off;
  const float* bet = P + a.be_off;
  const int n_out = a.ks * a.ks * cic * a.CO;

  float acc[WG3_ACC];
  int aoff[WG3_ACC], doff[WG3_ACC];
#pragma unroll
  for (int j = 0; j < WG3_ACC; ++j) {
    acc[j] = 0.f;
    const int o = tid + j * NT;
    if (o < n_out) {
      const int co = o % a.CO, ci = (o / a.CO) % cic, tap = o / (a.CO * cic);
      aoff[j] = ((tap / a.ks) * SW + tap % a.ks) * a.CIC + ci;
      doff[j] = co;
    } else {
      aoff[j] = -1;
      doff[j] = 0;
    }
  }
  const int b0 = sb * a.SB, b1 = min(a.B, b0 + a.SB);
  for (int b = b0; b < b1; ++b) {
    float mean = 0.f, rstd = 1.f;
    if (a.mode == 0 && a.ln) ln_coeffs(a.stats, (long long)net * a.B + b, (double)a.h * a.w * (double)a.CI, mean, rstd);
    __syncthreads();
    for (int idx = tid; idx < SH * SW * a.CIC; idx += NT) {
      const int ci = idx % a.CIC, pix = idx / a.CIC;
      const int gy = y0 - pad + pix / SW, gx = x0 - pad + pix % SW;
      float v = 0.f;
      if (ci < cic && gy >= 0 && gy < a.h && gx >= 0 && gx < a.w) {
        if (a.mode == 0) {
          const long long e = ((long long)gy * a.w + gx) * a.CI + c0 + ci;
          v = lrelu(a.x[(long long)net * a.x_net_stride + (long long)b * a.h * a.w * a.CI + e]);
          if (a.ln) v = (v - mean) * rstd * gam[e] + bet[e];
        } else {
          v = a.view.base[comp_off(a.view, a.mask, b, gy, gx, c0 + ci)];
        }
      }
      a_s[idx] = v;
    }
    const float* ds = a.dy + (long long)net * a.dy_net_stride + (long long)b * a.h * a.w * a.CO;
    for (int idx = tid; idx < a.TH * a.TW * a.CO; idx += NT) {
      const int co = idx % a.CO, pix = idx / a.CO;
      const int py = pix / a.TW, px = pix % a.TW;
      float v = 0.f;
      if (py < th && px < tw) v = ds[((long long)(y0 + py) * a.w + x0 + px) * a.CO + co];
      d_s[idx] = v;
    }
    __syncthreads();
    for (int py = 0; py < th; ++py)
      for (int px = 0; px < tw; ++px) {
        const float* ap = a_s + (py * SW + px) * a.CIC;
        const float* dp = d_s + (py * a.TW + px) * a.CO;
#pragma unroll
        for (int j = 0; j < WG3_ACC; ++j)
          if (aoff[j] >= 0) acc[j] = fmaf(ap[aoff[j]], dp[doff[j]], acc[j]);
      }
    if (chunk == 0) {
      float* gB = a.grads + (long long)net * a.net_stride + a.b_off;
      for (int co = tid; co < a.CO; co += NT) {
        float s = 0.f;
        for (int p = 0; p < a.TH * a.TW; ++p) s += d_s[p * a.CO + co];
        atomicAdd(gB + co, s);
      }
    }
  }
  float* gW = a.grads + (long long)net * a.net_stride + a.w_off;
#pragma unroll
  for (int j = 0; j < WG3_ACC; ++j) {
    const int o = tid + j * NT;
    if (o < n_out) {
      const int co = o % a.CO, ci = (o / a.CO) % cic, tap = o / (a.CO * cic);
      atomicAdd(gW + ((long long)tap * a.CI + c0 + ci) * a.CO + co, acc[j]);
    }
  }
}

static int launch_wgrad3(Wgrad3Args a, cudaStream_t st) {
  const int taps = a.ks * a.ks;
  if (taps * a.CO > 256 * WG3_ACC) return (int)cudaErrorInvalidConfiguration;
  a.CIC = std::max(1, std::min(a.CI, (256 * WG3_ACC) / (taps * a.CO)));
  a.CIC = std::min(a.CIC, 32);
  a.ci_chunks = (a.CI + a.CIC - 1) / a.CIC;
  a.TW = std::min(a.w, 32);
  const int pad = (a.ks - 1) / 2;
  a.TH = std::min(a.h, 16);
  auto bytes = [&](int th) { return (size_t)((th + 2 * pad) * (a.TW + 2 * pad) * a.CIC + th * a.TW * a.CO) * sizeof(float); };
  while (a.TH > 1 && bytes(a.TH) > 64 * 1024) --a.TH;
  if (bytes(a.TH) > 200 * 1024) return (int)cudaErrorInvalidConfiguration;
  a.tiles_y = (a.h + a.TH - 1) / a.TH;
  a.tiles_x = (a.w + a.TW - 1) / a.TW;
  const int tiles = a.tiles_y * a.tiles_x;
  const int want = std::max(1, 148 * 6 / (2 * a.ci_chunks * tiles));
  a.SB = std::max(1, (a.B + want - 1) / want);
  const int sbs = (a.B + a.SB - 1) / a.SB;
  const size_t smem = bytes(a.TH);
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)wgrad3_small_kernel, smem, cache));
  dim3 grid(tiles * sbs, a.ci_chunks, 2);
  wgrad3_small_kernel<<<grid, 256, smem, st>>>(a);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Warp-per-pixel kernels for the 3x3 convs that have a WIDE side (nk channels, spread over the lanes:
// coalesced row loads/stores) and a NARROW side (c1 or c2 <= 24 channels, broadcast scalars):
//   wgrad3_wide_kernel   weight + bias gradients of head (wide = LN(lrelu(x_R)) is cin, narrow = d raw is
//                        cout, neighbour at p - off) and stem (wide = dX0 is cout, narrow = u1c gathered
//                        through the mask from the saved flow state is cin, neighbour at p + off)
//   head_dgrad_kernel    d(LN_f output)[p, ci] = sum_tap sum_co draw[p - off, co] W[tap][ci][co]
//   stem_dgrad_kernel    G[u1 positions] += sum_net sum_tap sum_co dX0[p - off, co] W[tap][ci][co]
// Each activation row is read exactly once; the kernels are HBM-bound (4*nk bytes per pixel and net).
// ------------------------------------------------------------------------------------------
struct Wgrad3WideArgs {
  const float* wide;
  long long wide_net_stride;
  int CW, wide_act;             // wide_act: apply lrelu (+ LayerNorm) while loading (head)
  const float* narrow;          // dense [2][B][hw][CN] (head) ...
  long long narrow_net_stride;
  FlowView view;                // ... or gathered from the flow state through `mask` (stem; same for both nets)
  int mask, narrow_view, CN, sign;
  int wide_is_ci;               // head: dW[tap][wide][narrow];  stem: dW[tap][narrow][wide]
  const float* params;
  float* grads;
  long long net_stride, w_off, b_off, g_off, be_off;
  const double* stats;
  int B, h, w, ks, ln;
};

template <int WPL, int NC>
__global__ void __launch_bounds__(256) wgrad3_wide_kernel(const Wgrad3WideArgs a) {
  extern __shared__ __align__(16) float red_s[];   // [9][NC][CW]
  __shared__ float bias_s[256];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int net = blockIdx.z, n0 = blockIdx.y * NC;
  const int nc = min(NC, a.CN - n0);
  const int hw = a.h * a.w, pad = (a.ks - 1) / 2;
  const int ch0 = lane * WPL;
  const float* P = a.params + (long long)net * a.net_stride;
  const float* gam = P + a.g_off;
  const float* bet = P + a.be_off;
  for (int i = tid; i < 9 * NC * a.CW; i += 256) red_s[i] = 0.f;
  bias_s[tid] = 0.f;
  float acc[9][NC][WPL];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int j = 0; j < WPL; ++j) acc[t][c][j] = 0.f;
  float bw[WPL], bn[NC];
#pragma unroll
  for (int j = 0; j < WPL; ++j) bw[j] = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) bn[c] = 0.f;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    float mean = 0.f, rstd = 1.f;
    if (a.wide_act && a.ln) ln_coeffs(a.stats, (long long)net * a.B + b, (double)hw * (double)a.CW, mean, rstd);
    const float* wsrc = a.wide + (long long)net * a.wide_net_stride + (long long)b * hw * a.CW;
    const float* nsrc = a.narrow_view ? nullptr : a.narrow + (long long)net * a.narrow_net_stride + (long long)b * hw * a.CN;
    for (int p = wid; p < hw; p += 8) {
      const int y = p / a.w, x = p - y * a.w;
      float wv[WPL];
#pragma unroll
      for (int j = 0; j < WPL; ++j) {
        const int ch = ch0 + j;
        float v = 0.f;
        if (ch < a.CW) {
          const long long e = (long long)p * a.CW + ch;
          v = wsrc[e];
          if (a.wide_act) {
            v = lrelu(v);
            if (a.ln) v = (v - mean) * rstd * gam[e] + bet[e];
          }
        }
        wv[j] = v;
      }
      // all 9*NC neighbour scalars are fetched with predicated loads first (independent, in flight together)
      float nv[9][NC];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + a.sign * (ky - pad);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + a.sign * (kx - pad);
          const bool ok = ky < a.ks && kx < a.ks && yy >= 0 && yy < a.h && xx >= 0 && xx < a.w;   // warp-uniform
          const int yc = ok ? yy : y, xc = ok ? xx : x;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const int cc = c < nc ? n0 + c : n0;
            const float v = a.narrow_view ? a.view.base[comp_off(a.view, a.mask, b, yc, xc, cc)]
                                          : nsrc[(long long)(yc * a.w + xc) * a.CN + cc];
            nv[ky * 3 + kx][c] = (ok && c < nc) ? v : 0.f;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int j = 0; j < WPL; ++j) acc[t][c][j] = fmaf(nv[t][c], wv[j], acc[t][c][j]);
      if (pad * 3 + pad < 9) {
#pragma unroll
        for (int c = 0; c < NC; ++c) bn[c] += nv[4 * (pad != 0)][c];   // centre tap (index 4 for 3x3, 0 for 1x1)
      }
#pragma unroll
      for (int j = 0; j < WPL; ++j) bw[j] += wv[j];
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int j = 0; j < WPL; ++j)
        if (ch0 + j < a.CW) atomicAdd(&red_s[(t * NC + c) * a.CW + ch0 + j], acc[t][c][j]);
  // bias: head (wide_is_ci) -> sum of the narrow operand, stem -> sum of the wide operand (first narrow chunk only)
  if (a.wide_is_ci) {
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < NC; ++c) atomicAdd(&bias_s[c], bn[c]);
  } else if (n0 == 0) {
#pragma unroll
    for (int j = 0; j < WPL; ++j)
      if (ch0 + j < a.CW) atomicAdd(&bias_s[ch0 + j], bw[j]);
  }
  __syncthreads();
  float* gW = a.grads + (long long)net * a.net_stride + a.w_off;
  float* gB = a.grads + (long long)net * a.net_stride + a.b_off;
  const int taps = a.ks * a.ks;
  for (int i = tid; i < 9 * NC * a.CW; i += 256) {
    const int ch = i % a.CW, c = (i / a.CW) % NC, t = i / (a.CW * NC);
    const int ky = t / 3, kx = t % 3;
    if (c >= nc || ky >= a.ks || kx >= a.ks) continue;
    const int tap = ky * a.ks + kx;
    const long long o = a.wide_is_ci ? ((long long)tap * a.CW + ch) * a.CN + n0 + c : ((long long)tap * a.CN + n0 + c) * a.CW + ch;
    (void)taps;
    atomicAdd(gW + o, red_s[i]);
  }
  if (a.wide_is_ci) {
    if (tid < nc) atomicAdd(gB + n0 + tid, bias_s[tid]);
  } else if (n0 == 0) {
    if (tid < a.CW) atomicAdd(gB + tid, bias_s[tid]);
  }
}

template <int WPL, int NC>
static int launch_wgrad3_wide_t(const Wgrad3WideArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)9 * NC * a.CW * sizeof(float);
  const int chunks = (a.CN + NC - 1) / NC;
  const int gx = std::max(1, std::min(a.B, 148 * 4 / (2 * chunks) + 1));
  dim3 grid(gx, chunks, 2);
  wgrad3_wide_kernel<WPL, NC><<<grid, 256, smem, st>>>(a);
  return (int)cudaGetLastError();
}

// returns 1 when the shape is not covered (caller falls back to wgrad3_small_kernel)
static int launch_wgrad3_wide(const Wgrad3WideArgs& a, cudaStream_t st) {
  if (a.ks != 3 && a.ks != 1) return CNF_NOT_ELIGIBLE;
  if (a.CW > 256) return CNF_NOT_ELIGIBLE;
  if (a.CW <= 32) {
    if (a.CN >= 4) return launch_wgrad3_wide_t<1, 4>(a, st);
    if (a.CN >= 2) return launch_wgrad3_wide_t<1, 2>(a, st);
    return launch_wgrad3_wide_t<1, 1>(a, st);
  }
  if (a.CW <= 64) {
    if (a.CN >= 4) return launch_wgrad3_wide_t<2, 4>(a, st);
    if (a.CN >= 2) return launch_wgrad3_wide_t<2, 2>(a, st);
    return launch_wgrad3_wide_t<2, 1>(a, st);
  }
  if (a.CW <= 128) {
    if (a.CN >= 2) return launch_wgrad3_wide_t<4, 2>(a, st);
    return launch_wgrad3_wide_t<4, 1>(a, st);
  }
  return launch_wgrad3_wide_t<8, 1>(a, st);
}

struct Dgrad3Args {
  const float* in;      // head: d raw [2][B][hw][CN];  stem: dX0 [2][B][hw][CW]
  long long in_net_stride;
  const float* params;
  long long net_stride, w_off;
  int B, h, w, CW, CN, ks;
  float* out;           // head: [2][B][hw][CW]
  long long out_net_stride;
  FlowView view;        // stem: gradient buffer, positions of mask(., mask, compress=True)
  int mask;
};

// head: CW = nk outputs over the lanes, CN = c2 broadcast inputs.  smem: W as [tap][co][ci].
template <int WPL>
__global__ void __launch_bounds__(256) head_dgrad_kernel(const Dgrad3Args a) {
  extern __shared__ __align__(16) float w_s[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int net = blockIdx.y;
  const int hw = a.h * a.w, pad = (a.ks - 1) / 2, taps = a.ks * a.ks;
  const float* W = a.params + (long long)net * a.net_stride + a.w_off;   // [tap][ci = CW][co = CN]
  for (int i = tid; i < taps * a.CN * a.CW; i += 256) {
    const int ci = i % a.CW, co = (i / a.CW) % a.CN, tap = i / (a.CW * a.CN);
    w_s[i] = W[((long long)tap * a.CW + ci) * a.CN + co];
  }
  __syncthreads();
  const int ch0 = lane * WPL;
  const long long total = (long long)a.B * hw;
  for (long long it = (long long)blockIdx.x * 8 + wid; it < total; it += (long long)gridDim.x * 8) {
    const int b = (int)(it / hw), p = (int)(it - (long long)b * hw);
    const int y = p / a.w, x = p - y * a.w;
    const float* src = a.in + (long long)net * a.in_net_stride + (long long)b * hw * a.CN;
    float o[WPL];
#pragma unroll
    for (int j = 0; j < WPL; ++j) o[j] = 0.f;
    for (int co = 0; co < a.CN; ++co) {
      float dv[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y - (ky - pad);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x - (kx - pad);
          const bool ok = ky < a.ks && kx < a.ks && yy >= 0 && yy < a.h && xx >= 0 && xx < a.w;
          const float v = src[(long long)((ok ? yy : y) * a.w + (ok ? xx : x)) * a.CN + co];
          dv[ky * 3 + kx] = ok ? v : 0.f;
        }
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if (ky < a.ks && kx < a.ks) {
            const float* wt = w_s + ((ky * a.ks + kx) * a.CN + co) * a.CW;
#pragma unroll
            for (int j = 0; j < WPL; ++j)
              if (ch0 + j < a.CW) o[j] = fmaf(dv[ky * 3 + kx], wt[ch0 + j], o[j]);
          }
        }
    }
    float* dst = a.out + (long long)net * a.out_net_stride + it * a.CW;
#pragma unroll
    for (int j = 0; j < WPL; ++j)
      if (ch0 + j < a.CW) dst[ch0 + j] = o[j];
  }
}

// stem: CW = nk inputs over the lanes, CN = c1 outputs.  Scatter form: every dX0 row is read ONCE; its 9*c1 dot
// products with W[tap][ci][:] are reduced across the warp and added (atomics: neighbouring rows hit the same
// pixels) into the u1 positions of the gradient buffer at q + off(tap).  smem: W [tap][ci][co] of this net.
template <int WPL>
__global__ void __launch_bounds__(256) stem_dgrad_kernel(const Dgrad3Args a) {
  extern __shared__ __align__(16) float w_s[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int net = blockIdx.y;
  const int hw = a.h * a.w, pad = (a.ks - 1) / 2, taps = a.ks * a.ks;
  const int per_net = taps * a.CN * a.CW;
  for (int i = tid; i < per_net; i += 256) w_s[i] = a.params[(long long)net * a.net_stride + a.w_off + i];
  __syncthreads();
  const int ch0 = lane * WPL;
  const long long total = (long long)a.B * hw;
  const float* src = a.in + (long long)net * a.in_net_stride;
  for (long long it = (long long)blockIdx.x * 8 + wid; it < total; it += (long long)gridDim.x * 8) {
    const int b = (int)(it / hw), p = (int)(it - (long long)b * hw);
    const int y = p / a.w, x = p - y * a.w;
    float dv[WPL];
#pragma unroll
    for (int j = 0; j < WPL; ++j) dv[j] = ch0 + j < a.CW ? src[it * a.CW + ch0 + j] : 0.f;
    for (int ci = 0; ci < a.CN; ++ci) {
      float part[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float sacc = 0.f;
        if (t < taps) {
          const float* wt = w_s + (t * a.CN + ci) * a.CW;
#pragma unroll
          for (int j = 0; j < WPL; ++j)
            if (ch0 + j < a.CW) sacc = fmaf(dv[j], wt[ch0 + j], sacc);
        }
        part[t] = sacc;
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) part[t] = warp_sum(part[t]);
      // dX0[q] * W[tap] contributes to du1 at q + off(tap):  da[q', ci] = sum_tap dy[q' - off(tap)] W[tap]
      if (lane < taps) {
        float mine = part[0];
#pragma unroll
        for (int t = 1; t < 9; ++t) mine = lane == t ? part[t] : mine;
        const int ky = lane / a.ks, kx = lane - ky * a.ks;
        const int yy = y + (ky - pad), xx = x + (kx - pad);
        if (yy >= 0 && yy < a.h && xx >= 0 && xx < a.w)
          atomicAdd(a.view.base + comp_off(a.view, a.mask, b, yy, xx, ci), mine);
      }
    }
  }
}

static int launch_head_dgrad(const Dgrad3Args& a, cudaStream_t st) {
  const size_t smem = (size_t)a.ks * a.ks * a.CN * a.CW * sizeof(float);
  if (a.CW > 256 || smem > 160 * 1024 || (a.ks != 3 && a.ks != 1)) return CNF_NOT_ELIGIBLE;
  const long long total = (long long)a.B * a.h * a.w;
  dim3 grid((unsigned)std::min<long long>((total + 7) / 8, 148 * 8), 2);
#define CNF_LAUNCH_HD(WPL)                                                                                     \
  {                                                                                                            \
    static SmemAttrCache cache;                                                                                \
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)head_dgrad_kernel<WPL>, smem, cache));                         \
    head_dgrad_kernel<WPL><<<grid, 256, smem, st>>>(a);                                                        \
  }
  if (a.CW <= 32) CNF_LAUNCH_HD(1)
  else if (a.CW <= 64) CNF_LAUNCH_HD(2)
  else if (a.CW <= 128) CNF_LAUNCH_HD(4)
  else CNF_LAUNCH_HD(8)
#undef CNF_LAUNCH_HD
  return (int)cudaGetLastError();
}

static int launch_stem_dgrad(const Dgrad3Args& a, cudaStream_t st) {
  const size_t smem = (size_t)a.ks * a.ks * a.CN * a.CW * sizeof(float);
  if (a.CW > 256 || smem > 160 * 1024 || (a.ks != 3 && a.ks != 1)) return CNF_NOT_ELIGIBLE;
  const long long total = (long long)a.B * a.h * a.w;
  dim3 grid((unsigned)std::min<long long>((total + 7) / 8, 148 * 8), 2);
#define CNF_LAUNCH_SD(WPL)                                                                                     \
  {                                                                                                            \
    static SmemAttrCache cache;                                                                                \
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)stem_dgrad_kernel<WPL>, smem, cache));                         \
    stem_dgrad_kernel<WPL><<<grid, 256, smem, st>>>(a);                                                        \
  }
  if (a.CW <= 32) CNF_LAUNCH_SD(1)
  else if (a.CW <= 64) CNF_LAUNCH_SD(2)
  else if (a.CW <= 128) CNF_LAUNCH_SD(4)
  else CNF_LAUNCH_SD(8)
#undef CNF_LAUNCH_SD
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Grouped dilated 3x3 weight gradient of ONE dilation branch (F:387-413, F:565-590):
//   dW[g][ky][kx][ci][co] += sum_{b,q} a2[b, q + d*off(ky,kx), g*G + ci] dy[b, q, out_off + g*G + co]
// a2 = LN2(lrelu(y1)) applied while staging.  Thread = (channel c = g*G + ci, ky): 3 x G accumulators.
// ------------------------------------------------------------------------------------------
struct WgradGcArgs {
  const float* x;   // Y1 [2][B][hw][nk]
  long long x_net_stride;
  const float* dy;  // dY2 [2][B][hw][cat]
  long long dy_net_stride;
  const float* params;
  float* grads;
  long long net_stride, w_off, b_off, g_off, be_off;
  const double* stats;
  int B, h, w, nk, cat, ln;
  int dil, groups, out_off;
  int TH, TW, tiles_y, tiles_x, SB, teams;
  int gchunk;   // groups per CTA (blockIdx.y selects the chunk): wide branches whose channels exceed the thread budget are split
};

// threads per CTA: a thread of the 32-wide variant holds 3 x 32 accumulators and needs the larger register budget
__host__ __device__ constexpr int wgrad_gc_max_nt(int G, int KYS) { return KYS == 3 ? 256 : (G >= 32 ? 384 : 768); }

template <int G, int KYS>   // KYS = 3: a thread owns all 9 taps of its channel; KYS = 1: one kernel row (wide groups)
__global__ void __launch_bounds__(wgrad_gc_max_nt(G, KYS)) wgrad_gconv_kernel(const WgradGcArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, NT = blockDim.x;
  const int net = blockIdx.z;
  const int g0 = blockIdx.y * a.gchunk;                // first group of this CTA's chunk
  const int Cb = min(a.gchunk, a.groups - g0) * G;     // channels of the chunk
  const int cbase = g0 * G;
  const int tiles = a.tiles_y * a.tiles_x;
  const int tile = blockIdx.x % tiles, sb = blockIdx.x / tiles;
  const int y0 = (tile / a.tiles_x) * a.TH, x0 = (tile % a.tiles_x) * a.TW;
  const int th = min(a.TH, a.h - y0), tw = min(a.TW, a.w - x0);
  const int d = a.dil;
  const int SH = a.TH + 2 * d, SW = a.TW + 2 * d;
  float* a_s = smem;                                   // [SH][SW][Cb]
  float* d_s = smem + ((SH * SW * Cb + 3) & ~3);       // [TH][TW][Cb]
  const float* P = a.params + (long long)net * a.net_stride;
  const float* gam = P + a.g_off + cbase;
  const float* bet = P + a.be_off + cbase;
  const int tpt = (3 / KYS) * Cb;
  const int team = tid / tpt, t = tid % tpt;
  const bool worker = team < a.teams;
  const int c = t % Cb, ky0 = t / Cb;                  // ky0 = 0 when KYS == 3
  const int g = c / G;
  const bool vec = (Cb % 4 == 0) && (a.nk % 4 == 0) && (a.cat % 4 == 0) && (a.out_off % 4 == 0);

  float acc[KYS * 3][G];
  float bacc[G];
#pragma unroll
  for (int k = 0; k < KYS * 3; ++k)
#pragma unroll
    for (int j = 0; j < G; ++j) acc[k][j] = 0.f;
#pragma unroll
  for (int j = 0; j < G; ++j) bacc[j] = 0.f;
  const bool do_bias = worker && ky0 == 0 && (c % G) == 0;

  const int b0 = sb * a.SB, b1 = min(a.B, b0 + a.SB);
  for (int b = b0; b < b1; ++b) {
    float mean = 0.f, rstd = 1.f;
    if (a.ln) ln_coeffs(a.stats, (long long)net * a.B + b, (double)a.h * a.w * (double)a.nk, mean, rstd);
    const float* xs = a.x + (long long)net * a.x_net_stride + (long long)b * a.h * a.w * a.nk + cbase;
    const float* ds = a.dy + (long long)net * a.dy_net_stride + (long long)b * a.h * a.w * a.cat + a.out_off + cbase;
    __syncthreads();
    if (vec) {
      const int q = Cb >> 2;
      for (int idx = tid; idx < SH * SW * q; idx += NT) {
        const int cq = idx % q, pix = idx / q;
        const int sy = pix / SW, sx = pix - sy * SW;
        const int gy = y0 - d + sy, gx = x0 - d + sx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w) {
          const long long e = ((long long)gy * a.w + gx) * a.nk + 4 * cq;
          v = ld4(xs + e);
          v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w);
          if (a.ln) {
            const float4 gg = ld4(gam + e), be = ld4(bet + e);
            v.x = (v.x - mean) * rstd * gg.x + be.x;
            v.y = (v.y - mean) * rstd * gg.y + be.y;
            v.z = (v.z - mean) * rstd * gg.z + be.z;
            v.w = (v.w - mean) * rstd * gg.w + be.w;
          }
        }
        st4(a_s + pix * Cb + 4 * cq, v);
      }
      for (int idx = tid; idx < a.TH * a.TW * q; idx += NT) {
        const int cq = idx % q, pix = idx / q;
        const int py = pix / a.TW, px = pix - py * a.TW;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (py < th && px < tw) v = ld4(ds + ((long long)(y0 + py) * a.w + x0 + px) * a.cat + 4 * cq);
        st4(d_s + pix * Cb + 4 * cq, v);
      }
    } else {
      for (int idx = tid; idx < SH * SW * Cb; idx += NT) {
        const int ch = idx % Cb, pix = idx / Cb;
        const int gy = y0 - d + pix / SW, gx = x0 - d + pix % SW;
        float v = 0.f;
        if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w) {
          const long long e = ((long long)gy * a.w + gx) * a.nk + ch;
          v = lrelu(xs[e]);
          if (a.ln) v = (v - mean) * rstd * gam[e] + bet[e];
        }
        a_s[idx] = v;
      }
      for (int idx = tid; idx < a.TH * a.TW * Cb; idx += NT) {
        const int ch = idx % Cb, pix = idx / Cb;
        const int py = pix / a.TW, px = pix % a.TW;
        float v = 0.f;
        if (py < th && px < tw) v = ds[((long long)(y0 + py) * a.w + x0 + px) * a.cat + ch];
        d_s[idx] = v;
      }
    }
    __syncthreads();
    if (worker) {
      const int np = th * tw;
      for (int p = team; p < np; p += a.teams) {
        const int py = p / tw, px = p - py * tw;
        const float* dp = d_s + (py * a.TW + px) * Cb + g * G;
        float dv[G];
        if (G % 4 == 0) {
#pragma unroll
          for (int j = 0; j < G; j += 4) {
            const float4 t4 = ld4(dp + j);
            dv[j] = t4.x; dv[j + 1] = t4.y; dv[j + 2] = t4.z; dv[j + 3] = t4.w;
          }
        } else if (G == 2) {
          const float2 t2 = *reinterpret_cast<const float2*>(dp);
          dv[0] = t2.x; dv[G - 1] = t2.y;
        } else {
          dv[0] = dp[0];
        }
#pragma unroll
        for (int kyi = 0; kyi < KYS; ++kyi) {
          const float* ap = a_s + ((py + (ky0 + kyi) * d) * SW + px) * Cb + c;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float av = ap[kx * d * Cb];
#pragma unroll
            for (int j = 0; j < G; ++j) acc[kyi * 3 + kx][j] = fmaf(av, dv[j], acc[kyi * 3 + kx][j]);
          }
        }
        if (do_bias) {
#pragma unroll
          for (int j = 0; j < G; ++j) bacc[j] += dv[j];
        }
      }
    }
  }
  // reduce the teams through shared memory (the staging tiles are dead), then one global atomic per output and CTA
  __syncthreads();
  const int n_out = Cb * 9 * G;                        // packed [group][ky][kx][gin][gout] == [c / G][tap][c % G][co]
  float* red_w = smem;
  float* red_b = smem + n_out;
  for (int i = tid; i < n_out + Cb; i += NT) smem[i] = 0.f;
  __syncthreads();
  if (worker) {
    const int ci = c % G;
#pragma unroll
    for (int kyi = 0; kyi < KYS; ++kyi)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < G; ++j)
          atomicAdd(red_w + ((g * 9 + (ky0 + kyi) * 3 + kx) * G + ci) * G + j, acc[kyi * 3 + kx][j]);
    if (do_bias) {
#pragma unroll
      for (int j = 0; j < G; ++j) atomicAdd(red_b + g * G + j, bacc[j]);
    }
  }
  __syncthreads();
  float* gW = a.grads + (long long)net * a.net_stride + a.w_off + (long long)g0 * 9 * G * G;
  float* gB = a.grads + (long long)net * a.net_stride + a.b_off + cbase;
  for (int i = tid; i < n_out; i += NT) atomicAdd(gW + i, red_w[i]);
  for (int i = tid; i < Cb; i += NT) atomicAdd(gB + i, red_b[i]);
}

template <int G, int KYS>
static int launch_wgrad_gconv_t(WgradGcArgs a, cudaStream_t st) {
  const int d = a.dil;
  const int max_nt = wgrad_gc_max_nt(G, KYS);
  // all groups of the branch in one CTA when its channels fit the thread budget, else chunks of whole groups (16-byte aligned)
  a.gchunk = std::min(a.groups, max_nt / ((3 / KYS) * G));
  if (a.gchunk < 1) return CNF_NOT_ELIGIBLE;
  if (a.gchunk < a.groups && (a.gchunk * G) % 4) return CNF_NOT_ELIGIBLE;
  const int n_chunks = (a.groups + a.gchunk - 1) / a.gchunk;
  const int Cb = a.gchunk * G;
  const int tpt = (3 / KYS) * Cb;
  a.TW = std::min(a.w, 32);
  a.TH = std::min(a.h, 32);
  auto bytes = [&](int th) { return (size_t)((((th + 2 * d) * (a.TW + 2 * d) * Cb + 3) & ~3) + th * a.TW * Cb) * sizeof(float); };
  while (a.TH > 1 && bytes(a.TH) > 72 * 1024) --a.TH;
  if (bytes(a.TH) > 220 * 1024) return (int)cudaErrorInvalidConfiguration;
  a.tiles_y = (a.h + a.TH - 1) / a.TH;
  a.tiles_x = (a.w + a.TW - 1) / a.TW;
  const int tiles = a.tiles_y * a.tiles_x;
  a.teams = std::max(1, std::min(a.TH * a.TW, std::min(max_nt, KYS == 3 ? 256 : 384) / tpt));
  const int NT = std::max(64, ((tpt * a.teams + 31) / 32) * 32);
  const int want = std::max(1, 148 * 6 / (2 * tiles));
  a.SB = std::max(1, (a.B + want - 1) / want);
  const int sbs = (a.B + a.SB - 1) / a.SB;
  const size_t smem = std::max(bytes(a.TH), (size_t)(Cb * 9 * G + Cb) * sizeof(float));
  auto kern = wgrad_gconv_kernel<G, KYS>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  dim3 grid(tiles * sbs, n_chunks, 2);
  kern<<<grid, NT, smem, st>>>(a);
  return (int)cudaGetLastError();
}

template <int G>
static int launch_wgrad_gconv_g(const WgradGcArgs& a, cudaStream_t st) {
  if constexpr (G <= 8) {
    const int rc = launch_wgrad_gconv_t<G, 3>(a, st);
    if (rc != CNF_NOT_ELIGIBLE) return rc;
  }
  const int rc = launch_wgrad_gconv_t<G, 1>(a, st);
  return rc == CNF_NOT_ELIGIBLE ? (int)cudaErrorInvalidConfiguration : rc;
}

static int launch_wgrad_gconv(const WgradGcArgs& a, int G, cudaStream_t st) {
  switch (G) {
    case 1: return launch_wgrad_gconv_g<1>(a, st);
    case 2: return launch_wgrad_gconv_g<2>(a, st);
    case 4: return launch_wgrad_gconv_g<4>(a, st);
    case 8: return launch_wgrad_gconv_g<8>(a, st);
    case 16: return launch_wgrad_gconv_g<16>(a, st);
    case 32: return launch_wgrad_gconv_g<32>(a, st);      // config 5 (light): nk 64 / cardinality 2
    default: return (int)cudaErrorInvalidConfiguration;
  }
}

// ------------------------------------------------------------------------------------------
// Generic data gradient of the grouped dilated convs (fallback for group shapes the register-blocked
// gconv3 kernel does not cover): one thread per (net, b, pixel, input channel), all branches.
// ------------------------------------------------------------------------------------------
struct GcDgradBranch { int dil, groups, gin, gout, out_off; long long w_off; };
struct GcDgradArgs {
  const float* dy;   // [2][B][hw][cat]
  float* da;         // [2][B][hw][nk]
  const float* params;
  long long net_stride;
  int B, h, w, nk, cat, ks, n_br;
  unsigned mask;     // bit i = add branch i; the result is ADDED to da (the other branches were written by the specialised kernels)
  GcDgradBranch br[CNF_MAX_BRANCHES];
};

__global__ void __launch_bounds__(256) gconv_dgrad_naive_kernel(const GcDgradArgs a) {
  const long long per_net = (long long)a.B * a.h * a.w * a.nk;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * per_net) return;
  const int net = (int)(idx / per_net);
  const long long r = idx % per_net;
  const int c = (int)(r % a.nk);
  const long long pix = r / a.nk;
  const int x = (int)(pix % a.w), y = (int)((pix / a.w) % a.h), b = (int)(pix / ((long long)a.w * a.h));
  const int pad = (a.ks - 1) / 2;
  const float* P = a.params + (long long)net * a.net_stride;
  const float* src = a.dy + ((long long)net * a.B + b) * a.h * a.w * a.cat;
  float acc = 0.f;
  for (int i = 0; i < a.n_br; ++i) {
    const GcDgradBranch& br = a.br[i];
    if (!((a.mask >> i) & 1u)) continue;
    if (c >= br.groups * br.gin) continue;            // this branch reads only the first nk//d channels (F:579-586)
    const int g = c / br.gin, ci = c % br.gin;
    const float* W = P + br.w_off + (long long)g * a.ks * a.ks * br.gin * br.gout;
    for (int ky = 0; ky < a.ks; ++ky) {
      const int py = y - (ky - pad) * br.dil;
      if (py < 0 || py >= a.h) continue;
      for (int kx = 0; kx < a.ks; ++kx) {
        const int px = x - (kx - pad) * br.dil;
        if (px < 0 || px >= a.w) continue;
        const float* d = src + ((long long)py * a.w + px) * a.cat + br.out_off + g * br.gout;
        const float* wt = W + ((long long)(ky * a.ks + kx) * br.gin + ci) * br.gout;
        for (int co = 0; co < br.gout; ++co) acc = fmaf(d[co], wt[co], acc);
      }
    }
  }
  a.da[idx] += acc;
}

static int dgrad_gconv_naive(const cnf_coupling* c, int r, const float* params, const float* dY, float* dA, int B,
                             unsigned mask, cudaStream_t st) {
  const ResBlockLayout& L = c->rb[r];
  GcDgradArgs a = {};
  a.dy = dY; a.da = dA; a.params = params; a.net_stride = c->net_stride;
  a.B = B; a.h = c->h; a.w = c->w; a.nk = c->nk; a.cat = c->cat; a.ks = c->ks; a.n_br = (int)L.br.size();
  a.mask = mask;
  for (int i = 0; i < a.n_br; ++i) {
    const Branch& s = L.br[i];
    a.br[i].dil = s.dil; a.br[i].groups = s.groups; a.br[i].gin = s.gin; a.br[i].gout = s.gout;
    a.br[i].out_off = s.out_off; a.br[i].w_off = s.w_off;
  }
  const long long total = 2LL * B * c->hw() * c->nk;
  gconv_dgrad_naive_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Adam (keras, non-amsgrad; C:567, P:130): lr_t = lr sqrt(1-b2^t)/(1-b1^t) is computed by the caller.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr_t, float b1, float b2,
                                                   float eps, float gscale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps,
                float gscale, void* stream) {
  if (n <= 0) return 0;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_t, b1, b2, eps, gscale);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------
int64_t coupling_saved_bytes(const cnf_coupling* c, int64_t B) {
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  const int64_t hw = c->hw();
  int64_t b = 0;
  b += (int64_t)(c->R + 1) * al(2 * B * hw * c->nk * 4);
  b += (int64_t)c->R * al(2 * B * hw * c->nk * 4);
  b += (int64_t)c->R * al(2 * B * hw * c->cat * 4);
  b += al((int64_t)(c->n_ln() + 1) * 2 * B * 2 * 8);
  b += al(B * hw * c->c2 * 4);
  return b;
}

CouplingSaved carve_saved(const cnf_coupling* c, int64_t B, void* mem) {
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  const int64_t hw = c->hw();
  char* p = (char*)mem;
  CouplingSaved s;
  for (int r = 0; r <= c->R; ++r) { s.X.push_back((float*)p); p += al(2 * B * hw * c->nk * 4); }
  for (int r = 0; r < c->R; ++r) { s.Y1.push_back((float*)p); p += al(2 * B * hw * c->nk * 4); }
  for (int r = 0; r < c->R; ++r) { s.Y2.push_back((float*)p); p += al(2 * B * hw * c->cat * 4); }
  s.stats = (double*)p;
  p += al((int64_t)(c->n_ln() + 1) * 2 * B * 2 * 8);
  s.TH = (float*)p;
  s.state = nullptr;
  return s;
}

int64_t coupling_bwd_scratch_bytes(const cnf_coupling* c, int64_t B) {
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  const int64_t hw = c->hw();
  const int64_t wide = std::max(c->nk, c->cat);
  return al(2 * B * hw * c->c2 * 4) + al(2 * B * hw * c->nk * 4) + 2 * al(2 * B * hw * wide * 4) + al(4 * B * 8) +
         al(wgrad_tc_scratch_bytes(c->nk, c->cat, B, (int)hw));
}

int run_coupling_backward(const cnf_coupling* c, const float* params, float* grads, const CouplingSaved& sv,
                          FlowView g_view, FlowView s_view, int B, float inv_batch, void* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0) return 0;
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  const int hw = c->hw(), nk = c->nk, cat = c->cat, c2 = c->c2, R = c->R;
  const int64_t wide = std::max(nk, cat);
  char* sp = (char*)scratch;
  float* DR = (float*)sp; sp += al(2LL * B * hw * c2 * 4);
  float* GX = (float*)sp; sp += al(2LL * B * hw * nk * 4);
  float* GA = (float*)sp; sp += al(2LL * B * hw * wide * 4);
  float* GY = (float*)sp; sp += al(2LL * B * hw * wide * 4);
  double* bst = (double*)sp; sp += al(4LL * B * 8);
  float* WP = wgrad_tc_scratch_bytes(nk, cat, B, hw) ? (float*)sp : nullptr;   // per-CTA partials of the tensor-core weight gradients
  // 1x1 weight gradients: tcgen05 kernel + ordered reduction where the shape fits, else the FFMA kernel with fp32 atomics
  auto wgrad_pw = [&](const WgradPwArgs& a) -> int {
    if (!(c->paths & CNF_PATH_NO_TCGEN05)) {
      const int rc = launch_wgrad_pw_tc(a, WP, st);
      if (rc != CNF_NOT_ELIGIBLE) return rc;
    }
    return launch_wgrad_pw(a, st);
  };
  const long long slot = 2LL * B * 2;
  auto stats = [&](int i) -> const double* { return c->n_ln() ? sv.stats + slot * i : nullptr; };
  const long long ns = c->net_stride;

  {  // coupling law
    HeadBwdArgs a = {};
    a.g = g_view; a.s = s_view; a.mask_c = c->mask_c; a.B = B; a.h = c->h; a.w = c->w; a.c2 = c2;
    a.TH = sv.TH; a.tanh_w = params + c->tanh_w; a.dtanh_w = grads + c->tanh_w;
    a.DR = DR; a.dr_net_stride = (long long)B * hw * c2; a.invB = inv_batch;
    const long long total = (long long)B * hw * c2;
    head_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    CU_TRY(cudaGetLastError());
  }
  {  // head conv: weights, then data
    Wgrad3Args a = {};
    a.x = sv.X[R]; a.x_net_stride = (long long)B * hw * nk; a.mode = 0;
    a.dy = DR; a.dy_net_stride = (long long)B * hw * c2;
    a.params = params; a.grads = grads; a.net_stride = ns; a.w_off = c->head_w; a.b_off = c->head_b;
    a.g_off = c->lnf_g; a.be_off = c->lnf_b; a.stats = stats(3 * R);
    a.B = B; a.h = c->h; a.w = c->w; a.CI = nk; a.CO = c2; a.ks = c->ks; a.ln = c->ln;
    Wgrad3WideArgs ww = {};
    ww.wide = sv.X[R]; ww.wide_net_stride = (long long)B * hw * nk; ww.CW = nk; ww.wide_act = 1;
    ww.narrow = DR; ww.narrow_net_stride = (long long)B * hw * c2; ww.CN = c2; ww.sign = -1; ww.wide_is_ci = 1;
    ww.params = params; ww.grads = grads; ww.net_stride = ns; ww.w_off = c->head_w; ww.b_off = c->head_b;
    ww.g_off = c->lnf_g; ww.be_off = c->lnf_b; ww.stats = stats(3 * R);
    ww.B = B; ww.h = c->h; ww.w = c->w; ww.ks = c->ks; ww.ln = c->ln;
    int rc = launch_wgrad3_wide(ww, st);
    if (rc == CNF_NOT_ELIGIBLE) rc = launch_wgrad3(a, st);
    CU_TRY(rc);
    Dgrad3Args dg = {};
    dg.in = DR; dg.in_net_stride = (long long)B * hw * c2;
    dg.params = params; dg.net_stride = ns; dg.w_off = c->head_w;
    dg.B = B; dg.h = c->h; dg.w = c->w; dg.CW = nk; dg.CN = c2; dg.ks = c->ks;
    dg.out = GA; dg.out_net_stride = (long long)B * hw * nk;
    rc = launch_head_dgrad(dg, st);
    if (rc == CNF_NOT_ELIGIBLE) {
      Conv3tArgs t = {};
      t.in = DR; t.in_net_stride = (long long)B * hw * c2;
      t.params = params; t.net_stride = ns; t.w_off = c->head_w;
      t.B = B; t.h = c->h; t.w = c->w; t.CI = nk; t.CO = c2; t.ks = c->ks; t.mode = 0;
      t.out = GA; t.out_net_stride = (long long)B * hw * nk;
      const long long total = 2LL * B * hw * nk;
      conv3t_small_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t);
      rc = (int)cudaGetLastError();
    }
    CU_TRY(rc);
    CU_TRY(ln_backward(GA, sv.X[R], GX, params, grads, ns, c->lnf_g, c->lnf_b, stats(3 * R), bst, B, (long long)hw * nk,
                       c->ln, 0, st));
  }
  for (int r = R - 1; r >= 0; --r) {
    const ResBlockLayout& L = c->rb[r];
    {  // pw2: x_{r+1} = x_r + conv1x1(LN3(lrelu(y2)))
      WgradPwArgs a = {};
      a.x = sv.Y2[r]; a.x_net_stride = (long long)B * hw * cat;
      a.dy = GX; a.dy_net_stride = (long long)B * hw * nk;
      a.params = params; a.grads = grads; a.net_stride = ns; a.w_off = L.pw2_w; a.b_off = L.pw2_b;
      a.g_off = L.ln3_g; a.be_off = L.ln3_b; a.stats = stats(3 * r + 2);
      a.B = B; a.hw = hw; a.K = cat; a.N = nk; a.ln = c->ln;
      CU_TRY(wgrad_pw(a));
      CU_TRY(dgrad_pw(params, ns, L.pw2_w, GX, GA, B, hw, cat, nk, st, c->paths));
      CU_TRY(ln_backward(GA, sv.Y2[r], GY, params, grads, ns, L.ln3_g, L.ln3_b, stats(3 * r + 2), bst, B,
                         (long long)hw * cat, c->ln, 0, st));
    }
    {  // grouped dilated convs
      for (const Branch& s : L.br) {
        if (s.gin != s.gout || c->ks != 3) return (int)cudaErrorInvalidConfiguration;
        WgradGcArgs a = {};
        a.x = sv.Y1[r]; a.x_net_stride = (long long)B * hw * nk;
        a.dy = GY; a.dy_net_stride = (long long)B * hw * cat;
        a.params = params; a.grads = grads; a.net_stride = ns; a.w_off = s.w_off; a.b_off = s.b_off;
        a.g_off = L.ln2_g; a.be_off = L.ln2_b; a.stats = stats(3 * r + 1);
        a.B = B; a.h = c->h; a.w = c->w; a.nk = nk; a.cat = cat; a.ln = c->ln;
        a.dil = s.dil; a.groups = s.groups; a.out_off = s.out_off;
        CU_TRY(launch_wgrad_gconv(a, s.gin, st));
      }
      {
        unsigned leftover = 0;
        CU_TRY(dgrad_gconv(c, r, params, GY, GA, B, st, &leftover));
        if (leftover) CU_TRY(dgrad_gconv_naive(c, r, params, GY, GA, B, leftover, st));
      }
      CU_TRY(ln_backward(GA, sv.Y1[r], GY, params, grads, ns, L.ln2_g, L.ln2_b, stats(3 * r + 1), bst, B,
                         (long long)hw * nk, c->ln, 0, st));
    }
    {  // pw1, then the residual branch joins the skip path: GX += d x_r
      WgradPwArgs a = {};
      a.x = sv.X[r]; a.x_net_stride = (long long)B * hw * nk;
      a.dy = GY; a.dy_net_stride = (long long)B * hw * nk;
      a.params = params; a.grads = grads; a.net_stride = ns; a.w_off = L.pw1_w; a.b_off = L.pw1_b;
      a.g_off = L.ln1_g; a.be_off = L.ln1_b; a.stats = stats(3 * r);
      a.B = B; a.hw = hw; a.K = nk; a.N = nk; a.ln = c->ln;
      CU_TRY(wgrad_pw(a));
      CU_TRY(dgrad_pw(params, ns, L.pw1_w, GY, GA, B, hw, nk, nk, st, c->paths));
      CU_TRY(ln_backward(GA, sv.X[r], GX, params, grads, ns, L.ln1_g, L.ln1_b, stats(3 * r), bst, B, (long long)hw * nk,
                         c->ln, 1, st));
    }
  }
  {  // stem: weights from the saved layer input, data gradient added into the u1 half of G
    Wgrad3Args a = {};
    a.view = s_view; a.mask = c->mask; a.mode = 1;
    a.dy = GX; a.dy_net_stride = (long long)B * hw * nk;
    a.params = params; a.grads = grads; a.net_stride = ns; a.w_off = c->stem_w; a.b_off = c->stem_b;
    a.B = B; a.h = c->h; a.w = c->w; a.CI = c->c1; a.CO = nk; a.ks = c->ks; a.ln = 0;
    Wgrad3WideArgs ww = {};
    ww.wide = GX; ww.wide_net_stride = (long long)B * hw * nk; ww.CW = nk; ww.wide_act = 0;
    ww.view = s_view; ww.mask = c->mask; ww.narrow_view = 1; ww.CN = c->c1; ww.sign = +1; ww.wide_is_ci = 0;
    ww.params = params; ww.grads = grads; ww.net_stride = ns; ww.w_off = c->stem_w; ww.b_off = c->stem_b;
    ww.B = B; ww.h = c->h; ww.w = c->w; ww.ks = c->ks; ww.ln = 0;
    int rc = launch_wgrad3_wide(ww, st);
    if (rc == CNF_NOT_ELIGIBLE) rc = launch_wgrad3(a, st);
    CU_TRY(rc);
    Dgrad3Args dg = {};
    dg.in = GX; dg.in_net_stride = (long long)B * hw * nk;
    dg.params = params; dg.net_stride = ns; dg.w_off = c->stem_w;
    dg.B = B; dg.h = c->h; dg.w = c->w; dg.CW = nk; dg.CN = c->c1; dg.ks = c->ks;
    dg.view = g_view; dg.mask = c->mask;
    rc = launch_stem_dgrad(dg, st);
    if (rc == CNF_NOT_ELIGIBLE) {
      Conv3tArgs t = {};
      t.in = GX; t.in_net_stride = (long long)B * hw * nk;
      t.params = params; t.net_stride = ns; t.w_off = c->stem_w;
      t.B = B; t.h = c->h; t.w = c->w; t.CI = c->c1; t.CO = nk; t.ks = c->ks; t.mode = 1;
      t.view = g_view; t.mask = c->mask;
      const long long total = (long long)B * hw * c->c1;
      conv3t_small_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t);
      rc = (int)cudaGetLastError();
    }
    CU_TRY(rc);
  }
  return 0;
}

}  // namespace cnf
