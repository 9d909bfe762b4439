// s/t sub-network kernels of one coupling layer (fp32-exact path) and their orchestration.
//
// What the reference computes (conv_cINN_make_model.py M:1076-1213 with
// conv_cINN_base_functions.py F:330-362, F:364-413, F:501-627), per net (A and b):
//   x = Conv3x3(u1c)                                    "stem"  (M:1114, M:1159)
//   R x { a = LN(LReLU(x)); y1 = Conv1x1(a)             "pw1"   (F:552-563)
//         b = LN(LReLU(y1)); y2 = concat_d GroupedDilatedConv3x3(b[..., :nk//d])  "gconv" (F:565-590)
//         c = LN(LReLU(y2)); x = x + Conv1x1(c) }       "pw2"   (F:604-625)
//   out = Conv3x3(LN(LReLU(x)))                         "head"  (M:1133-1150); A = w*tanh(out) (M:1198)
// followed by the coupling law and the per-sample log-det (M:1307-1326 / M:1379-1394).
//
// How it is laid out here: LayerNorm spans the whole sample, so every conv kernel (i) applies the
// LReLU+LN of its INPUT while staging it into shared memory, using the per-sample (sum, sumsq) the
// producing kernel left behind, and (ii) accumulates the (sum, sumsq) of LReLU(output) for its
// consumer.  Both nets run in the same launches (blockIdx.z / in-CTA).  The masked gather of u1
// (M:723-759) is folded into the stem's loads, and the head's epilogue applies tanh*w, exp, the
// affine law and the decompress scatter (M:903-1071) straight into the flow buffer, so s, t and
// the compressed tensors never exist in HBM.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "cnf_internal.h"
#include "device_utils.cuh"

namespace cnf {

// ------------------------------------------------------------------------------------------
// 1. GEMM-shaped convs: stem (implicit im2col of the masked input) and 1x1 convs.
//    out[m, n] = sum_k A[m, k] W[k, n] + bias[n] (+ res[m, n]),  m = pixel of ONE sample.
// ------------------------------------------------------------------------------------------
struct GemmArgs {
  // A operand
  const float* in;            // PW: [2][B][hw][K]
  long long in_net_stride;
  FlowView view;              // STEM: masked gather source
  int mask, h, w, c1, ks;
  // parameters
  const float* params;        // layer base; net n at params + n*net_stride
  long long net_stride, w_off, b_off, g_off, be_off;
  const double* stats_in;     // [2][B][2] or nullptr
  double* stats_out;          // [2][B][2] or nullptr
  float* out;                 // [2][B][hw][N]
  const float* res;           // residual, same layout as out, may alias out; or nullptr
  long long out_net_stride;
  int B, hw, K, N, KC, ln;
  // backward-pass uses of gemm_kernel: raw_in = no LReLU/LN on A; w_trans = B[k][n] = W[n*ldw + k]; no_bias
  int raw_in, w_trans, ldw, no_bias;
  int dbg;   // timing experiments only (-DCNF_DEBUG builds, CNF_PW_DBG): 1 skip epilogue stores, 2 skip cp.async, 4 skip MMAs, 8 skip transform math
  int paths; // CNF_PATH_* bits of the layer: kernel families NOT to use
};

template <int TN, int NQ, int RM, bool STEM>
__global__ void __launch_bounds__(128) gemm_kernel(const GemmArgs a) {
  constexpr int NT = 128;
  constexpr int CT = TN / (4 * NQ);  // threads along n
  constexpr int RT = NT / CT;        // threads along m
  constexpr int TM = RM * RT;
  constexpr int RN = 4 * NQ;
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[64];

  const int tid = threadIdx.x;
  const int tn = tid % CT, tm = tid / CT;
  const int tiles_n = (a.N + TN - 1) / TN;
  const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
  const int b = blockIdx.y, net = blockIdx.z;
  const int m0 = tile_m * TM, n0 = tile_n * TN;
  const int KS = a.KC + 4;
  float* As = smem;             // [TM][KS]
  float* Bs = smem + TM * KS;   // [KC][TN]

  const float* P = a.params + (long long)net * a.net_stride;
  const float* Wg = P + a.w_off;
  const bool n_vec = (a.N % 4) == 0;
  const bool k_vec = (a.K % 4) == 0;

  float mean = 0.f, rstd = 1.f;
  if (!STEM && a.ln) ln_coeffs(a.stats_in, (long long)net * a.B + b, (double)a.hw * (double)a.K, mean, rstd);

  float acc[RM][RN];
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int c = 0; c < RN; ++c) acc[r][c] = 0.f;

  for (int kc0 = 0; kc0 < a.K; kc0 += a.KC) {
    const int kc = min(a.KC, ((a.K - kc0) + 3) & ~3);  // padded length of this K chunk
    __syncthreads();
    // ---- weights: Bs[k][n] = W[kc0+k][n0+n] (zero outside)
    for (int idx = tid; idx < kc * (TN / 4); idx += NT) {
      const int k = idx / (TN / 4), nq = idx % (TN / 4);
      const int gk = kc0 + k, n = n0 + nq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < a.K) {
        if (a.w_trans) {
          if (n + 0 < a.N) v.x = Wg[(long long)(n + 0) * a.ldw + gk];
          if (n + 1 < a.N) v.y = Wg[(long long)(n + 1) * a.ldw + gk];
          if (n + 2 < a.N) v.z = Wg[(long long)(n + 2) * a.ldw + gk];
          if (n + 3 < a.N) v.w = Wg[(long long)(n + 3) * a.ldw + gk];
        } else {
          const float* src = Wg + (long long)gk * a.N + n;
          if (n_vec && n + 3 < a.N) {
            v = ld4(src);
          } else {
            if (n + 0 < a.N) v.x = src[0];
            if (n + 1 < a.N) v.y = src[1];
            if (n + 2 < a.N) v.z = src[2];
            if (n + 3 < a.N) v.w = src[3];
          }
        }
      }
      st4(&Bs[k * TN + nq * 4], v);
    }
    // ---- activations
    if (STEM) {
      const int pad = (a.ks - 1) / 2;
      for (int idx = tid; idx < TM * kc; idx += NT) {
        const int m = idx / kc, k = idx % kc;
        const int gk = kc0 + k, p = m0 + m;
        float v = 0.f;
        if (p < a.hw && gk < a.K) {
          const int y = p / a.w, x = p % a.w;
          const int tap = gk / a.c1, ci = gk % a.c1;
          const int iy = y + tap / a.ks - pad, ix = x + tap % a.ks - pad;
          if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w)
            v = a.view.base[comp_off(a.view, a.mask, b, iy, ix, ci)];
        }
        As[m * KS + k] = v;
      }
    } else {
      const float* src_s = a.in + (long long)net * a.in_net_stride + (long long)b * a.hw * a.K;
      const float* gam = P + a.g_off;
      const float* bet = P + a.be_off;
      if (k_vec) {
        const int kq_n = kc / 4;
        for (int idx = tid; idx < TM * kq_n; idx += NT) {
          const int m = idx / kq_n, kq = idx % kq_n;
          const int p = m0 + m;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p < a.hw) {
            const long long e = (long long)p * a.K + kc0 + kq * 4;
            v = ld4(src_s + e);
            if (!a.raw_in) { v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w); }
            if (a.ln) {
              const float4 g = ld4(gam + e), be = ld4(bet + e);
              v.x = (v.x - mean) * rstd * g.x + be.x;
              v.y = (v.y - mean) * rstd * g.y + be.y;
              v.z = (v.z - mean) * rstd * g.z + be.z;
              v.w = (v.w - mean) * rstd * g.w + be.w;
            }
          }
          st4(&As[m * KS + kq * 4], v);
        }
      } else {
        for (int idx = tid; idx < TM * kc; idx += NT) {
          const int m = idx / kc, k = idx % kc;
          const int gk = kc0 + k, p = m0 + m;
          float v = 0.f;
          if (p < a.hw && gk < a.K) {
            const long long e = (long long)p * a.K + gk;
            v = a.raw_in ? src_s[e] : lrelu(src_s[e]);
            if (a.ln) v = (v - mean) * rstd * gam[e] + bet[e];
          }
          As[m * KS + k] = v;
        }
      }
    }
    __syncthreads();
    // ---- FFMA tile
    for (int k4 = 0; k4 < kc; k4 += 4) {
      float4 av[RM];
#pragma unroll
      for (int r = 0; r < RM; ++r) av[r] = ld4(&As[(tm + RT * r) * KS + k4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float bv[RN];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 t = ld4(&Bs[(k4 + kk) * TN + q * (TN / NQ) + tn * 4]);
          bv[q * 4 + 0] = t.x; bv[q * 4 + 1] = t.y; bv[q * 4 + 2] = t.z; bv[q * 4 + 3] = t.w;
        }
#pragma unroll
        for (int r = 0; r < RM; ++r) {
          const float ar = kk == 0 ? av[r].x : kk == 1 ? av[r].y : kk == 2 ? av[r].z : av[r].w;
#pragma unroll
          for (int c = 0; c < RN; ++c) acc[r][c] = fmaf(ar, bv[c], acc[r][c]);
        }
      }
    }
  }

  // ---- epilogue: bias (+ residual), store, stats of LReLU(out) for the consumer's LayerNorm
  const float* bias = P + a.b_off;
  float* out_s = a.out + (long long)net * a.out_net_stride + (long long)b * a.hw * a.N;
  const float* res_s = a.res ? a.res + (long long)net * a.out_net_stride + (long long)b * a.hw * a.N : nullptr;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int r = 0; r < RM; ++r) {
    const int p = m0 + tm + RT * r;
    if (p >= a.hw) continue;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int n = n0 + q * (TN / NQ) + tn * 4;
      if (n >= a.N) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = acc[r][q * 4 + j];
      const long long e = (long long)p * a.N + n;
      if (n_vec) {
        const float4 bb = a.no_bias ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(bias + n);
        o[0] += bb.x; o[1] += bb.y; o[2] += bb.z; o[3] += bb.w;
        if (res_s) {
          const float4 rr = ld4(res_s + e);
          o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
        }
        st4(out_s + e, make_float4(o[0], o[1], o[2], o[3]));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float l = lrelu(o[j]);
          s1 += l;
          s2 += l * l;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < a.N) {
            float v = o[j] + (a.no_bias ? 0.f : bias[n + j]);
            if (res_s) v += res_s[e + j];
            out_s[e + j] = v;
            const float l = lrelu(v);
            s1 += l;
            s2 += l * l;
          }
        }
      }
    }
  }
  if (a.stats_out) {
    double d1, d2;
    block_sum2(s1, s2, red, d1, d2);
    if (tid == 0) {
      double* so = a.stats_out + 2 * ((long long)net * a.B + b);
      atomicAdd(so, d1);
      atomicAdd(so + 1, d2);
    }
  }
}

}  // namespace cnf
#include "tc_kernels.cuh"
namespace cnf {

// ------------------------------------------------------------------------------------------
// 1b. 1x1 convs, multi-sample tiles.  A CTA owns PT pixels x S samples (M = PT*S = 256 rows):
//     gamma/beta of a pixel slot are loaded ONCE and shared by the S samples, and the S activation
//     loads of a slot are independent (S+2 128-bit loads in flight per thread), which removes both the
//     3x L2 traffic and the load-latency serialisation of the per-sample tiling above.
//     Rows are m = s*PT + p; thread (tm, tn) owns rows tm + PT*r, r = 0..S-1, i.e. pixel tm of every
//     sample, so LayerNorm statistics fall out per (thread, r) without a segmented reduction.
// ------------------------------------------------------------------------------------------
template <int TN, int NQ>
__global__ void __launch_bounds__(256, 2) pw_kernel(const GemmArgs a) {
  constexpr int NT = 256;
  constexpr int RN = 4 * NQ;
  constexpr int CT = TN / RN;   // threads along n
  constexpr int PT = NT / CT;   // pixels per tile (= threads along m)
  constexpr int S = 8;          // samples per tile (= rows per thread)
  extern __shared__ __align__(16) float smem[];
  __shared__ float mr[S][2];
  __shared__ float red[NT / 32][2 * S];

  const int tid = threadIdx.x;
  const int tn = tid % CT, tm = tid / CT;
  const int tiles_n = (a.N + TN - 1) / TN;
  const int tile_p = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
  const int s0 = blockIdx.y * S, net = blockIdx.z;
  const int p0 = tile_p * PT, n0 = tile_n * TN;
  const int KS = a.KC + 4;
  float* As = smem;                 // [S*PT][KS]
  float* Bs = smem + S * PT * KS;   // [KC][TN]
  const int ns = min(S, a.B - s0);  // valid samples in this tile

  const float* P = a.params + (long long)net * a.net_stride;
  const float* Wg = P + a.w_off;
  const float* gam = P + a.g_off;
  const float* bet = P + a.be_off;
  const bool n_vec = (a.N % 4) == 0;

  if (tid < S) {
    float mean = 0.f, rstd = 1.f;
    if (a.ln && tid < ns) ln_coeffs(a.stats_in, (long long)net * a.B + s0 + tid, (double)a.hw * (double)a.K, mean, rstd);
    mr[tid][0] = mean;
    mr[tid][1] = rstd;
  }

  float acc[S][RN];
#pragma unroll
  for (int r = 0; r < S; ++r)
#pragma unroll
    for (int c = 0; c < RN; ++c) acc[r][c] = 0.f;

  const float* src_n = a.in + (long long)net * a.in_net_stride;
  for (int kc0 = 0; kc0 < a.K; kc0 += a.KC) {
    const int kc = min(a.KC, a.K - kc0);  // K % 4 == 0 is a launch precondition
    __syncthreads();
    for (int idx = tid; idx < kc * (TN / 4); idx += NT) {
      const int k = idx / (TN / 4), nq = idx % (TN / 4);
      const int n = n0 + nq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = Wg + (long long)(kc0 + k) * a.N + n;
      if (n_vec && n + 3 < a.N) {
        v = ld4(src);
      } else {
        if (n + 0 < a.N) v.x = src[0];
        if (n + 1 < a.N) v.y = src[1];
        if (n + 2 < a.N) v.z = src[2];
        if (n + 3 < a.N) v.w = src[3];
      }
      st4(&Bs[k * TN + nq * 4], v);
    }
    const int kq_n = kc >> 2;
    for (int slot = tid; slot < PT * kq_n; slot += NT) {
      const int p = slot / kq_n, kq = slot - p * kq_n;
      const int gp = p0 + p;
      float4 g = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool pv = gp < a.hw;
      const long long e = (long long)gp * a.K + kc0 + kq * 4;
      if (pv && a.ln) {
        g = ld4(gam + e);
        be = ld4(bet + e);
      }
#pragma unroll
      for (int h = 0; h < S; h += 4) {   // two batches of 4 independent 128-bit loads (register budget)
        float4 xv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (pv && h + r < ns) xv[r] = ld4(src_n + ((long long)(s0 + h + r) * a.hw) * a.K + e);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pv && h + r < ns) {
            const float mean = mr[h + r][0], rstd = mr[h + r][1];
            v = xv[r];
            v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w);
            if (a.ln) {
              v.x = (v.x - mean) * rstd * g.x + be.x;
              v.y = (v.y - mean) * rstd * g.y + be.y;
              v.z = (v.z - mean) * rstd * g.z + be.z;
              v.w = (v.w - mean) * rstd * g.w + be.w;
            }
          }
          st4(&As[((h + r) * PT + p) * KS + kq * 4], v);
        }
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int k4 = 0; k4 < kc; k4 += 4) {
      float4 av[S];
#pragma unroll
      for (int r = 0; r < S; ++r) av[r] = ld4(&As[(tm + PT * r) * KS + k4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float bv[RN];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 t = ld4(&Bs[(k4 + kk) * TN + q * (TN / NQ) + tn * 4]);
          bv[q * 4 + 0] = t.x; bv[q * 4 + 1] = t.y; bv[q * 4 + 2] = t.z; bv[q * 4 + 3] = t.w;
        }
#pragma unroll
        for (int r = 0; r < S; ++r) {
          const float ar = kk == 0 ? av[r].x : kk == 1 ? av[r].y : kk == 2 ? av[r].z : av[r].w;
#pragma unroll
          for (int c = 0; c < RN; ++c) acc[r][c] = fmaf(ar, bv[c], acc[r][c]);
        }
      }
    }
  }

  // epilogue: bias (+ residual), store, per-sample stats of LReLU(out)
  const float* bias = P + a.b_off;
  const int gp = p0 + tm;
  float st1[S], st2[S];
#pragma unroll
  for (int r = 0; r < S; ++r) {
    st1[r] = 0.f;
    st2[r] = 0.f;
    if (gp >= a.hw || r >= ns) continue;
    const long long row = ((long long)(s0 + r) * a.hw + gp) * a.N;
    float* out_r = a.out + (long long)net * a.out_net_stride + row;
    const float* res_r = a.res ? a.res + (long long)net * a.out_net_stride + row : nullptr;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int n = n0 + q * (TN / NQ) + tn * 4;
      if (n >= a.N) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = acc[r][q * 4 + j];
      if (n_vec) {
        const float4 bb = ld4(bias + n);
        o[0] += bb.x; o[1] += bb.y; o[2] += bb.z; o[3] += bb.w;
        if (res_r) {
          const float4 rr = ld4(res_r + n);
          o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
        }
        st4(out_r + n, make_float4(o[0], o[1], o[2], o[3]));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float l = lrelu(o[j]);
          st1[r] += l;
          st2[r] += l * l;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < a.N) {
            float v = o[j] + bias[n + j];
            if (res_r) v += res_r[n + j];
            out_r[n + j] = v;
            const float l = lrelu(v);
            st1[r] += l;
            st2[r] += l * l;
          }
      }
    }
  }
  if (a.stats_out) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int r = 0; r < S; ++r) {
      const float v1 = warp_sum(st1[r]), v2 = warp_sum(st2[r]);
      if (lane == 0) {
        red[wid][2 * r] = v1;
        red[wid][2 * r + 1] = v2;
      }
    }
    __syncthreads();
    if (tid < 2 * S && (tid >> 1) < ns) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) t += (double)red[w][tid];
      atomicAdd(a.stats_out + 2 * ((long long)net * a.B + s0 + (tid >> 1)) + (tid & 1), t);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 2. Grouped dilated 3x3 convs of one residual block, all dilation branches in one launch,
//    written straight into the concat layout (F:397-411, F:577-590).
// ------------------------------------------------------------------------------------------
struct GconvBranch {
  int dil, groups, gin, gout, out_off, first_item;
  int in_off;   // first input channel of the branch inside the input tensor (0 in the forward pass)
  long long w_off, b_off;
};
struct GconvArgs {
  const float* in;    // [2][B][hw][Cin]
  float* out;         // [2][B][hw][Cout]
  long long in_net_stride, out_net_stride;
  const float* params;
  long long net_stride, g_off, be_off;
  const double* stats_in;
  double* stats_out;
  int B, h, w, Cin, Cout, ln, ks;
  int TH, TW, tiles_y, tiles_x;
  int dbg;  // debug (-DCNF_DEBUG builds): bit0 skip global loads in staging, bit1 skip the FFMA taps
  int paths;  // CNF_PATH_* bits of the layer: kernel families NOT to use
  // backward-input mode: raw input (no LReLU/LN), flipped+transposed weights, accumulate into `out`, no bias, no stats
  int bwd;
  int n_br;
  GconvBranch br[CNF_MAX_BRANCHES];
};

constexpr int GC_NT = 256;
constexpr int GC_PX = 4;  // pixels per thread

__host__ __device__ inline int gc_stride(int gin) { return gin + ((gin % 8) == 0 ? 4 : 0); }

struct GcGeom {
  int dil, ks, gin, gout, cbase, w, Cout;  // cbase = first output channel of this group in the concat
};

template <int G>
__device__ __forceinline__ void gconv_compute(const GcGeom q, const float* in_s, const float* w_s,
                                              const float* b_s, int SW, int GS, int th, int tw, int y0,
                                              int x0, float* out_s, float& s1, float& s2) {
  const int tid = threadIdx.x;
  const int TP = th * tw;
  const int npx = (TP + GC_NT - 1) / GC_NT;
  const int d = q.dil, ks = q.ks;
  float acc[GC_PX][G];
  int poff[GC_PX];
#pragma unroll
  for (int j = 0; j < GC_PX; ++j) {
    const int p = min(tid + j * GC_NT, TP - 1);
    poff[j] = ((p / tw) * SW + (p % tw)) * GS;
#pragma unroll
    for (int co = 0; co < G; ++co) acc[j][co] = 0.f;
  }
  for (int ky = 0; ky < ks; ++ky) {
    for (int kx = 0; kx < ks; ++kx) {
      const int toff = (ky * d * SW + kx * d) * GS;
      const float* wt = w_s + (ky * ks + kx) * G * G;
      float xv[GC_PX][G];
#pragma unroll
      for (int j = 0; j < GC_PX; ++j) {
        if (j < npx) {
          const float* src = in_s + poff[j] + toff;
          if (G % 4 == 0) {
#pragma unroll
            for (int c4 = 0; c4 < G; c4 += 4) {
              const float4 t = ld4(src + c4);
              xv[j][c4] = t.x; xv[j][c4 + 1] = t.y; xv[j][c4 + 2] = t.z; xv[j][c4 + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int c = 0; c < G; ++c) xv[j][c] = src[c];
          }
        }
      }
#pragma unroll
      for (int ci = 0; ci < G; ++ci) {
        float wv[G];
        if (G % 4 == 0) {
#pragma unroll
          for (int c4 = 0; c4 < G; c4 += 4) {
            const float4 t = ld4(wt + ci * G + c4);
            wv[c4] = t.x; wv[c4 + 1] = t.y; wv[c4 + 2] = t.z; wv[c4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < G; ++c) wv[c] = wt[ci * G + c];
        }
#pragma unroll
        for (int j = 0; j < GC_PX; ++j) {
          if (j < npx) {
#pragma unroll
            for (int co = 0; co < G; ++co) acc[j][co] = fmaf(xv[j][ci], wv[co], acc[j][co]);
          }
        }
      }
    }
  }
  const int cbase = q.cbase;
  const bool vec = (G % 4 == 0) && (q.Cout % 4 == 0) && (cbase % 4 == 0);
#pragma unroll
  for (int j = 0; j < GC_PX; ++j) {
    const int p = tid + j * GC_NT;
    if (j < npx && p < TP) {
      const int y = y0 + p / tw, x = x0 + p % tw;
      float* dst = out_s + ((long long)y * q.w + x) * q.Cout + cbase;
      float o[G];
#pragma unroll
      for (int co = 0; co < G; ++co) {
        o[co] = acc[j][co] + b_s[co];
        const float l = lrelu(o[co]);
        s1 += l;
        s2 += l * l;
      }
      if (vec) {
#pragma unroll
        for (int c4 = 0; c4 < G; c4 += 4) st4(dst + c4, make_float4(o[c4], o[c4 + 1], o[c4 + 2], o[c4 + 3]));
      } else {
#pragma unroll
        for (int co = 0; co < G; ++co) dst[co] = o[co];
      }
    }
  }
}

// any (gin, gout): one pixel per pass, outputs in chunks of 8 (cardinality 1, wide groups)
__device__ __noinline__ void gconv_compute_generic(const GcGeom q, const float* in_s, const float* w_s,
                                                   const float* b_s, int SW, int GS, int th, int tw, int y0,
                                                   int x0, float* out_s, float* s12) {
  float s1 = 0.f, s2 = 0.f;
  const int TP = th * tw;
  const int d = q.dil, ks = q.ks, gin = q.gin, gout = q.gout;
  const int cbase = q.cbase;
  for (int p = threadIdx.x; p < TP; p += GC_NT) {
    const int poff = ((p / tw) * SW + (p % tw)) * GS;
    const int y = y0 + p / tw, x = x0 + p % tw;
    float* dst = out_s + ((long long)y * q.w + x) * q.Cout + cbase;
    for (int co0 = 0; co0 < gout; co0 += 8) {
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      for (int ky = 0; ky < ks; ++ky)
        for (int kx = 0; kx < ks; ++kx) {
          const float* src = in_s + poff + (ky * d * SW + kx * d) * GS;
          const float* wt = w_s + (long long)(ky * ks + kx) * gin * gout;
          for (int ci = 0; ci < gin; ++ci) {
            const float xv = src[ci];
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (co0 + c < gout) acc[c] = fmaf(xv, wt[ci * gout + co0 + c], acc[c]);
          }
        }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (co0 + c < gout) {
          const float o = acc[c] + b_s[co0 + c];
          dst[co0 + c] = o;
          const float l = lrelu(o);
          s1 += l;
          s2 += l * l;
        }
    }
  }
  s12[0] = s1;
  s12[1] = s2;
}

__global__ void __launch_bounds__(GC_NT) gconv_kernel(const GconvArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[64];
  const int tid = threadIdx.x;
  const int b = blockIdx.y, net = blockIdx.z;
  // select this CTA's branch with scalar copies (static indices keep the params in the constant bank)
  int br_dil = a.br[0].dil, br_groups = a.br[0].groups, br_gin = a.br[0].gin, br_gout = a.br[0].gout;
  int br_out_off = a.br[0].out_off, br_first = 0;
  long long br_w_off = a.br[0].w_off, br_b_off = a.br[0].b_off;
#pragma unroll
  for (int i = 1; i < CNF_MAX_BRANCHES; ++i) {
    if (i < a.n_br && (int)blockIdx.x >= a.br[i].first_item) {
      br_dil = a.br[i].dil; br_groups = a.br[i].groups; br_gin = a.br[i].gin; br_gout = a.br[i].gout;
      br_out_off = a.br[i].out_off; br_first = a.br[i].first_item;
      br_w_off = a.br[i].w_off; br_b_off = a.br[i].b_off;
    }
  }

  const int local = blockIdx.x - br_first;
  const int tiles = a.tiles_y * a.tiles_x;
  const int g = local / tiles, tile = local % tiles;
  const int y0 = (tile / a.tiles_x) * a.TH, x0 = (tile % a.tiles_x) * a.TW;
  const int th = min(a.TH, a.h - y0), tw = min(a.TW, a.w - x0);
  const int halo = br_dil * (a.ks - 1) / 2;
  const int SH = th + 2 * halo, SW = tw + 2 * halo;
  const int gin = br_gin, gout = br_gout;
  const int GS = gc_stride(gin);
  float* in_s = smem;                                   // [SH*SW][GS]
  float* w_s = in_s + (((long long)SH * SW * GS + 3) & ~3);  // [ks*ks][gin][gout]
  float* b_s = w_s + ((a.ks * a.ks * gin * gout + 3) & ~3);

  const float* P = a.params + (long long)net * a.net_stride;
  float mean = 0.f, rstd = 1.f;
  if (a.ln) ln_coeffs(a.stats_in, (long long)net * a.B + b, (double)a.h * a.w * (double)a.Cin, mean, rstd);

  // stage weights + bias of this group
  {
    const float* wsrc = P + br_w_off + (long long)g * a.ks * a.ks * gin * gout;
    for (int i = tid; i < a.ks * a.ks * gin * gout; i += GC_NT) w_s[i] = wsrc[i];
    const float* bsrc = P + br_b_off + g * gout;
    for (int i = tid; i < gout; i += GC_NT) b_s[i] = bsrc[i];
  }
  // stage the LN(LReLU(.)) input tile with its halo; zero padding is applied AFTER the LayerNorm
  {
    const float* src_s = a.in + (long long)net * a.in_net_stride + (long long)b * a.h * a.w * a.Cin;
    const float* gam = P + a.g_off;
    const float* bet = P + a.be_off;
    const int cin0 = g * gin;  // F:402: group j reads channels [j*_d, (j+1)*_d) (cardinality 1: all)
    const int n_stage = SH * SW * gin;
    for (int idx = tid; idx < n_stage; idx += GC_NT) {
      const int ci = idx % gin, pix = idx / gin;
      const int gy = y0 - halo + pix / SW, gx = x0 - halo + pix % SW;
      float v = 0.f;
      if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w) {
        const long long e = ((long long)gy * a.w + gx) * a.Cin + cin0 + ci;
        v = lrelu(src_s[e]);
        if (a.ln) v = (v - mean) * rstd * gam[e] + bet[e];
      }
      in_s[pix * GS + ci] = v;
    }
  }
  __syncthreads();

  float* out_s = a.out + (long long)net * a.out_net_stride + (long long)b * a.h * a.w * a.Cout;
  float s1 = 0.f, s2 = 0.f;
  GcGeom q;
  q.dil = br_dil; q.ks = a.ks; q.gin = gin; q.gout = gout; q.cbase = br_out_off + g * gout; q.w = a.w; q.Cout = a.Cout;
  const int G = (gin == gout) ? gin : 0;
  switch (G) {
    case 1: gconv_compute<1>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
    case 2: gconv_compute<2>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
    case 4: gconv_compute<4>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
    case 8: gconv_compute<8>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
    default: {
      float s12[2];
      gconv_compute_generic(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s12);
      s1 = s12[0];
      s2 = s12[1];
      break;
    }
  }
  if (a.stats_out) {
    double d1, d2;
    block_sum2(s1, s2, red, d1, d2);
    if (tid == 0) {
      double* so = a.stats_out + 2 * ((long long)net * a.B + b);
      atomicAdd(so, d1);
      atomicAdd(so + 1, d2);
    }
  }
}


// ------------------------------------------------------------------------------------------
// 2b. Grouped dilated convs, register-blocked: 128-thread CTAs; the weights of the current tap
//     (G x G) live in registers and are reused by up to 8 pixels per thread (FFMA:LDS.128 = 16:1 at
//     G = 8); the LN(LReLU(.)) tile is staged with 128/64/32-bit vector slots, four slots per thread
//     in flight.  Used when every branch has gin == gout in {1,2,4,8}; other shapes take gconv_kernel.
// ------------------------------------------------------------------------------------------
constexpr int GC2_NT = 128;
constexpr int GC2_PX = 8;

template <int V> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<1> { using T = float; };

template <int V>
__device__ __forceinline__ void gc2_stage(const float* __restrict__ src_s, const float* __restrict__ gam,
                                          const float* __restrict__ bet, float* __restrict__ in_s, int gin, int GS,
                                          int cin0, int Cin, int h, int w, int y0, int x0, int halo, int SH, int SW,
                                          int ln, float mean, float rstd) {
  using T = typename VecT<V>::T;
  constexpr int U = 4;
  const int vpp = gin / V;  // vector slots per pixel
  const int n_slots = SH * SW * vpp;
  for (int base = threadIdx.x; base < n_slots; base += GC2_NT * U) {
    T xv[U], gv[U], bv[U];
    bool ok[U];
    int dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int slot = base + u * GC2_NT;
      ok[u] = false;
      dst[u] = -1;
      if (slot < n_slots) {
        const int pix = slot / vpp, cv = slot - pix * vpp;
        const int sy = pix / SW, sx = pix - sy * SW;
        const int gy = y0 - halo + sy, gx = x0 - halo + sx;
        dst[u] = pix * GS + cv * V;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
          ok[u] = true;
          const long long e = ((long long)gy * w + gx) * Cin + cin0 + cv * V;
          xv[u] = *reinterpret_cast<const T*>(src_s + e);
          if (ln) {
            gv[u] = *reinterpret_cast<const T*>(gam + e);
            bv[u] = *reinterpret_cast<const T*>(bet + e);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (dst[u] < 0) continue;
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.f;
      if (ok[u]) {
        const float* xf = reinterpret_cast<const float*>(&xv[u]);
        const float* gf = reinterpret_cast<const float*>(&gv[u]);
        const float* bf = reinterpret_cast<const float*>(&bv[u]);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float v = lrelu(xf[i]);
          if (ln) v = (v - mean) * rstd * gf[i] + bf[i];
          o[i] = v;
        }
      }
      *reinterpret_cast<T*>(in_s + dst[u]) = *reinterpret_cast<const T*>(o);
    }
  }
}

template <int G>
__device__ __forceinline__ void gc2_compute(const GcGeom q, const float* __restrict__ in_s, const float* __restrict__ w_s,
                                            const float* __restrict__ b_s, int SW, int GS, int th, int tw, int y0,
                                            int x0, float* __restrict__ out_s, float& s1, float& s2) {
  const int tid = threadIdx.x;
  const int TP = th * tw;
  const int npx = (TP + GC2_NT - 1) / GC2_NT;
  const int d = q.dil, ks = q.ks;
  float acc[GC2_PX][G];
  int poff[GC2_PX];
#pragma unroll
  for (int j = 0; j < GC2_PX; ++j) {
    const int p = min(tid + j * GC2_NT, TP - 1);
    poff[j] = ((p / tw) * SW + (p % tw)) * GS;
#pragma unroll
    for (int co = 0; co < G; ++co) acc[j][co] = 0.f;
  }
  for (int ky = 0; ky < ks; ++ky) {
    for (int kx = 0; kx < ks; ++kx) {
      const int toff = (ky * d * SW + kx * d) * GS;
      const float* wt = w_s + (ky * ks + kx) * G * G;
      float wv[G][G];
#pragma unroll
      for (int ci = 0; ci < G; ++ci) {
        if (G % 4 == 0) {
#pragma unroll
          for (int c4 = 0; c4 < G; c4 += 4) {
            const float4 t = ld4(wt + ci * G + c4);
            wv[ci][c4] = t.x; wv[ci][c4 + 1] = t.y; wv[ci][c4 + 2] = t.z; wv[ci][c4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < G; ++c) wv[ci][c] = wt[ci * G + c];
        }
      }
#pragma unroll
      for (int j = 0; j < GC2_PX; ++j) {
        if (j < npx) {
          const float* src = in_s + poff[j] + toff;
          float xv[G];
          if (G % 4 == 0) {
#pragma unroll
            for (int c4 = 0; c4 < G; c4 += 4) {
              const float4 t = ld4(src + c4);
              xv[c4] = t.x; xv[c4 + 1] = t.y; xv[c4 + 2] = t.z; xv[c4 + 3] = t.w;
            }
          } else if (G == 2) {
            const float2 t = *reinterpret_cast<const float2*>(src);
            xv[0] = t.x; xv[G - 1] = t.y;
          } else {
            xv[0] = src[0];
          }
#pragma unroll
          for (int ci = 0; ci < G; ++ci)
#pragma unroll
            for (int co = 0; co < G; ++co) acc[j][co] = fmaf(xv[ci], wv[ci][co], acc[j][co]);
        }
      }
    }
  }
  const int cbase = q.cbase;
  const bool vec = (G % 4 == 0) && (q.Cout % 4 == 0) && (cbase % 4 == 0);
#pragma unroll
  for (int j = 0; j < GC2_PX; ++j) {
    const int p = tid + j * GC2_NT;
    if (j < npx && p < TP) {
      const int y = y0 + p / tw, x = x0 + p % tw;
      float* dst = out_s + ((long long)y * q.w + x) * q.Cout + cbase;
      float o[G];
#pragma unroll
      for (int co = 0; co < G; ++co) {
        o[co] = acc[j][co] + b_s[co];
        const float l = lrelu(o[co]);
        s1 += l;
        s2 += l * l;
      }
      if (vec) {
#pragma unroll
        for (int c4 = 0; c4 < G; c4 += 4) st4(dst + c4, make_float4(o[c4], o[c4 + 1], o[c4 + 2], o[c4 + 3]));
      } else {
#pragma unroll
        for (int co = 0; co < G; ++co) dst[co] = o[co];
      }
    }
  }
}

__global__ void __launch_bounds__(GC2_NT) gconv2_kernel(const GconvArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[64];
  const int tid = threadIdx.x;
  const int b = blockIdx.y, net = blockIdx.z;
  int br_dil = a.br[0].dil, br_gin = a.br[0].gin, br_out_off = a.br[0].out_off, br_first = 0;
  long long br_w_off = a.br[0].w_off, br_b_off = a.br[0].b_off;
#pragma unroll
  for (int i = 1; i < CNF_MAX_BRANCHES; ++i) {
    if (i < a.n_br && (int)blockIdx.x >= a.br[i].first_item) {
      br_dil = a.br[i].dil; br_gin = a.br[i].gin; br_out_off = a.br[i].out_off; br_first = a.br[i].first_item;
      br_w_off = a.br[i].w_off; br_b_off = a.br[i].b_off;
    }
  }
  const int local = blockIdx.x - br_first;
  const int tiles = a.tiles_y * a.tiles_x;
  const int g = local / tiles, tile = local % tiles;
  const int y0 = (tile / a.tiles_x) * a.TH, x0 = (tile % a.tiles_x) * a.TW;
  const int th = min(a.TH, a.h - y0), tw = min(a.TW, a.w - x0);
  const int halo = br_dil * (a.ks - 1) / 2;
  const int SH = th + 2 * halo, SW = tw + 2 * halo;
  const int G = br_gin;
  const int GS = gc_stride(G);
  float* in_s = smem;
  float* w_s = in_s + (((long long)SH * SW * GS + 3) & ~3);
  float* b_s = w_s + ((a.ks * a.ks * G * G + 3) & ~3);

  const float* P = a.params + (long long)net * a.net_stride;
  float mean = 0.f, rstd = 1.f;
  if (a.ln) ln_coeffs(a.stats_in, (long long)net * a.B + b, (double)a.h * a.w * (double)a.Cin, mean, rstd);
  {
    const float* wsrc = P + br_w_off + (long long)g * a.ks * a.ks * G * G;
    for (int i = tid; i < a.ks * a.ks * G * G; i += GC2_NT) w_s[i] = wsrc[i];
    const float* bsrc = P + br_b_off + g * G;
    if (tid < G) b_s[tid] = bsrc[tid];
  }
  const float* src_s = a.in + (long long)net * a.in_net_stride + (long long)b * a.h * a.w * a.Cin;
  const float* gam = P + a.g_off;
  const float* bet = P + a.be_off;
  const int cin0 = g * G;
  const bool al4 = (a.Cin % 4) == 0, al2 = (a.Cin % 2) == 0;
  if (G % 4 == 0 && al4) gc2_stage<4>(src_s, gam, bet, in_s, G, GS, cin0, a.Cin, a.h, a.w, y0, x0, halo, SH, SW, a.ln, mean, rstd);
  else if (G % 2 == 0 && al2) gc2_stage<2>(src_s, gam, bet, in_s, G, GS, cin0, a.Cin, a.h, a.w, y0, x0, halo, SH, SW, a.ln, mean, rstd);
  else gc2_stage<1>(src_s, gam, bet, in_s, G, GS, cin0, a.Cin, a.h, a.w, y0, x0, halo, SH, SW, a.ln, mean, rstd);
  __syncthreads();

  float* out_s = a.out + (long long)net * a.out_net_stride + (long long)b * a.h * a.w * a.Cout;
  float s1 = 0.f, s2 = 0.f;
  GcGeom q;
  q.dil = br_dil; q.ks = a.ks; q.gin = G; q.gout = G; q.cbase = br_out_off + g * G; q.w = a.w; q.Cout = a.Cout;
  switch (G) {
    case 1: gc2_compute<1>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
    case 2: gc2_compute<2>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
    case 4: gc2_compute<4>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
    default: gc2_compute<8>(q, in_s, w_s, b_s, SW, GS, th, tw, y0, x0, out_s, s1, s2); break;
  }
  if (a.stats_out) {
    double d1, d2;
    block_sum2(s1, s2, red, d1, d2);
    if (tid == 0) {
      double* so = a.stats_out + 2 * ((long long)net * a.B + b);
      atomicAdd(so, d1);
      atomicAdd(so + 1, d2);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 2c. Grouped dilated convs, one branch per launch, templated on group width G and pixels per
//     thread PX (no per-pixel branches; out-of-range pixels are clamped and simply not stored).
//     Tap loops are NOT unrolled (the body is ~9 KB of SASS and stays in the instruction cache);
//     staging walks rows and vector slots without integer divisions; the LayerNorm coefficients are
//     computed once per CTA (fp64) and broadcast through shared memory.
// ------------------------------------------------------------------------------------------
template <int G, int PX, int S>
__global__ void __launch_bounds__(320, 2) gconv3_kernel(const GconvArgs a) {
  constexpr int V = G >= 4 ? 4 : G;        // staging vector width (floats)
  constexpr int VPP = G / V;               // vector slots per pixel
  constexpr int GS = G + ((G % 8) == 0 ? 4 : 0);
  using T = typename VecT<V>::T;
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32][2];
  __shared__ float mr[S][2];
  const int tid = threadIdx.x, NT = blockDim.x, NT1 = NT / S;   // NT1: threads per sample (multiple of 32)
  const int lane = tid & 31, wid = tid >> 5, nw = NT >> 5;
  const int b0 = blockIdx.y * S, net = blockIdx.z;
  const int ns = min(S, a.B - b0);
  const GconvBranch& br = a.br[0];
  const int tiles = a.tiles_y * a.tiles_x;
  const int g = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int y0 = (tile / a.tiles_x) * a.TH, x0 = (tile % a.tiles_x) * a.TW;
  const int th = min(a.TH, a.h - y0), tw = min(a.TW, a.w - x0);
  const int d = br.dil;
  const int halo = d;                      // ksize 3
  const int SH = th + 2 * halo, SW = tw + 2 * halo;
  const int in_sz = (SH * SW * GS + 3) & ~3;
  float* in_s = smem;                                        // [S][SH*SW][GS]
  float* w_s = in_s + S * in_sz;                             // [9][G][G]
  float* b_s = w_s + ((9 * G * G + 3) & ~3);

  const float* P = a.params + (long long)net * a.net_stride;
  if (tid < S) {
    float mean = 0.f, rstd = 1.f;
    if (a.ln && tid < ns) ln_coeffs(a.stats_in, (long long)net * a.B + b0 + tid, (double)a.h * a.w * (double)a.Cin, mean, rstd);
    mr[tid][0] = mean;
    mr[tid][1] = rstd;
  }
  {
    const float* wsrc = P + br.w_off + (long long)g * 9 * G * G;
    if (!a.bwd) {
      for (int i = tid; i < 9 * G * G; i += NT) w_s[i] = wsrc[i];
    } else {
      // dA[q, ci] = sum_tap sum_co dOut[q - off(tap), co] W[tap][ci][co]: a forward-form conv with taps flipped
      // and (ci, co) swapped: w'[tap'][co][ci] = W[8 - tap'][ci][co]
      for (int i = tid; i < 9 * G * G; i += NT) {
        const int tap = i / (G * G), r2 = i % (G * G), in_c = r2 / G, out_c = r2 % G;
        w_s[i] = wsrc[((8 - tap) * G + out_c) * G + in_c];
      }
    }
    if (tid < G) b_s[tid] = a.bwd ? 0.f : P[br.b_off + g * G + tid];
  }
  __syncthreads();
  {
    const long long sample_stride = (long long)a.h * a.w * a.Cin;
    const float* src_s = a.in + (long long)net * a.in_net_stride + (long long)b0 * sample_stride;
    const float* gam = P + a.g_off;
    const float* bet = P + a.be_off;
    const int cin0 = br.in_off + g * G;
    const int row_slots = SW * VPP;
    for (int sy = wid; sy < SH; sy += nw) {
      const int gy = y0 - halo + sy;
      const bool rowok = gy >= 0 && gy < a.h;
      for (int sl0 = 0; sl0 < row_slots; sl0 += 64) {       // two slots per lane in flight
        T xv[2][S], gv[2], bv[2];
        bool ok[2];
        int sl[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          sl[u] = sl0 + u * 32 + lane;
          const int sx = sl[u] / VPP, cv = sl[u] % VPP;
          const int gx = x0 - halo + sx;
          ok[u] = rowok && sl[u] < row_slots && gx >= 0 && gx < a.w && !(a.dbg & 1);
          if (ok[u]) {
            const long long e = ((long long)gy * a.w + gx) * a.Cin + cin0 + cv * V;
#pragma unroll
            for (int q = 0; q < S; ++q)
              if (q < ns) xv[u][q] = *reinterpret_cast<const T*>(src_s + q * sample_stride + e);
            if (a.ln) {
              gv[u] = *reinterpret_cast<const T*>(gam + e);
              bv[u] = *reinterpret_cast<const T*>(bet + e);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (sl[u] >= row_slots) continue;
          const int sx = sl[u] / VPP, cv = sl[u] % VPP;
          const float* gf = reinterpret_cast<const float*>(&gv[u]);
          const float* bf = reinterpret_cast<const float*>(&bv[u]);
#pragma unroll
          for (int q = 0; q < S; ++q) {
            float o[V];
#pragma unroll
            for (int i = 0; i < V; ++i) o[i] = 0.f;
            if (ok[u] && q < ns) {
              const float* xf = reinterpret_cast<const float*>(&xv[u][q]);
              const float mean = mr[q][0], rstd = mr[q][1];
#pragma unroll
              for (int i = 0; i < V; ++i) {
                float v = a.bwd ? xf[i] : lrelu(xf[i]);
                if (a.ln) v = (v - mean) * rstd * gf[i] + bf[i];
                o[i] = v;
              }
            }
            *reinterpret_cast<T*>(in_s + q * in_sz + (sy * SW + sx) * GS + cv * V) = *reinterpret_cast<const T*>(o);
          }
        }
      }
    }
  }
  __syncthreads();

  const int sidx = tid / NT1, t1 = tid - sidx * NT1;   // warp-uniform: NT1 is a multiple of 32
  const int TP = th * tw;
  float s1 = 0.f, s2 = 0.f;
  if (sidx < ns) {
    const float* in_q = in_s + sidx * in_sz;
    float acc[PX][G];
    int poff[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      const int p = min(t1 + j * NT1, TP - 1);
      poff[j] = ((p / tw) * SW + (p % tw)) * GS;
#pragma unroll
      for (int co = 0; co < G; ++co) acc[j][co] = 0.f;
    }
#pragma unroll 1
    for (int tap = 0; tap < ((a.dbg & 2) ? 0 : 9); ++tap) {
      const int ky = tap / 3, kx = tap - ky * 3;
      const int toff = (ky * d * SW + kx * d) * GS;
      const float* wt = w_s + tap * G * G;
      float xv[PX][G];
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        const float* src = in_q + poff[j] + toff;
        if (G % 4 == 0) {
#pragma unroll
          for (int c4 = 0; c4 < G; c4 += 4) {
            const float4 t = ld4(src + c4);
            xv[j][c4] = t.x; xv[j][c4 + 1] = t.y; xv[j][c4 + 2] = t.z; xv[j][c4 + 3] = t.w;
          }
        } else if (G == 2) {
          const float2 t = *reinterpret_cast<const float2*>(src);
          xv[j][0] = t.x; xv[j][G - 1] = t.y;
        } else {
          xv[j][0] = src[0];
        }
      }
#pragma unroll
      for (int ci = 0; ci < G; ++ci) {
        float wv[G];
        if (G % 4 == 0) {
#pragma unroll
          for (int c4 = 0; c4 < G; c4 += 4) {
            const float4 t = ld4(wt + ci * G + c4);
            wv[c4] = t.x; wv[c4 + 1] = t.y; wv[c4 + 2] = t.z; wv[c4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < G; ++c) wv[c] = wt[ci * G + c];
        }
#pragma unroll
        for (int j = 0; j < PX; ++j)
#pragma unroll
          for (int co = 0; co < G; ++co) acc[j][co] = fmaf(xv[j][ci], wv[co], acc[j][co]);
      }
    }

    float* out_s = a.out + (long long)net * a.out_net_stride + (long long)(b0 + sidx) * a.h * a.w * a.Cout;
    const int cbase = br.out_off + g * G;
    const bool vec = (G % 4 == 0) && (a.Cout % 4 == 0) && (cbase % 4 == 0);
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      const int p = t1 + j * NT1;
      if (p < TP) {
        const int y = y0 + p / tw, x = x0 + p % tw;
        float* dst = out_s + ((long long)y * a.w + x) * a.Cout + cbase;
        float o[G];
#pragma unroll
        for (int co = 0; co < G; ++co) {
          o[co] = acc[j][co] + b_s[co];
          if (a.bwd) o[co] += dst[co];      // branches overlap on input channels: accumulate (launches are sequential)
          const float l = lrelu(o[co]);
          s1 += l;
          s2 += l * l;
        }
        if (vec) {
#pragma unroll
          for (int c4 = 0; c4 < G; c4 += 4) st4(dst + c4, make_float4(o[c4], o[c4 + 1], o[c4 + 2], o[c4 + 3]));
        } else {
#pragma unroll
          for (int co = 0; co < G; ++co) dst[co] = o[co];
        }
      }
    }
  }
  if (a.stats_out) {
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      red[wid][0] = s1;
      red[wid][1] = s2;
    }
    __syncthreads();
    if (tid < 2 * S && (tid >> 1) < ns) {
      const int q = tid >> 1, which = tid & 1, w1 = NT1 >> 5;
      double t = 0.0;
      for (int i = 0; i < w1; ++i) t += (double)red[q * w1 + i][which];
      atomicAdd(a.stats_out + 2 * ((long long)net * a.B + b0 + q) + which, t);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 3. Head conv (both nets) fused with tanh*w, exp, the affine coupling law, the decompress
//    scatter into the flow buffer and the per-sample log-det (M:1133-1150, M:1198, M:1307-1326,
//    M:1379-1394).  One thread per output pixel, all c2 channels of both nets.
// ------------------------------------------------------------------------------------------
struct HeadArgs {
  const float* in;  // X [2][B][hw][nk]
  long long in_net_stride;
  const float* params;
  long long net_stride, g_off, be_off, w_off, b_off, tanh_off;
  const double* stats_in;
  int B, h, w, nk, c2, ln, ks, mode;
  FlowView view;   // where u2/v2 live (complement mask of the layer)
  int mask_c;
  double* logdet;  // [B], accumulated (forward only)
  float *outA, *outB;  // HEAD_EMIT: [B][hw][c2]
  int TH, CC;
  int paths;  // CNF_PATH_* bits of the layer: kernel families NOT to use
};

constexpr int HD_NT = 256;

template <int C2T>
__global__ void __launch_bounds__(HD_NT) head_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[64];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int y0 = blockIdx.x * a.TH;
  const int th = min(a.TH, a.h - y0);
  const int pad = (a.ks - 1) / 2;
  const int SH = th + 2 * pad, SW = a.w + 2 * pad;
  const int CC = a.CC, CS = CC + 4;
  const int in_sz = (SH * SW * CS + 3) & ~3;
  const int w_sz = a.ks * a.ks * C2T * CC;
  float* in_s = smem;               // [2][SH*SW][CS]
  float* w_s = smem + 2 * in_sz;    // [2][ks*ks][C2T][CC]

  const bool valid = tid < th * a.w;
  const int py = valid ? tid / a.w : 0, px = valid ? tid % a.w : 0;

  __shared__ float mr_s[2][2];
  if (tid < 2) {
    float m_ = 0.f, r_ = 1.f;
    if (a.ln) ln_coeffs(a.stats_in, (long long)tid * a.B + b, (double)a.h * a.w * (double)a.nk, m_, r_);
    mr_s[tid][0] = m_;
    mr_s[tid][1] = r_;
  }
  __syncthreads();
  const float mean[2] = {mr_s[0][0], mr_s[1][0]}, rstd[2] = {mr_s[0][1], mr_s[1][1]};
  float acc[2][C2T];
#pragma unroll
  for (int net = 0; net < 2; ++net)
#pragma unroll
    for (int co = 0; co < C2T; ++co) acc[net][co] = 0.f;

  for (int c0 = 0; c0 < a.nk; c0 += CC) {
    __syncthreads();
    for (int net = 0; net < 2; ++net) {
      const float* P = a.params + (long long)net * a.net_stride;
      const float* src_s = a.in + (long long)net * a.in_net_stride + (long long)b * a.h * a.w * a.nk;
      const float* gam = P + a.g_off;
      const float* bet = P + a.be_off;
      float* dst = in_s + net * in_sz;
      if ((a.nk & 3) == 0) {
        // 128-bit slots, rows walked by warps, two slots per lane in flight (no integer divisions by runtime values
        // other than the quad count, 6 independent loads per lane per step)
        const int lane = tid & 31, wid = tid >> 5;
        const int qpp = CC >> 2;                  // quads per pixel in this chunk
        const int row_slots = SW * qpp;
        for (int sy = wid; sy < SH; sy += HD_NT / 32) {
          const int gy = y0 - pad + sy;
          const bool rowok = gy >= 0 && gy < a.h;
          for (int sl0 = 0; sl0 < row_slots; sl0 += 64) {
            float4 xv[2], gv[2], bv[2];
            bool ok[2];
            int sl[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              sl[u] = sl0 + u * 32 + lane;
              const int sx = sl[u] / qpp, cq = sl[u] - sx * qpp;
              const int gx = sx - pad;
              ok[u] = rowok && sl[u] < row_slots && gx >= 0 && gx < a.w && c0 + 4 * cq < a.nk;
              if (ok[u]) {
                const long long e = ((long long)gy * a.w + gx) * a.nk + c0 + 4 * cq;
                xv[u] = ld4(src_s + e);
                if (a.ln) {
                  gv[u] = ld4(gam + e);
                  bv[u] = ld4(bet + e);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (sl[u] >= row_slots) continue;
              const int sx = sl[u] / qpp, cq = sl[u] - sx * qpp;
              float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok[u]) {
                o.x = lrelu(xv[u].x); o.y = lrelu(xv[u].y); o.z = lrelu(xv[u].z); o.w = lrelu(xv[u].w);
                if (a.ln) {
                  o.x = (o.x - mean[net]) * rstd[net] * gv[u].x + bv[u].x;
                  o.y = (o.y - mean[net]) * rstd[net] * gv[u].y + bv[u].y;
                  o.z = (o.z - mean[net]) * rstd[net] * gv[u].z + bv[u].z;
                  o.w = (o.w - mean[net]) * rstd[net] * gv[u].w + bv[u].w;
                }
              }
              st4(dst + (sy * SW + sx) * CS + 4 * cq, o);
            }
          }
        }
      } else {
        for (int idx = tid; idx < SH * SW * CC; idx += HD_NT) {
          const int ci = idx % CC, pix = idx / CC;
          const int gy = y0 - pad + pix / SW, gx = pix % SW - pad;
          float v = 0.f;
          if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w && c0 + ci < a.nk) {
            const long long e = ((long long)gy * a.w + gx) * a.nk + c0 + ci;
            v = lrelu(src_s[e]);
            if (a.ln) v = (v - mean[net]) * rstd[net] * gam[e] + bet[e];
          }
          dst[pix * CS + ci] = v;
        }
      }
      // weights: w_s[net][tap][co][ci] = W[tap][c0+ci][co]   (HWIO: ((tap*nk)+cin)*c2 + co)
      const float* Wg = P + a.w_off;
      float* wd = w_s + net * w_sz;
      for (int idx = tid; idx < w_sz; idx += HD_NT) {
        const int ci = idx % CC, co = (idx / CC) % C2T, tap = idx / (CC * C2T);
        float v = 0.f;
        if (co < a.c2 && c0 + ci < a.nk) v = Wg[((long long)tap * a.nk + c0 + ci) * a.c2 + co];
        wd[idx] = v;
      }
    }
    __syncthreads();
    if (valid) {
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        const float* src = in_s + net * in_sz;
        const float* wn = w_s + net * w_sz;
        for (int ky = 0; ky < a.ks; ++ky)
          for (int kx = 0; kx < a.ks; ++kx) {
            const float* s = src + ((py + ky) * SW + px + kx) * CS;
            const float* wt = wn + (ky * a.ks + kx) * C2T * CC;
            for (int c4 = 0; c4 < CC; c4 += 4) {
              const float4 xv = ld4(s + c4);
#pragma unroll
              for (int co = 0; co < C2T; ++co) {
                const float4 wv = ld4(wt + co * CC + c4);
                acc[net][co] = fmaf(xv.x, wv.x, acc[net][co]);
                acc[net][co] = fmaf(xv.y, wv.y, acc[net][co]);
                acc[net][co] = fmaf(xv.z, wv.z, acc[net][co]);
                acc[net][co] = fmaf(xv.w, wv.w, acc[net][co]);
              }
            }
          }
      }
    }
  }

  float ld = 0.f;
  if (valid) {
    const float* PA = a.params;
    const float* PB = a.params + a.net_stride;
    const float tw_ = PA[a.tanh_off];
    const int y = y0 + py, x = px;
#pragma unroll
    for (int co = 0; co < C2T; ++co) {
      if (co < a.c2) {
        const float A = tw_ * tanhf(acc[0][co] + PA[a.b_off + co]);  // M:1198, M:114-116
        const float t = acc[1][co] + PB[a.b_off + co];
        if (a.mode == HEAD_EMIT || a.mode == HEAD_EMIT_TANH) {
          const long long e = ((long long)b * a.h * a.w + (long long)y * a.w + x) * a.c2 + co;
          a.outA[e] = a.mode == HEAD_EMIT ? A : tanhf(acc[0][co] + PA[a.b_off + co]);
          a.outB[e] = t;
        } else {
          float* ptr = a.view.base + comp_off(a.view, a.mask_c, b, y, x, co);
          const float u2 = *ptr;
          if (a.mode == HEAD_FWD) {
            *ptr = __fadd_rn(__fmul_rn(expf(A), u2), t);                  // M:1307, M:1230-1231
            ld += A;                                                      // M:1323
            if (a.outA)   // training: keep tanh(raw_A) for the backward pass
              a.outA[((long long)b * a.h * a.w + (long long)y * a.w + x) * a.c2 + co] = tanhf(acc[0][co] + PA[a.b_off + co]);
          } else {
            *ptr = __fmul_rn(__frcp_rn(expf(A)), __fsub_rn(u2, t));       // M:1379, M:1250-1251
          }
        }
      }
    }
  }
  if (a.mode == HEAD_FWD && a.logdet) {
    double d1, d2;
    block_sum2(ld, 0.f, red, d1, d2);
    if (tid == 0) atomicAdd(a.logdet + b, d1);
  }
}

// ------------------------------------------------------------------------------------------
// 1c. Stem conv, dedicated form (M:1106-1111, M:1150-1155): out[net][b][p][:] = bias + sum_k u1c[p + off(k)] W[k][:] with
//     K = 9 c1 (9 .. 108) inputs per pixel.  The generic GEMM staged an im2col tile per CTA with two integer divisions per
//     element; here one CTA owns one sample (both nets): the compressed input (masked gather from the flow buffer) is staged
//     once with a zero halo, both weight matrices live in smem, a lane owns one output-channel quad of PX pixels
//     (32 accumulators for the two nets) and the stores are fully coalesced 128-bit rows.  Write-bound.
// ------------------------------------------------------------------------------------------
template <int PX>
__global__ void __launch_bounds__(256) stem2_kernel(const GemmArgs a) {
  extern __shared__ __align__(16) float st_smem[];
  __shared__ float red[64];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.x;
  const int SW = a.w + 2, SH = a.h + 2;
  const int in_sz = (SH * SW * a.c1 + 3) & ~3;
  float* in_s = st_smem;                     // [SH][SW][c1]
  float* w_s = in_s + in_sz;                 // [2][K][N]
  float* b_s = w_s + 2 * a.K * a.N;          // [2][N]
  for (int idx = tid; idx < SH * SW * a.c1; idx += 256) {
    const int ci = idx % a.c1, pix = idx / a.c1;
    const int iy = pix / SW - 1, ix = pix % SW - 1;
    float v = 0.f;
    if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) v = a.view.base[comp_off(a.view, a.mask, b, iy, ix, ci)];
    in_s[idx] = v;
  }
  for (int net = 0; net < 2; ++net) {
    const float* P = a.params + (long long)net * a.net_stride;
    for (int idx = tid; idx < a.K * a.N; idx += 256) w_s[net * a.K * a.N + idx] = P[a.w_off + idx];
    if (tid < a.N) b_s[net * a.N + tid] = P[a.b_off + tid];
  }
  __syncthreads();
  const int NQ = a.N >> 2, PPW = 32 / NQ;    // lanes per pixel (4, 8, 16), pixel columns per warp
  const int quad = lane % NQ, pcol = lane / NQ;
  const float4 bias0 = ld4(b_s + quad * 4), bias1 = ld4(b_s + a.N + quad * 4);
  float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
  const int per_it = PPW * PX;
  // small batches of large planes: gridDim.y CTAs share a sample (each stages the whole masked input, which is small, and
  // takes every gridDim.y-th group of 8 warp tiles), so that the launch still fills the machine
  for (int it = wid + 8 * blockIdx.y; it * per_it < a.hw; it += 8 * gridDim.y) {
    int pb[PX];
    bool ok[PX];
    float4 acc[2][PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      const int p = it * per_it + j * PPW + pcol;
      ok[j] = p < a.hw;
      const int pc = min(p, a.hw - 1);
      const int y = pc / a.w, x = pc - y * a.w;
      pb[j] = (y * SW + x) * a.c1;
      acc[0][j] = bias0;
      acc[1][j] = bias1;
    }
    int k = 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int toff = (ky * SW + kx) * a.c1;
        for (int ci = 0; ci < a.c1; ++ci, ++k) {
          const float4 w0 = ld4(w_s + k * a.N + quad * 4), w1 = ld4(w_s + (a.K + k) * a.N + quad * 4);
#pragma unroll
          for (int j = 0; j < PX; ++j) {
            const float xv = in_s[pb[j] + toff + ci];
            acc[0][j].x = fmaf(xv, w0.x, acc[0][j].x); acc[0][j].y = fmaf(xv, w0.y, acc[0][j].y);
            acc[0][j].z = fmaf(xv, w0.z, acc[0][j].z); acc[0][j].w = fmaf(xv, w0.w, acc[0][j].w);
            acc[1][j].x = fmaf(xv, w1.x, acc[1][j].x); acc[1][j].y = fmaf(xv, w1.y, acc[1][j].y);
            acc[1][j].z = fmaf(xv, w1.z, acc[1][j].z); acc[1][j].w = fmaf(xv, w1.w, acc[1][j].w);
          }
        }
      }
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      float* out_s = a.out + (long long)net * a.out_net_stride + (long long)b * a.hw * a.N;
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        if (!ok[j]) continue;
        const int p = it * per_it + j * PPW + pcol;
        const float4 o = acc[net][j];
        __stcs(reinterpret_cast<float4*>(out_s + (long long)p * a.N + quad * 4), o);
        float l;
        l = fmaxf(o.x, CNF_LRELU_SLOPE * o.x); s1[net] += l; s2[net] = fmaf(l, l, s2[net]);
        l = fmaxf(o.y, CNF_LRELU_SLOPE * o.y); s1[net] += l; s2[net] = fmaf(l, l, s2[net]);
        l = fmaxf(o.z, CNF_LRELU_SLOPE * o.z); s1[net] += l; s2[net] = fmaf(l, l, s2[net]);
        l = fmaxf(o.w, CNF_LRELU_SLOPE * o.w); s1[net] += l; s2[net] = fmaf(l, l, s2[net]);
      }
    }
  }
  if (a.stats_out) {
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      double d1, d2;
      block_sum2(s1[net], s2[net], red, d1, d2);
      if (tid == 0) {
        double* so = a.stats_out + 2 * ((long long)net * a.B + b);
        atomicAdd(so, d1);
        atomicAdd(so + 1, d2);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// 3b. Head conv for narrow outputs (c2 <= 2: the channel-mask layers), streaming form.  out[p][co] = sum_tap sum_c
//     x[p + off(tap)][c] W[tap][c][co] is evaluated as (i) d[q][tap][co] = sum_c x[q][c] W[tap][c][co] for every pixel q
//     -- every activation row is read from HBM exactly once, fully coalesced, LReLU+LayerNorm in registers, no halo
//     staging of the nk-channel tensor -- and (ii) out[p] = sum_tap d[p + off(tap)][tap] from shared memory, followed by
//     the same tanh*w / exp / coupling law / scatter / log-det epilogue as head_kernel.  One CTA per sample, both nets.
//     A lane owns one channel quad of a pixel (W of its quad for the 9 taps lives in registers); the nk/4 lanes of a pixel
//     are reduced with shuffles.  (A variant with one net per pass and 3-4 pixel groups in flight spilled at 128 registers
//     and was slower: 120 vs 86 us.)
// ------------------------------------------------------------------------------------------
template <int C2T>
__global__ void __launch_bounds__(256, C2T == 1 ? 2 : 1) head2_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float h2_smem[];
  __shared__ float red[64];
  __shared__ float mr_s[2][2];
  constexpr int NV = 9 * C2T;                    // partial sums per pixel and net
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.x;
  const int hw = a.h * a.w;
  const int LPP = a.nk >> 2, PPW = 32 / LPP;     // lanes per pixel (4, 8 or 16), pixels per warp step
  const int cq = lane % LPP, pl = lane / LPP;
  float* d_s = h2_smem;                          // [2][hw][NV]

  if (tid < 2) {
    float m_ = 0.f, r_ = 1.f;
    if (a.ln) ln_coeffs(a.stats_in, (long long)tid * a.B + b, (double)hw * (double)a.nk, m_, r_);
    mr_s[tid][0] = r_;
    mr_s[tid][1] = -m_ * r_;
  }
  // this lane's weights: wr[net][tap * C2T + co][i] = W[tap][cq * 4 + i][co]
  float4 wr[2][NV];
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    const float* Wg = a.params + (long long)net * a.net_stride + a.w_off;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int tap = v / C2T, co = v - tap * C2T;
      float w4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w4[i] = co < a.c2 ? Wg[((long long)tap * a.nk + cq * 4 + i) * a.c2 + co] : 0.f;
      wr[net][v] = make_float4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
  __syncthreads();
  const float sc[2] = {mr_s[0][0], mr_s[1][0]}, sh[2] = {mr_s[0][1], mr_s[1][1]};
  const long long sample = (long long)b * hw * a.nk;
  const int nsteps = (hw + PPW - 1) / PPW;
  for (int g = wid; g < nsteps; g += 8) {
    const int p = g * PPW + pl;
    const bool ok = p < hw;
    const long long e = (long long)min(p, hw - 1) * a.nk + cq * 4;
    float4 x[2], gm[2], be[2];
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      x[net] = __ldcs(reinterpret_cast<const float4*>(a.in + (long long)net * a.in_net_stride + sample + e));
      if (a.ln) {
        const float* P = a.params + (long long)net * a.net_stride;
        gm[net] = ld4(P + a.g_off + e);
        be[net] = ld4(P + a.be_off + e);
      }
    }
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      float4 v = x[net];
      v.x = fmaxf(v.x, CNF_LRELU_SLOPE * v.x); v.y = fmaxf(v.y, CNF_LRELU_SLOPE * v.y);
      v.z = fmaxf(v.z, CNF_LRELU_SLOPE * v.z); v.w = fmaxf(v.w, CNF_LRELU_SLOPE * v.w);
      if (a.ln) {
        v.x = fmaf(fmaf(v.x, sc[net], sh[net]), gm[net].x, be[net].x);
        v.y = fmaf(fmaf(v.y, sc[net], sh[net]), gm[net].y, be[net].y);
        v.z = fmaf(fmaf(v.z, sc[net], sh[net]), gm[net].z, be[net].z);
        v.w = fmaf(fmaf(v.w, sc[net], sh[net]), gm[net].w, be[net].w);
      }
      float part[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k)
        part[k] = fmaf(v.x, wr[net][k].x, fmaf(v.y, wr[net][k].y, fmaf(v.z, wr[net][k].z, v.w * wr[net][k].w)));
      for (int m = LPP >> 1; m > 0; m >>= 1) {
#pragma unroll
        for (int k = 0; k < NV; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], m);
      }
      if (cq == 0 && ok) {
        float* dd = d_s + ((long long)net * hw + p) * NV;
#pragma unroll
        for (int k = 0; k < NV; ++k) dd[k] = part[k];
      }
    }
  }
  __syncthreads();

  float ld = 0.f;
  const float* PA = a.params;
  const float* PB = a.params + a.net_stride;
  const float tw_ = PA[a.tanh_off];
  for (int p = tid; p < hw; p += 256) {
    const int y = p / a.w, x = p - y * a.w;
    float acc[2][C2T];
#pragma unroll
    for (int net = 0; net < 2; ++net)
#pragma unroll
      for (int co = 0; co < C2T; ++co) acc[net][co] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y + ky - 1;
      if (iy < 0 || iy >= a.h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x + kx - 1;
        if (ix < 0 || ix >= a.w) continue;
        const int q = iy * a.w + ix;
#pragma unroll
        for (int net = 0; net < 2; ++net)
#pragma unroll
          for (int co = 0; co < C2T; ++co) acc[net][co] += d_s[((long long)net * hw + q) * NV + (ky * 3 + kx) * C2T + co];
      }
    }
#pragma unroll
    for (int co = 0; co < C2T; ++co) {
      if (co < a.c2) {
        const float raw = acc[0][co] + PA[a.b_off + co];
        const float A = tw_ * tanhf(raw);                                // M:1198, M:114-116
        const float t = acc[1][co] + PB[a.b_off + co];
        if (a.mode == HEAD_EMIT || a.mode == HEAD_EMIT_TANH) {
          const long long e = ((long long)b * hw + p) * a.c2 + co;
          a.outA[e] = a.mode == HEAD_EMIT ? A : tanhf(raw);
          a.outB[e] = t;
        } else {
          float* ptr = a.view.base + comp_off(a.view, a.mask_c, b, y, x, co);
          const float u2 = *ptr;
          if (a.mode == HEAD_FWD) {
            *ptr = __fadd_rn(__fmul_rn(expf(A), u2), t);                  // M:1307, M:1230-1231
            ld += A;                                                      // M:1323
            if (a.outA) a.outA[((long long)b * hw + p) * a.c2 + co] = tanhf(raw);   // training: tanh(raw_A) for the backward pass
          } else {
            *ptr = __fmul_rn(__frcp_rn(expf(A)), __fsub_rn(u2, t));       // M:1379, M:1250-1251
          }
        }
      }
    }
  }
  if (a.mode == HEAD_FWD && a.logdet) {
    double d1, d2;
    block_sum2(ld, 0.f, red, d1, d2);
    if (tid == 0) atomicAdd(a.logdet + b, d1);
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
#define CU_TRY(x)                      \
  do {                                 \
    cudaError_t e_ = (x);              \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

template <int TN, int NQ, int RM, bool STEM>
static int launch_gemm_t(const GemmArgs& a, cudaStream_t st) {
  constexpr int CT = TN / (4 * NQ), RT = 128 / CT, TM = RM * RT;
  const size_t smem = ((size_t)TM * (a.KC + 4) + (size_t)a.KC * TN) * sizeof(float);
  auto kern = gemm_kernel<TN, NQ, RM, STEM>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  const int tiles_m = (a.hw + TM - 1) / TM, tiles_n = (a.N + TN - 1) / TN;
  dim3 grid(tiles_m * tiles_n, a.B, 2);
  kern<<<grid, 128, smem, st>>>(a);
  return (int)cudaGetLastError();
}

static int launch_stem2(const GemmArgs& a, cudaStream_t st) {
  const size_t smem = ((((size_t)(a.h + 2) * (a.w + 2) * a.c1 + 3) & ~(size_t)3) + 2 * (size_t)a.K * a.N + 2 * a.N) * sizeof(float);
  auto kern = stem2_kernel<4>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  // large planes come with small batches (64x64: 64, 128x128: 16 per GPU): one CTA per 1024 pixels.  The split depends on
  // the plane only, never on the batch, so a sample's statistics are reduced the same way whatever batch it is part of.
  const int warp_px = (32 / (a.N >> 2)) * 4;                       // pixels of one warp tile
  const int rounds = (a.hw + 8 * warp_px - 1) / (8 * warp_px);     // groups of 8 warp tiles per sample
  const int split = std::max(1, std::min(rounds, a.hw / 1024));
  kern<<<dim3(a.B, split), 256, smem, st>>>(a);
  return (int)cudaGetLastError();
}

template <bool STEM>
static int launch_gemm(GemmArgs a, cudaStream_t st) {
  if (STEM) {
    const bool s2 = !(a.paths & CNF_PATH_NO_STEM2);
    const size_t smem = ((size_t)(a.h + 2) * (a.w + 2) * a.c1 + 4 + 2 * (size_t)a.K * a.N + 2 * a.N) * sizeof(float);
    if (s2 && a.ks == 3 && a.K == 9 * a.c1 && (a.N == 16 || a.N == 32 || a.N == 64) && !a.w_trans && !a.no_bias && !a.res &&
        smem <= 200 * 1024)
      return launch_stem2(a, st);
  }
  const int Kp = (a.K + 3) & ~3;
  a.KC = std::min(Kp, 128);
  if (a.N > 32) return launch_gemm_t<64, 2, 8, STEM>(a, st);
  if (a.N > 16) return launch_gemm_t<32, 1, 8, STEM>(a, st);
  return launch_gemm_t<16, 1, 4, STEM>(a, st);
}

template <int TN, int NQ>
static int launch_pw_t(GemmArgs a, cudaStream_t st) {
  constexpr int CT = TN / (4 * NQ), PT = 256 / CT, S = 8;
  const int nchunks = (a.K + 63) / 64;
  a.KC = (((a.K + nchunks - 1) / nchunks) + 3) & ~3;
  const size_t smem = ((size_t)S * PT * (a.KC + 4) + (size_t)a.KC * TN) * sizeof(float);
  auto kern = pw_kernel<TN, NQ>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  const int tiles_p = (a.hw + PT - 1) / PT, tiles_n = (a.N + TN - 1) / TN;
  dim3 grid(tiles_p * tiles_n, (a.B + S - 1) / S, 2);
  kern<<<grid, 256, smem, st>>>(a);
  return (int)cudaGetLastError();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link); nullptr if unavailable
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      return (tmap_encode_fn)ptr;
    return (tmap_encode_fn) nullptr;
  }();
  return fn;
}

// Tensor maps of the TMA producer of pw_tc3_kernel (tc_kernels.cuh, Tc3Maps): false when a pointer / stride does not meet
// the 16-byte rules of a tensor map (the caller then uses the cp.async producer).
static bool tc3_make_maps(const GemmArgs& a, Tc3Maps* m) {
  tmap_encode_fn fn = tmap_encoder();
  if (!fn) return false;
  auto ok16 = [](long long bytes) { return bytes > 0 && (bytes % 16) == 0; };
  const long long row = (long long)a.K * 4, img = (long long)a.hw * a.K * 4;
  if (!ok16(row) || ((uintptr_t)a.in & 15) || !ok16(a.in_net_stride * 4)) return false;
  const cuuint32_t one[4] = {1, 1, 1, 1};
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)a.K, (cuuint64_t)a.hw, (cuuint64_t)a.B, 2};
    const cuuint64_t gstr[3] = {(cuuint64_t)row, (cuuint64_t)img, (cuuint64_t)a.in_net_stride * 4};
    const cuuint32_t box[4] = {32, 32, 4, 1};
    if (fn(&m->x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.in, gdim, gstr, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  if (a.ln) {
    const float* g = a.params + a.g_off;
    const float* b = a.params + a.be_off;
    if (((uintptr_t)g & 15) || ((uintptr_t)b & 15) || !ok16(a.net_stride * 4)) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)a.K, (cuuint64_t)a.hw, 2};
    const cuuint64_t gstr[2] = {(cuuint64_t)row, (cuuint64_t)a.net_stride * 4};
    const cuuint32_t box[3] = {32, 32, 1};
    for (int i = 0; i < 2; ++i)
      if (fn(i ? &m->b : &m->g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)(i ? b : g), gdim, gstr, box, one,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
  }
  return true;
}

// warp-specialised persistent tcgen05 1x1 conv; CNF_NOT_ELIGIBLE when the resident-W image does not fit (or, TMA, when the
// tensors do not meet the tensor-map alignment rules)
template <int N, int TW, int NST, bool PADN, bool ATM, int KCH = 32, bool TMA = false>
static int launch_pw_tc3_tw(const GemmArgs& a, cudaStream_t st) {
  const int nchunks = (a.K + KCH - 1) / KCH;
  const size_t ops = ATM ? 0 : (size_t)NST * 2 * 128 * 32;           // operand stages: shared memory, or tensor memory (ATM)
  const size_t depth = (ATM && KCH == 32) ? 5 : 3;
  const size_t smem = (ops + depth * 6 * (32 * KCH / 4) * 4 + (size_t)nchunks * 2 * N * KCH) * sizeof(float);
  if (smem > 225 * 1024) return CNF_NOT_ELIGIBLE;
  Tc3Maps maps{};
  if (TMA && !tc3_make_maps(a, &maps)) return CNF_NOT_ELIGIBLE;
  auto kern = pw_tc3_kernel<N, TW, NST, PADN, ATM, KCH, TMA>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  int n_sm = 0;
  CU_TRY((cudaError_t)device_sm_count(&n_sm));
  const int tiles_p = (a.hw + 31) / 32, tiles_s = (a.B + 3) / 4;
  const int grid = std::min(2 * tiles_p * tiles_s, n_sm & ~1);
  GemmArgs b = a;
  { static int dbg = -1; if (dbg < 0) dbg = knob_int("PW_DBG", 0); b.dbg = dbg; }
  kern<<<grid, (TW + 1 + 8 + (TMA ? 1 : 4)) * 32, smem, st>>>(b, tiles_p, tiles_s, maps);
  return (int)cudaGetLastError();
}

// N <= 64: the A operand stages live in tensor memory (3 stages of 64 columns beside the two accumulators) and the raw ring is
// 5 deep.  N = 128 (zero-padded data gradients): operand stages in shared memory, as many as fit beside the resident weights.
template <int N, bool PADN>
static int launch_pw_tc3_p(const GemmArgs& a, cudaStream_t st) {
  if constexpr (N <= 64) {
    static int atm = -1;
    if (atm < 0) atm = knob_int("PW_ATM", 1);
    if (atm) {
      // 32-channel chunks, 3 TMEM stages of 64 columns.  (64-channel chunks with 2 stages of 128 columns, KCH = 64, measured
      // slower: 77.7 / 133.0 us against 70.8 / 112.7 us, profiles/r02_summary.md)
      static int tma = -1;
      if (tma < 0) tma = knob_int("PW_TMA", 1);
      if (tma) {     // raw ring filled by bulk tensor copies (one producer thread)
        const int rc = launch_pw_tc3_tw<N, 8, 3, PADN, true, 32, true>(a, st);
        if (rc != CNF_NOT_ELIGIBLE) return rc;
      }
      const int rc = launch_pw_tc3_tw<N, 8, 3, PADN, true>(a, st);
      if (rc != CNF_NOT_ELIGIBLE) return rc;
    }
  }
  static int nst = -1;
  if (nst < 0) nst = knob_int("PW_NST", 4);
  int rc = CNF_NOT_ELIGIBLE;
  if (nst >= 4) rc = launch_pw_tc3_tw<N, 8, 4, PADN, false>(a, st);
  if (rc == CNF_NOT_ELIGIBLE && nst >= 3) rc = launch_pw_tc3_tw<N, 8, 3, PADN, false>(a, st);
  if (rc == CNF_NOT_ELIGIBLE) rc = launch_pw_tc3_tw<N, 8, 2, PADN, false>(a, st);
  return rc;
}

template <int N>
static int launch_pw_tc3_t(const GemmArgs& a, cudaStream_t st) {
  return a.N == N ? launch_pw_tc3_p<N, false>(a, st) : launch_pw_tc3_p<N, true>(a, st);
}

// tcgen05 3xTF32 kernel on the smallest UMMA tile (N = 16, 32, 64, 128 columns) that holds the a.N outputs (the extra
// columns are zero weights and are not stored); CNF_NOT_ELIGIBLE when the shape is outside the tile family
static int launch_pw_tc3(const GemmArgs& a, cudaStream_t st) {
  if (a.K % 4 || a.N % 8 || a.N < 8 || a.N > 128) return CNF_NOT_ELIGIBLE;   // K % 8 == 4: the last K-step is zero-padded
  if (a.N <= 16) return a.N == 16 ? launch_pw_tc3_t<16>(a, st) : CNF_NOT_ELIGIBLE;   // N = 8: no 8-column epilogue split
  if (a.N <= 32) return launch_pw_tc3_t<32>(a, st);
  if (a.N <= 64) return launch_pw_tc3_t<64>(a, st);
  return launch_pw_tc3_t<128>(a, st);
}

// 1x1 conv dispatcher: tcgen05 3xTF32 kernel when the shape fits its UMMA tile family (N <= 128, K % 8 == 0, resident
// weights fit shared memory); otherwise (or when a.paths excludes it) the FFMA multi-sample kernel (K % 4 == 0), then
// the generic GEMM kernel.
static int launch_pw(const GemmArgs& a, cudaStream_t st) {
  if (!(a.paths & CNF_PATH_NO_TCGEN05)) {
    const int rc = launch_pw_tc3(a, st);
    if (rc != CNF_NOT_ELIGIBLE) return rc;
  }
  if (!(a.paths & CNF_PATH_NO_PW_FFMA) && a.K % 4 == 0 && a.B <= 65535 * 8) {
    if (a.N > 32) return launch_pw_t<64, 2>(a, st);
    return launch_pw_t<32, 1>(a, st);
  }
  return launch_gemm<false>(a, st);
}

template <int G, int PX, int S>
static int launch_gconv3_t(const GconvArgs& a, int NT, size_t smem, cudaStream_t st) {
  auto kern = gconv3_kernel<G, PX, S>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  dim3 grid(a.br[0].groups * a.tiles_y * a.tiles_x, (a.B + S - 1) / S, 2);
  kern<<<grid, NT * S, smem, st>>>(a);
  return (int)cudaGetLastError();
}

template <int G, int S>
static int launch_gconv3_g(const GconvArgs& a, int PX, int NT, size_t smem, cudaStream_t st) {
  switch (PX) {
    case 1: return launch_gconv3_t<G, 1, S>(a, NT, smem, st);
    case 2: return launch_gconv3_t<G, 2, S>(a, NT, smem, st);
    case 3: return launch_gconv3_t<G, 3, S>(a, NT, smem, st);
    case 4: return launch_gconv3_t<G, 4, S>(a, NT, smem, st);
    case 5: return launch_gconv3_t<G, 5, S>(a, NT, smem, st);
    case 6: return launch_gconv3_t<G, 6, S>(a, NT, smem, st);
    case 7: return launch_gconv3_t<G, 7, S>(a, NT, smem, st);
    default: return launch_gconv3_t<G, 8, S>(a, NT, smem, st);
  }
}

// One branch with gin == gout in {1,2,4,8}, ksize 3.  CNF_NOT_ELIGIBLE for other shapes.
static int launch_gconv3_branch(const GconvArgs& g, int bi, cudaStream_t st) {
  const GconvBranch& br = g.br[bi];
  const int G = br.gin;
  if (g.ks != 3 || br.gin != br.gout || !(G == 1 || G == 2 || G == 4 || G == 8) || g.B > 65535) return CNF_NOT_ELIGIBLE;
  if ((G >= 4 && g.Cin % 4) || (G == 2 && g.Cin % 2)) return CNF_NOT_ELIGIBLE;
  GconvArgs a = g;
  a.n_br = 1;
  a.br[0] = br;
  a.br[0].first_item = 0;
  a.dbg = 0;
  a.TH = std::min(a.h, 32);
  a.TW = std::min(a.w, 32);
  a.tiles_y = (a.h + a.TH - 1) / a.TH;
  a.tiles_x = (a.w + a.TW - 1) / a.TW;
  const int TP = a.TH * a.TW;
  // pixels per thread: thread slots nt*px (idle slots still issue), each tap re-reads the group's weights once per
  // thread, so the per-pixel cost falls like (1 + alpha/px); 32-thread samples are allowed (two samples share a CTA)
  static float alpha = -1.f;
  if (alpha < 0.f) alpha = knob_float("GC_ALPHA", 2.0f);
  int best_px = 8, best_nt = 128;
  float best_cost = 1e30f;
  for (int px = 8; px >= 1; --px) {
    int nt = ((TP + px - 1) / px + 31) / 32 * 32;
    if (nt > 256) continue;
    const float cost = (float)nt * px * (1.f + alpha / px);
    if (cost < best_cost) { best_cost = cost; best_px = px; best_nt = nt; }
  }
  const int halo = br.dil;
  const int GS = G + ((G % 8) == 0 ? 4 : 0);
  const size_t in_sz = (((size_t)(a.TH + 2 * halo) * (a.TW + 2 * halo) * GS) + 3) & ~(size_t)3;
  const size_t tail = ((9 * G * G + 3) & ~3) + G + 4;
  // two samples per CTA share gamma/beta loads and double the warps that walk the staging rows
  static int s_env = -1;
  if (s_env < 0) s_env = knob_int("GC_S", 2);
  const bool two = s_env >= 2 && a.B >= 2 && (2 * in_sz + tail) * sizeof(float) <= 110 * 1024 && 2 * best_nt <= 320;
  if (!two && best_nt < 64) best_nt = 64;
  const size_t smem = ((two ? 2 : 1) * in_sz + tail) * sizeof(float);
  if (smem > 227 * 1024) return CNF_NOT_ELIGIBLE;
  if (two) {
    switch (G) {
      case 1: return launch_gconv3_g<1, 2>(a, best_px, best_nt, smem, st);
      case 2: return launch_gconv3_g<2, 2>(a, best_px, best_nt, smem, st);
      case 4: return launch_gconv3_g<4, 2>(a, best_px, best_nt, smem, st);
      default: return launch_gconv3_g<8, 2>(a, best_px, best_nt, smem, st);
    }
  }
  switch (G) {
    case 1: return launch_gconv3_g<1, 1>(a, best_px, best_nt, smem, st);
    case 2: return launch_gconv3_g<2, 1>(a, best_px, best_nt, smem, st);
    case 4: return launch_gconv3_g<4, 1>(a, best_px, best_nt, smem, st);
    default: return launch_gconv3_g<8, 1>(a, best_px, best_nt, smem, st);
  }
}

static int launch_gconv_ffma(GconvArgs a, cudaStream_t st);

}  // namespace cnf
#include "gconv_oct.cuh"
#include "gconv_tc.cuh"
namespace cnf {

static int launch_gconv(GconvArgs a, cudaStream_t st) {
  if (!(a.paths & CNF_PATH_NO_OCTET)) {
    // all dilation branches in one persistent launch (gconv_oct.cuh); shapes it does not cover fall through
    const int rc = launch_gconv_oct(a, st);
    if (rc != CNF_NOT_ELIGIBLE) return rc;
  }
  GconvArgs rest = a;
  rest.n_br = 0;
  for (int i = 0; i < a.n_br; ++i) {
    // groups of 16 / 32 channels: implicit GEMM on the tensor cores (gconv_tc.cuh); narrower groups: FFMA2 per-branch kernel
    int rc = (a.paths & CNF_PATH_NO_TCGEN05) ? CNF_NOT_ELIGIBLE : launch_gconv_tc_branch(a, i, st);
    if (rc == CNF_NOT_ELIGIBLE) rc = (a.paths & CNF_PATH_NO_BRANCH) ? CNF_NOT_ELIGIBLE : launch_gconv3_branch(a, i, st);
    if (rc == CNF_NOT_ELIGIBLE) rest.br[rest.n_br++] = a.br[i];
    else if (rc != 0) return rc;
  }
  if (rest.n_br) return launch_gconv_ffma(rest, st);
  return 0;
}

static int launch_gconv_ffma(GconvArgs a, cudaStream_t st) {
  bool v2 = !(a.paths & CNF_PATH_NO_GCONV2);
  for (int i = 0; i < a.n_br; ++i) {
    const int g = a.br[i].gin;
    if (a.br[i].gin != a.br[i].gout || !(g == 1 || g == 2 || g == 4 || g == 8)) v2 = false;
  }
  // whole plane when it fits the pixels-per-thread budget, else 32x32 tiles
  a.TH = std::min(a.h, 32);
  a.TW = std::min(a.w, 32);
  a.tiles_y = (a.h + a.TH - 1) / a.TH;
  a.tiles_x = (a.w + a.TW - 1) / a.TW;
  const int tiles = a.tiles_y * a.tiles_x;
  int items = 0;
  size_t smem = 0;
  for (int i = 0; i < a.n_br; ++i) {
    a.br[i].first_item = items;
    items += a.br[i].groups * tiles;
    const int halo = a.br[i].dil * (a.ks - 1) / 2;
    const size_t in_sz = (((size_t)(a.TH + 2 * halo) * (a.TW + 2 * halo) * gc_stride(a.br[i].gin)) + 3) & ~(size_t)3;
    const size_t w_sz = ((size_t)a.ks * a.ks * a.br[i].gin * a.br[i].gout + 3) & ~(size_t)3;
    smem = std::max(smem, (in_sz + w_sz + a.br[i].gout + 4) * sizeof(float));
  }
  if (smem > 227 * 1024) return (int)cudaErrorInvalidConfiguration;
  dim3 grid(items, a.B, 2);
  if (v2) {
    static SmemAttrCache cache2;
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)gconv2_kernel, smem, cache2));
    gconv2_kernel<<<grid, GC2_NT, smem, st>>>(a);
    return (int)cudaGetLastError();
  }
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)gconv_kernel, smem, cache));
  gconv_kernel<<<grid, GC_NT, smem, st>>>(a);
  return (int)cudaGetLastError();
}

template <int C2T>
static int launch_head_t(const HeadArgs& a, cudaStream_t st) {
  const int pad = (a.ks - 1) / 2;
  const size_t in_sz = (((size_t)(a.TH + 2 * pad) * (a.w + 2 * pad) * (a.CC + 4)) + 3) & ~(size_t)3;
  const size_t smem = (2 * in_sz + 2 * (size_t)a.ks * a.ks * C2T * a.CC) * sizeof(float);
  if (smem > 227 * 1024) return (int)cudaErrorInvalidConfiguration;
  auto kern = head_kernel<C2T>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  dim3 grid((a.h + a.TH - 1) / a.TH, a.B);
  kern<<<grid, HD_NT, smem, st>>>(a);
  return (int)cudaGetLastError();
}

template <int C2T>
static int launch_head2_t(const HeadArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)2 * a.h * a.w * 9 * C2T * sizeof(float);
  auto kern = head2_kernel<C2T>;
  static SmemAttrCache cache;
  CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)kern, smem, cache));
  kern<<<a.B, 256, smem, st>>>(a);
  return (int)cudaGetLastError();
}

static int launch_head(HeadArgs a, cudaStream_t st) {
  // narrow heads (channel-mask layers): streaming kernel, every activation row read once
  const bool h2 = !(a.paths & CNF_PATH_NO_HEAD2);
  if (h2 && a.ks == 3 && a.c2 <= 2 && (a.nk == 16 || a.nk == 32 || a.nk == 64) &&
      (size_t)2 * a.h * a.w * 9 * (a.c2 <= 1 ? 1 : 2) * sizeof(float) <= 100 * 1024)
    return a.c2 <= 1 ? launch_head2_t<1>(a, st) : launch_head2_t<2>(a, st);
  if (a.w > HD_NT) return (int)cudaErrorInvalidConfiguration;
  const int n_tiles = (a.h * a.w + HD_NT - 1) / HD_NT;
  a.TH = (a.h + n_tiles - 1) / n_tiles;
  while (a.TH * a.w > HD_NT) --a.TH;
  a.CC = std::min(16, (a.nk + 3) & ~3);
  if (a.c2 <= 1) return launch_head_t<1>(a, st);
  if (a.c2 <= 2) return launch_head_t<2>(a, st);
  if (a.c2 <= 4) return launch_head_t<4>(a, st);
  if (a.c2 <= 8) return launch_head_t<8>(a, st);
  if (a.c2 <= 16) return launch_head_t<16>(a, st);
  if (a.c2 <= 32) return launch_head_t<32>(a, st);
  return (int)cudaErrorInvalidConfiguration;
}

// One coupling layer on `B` samples.  in_view/in_mask: where u1 is gathered from; out_view: where
// u2 is read and v2 written (the same buffer for the in-place flow).  mode: HeadMode.
int run_coupling(const cnf_coupling* c, const float* params, FlowView in_view, int in_mask, FlowView out_view,
                 int B, int mode, double* logdet_acc, float* outA, float* outB, void* ws, void* stream,
                 const CouplingSaved* sv, int extra_excluded_paths) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0) return 0;
  const int paths = c->paths | extra_excluded_paths;
  if (!sv && !(paths & CNF_PATH_NO_RESIDENT) && (mode == HEAD_FWD || mode == HEAD_INV)) {
    const int rc = launch_fused_coupling(c, params, in_view, in_mask, out_view, B, mode, logdet_acc, ws, stream);
    if (rc != CNF_NOT_ELIGIBLE) return rc;
  }
  CouplingWorkspace W = {};
  if (!sv) W = carve_ws(c, B, ws);
  else W.stats = sv->stats;
  const int hw = c->hw(), nk = c->nk, cat = c->cat;
  const int n_ln = c->n_ln();
  const long long slot = 2LL * B * 2;  // doubles per stats slot
  if (n_ln) CU_TRY(cudaMemsetAsync(W.stats, 0, sizeof(double) * slot * n_ln, st));
  auto stats = [&](int i) -> double* { return n_ln ? W.stats + slot * i : nullptr; };
  // inference: one residual-stream buffer updated in place; training (sv): every stage keeps its own buffer
  auto Xb = [&](int r) -> float* { return sv ? sv->X[r] : W.X; };
  auto Y1b = [&](int r) -> float* { return sv ? sv->Y1[r] : W.Y1; };
  auto Y2b = [&](int r) -> float* { return sv ? sv->Y2[r] : W.Y2; };
  if (sv && mode == HEAD_FWD) outA = sv->TH;   // tanh(raw_A) for the backward pass

  {  // stem
    GemmArgs a = {};
    a.view = in_view; a.mask = in_mask; a.h = c->h; a.w = c->w; a.c1 = c->c1; a.ks = c->ks;
    a.params = params; a.net_stride = c->net_stride; a.w_off = c->stem_w; a.b_off = c->stem_b;
    a.stats_out = stats(0);
    a.out = Xb(0); a.out_net_stride = (long long)B * hw * nk;
    a.B = B; a.hw = hw; a.K = c->ks * c->ks * c->c1; a.N = nk; a.ln = 0; a.paths = paths;
    CU_TRY((cudaError_t)launch_gemm<true>(a, st));
  }
  for (int r = 0; r < c->R; ++r) {
    const ResBlockLayout& L = c->rb[r];
    {  // pw1: X -> Y1
      GemmArgs a = {};
      a.in = Xb(r); a.in_net_stride = (long long)B * hw * nk;
      a.params = params; a.net_stride = c->net_stride; a.w_off = L.pw1_w; a.b_off = L.pw1_b;
      a.g_off = L.ln1_g; a.be_off = L.ln1_b;
      a.stats_in = stats(3 * r); a.stats_out = stats(3 * r + 1);
      a.out = Y1b(r); a.out_net_stride = (long long)B * hw * nk;
      a.B = B; a.hw = hw; a.K = nk; a.N = nk; a.ln = c->ln; a.paths = paths;
      CU_TRY((cudaError_t)launch_pw(a, st));
    }
    {  // grouped dilated convs: Y1 -> Y2
      GconvArgs a = {};
      a.in = Y1b(r); a.in_net_stride = (long long)B * hw * nk; a.Cin = nk;
      a.out = Y2b(r); a.out_net_stride = (long long)B * hw * cat; a.Cout = cat;
      a.params = params; a.net_stride = c->net_stride; a.g_off = L.ln2_g; a.be_off = L.ln2_b;
      a.stats_in = stats(3 * r + 1); a.stats_out = stats(3 * r + 2);
      a.B = B; a.h = c->h; a.w = c->w; a.ln = c->ln; a.ks = c->ks; a.paths = paths;
      a.n_br = (int)L.br.size();
      for (int i = 0; i < a.n_br; ++i) {
        const Branch& s = L.br[i];
        a.br[i].dil = s.dil; a.br[i].groups = s.groups; a.br[i].gin = s.gin; a.br[i].gout = s.gout;
        a.br[i].out_off = s.out_off; a.br[i].w_off = s.w_off; a.br[i].b_off = s.b_off;
      }
      CU_TRY((cudaError_t)launch_gconv(a, st));
    }
    {  // pw2 + residual: Y2 (+X) -> X
      GemmArgs a = {};
      a.in = Y2b(r); a.in_net_stride = (long long)B * hw * cat;
      a.params = params; a.net_stride = c->net_stride; a.w_off = L.pw2_w; a.b_off = L.pw2_b;
      a.g_off = L.ln3_g; a.be_off = L.ln3_b;
      a.stats_in = stats(3 * r + 2); a.stats_out = stats(3 * r + 3);
      a.out = Xb(r + 1); a.res = Xb(r); a.out_net_stride = (long long)B * hw * nk;
      a.B = B; a.hw = hw; a.K = cat; a.N = nk; a.ln = c->ln; a.paths = paths;
      CU_TRY((cudaError_t)launch_pw(a, st));
    }
  }
  {  // head + coupling
    HeadArgs a = {};
    a.in = Xb(c->R); a.in_net_stride = (long long)B * hw * nk;
    a.params = params; a.net_stride = c->net_stride; a.g_off = c->lnf_g; a.be_off = c->lnf_b;
    a.w_off = c->head_w; a.b_off = c->head_b; a.tanh_off = c->tanh_w;
    a.stats_in = stats(3 * c->R);
    a.B = B; a.h = c->h; a.w = c->w; a.nk = nk; a.c2 = c->c2; a.ln = c->ln; a.ks = c->ks; a.mode = mode; a.paths = paths;
    a.view = out_view; a.mask_c = c->mask_c;
    a.logdet = logdet_acc; a.outA = outA; a.outB = outB;
    CU_TRY((cudaError_t)launch_head(a, st));
  }
  return 0;
}

// (sum, sum of squares) of LReLU(x) per sample and net: the LayerNorm statistics a producing kernel would have left
// behind (stand-alone residual block entry point).  x is [2][B][n]; stats [2][B][2] (overwritten).
__global__ void __launch_bounds__(256) lrelu_stats_kernel(const float* __restrict__ x, long long n, int B, double* __restrict__ stats) {
  __shared__ float red[64];
  const int b = blockIdx.x, net = blockIdx.y;
  const float* p = x + ((long long)net * B + b) * n;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float l = lrelu(p[i]);
    s1 += l;
    s2 = fmaf(l, l, s2);
  }
  double d1, d2;
  block_sum2(s1, s2, red, d1, d2);
  if (threadIdx.x == 0) {
    stats[2 * ((long long)net * B + b)] = d1;
    stats[2 * ((long long)net * B + b) + 1] = d2;
  }
}

// One dilated residual block (F:501-627) of both nets on stand-alone activations: Xin, Xout are [2][B][hw][nk]
// (net A first; Xout may alias Xin).  The same three launches run_coupling issues for block r.
int run_residual_block(const cnf_coupling* c, const float* params, int r, const float* Xin, float* Xout, int B, void* ws,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0) return 0;
  if (r < 0 || r >= c->R) return (int)cudaErrorInvalidValue;
  CouplingWorkspace W = carve_ws(c, B, ws);
  const int hw = c->hw(), nk = c->nk, cat = c->cat;
  const long long slot = 2LL * B * 2;
  const int n_ln = c->n_ln();
  auto stats = [&](int i) -> double* { return n_ln ? W.stats + slot * i : nullptr; };
  if (n_ln) {
    CU_TRY(cudaMemsetAsync(W.stats, 0, sizeof(double) * slot * 4, st));
    lrelu_stats_kernel<<<dim3(B, 2), 256, 0, st>>>(Xin, (long long)hw * nk, B, stats(0));
    CU_TRY(cudaGetLastError());
  }
  const ResBlockLayout& L = c->rb[r];
  {
    GemmArgs a = {};
    a.in = Xin; a.in_net_stride = (long long)B * hw * nk;
    a.params = params; a.net_stride = c->net_stride; a.w_off = L.pw1_w; a.b_off = L.pw1_b;
    a.g_off = L.ln1_g; a.be_off = L.ln1_b;
    a.stats_in = stats(0); a.stats_out = stats(1);
    a.out = W.Y1; a.out_net_stride = (long long)B * hw * nk;
    a.B = B; a.hw = hw; a.K = nk; a.N = nk; a.ln = c->ln; a.paths = c->paths;
    CU_TRY((cudaError_t)launch_pw(a, st));
  }
  {
    GconvArgs a = {};
    a.in = W.Y1; a.in_net_stride = (long long)B * hw * nk; a.Cin = nk;
    a.out = W.Y2; a.out_net_stride = (long long)B * hw * cat; a.Cout = cat;
    a.params = params; a.net_stride = c->net_stride; a.g_off = L.ln2_g; a.be_off = L.ln2_b;
    a.stats_in = stats(1); a.stats_out = stats(2);
    a.B = B; a.h = c->h; a.w = c->w; a.ln = c->ln; a.ks = c->ks; a.paths = c->paths;
    a.n_br = (int)L.br.size();
    for (int i = 0; i < a.n_br; ++i) {
      const Branch& s = L.br[i];
      a.br[i].dil = s.dil; a.br[i].groups = s.groups; a.br[i].gin = s.gin; a.br[i].gout = s.gout;
      a.br[i].out_off = s.out_off; a.br[i].w_off = s.w_off; a.br[i].b_off = s.b_off;
    }
    CU_TRY((cudaError_t)launch_gconv(a, st));
  }
  {
    GemmArgs a = {};
    a.in = W.Y2; a.in_net_stride = (long long)B * hw * cat;
    a.params = params; a.net_stride = c->net_stride; a.w_off = L.pw2_w; a.b_off = L.pw2_b;
    a.g_off = L.ln3_g; a.be_off = L.ln3_b;
    a.stats_in = stats(2); a.stats_out = stats(3);
    a.out = Xout; a.res = Xin; a.out_net_stride = (long long)B * hw * nk;
    a.B = B; a.hw = hw; a.K = cat; a.N = nk; a.ln = c->ln; a.paths = c->paths;
    CU_TRY((cudaError_t)launch_pw(a, st));
  }
  return 0;
}

int read_tc3_clocks(long long* out, int n) {
  CU_TRY(cudaDeviceSynchronize());
  return (int)cudaMemcpyFromSymbol(out, g_tc3_clk, sizeof(long long) * std::min(n, 8192));
}

// Measurement hook (cnf_debug_pw_conv): one 1x1-conv launch of residual block 0 on the current workspace.
int run_pw_only(const cnf_coupling* c, const float* params, int B, int which, void* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (c->R < 1) return (int)cudaErrorInvalidValue;
  CouplingWorkspace W = carve_ws(c, B, ws);
  const int hw = c->hw(), nk = c->nk, cat = c->cat;
  const long long slot = 2LL * B * 2;
  const ResBlockLayout& L = c->rb[0];
  GemmArgs a = {};
  a.params = params; a.net_stride = c->net_stride;
  a.B = B; a.hw = hw; a.N = nk; a.ln = c->ln; a.paths = c->paths;
  a.out_net_stride = (long long)B * hw * nk;
  // the statistics outputs go to the spare slot (index n_ln) so that repeated launches do not disturb the layer
  double* spare = c->n_ln() ? W.stats + slot * c->n_ln() : nullptr;
  if (spare) CU_TRY(cudaMemsetAsync(spare, 0, sizeof(double) * slot, st));
  if (which == 2) {  // the grouped dilated convs of block 0: Y1 -> Y2 (Y2 is simply rewritten with the same values)
    GconvArgs g = {};
    g.in = W.Y1; g.in_net_stride = (long long)B * hw * nk; g.Cin = nk;
    g.out = W.Y2; g.out_net_stride = (long long)B * hw * cat; g.Cout = cat;
    g.params = params; g.net_stride = c->net_stride; g.g_off = L.ln2_g; g.be_off = L.ln2_b;
    g.stats_in = c->n_ln() ? W.stats + slot : nullptr; g.stats_out = spare;
    g.B = B; g.h = c->h; g.w = c->w; g.ln = c->ln; g.ks = c->ks; g.paths = c->paths;
    g.n_br = (int)L.br.size();
    for (int i = 0; i < g.n_br; ++i) {
      const Branch& s = L.br[i];
      g.br[i].dil = s.dil; g.br[i].groups = s.groups; g.br[i].gin = s.gin; g.br[i].gout = s.gout;
      g.br[i].out_off = s.out_off; g.br[i].w_off = s.w_off; g.br[i].b_off = s.b_off;
    }
    return launch_gconv(g, st);
  }
  if (which == 0) {
    a.in = W.X; a.in_net_stride = (long long)B * hw * nk; a.K = nk;
    a.w_off = L.pw1_w; a.b_off = L.pw1_b; a.g_off = L.ln1_g; a.be_off = L.ln1_b;
    a.stats_in = c->n_ln() ? W.stats : nullptr; a.stats_out = spare;
    a.out = W.Y1;
  } else {
    a.in = W.Y2; a.in_net_stride = (long long)B * hw * cat; a.K = cat;
    a.w_off = L.pw2_w; a.b_off = L.pw2_b; a.g_off = L.ln3_g; a.be_off = L.ln3_b;
    a.stats_in = c->n_ln() ? W.stats + slot * 2 : nullptr; a.stats_out = spare;
    a.out = W.Y1; a.res = W.X;     // out-of-place so that X is not accumulated into
  }
  return launch_pw(a, st);
}


// ---- backward-pass uses of the forward kernels (data gradients), called from bwd_kernels.cu ----
// 1x1 conv: dA[2][B][hw][K] = dY[2][B][hw][N] * W^T   (W is [K][N] at w_off)
int dgrad_pw(const float* params, long long net_stride, long long w_off, const float* dY, float* dA, int B, int hw,
             int K, int N, void* stream, int paths) {
  GemmArgs a = {};
  a.in = dY; a.in_net_stride = (long long)B * hw * N;
  a.params = params; a.net_stride = net_stride; a.w_off = w_off;
  a.out = dA; a.out_net_stride = (long long)B * hw * K;
  a.B = B; a.hw = hw; a.K = N; a.N = K; a.ln = 0;
  a.raw_in = 1; a.w_trans = 1; a.ldw = N; a.no_bias = 1; a.paths = paths;
  if (!(paths & CNF_PATH_NO_TCGEN05)) {   // the forward tensor-core kernel with the weights read transposed
    const int rc = launch_pw_tc3(a, (cudaStream_t)stream);
    if (rc != CNF_NOT_ELIGIBLE) return rc;
  }
  return launch_gemm<false>(a, (cudaStream_t)stream);
}

// grouped dilated convs: dA[2][B][hw][nk] = sum over branches of conv^T(dY[2][B][hw][cat]).  Branches are handled by the
// tensor-core kernel (groups of 16 / 32), else by gconv3 (groups of 1 / 2 / 4 / 8); bit i of *leftover is set for a branch
// neither covers (the caller adds it with the generic kernel).
int dgrad_gconv(const cnf_coupling* c, int r, const float* params, const float* dY, float* dA, int B, void* stream,
                unsigned* leftover) {
  cudaStream_t st = (cudaStream_t)stream;
  const ResBlockLayout& L = c->rb[r];
  const int hw = c->hw(), nk = c->nk, cat = c->cat;
  *leftover = 0;
  CU_TRY(cudaMemsetAsync(dA, 0, sizeof(float) * 2 * (size_t)B * hw * nk, st));
  GconvArgs a = {};
  a.in = dY; a.in_net_stride = (long long)B * hw * cat; a.Cin = cat;
  a.out = dA; a.out_net_stride = (long long)B * hw * nk; a.Cout = nk;
  a.params = params; a.net_stride = c->net_stride;
  a.B = B; a.h = c->h; a.w = c->w; a.ln = 0; a.ks = c->ks; a.bwd = 1;
  a.n_br = (int)L.br.size();
  for (int i = 0; i < a.n_br; ++i) {
    const Branch& s = L.br[i];
    a.br[i].dil = s.dil; a.br[i].groups = s.groups; a.br[i].gin = s.gin; a.br[i].gout = s.gout;
    a.br[i].in_off = s.out_off; a.br[i].out_off = 0; a.br[i].w_off = s.w_off; a.br[i].b_off = s.b_off;
  }
  for (int i = 0; i < a.n_br; ++i) {
    int rc = (c->paths & CNF_PATH_NO_TCGEN05) ? CNF_NOT_ELIGIBLE : launch_gconv_tc_branch(a, i, st);
    if (rc == CNF_NOT_ELIGIBLE) rc = launch_gconv3_branch(a, i, st);
    if (rc == CNF_NOT_ELIGIBLE) *leftover |= 1u << i;
    else if (rc) return rc;
  }
  return 0;
}

}  // namespace cnf
