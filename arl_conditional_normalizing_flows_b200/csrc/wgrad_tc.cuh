// 1x1-conv weight gradient on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only:
//   dW[k][n] = sum_{b,p} a[b,p,k] dy[b,p,n],  db[n] = sum_{b,p} dy[b,p,n],  a = LN(lrelu(x)) applied on load.
// Reference: the tf.GradientTape block of cFlow.train_step (conv_cINN_make_model.py M:1863-1871) applied to the Conv2D 1x1
// layers of dilated_residual_block (conv_cINN_base_functions.py F:561-575, F:612-625).
//
// The reduction runs over the B * hw pixels (200 k at config 2), so the GEMM is  D[K x N] = A^T[K x pixels] dY[pixels x N]
// with the PIXELS as the UMMA K dimension.  Both operands are K-major in shared memory, i.e. transposed with respect to
// the [pixel][channel] tensors in HBM: a thread loads a 4-pixel x 4-channel block (four 128-bit loads), applies LReLU +
// LayerNorm, splits hi / lo (3xTF32, as in tc_kernels.cuh) and stores four 128-bit rows "4 pixels of one channel" -- the
// transpose happens in registers.  A stage is 32 pixels of one sample: A^T is 128 rows (channels, zero-extended: rows
// >= K produce accumulator rows nobody reads) x 32, dY^T is [hi rows | lo rows] x 32.  Per 8-pixel K-step one thread issues
// hi(A) x [hi(dY) | lo(dY)] (width 2 NB) and lo(A) x hi(dY) (width NB).  The stage loop is software-pipelined over two
// buffers (the MMAs of stage i overlap the loads + transform of stage i + 1); the 128 x NB accumulator stays in TMEM for
// the whole CTA and is written ONCE, to a per-CTA partial; wgrad_tc_reduce_kernel adds the partials in a fixed order, so the
// weight gradients are bit-reproducible (the FFMA kernel it replaces used fp32 atomics).
#pragma once

#include "tc_common.cuh"

namespace cnf {

constexpr int WGT_PX = 32;                 // pixels per stage
constexpr int WGT_WT = 8;                  // worker warps
constexpr int WGT_NT = (WGT_WT + 1) * 32;  // + the MMA-issuing warp
constexpr int WGT_MAX_NC = 160;            // CTAs per net (the partial buffer is sized for them)

struct WgtArgs {
  const float* x;             // [2][B][hw][K]
  const float* dy;            // [2][B][hw][N]
  long long x_net_stride, dy_net_stride;
  const float* params;
  long long net_stride, g_off, be_off;
  const double* stats;
  float* partial;             // [2][nc][K * N + N]
  int B, hw, K, N, ln;
  int sps, n_stages, nc;      // stages per sample, stages per net, CTAs per net
};

template <int NB>              // dY channels of the UMMA tile: 16, 32 or 64 (N <= NB)
__global__ void __launch_bounds__(WGT_NT, 2) wgrad_tc_kernel(const WgtArgs a) {
  constexpr int KQ8 = WGT_PX / 4;              // 16-byte K chunks (4 pixels) per stage = 8
  constexpr int A_ST = 128 * WGT_PX;           // floats of one A^T image (hi or lo)
  constexpr int B_ST = NB * WGT_PX;
  constexpr int BUF = 2 * A_ST + 2 * B_ST;     // floats per buffer: A hi, A lo, dY hi, dY lo
  constexpr uint32_t TMEM_COLS = 2 * NB < 32 ? 32 : 2 * NB;
  extern __shared__ __align__(128) float wgt_smem[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[NB];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int net = blockIdx.y, rank = blockIdx.x;
  const float* P = a.params + (long long)net * a.net_stride;
  const float* xs_n = a.x + (long long)net * a.x_net_stride;
  const float* dy_n = a.dy + (long long)net * a.dy_net_stride;
  const float* gam = P + a.g_off;
  const float* bet = P + a.be_off;
  const int KQ = a.K >> 2, NQ = a.N >> 2;
  const double n_ln = (double)a.hw * (double)a.K;

  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid < NB) bias_s[tid] = 0.f;

  // K-major unit (16 bytes = 4 pixels) of row m, pixel quad kq: (m % 8) + 8 kq + 8 KQ8 (m / 8)
  auto unit = [](int m, int kq) { return (m & 7) + 8 * kq + 8 * KQ8 * (m >> 3); };
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);   // this thread's share of db (its dY block has a fixed channel quad)

  // One A^T block (channel quad c4, pixel quad rq) and at most one dY^T block per thread (K <= 128, N <= 64: at most 256 / 128
  // blocks per stage), so ALL global loads of a stage -- 4 x, 4 gamma, 4 beta, 4 dY, the sample's two LayerNorm sums -- are issued
  // before the first use; a thread's dY channel quad is the same in every stage (its share of db stays in registers).
  const bool has_a = tid < KQ * KQ8, has_d = tid < NQ * KQ8;
  const int c4 = has_a ? tid % KQ : 0, rq = has_a ? tid / KQ : 0;
  const int n4 = has_d ? tid % NQ : 0, rqd = has_d ? tid / NQ : 0;
  auto transform = [&](int st, int buf) {
    const int b = st / a.sps, p0 = (st - b * a.sps) * WGT_PX;
    const int np = min(WGT_PX, a.hw - p0);
    float* A_hi = wgt_smem + buf * BUF;
    float* A_lo = A_hi + A_ST;
    float* B_hi = A_lo + A_ST;
    float* B_lo = B_hi + B_ST;
    const float* xs = xs_n + ((long long)b * a.hw + p0) * a.K;
    const float* ds = dy_n + ((long long)b * a.hw + p0) * a.N;
    const float* gs = gam + (long long)p0 * a.K;
    const float* bs = bet + (long long)p0 * a.K;
    float4 v[4], g[4], be[4], dv[4];
    if (has_a) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = min(rq * 4 + j, np - 1);
        const long long e = (long long)p * a.K + 4 * c4;
        v[j] = ld4(xs + e);
        if (a.ln) { g[j] = ld4(gs + e); be[j] = ld4(bs + e); }
      }
    }
    if (has_d) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dv[j] = ld4(ds + (long long)min(rqd * 4 + j, np - 1) * a.N + 4 * n4);
    }
    float mean = 0.f, rstd = 1.f;
    if (a.ln) ln_coeffs(a.stats, (long long)net * a.B + b, n_ln, mean, rstd);
    const float sc = rstd, sh = -mean * rstd;
    if (has_a) {
      float t[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 w = v[j];
        w.x = fmaxf(w.x, CNF_LRELU_SLOPE * w.x); w.y = fmaxf(w.y, CNF_LRELU_SLOPE * w.y);
        w.z = fmaxf(w.z, CNF_LRELU_SLOPE * w.z); w.w = fmaxf(w.w, CNF_LRELU_SLOPE * w.w);
        if (a.ln) {
          w.x = fmaf(fmaf(w.x, sc, sh), g[j].x, be[j].x);
          w.y = fmaf(fmaf(w.y, sc, sh), g[j].y, be[j].y);
          w.z = fmaf(fmaf(w.z, sc, sh), g[j].z, be[j].z);
          w.w = fmaf(fmaf(w.w, sc, sh), g[j].w, be[j].w);
        }
        if (rq * 4 + j >= np) w = make_float4(0.f, 0.f, 0.f, 0.f);   // beyond the sample: contributes nothing
        t[j][0] = w.x; t[j][1] = w.y; t[j][2] = w.z; t[j][3] = w.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {          // channel 4 c4 + i: its 4 pixels as one 16-byte K chunk
        float4 hi, lo;
        tf32_split(t[0][i], hi.x, lo.x); tf32_split(t[1][i], hi.y, lo.y);
        tf32_split(t[2][i], hi.z, lo.z); tf32_split(t[3][i], hi.w, lo.w);
        const int u = unit(4 * c4 + i, rq);
        st4(A_hi + 4 * u, hi);
        st4(A_lo + 4 * u, lo);
      }
    }
    if (has_d) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (rqd * 4 + j >= np) dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        bsum.x += dv[j].x; bsum.y += dv[j].y; bsum.z += dv[j].z; bsum.w += dv[j].w;
      }
      const float t[4][4] = {{dv[0].x, dv[0].y, dv[0].z, dv[0].w}, {dv[1].x, dv[1].y, dv[1].z, dv[1].w},
                             {dv[2].x, dv[2].y, dv[2].z, dv[2].w}, {dv[3].x, dv[3].y, dv[3].z, dv[3].w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 hi, lo;
        tf32_split(t[0][i], hi.x, lo.x); tf32_split(t[1][i], hi.y, lo.y);
        tf32_split(t[2][i], hi.z, lo.z); tf32_split(t[3][i], hi.w, lo.w);
        const int u = unit(4 * n4 + i, rqd);
        st4(B_hi + 4 * u, hi);
        st4(B_lo + 4 * u, lo);
      }
    }
  };

  // rows of A^T beyond K and of dY^T beyond N are never written; they only feed accumulator rows / columns nobody reads
  // (rows and columns of a GEMM are independent), but the tensor core should not be handed undefined bits: zero once
  for (int i = tid; i < 2 * BUF; i += WGT_NT) wgt_smem[i] = 0.f;
  __syncthreads();

  int st = rank, buf = 0;
  const bool any = st < a.n_stages;
  if (any && wid < WGT_WT) transform(st, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  uint32_t phase = 0;
  bool first = true;

  for (; st < a.n_stages; st += a.nc, buf ^= 1) {
    if (wid == WGT_WT) {
      if (elect_one()) {
        tc_fence_after();
        constexpr uint32_t idesc2 = umma_idesc_tf32(2 * NB), idesc1 = umma_idesc_tf32(NB);
        constexpr uint32_t LBO = 128, SBO = KQ8 * 128;
        const uint32_t a_hi = smem_u32(wgt_smem + buf * BUF), a_lo = a_hi + A_ST * 4, b_hi = a_lo + A_ST * 4;
        const uint64_t dah0 = umma_desc(a_hi, LBO, SBO), dal0 = umma_desc(a_lo, LBO, SBO), dbh0 = umma_desc(b_hi, LBO, SBO);
#pragma unroll
        for (int ks = 0; ks < WGT_PX / 8; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 2 * LBO) >> 4);
          umma_tf32(tmem_d, dah0 + adv, dbh0 + adv, idesc2, first ? 0u : 1u);
          umma_tf32(tmem_d, dal0 + adv, dbh0 + adv, idesc1, 1u);
          first = false;
        }
        umma_commit(&bar_mma);
      }
    } else {
      const int nx = st + a.nc;
      if (nx < a.n_stages) transform(nx, buf ^ 1);     // overlaps the MMAs of `st`
      fence_async_smem();
      mbar_wait(&bar_mma, phase);                       // the MMAs of `st` are done: its buffer may be refilled next time
    }
    phase ^= 1;
    __syncthreads();
  }

  // ---- epilogue: accumulator row = channel k (TMEM lane), columns [0, NB) + [NB, 2 NB) -> this CTA's partial
  float* part = a.partial + ((long long)net * a.nc + rank) * ((long long)a.K * a.N + a.N);
  if (wid < WGT_WT) {
    tc_fence_after();
    constexpr int NH = NB / 2;
    const int quarter = wid & 3, half = wid >> 2;
    const int c0 = half * NH;
    const int k = quarter * 32 + lane;
    const uint32_t t0 = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
#pragma unroll
    for (int cb = 0; cb < NH; cb += 8) {
      float v0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, v1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (any) {
        tmem_ld<8>(t0 + cb, v0);
        tmem_ld<8>(t0 + NB + cb, v1);
      }
      if (k < a.K) {
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const int n = c0 + cb + i;
          if (n < a.N) st4(part + (long long)k * a.N + n, make_float4(v0[i] + v1[i], v0[i + 1] + v1[i + 1], v0[i + 2] + v1[i + 2], v0[i + 3] + v1[i + 3]));
        }
      }
    }
    tc_fence_before();
    // db: the eight pixel-quad threads of a channel quad add up in shared memory (fixed set of addends; fp32 atomics on 8
    // values: order-dependent in the last bit only within this CTA's partial, the cross-CTA sum is ordered)
    if (has_d) {
      atomicAdd(&bias_s[4 * n4], bsum.x); atomicAdd(&bias_s[4 * n4 + 1], bsum.y);
      atomicAdd(&bias_s[4 * n4 + 2], bsum.z); atomicAdd(&bias_s[4 * n4 + 3], bsum.w);
    }
  }
  __syncthreads();
  if (tid < a.N) part[(long long)a.K * a.N + tid] = bias_s[tid];
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// grads[w_off + i] += sum over the CTAs' partials in CTA order (i < K N), grads[b_off + n] likewise
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float* __restrict__ partial, int nc, int KN, int N,
                                                              float* __restrict__ grads, long long net_stride, long long w_off,
                                                              long long b_off) {
  const int net = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= KN + N) return;
  const float* p = partial + (long long)net * nc * (KN + N) + i;
  float s = 0.f;
  for (int c = 0; c < nc; ++c) s += p[(long long)c * (KN + N)];
  float* g = grads + (long long)net * net_stride + (i < KN ? w_off + i : b_off + (i - KN));
  *g += s;        // this launch is the only writer of these entries; the stream orders it after earlier contributions
}

// bytes of per-CTA partials for the 1x1 convs of a layer with nk kernels and `cat` concatenated channels (0 if not eligible)
inline int64_t wgrad_tc_scratch_bytes(int nk, int cat, int64_t B, int hw) {
  const int K = cat > nk ? cat : nk;
  if (nk > 64 || nk % 4 || K > 128) return 0;
  const int64_t stages = B * ((hw + WGT_PX - 1) / WGT_PX);
  const int64_t nc = stages < WGT_MAX_NC ? stages : WGT_MAX_NC;
  return 2 * nc * ((int64_t)K * nk + nk) * (int64_t)sizeof(float);
}

}  // namespace cnf
