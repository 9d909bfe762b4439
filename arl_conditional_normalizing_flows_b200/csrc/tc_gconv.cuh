// Grouped dilated 3x3 convolutions on tcgen05 (sm_100a): implicit GEMM by shifted shared-memory descriptors.
//
// One CTA owns 16 consecutive channels of one dilation branch (16/G groups) for a band of TH image rows of
// one (sample, net).  The LN(LReLU(.)) input band with its halo is staged ONCE, split hi/lo (3xTF32, see
// tc_kernels.cuh), in a layout that is planar by channel quad:
//       unit(16 B) = q + Qp * cq,     q = linear index in the padded band (pitch SW), cq = channel quad 0..3
// This is a legal K-major no-swizzle UMMA operand with SBO = 128 B (8-row groups are contiguous, i.e. row m
// sits at m*16 B) and LBO = Qp*16 B, so the A operand of tap (ky,kx) is the SAME buffer with the start
// address advanced by (ky*d*SW + kx*d)*16 B: no im2col copy, 9 taps = 9 descriptor offsets.
// Rows of the accumulator are padded-pitch pixels (the 2*halo garbage columns per row are never stored).
// B is the per-tap 16x16 block-diagonal weight matrix of the 16/G groups (zero off the diagonal blocks).
// Per 128-pixel M-tile: 9 taps x 2 K-steps x 3 split terms = 54 tcgen05.mma (M=128, N=16, K=8).
#pragma once

#include "tc_kernels.cuh"

namespace cnf {

struct GcTcArgs {
  const float* in;     // [2][B][hw][Cin]
  float* out;          // [2][B][hw][Cout]
  long long in_net_stride, out_net_stride;
  const float* params;
  long long net_stride, g_off, be_off;
  const double* stats_in;
  double* stats_out;
  int B, h, w, Cin, Cout, ln;
  int TH, n_bands, Qp_max;
  int dbg;   // debug: bit0 skip staging loads, bit1 skip MMAs, bit2 skip epilogue stores
  int n_br;
  GconvBranch br[CNF_MAX_BRANCHES];   // first_item counts 16-channel slabs x bands
};

constexpr int GCT_NT = 256;
constexpr int GCT_MT = 4;   // M-tiles (128 padded-pitch pixels each) per CTA

__global__ void __launch_bounds__(GCT_NT, 2) gconv_tc_kernel(const GcTcArgs a) {
  extern __shared__ __align__(128) float gct_smem[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[GCT_NT / 32][2];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, net = blockIdx.z;
  int dil = a.br[0].dil, G = a.br[0].gin, out_off = a.br[0].out_off, first = 0;
  long long w_off = a.br[0].w_off, b_off = a.br[0].b_off;
#pragma unroll
  for (int i = 1; i < CNF_MAX_BRANCHES; ++i) {
    if (i < a.n_br && (int)blockIdx.x >= a.br[i].first_item) {
      dil = a.br[i].dil; G = a.br[i].gin; out_off = a.br[i].out_off; first = a.br[i].first_item;
      w_off = a.br[i].w_off; b_off = a.br[i].b_off;
    }
  }
  const int local = blockIdx.x - first;
  const int slab = local / a.n_bands, band = local % a.n_bands;   // 16-channel slab of the branch, row band
  const int c0 = slab * 16;                   // first channel of the slab inside the branch (input == output index)
  const int y0 = band * a.TH;
  const int th = min(a.TH, a.h - y0);
  const int halo = dil;                       // ksize 3: dilation * (3 - 1) / 2
  const int SW = a.w + 2 * halo, SH = th + 2 * halo;
  const int Q = SH * SW;
  const int Qp = a.Qp_max;                    // plane pitch in 16-B units (== 2 mod 8: conflict-free staging)

  float* A_hi = gct_smem;                     // [4 quads][Qp][4]
  float* A_lo = A_hi + 16 * Qp;
  float* B_hi = A_lo + 16 * Qp;               // [9 taps][16 n][16 k] canonical K-major
  float* B_lo = B_hi + 9 * 256;

  const float* P = a.params + (long long)net * a.net_stride;
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) tmem_alloc(&tmem_slot, 64);
  float mean = 0.f, rstd = 1.f;
  if (a.ln) ln_coeffs(a.stats_in, (long long)net * a.B + b, (double)a.h * a.w * (double)a.Cin, mean, rstd);

  // ---- B: block-diagonal per-tap weights.  Packed kernels are [group][tap][ci][co] (G x G each).
  {
    const float* Wg = P + w_off;
    const int g0 = c0 / G;
    for (int idx = tid; idx < 9 * 256; idx += GCT_NT) {
      const int tap = idx >> 8, n = (idx >> 4) & 15, k = idx & 15;
      float wv = 0.f;
      if (n / G == k / G) wv = Wg[(((long long)(g0 + n / G) * 9 + tap) * G + (k % G)) * G + (n % G)];
      float hh, ll;
      tf32_split(wv, hh, ll);
      const int off = tap * 256 + 4 * ((n & 7) + 8 * (k >> 2) + 32 * (n >> 3)) + (k & 3);
      B_hi[off] = hh;
      B_lo[off] = ll;
    }
  }
  // ---- A: stage LN(LReLU(x)) of the band + halo, zero outside the image, hi/lo split, quad-planar
  {
    const float* src_s = a.in + (long long)net * a.in_net_stride + (long long)b * a.h * a.w * a.Cin;
    const float* gam = P + a.g_off;
    const float* bet = P + a.be_off;
    const int cq = lane & 3;
    constexpr int U = 4;
    const int n_pix = Q;
    for (int base = (tid >> 2); base < n_pix; base += (GCT_NT / 4) * U) {
      float4 xv[U], gv[U], bv[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = base + u * (GCT_NT / 4);
        ok[u] = false;
        if (q < n_pix) {
          const int sy = q / SW, sx = q - sy * SW;
          const int gy = y0 - halo + sy, gx = sx - halo;
          if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w && !(a.dbg & 1)) {
            ok[u] = true;
            const long long e = ((long long)gy * a.w + gx) * a.Cin + c0 + 4 * cq;
            xv[u] = ld4(src_s + e);
            if (a.ln) {
              gv[u] = ld4(gam + e);
              bv[u] = ld4(bet + e);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = base + u * (GCT_NT / 4);
        if (q >= n_pix) continue;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok[u]) {
          v[0] = lrelu(xv[u].x); v[1] = lrelu(xv[u].y); v[2] = lrelu(xv[u].z); v[3] = lrelu(xv[u].w);
          if (a.ln) {
            v[0] = (v[0] - mean) * rstd * gv[u].x + bv[u].x;
            v[1] = (v[1] - mean) * rstd * gv[u].y + bv[u].y;
            v[2] = (v[2] - mean) * rstd * gv[u].z + bv[u].z;
            v[3] = (v[3] - mean) * rstd * gv[u].w + bv[u].w;
          }
        }
        float hh[4], ll[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) tf32_split(v[i], hh[i], ll[i]);
        const int unit = q + Qp * cq;
        st4(A_hi + 4 * unit, make_float4(hh[0], hh[1], hh[2], hh[3]));
        st4(A_lo + 4 * unit, make_float4(ll[0], ll[1], ll[2], ll[3]));
      }
    }
    // rows read past Q by the last M-tile / largest shift must be finite: zero the tail of every plane
    const int tail0 = Q, tail1 = Qp;
    for (int idx = tid; idx < 4 * (tail1 - tail0); idx += GCT_NT) {
      const int cqq = idx / (tail1 - tail0), q = tail0 + idx % (tail1 - tail0);
      st4(A_hi + 4 * (q + Qp * cqq), make_float4(0.f, 0.f, 0.f, 0.f));
      st4(A_lo + 4 * (q + Qp * cqq), make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const int n_out = th * SW;                       // accumulator rows that can hold real pixels
  const int n_mt = (n_out + 127) / 128;            // <= GCT_MT (launch precondition)

  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32(16);
    const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);
    const uint32_t lbo_a = (uint32_t)Qp * 16u;
    for (int mt = 0; mt < ((a.dbg & 2) ? 0 : n_mt); ++mt) {
      bool first_mma = true;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t shift = (uint32_t)(mt * 128 + ky * dil * SW + kx * dil) * 16u;
          const uint32_t boff = (uint32_t)(ky * 3 + kx) * 1024u;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t dah = umma_desc(a_hi + shift + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t dal = umma_desc(a_lo + shift + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t dbh = umma_desc(b_hi + boff + ks * 256, 128, 512);
            const uint64_t dbl = umma_desc(b_lo + boff + ks * 256, 128, 512);
            umma_tf32(tmem_d + mt * 16, dah, dbh, idesc, first_mma ? 0u : 1u);
            umma_tf32(tmem_d + mt * 16, dal, dbh, idesc, 1u);
            umma_tf32(tmem_d + mt * 16, dah, dbl, idesc, 1u);
            first_mma = false;
          }
        }
    }
    umma_commit(&bar_done);
  }

  // ---- epilogue: warp (quarter, parity) reads rows quarter*32+lane of M-tiles parity, parity+2
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  const int quarter = wid & 3;
  float* out_s = a.out + (long long)net * a.out_net_stride + (long long)b * a.h * a.w * a.Cout;
  const float* bias = P + b_off + c0;
  const int cbase = out_off + c0;
  float s1 = 0.f, s2 = 0.f;
  for (int mt = wid >> 2; mt < n_mt; mt += 2) {
    float v[16];
    tmem_ld<16>(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * 16), v);
    const int q = mt * 128 + quarter * 32 + lane;
    const int y = q / SW, x = q - y * SW;
    if (q < n_out && x < a.w && !(a.dbg & 4)) {
      float* dst = out_s + ((long long)(y0 + y) * a.w + x) * a.Cout + cbase;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 bb = ld4(bias + j);
        const float o0 = v[j] + bb.x, o1 = v[j + 1] + bb.y, o2 = v[j + 2] + bb.z, o3 = v[j + 3] + bb.w;
        st4(dst + j, make_float4(o0, o1, o2, o3));
        float l;
        l = lrelu(o0); s1 += l; s2 += l * l;
        l = lrelu(o1); s1 += l; s2 += l * l;
        l = lrelu(o2); s1 += l; s2 += l * l;
        l = lrelu(o3); s1 += l; s2 += l * l;
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
  }
  tc_fence_before();
  __syncthreads();
  if (a.stats_out && tid < 2) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < GCT_NT / 32; ++i) t += (double)red[i][tid];
    atomicAdd(a.stats_out + 2 * ((long long)net * a.B + b) + tid, t);
  }
  if (wid == 0) tmem_dealloc(tmem_d, 64);
}

}  // namespace cnf
