// Fused grouped dilated 3x3 convs of one residual block (F:364-413, F:577-590): ALL dilation branches in one
// persistent, double-buffered launch.  Included by stnet_kernels.cu after tc_kernels.cuh (cp.async / mbarrier helpers).
//
// Work decomposition: a CTA owns one OCTET of input channels (8 consecutive channels of the 1x1-conv output) of one
// net and walks over batch items (S samples each).  Branch d reads the first nk/d channels (Q3), so octet o feeds
// every branch with nk/d > 8 o: the d = 1 group(s) of the octet, and -- for the low octets -- the d = 2, 4, ... groups
// that read the same channels.  The activation octet is therefore fetched from HBM/L2 ONCE for all branches
// (the per-branch kernels fetched it once per branch), gamma/beta of the octet and the octet's weights stay resident in
// shared memory for the whole launch, and the next item's raw tile (zero halo included) arrives by TMA
// (cp.async.bulk.tensor.4d with the 32-byte swizzle, completion on an mbarrier; 16-byte cp.async pieces where the shape does
// not allow a tensor map) while the current item is transformed (LReLU + LayerNorm, in place) and convolved.  Octets that feed more branches get proportionally more CTAs
// (host-side split, launch_gconv_oct).  With TMA tiles and a full-size CTA the kernel is warp-specialised (gconv_oct_body_ws): two
// producer warps do the waiting, the LayerNorm coefficients and the in-place transform of item i + 1 while eight consumer warps
// convolve item i (mbarrier hand-offs both ways).
//
// Compute: a thread owns a COLUMN SEGMENT of OCT_PX = 7 output pixels spaced `dil` rows apart x the full octet (8 in, 8 out).
// For each kx it loads the 9 input rows the segment's 3 ky taps touch ONCE (9 x 128-bit shared loads per channel quad) and
// reuses them from registers for the three ky taps, so every staged pixel is read ~3.9x instead of 9x.  The weights are
// block-diagonal with G x G blocks (G = group width in {1, 2, 4, 8}) and are read with warp-uniform (broadcast) addresses.
// The FMAs are FFMA2 over output-channel PAIRS with the activation as broadcast scalar operand (no extra accumulators, no
// spills at 8 warps x <= 200 registers).  Lanes of a warp are consecutive columns; the two channel quads of a pixel are
// stored swapped when bit 2 of the pixel index is set (= the TMA 32-byte swizzle), which makes every 8-lane phase of the
// 128-bit loads conflict-free without padding the tile.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

namespace cnf {

constexpr int OCT_MAX = 32;   // octets per tensor (nk <= 256)
#ifndef OCT_PX_N
#define OCT_PX_N 7
#endif
constexpr int OCT_PX = OCT_PX_N;   // rows per column segment
constexpr int OCT_MAXT = 256;      // threads per CTA (8 warps: 2 per scheduler, up to 255 registers each)

struct OctBranch {
  int dil, G, noct, out_off;  // noct = branch channels / 8; out_off = first channel of the branch in the concat
  int nsr;                    // column segments (4 rows spaced dil apart) per row residue class: ceil(ceil(h / dil) / 4)
  long long w_off, b_off;
  int w_smem;                 // float offset of this branch's weights inside the weight area
};

struct OctArgs {
  const float* in;
  float* out;
  long long in_net_stride, out_net_stride;
  const float* params;
  long long net_stride, g_off, be_off;
  const double* stats_in;
  double* stats_out;
  int B, h, w, Cin, Cout, ln;
  int S, halo, SW, SHW, n_items, n_oct, nsps;   // nsps = column-segment slots per sample (max over the branches)
  int tps;                                      // threads per sample: nsps rounded up to whole warps (nsps itself if < 32)
  int n_br;
  int tma;                                      // 1: the tiles (zero halo included) arrive by TMA (cp.async.bulk.tensor.4d, 32-byte swizzle)
  int vec8;                                     // 256-bit stores: out 32-byte aligned, Cout % 8 == 0
  int nbuf;                                     // 2: cp.async double buffering inside the CTA; 1: single buffer, two CTAs per SM overlap
  int ws;                                       // 1: warp-specialised body (TMA tiles, two buffers): OCT_PROD extra producer threads
  int dbg;                                      // CNF_OCT_DBG (timing experiments, wrong results): 1 skip branches, 2 transform, 4 copies, 8 epilogue, 16 coeffs, 32 weight loads (G = 8), 128 clock stamps
  OctBranch br[CNF_MAX_BRANCHES];
  unsigned short cta_first[OCT_MAX + 1];        // CTAs [cta_first[o], cta_first[o+1]) of a net own octet o
};

__host__ __device__ inline int oct_w_floats(int G) { return 9 * 2 * (4 * G + 4); }   // [tap][half][4 rows x G + pad 4]

// d.xy += a.xy * b.xy: one FFMA2 (sm_100 packed fp32 FMA).  A three-register FFMA issues every other cycle per scheduler
// on this part, FFMA2 does two FMAs in the same slot, so the packed form is what reaches the fp32 peak.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {   // a.xy * b.xy in one issue slot
  unsigned long long dd;
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(dd) : "l"(aa), "l"(bb));
  return *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ float2 fadd2(const float2 a, const float2 b) {
  unsigned long long dd;
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dd) : "l"(aa), "l"(bb));
  return *reinterpret_cast<float2*>(&dd);
}

// position of weight (ci, co) of a tap inside its quad block (ci & 4 selects the block): rows of G outputs per input channel
#ifndef OCT_PACKED
#define OCT_PACKED 1   // 1: FFMA2 over output-channel pairs with the activation as broadcast scalar operand; 0: scalar FFMA
#endif
__host__ __device__ inline int oct_w_pos(int G, int ci, int co) { return (ci & 3) * G + co; }

// byte-free helper: float offset of (pixel P, logical quad hq) inside a swizzled tile
__device__ __forceinline__ int oct_xoff(int P, int hq) { return P * 8 + ((hq ^ ((P >> 2) & 1)) << 2); }

// one branch of one item for this thread's column segment: output pixels (y0 + j dil, x), j < 4.  P0 = tile pixel index of
// (y0 - dil, x - dil) (first input row, kx = 0); rows are rstep = dil * SW pixels apart; pixel indices are clamped to Pmax
// (segments that stick out of the image read in-bounds garbage that is never stored).
template <int G>
__device__ __forceinline__ void oct_branch(const float* __restrict__ xb, const float* __restrict__ wb,
                                           const float* __restrict__ bias, int P0, int Pmax, int rstep, int dil,
                                           float* __restrict__ dst, long long jstride, int vmask, bool vec8, float2& s1,
                                           float2& s2, bool wskip = false) {
  constexpr int HS = 4 * G + 4;                 // floats per (tap, quad) weight block (4 of them padding)
  // acc[j][p] = output channels (2p, 2p+1) of pixel j, started at the bias.  A three-register FFMA issues every other cycle
  // on this part; FFMA2 (two FMAs per issue, the activation as broadcast scalar operand, the weight pair straight from a
  // 128-bit load) is what reaches the fp32 peak.
  float2 acc[OCT_PX][4];
  {
    const float4 b0 = ld4(bias), b1 = ld4(bias + 4);
#pragma unroll
    for (int j = 0; j < OCT_PX; ++j) {
      acc[j][0] = make_float2(b0.x, b0.y); acc[j][1] = make_float2(b0.z, b0.w);
      acc[j][2] = make_float2(b1.x, b1.y); acc[j][3] = make_float2(b1.z, b1.w);
    }
  }
  // float offset of pixel (row m, kx = 0) of the segment, clamped ONCE per row so that the kx = 1, 2 pixels (+ dil, + 2 dil)
  // stay inside the tile; every load that feeds a stored output is below the clamp (the last valid pixel of a row is the
  // kx = 2 one).  The swizzle term is bit 2 of the pixel index = bit 5 of the float offset.
  int row8[OCT_PX + 2];
#pragma unroll
  for (int m = 0; m < OCT_PX + 2; ++m) row8[m] = min(P0 + m * rstep, Pmax - 2 * dil) * 8;
#pragma unroll 1
  for (int kx = 0; kx < 3; ++kx) {
    const int k8 = kx * dil * 8;
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
      float4 xw[OCT_PX + 2];                     // logical quad hq of the OCT_PX + 2 input rows
#pragma unroll
      for (int m = 0; m < OCT_PX + 2; ++m) {
        const int p8 = row8[m] + k8;
        xw[m] = ld4(xb + p8 + (((p8 >> 3) & 4) ^ (hq << 2)));
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* wr = wb + ((ky * 3 + kx) * 2 + hq) * HS;
        if constexpr (G == 1) {
          // depthwise: channels (2k, 2k+1) of the quad times their own weights
          const float4 t = ld4(wr);
#pragma unroll
          for (int j = 0; j < OCT_PX; ++j) {
            const float4 xr = xw[j + ky];
            ffma2(acc[j][hq * 2], make_float2(xr.x, xr.y), make_float2(t.x, t.y));
            ffma2(acc[j][hq * 2 + 1], make_float2(xr.z, xr.w), make_float2(t.z, t.w));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {            // input channel hq * 4 + i; its row of G output weights
            float2 wv[G / 2];
            if constexpr (G == 8) {
              float4 t0 = make_float4(0.5f, 0.25f, 0.125f, 0.0625f), t1 = t0;
              if (!wskip) { t0 = ld4(wr + i * 8); t1 = ld4(wr + i * 8 + 4); }
              wv[0] = make_float2(t0.x, t0.y); wv[1] = make_float2(t0.z, t0.w);
              wv[2] = make_float2(t1.x, t1.y); wv[3] = make_float2(t1.z, t1.w);
            } else if constexpr (G == 4) {
              const float4 t0 = ld4(wr + i * 4);
              wv[0] = make_float2(t0.x, t0.y); wv[1] = make_float2(t0.z, t0.w);
            } else {
              wv[0] = *reinterpret_cast<const float2*>(wr + i * 2);
            }
            // first output pair of this input channel's group: G == 8: 0; G == 4: quad hq; G == 2: group i / 2 of the quad
            constexpr int NP = G / 2;
            const int base = G == 8 ? 0 : G == 4 ? hq * 2 : hq * 2 + (i >> 1);
#pragma unroll
            for (int j = 0; j < OCT_PX; ++j) {
              const float4 xr = xw[j + ky];
              const float xv = i == 0 ? xr.x : i == 1 ? xr.y : i == 2 ? xr.z : xr.w;
#pragma unroll
              for (int c = 0; c < NP; ++c) {
#if OCT_PACKED
                ffma2(acc[j][base + c], make_float2(xv, xv), wv[c]);
#else
                acc[j][base + c].x = fmaf(xv, wv[c].x, acc[j][base + c].x);
                acc[j][base + c].y = fmaf(xv, wv[c].y, acc[j][base + c].y);
#endif
              }
            }
          }
        }
      }
    }
  }
  // epilogue: LReLU statistics of the outputs in packed form (one issue slot per channel pair), one 256-bit store per
  // pixel (32 lanes = 32 different lines either way; half the store instructions and tag look-ups)
  const float2 slope = make_float2(CNF_LRELU_SLOPE, CNF_LRELU_SLOPE);
#pragma unroll
  for (int j = 0; j < OCT_PX; ++j) {
    if (!((vmask >> j) & 1)) continue;
    float* d = dst + j * jstride;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float2 t = fmul2(acc[j][c], slope);
      const float2 l = make_float2(fmaxf(acc[j][c].x, t.x), fmaxf(acc[j][c].y, t.y));
      s1 = fadd2(s1, l);
      ffma2(s2, l, l);
    }
    if (vec8) {
      const float o[8] = {acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y, acc[j][2].x, acc[j][2].y, acc[j][3].x, acc[j][3].y};
      st8(d, o);
    } else {
      st4(d, make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y));
      st4(d + 4, make_float4(acc[j][2].x, acc[j][2].y, acc[j][3].x, acc[j][3].y));
    }
  }
}

#define OCT_STAMP(slot)                                                                                          \
  do {                                                                                                           \
    if ((a.dbg & 128) && rank == 0 && net == 0 && tid == 0 && it < 32 && o < 8) g_tc3_clk[(o * 32 + it) * 8 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ void gconv_oct_body(const OctArgs& a, const CUtensorMap* tmap) {
  extern __shared__ __align__(1024) float oct_smem[];
  __shared__ __align__(8) uint64_t tma_bar[2];
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const int net = blockIdx.y;
  const int hw = a.h * a.w;
  // which octet, and this CTA's rank among the CTAs of the octet
  int o = 0;
  while (o + 1 < a.n_oct && (int)blockIdx.x >= (int)a.cta_first[o + 1]) ++o;
  const int rank = blockIdx.x - a.cta_first[o], nshare = a.cta_first[o + 1] - a.cta_first[o];

  const int xsz = a.S * a.SHW * 8;                          // floats per x buffer
  float* xbuf = oct_smem;                                   // [2][S][SHW][8]
  float* gb = xbuf + a.nbuf * xsz;                               // [2][hw][8] gamma, beta of the octet
  float* w_s = gb + (a.ln ? 2 * hw * 8 : 0);                // per branch [9][2][4G+4]
  int wtot = 0;
  for (int b = 0; b < a.n_br; ++b) wtot += oct_w_floats(a.br[b].G);
  float* b_s = w_s + wtot;                                  // [n_br][8]
  float* mr = b_s + a.n_br * 8;                             // [2][S][2]  (rstd, -mean rstd) per buffer parity
  float* red = mr + 4 * a.S;                                // [NT][2] per-thread, or [NT/32][4] per-warp partial sums
  unsigned* segtab = reinterpret_cast<unsigned*>(red + 2 * NT);            // [n_br][NT] this thread's column segment per branch
  unsigned short* pt = reinterpret_cast<unsigned short*>(segtab + a.n_br * NT);   // [hw] pixel offset inside a sample tile

  const float* P = a.params + (long long)net * a.net_stride;
  const float* src_n = a.in + (long long)net * a.in_net_stride + o * 8;
  float* out_n = a.out + (long long)net * a.out_net_stride + o * 8;

  // ---- one-time setup: zero tiles (the halo stays zero for the whole launch), tables, weights, gamma/beta ----
  if (!a.tma) {
    for (int i = tid; i < a.nbuf * xsz / 4; i += NT) st4(xbuf + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
  } else if (tid == 0) {
    mbar_init(&tma_bar[0], 1);
    mbar_init(&tma_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int p = tid; p < hw; p += NT) {
    const int y = p / a.w, x = p - y * a.w;
    pt[p] = (unsigned short)((y + a.halo) * a.SW + x + a.halo);
  }
  for (int b = 0; b < a.n_br; ++b) {
    const OctBranch& br = a.br[b];
    if (o >= br.noct) continue;
    const int G = br.G, HS = 4 * G + 4;
    float* wd = w_s + br.w_smem;
    const float* wsrc = P + br.w_off + (long long)o * (8 / G) * 9 * G * G;   // groups o*8/G .. of [group][tap][ci][co]
    for (int i = tid; i < 9 * 8 * G; i += NT) {
      const int co = i % G, ci = (i / G) % 8, tap = i / (8 * G);
      const int grp = ci / G, cig = ci - grp * G;
      wd[tap * 2 * HS + (ci >> 2) * HS + oct_w_pos(G, ci, co)] = wsrc[((grp * 9 + tap) * G + cig) * G + co];
    }
    if (tid < 8) b_s[b * 8 + tid] = P[br.b_off + o * 8 + tid];
  }
  if (a.ln) {
    const float* gam = P + a.g_off + o * 8;
    const float* bet = P + a.be_off + o * 8;
    for (int i0 = tid; i0 < hw * 2; i0 += 4 * NT) {         // 8 global loads in flight per thread
      float4 gv[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = min(i0 + u * NT, hw * 2 - 1);
        gv[u] = ld4(gam + (long long)(i >> 1) * a.Cin + (i & 1) * 4);
        bv[u] = ld4(bet + (long long)(i >> 1) * a.Cin + (i & 1) * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * NT;
        if (i < hw * 2) {
          st4(gb + (i >> 1) * 8 + (i & 1) * 4, gv[u]);
          st4(gb + hw * 8 + (i >> 1) * 8 + (i & 1) * 4, bv[u]);
        }
      }
    }
  }

  // this thread: sample q of the item, slot sl of the sample (the slot -> column-segment map depends on the branch)
  // (a sample's threads are whole warps, so its reduction order -- and therefore every bit of the result -- does not
  // depend on the batch size or on the sample's position inside the item)
  const int q = tid / a.tps, sl = tid - q * a.tps;
  const bool tactive = q < a.S && sl < a.nsps;
  const int seg_t = sl / a.w, seg_x = sl - seg_t * a.w;
  // slot -> (column x, row residue r mod dil, segment k): rows y0 + j dil with y0 = r + OCT_PX k dil.  The map does not depend
  // on the item, so it is packed once per branch: bit 31 = the thread has work, bits 8-14 = valid rows, bits 0-7 = y0
  for (int b = 0; b < a.n_br; ++b) {
    const OctBranch& br = a.br[b];
    const int d = br.dil;
    const int r = seg_t / br.nsr, k = seg_t - r * br.nsr;
    const int y0 = r + OCT_PX * k * d;
    unsigned e = 0;
    if (tactive && o < br.noct && r < d && y0 < a.h) {
      int vmask = 0;
#pragma unroll
      for (int j = 0; j < OCT_PX; ++j) vmask |= (y0 + j * d < a.h && !(a.dbg & 8)) ? (1 << j) : 0;
      e = 0x80000000u | ((unsigned)vmask << 8) | (unsigned)y0;
    }
    segtab[b * NT + tid] = e;     // read back by this thread only
  }
  const double inv_n = 1.0 / ((double)hw * (double)a.Cin);   // mean and centred variance in fp64 (see ln_coeffs)
  auto coeffs = [&](int item, int par) {   // LayerNorm coefficients of the samples of `item` -> mr[par]
    if (tid < a.S) {
      float sc = 1.f, sh = 0.f;
      const int s = item * a.S + tid;
      if (a.ln && s < a.B) {
        const double* sp = a.stats_in + 2 * ((long long)net * a.B + s);
        const double md = sp[0] * inv_n;
        const float m = (float)md;
        const float var = fmaxf((float)(sp[1] * inv_n - md * md), 0.f);
        sc = 1.0f / sqrtf(var + (float)CNF_LN_EPS);
        sh = -m * sc;
      }
      mr[(par * a.S + tid) * 2] = sc;
      mr[(par * a.S + tid) * 2 + 1] = sh;
    }
  };
  auto issue = [&](int item, int par) {    // the raw octet rows of `item` -> buffer `par` (TMA tiles or cp.async pieces)
    const int b0 = item * a.S, ns = min(a.S, a.B - b0);
    float* xb = xbuf + par * xsz;
    if (a.tma) {
      // one thread: order the CTA's earlier generic-proxy accesses to the buffer before the async-proxy writes, arm the
      // barrier with the byte count, one box [SH][SW][8 channels] per sample; out-of-image rows / columns are zero-filled
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bar = smem_u32(&tma_bar[par]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(ns * a.SHW * 32)) : "memory");
        for (int s = 0; s < ns; ++s) {
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(smem_u32(xb + s * a.SHW * 8)), "l"((unsigned long long)tmap), "r"(o * 8), "r"(-a.halo), "r"(-a.halo),
              "r"(net * a.B + b0 + s), "r"(bar)
              : "memory");
        }
      }
      return;
    }
    for (int s = 0; s < ns; ++s) {
      const float* src = src_n + (long long)(b0 + s) * hw * a.Cin;
      float* dst = xb + s * a.SHW * 8;
      for (int i0 = tid; i0 < hw * 2; i0 += 4 * NT) {       // 4 independent table look-ups in flight
        int off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = min(i0 + u * NT, hw * 2 - 1);
          off[u] = oct_xoff(s * a.SHW + pt[i >> 1], i & 1) - s * a.SHW * 8;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * NT;
          if (i < hw * 2) cp_async16_cg(dst + off[u], src + (long long)(i >> 1) * a.Cin + (i & 1) * 4);
        }
      }
    }
  };
  __syncthreads();                         // pt[] and the zero fill are visible before the first copies land
  int item = rank;
  const bool dbl = a.nbuf == 2;
  const int ctid = tid - (NT - 32);          // lanes of the last warp compute the next item's LayerNorm coefficients
  double nx0 = 0.0, nx1 = 1.0;
  if (dbl && item < a.n_items) {
    issue(item, 0);
    coeffs(item, 0);
  }
  cp_async_commit();

  for (int it = 0; item < a.n_items; ++it, item += nshare) {
    const int par = dbl ? (it & 1) : 0;
    const int next = item + nshare;
    OCT_STAMP(0);
    if (dbl) {
      if (next < a.n_items && !(a.dbg & 4)) issue(next, par ^ 1);
      cp_async_commit();
      // LayerNorm sums of the NEXT item: loaded now by a few lanes of the last warp, turned into coefficients after the
      // branches (the load latency never sits in front of a barrier)
      if (ctid >= 0 && ctid < a.S && a.ln && next < a.n_items && next * a.S + ctid < a.B && !(a.dbg & 16)) {
        const double* sp = a.stats_in + 2 * ((long long)net * a.B + next * a.S + ctid);
        nx0 = sp[0];
        nx1 = sp[1];
      }
      if (a.tma) mbar_wait(&tma_bar[par], (it >> 1) & 1);   // the tile(s) of `item` have landed (async proxy -> visible)
      else cp_async_wait<1>();             // this thread's copies of `item` have landed
    } else {
      if (!(a.dbg & 4)) issue(item, 0);    // single buffer: the SM's other CTA computes while these copies fly
      cp_async_commit();
      coeffs(item, 0);
      cp_async_wait<0>();
    }
    OCT_STAMP(1);
    __syncthreads();                       // ... and everybody else's; mr[par] is visible
    OCT_STAMP(2);
    const int b0 = item * a.S, ns = min(a.S, a.B - b0);
    float* xb = xbuf + par * xsz;
    // ---- LReLU + LayerNorm in place on the interior pixels (4 float4 units per thread in flight) ----
    for (int s = 0; s < ((a.dbg & 2) ? 0 : ns); ++s) {
      const float sc = mr[(par * a.S + s) * 2], sh = mr[(par * a.S + s) * 2 + 1];
      float* xs = xb + s * a.SHW * 8;
      for (int i0 = tid; i0 < hw * 2; i0 += 4 * NT) {
        float4 v[4], g[4], be[4];
        float* px[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = min(i0 + u * NT, hw * 2 - 1);
          const int p = i >> 1, qd = (i & 1) * 4;
          px[u] = xb + oct_xoff(s * a.SHW + pt[p], i & 1);
          v[u] = ld4(px[u]);
          if (a.ln) {
            g[u] = ld4(gb + p * 8 + qd);
            be[u] = ld4(gb + hw * 8 + p * 8 + qd);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float4 t = v[u];
          t.x = lrelu(t.x); t.y = lrelu(t.y); t.z = lrelu(t.z); t.w = lrelu(t.w);
          if (a.ln) {
            t.x = fmaf(fmaf(t.x, sc, sh), g[u].x, be[u].x);
            t.y = fmaf(fmaf(t.y, sc, sh), g[u].y, be[u].y);
            t.z = fmaf(fmaf(t.z, sc, sh), g[u].z, be[u].z);
            t.w = fmaf(fmaf(t.w, sc, sh), g[u].w, be[u].w);
          }
          if (i0 + u * NT < hw * 2) st4(px[u], t);
        }
      }
    }
    OCT_STAMP(3);
    __syncthreads();
    OCT_STAMP(4);
    // ---- the branches that read this octet ----
    const bool live = tactive && q < ns && !(a.dbg & 1);
    float2 s1p = make_float2(0.f, 0.f), s2p = make_float2(0.f, 0.f);
    if (live) {
      const int Pmax = (q + 1) * a.SHW - 1;
      for (int b = 0; b < a.n_br; ++b) {
        const unsigned e = segtab[b * NT + tid];
        if (!(e >> 31)) continue;
        const OctBranch& br = a.br[b];
        const int d = br.dil, x = seg_x;
        const int y0 = (int)(e & 255u), vmask = (int)((e >> 8) & 127u);
        const int P0 = q * a.SHW + (y0 - d + a.halo) * a.SW + x - d + a.halo;
        const float* wb = w_s + br.w_smem;
        float* dptr = out_n + ((long long)(b0 + q) * hw + (long long)y0 * a.w + x) * a.Cout + br.out_off;
        const long long jstride = (long long)d * a.w * a.Cout;
        const bool vec8 = a.vec8 && !(br.out_off & 7);
        switch (br.G) {
          case 8: oct_branch<8>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p, (a.dbg & 32) != 0); break;
          case 4: oct_branch<4>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p); break;
          case 2: oct_branch<2>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p); break;
          default: oct_branch<1>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p); break;
        }
      }
    }
    const float s1 = s1p.x + s1p.y, s2 = s2p.x + s2p.y;
    if (dbl && ctid >= 0 && ctid < a.S && next < a.n_items && !(a.dbg & 16)) {
      float sc = 1.f, sh = 0.f;
      if (a.ln && next * a.S + ctid < a.B) {
        const double md = nx0 * inv_n;
        const float m = (float)md;
        const float var = fmaxf((float)(nx1 * inv_n - md * md), 0.f);
        sc = 1.0f / sqrtf(var + (float)CNF_LN_EPS);
        sh = -m * sc;
      }
      mr[((par ^ 1) * a.S + ctid) * 2] = sc;      // read after the next iteration's first barrier
      mr[((par ^ 1) * a.S + ctid) * 2 + 1] = sh;
    }
    OCT_STAMP(5);
    // statistics of LReLU(out): fixed-order reduction (bit-reproducible run to run), fp64 across CTAs
    if (a.stats_out) {
      if (a.nsps >= 32) {
        const float a1 = warp_sum(s1), a2 = warp_sum(s2);       // the warp belongs to one sample
        if (lane == 0) {
          red[2 * (tid >> 5)] = a1;
          red[2 * (tid >> 5) + 1] = a2;
        }
      } else {
        red[2 * tid] = s1;
        red[2 * tid + 1] = s2;
      }
    }
    __syncthreads();                       // everyone is done with xb (the next iteration refills it); red is complete
    OCT_STAMP(6);
    const int ft = NT >= 64 ? tid - 32 : tid;   // flushed by warp 1: thread 0 goes straight to the next TMA issue
    if (a.stats_out && ft >= 0 && ft < 2 * a.S && (ft >> 1) < ns) {
      const int s = ft >> 1, which = ft & 1;
      double t = 0.0;
      const int cnt = a.nsps >= 32 ? (a.tps >> 5) : a.nsps;      // partial sums of sample s: per warp or per thread
      for (int k = 0; k < cnt; ++k) t += (double)red[2 * (s * cnt + k) + which];
      atomicAdd(a.stats_out + 2 * ((long long)net * a.B + b0 + s) + which, t);
    }
  }
  cp_async_wait<0>();
}

// Warp-specialised variant (TMA tiles, two buffers): the LAST two warps of the CTA are producers.  They wait for the TMA tile
// of item i + 1, turn the per-sample LayerNorm sums into coefficients, apply LReLU + LayerNorm in place and hand the buffer to
// the consumer warps through an mbarrier, all while the consumers convolve item i; as soon as the consumers release a buffer
// the producers re-arm its TMA.  The transform (2.6 k cycles per item on the whole CTA, clock stamps of round 1) and the
// wait for the copies (0.8 k) no longer sit in front of the branches.  Consumer threads keep the thread -> column-segment
// map and the fixed-order statistics of the single-role body, so results are bit-identical to it.
constexpr int OCT_PROD = 64;   // producer threads (2 warps: 200 registers per thread stay available to the consumers)

__device__ __forceinline__ void gconv_oct_body_ws(const OctArgs& a, const CUtensorMap* tmap) {
  extern __shared__ __align__(1024) float oct_smem[];
  __shared__ __align__(8) uint64_t tma_bar[2], xf_full[2], x_free[2];
  const int tid = threadIdx.x, NT = blockDim.x - OCT_PROD, lane = tid & 31;   // NT = consumer threads
  const int net = blockIdx.y;
  const int hw = a.h * a.w;
  int o = 0;
  while (o + 1 < a.n_oct && (int)blockIdx.x >= (int)a.cta_first[o + 1]) ++o;
  const int rank = blockIdx.x - a.cta_first[o], nshare = a.cta_first[o + 1] - a.cta_first[o];

  const int xsz = a.S * a.SHW * 8;                          // floats per x buffer
  float* xbuf = oct_smem;                                   // [2][S][SHW][8]
  float* gb = xbuf + 2 * xsz;                               // [2][hw][8] gamma, beta of the octet
  float* w_s = gb + (a.ln ? 2 * hw * 8 : 0);                // per branch [9][2][4G+4]
  int wtot = 0;
  for (int b = 0; b < a.n_br; ++b) wtot += oct_w_floats(a.br[b].G);
  float* b_s = w_s + wtot;                                  // [n_br][8]
  float* mr = b_s + a.n_br * 8;                             // [2][S][2]  (rstd, -mean rstd) per buffer parity
  float* red = mr + 4 * a.S;                                // [2 parities][OCT_MAXT][2] partial sums
  unsigned* segtab = reinterpret_cast<unsigned*>(red + 4 * OCT_MAXT);              // [n_br][OCT_MAXT]
  unsigned short* pt = reinterpret_cast<unsigned short*>(segtab + a.n_br * OCT_MAXT);   // [hw] pixel offset inside a sample tile

  const float* P = a.params + (long long)net * a.net_stride;
  float* out_n = a.out + (long long)net * a.out_net_stride + o * 8;
  const int NTA = blockDim.x;

  // ---- one-time setup (all threads): barriers, tables, weights, gamma/beta ----
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tma_bar[i], 1);
      mbar_init(&xf_full[i], OCT_PROD / 32);       // one arrival per producer warp
      mbar_init(&x_free[i], NT / 32);              // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int p = tid; p < hw; p += NTA) {
    const int y = p / a.w, x = p - y * a.w;
    pt[p] = (unsigned short)((y + a.halo) * a.SW + x + a.halo);
  }
  for (int b = 0; b < a.n_br; ++b) {
    const OctBranch& br = a.br[b];
    if (o >= br.noct) continue;
    const int G = br.G, HS = 4 * G + 4;
    float* wd = w_s + br.w_smem;
    const float* wsrc = P + br.w_off + (long long)o * (8 / G) * 9 * G * G;
    for (int i = tid; i < 9 * 8 * G; i += NTA) {
      const int co = i % G, ci = (i / G) % 8, tap = i / (8 * G);
      const int grp = ci / G, cig = ci - grp * G;
      wd[tap * 2 * HS + (ci >> 2) * HS + oct_w_pos(G, ci, co)] = wsrc[((grp * 9 + tap) * G + cig) * G + co];
    }
    if (tid < 8) b_s[b * 8 + tid] = P[br.b_off + o * 8 + tid];
  }
  if (a.ln) {
    const float* gam = P + a.g_off + o * 8;
    const float* bet = P + a.be_off + o * 8;
    for (int i0 = tid; i0 < hw * 2; i0 += 4 * NTA) {
      float4 gv[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = min(i0 + u * NTA, hw * 2 - 1);
        gv[u] = ld4(gam + (long long)(i >> 1) * a.Cin + (i & 1) * 4);
        bv[u] = ld4(bet + (long long)(i >> 1) * a.Cin + (i & 1) * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * NTA;
        if (i < hw * 2) {
          st4(gb + (i >> 1) * 8 + (i & 1) * 4, gv[u]);
          st4(gb + hw * 8 + (i >> 1) * 8 + (i & 1) * 4, bv[u]);
        }
      }
    }
  }
  const int q = tid / a.tps, sl = tid - q * a.tps;
  const bool tactive = tid < NT && q < a.S && sl < a.nsps;
  const int seg_t = sl / a.w, seg_x = sl - seg_t * a.w;
  if (tid < NT) {
    for (int b = 0; b < a.n_br; ++b) {
      const OctBranch& br = a.br[b];
      const int d = br.dil;
      const int r = seg_t / br.nsr, k = seg_t - r * br.nsr;
      const int y0 = r + OCT_PX * k * d;
      unsigned e = 0;
      if (tactive && o < br.noct && r < d && y0 < a.h) {
        int vmask = 0;
#pragma unroll
        for (int j = 0; j < OCT_PX; ++j) vmask |= (y0 + j * d < a.h) ? (1 << j) : 0;
        e = 0x80000000u | ((unsigned)vmask << 8) | (unsigned)y0;
      }
      segtab[b * OCT_MAXT + tid] = e;
    }
  }
  __syncthreads();

  if (tid >= NT) {
    // =============================== producers ===============================
    const int ptid = tid - NT;
    const double inv_n = 1.0 / ((double)hw * (double)a.Cin);   // mean and centred variance in fp64 (see ln_coeffs)
    auto issue = [&](int item, int par) {
      if (ptid == 0) {
        const int b0 = item * a.S, ns = min(a.S, a.B - b0);
        float* xb = xbuf + par * xsz;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the consumers' generic reads of this buffer come first
        const uint32_t bar = smem_u32(&tma_bar[par]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(ns * a.SHW * 32)) : "memory");
        for (int s = 0; s < ns; ++s) {
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(smem_u32(xb + s * a.SHW * 8)), "l"((unsigned long long)tmap), "r"(o * 8), "r"(-a.halo), "r"(-a.halo),
              "r"(net * a.B + b0 + s), "r"(bar)
              : "memory");
        }
      }
    };
    int item = rank;
    if (item < a.n_items) issue(item, 0);
    for (int it = 0; item < a.n_items; ++it, item += nshare) {
      const int par = it & 1;
      const int b0 = item * a.S, ns = min(a.S, a.B - b0);
      // LayerNorm coefficients of this item's samples (the loads fly while the tile lands)
      if (ptid < a.S) {
        float sc = 1.f, sh = 0.f;
        const int s = b0 + ptid;
        if (a.ln && s < a.B) {
          const double* sp = a.stats_in + 2 * ((long long)net * a.B + s);
          const double md = sp[0] * inv_n;
          const float m = (float)md;
          const float var = fmaxf((float)(sp[1] * inv_n - md * md), 0.f);
          sc = 1.0f / sqrtf(var + (float)CNF_LN_EPS);
          sh = -m * sc;
        }
        mr[(par * a.S + ptid) * 2] = sc;
        mr[(par * a.S + ptid) * 2 + 1] = sh;
      }
      named_bar_sync(2, OCT_PROD);                 // coefficients visible to both producer warps
      mbar_wait(&tma_bar[par], (it >> 1) & 1);     // the tile(s) of `item` have landed (async proxy -> visible)
      float* xb = xbuf + par * xsz;
      for (int s = 0; s < ns; ++s) {
        const float sc = mr[(par * a.S + s) * 2], sh = mr[(par * a.S + s) * 2 + 1];
        for (int i0 = ptid; i0 < hw * 2; i0 += 8 * OCT_PROD) {   // 8 units in flight per thread
          float4 v[8], g[8], be[8];
          float* px[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = min(i0 + u * OCT_PROD, hw * 2 - 1);
            const int p = i >> 1, qd = (i & 1) * 4;
            px[u] = xb + oct_xoff(s * a.SHW + pt[p], i & 1);
            v[u] = ld4(px[u]);
            if (a.ln) {
              g[u] = ld4(gb + p * 8 + qd);
              be[u] = ld4(gb + hw * 8 + p * 8 + qd);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 t = v[u];
            t.x = lrelu(t.x); t.y = lrelu(t.y); t.z = lrelu(t.z); t.w = lrelu(t.w);
            if (a.ln) {
              t.x = fmaf(fmaf(t.x, sc, sh), g[u].x, be[u].x);
              t.y = fmaf(fmaf(t.y, sc, sh), g[u].y, be[u].y);
              t.z = fmaf(fmaf(t.z, sc, sh), g[u].z, be[u].z);
              t.w = fmaf(fmaf(t.w, sc, sh), g[u].w, be[u].w);
            }
            if (i0 + u * OCT_PROD < hw * 2) st4(px[u], t);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&xf_full[par]);   // transformed buffer -> consumers (release)
      // re-arm the other buffer for the item after next as soon as the consumers have released it
      const int next = item + nshare;
      if (next < a.n_items) {
        if (it >= 1) mbar_wait(&x_free[par ^ 1], ((it - 1) >> 1) & 1);   // consumers are done with item it - 1
        issue(next, par ^ 1);
      }
    }
    return;
  }

  // =============================== consumers ===============================
  int item = rank;
  for (int it = 0; item < a.n_items; ++it, item += nshare) {
    const int par = it & 1;
    const int b0 = item * a.S, ns = min(a.S, a.B - b0);
    float* xb = xbuf + par * xsz;
    mbar_wait(&xf_full[par], (it >> 1) & 1);       // LReLU + LayerNorm applied to the whole buffer (acquire)
    const bool live = tactive && q < ns;
    float2 s1p = make_float2(0.f, 0.f), s2p = make_float2(0.f, 0.f);
    if (live) {
      const int Pmax = (q + 1) * a.SHW - 1;
      for (int b = 0; b < a.n_br; ++b) {
        const unsigned e = segtab[b * OCT_MAXT + tid];
        if (!(e >> 31)) continue;
        const OctBranch& br = a.br[b];
        const int d = br.dil, x = seg_x;
        const int y0 = (int)(e & 255u), vmask = (int)((e >> 8) & 127u);
        const int P0 = q * a.SHW + (y0 - d + a.halo) * a.SW + x - d + a.halo;
        const float* wb = w_s + br.w_smem;
        float* dptr = out_n + ((long long)(b0 + q) * hw + (long long)y0 * a.w + x) * a.Cout + br.out_off;
        const long long jstride = (long long)d * a.w * a.Cout;
        const bool vec8 = a.vec8 && !(br.out_off & 7);
        switch (br.G) {
          case 8: oct_branch<8>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p); break;
          case 4: oct_branch<4>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p); break;
          case 2: oct_branch<2>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p); break;
          default: oct_branch<1>(xb, wb, b_s + b * 8, P0, Pmax, d * a.SW, d, dptr, jstride, vmask, vec8, s1p, s2p); break;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&x_free[par]);      // this warp no longer reads the buffer (release)
    const float s1 = s1p.x + s1p.y, s2 = s2p.x + s2p.y;
    // statistics of LReLU(out): fixed-order reduction (bit-reproducible run to run), fp64 across CTAs
    if (a.stats_out) {
      float* rd = red + par * 2 * OCT_MAXT;        // two parities: the flush of item i and the writes of item i + 1 never meet
      if (a.nsps >= 32) {
        const float a1 = warp_sum(s1), a2 = warp_sum(s2);       // the warp belongs to one sample
        if (lane == 0) {
          rd[2 * (tid >> 5)] = a1;
          rd[2 * (tid >> 5) + 1] = a2;
        }
      } else {
        rd[2 * tid] = s1;
        rd[2 * tid + 1] = s2;
      }
      named_bar_sync(1, NT);
      const int ft = NT >= 64 ? tid - 32 : tid;   // flushed by warp 1
      if (ft >= 0 && ft < 2 * a.S && (ft >> 1) < ns) {
        const int s = ft >> 1, which = ft & 1;
        double t = 0.0;
        const int cnt = a.nsps >= 32 ? (a.tps >> 5) : a.nsps;      // partial sums of sample s: per warp or per thread
        for (int k = 0; k < cnt; ++k) t += (double)rd[2 * (s * cnt + k) + which];
        atomicAdd(a.stats_out + 2 * ((long long)net * a.B + b0 + s) + which, t);
      }
    }
  }
}

// 8 warps (2 per scheduler, each scheduler owns 16 K registers): up to 255 registers per thread.  The shared-memory carve-out
// leaves almost no L1, so a spilled value costs an L2 round trip -- the kernel must not spill (ptxas -v: 0 bytes).
template <int MAXT>
__global__ void __launch_bounds__(MAXT + OCT_PROD, 1) gconv_oct_kernel(const OctArgs a, const __grid_constant__ CUtensorMap tmap) {
  if (a.ws) gconv_oct_body_ws(a, &tmap);
  else gconv_oct_body(a, &tmap);
}

// 4-D tensor map over the input [2 nets * B][h][w][Cin] with box [1][h + 2 halo][w + 2 halo][8] and the 32-byte swizzle (16-byte
// chunk index ^= address bit 7, i.e. the tile's quad swizzle).  Returns false if the driver entry point is unavailable.
static bool oct_make_tmap(CUtensorMap* tm, const float* in, int rows, int h, int w, int Cin, int SH, int SW) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (encode_fn)ptr;
  }
  if (!fn) return false;
  const cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)rows};
  const cuuint64_t gstr[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)w * Cin * 4, (cuuint64_t)h * w * Cin * 4};
  const cuuint32_t box[4] = {8, (cuuint32_t)SW, (cuuint32_t)SH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Host side: eligibility, shared-memory budget, CTA split.  Returns CNF_NOT_ELIGIBLE if the shape is not covered (caller
// falls back to the per-branch kernels), 0 on success, a cudaError otherwise.
static int launch_gconv_oct(const GconvArgs& g, cudaStream_t st) {
  if (g.bwd || g.ks != 3 || g.n_br < 1) return CNF_NOT_ELIGIBLE;
  if ((g.Cin % 4) || (g.Cout % 4)) return CNF_NOT_ELIGIBLE;
  OctArgs a{};
  a.in = g.in; a.out = g.out; a.in_net_stride = g.in_net_stride; a.out_net_stride = g.out_net_stride;
  a.params = g.params; a.net_stride = g.net_stride; a.g_off = g.g_off; a.be_off = g.be_off;
  a.stats_in = g.stats_in; a.stats_out = g.stats_out;
  a.B = g.B; a.h = g.h; a.w = g.w; a.Cin = g.Cin; a.Cout = g.Cout; a.ln = g.ln;
  a.n_br = g.n_br;
  a.vec8 = (((uintptr_t)g.out & 31) == 0 && (g.Cout % 8) == 0 && ((g.out_net_stride * 4) % 32) == 0) ? 1 : 0;
  { static int d = -1; if (d < 0) d = knob_int("OCT_DBG", 0); a.dbg = d; }   // ablation modes: -DCNF_DEBUG builds only
  int halo = 0, n_oct = 0, wtot = 0;
  for (int i = 0; i < g.n_br; ++i) {
    const GconvBranch& b = g.br[i];
    const int G = b.gin, ch = b.groups * b.gin;
    if (b.gin != b.gout || !(G == 1 || G == 2 || G == 4 || G == 8) || (ch % 8) || b.in_off != 0 || (b.out_off % 4)) return CNF_NOT_ELIGIBLE;
    if (ch > g.Cin) return CNF_NOT_ELIGIBLE;
    OctBranch& ob = a.br[i];
    ob.dil = b.dil; ob.G = G; ob.noct = ch / 8; ob.out_off = b.out_off; ob.w_off = b.w_off; ob.b_off = b.b_off;
    ob.w_smem = wtot;
    wtot += oct_w_floats(G);
    halo = std::max(halo, b.dil);
    n_oct = std::max(n_oct, ob.noct);
  }
  if (n_oct > OCT_MAX || n_oct < 1) return CNF_NOT_ELIGIBLE;
  const int hw = g.h * g.w;
  a.halo = halo; a.SW = g.w + 2 * halo; a.SHW = (g.h + 2 * halo) * a.SW; a.n_oct = n_oct;
  if ((long long)a.SHW >= 65536 / 1 || g.h > 255) return CNF_NOT_ELIGIBLE;     // pt[] is 16 bit, y0 of a segment 8 bit
  a.nsps = 0;
  for (int i = 0; i < g.n_br; ++i) {
    const int d = a.br[i].dil;
    a.br[i].nsr = ((g.h + d - 1) / d + OCT_PX - 1) / OCT_PX;
    a.nsps = std::max(a.nsps, g.w * d * a.br[i].nsr);
  }
  a.tps = a.nsps < 32 ? a.nsps : (a.nsps + 31) / 32 * 32;
  // samples per item: as many as keep <= 512 threads busy and fit the shared-memory budget
  const size_t budget = 227 * 1024 - 1024;
  static int nbuf_env = 0;
  if (!nbuf_env) nbuf_env = knob_int("OCT_NBUF", 2) == 1 ? 1 : 2;
  a.nbuf = nbuf_env;
  auto smem_for = [&](int S) {
    size_t f = (size_t)a.nbuf * S * a.SHW * 8 + (g.ln ? (size_t)2 * hw * 8 : 0) + wtot + (size_t)g.n_br * 8 + 4 * S + 4 * OCT_MAXT +
               (size_t)g.n_br * OCT_MAXT;
    return f * sizeof(float) + (((size_t)hw * 2 + 15) & ~(size_t)15);
  };
  if (a.tps > OCT_MAXT || smem_for(1) > budget) return CNF_NOT_ELIGIBLE;
  int nsm = 0;
  if (device_sm_count(&nsm) != 0 || nsm < 2) nsm = 148;
  // walk down from the largest S (<= 512 threads, fits shared memory) until every CTA slot has >= 3 items to pipeline,
  // but keep at least ~96 busy threads per CTA
  int S = 0, NT = 0, slots = 0;
  static int max_per_sm = 0;
  if (!max_per_sm) max_per_sm = std::max(1, knob_int("OCT_PER_SM", 4));
  static int s_cap = 0;
  if (!s_cap) s_cap = std::max(1, knob_int("OCT_S", 512));
  for (int s = std::max(1, std::min(std::min(OCT_MAXT / a.tps, a.nbuf == 1 ? 1 : s_cap), g.B)); s >= 1; --s) {
    if (smem_for(s) > budget) continue;
    if (S && s * a.tps < 96) break;
    S = s;
    NT = std::min(OCT_MAXT, (s * a.tps + 31) / 32 * 32);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(max_per_sm, (227 * 1024) / (smem_for(s) + 1024)), 65536 / (NT * 200)));
    slots = std::min(1023, nsm / 2 * per_sm);
    if (((g.B + s - 1) / s) * n_oct >= 3 * slots) break;
  }
  a.S = S;
  a.n_items = (g.B + S - 1) / S;
  // CTA split: half of the slots per net; octet o gets CTAs in proportion to its FFMA work
  int per_net = std::max(n_oct, slots);
  if (per_net > n_oct * a.n_items) per_net = std::max(n_oct, n_oct * a.n_items);
  static float ovh = -1.f;
  if (ovh < 0.f) ovh = knob_float("OCT_OVH", 24.f);
  float work[OCT_MAX];
  int nc[OCT_MAX];
  for (int o = 0; o < n_oct; ++o) {
    work[o] = ovh;                                             // per-item staging / transform / barrier overhead in the same units
    for (int i = 0; i < g.n_br; ++i)
      if (o < a.br[i].noct) work[o] += 8.f * a.br[i].G + 6.f;  // FFMA per pixel-tap + epilogue share
    nc[o] = 1;
  }
  for (int left = per_net - n_oct; left > 0; --left) {         // greedy: relieve the octet with the longest schedule
    int best = 0;
    float worst = -1.f;
    for (int o = 0; o < n_oct; ++o) {
      const float t = work[o] * (float)((a.n_items + nc[o] - 1) / nc[o]);
      if (t > worst) { worst = t; best = o; }
    }
    if ((a.n_items + nc[best] - 1) / nc[best] <= 1) break;
    ++nc[best];
  }
  int tot = 0;
  for (int o = 0; o < n_oct; ++o) { a.cta_first[o] = (unsigned short)tot; tot += nc[o]; }
  a.cta_first[n_oct] = (unsigned short)tot;
  if (tot > 1023) return CNF_NOT_ELIGIBLE;
  const size_t smem = smem_for(S);
  static int verbose = -1;
  if (verbose < 0) verbose = knob_int("OCT_VERBOSE", 0);
  if (verbose > 0) {
    fprintf(stderr, "[gconv_oct] B=%d %dx%dx%d->%d nbuf=%d S=%d NT=%d items=%d octets=%d ctas/net=%d smem=%zu halo=%d split=%d,%d,..,%d\n", g.B, g.h, g.w, g.Cin,
            g.Cout, a.nbuf, S, NT, a.n_items, n_oct, tot, smem, halo, nc[0], n_oct > 1 ? nc[1] : 0, nc[n_oct - 1]);
  }
  {
    static SmemAttrCache cache;
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    const bool first = dev >= 0 && dev < CNF_MAX_DEVICES && cache.set[dev] == 0;
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)gconv_oct_kernel<OCT_MAXT>, budget, cache));
    // several small CTAs per SM only co-reside if the carve-out is the maximum
    if (first) CU_TRY(cudaFuncSetAttribute(gconv_oct_kernel<OCT_MAXT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  if (verbose > 0) {
    int occ = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gconv_oct_kernel<OCT_MAXT>, NT, smem);
    fprintf(stderr, "[gconv_oct] resident CTAs per SM at this launch shape: %d\n", occ);
  }
  if (verbose > 0) --verbose;
  // TMA tiles: nets stacked contiguously, 16-byte aligned strides, boxes <= 256 per dimension, buffers on 256-byte phases
  static int tma_env = -1;
  if (tma_env < 0) tma_env = knob_int("OCT_TMA", 1) ? 1 : 0;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const int SH = g.h + 2 * halo;
  a.tma = 0;
  if (tma_env && a.nbuf == 2 && g.in_net_stride == (long long)g.B * hw * g.Cin && (((uintptr_t)g.in) & 15) == 0 && SH <= 256 &&
      a.SW <= 256 && (a.SHW % 4) == 0 && ((S * a.SHW) % 8) == 0)
    a.tma = oct_make_tmap(&tm, g.in, 2 * g.B, g.h, g.w, g.Cin, SH, a.SW) ? 1 : 0;
  static int ws_env = -1;
  if (ws_env < 0) ws_env = knob_int("OCT_WS", 1) ? 1 : 0;
  a.ws = (a.tma && a.nbuf == 2 && ws_env && NT == OCT_MAXT) ? 1 : 0;   // producer warps: TMA tiles, one full-size CTA per SM
  gconv_oct_kernel<OCT_MAXT><<<dim3(tot, 2), NT + (a.ws ? OCT_PROD : 0), smem, st>>>(a, tm);
  return (int)cudaGetLastError();
}

}  // namespace cnf
