// Activation-resident coupling layer: the two s/t networks of ONE sample, the coupling law, the decompress scatter
// and the per-sample log-det in ONE launch, with every intermediate tensor of the nets kept in shared memory.
//
// Reference being replaced: coupling_layer.forward_and_Jacobian / backward (conv_cINN_make_model.py M:1258-1394) with
// coupling_function (M:1076-1213), dilated_residual_block (conv_cINN_base_functions.py F:501-627), grouped_convolution
// (F:364-413), add_common_layers (F:330-362) and tanh_scaling_layer (M:97-122).
//
// Why: on the 14x14 / 7x7 (16x16 / 8x8) levels one sample's residual stream is 3-32 KB per net, so the layer-per-kernel
// path (stnet_kernels.cu: 11 launches per coupling layer, activations through HBM/L2) is bound by launch latency and
// pipeline fill, not by work: 35 % of a config-2 step for 7 % of its FLOPs (profiles/r01i_summary.md).  Here a CTA owns
// one sample: it gathers the masked input once per net, runs stem -> R x {LN, 1x1, LN, grouped dilated 3x3, LN, 1x1 +
// residual} -> LN -> head for net b and then net A with X / Y1 / Y2 resident in shared memory (rows padded by 4 floats:
// 8 consecutive pixels x 16 B hit 32 distinct banks for every channel count that is a multiple of 8), the LayerNorm
// sums are CTA reductions (fixed order: results do not depend on the batch size or the position in the batch), gamma /
// beta stream from L2 once per sample (first batch in flight across the reduction barrier), the stage weights are
// prefetched with cp.async behind the normalisation pass of the previous stage, and the head epilogue applies w*tanh,
// exp, the affine law, the scatter and the log-det.  HBM traffic per sample: the masked half of the flow state in, the
// other half in and out.
//
// All FMAs are FFMA2 (fma.rn.f32x2 over output-channel pairs, the activation as the broadcast operand): a three-register
// FFMA issues every other cycle per scheduler on sm_100, the packed form is what reaches the fp32 peak.
// The shapes of BASELINE configs 2 and 3 are compiled in (FzStatic: every index computation folds to constants, the
// grouped-conv variants that a layer does not use are not instantiated); any other eligible layer runs the same code with
// run-time shapes (FzDynamic).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "cnf_internal.h"
#include "device_utils.cuh"

namespace cnf {

constexpr int FZ_MAXR = 6;
constexpr int FZ_MAXBR = 5;

struct FusedRB {
  long long ln1_g, ln1_b, pw1_w, pw1_b, ln2_g, ln2_b, ln3_g, ln3_b, pw2_w, pw2_b;
  long long gw[FZ_MAXBR], gb[FZ_MAXBR];
};

struct FusedBranch {
  int dil, g, ch, out_off;      // dilation, group width (gin = gout), channels of the branch, offset in the concat
  int nseg, seglen, ncls;       // row classes (rows congruent mod dil) cut into nseg segments of <= seglen rows
  int ntpc, wbpc, per;          // lane tasks per 4-channel chunk, warp blocks per chunk, lane tasks per warp block
  int sw, sb;                   // offsets (floats) of the branch's weights / biases inside the smem weight buffer
  int wt0, nwt;                 // first warp task of this branch, number of warp tasks
};

// geometry of a branch given everything before it (wt0: warp tasks so far, sw0: weight floats so far)
__host__ __device__ inline FusedBranch fz_branch_geom(int h, int w, int nr, int dil, int g, int ch, int out_off, int wt0,
                                                      int sw0) {
  FusedBranch f = {};
  f.dil = dil; f.g = g; f.ch = ch; f.out_off = out_off;
  f.ncls = dil < h ? dil : h;
  const int rows = (h + dil - 1) / dil;
  f.nseg = (rows + nr - 1) / nr;
  f.seglen = (rows + f.nseg - 1) / f.nseg;
  f.ntpc = f.ncls * f.nseg * w;
  f.wbpc = (f.ntpc + 31) / 32;
  f.per = (f.ntpc + f.wbpc - 1) / f.wbpc;
  f.wt0 = wt0;
  f.nwt = (ch / 4) * f.wbpc;
  f.sw = sw0;
  f.sb = sw0 + 9 * g * ch;
  return f;
}

struct FusedArgs {
  FlowView in_view, out_view;
  const float* params;
  long long net_stride;
  long long stem_w, stem_b, lnf_g, lnf_b, head_w, head_b, tanh_off;
  FusedRB rb[FZ_MAXR];
  FusedBranch br[FZ_MAXBR];
  int n_br, n_wt;               // branches, warp tasks of the grouped-conv stage
  int B, h, w, hw, nk, cat, c1, c2, R, in_mask, mask_c, mode;
  int ldx, ldy2;                // padded row strides (floats) of X / Y1 and of Y2
  int off_y1, off_y2, off_w;    // smem offsets (floats)
  int pw_variant, nr_variant, c2t;
  int lognl;                    // log2(nk / 4): output quads per warp in the GEMM-shaped stages
  unsigned div_w_m, div_qnk_m, div_qcat_m, div_c1_m, div_sw_m;   // ceil(2^32 / d): n / d == umulhi(n, m), n * d < 2^32
  double* logdet;
  float* tscratch;              // [B][hw][c2]: t = net b output, parked while net A runs
};

// ---- shape policies --------------------------------------------------------------------------------------------
struct FzDynamic {
  static constexpr bool kStatic = false;
};

// dilations are 1, 2, 4 (the planner's list for k = 3, M:1553-1617); G0..G2 = group widths of the branches (0 = absent)
template <int H_, int W_, int NK_, int C1_, int C2_, int G0_, int G1_, int G2_, int NT_, int PXT_, int NR_>
struct FzStatic {
  static constexpr bool kStatic = true;
  static constexpr int H = H_, W = W_, HW = H_ * W_, NK = NK_, C1 = C1_, C2 = C2_;
  static constexpr int NBR = (G0_ > 0) + (G1_ > 0) + (G2_ > 0);
  static constexpr int CAT = NK_ + (G1_ > 0 ? NK_ / 2 : 0) + (G2_ > 0 ? NK_ / 4 : 0);
  static constexpr int LDX = NK_ + 4, LDY2 = (CAT > NK_ ? CAT : NK_) + 4;
  static constexpr int NT = NT_, PXT = PXT_, NR = NR_;
  static constexpr int LOGNL = NK_ == 8 ? 1 : NK_ == 16 ? 2 : NK_ == 32 ? 3 : 4;   // log2(NK / 4)
  static constexpr int C2T = C2_ <= 2 ? 2 : C2_ <= 4 ? 4 : 8;
  __host__ __device__ static constexpr int g(int i) { return i == 0 ? G0_ : i == 1 ? G1_ : G2_; }
};

template <class S>
struct Dm {   // dimension accessors: constants for FzStatic, kernel arguments for FzDynamic
  const FusedArgs& a;
  __device__ __forceinline__ int h() const { if constexpr (S::kStatic) return S::H; else return a.h; }
  __device__ __forceinline__ int w() const { if constexpr (S::kStatic) return S::W; else return a.w; }
  __device__ __forceinline__ int hw() const { if constexpr (S::kStatic) return S::HW; else return a.hw; }
  __device__ __forceinline__ int nk() const { if constexpr (S::kStatic) return S::NK; else return a.nk; }
  __device__ __forceinline__ int cat() const { if constexpr (S::kStatic) return S::CAT; else return a.cat; }
  __device__ __forceinline__ int c1() const { if constexpr (S::kStatic) return S::C1; else return a.c1; }
  __device__ __forceinline__ int c2() const { if constexpr (S::kStatic) return S::C2; else return a.c2; }
  __device__ __forceinline__ int ldx() const { if constexpr (S::kStatic) return S::LDX; else return a.ldx; }
  __device__ __forceinline__ int ldy2() const { if constexpr (S::kStatic) return S::LDY2; else return a.ldy2; }
  __device__ __forceinline__ int nt() const { if constexpr (S::kStatic) return S::NT; else return (int)blockDim.x; }
  __device__ __forceinline__ int n_br() const { if constexpr (S::kStatic) return S::NBR; else return a.n_br; }
  __device__ __forceinline__ int lognl() const { if constexpr (S::kStatic) return S::LOGNL; else return a.lognl; }
  // divisions by shape constants (n >= 0)
  __device__ __forceinline__ int div_w(int n) const { if constexpr (S::kStatic) return n / S::W; else return (int)__umulhi((unsigned)n, a.div_w_m); }
  __device__ __forceinline__ int div_qnk(int n) const { if constexpr (S::kStatic) return n / (S::NK / 4); else return (int)__umulhi((unsigned)n, a.div_qnk_m); }
  __device__ __forceinline__ int div_qcat(int n) const { if constexpr (S::kStatic) return n / (S::CAT / 4); else return (int)__umulhi((unsigned)n, a.div_qcat_m); }
  __device__ __forceinline__ int div_c1(int n) const { if constexpr (S::kStatic) return n / S::C1; else return (a.div_c1_m ? (int)__umulhi((unsigned)n, a.div_c1_m) : n); }
  __device__ __forceinline__ int div_sw(int n) const { if constexpr (S::kStatic) return n / (S::W + 2); else return (int)__umulhi((unsigned)n, a.div_sw_m); }
  __device__ __forceinline__ FusedBranch br(int i) const {
    if constexpr (S::kStatic) {
      int wt0 = 0, sw0 = 0, off = 0;
      FusedBranch f = {};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j < S::NBR && j <= i) {
          f = fz_branch_geom(S::H, S::W, S::NR, 1 << j, S::g(j), S::NK >> j, off, wt0, sw0);
          wt0 += f.nwt; sw0 = f.sb + f.ch; off += f.ch;
        }
      }
      return f;
    } else {
      return a.br[i];
    }
  }
  __device__ __forceinline__ int n_wt() const {
    if constexpr (S::kStatic) { const FusedBranch f = br(S::NBR - 1); return f.wt0 + f.nwt; }
    else return a.n_wt;
  }
};

__device__ __forceinline__ unsigned fz_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fz_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fz_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fz_cp_commit_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// global -> smem copy of n floats (src 16-byte aligned, dst 16-byte aligned); tail (n % 4) with plain loads
__device__ __forceinline__ void fz_stage(float* dst, const float* __restrict__ src, int n, int tid, int nt) {
  const int n4 = n >> 2;
  for (int i = tid; i < n4; i += nt) fz_cp16(dst + 4 * i, src + 4 * i);
  for (int i = (n4 << 2) + tid; i < n; i += nt) dst[i] = src[i];
}

// d.xy += a.xy * b.xy: one FFMA2
__device__ __forceinline__ void fz_ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void fz_bfma4(float2& d0, float2& d1, const float s, const float4 w) {   // (d0, d1) += s * w
  const float2 ss = make_float2(s, s);
  fz_ffma2(d0, ss, make_float2(w.x, w.y));
  fz_ffma2(d1, ss, make_float2(w.z, w.w));
}

__device__ __forceinline__ void fz_stat2(const float2 o, float& s1, float& s2) {
  float l;
  l = fmaxf(o.x, CNF_LRELU_SLOPE * o.x); s1 += l; s2 = fmaf(l, l, s2);
  l = fmaxf(o.y, CNF_LRELU_SLOPE * o.y); s1 += l; s2 = fmaf(l, l, s2);
}

// --------------------------------------------------------------------------------------------------------------
// GEMM-shaped stages.  out[p][n] = bias[n] + sum_k A(p, k) W[k][n] (+ res[p][n]).
// A warp owns a tile of (PL * PXT) pixels x all N output channels, as an outer product over its lanes: lane = pgl * NL +
// nbl with NL = N / 4 output quads and PL = 32 / NL pixel lanes; a thread owns the pixels tile0 + pgl + PL * j (j < PXT)
// x the 4 outputs of quad nbl.  Shared memory serves the DISTINCT words of a warp-wide load (measured,
// tools/microbench/lds_patterns.cu: 128 B of distinct 16-byte words per cycle whatever lanes ask for them), so an A load
// costs PL x 16 B and a W load NL x 16 B: (PXT + 4) wavefronts per 4 k for 4 PXT FFMA2 per lane.
//   STEM = false: A(p, k) = Ain[p * lda + k]           (1x1 conv, K = channels)
//   STEM = true : A(p, (tap, ci)) = Ain[((y + ky) * (w + 2) + x + kx) * c1 + ci]   (3x3 conv on the zero-haloed input)
// --------------------------------------------------------------------------------------------------------------
template <class S, int PXT, bool STEM>
__device__ __forceinline__ void fz_gemm_stage(const FusedArgs& a, const float* __restrict__ Ain, int lda, int K,
                                              const float* __restrict__ Ws, const float* __restrict__ bs,
                                              float* __restrict__ Out, const float* __restrict__ Res, float& s1, float& s2) {
  const Dm<S> d{a};
  const int N = d.nk(), ldo = d.ldx(), hw = d.hw();
  const int lognl = d.lognl(), NL = 1 << lognl, PL = 32 >> lognl;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = d.nt() >> 5;
  const int nbl = lane & (NL - 1), pgl = lane >> lognl;
  const int tile_px = PL * PXT;
  const int ntiles = (hw + tile_px - 1) / tile_px;
  const float* wcol = Ws + 4 * nbl;
  const float4 bv = ld4(bs + 4 * nbl);
  for (int tile = wid; tile < ntiles; tile += nw) {
    const int p0 = tile * tile_px + pgl;
    int pbase[PXT];
    float2 acc[PXT][2];
#pragma unroll
    for (int j = 0; j < PXT; ++j) {
      const int p = p0 + PL * j;
      const int pc = p < hw ? p : hw - 1;
      if (STEM) {
        const int y = d.div_w(pc), x = pc - y * d.w();
        pbase[j] = (y * (d.w() + 2) + x) * d.c1();
      } else {
        pbase[j] = pc * lda;
      }
      acc[j][0] = make_float2(bv.x, bv.y);
      acc[j][1] = make_float2(bv.z, bv.w);
    }
    if (STEM) {
      const int c1 = d.c1(), SW = d.w() + 2;
      const int vec = (c1 & 3) == 0 ? 4 : (c1 & 1) == 0 ? 2 : 1;
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - 3 * ky;
        const float* ap = Ain + (ky * SW + kx) * c1;
        const float* wp = wcol + tap * c1 * N;
        if (vec == 4) {
#pragma unroll 1
          for (int ci = 0; ci < c1; ci += 4) {
            const float4 w0 = ld4(wp + ci * N), w1 = ld4(wp + (ci + 1) * N), w2 = ld4(wp + (ci + 2) * N), w3 = ld4(wp + (ci + 3) * N);
#pragma unroll
            for (int j = 0; j < PXT; ++j) {
              const float4 xv = ld4(ap + pbase[j] + ci);
              fz_bfma4(acc[j][0], acc[j][1], xv.x, w0);
              fz_bfma4(acc[j][0], acc[j][1], xv.y, w1);
              fz_bfma4(acc[j][0], acc[j][1], xv.z, w2);
              fz_bfma4(acc[j][0], acc[j][1], xv.w, w3);
            }
          }
        } else if (vec == 2) {
#pragma unroll 1
          for (int ci = 0; ci < c1; ci += 2) {
            const float4 w0 = ld4(wp + ci * N), w1 = ld4(wp + (ci + 1) * N);
#pragma unroll
            for (int j = 0; j < PXT; ++j) {
              const float2 xv = *reinterpret_cast<const float2*>(ap + pbase[j] + ci);
              fz_bfma4(acc[j][0], acc[j][1], xv.x, w0);
              fz_bfma4(acc[j][0], acc[j][1], xv.y, w1);
            }
          }
        } else {
#pragma unroll 1
          for (int ci = 0; ci < c1; ++ci) {
            const float4 w0 = ld4(wp + ci * N);
#pragma unroll
            for (int j = 0; j < PXT; ++j) fz_bfma4(acc[j][0], acc[j][1], ap[pbase[j] + ci], w0);
          }
        }
      }
    } else {
#pragma unroll 2
      for (int k0 = 0; k0 < K; k0 += 4) {
        float4 av[PXT];
#pragma unroll
        for (int j = 0; j < PXT; ++j) av[j] = ld4(Ain + pbase[j] + k0);
        const float4 w0 = ld4(wcol + k0 * N), w1 = ld4(wcol + (k0 + 1) * N), w2 = ld4(wcol + (k0 + 2) * N), w3 = ld4(wcol + (k0 + 3) * N);
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
          fz_bfma4(acc[j][0], acc[j][1], av[j].x, w0);
          fz_bfma4(acc[j][0], acc[j][1], av[j].y, w1);
          fz_bfma4(acc[j][0], acc[j][1], av[j].z, w2);
          fz_bfma4(acc[j][0], acc[j][1], av[j].w, w3);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PXT; ++j) {
      const int p = p0 + PL * j;
      if (p >= hw) continue;
      float4 v = make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y);
      if (Res) {
        const float4 r = ld4(Res + p * ldo + 4 * nbl);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      st4(Out + p * ldo + 4 * nbl, v);
      fz_stat2(make_float2(v.x, v.y), s1, s2);
      fz_stat2(make_float2(v.z, v.w), s1, s2);
    }
  }
}

template <class S, bool STEM>
__device__ __forceinline__ void fz_gemm_dispatch(const FusedArgs& a, const float* Ain, int lda, int K, const float* Ws,
                                                 const float* bs, float* Out, const float* Res, float& s1, float& s2) {
  if constexpr (S::kStatic) {
    fz_gemm_stage<S, S::PXT, STEM>(a, Ain, lda, K, Ws, bs, Out, Res, s1, s2);
  } else {
    switch (a.pw_variant) {
      case 0: fz_gemm_stage<S, 7, STEM>(a, Ain, lda, K, Ws, bs, Out, Res, s1, s2); break;
      case 1: fz_gemm_stage<S, 4, STEM>(a, Ain, lda, K, Ws, bs, Out, Res, s1, s2); break;
      case 2: fz_gemm_stage<S, 2, STEM>(a, Ain, lda, K, Ws, bs, Out, Res, s1, s2); break;
      default: fz_gemm_stage<S, 1, STEM>(a, Ain, lda, K, Ws, bs, Out, Res, s1, s2); break;
    }
  }
}

// --------------------------------------------------------------------------------------------------------------
// Grouped dilated 3x3 convs (F:389-411, F:577-590).  A lane owns 4 consecutive output channels of one branch (one
// group for G = 4, half a group for G = 8, two groups for G = 2, four for G = 1) on a column segment of <= NR output
// rows spaced `dil` apart: per kx it loads the NR + 2 input rows once and reuses them for the 3 ky taps.  The lanes of
// a warp share the chunk (warp-uniform weight loads) and walk consecutive x.
// --------------------------------------------------------------------------------------------------------------
template <class S, int G, int NR>
__device__ __forceinline__ void fz_gconv_task(const FusedArgs& a, const FusedBranch& br, const float* __restrict__ In,
                                              const float* __restrict__ Wb, float* __restrict__ Out, int chunk, int x,
                                              int row0, int nrows, float& s1, float& s2) {
  constexpr int NIN = G == 8 ? 2 : 1;
  const Dm<S> d{a};
  const int dil = br.dil, ldi = d.ldx(), ldo = d.ldy2(), h = d.h(), w = d.w();
  const float* Wbr = Wb + br.sw;
  const int cbase = G == 8 ? (chunk >> 1) * 8 : chunk * 4;
  float2 acc[NR][2];
  {
    const float4 bv = ld4(Wb + br.sb + chunk * 4);
#pragma unroll
    for (int j = 0; j < NR; ++j) { acc[j][0] = make_float2(bv.x, bv.y); acc[j][1] = make_float2(bv.z, bv.w); }
  }
  const int rstep = dil * w * ldi;
#pragma unroll 1
  for (int kx = 0; kx < 3; ++kx) {
    const int ix = x + (kx - 1) * dil;
    if (ix < 0 || ix >= w) continue;
    const float* col = In + ((row0 - dil) * w + ix) * ldi + cbase;   // input row j of the segment at col + j * rstep
#pragma unroll
    for (int half = 0; half < NIN; ++half) {
      float4 in[NR + 2];
#pragma unroll
      for (int j = 0; j < NR + 2; ++j) {
        const int iy = row0 + (j - 1) * dil;
        in[j] = (iy >= 0 && iy < h) ? ld4(col + j * rstep + 4 * half) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int tap = ky * 3 + kx;
        if constexpr (G == 8 || G == 4) {
          // W[group][tap][ci][co]: rows ci = 4 * half .. + 3 of the group, the 4 columns of this chunk
          const float* wp = Wbr + (((G == 8 ? (chunk >> 1) : chunk) * 9 + tap) * G + 4 * half) * G + (G == 8 ? (chunk & 1) * 4 : 0);
          const float4 w0 = ld4(wp), w1 = ld4(wp + G), w2 = ld4(wp + 2 * G), w3 = ld4(wp + 3 * G);
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            fz_bfma4(acc[j][0], acc[j][1], in[j + ky].x, w0);
            fz_bfma4(acc[j][0], acc[j][1], in[j + ky].y, w1);
            fz_bfma4(acc[j][0], acc[j][1], in[j + ky].z, w2);
            fz_bfma4(acc[j][0], acc[j][1], in[j + ky].w, w3);
          }
        } else if constexpr (G == 2) {
          // groups 2 * chunk and 2 * chunk + 1; W[g][tap] = (ci0co0, ci0co1, ci1co0, ci1co1)
          const float4 wa = ld4(Wbr + ((2 * chunk) * 9 + tap) * 4), wb = ld4(Wbr + ((2 * chunk + 1) * 9 + tap) * 4);
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            const float4 v = in[j + ky];
            fz_ffma2(acc[j][0], make_float2(v.x, v.x), make_float2(wa.x, wa.y));
            fz_ffma2(acc[j][0], make_float2(v.y, v.y), make_float2(wa.z, wa.w));
            fz_ffma2(acc[j][1], make_float2(v.z, v.z), make_float2(wb.x, wb.y));
            fz_ffma2(acc[j][1], make_float2(v.w, v.w), make_float2(wb.z, wb.w));
          }
        } else {
          // depthwise: groups 4 * chunk .. + 3, W[g][tap] scalar
          const float2 w01 = make_float2(Wbr[(4 * chunk) * 9 + tap], Wbr[(4 * chunk + 1) * 9 + tap]);
          const float2 w23 = make_float2(Wbr[(4 * chunk + 2) * 9 + tap], Wbr[(4 * chunk + 3) * 9 + tap]);
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            const float4 v = in[j + ky];
            fz_ffma2(acc[j][0], make_float2(v.x, v.y), w01);
            fz_ffma2(acc[j][1], make_float2(v.z, v.w), w23);
          }
        }
      }
    }
  }
  float* orow = Out + (row0 * w + x) * ldo + br.out_off + chunk * 4;
  const int ostep = dil * w * ldo;
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    if (j < nrows) {
      st4(orow + j * ostep, make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y));
      fz_stat2(acc[j][0], s1, s2);
      fz_stat2(acc[j][1], s1, s2);
    }
  }
}

template <class S, int NR>
__device__ __forceinline__ void fz_gconv_branch(const FusedArgs& a, const FusedBranch& br, int rem, const float* In,
                                                const float* Wb, float* Out, float& s1, float& s2) {
  const Dm<S> d{a};
  const int lane = threadIdx.x & 31;
  const int chunk = rem / br.wbpc, blk = rem - chunk * br.wbpc;
  const int id = blk * br.per + lane;
  if (lane >= br.per || id >= br.ntpc) return;
  const int cs = d.div_w(id), x = id - cs * d.w();
  const int cls = cs / br.nseg, seg = cs - cls * br.nseg;
  const int nrows_c = (d.h() - cls + br.dil - 1) / br.dil;
  const int nrows = min(br.seglen, nrows_c - seg * br.seglen);
  if (nrows <= 0) return;
  const int row0 = cls + seg * br.seglen * br.dil;
  switch (br.g) {
    case 8: fz_gconv_task<S, 8, NR>(a, br, In, Wb, Out, chunk, x, row0, nrows, s1, s2); break;
    case 4: fz_gconv_task<S, 4, NR>(a, br, In, Wb, Out, chunk, x, row0, nrows, s1, s2); break;
    case 2: fz_gconv_task<S, 2, NR>(a, br, In, Wb, Out, chunk, x, row0, nrows, s1, s2); break;
    default: fz_gconv_task<S, 1, NR>(a, br, In, Wb, Out, chunk, x, row0, nrows, s1, s2); break;
  }
}

template <class S, int NR>
__device__ __forceinline__ void fz_gconv_stage(const FusedArgs& a, const float* __restrict__ In, const float* __restrict__ Wb,
                                               float* __restrict__ Out, float& s1, float& s2) {
  const Dm<S> d{a};
  const int wid = threadIdx.x >> 5, nw = d.nt() >> 5;
  const int n_wt = d.n_wt();
  for (int wt = wid; wt < n_wt; wt += nw) {
    if constexpr (S::kStatic) {
#pragma unroll
      for (int i = 0; i < S::NBR; ++i) {
        const FusedBranch br = d.br(i);   // folds to constants
        if (wt >= br.wt0 && wt < br.wt0 + br.nwt) fz_gconv_branch<S, NR>(a, br, wt - br.wt0, In, Wb, Out, s1, s2);
      }
    } else {
      int bi = 0;
#pragma unroll
      for (int i = 1; i < FZ_MAXBR; ++i)
        if (i < a.n_br && wt >= a.br[i].wt0) bi = i;
      fz_gconv_branch<S, NR>(a, a.br[bi], wt - a.br[bi].wt0, In, Wb, Out, s1, s2);
    }
  }
}

// --------------------------------------------------------------------------------------------------------------
// LayerNorm (F:350-360 -> keras LayerNormalization: biased variance over the whole sample, eps 1e-3).
// --------------------------------------------------------------------------------------------------------------
struct FzRed {
  double part[2][32];
  float mean, rstd;
  float ldpart[32];
};

constexpr int FZ_GB = 4;   // gamma/beta quads in flight per thread
struct FzGB { float4 g[FZ_GB], b[FZ_GB]; };

__device__ __forceinline__ void fz_gb_load(FzGB& r, const float* __restrict__ gam, const float* __restrict__ bet, int i0,
                                           int nt, int n4) {
#pragma unroll
  for (int u = 0; u < FZ_GB; ++u) {
    const int i = min(i0 + u * nt, n4 - 1);   // clamped: the loads are unconditional, the stores are not
    r.g[u] = __ldg(reinterpret_cast<const float4*>(gam) + i);
    r.b[u] = __ldg(reinterpret_cast<const float4*>(bet) + i);
  }
}

// Statistics of the stage that just finished (CTA reduction in a fixed order; mean and centred variance in fp64), then
// dst[p][c] = LN(LReLU(src[p][c])) * gamma + beta over the hw x C tensor (rows padded to lds / ldd floats; gamma / beta
// are the flat [hw * C] keras vectors).  `between` runs after the first barrier (the weight prefetch of the next
// stage); the first gamma / beta batch is in flight across the second barrier.
template <class S, bool CAT, typename F>
__device__ __forceinline__ void fz_norm(const FusedArgs& a, float s1, float s2, FzRed& red, const float* src, int lds,
                                        float* dst, int ldd, const float* __restrict__ gam, const float* __restrict__ bet,
                                        F&& between) {
  const Dm<S> d{a};
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nt = d.nt(), nw = nt >> 5;
  const int C = CAT ? d.cat() : d.nk();
  const int Q = C >> 2, n4 = d.hw() * Q;
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    red.part[0][wid] = (double)s1;
    red.part[1][wid] = (double)s2;
  }
  __syncthreads();
  between();
  FzGB gb;
  fz_gb_load(gb, gam, bet, tid, nt, n4);
  if (wid == 0) {
    double d1 = lane < nw ? red.part[0][lane] : 0.0, d2 = lane < nw ? red.part[1][lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if (lane == 0) {
      const double n = (double)d.hw() * (double)C;
      const double mean = d1 / n;
      double var = d2 / n - mean * mean;
      var = var > 0.0 ? var : 0.0;
      red.mean = (float)mean;
      red.rstd = (float)(1.0 / sqrt(var + (double)CNF_LN_EPS));
    }
  }
  __syncthreads();
  const float mean = red.mean, rstd = red.rstd;
  for (int i0 = tid; i0 < n4; i0 += FZ_GB * nt) {
#pragma unroll
    for (int u = 0; u < FZ_GB; ++u) {
      const int i = i0 + u * nt;
      if (i < n4) {
        const int p = CAT ? d.div_qcat(i) : d.div_qnk(i);
        const int c = (i - p * Q) << 2;
        float4 v = ld4(src + p * lds + c);
        v.x = fmaxf(v.x, CNF_LRELU_SLOPE * v.x); v.y = fmaxf(v.y, CNF_LRELU_SLOPE * v.y);
        v.z = fmaxf(v.z, CNF_LRELU_SLOPE * v.z); v.w = fmaxf(v.w, CNF_LRELU_SLOPE * v.w);
        v.x = fmaf((v.x - mean) * rstd, gb.g[u].x, gb.b[u].x);
        v.y = fmaf((v.y - mean) * rstd, gb.g[u].y, gb.b[u].y);
        v.z = fmaf((v.z - mean) * rstd, gb.g[u].z, gb.b[u].z);
        v.w = fmaf((v.w - mean) * rstd, gb.g[u].w, gb.b[u].w);
        st4(dst + p * ldd + c, v);
      }
    }
    if (i0 + FZ_GB * nt < n4) fz_gb_load(gb, gam, bet, i0 + FZ_GB * nt, nt, n4);
  }
}

// --------------------------------------------------------------------------------------------------------------
// Head 3x3 conv (nk -> c2) on the normalised residual stream + the layer's epilogue (M:1133-1150, M:1198, M:1307-1326,
// M:1379-1394).  One thread per output pixel; the weights are staged as [tap][k][C2T] (zero-padded to C2T columns).
// --------------------------------------------------------------------------------------------------------------
template <class S, int C2T>
__device__ __forceinline__ void fz_head_stage(const FusedArgs& a, const float* __restrict__ Xt, const float* __restrict__ Wh,
                                              const float* __restrict__ bh, int net, int b, float tanh_w, float& ld) {
  const Dm<S> d{a};
  const int ldx = d.ldx(), hw = d.hw(), h = d.h(), w = d.w(), nk = d.nk(), c2 = d.c2();
  for (int p = threadIdx.x; p < hw; p += d.nt()) {
    const int y = d.div_w(p), x = p - y * w;
    float2 acc[C2T / 2];
#pragma unroll
    for (int c = 0; c < C2T / 2; ++c) acc[c] = make_float2(bh[2 * c], bh[2 * c + 1]);
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y + ky - 1;
      if (iy < 0 || iy >= h) continue;
#pragma unroll 1
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x + kx - 1;
        if (ix < 0 || ix >= w) continue;
        const float* row = Xt + (iy * w + ix) * ldx;
        const float* wt = Wh + (ky * 3 + kx) * nk * C2T;
#pragma unroll 2
        for (int k0 = 0; k0 < nk; k0 += 4) {
          const float4 v = ld4(row + k0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float xv = kk == 0 ? v.x : kk == 1 ? v.y : kk == 2 ? v.z : v.w;
            const float* wr = wt + (k0 + kk) * C2T;
            if constexpr (C2T >= 4) {
#pragma unroll
              for (int q = 0; q < C2T / 4; ++q) fz_bfma4(acc[2 * q], acc[2 * q + 1], xv, ld4(wr + 4 * q));
            } else {
              fz_ffma2(acc[0], make_float2(xv, xv), *reinterpret_cast<const float2*>(wr));
            }
          }
        }
      }
    }
    float* tp = a.tscratch + ((long long)b * hw + p) * c2;
#pragma unroll
    for (int co = 0; co < C2T; ++co) {
      if (co < c2) {
        const float raw = (co & 1) ? acc[co >> 1].y : acc[co >> 1].x;
        if (net == 1) {
          tp[co] = raw;
        } else {
          const float A = tanh_w * tanhf(raw);                              // M:1198, M:114-116
          const float t = tp[co];
          float* ptr = a.out_view.base + comp_off(a.out_view, a.mask_c, b, y, x, co);
          const float u2 = *ptr;
          if (a.mode == HEAD_FWD) {
            *ptr = __fadd_rn(__fmul_rn(expf(A), u2), t);                    // M:1307, M:1230-1231
            ld += A;                                                        // M:1323
          } else {
            *ptr = __fmul_rn(__frcp_rn(expf(A)), __fsub_rn(u2, t));         // M:1379, M:1250-1251
          }
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------------------
// The kernel: one CTA per sample.
// --------------------------------------------------------------------------------------------------------------
template <class S>
__device__ __forceinline__ void fz_body(const FusedArgs& a, float* fz_smem, FzRed& red) {
  const Dm<S> d{a};
  float* X = fz_smem;
  float* Y1 = fz_smem + a.off_y1;
  float* Y2 = fz_smem + a.off_y2;
  float* Wb = fz_smem + a.off_w;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, nt = d.nt();
  const int hw = d.hw(), nk = d.nk(), cat = d.cat(), ldx = d.ldx(), ldy2 = d.ldy2(), c1 = d.c1(), c2 = d.c2();
  float ld = 0.f;

#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int net = 1 - pass;   // net b first: its output t is parked while net A runs
    const float* P = a.params + (long long)net * a.net_stride;
    // ---- masked gather of u1 (M:723-759) with a zero halo -> Y1 region; stem weights -> Wb
    {
      const int SW = d.w() + 2, SH = d.h() + 2;
      const int n_in = SH * SW * c1;
      for (int idx = tid; idx < n_in; idx += nt) {
        const int pix = d.div_c1(idx), ci = idx - pix * c1;
        const int sy = d.div_sw(pix), sx = pix - sy * SW;
        const int iy = sy - 1, ix = sx - 1;
        float v = 0.f;
        if (iy >= 0 && iy < d.h() && ix >= 0 && ix < d.w()) v = a.in_view.base[comp_off(a.in_view, a.in_mask, b, iy, ix, ci)];
        Y1[idx] = v;
      }
      fz_stage(Wb, P + a.stem_w, 9 * c1 * nk, tid, nt);
      fz_stage(Wb + 9 * c1 * nk, P + a.stem_b, nk, tid, nt);
      fz_cp_commit_wait();
      __syncthreads();
    }
    float s1 = 0.f, s2 = 0.f;
    fz_gemm_dispatch<S, true>(a, Y1, 0, 9 * c1, Wb, Wb + 9 * c1 * nk, X, nullptr, s1, s2);

#pragma unroll 1
    for (int r = 0; r < a.R; ++r) {
      const FusedRB& L = a.rb[r];
      // ---- a = LN1(LReLU(X)) -> Y2 region (row stride ldx); pw1 weights behind it
      fz_norm<S, false>(a, s1, s2, red, X, ldx, Y2, ldx, P + L.ln1_g, P + L.ln1_b, [&] {
        fz_stage(Wb, P + L.pw1_w, nk * nk, tid, nt);
        fz_stage(Wb + nk * nk, P + L.pw1_b, nk, tid, nt);
      });
      fz_cp_commit_wait();
      __syncthreads();
      s1 = 0.f; s2 = 0.f;
      fz_gemm_dispatch<S, false>(a, Y2, ldx, nk, Wb, Wb + nk * nk, Y1, nullptr, s1, s2);
      // ---- Y1 <- LN2(LReLU(Y1)) in place; grouped-conv weights behind it
      fz_norm<S, false>(a, s1, s2, red, Y1, ldx, Y1, ldx, P + L.ln2_g, P + L.ln2_b, [&] {
#pragma unroll
        for (int i = 0; i < (S::kStatic ? 3 : FZ_MAXBR); ++i) {
          if (i < d.n_br()) {
            const FusedBranch br = d.br(i);
            fz_stage(Wb + br.sw, P + L.gw[i], 9 * br.g * br.ch, tid, nt);
            fz_stage(Wb + br.sb, P + L.gb[i], br.ch, tid, nt);
          }
        }
      });
      fz_cp_commit_wait();
      __syncthreads();
      s1 = 0.f; s2 = 0.f;
      if constexpr (S::kStatic) {
        fz_gconv_stage<S, S::NR>(a, Y1, Wb, Y2, s1, s2);
      } else {
        if (a.nr_variant == 7) fz_gconv_stage<S, 7>(a, Y1, Wb, Y2, s1, s2);
        else fz_gconv_stage<S, 8>(a, Y1, Wb, Y2, s1, s2);
      }
      // ---- Y2 <- LN3(LReLU(Y2)) in place; pw2 weights behind it
      fz_norm<S, true>(a, s1, s2, red, Y2, ldy2, Y2, ldy2, P + L.ln3_g, P + L.ln3_b, [&] {
        fz_stage(Wb, P + L.pw2_w, cat * nk, tid, nt);
        fz_stage(Wb + cat * nk, P + L.pw2_b, nk, tid, nt);
      });
      fz_cp_commit_wait();
      __syncthreads();
      s1 = 0.f; s2 = 0.f;
      fz_gemm_dispatch<S, false>(a, Y2, ldy2, cat, Wb, Wb + cat * nk, X, X, s1, s2);   // + residual (F:625)
    }
    // ---- X <- LNf(LReLU(X)) in place; head weights [tap][k][C2T] (zero-padded columns) behind it
    int c2t;
    if constexpr (S::kStatic) c2t = S::C2T; else c2t = a.c2t;
    fz_norm<S, false>(a, s1, s2, red, X, ldx, X, ldx, P + a.lnf_g, P + a.lnf_b, [&] {
      const int nwh = 9 * nk * c2t;
      const float* Wg = P + a.head_w;
      for (int i = tid; i < nwh; i += nt) {
        const int row = i / c2t, co = i - row * c2t;
        Wb[i] = co < c2 ? Wg[row * c2 + co] : 0.f;
      }
      if (tid < c2t) Wb[nwh + tid] = tid < c2 ? P[a.head_b + tid] : 0.f;
    });
    __syncthreads();
    const float tanh_w = a.params[a.tanh_off];
    const float* Wh = Wb;
    const float* bh = Wb + 9 * nk * c2t;
    if constexpr (S::kStatic) {
      fz_head_stage<S, S::C2T>(a, X, Wh, bh, net, b, tanh_w, ld);
    } else {
      if (c2t == 2) fz_head_stage<S, 2>(a, X, Wh, bh, net, b, tanh_w, ld);
      else if (c2t == 4) fz_head_stage<S, 4>(a, X, Wh, bh, net, b, tanh_w, ld);
      else fz_head_stage<S, 8>(a, X, Wh, bh, net, b, tanh_w, ld);
    }
    __syncthreads();   // the buffers are reused by the next net
  }
  if (a.mode == HEAD_FWD && a.logdet) {
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    ld = warp_sum(ld);
    if (lane == 0) red.ldpart[wid] = ld;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < nw; ++i) s += (double)red.ldpart[i];
      atomicAdd(a.logdet + b, s);
    }
  }
}

template <class S, int NTB, int MINB>
__global__ void __launch_bounds__(NTB, MINB) fused_coupling_kernel(const FusedArgs a) {
  extern __shared__ __align__(16) float fz_smem[];
  __shared__ FzRed red;
  fz_body<S>(a, fz_smem, red);
}

// --------------------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------------------
static unsigned recip32(unsigned d) { return d <= 1 ? 0u : (unsigned)((0x100000000ULL + d - 1) / d); }   // 0: divisor 1

constexpr size_t kFzStatic = 1024;   // static shared memory of the kernel (FzRed), rounded up

template <class S, int NTB, int MINB>
static int fz_launch(const FusedArgs& a, int B, int nt, size_t smem, cudaStream_t st) {
  static SmemAttrCache cache;   // per instantiation, per device
  const int e = ensure_dynamic_smem((const void*)fused_coupling_kernel<S, NTB, MINB>, 227 * 1024 - kFzStatic, cache);
  if (e) return e;
  fused_coupling_kernel<S, NTB, MINB><<<B, nt, smem, st>>>(a);
  return (int)cudaGetLastError();
}

// compiled-in shapes: H, W, NK, C1, C2, G0, G1, G2, NT, PXT, NR
using FzCfg2A = FzStatic<14, 14, 32, 4, 4, 4, 2, 1, 256, 7, 7>;    // config 2 level 0 checkerboard: 7 warp tiles of 28 pixels
using FzCfg2B = FzStatic<14, 14, 32, 2, 2, 8, 4, 0, 256, 7, 7>;    // config 2 level 1 channel
using FzCfg2C = FzStatic<7, 7, 16, 8, 8, 4, 2, 0, 128, 2, 7>;      // config 2 level 1 checkerboard: 4 warp tiles of 16 pixels
using FzCfg3A = FzStatic<16, 16, 32, 8, 8, 4, 2, 1, 512, 4, 8>;    // config 3 level 0 checkerboard: 16 warp tiles of 16 pixels
using FzCfg3B = FzStatic<16, 16, 32, 4, 4, 8, 4, 0, 512, 4, 8>;    // config 3 level 1 channel

template <class S>
static bool fz_matches(const cnf_coupling* c) {
  if (c->h != S::H || c->w != S::W || c->nk != S::NK || c->c1 != S::C1 || c->c2 != S::C2) return false;
  if ((int)c->dil.size() != S::NBR) return false;
  for (int i = 0; i < S::NBR; ++i) {
    const Branch& s = c->rb[0].br[i];
    if (s.dil != (1 << i) || s.gin != S::g(i) || s.gout != S::g(i) || s.channels != (S::NK >> i)) return false;
  }
  return true;
}

// CNF_NOT_ELIGIBLE: this layer is not covered by the activation-resident kernel (the caller runs the layer-per-kernel path)
int launch_fused_coupling(const cnf_coupling* c, const float* params, FlowView in_view, int in_mask, FlowView out_view,
                          int B, int mode, double* logdet_acc, void* ws, void* stream, bool dry_run) {
  if (mode != HEAD_FWD && mode != HEAD_INV) return CNF_NOT_ELIGIBLE;
  if (!c->ln || c->ks != 3 || c->R < 1 || c->R > FZ_MAXR) return CNF_NOT_ELIGIBLE;
  if (in_mask < 0 || in_mask > 3) return CNF_NOT_ELIGIBLE;
  const int nk = c->nk, cat = c->cat, hw = c->hw();
  if (!(nk == 8 || nk == 16 || nk == 32 || nk == 64) || cat % 4 || c->c2 > 8 || c->c1 > 16) return CNF_NOT_ELIGIBLE;
  if ((int)c->dil.size() > FZ_MAXBR || c->h > 16 || c->w > 32) return CNF_NOT_ELIGIBLE;
  FusedArgs a = {};
  a.in_view = in_view; a.out_view = out_view; a.in_mask = in_mask; a.mask_c = c->mask_c; a.mode = mode;
  a.params = params; a.net_stride = c->net_stride;
  a.stem_w = c->stem_w; a.stem_b = c->stem_b; a.lnf_g = c->lnf_g; a.lnf_b = c->lnf_b;
  a.head_w = c->head_w; a.head_b = c->head_b; a.tanh_off = c->tanh_w;
  a.B = B; a.h = c->h; a.w = c->w; a.hw = hw; a.nk = nk; a.cat = cat; a.c1 = c->c1; a.c2 = c->c2; a.R = c->R;
  a.logdet = logdet_acc;
  a.tscratch = (float*)ws;
  a.ldx = nk + 4;
  a.ldy2 = std::max(cat, nk) + 4;
  a.c2t = c->c2 <= 2 ? 2 : c->c2 <= 4 ? 4 : 8;
  a.nr_variant = (c->h % 7 == 0) ? 7 : 8;
  // grouped-conv task geometry
  a.n_br = (int)c->dil.size();
  int wt = 0, gw_floats = 0;
  for (int i = 0; i < a.n_br; ++i) {
    const Branch& s = c->rb[0].br[i];
    if (s.gin != s.gout || !(s.gin == 1 || s.gin == 2 || s.gin == 4 || s.gin == 8) || s.channels % 4) return CNF_NOT_ELIGIBLE;
    a.br[i] = fz_branch_geom(c->h, c->w, a.nr_variant, s.dil, s.gin, s.channels, s.out_off, wt, gw_floats);
    wt += a.br[i].nwt;
    gw_floats = a.br[i].sb + s.channels;
  }
  a.n_wt = wt;
  for (int r = 0; r < c->R; ++r) {
    const ResBlockLayout& L = c->rb[r];
    FusedRB& f = a.rb[r];
    f.ln1_g = L.ln1_g; f.ln1_b = L.ln1_b; f.pw1_w = L.pw1_w; f.pw1_b = L.pw1_b;
    f.ln2_g = L.ln2_g; f.ln2_b = L.ln2_b; f.ln3_g = L.ln3_g; f.ln3_b = L.ln3_b;
    f.pw2_w = L.pw2_w; f.pw2_b = L.pw2_b;
    for (int i = 0; i < a.n_br; ++i) { f.gw[i] = L.br[i].w_off; f.gb[i] = L.br[i].b_off; }
  }
  // shared-memory carve-up
  const int sz_x = hw * a.ldx, sz_y2 = hw * a.ldy2;
  if ((c->h + 2) * (c->w + 2) * c->c1 > sz_x) return CNF_NOT_ELIGIBLE;
  int wmax = 9 * c->c1 * nk + nk;
  wmax = std::max(wmax, nk * nk + nk);
  wmax = std::max(wmax, gw_floats);
  wmax = std::max(wmax, cat * nk + nk);
  wmax = std::max(wmax, 9 * nk * a.c2t + a.c2t);
  wmax = (wmax + 3) & ~3;
  a.off_y1 = sz_x; a.off_y2 = 2 * sz_x; a.off_w = 2 * sz_x + sz_y2;
  const size_t smem = (size_t)(a.off_w + wmax) * sizeof(float);
  if (smem + kFzStatic > 227 * 1024) return CNF_NOT_ELIGIBLE;
  a.div_w_m = recip32((unsigned)c->w);
  a.div_qnk_m = recip32((unsigned)(nk / 4));
  a.div_qcat_m = recip32((unsigned)(cat / 4));
  a.div_c1_m = recip32((unsigned)c->c1);
  a.div_sw_m = recip32((unsigned)(c->w + 2));
  if (dry_run) return 0;   // eligibility query
  cudaStream_t st = (cudaStream_t)stream;
  const bool two = 2 * (smem + kFzStatic + 1024) <= 228 * 1024;   // two CTAs per SM fit
  if (two && fz_matches<FzCfg2A>(c)) return fz_launch<FzCfg2A, 256, 2>(a, B, 256, smem, st);
  if (two && fz_matches<FzCfg2B>(c)) return fz_launch<FzCfg2B, 256, 2>(a, B, 256, smem, st);
  if (fz_matches<FzCfg2C>(c)) return fz_launch<FzCfg2C, 128, 4>(a, B, 128, smem, st);
  if (fz_matches<FzCfg3A>(c)) return fz_launch<FzCfg3A, 512, 1>(a, B, 512, smem, st);
  if (fz_matches<FzCfg3B>(c)) return fz_launch<FzCfg3B, 512, 1>(a, B, 512, smem, st);
  // run-time shapes.  threads: two CTAs of 256 per SM when they fit, else one CTA of 512; small planes use fewer
  int nt = two ? 256 : 512;
  if (hw <= 64) nt = 128;
  {
    // pixel tiling of the GEMM-shaped stages: a warp tile is (32 / NL) * PXT pixels; fewest rounds x tile work
    a.lognl = nk == 8 ? 1 : nk == 16 ? 2 : nk == 32 ? 3 : 4;
    const int PL = 32 >> a.lognl, nw = nt / 32;
    static const int PXT[4] = {7, 4, 2, 1};
    long best = -1;
    for (int v = 0; v < 4; ++v) {
      const int tiles = (hw + PL * PXT[v] - 1) / (PL * PXT[v]);
      const long cost = (long)((tiles + nw - 1) / nw) * (PXT[v] + 1);
      if (best < 0 || cost < best) { best = cost; a.pw_variant = v; }
    }
  }
  return fz_launch<FzDynamic, 512, 1>(a, B, nt, smem, st);
}

}  // namespace cnf
