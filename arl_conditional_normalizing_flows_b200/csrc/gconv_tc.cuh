// Grouped dilated 3x3 conv for WIDE groups (16 or 32 channels per group) as an implicit GEMM on the 5th-generation tensor
// cores (tcgen05 / TMEM), sm_100a only.  Reference: grouped_convolution (conv_cINN_base_functions.py F:364-413) inside
// dilated_residual_block (F:577-590) with add_common_layers (F:330-362) applied to its input.
//
// Where it is used: the builder-chosen hyper-parameters of BASELINE configs 4 and 5 give group widths of
// 16 / 32 channels on the full-resolution layers (nk 64 / cardinality 4 or 2, nk 128 / 256 with cardinality 8).  Those
// branches are dense enough for the tensor pipe: a 32 x 32 block per tap is 16 x the work per operand byte of the 8-wide
// groups of config 2, which stay on the FFMA2 octet kernel (gconv_oct.cuh; DESIGN.md section 3 explains the split).
//
// One CTA item = (net, sample, group, 16 x 8 output pixels).  The input window (16 + 2 dil) x (8 + 2 dil) pixels x G channels
// is read once from global memory, LReLU + LayerNorm are applied on the way (gamma / beta per element, zero outside the
// image: Keras pads the NORMALISED tensor), and every value is split x = hi + lo (hi = TF32 truncation, exactly what
// kind::tf32 reads) into two channel-planar shared-memory images: [4-channel chunk][window pixel][16 bytes].  In that
// layout 8 consecutive window pixels of one chunk are one contiguous 128-byte UMMA core matrix (K-major, no swizzle), the
// next image row is SBO = SW * 16 bytes further and the next channel chunk LBO = one plane further, so the A operand of
// tap (ky, kx) is the SAME image addressed at start + ((ky dil) SW + kx dil) * 16 bytes: the nine taps are nine descriptor
// offsets, nothing is gathered or copied.  Per tap and 8-channel K-step one thread issues hi(A) x [hi(W) | lo(W)] (width
// 2 G) and lo(A) x hi(W) (width G): 3xTF32, fp32-exact to ~1e-6.  The 128 x G accumulator lives in TMEM; the epilogue adds the
// two halves, the bias, accumulates the LayerNorm statistics of LReLU(out) for the consumer and stores G contiguous floats
// per pixel.  The group's weights (9 taps, hi and lo) stay in shared memory while the CTA walks over its items.
#pragma once

namespace cnf {

constexpr int GTC_TH = 16, GTC_TW = 8;     // output pixels of an item: 16 rows x 8 columns = the 128 rows of the UMMA tile
constexpr int GTC_WT = 8;                  // worker warps (transform + epilogue: two per TMEM lane quarter)
constexpr int GTC_NT = (GTC_WT + 1) * 32;  // + one warp whose lane 0 issues the MMAs

struct GtcArgs {
  const float* in;            // [2][B][h][w][Cin]
  float* out;                 // [2][B][h][w][Cout]
  long long in_net_stride, out_net_stride;
  const float* params;
  long long net_stride, g_off, be_off, w_off, b_off;
  const double* stats_in;
  double* stats_out;
  int B, h, w, Cin, Cout, ln;
  int dil, groups, out_off;
  int tiles_y, tiles_x, n_items;   // items per (net, group): B * tiles_y * tiles_x
  int ctas_per_ng;                 // CTAs that share one (net, group)
  int SW, NPX, NPXP;               // window width, window pixels, plane stride in pixels (= 1 mod 8: conflict-free chunk stores)
  int vec4;                        // 1: 128-bit output stores (Cout % 4 == 0), 0: 64-bit (Cout % 4 == 2, e.g. cat = 62)
  // data-gradient mode (training): the input is the raw gradient slice of the branch (no LReLU / LayerNorm, channel offset
  // in_off), the weights are read flipped and transposed (dx[p] = sum_tap dy[p - off(tap)] W[tap]^T), the result is ADDED to
  // `out` (the branches of a block are separate launches on one stream), no bias, no statistics
  int bwd, in_off;
};

// The item loop is software-pipelined over two plane buffers: while the tensor core works through the 9 x G / 8 x 2 MMAs
// of item i (asynchronous, issued by lane 0 of the extra warp), the eight worker warps transform the window of item i + 1
// into the other buffer; they then wait for the MMAs' mbarrier, read the accumulator and store item i.
template <int G>
__global__ void __launch_bounds__(GTC_NT, 1) gconv_tc_kernel(const GtcArgs a) {
  constexpr int NQ = G / 4;                    // 4-channel chunks of the group
  constexpr int WT = 2 * G * G;                // floats of one tap's weights: [hi rows | lo rows] x G inputs
  constexpr int NW = GTC_WT * 32;              // worker threads
  constexpr uint32_t TMEM_COLS = 2 * G < 32 ? 32 : 2 * G;
  extern __shared__ __align__(128) float gtc_smem[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[G];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ng = blockIdx.x / a.ctas_per_ng, rank = blockIdx.x - ng * a.ctas_per_ng;
  const int net = ng / a.groups, grp = ng - net * a.groups;
  const int plane = a.NPXP * 4;                // floats per chunk plane
  const int bufsz = 2 * NQ * plane;            // floats per buffer: hi planes, lo planes
  float* Abuf = gtc_smem;                      // [2 buffers][hi | lo][NQ][NPXP][4]
  float* Bw = Abuf + 2 * bufsz;                // [9][hi | lo][G rows][G] K-major core-matrix layout

  const float* P = a.params + (long long)net * a.net_stride;
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid < G) bias_s[tid] = a.bwd ? 0.f : P[a.b_off + grp * G + tid];
  // the group's weights W[tap][ci][co] -> B operand rows n = co, columns k = ci: 16-byte unit (n % 8) + 8 (k / 4) + 8 NQ (n / 8)
  {
    const float* Wg = P + a.w_off + (long long)grp * 9 * G * G;
    for (int i = tid; i < 9 * G * G; i += GTC_NT) {
      const int tap = i / (G * G), r = i - tap * G * G;
      const int ci = r / G, co = r - ci * G;
      float hi, lo;
      tf32_split(Wg[i], hi, lo);
      // forward: rows n = co, columns k = ci of tap; data gradient: rows n = ci, columns k = co of the mirrored tap
      const int n = a.bwd ? ci : co, k = a.bwd ? co : ci, tp = a.bwd ? 8 - tap : tap;
      const int off = tp * WT + 4 * ((n & 7) + 8 * (k >> 2) + 8 * NQ * (n >> 3)) + (k & 3);
      Bw[off] = hi;
      Bw[off + G * G] = lo;
    }
  }

  const float* src_n = a.in + (long long)net * a.in_net_stride + a.in_off + grp * G;
  float* out_n = a.out + (long long)net * a.out_net_stride + a.out_off + grp * G;
  const float* gam = P + a.g_off + grp * G;
  const float* bet = P + a.be_off + grp * G;
  const int d = a.dil, SW = a.SW;
  const int tiles = a.tiles_y * a.tiles_x;
  const double n_ln = (double)a.h * (double)a.w * (double)a.Cin;

  // window of item `it` -> LReLU + LayerNorm -> hi / lo planes of buffer `buf`; 8 consecutive lanes fetch the NQ <= 8 chunks
  // of one pixel (worker threads only)
  auto transform = [&](int it, int buf) {
    const int b = it / tiles, t = it - b * tiles;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    const int y0 = ty * GTC_TH, x0 = tx * GTC_TW;
    float mean = 0.f, rstd = 1.f;
    if (a.ln) ln_coeffs(a.stats_in, (long long)net * a.B + b, n_ln, mean, rstd);
    const float sc = rstd, sh = -mean * rstd;
    float* A_hi = Abuf + buf * bufsz;
    float* A_lo = A_hi + NQ * plane;
    const float* src_b = src_n + (long long)b * a.h * a.w * a.Cin;
    constexpr int UB = 4;                                  // units in flight per thread (12 independent 128-bit loads)
    const int n_units = a.NPX * NQ;
    for (int u0 = tid; u0 < n_units; u0 += UB * NW) {
      float4 v[UB], g[UB], be[UB];
      int q[UB], c4[UB];
      bool in[UB];
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        const int u = min(u0 + k * NW, n_units - 1);
        q[k] = u / NQ; c4[k] = u - q[k] * NQ;
        const int wy = q[k] / SW, wx = q[k] - wy * SW;
        const int iy = y0 - d + wy, ix = x0 - d + wx;
        in[k] = iy >= 0 && iy < a.h && ix >= 0 && ix < a.w;
        const long long e = in[k] ? ((long long)iy * a.w + ix) * a.Cin + 4 * c4[k] : 0;
        v[k] = ld4(src_b + e);
        if (a.ln) { g[k] = ld4(gam + e); be[k] = ld4(bet + e); }
      }
#pragma unroll
      for (int k = 0; k < UB; ++k) {
        if (u0 + k * NW >= n_units) continue;
        float4 t = v[k];
        if (!a.bwd) {
          t.x = fmaxf(t.x, CNF_LRELU_SLOPE * t.x); t.y = fmaxf(t.y, CNF_LRELU_SLOPE * t.y);
          t.z = fmaxf(t.z, CNF_LRELU_SLOPE * t.z); t.w = fmaxf(t.w, CNF_LRELU_SLOPE * t.w);
        }
        if (a.ln) {
          t.x = fmaf(fmaf(t.x, sc, sh), g[k].x, be[k].x);
          t.y = fmaf(fmaf(t.y, sc, sh), g[k].y, be[k].y);
          t.z = fmaf(fmaf(t.z, sc, sh), g[k].z, be[k].z);
          t.w = fmaf(fmaf(t.w, sc, sh), g[k].w, be[k].w);
        }
        if (!in[k]) t = make_float4(0.f, 0.f, 0.f, 0.f);   // outside the image: Keras pads the NORMALISED tensor with zeros
        float4 hi, lo;
        tf32_split(t.x, hi.x, lo.x); tf32_split(t.y, hi.y, lo.y);
        tf32_split(t.z, hi.z, lo.z); tf32_split(t.w, hi.w, lo.w);
        st4(A_hi + c4[k] * plane + q[k] * 4, hi);
        st4(A_lo + c4[k] * plane + q[k] * 4, lo);
      }
    }
  };

  int it = rank, buf = 0;
  if (it < a.n_items && wid < GTC_WT) transform(it, 0);
  fence_async_smem();                      // generic-proxy writes (planes, weights) -> async proxy (the tensor core reads them)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  uint32_t phase = 0;

  for (; it < a.n_items; it += a.ctas_per_ng, buf ^= 1) {
    if (wid == GTC_WT) {
      // ---- 9 taps x G / 8 K-steps x 2 MMAs on buffer `buf`, one (elected) thread
      if (elect_one()) {
        tc_fence_after();
        constexpr uint32_t idesc2 = umma_idesc_tf32(2 * G), idesc1 = umma_idesc_tf32(G);
        const uint32_t lbo_a = (uint32_t)a.NPXP * 16u, sbo_a = (uint32_t)SW * 16u;
        const uint32_t a_hi = smem_u32(Abuf + buf * bufsz), a_lo = a_hi + (uint32_t)(NQ * plane * 4), b0 = smem_u32(Bw);
        const uint64_t dah0 = umma_desc(a_hi, lbo_a, sbo_a), dal0 = umma_desc(a_lo, lbo_a, sbo_a);
        const uint64_t db0 = umma_desc(b0, 128u, (uint32_t)NQ * 128u);
        bool first = true;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - 3 * ky;
          const uint32_t shift = (uint32_t)((ky * d) * SW + kx * d);          // window pixels = 16-byte units
#pragma unroll
          for (int ks = 0; ks < G / 8; ++ks) {
            const uint64_t adv_a = (uint64_t)(shift + (uint32_t)(2 * ks) * (uint32_t)a.NPXP);
            const uint64_t adv_b = (uint64_t)((uint32_t)(tap * WT * 4 + ks * 256) >> 4);
            umma_tf32(tmem_d, dah0 + adv_a, db0 + adv_b, idesc2, first ? 0u : 1u);
            umma_tf32(tmem_d, dal0 + adv_a, db0 + adv_b, idesc1, 1u);
            first = false;
          }
        }
        umma_commit(&bar_mma);
      }
    } else {
      const int nx = it + a.ctas_per_ng;
      if (nx < a.n_items) transform(nx, buf ^ 1);       // overlaps the MMAs of `it`
      fence_async_smem();
      // ---- epilogue: TMEM lane = output pixel (row = lane / 8 of the quarter, column = lane % 8); warps w and w + 4
      // share the lane quarter w % 4 and take half of the G channels each
      mbar_wait(&bar_mma, phase);
      tc_fence_after();
      const int b = it / tiles, t = it - b * tiles;
      const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
      const int y0 = ty * GTC_TH, x0 = tx * GTC_TW;
      constexpr int NH = G / 2;
      const int quarter = wid & 3, half = wid >> 2;
      const int c0 = half * NH;
      const int row = quarter * 4 + (lane >> 3), col = lane & 7;
      const int oy = y0 + row, ox = x0 + col;
      float s1 = 0.f, s2 = 0.f;
      float acc[NH];
      const uint32_t t0 = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
#pragma unroll
      for (int cb = 0; cb < NH; cb += 8) {
        float v0[8], v1[8];
        tmem_ld<8>(t0 + cb, v0);
        tmem_ld<8>(t0 + G + cb, v1);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[cb + i] = v0[i] + v1[i] + bias_s[c0 + cb + i];
      }
      tc_fence_before();                   // ordered before the barrier below: the next item's MMAs restart the accumulator
      if (oy < a.h && ox < a.w) {
        float* o = out_n + (((long long)b * a.h + oy) * a.w + ox) * a.Cout + c0;
#pragma unroll
        for (int j = 0; j < NH; j += 4) {
          if (a.bwd) {   // accumulate onto the other branches' contributions (vec4 layout: Cout = nk, a multiple of 4)
            const float4 old = ld4(o + j);
            acc[j] += old.x; acc[j + 1] += old.y; acc[j + 2] += old.z; acc[j + 3] += old.w;
          }
          if (a.vec4) {
            st4(o + j, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
          } else {
            *reinterpret_cast<float2*>(o + j) = make_float2(acc[j], acc[j + 1]);
            *reinterpret_cast<float2*>(o + j + 2) = make_float2(acc[j + 2], acc[j + 3]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float l = fmaxf(acc[j + i], CNF_LRELU_SLOPE * acc[j + i]);
            s1 += l;
            s2 = fmaf(l, l, s2);
          }
        }
      }
      if (a.stats_out) {
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
          double* so = a.stats_out + 2 * ((long long)net * a.B + b);
          atomicAdd(so, (double)s1);
          atomicAdd(so + 1, (double)s2);
        }
      }
    }
    phase ^= 1;
    // the MMAs of `it` have completed (every worker waited for their mbarrier), its accumulator has been read, the planes of
    // the next item are written: the next iteration may issue
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// One branch with gin == gout in {16, 32}, ksize 3; forward, or (g.bwd) its data gradient.  CNF_NOT_ELIGIBLE for anything else.
static int launch_gconv_tc_branch(const GconvArgs& g, int bi, cudaStream_t st) {
  const GconvBranch& br = g.br[bi];
  const int G = br.gin;
  if (g.ks != 3 || br.gin != br.gout || !(G == 16 || G == 32) || (br.in_off % 4)) return CNF_NOT_ELIGIBLE;
  if (!g.bwd && br.in_off != 0) return CNF_NOT_ELIGIBLE;
  if ((g.Cin % 4) || (g.Cout % 2) || (br.out_off % 4) || br.in_off + br.groups * G > g.Cin) return CNF_NOT_ELIGIBLE;
  if (g.bwd && ((g.Cout % 4) || g.ln)) return CNF_NOT_ELIGIBLE;
  if ((((uintptr_t)g.in) & 15) || (((uintptr_t)g.out) & 15) || ((g.in_net_stride * 4) % 16) || ((g.out_net_stride * 4) % 16)) return CNF_NOT_ELIGIBLE;
  GtcArgs a{};
  a.in = g.in; a.out = g.out; a.in_net_stride = g.in_net_stride; a.out_net_stride = g.out_net_stride;
  a.params = g.params; a.net_stride = g.net_stride; a.g_off = g.g_off; a.be_off = g.be_off;
  a.w_off = br.w_off; a.b_off = br.b_off;
  a.stats_in = g.stats_in; a.stats_out = g.stats_out;
  a.B = g.B; a.h = g.h; a.w = g.w; a.Cin = g.Cin; a.Cout = g.Cout; a.ln = g.ln;
  a.dil = br.dil; a.groups = br.groups; a.out_off = br.out_off;
  a.bwd = g.bwd ? 1 : 0; a.in_off = br.in_off;
  if (g.bwd) a.stats_out = nullptr;
  a.tiles_y = (g.h + GTC_TH - 1) / GTC_TH;
  a.tiles_x = (g.w + GTC_TW - 1) / GTC_TW;
  a.n_items = g.B * a.tiles_y * a.tiles_x;
  a.SW = GTC_TW + 2 * br.dil;
  a.NPX = (GTC_TH + 2 * br.dil) * a.SW;
  a.NPXP = ((a.NPX + 7) & ~7) + 1;
  a.vec4 = (g.Cout % 4) == 0 ? 1 : 0;
  // descriptor fields are 14 bits of 16-byte units
  if ((long long)a.NPXP >= 16384 || a.SW >= 16384) return CNF_NOT_ELIGIBLE;
  const size_t smem = ((size_t)2 * 2 * (G / 4) * a.NPXP * 4 + (size_t)9 * 2 * G * G) * sizeof(float);   // two plane buffers + weights
  if (smem > 200 * 1024) return CNF_NOT_ELIGIBLE;
  int n_sm = 0;
  CU_TRY((cudaError_t)device_sm_count(&n_sm));
  const int ngs = 2 * br.groups;
  a.ctas_per_ng = std::max(1, std::min(a.n_items, (2 * n_sm + ngs - 1) / ngs));   // ~2 CTAs per SM: phases of co-resident CTAs overlap
  const int grid = ngs * a.ctas_per_ng;
  if (G == 16) {
    static SmemAttrCache cache;
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)gconv_tc_kernel<16>, smem, cache));
    gconv_tc_kernel<16><<<grid, GTC_NT, smem, st>>>(a);
  } else {
    static SmemAttrCache cache;
    CU_TRY((cudaError_t)ensure_dynamic_smem((const void*)gconv_tc_kernel<32>, smem, cache));
    gconv_tc_kernel<32><<<grid, GTC_NT, smem, st>>>(a);
  }
  return (int)cudaGetLastError();
}

}  // namespace cnf
