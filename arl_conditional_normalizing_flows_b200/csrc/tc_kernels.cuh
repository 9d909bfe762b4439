// tcgen05 / TMEM 1x1-conv kernel (sm_100a only).
//
// out[m, n] = sum_k LN(LReLU(x))[m, k] * W[k, n] + bias[n] (+ res[m, n]) on the 5th-generation tensor
// cores with fp32-level accuracy: every fp32 operand is split x = hi + lo with hi = x truncated to TF32
// (low 13 mantissa bits cleared, exactly what kind::tf32 reads) and lo = x - hi (exact in fp32); per 8-wide K-step
// hi(A) x [hi(W) | lo(W)] (ONE tcgen05.mma of width 2 N) and lo(A) x hi(W) (width N) accumulate into a 2 N-column fp32
// TMEM tile whose halves the epilogue adds (3xTF32; the dropped lo*lo term is ~2^-20 relative).
//
// A CTA owns 128 rows = 4 samples x 32 pixels (gamma/beta of a pixel are fetched once per tile for the 4 samples) and is
// persistent over tiles.  K is consumed in chunks of 32 channels: producer warps cp.async the raw rows (+ gamma, beta) into a
// swizzled shared-memory ring; transform warps apply LReLU + LayerNorm + the hi/lo split and hand the A operand to the tensor
// core -- for N <= 64 by writing it straight into TENSOR MEMORY (tcgen05.st; the MMA takes A from TMEM), for N = 128 through
// K-major no-swizzle operand stages in shared memory (core matrix = 8 rows x 16 B; LBO = 128 B between K chunks, SBO = 1 KB
// between 8-row groups); W (hi and lo) stays resident in shared memory in that same layout; one thread issues the MMAs and a
// tcgen05.commit per chunk that releases the stage.  The epilogue reads the accumulator with tcgen05.ld, adds the halves onto
// the prefetched residual, the bias, stores and accumulates the LayerNorm statistics of LReLU(out) for the consumer.
// The same kernel computes the 1x1 DATA gradients (a.raw_in: operand taken as it is; a.w_trans: W read transposed; zero-padded
// UMMA tile when the output width is not 16 / 32 / 64 / 128).
#pragma once

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_runtime.h>
#include <stdint.h>

#include "cnf_internal.h"
#include "device_utils.cuh"
#include "tc_common.cuh"

namespace cnf {

// ---------------------------------------------------------------------------------------------
// Warp-specialised persistent variant (the one the flow uses):
//   warps 0-7   transform: cp.async ring -> LReLU + LayerNorm + hi/lo split -> A operand stages (tensor memory, or shared)
//   warp  8     MMA issuer (one lane): waits full[stage], issues 2 tcgen05.mma per K-step, commits free[stage]
//   warps 9-16  epilogue: wait tmem_full[buf], tcgen05.ld, + bias (+ residual), store, LN statistics,
//               arrive tmem_empty[buf]   (epilogue warp w owns TMEM lanes 32*(w%4).. = sample w%4)
//   warps 17-20 producers: cp.async of the raw rows, gamma, beta; LayerNorm coefficients of the tile's samples
// The accumulator is double-buffered in TMEM (2 x 2 N columns), so the epilogue of tile t overlaps the
// transform + MMA of tile t+1, and nothing but the ring depth bounds the loads in flight.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA variant of the producer role: tensor maps over the input [K, hw, B, 2 nets] (box 32 channels x 32 pixels x 4 samples) and
// over gamma / beta [K, hw, 2 nets] (box 32 x 32), all with the 128-byte swizzle -- exactly the raw-ring layout the transform
// warps read (16-byte unit q of row r at r * 8 + (q ^ (r & 7))); rows / channels / samples beyond the tensor arrive as zeros.
struct Tc3Maps {
  CUtensorMap x, g, b;
};
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(smem_u32(dst)), "l"((unsigned long long)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"((unsigned long long)tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// timing instrumentation (CNF_PW_DBG bit 128): CTA 0 records clock64() at the hand-off points of its first chunks
__device__ long long g_tc3_clk[8192];
#define TC3_STAMP(role, idx, slot)                                                                     \
  do {                                                                                                 \
    if ((a.dbg & 128) && blockIdx.x == 0 && (idx) < 64) g_tc3_clk[((role) * 64 + (idx)) * 8 + (slot)] = clock64(); \
  } while (0)

// the calling thread's earlier cp.async operations arrive on the mbarrier when they complete (no pending-count increment)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void st8(float* p, const float* v) {   // 256-bit store (sm_100: STG.E.256), p 32-byte aligned
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld8(const float* p, float* v) {   // 256-bit load
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}

// TW = 8 transform warps (one 16-byte unit of each of the 4 samples per thread and K-chunk); NST operand stages; PADN:
// a.N < N outputs (the other columns of the UMMA tile are zero weights and are not stored).
// Measured and dropped (profiles/r02_summary.md): two transform groups on alternate K-chunks (16 + 1 + 4 + 2 warps): the four
// epilogue warps then hold a TMEM buffer across their first store pass and the tile period grows from 6.0 k to 7.5 k cycles.
// Also measured and dropped, after the MMA issuer became an elected thread with straight-line UTCHMMAs (its chain fell from ~1750
// to ~900 cycles per chunk, profiles/r02x_clocks_*): the eight transform warps as two groups of four on ALTERNATE chunks (each
// thread transforms its row's whole chunk; LayerNorm coefficients in a ring indexed by tile): 73.0 / 116.0 us against 69.9 /
// 109.1 us -- the transform role is paced by shared-memory / MIO throughput under the producers' cp.async traffic, not by the
// latency of its hand-off chain, so running two chains in parallel buys nothing.  Likewise sixteen transform warps (one pass of
// two quads per thread and chunk, 72 registers): 70.5 / 116.5 us, and 64-channel chunks again: 77.9 / 127.2 us.  Clock stamps of
// every variant show the same ~2000 cycles per 16 KB chunk in the transform role: ~200 + ~280 for the two (non-blocking)
// mbarrier waits, 500-900 for 12 shared loads, the math and 4 tcgen05.st, ~100-200 for tcgen05.wait::st + fence + __syncwarp and
// ~250 for the two arrives (profiles/r02x_clocks_*).
// ATM: the transformed A operand (hi and lo, 32 + 32 columns per stage) is written straight into TENSOR MEMORY with tcgen05.st
// and the MMAs read it from there (tcgen05.mma with a TMEM A operand): no operand stages in shared memory, no proxy fence,
// the shared-memory port only carries the raw ring and the resident weights, and the ring is 5 deep.  Needs 4 N + 64 NST <= 512
// columns, i.e. N <= 64.
// KCH: channels per K-chunk (32, or 64 with ATM: the MMA issuer and the transform warps pay their per-chunk hand-off costs half
// as often; clock stamps put those at ~1000 of ~2300 cycles per 32-channel chunk).
// TMA: the raw ring is filled by ONE producer thread with three bulk tensor copies per chunk (input box of 4 samples, gamma,
// beta; completion by transaction bytes on raw_full) instead of 128 threads x 12 cp.async of 16 bytes: the per-thread copies
// kept the load/store (MIO) queue of every scheduler busy ~70 % of the time (clock stamps: 700-3300 cycles to issue one
// chunk), and every shared load and mbarrier operation of the transform warps queued behind them.
template <int N, int TW, int NST, bool PADN = false, bool ATM = false, int KCH = 32, bool TMA = false>
__global__ void __launch_bounds__((TW + 1 + 8 + (TMA ? 1 : 4)) * 32, 1)
    pw_tc3_kernel(const GemmArgs a, const int tiles_p, const int tiles_s, const __grid_constant__ Tc3Maps maps) {
  static_assert(KCH == 32 || (KCH == 64 && ATM), "64-channel chunks need the TMEM operand path");
  static_assert(!TMA || (ATM && KCH == 32), "the tensor maps describe 32-channel chunks at the start of the dynamic shared memory");
  constexpr int QPR = KCH / 4;                 // 16-byte quads per row of a chunk
  constexpr int NTT = 32 * QPR;                // 16-byte units per sample and K-chunk (32 rows x QPR quads)
  constexpr int NTH = TW * 32;                 // transform threads
  constexpr int SPT = 4 * 256 / NTH;           // samples per transform thread
  static_assert(SPT == 4, "8 transform warps");
  constexpr int S = 4, PT = 32, M = 128;
  constexpr int KC = KCH;
  constexpr int DEPTH = (ATM && KCH == 32) ? 5 : 3;
  constexpr int A_ST = M * KC, B_ST = N * KC;
  constexpr int RAW = 6 * NTT * 4;
  // one accumulator buffer = 2 N columns: [0, N) hi*hi + lo*hi, [N, 2N) hi*lo (the epilogue adds the halves); two buffers
  constexpr uint32_t TMEM_NEED = 4 * N + (ATM ? 2 * KC * NST : 0);
  static_assert(TMEM_NEED <= 512, "tensor memory");
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  constexpr uint32_t A_TM = 4 * N;             // first column of the A stages (ATM): [stage][hi KC | lo KC]
  extern __shared__ __align__(1024) float tc3_smem[];   // 1024: the 128-byte swizzle of the TMA variant is address based
  const int nchunks = (a.K + KC - 1) / KC;
  float* opsA = tc3_smem;                      // [NST stages][hi, lo][A_ST]
  float* raw = opsA + (ATM ? 0 : NST * 2 * A_ST);    // [DEPTH][6][NTT] float4
  float* Bres = raw + DEPTH * RAW;             // [nchunks][hi, lo][B_ST]
  constexpr int PWARPS = TMA ? 1 : 4, NPT = PWARPS * 32;  // producer (copy) warps / threads
  constexpr int EWARPS = 8;                     // epilogue warps: two per TMEM lane quarter, half of the columns each
  __shared__ __align__(8) uint64_t bar_full[NST], bar_free[NST], bar_tfull[2], bar_tempty[2], raw_full[DEPTH], raw_free[DEPTH];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float2 cf_s[DEPTH][S];   // LayerNorm coefficients (rstd, -mean rstd) of the tile whose first chunk sits in raw slot i
  __shared__ __align__(16) float bias_s[N];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int net = blockIdx.x & 1;
  const int cta = blockIdx.x >> 1, ncta = (gridDim.x + 1 - net) >> 1;
  const int ntiles = tiles_p * tiles_s;
  const int my_tiles = cta < ntiles ? (ntiles - 1 - cta) / ncta + 1 : 0;

  const float* P = a.params + (long long)net * a.net_stride;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&bar_full[i], TW);             // one arrival per transform warp
      mbar_init(&bar_free[i], 1);
    }
    for (int i = 0; i < DEPTH; ++i) {
      mbar_init(&raw_full[i], TMA ? 1 : NPT);  // every producer thread's copies of the chunk have landed (TMA: one arrival + bytes)
      mbar_init(&raw_free[i], TW);             // every transform warp has read the slot
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], EWARPS);       // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == TW) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid < N) bias_s[tid] = (tid < a.N && !a.no_bias) ? P[a.b_off + tid] : 0.f;   // N >= a.N: zero-padded output columns
  // resident B operand (all threads help)
  {
    const float* Wg = P + a.w_off;
    const int bnr = lane & 7, bkr = lane >> 3;
    for (int it = wid; it < ((a.dbg & 64) ? 0 : nchunks * (N / 8) * (KC / 4)); it += TW + 1 + EWARPS + PWARPS) {
      const int c = it / ((N / 8) * (KC / 4)), r = it % ((N / 8) * (KC / 4));
      const int ng = r % (N / 8), kq = r / (N / 8);
      const int n = ng * 8 + bnr, k = c * KC + kq * 4 + bkr;
      float w = 0.f;
      if (k < a.K && n < a.N) w = a.w_trans ? Wg[(long long)n * a.ldw + k] : Wg[(long long)k * a.N + n];
      float h, l;
      tf32_split(w, h, l);
      const int off = c * 2 * B_ST + 4 * (bnr + 8 * kq + (KC / 4) * 8 * ng) + bkr;
      Bres[off] = h;
      Bres[off + B_ST] = l;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (wid < TW) {
    // =============================== transform warps ===============================
    if constexpr (ATM) {
      // warp w: TMEM lane quarter w & 3 = sample of the tile, lane = pixel; warps w and w + 4 take half of the chunk's
      // channels each.  A thread reads its row's quads from the swizzled raw slot (8 lanes = 8 rows -> 8 distinct 16-byte
      // bank slots), applies LReLU + LayerNorm, splits hi / lo and writes 2 x 8 + 2 x 8 columns of its own TMEM lane.
      const int quarter = wid & 3, khalf = wid >> 2;
      const float slope = a.raw_in ? 1.f : CNF_LRELU_SLOPE;
      const uint32_t lane_addr = tmem_d + ((uint32_t)(quarter * 32) << 16);
      int gi = 0, slot = 0;
      const int dq = ncta / tiles_p, dp = ncta - dq * tiles_p;
      int tq = cta / tiles_p, tp = cta - tq * tiles_p;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int s0 = tq * S, p0 = tp * PT;
        const int ns = min(S, a.B - s0);
        tq += dq; tp += dp;
        if (tp >= tiles_p) { tp -= tiles_p; ++tq; }
        const bool ok = (p0 + lane) < a.hw && quarter < ns;
        const bool use_ln = ok && a.ln;
        float sc = 1.f, sh = 0.f;
        for (int c = 0; c < nchunks; ++c, ++gi) {
          const int stage = gi % NST;
          const int kc = min(KC, a.K - c * KC);
          if (tid == 0) TC3_STAMP(0, gi, 0);
          mbar_wait(&raw_full[slot], (gi / DEPTH) & 1);   // the producer warps' copies of chunk gi have landed
          if (tid == 0) TC3_STAMP(0, gi, 3);
          if (c == 0) { const float2 cf = cf_s[slot][quarter]; sc = cf.x; sh = cf.y; }
          if (gi >= NST) mbar_wait(&bar_free[stage], ((gi / NST) - 1) & 1);   // the MMAs that read this A stage are done
          tc_fence_after();
          if (tid == 0) TC3_STAMP(0, gi, 4);
          const float* src = raw + slot * RAW;
          const uint32_t a_hi = lane_addr + A_TM + stage * 2 * KC + (KC / 2) * khalf, a_lo = a_hi + KC;
#pragma unroll
          for (int h2 = 0; h2 < QPR / 4; ++h2) {
            const int q0 = (QPR / 2) * khalf + 2 * h2;    // first of the two channel quads of this pass
            float hi[8], lo[8];
#pragma unroll
            for (int jq = 0; jq < 2; ++jq) {
              const int unit = lane * QPR + ((q0 + jq) ^ (lane & 7));
              const float4 xv = ld4(src + (quarter * NTT + unit) * 4);
              float4 g = ld4(src + (4 * NTT + unit) * 4), be = ld4(src + (5 * NTT + unit) * 4);
              g.x = use_ln ? g.x : 1.f; g.y = use_ln ? g.y : 1.f; g.z = use_ln ? g.z : 1.f; g.w = use_ln ? g.w : 1.f;
              be.x = use_ln ? be.x : 0.f; be.y = use_ln ? be.y : 0.f; be.z = use_ln ? be.z : 0.f; be.w = use_ln ? be.w : 0.f;
              float v[4];
              v[0] = fmaf(fmaf(fmaxf(xv.x, slope * xv.x), sc, sh), g.x, be.x);
              v[1] = fmaf(fmaf(fmaxf(xv.y, slope * xv.y), sc, sh), g.y, be.y);
              v[2] = fmaf(fmaf(fmaxf(xv.z, slope * xv.z), sc, sh), g.z, be.z);
              v[3] = fmaf(fmaf(fmaxf(xv.w, slope * xv.w), sc, sh), g.w, be.w);
              const bool valid = ok && (q0 + jq) * 4 < kc && !(a.dbg & 8);   // rows / channels beyond the tensor feed zeros
#pragma unroll
              for (int i = 0; i < 4; ++i) tf32_split(valid ? v[i] : 0.f, hi[4 * jq + i], lo[4 * jq + i]);
            }
            tmem_st8(a_hi + 8 * h2, hi);
            tmem_st8(a_lo + 8 * h2, lo);
          }
          if (tid == 0) TC3_STAMP(0, gi, 5);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (tid == 0) TC3_STAMP(0, gi, 6);
          if (lane == 0) {
            mbar_arrive(&raw_free[slot]);              // the ring slot may be refilled
            mbar_arrive(&bar_full[stage]);
          }
          if (tid == 0) TC3_STAMP(0, gi, 7);
          slot = slot + 1 == DEPTH ? 0 : slot + 1;
        }
      }
    } else {
    const int wl = wid & 7, sh0 = 0;           // transform warp index; first sample this thread handles
    const int ar = lane & 7, akq = (lane >> 3) + 4 * (wl >> 2);
    const int ap = 8 * (wl & 3) + ar;
    // unit (row ap, quad akq) was written by a producer thread at the swizzled position below (bank-conflict-free
    // for 8 lanes = 8 rows of one quad)
    const int runit = ap * 8 + (akq ^ (ap & 7));
    const float slope = a.raw_in ? 1.f : CNF_LRELU_SLOPE;   // data-gradient use: the operand is taken as it is
    int gi = 0, slot = 0, r = cta;
    // LayerNorm coefficients (rstd, -mean rstd) of the tile's samples: computed once per tile by four producer threads (fp64
    // mean / centred variance) and handed over in shared memory with the tile's first chunk -- no global loads, no fp64 and
    // no extra live registers in the transform warps (the per-thread version cost ~1400 cycles per tile, clock stamps)
    // tile index r = tq * tiles_p + tp is advanced by ncta without divisions
    const int dq = ncta / tiles_p, dp = ncta - dq * tiles_p;
    int tq = cta / tiles_p, tp = cta - tq * tiles_p;
    for (int tl = 0; tl < my_tiles; ++tl, r += ncta) {
      const int s0 = tq * S, p0 = tp * PT;
      const int ns = min(S, a.B - s0);
      tq += dq; tp += dp;
      if (tp >= tiles_p) { tp -= tiles_p; ++tq; }
      float2 cf[SPT] = {};
      if (tid == 0) TC3_STAMP(0, gi, 1);
      const bool pv = (p0 + ap) < a.hw;
      for (int c = 0; c < nchunks; ++c, ++gi) {
        const int stage = gi % NST;
        float* As_hi = opsA + stage * 2 * A_ST;
        float* As_lo = As_hi + A_ST;
        const int kc = min(KC, a.K - c * KC);
        if (tid == 0) TC3_STAMP(0, gi, 0);
        mbar_wait(&raw_full[slot], (gi / DEPTH) & 1);   // the producer warps' copies of chunk gi have landed
        if (c == 0) {
          const float4 c01 = *reinterpret_cast<const float4*>(&cf_s[slot][0]), c23 = *reinterpret_cast<const float4*>(&cf_s[slot][2]);
          cf[0] = make_float2(c01.x, c01.y); cf[1] = make_float2(c01.z, c01.w);
          cf[2] = make_float2(c23.x, c23.y); cf[3] = make_float2(c23.z, c23.w);
        }
        if (tid == 0) TC3_STAMP(0, gi, 3);
        if (gi >= NST) mbar_wait(&bar_free[stage], ((gi / NST) - 1) & 1);   // the MMAs that read this operand stage are done
        if (tid == 0) TC3_STAMP(0, gi, 4);
        if (akq * 4 < kc && !(a.dbg & 32)) {
          // every shared-memory read of this chunk is issued up front into its own registers (no branches between
          // them: slots that were not copied hold stale bits and are discarded by the selects below), so the
          // LDS latencies overlap instead of forming one dependent chain per sample
          const float* src = raw + slot * RAW;
          float4 xv[SPT];
#pragma unroll
          for (int j = 0; j < SPT; ++j) xv[j] = ld4(src + ((sh0 + j) * NTT + runit) * 4);
          float4 g = ld4(src + (4 * NTT + runit) * 4), be = ld4(src + (5 * NTT + runit) * 4);
          const bool use_ln = pv && a.ln;
          g.x = use_ln ? g.x : 1.f; g.y = use_ln ? g.y : 1.f; g.z = use_ln ? g.z : 1.f; g.w = use_ln ? g.w : 1.f;
          be.x = use_ln ? be.x : 0.f; be.y = use_ln ? be.y : 0.f; be.z = use_ln ? be.z : 0.f; be.w = use_ln ? be.w : 0.f;
#pragma unroll
          for (int j = 0; j < SPT; ++j) {
            const int s = sh0 + j;
            const bool ok = pv && s < ns && !(a.dbg & 8);
            const float sc = cf[j].x, sh = cf[j].y;
            float v[4];
            v[0] = fmaxf(xv[j].x, slope * xv[j].x);
            v[1] = fmaxf(xv[j].y, slope * xv[j].y);
            v[2] = fmaxf(xv[j].z, slope * xv[j].z);
            v[3] = fmaxf(xv[j].w, slope * xv[j].w);
            v[0] = fmaf(fmaf(v[0], sc, sh), g.x, be.x);
            v[1] = fmaf(fmaf(v[1], sc, sh), g.y, be.y);
            v[2] = fmaf(fmaf(v[2], sc, sh), g.z, be.z);
            v[3] = fmaf(fmaf(v[3], sc, sh), g.w, be.w);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ok ? v[i] : 0.f;
            float h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) tf32_split(v[i], h[i], l[i]);
            const int unit = ar + 8 * akq + 64 * ((s * PT + ap) >> 3);
            st4(As_hi + 4 * unit, make_float4(h[0], h[1], h[2], h[3]));
            st4(As_lo + 4 * unit, make_float4(l[0], l[1], l[2], l[3]));
          }
        } else if (akq * 4 < ((kc + 7) & ~7)) {
          // K % 8 == 4: the second half of the last 8-wide K-step does not exist; it is fed as zeros (its weights are zero too)
#pragma unroll
          for (int j = 0; j < SPT; ++j) {
            const int unit = ar + 8 * akq + 64 * (((sh0 + j) * PT + ap) >> 3);
            st4(As_hi + 4 * unit, make_float4(0.f, 0.f, 0.f, 0.f));
            st4(As_lo + 4 * unit, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
        if (tid == 0) TC3_STAMP(0, gi, 5);
        fence_async_smem();          // my generic-proxy writes -> async proxy ...
        __syncwarp();                // ... for every lane of the warp, then ONE arrival per warp signals the MMA warp
        if (tid == 0) TC3_STAMP(0, gi, 6);
        if (lane == 0) {
          mbar_arrive(&raw_free[slot]);              // the ring slot may be refilled
          mbar_arrive(&bar_full[stage]);
        }
        slot = slot + 1 == DEPTH ? 0 : slot + 1;
        if (tid == 0) TC3_STAMP(0, gi, 7);
      }
    }
    }
  } else if (wid >= TW + 1 + EWARPS) {
    // =============================== producer (copy) warps ===============================
    // 8 consecutive lanes fetch the 8 16-byte pieces of ONE 128-byte row segment (one L2 line -> one shared-memory
    // wavefront); the piece (row, q) lands in unit row*8 + (q ^ (row & 7)).  Each producer thread owns 2 of the 256
    // units of a chunk; completion is signalled through raw_full[slot] (cp.async.mbarrier.arrive.noinc).
    const float* src_n = a.in + (long long)net * a.in_net_stride;
    const float* gam = P + a.g_off;
    const float* bet = P + a.be_off;
    const int pt = tid - (TW + 1 + EWARPS) * 32;
    int idx = 0, pslot = 0, r = cta;
    if constexpr (TMA) {
      // one thread: waits for the slot, turns the tile's LayerNorm sums into coefficients (the four samples' sums are loaded
      // one tile ahead), then three bulk tensor copies complete the slot's barrier by transaction bytes
      if (pt == 0) {
        const double inv_n = 1.0 / ((double)a.hw * (double)a.K);   // mean and centred variance in fp64 (see ln_coeffs)
        double st[S][2];
        auto load_stats = [&](int rt) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            st[s][0] = 0.0; st[s][1] = 1.0;
            const int sn = (rt / tiles_p) * S + s;
            if (a.ln && rt < ntiles && sn < a.B) {
              const double* sp = a.stats_in + 2 * ((long long)net * a.B + sn);
              st[s][0] = sp[0];
              st[s][1] = sp[1];
            }
          }
        };
        const uint32_t chunk_bytes = (uint32_t)((a.ln ? 6 : 4) * NTT * 16);
        load_stats(r);
        for (int tl = 0; tl < my_tiles; ++tl, r += ncta) {
          const int s0 = (r / tiles_p) * S, p0 = (r % tiles_p) * PT;
          for (int c = 0; c < nchunks; ++c, ++idx) {
            if (pt == 0) TC3_STAMP(3, idx, 0);
            if (idx >= DEPTH) mbar_wait(&raw_free[pslot], ((idx / DEPTH) - 1) & 1);
            if (pt == 0) TC3_STAMP(3, idx, 1);
            if (c == 0) {
#pragma unroll
              for (int s = 0; s < S; ++s) {
                // rsqrtf (2 ulp) instead of the IEEE 1/sqrt sequence; the difference is ~1e-7 relative
                const double md = st[s][0] * inv_n;
                const float m_ = (float)md;
                const float var = fmaxf((float)(st[s][1] * inv_n - md * md), 0.f);
                const float sc = rsqrtf(var + (float)CNF_LN_EPS);
                cf_s[pslot][s] = a.ln ? make_float2(sc, -m_ * sc) : make_float2(1.f, 0.f);
              }
              load_stats(r + ncta);
            }
            float* dst = raw + pslot * RAW;
            fence_async_smem();   // the transform warps' generic reads of this slot (ordered by raw_free) before the async writes
            mbar_arrive_expect_tx(&raw_full[pslot], chunk_bytes);   // release: the coefficients above are visible with the phase
            tma_load_4d(dst, &maps.x, c * KC, p0, s0, net, &raw_full[pslot]);
            if (a.ln) {
              tma_load_3d(dst + 4 * NTT * 4, &maps.g, c * KC, p0, net, &raw_full[pslot]);
              tma_load_3d(dst + 5 * NTT * 4, &maps.b, c * KC, p0, net, &raw_full[pslot]);
            }
            if (pt == 0) TC3_STAMP(3, idx, 2);
            pslot = pslot + 1 == DEPTH ? 0 : pslot + 1;
          }
        }
      }
    } else {
      // threads 0-3: (sum, sumsq) of sample pt of the NEXT tile, loaded one tile ahead
      const double inv_n = 1.0 / ((double)a.hw * (double)a.K);   // mean and centred variance in fp64 (see ln_coeffs)
      double st_s = 0.0, st_q = 1.0;
      auto load_stats = [&](int rt) {
        st_s = 0.0; st_q = 1.0;
        const int sn = (rt / tiles_p) * S + pt;
        if (a.ln && pt < S && rt < ntiles && sn < a.B) {
          const double* sp = a.stats_in + 2 * ((long long)net * a.B + sn);
          st_s = sp[0];
          st_q = sp[1];
        }
      };
      load_stats(r);
      for (int tl = 0; tl < my_tiles; ++tl, r += ncta) {
        const int s0 = (r / tiles_p) * S, p0 = (r % tiles_p) * PT;
        const int ns = min(S, a.B - s0);
        for (int c = 0; c < nchunks; ++c, ++idx) {
          if (pt == 0) TC3_STAMP(3, idx, 0);
          if (idx >= DEPTH) mbar_wait(&raw_free[pslot], ((idx / DEPTH) - 1) & 1);
          if (pt == 0) TC3_STAMP(3, idx, 1);
          if (c == 0 && pt < S) {
            // rsqrtf (2 ulp) instead of the IEEE 1/sqrt sequence; the difference is ~1e-7 relative
            const double md = st_s * inv_n;
            const float m_ = (float)md;
            const float var = fmaxf((float)(st_q * inv_n - md * md), 0.f);
            const float sc = rsqrtf(var + (float)CNF_LN_EPS);
            cf_s[pslot][pt] = a.ln ? make_float2(sc, -m_ * sc) : make_float2(1.f, 0.f);
            load_stats(r + ncta);
            __threadfence_block();   // ordered before this thread's arrival on raw_full[pslot] below
          }
          const int k0 = c * KC, kc = min(KC, a.K - k0);
          float* dst = raw + pslot * RAW;
  #pragma unroll
          for (int h = 0; h < NTT / NPT; ++h) {
            const int u = pt + h * NPT;
            const int cq = u % QPR, crow = u / QPR;
            const int gp = p0 + crow;
            if (cq * 4 < kc && gp < a.hw && !(a.dbg & 2)) {
              const int cunit = crow * QPR + (cq ^ (crow & 7));
              const long long e = (long long)gp * a.K + k0 + cq * 4;
  #pragma unroll
              for (int s = 0; s < S; ++s)
                if (s < ns) cp_async16_cg(dst + (s * NTT + cunit) * 4, src_n + ((long long)(s0 + s) * a.hw) * a.K + e);
              if (a.ln) {
                cp_async16_ca(dst + (4 * NTT + cunit) * 4, gam + e);
                cp_async16_ca(dst + (5 * NTT + cunit) * 4, bet + e);
              }
            }
          }
          cp_async_mbar_arrive_noinc(&raw_full[pslot]);
          if (pt == 0) TC3_STAMP(3, idx, 2);
          pslot = pslot + 1 == DEPTH ? 0 : pslot + 1;
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
  } else if (wid == TW) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      // B = [hi rows | lo rows] is one K-major operand of 2 N rows (the two blocks are adjacent and share the layout), so
      // hi(A) * [hi(B) | lo(B)] is ONE instruction of width 2 N and the A operand is read twice per K-step instead of
      // three times; lo(A) * hi(B) accumulates into the first N columns
      constexpr uint32_t idesc2 = umma_idesc_tf32(2 * N), idesc1 = umma_idesc_tf32(N);
      constexpr uint32_t LBO = 128, SBO = (KC / 4) * 128;
      const uint32_t a_base = smem_u32(opsA), b_base = smem_u32(Bres);
      int gi = 0;
      const int n_total = my_tiles * nchunks;
      bool ready = false;                      // the next chunk's operands were already seen complete (early poll below)
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int buf = tl & 1;
        if (tl >= 2) mbar_wait(&bar_tempty[buf], ((tl >> 1) - 1) & 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_d + buf * 2 * N;
        for (int c = 0; c < nchunks; ++c, ++gi) {
          const int stage = gi % NST;
          const int kc = min(KC, a.K - c * KC);
          TC3_STAMP(1, gi, 0);
          if (!ready) mbar_wait(&bar_full[stage], (gi / NST) & 1);
          TC3_STAMP(1, gi, 1);
          tc_fence_after();
          const uint32_t a_hi = a_base + stage * 2 * A_ST * 4, a_lo = a_hi + A_ST * 4;
          const uint32_t b_hi = b_base + c * 2 * B_ST * 4;
          // the descriptors of a K-step differ from those of step 0 only in the start-address field (16-byte units;
          // shared-memory addresses stay below 2^18, so the 14-bit field never carries): one 64-bit add each
          const uint64_t dah0 = umma_desc(a_hi, LBO, SBO), dal0 = umma_desc(a_lo, LBO, SBO), dbh0 = umma_desc(b_hi, LBO, SBO);
          // one K-step: hi(A) [hi(B) | lo(B)] (width 2 N) then lo(A) hi(B) (width N)
          auto kstep = [&](const int ks, const uint32_t acc) {
            const uint64_t adv = (uint64_t)((ks * 2 * LBO) >> 4);
            if constexpr (ATM) {
              const uint32_t at = tmem_d + A_TM + stage * 2 * KC + 8 * ks;   // hi columns of this K-step; lo KC further
              umma_tf32_ts(d_addr, at, dbh0 + adv, idesc2, acc);
              umma_tf32_ts(d_addr, at + KC, dbh0 + adv, idesc1, 1);
            } else {
              umma_tf32(d_addr, dah0 + adv, dbh0 + adv, idesc2, acc);
              umma_tf32(d_addr, dal0 + adv, dbh0 + adv, idesc1, 1);
            }
          };
#ifdef CNF_DEBUG
          const int nks = (a.dbg & 4) ? (c == 0 ? 1 : 0) : (kc + 7) / 8;
#else
          const int nks = (kc + 7) / 8;
#endif
          if (nks == KC / 8) {                    // full chunk: straight-line code, every offset an immediate
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) kstep(ks, ks ? 1u : (uint32_t)(c != 0));
          } else {
            for (int ks = 0; ks < nks; ++ks) kstep(ks, (uint32_t)((c | ks) != 0));
          }
          TC3_STAMP(1, gi, 2);
          // poll the NEXT chunk's barrier before the commits: the ~250-cycle latency of the try_wait overlaps them (this
          // thread's chain -- wait, 8 MMAs, commits -- paces the kernel, clock stamps in profiles/r02t_*)
          ready = gi + 1 < n_total && mbar_try_wait(&bar_full[(gi + 1) % NST], ((gi + 1) / NST) & 1);
          umma_commit(&bar_free[stage]);
          if (c == nchunks - 1) umma_commit(&bar_tfull[buf]);
          TC3_STAMP(1, gi, 3);
        }
      }
    }
  } else if (wid < TW + 1 + EWARPS) {
    // =============================== epilogue warps ===============================
    // warp e: TMEM lane quarter (wid & 3) == sample index (PT == 32), column half e >> 2.  The accumulator values are
    // pulled into registers and the TMEM buffer is released BEFORE bias / residual / statistics / 256-bit stores.
    constexpr bool SPLIT = EWARPS == 8 && N >= 32;   // two warps per lane quarter, half of the columns each
    constexpr int NH = SPLIT ? N / 2 : N;        // columns per epilogue warp
    constexpr int CH = NH > 32 ? 32 : NH;        // ... handled CH at a time (registers)
    const int quarter = wid & 3, half = (wid - (TW + 1)) >> 2;
    const bool has_cols = SPLIT || half == 0;
    const int c0 = SPLIT ? half * NH : 0;
    int r = cta;
    for (int tl = 0; tl < my_tiles; ++tl, r += ncta) {
      const int buf = tl & 1;
      const int s0 = (r / tiles_p) * S, p0 = (r % tiles_p) * PT;
      const int ns = min(S, a.B - s0);
      const int egp = p0 + lane;
      const bool erow = egp < a.hw && quarter < ns && has_cols;
      const long long row = ((long long)(s0 + quarter) * a.hw + egp) * a.N + c0;
      float* out_r = a.out + (long long)net * a.out_net_stride + row;
      const float* res_r = a.res ? a.res + (long long)net * a.out_net_stride + row : nullptr;
      // acc starts as the residual row segment (in flight while the MMAs of this tile run) or zero; the two accumulator
      // halves are added to it 16 columns at a time (two tcgen05.ld in flight)
      float acc[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) acc[j] = 0.f;
      if (res_r && erow) {
#pragma unroll
        for (int j = 0; j < CH; j += 8)
          if (!PADN || c0 + j < a.N) ld8(res_r + j, acc + j);
      }
      if (wid == TW + 1 && lane == 0) TC3_STAMP(2, tl, 0);
      mbar_wait(&bar_tfull[buf], (tl >> 1) & 1);
      if (wid == TW + 1 && lane == 0) TC3_STAMP(2, tl, 1);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int cc = 0; cc < NH; cc += CH) {
        if (cc > 0) {
#pragma unroll
          for (int j = 0; j < CH; ++j) acc[j] = 0.f;
          if (res_r && erow) {
#pragma unroll
            for (int j = 0; j < CH; j += 8)
              if (!PADN || c0 + cc + j < a.N) ld8(res_r + cc + j, acc + j);
          }
        }
        if (has_cols && !(a.dbg & 16)) {
          const uint32_t t0 = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 2 * N + c0 + cc);
#pragma unroll
          for (int cb = 0; cb < CH; cb += 16) {
            uint32_t u0[16], u1[16];
            tmem_ld16_nowait(t0 + cb, u0);
            tmem_ld16_nowait(t0 + N + cb, u1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[cb + i] += __uint_as_float(u0[i]) + __uint_as_float(u1[i]);
          }
        }
        if (cc + CH >= NH) {
          // the accumulator values are in registers: release the TMEM buffer BEFORE bias / statistics / stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_tempty[buf]);   // accumulator buffer may be overwritten
          if (wid == TW + 1 && lane == 0) TC3_STAMP(2, tl, 2);
        }
        if (erow && !(a.dbg & 16)) {
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            if (PADN && c0 + cc + j >= a.N) continue;      // zero-padded columns of the UMMA tile
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i] = acc[j + i] + bias_s[c0 + cc + j + i];
              const float l = fmaxf(o[i], CNF_LRELU_SLOPE * o[i]);
              s1 += l;
              s2 = fmaf(l, l, s2);
            }
            if (!(a.dbg & 1)) st8(out_r + cc + j, o);
          }
        }
      }
      if (a.stats_out) {
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0 && quarter < ns && has_cols) {
          double* so = a.stats_out + 2 * ((long long)net * a.B + s0 + quarter);
          atomicAdd(so, (double)s1);
          atomicAdd(so + 1, (double)s2);
        }
      }
      if (wid == TW + 1 && lane == 0) TC3_STAMP(2, tl, 3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == TW) tmem_dealloc(tmem_d, TMEM_COLS);
}

}  // namespace cnf
