// How fast does one thread get tcgen05.mma.kind::tf32 (M = 128, K = 8) instructions through the tensor pipe, as a function of
// the tile width N, of the A source (tensor memory / shared memory) and -- the question behind this file -- of how many
// INDEPENDENT accumulators consecutive instructions rotate over?  pw_tc3_kernel issues, per K-step, N = 128 and N = 64 into
// the same columns; its clock stamps show ~150 cycles per instruction where the issue floor is N / 2 = 64 / 32 cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../arl_conditional_normalizing_flows_b200/csrc -I../../include \
//        -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace cnf;

struct Pat {
  int len;
  int n[8];      // tile width of instruction i % len
  int dcol[8];   // accumulator column offset
};

// patterns are compile-time so that the issue loop is what a tuned kernel would run: fully unrolled, descriptors and
// instruction descriptors folded into uniform-register immediates, no memory traffic between two tcgen05.mma
template <int LEN, int N0, int D0, int N1, int D1, int N2, int D2, int N3, int D3, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k(int ni, int reps, long long* out) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, wid = tid >> 5;
  for (int i = tid; i < 12288; i += 128) smem[i] = 0.f;   // A: 128 x 32 floats (16 KB), B: 256 x 32 floats (32 KB)
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) tmem_alloc(&slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 4096);
    constexpr uint32_t LBO = 128, SBO = 8 * 128;      // KC = 32 channels per chunk, as in pw_tc3_kernel
    const uint64_t da = umma_desc(a_base, LBO, SBO), db = umma_desc(b_base, LBO, SBO);
    constexpr int NS[4] = {N0, N1, N2, N3}, DS[4] = {D0, D1, D2, D3};
    long long best_issue = 1ll << 60, best_total = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int i0 = 0; i0 < ni; i0 += 4 * LEN) {
#pragma unroll
        for (int j = 0; j < 4 * LEN; ++j) {
          constexpr uint64_t one = 1;
          const int p = j % LEN;
          const uint32_t idesc = umma_idesc_tf32((uint32_t)NS[p]);
          const uint64_t adv = (uint64_t)((((j / LEN) & 3) * 2 * LBO) >> 4) * one;
          if (A_TMEM) umma_tf32_ts(tm + DS[p], tm + 448 + 8 * ((j / LEN) & 3), db + adv, idesc, (i0 | j) >= LEN);
          else umma_tf32(tm + DS[p], da + adv, db + adv, idesc, (i0 | j) >= LEN);
        }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, r & 1);
      const long long t2 = clock64();
      tc_fence_after();
      if (t1 - t0 < best_issue) best_issue = t1 - t0;
      if (t2 - t0 < best_total) best_total = t2 - t0;
    }
    if (blockIdx.x == 0) {
      out[0] = best_issue;
      out[1] = best_total;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc(tm, 512);
}

template <int LEN, int N0, int D0, int N1, int D1, int N2, int D2, int N3, int D3>
static void run(const char* name, int grid, long long* d) {
  for (int a_tmem = 1; a_tmem >= 0; --a_tmem) {
    const int ni = 64 * LEN * 4, reps = 5;
    auto kern = a_tmem ? k<LEN, N0, D0, N1, D1, N2, D2, N3, D3, true> : k<LEN, N0, D0, N1, D1, N2, D2, N3, D3, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    kern<<<grid, 128, 49152>>>(ni, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e));
      return;
    }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const int NS[4] = {N0, N1, N2, N3};
    double floor_c = 0;
    for (int i = 0; i < LEN; ++i) floor_c += NS[i] / 2.0;
    floor_c /= LEN;
    printf("%-58s A=%s grid=%3d  issue %6.1f  total %6.1f cycles per instruction (floor N/2 = %5.1f)\n", name,
           a_tmem ? "tmem" : "smem", grid, (double)h[0] / ni, (double)h[1] / ni, floor_c);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<1, 64, 0, 0, 0, 0, 0, 0, 0>("N= 64, one accumulator", 1, d);
  run<2, 64, 0, 64, 64, 0, 0, 0, 0>("N= 64, two accumulators", 1, d);
  run<4, 64, 0, 64, 64, 64, 128, 64, 192>("N= 64, four accumulators", 1, d);
  run<1, 128, 0, 0, 0, 0, 0, 0, 0>("N=128, one accumulator", 1, d);
  run<2, 128, 0, 128, 128, 0, 0, 0, 0>("N=128, two accumulators", 1, d);
  run<1, 256, 0, 0, 0, 0, 0, 0, 0>("N=256, one accumulator", 1, d);
  run<2, 128, 0, 64, 0, 0, 0, 0, 0>("today: N=128 then N=64 into the same columns", 1, d);
  run<2, 128, 0, 64, 128, 0, 0, 0, 0>("N=128 then N=64 into its own columns", 1, d);
  run<3, 64, 0, 64, 64, 64, 128, 0, 0>("three N=64 into three accumulators", 1, d);
  run<4, 128, 0, 64, 128, 128, 192, 64, 320>("even/odd K-steps: (128, 64) x 2, four accumulators", 1, d);
  run<2, 128, 0, 64, 0, 0, 0, 0, 0>("today's pattern on every SM", 148, d);
  run<3, 64, 0, 64, 64, 64, 128, 0, 0>("three N=64 on every SM", 148, d);
  return 0;
}
