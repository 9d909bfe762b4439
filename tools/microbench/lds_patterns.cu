// Shared-memory wavefront cost of 128-bit loads under different lane -> address patterns (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int pattern, int iters, long long* out, float* sink) {
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int off;  // float offset of this lane's 16-byte word
  switch (pattern) {
    case 0: off = 0; break;                                  // warp-uniform
    case 1: off = lane * 4; break;                           // 32 distinct, contiguous (512 B)
    case 2: off = (lane & 3) * 36; break;                    // 4 distinct (stride 36 floats), every quarter-warp reads all 4
    case 3: off = (lane >> 2) * 4; break;                    // 8 distinct contiguous (128 B), 4 neighbouring lanes share one
    case 4: off = (lane >> 3) * 4; break;                    // 4 distinct, one per quarter-warp
    case 5: off = (lane & 7) * 36; break;                    // 8 distinct (stride 36), every quarter-warp reads all 8 (128 B)
    case 6: off = (lane & 15) * 36; break;                   // 16 distinct (256 B), halves identical
    case 7: off = (lane & 7) * 4 + (lane >> 3) * 1024; break; // 32 distinct, quarter q in row q: same banks across quarters
    default: off = lane * 36; break;                         // 32 distinct, stride 36 floats
  }
  float4 acc = make_float4(0, 0, 0, 0);
  const float* p = s + off;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p + ((u * 64 + (i & 15) * 4) & 1023))));
      acc.x += v.x;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
  const char* names[] = {"uniform", "32 distinct contiguous", "4 distinct, all in every quarter", "8 distinct contiguous, 4 lanes share",
                         "4 distinct, one per quarter", "8 distinct stride 36, every quarter all 8", "16 distinct stride 36, halves identical",
                         "32 distinct, quarters on the same banks", "32 distinct stride 36"};
  for (int nw = 32; nw <= 32; nw *= 4)
    for (int pat = 0; pat < 9; ++pat) {
      k<<<1, 32 * nw>>>(pat, 1, d, sink);
      k<<<1, 32 * nw>>>(pat, 256, d, sink);
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("warps %d  %-45s %.2f SM cycles per warp-wide LDS.128\n", nw, names[pat], (double)h / (256.0 * 16) / nw);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
