// calibration: 32-bit shared loads with 1-, 2-, 4-, 32-way bank conflicts, and 128-bit loads, 32 warps
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC>
__global__ void k(int stride, int iters, long long* out, float* sink) {
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float* p = s + (lane * stride) % 4096;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const unsigned a = (unsigned)__cvta_generic_to_shared(p + ((u * 64 + (i & 15) * 4) & 1023));
      if (VEC == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); acc += v.x; }
      else if (VEC == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); acc += v.x; }
      else { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); acc += v; }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 8192);
  const int nw = 32;
  int strides[] = {0, 1, 2, 4, 8, 32, 36};
  for (int vec = 1; vec <= 4; vec *= 2)
    for (int s : strides) {
      if ((s % vec) && s) continue;
      for (int rep = 0; rep < 2; ++rep) {
        if (vec == 1) k<1><<<1, 32 * nw>>>(s, 256, d, sink);
        else if (vec == 2) k<2><<<1, 32 * nw>>>(s, 256, d, sink);
        else k<4><<<1, 32 * nw>>>(s, 256, d, sink);
      }
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("LDS.%d lane stride %2d floats: %.2f SM cycles per warp-wide load (block time / loads)\n", 32 * vec, s, (double)h / (256.0 * 16) / nw);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
