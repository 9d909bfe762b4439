// fp32 FMA issue-rate microbenchmark for sm_100a: scalar FFMA vs packed FFMA2, register vs constant-bank operands.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu ; run: ./ffma_rate
#include <cuda_runtime.h>
#include <cstdio>
__constant__ float2 cw[64];
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE>
__global__ void k(float* out, const float* in, int iters, int cbase) {
  float2 acc[16];
  float2 w[8];
  float x[4];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int i = 0; i < 8; ++i) w[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 8 + i]);
  for (int i = 0; i < 4; ++i) x[i] = in[threadIdx.x + 32 + i];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) {            // scalar FFMA, 3 registers
          acc[i].x = fmaf(x[r], w[i & 7].x, acc[i].x);
          acc[i].y = fmaf(x[r], w[i & 7].y, acc[i].y);
        } else if (MODE == 1) {     // FFMA2, scalar broadcast A, register pair B
          ffma2(acc[i], make_float2(x[r], x[r]), w[i & 7]);
        } else if (MODE == 2) {     // FFMA2, pair A, pair B
          ffma2(acc[i], make_float2(x[r], x[(r + 1) & 3]), w[i & 7]);
        } else if (MODE == 3) {     // scalar FFMA, B from the constant bank (uniform index)
          acc[i].x = fmaf(x[r], cw[cbase + (i & 7)].x, acc[i].x);
          acc[i].y = fmaf(x[r], cw[cbase + (i & 7)].y, acc[i].y);
        } else {                    // FFMA2, scalar broadcast A, B from the constant bank
          ffma2(acc[i], make_float2(x[r], x[r]), cw[cbase + (i & 7)]);
        }
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int threads, int sms, float ghz) {
  float *out, *in;
  cudaMalloc(&out, sizeof(float) * sms * threads);
  cudaMalloc(&in, sizeof(float) * 2048);
  cudaMemset(in, 0, sizeof(float) * 2048);
  const int iters = 20000;
  k<MODE><<<sms, threads>>>(out, in, 10, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms, threads>>>(out, in, iters, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)iters * 4 * 16 * 2 * threads;          // per SM
  printf("%-48s %4d thr/SM: %7.1f FMA/clk/SM (%.3f ms, err %d)\n", name, threads, fma / (ms * 1e-3 * ghz * 1e9), ms, (int)cudaGetLastError());
  cudaFree(out); cudaFree(in);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const float ghz = p.clockRate * 1e-6f;
  printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  float2 h[64]; for (int i = 0; i < 64; ++i) h[i] = make_float2(0.f, 0.f);
  cudaMemcpyToSymbol(cw, h, sizeof(h));
  for (int t : {128, 256, 512, 1024}) {
    run<0>("FFMA  3 registers", t, p.multiProcessorCount, ghz);
    run<1>("FFMA2 scalar-broadcast A, register-pair B", t, p.multiProcessorCount, ghz);
    run<2>("FFMA2 register-pair A and B", t, p.multiProcessorCount, ghz);
    run<3>("FFMA  B from constant bank", t, p.multiProcessorCount, ghz);
    run<4>("FFMA2 scalar-broadcast A, B from constant bank", t, p.multiProcessorCount, ghz);
  }
  return 0;
}
