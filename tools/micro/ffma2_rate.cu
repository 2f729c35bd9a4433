// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, register operands, 16 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu ; prints FMA / clk / SM for both.
#include <cstdio>
#include <cuda_runtime.h>
template <bool PACKED>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) {
        unsigned long long d, x = *reinterpret_cast<unsigned long long*>(&acc[i]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(*reinterpret_cast<const unsigned long long*>(&aa)),
                     "l"(*reinterpret_cast<const unsigned long long*>(&bb)));
        acc[i] = *reinterpret_cast<float2*>(&d);
      } else {
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(acc[i].x) : "f"(acc[i].x), "f"(aa.x), "f"(bb.x));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(acc[i].y) : "f"(acc[i].y), "f"(aa.y), "f"(bb.y));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
  const int iters = 1 << 16;
  for (int packed = 0; packed < 2; ++packed) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (packed) k<true><<<sms * 8, 256>>>(out, 0.999f, 1e-3f, iters); else k<false><<<sms * 8, 256>>>(out, 0.999f, 1e-3f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)sms * 8 * 256 * iters * 16;
      if (rep == 2) printf("%s: %.3f ms  %.1f FMA/ns/SM  (at max clock %.0f MHz: %.1f FMA/clk/SM)\n", packed ? "FFMA2" : "FFMA ", ms, fma / (ms * 1e6) / sms,
                           clk_khz / 1e3, fma / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
