// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (one CTA, W warps, each warp on its own 32-lane quadrant).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../uncertainty-vit_b200/csrc/ptx_sm100.cuh"

template <int MODE>   // 0: ld x32 + wait each ; 1: 4 x ld x32 then one wait ; 2: st x32 (wait at end) ; 3: ld x16+wait
__global__ void k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 32 + (warp >> 2) * 128) & 511) & ~31u;
    if (MODE == 0) {
      ptx::tmem_ld_x32_sync(base + col, r);
      acc += r[0] ^ r[31];
    } else if (MODE == 1) {
      uint32_t a[32], b[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(a[0]),"=r"(a[1]),"=r"(a[2]),"=r"(a[3]),"=r"(a[4]),"=r"(a[5]),"=r"(a[6]),"=r"(a[7]),"=r"(a[8]),"=r"(a[9]),"=r"(a[10]),"=r"(a[11]),"=r"(a[12]),"=r"(a[13]),"=r"(a[14]),"=r"(a[15]),
          "=r"(a[16]),"=r"(a[17]),"=r"(a[18]),"=r"(a[19]),"=r"(a[20]),"=r"(a[21]),"=r"(a[22]),"=r"(a[23]),"=r"(a[24]),"=r"(a[25]),"=r"(a[26]),"=r"(a[27]),"=r"(a[28]),"=r"(a[29]),"=r"(a[30]),"=r"(a[31]) : "r"(base + col));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
        : "=r"(b[0]),"=r"(b[1]),"=r"(b[2]),"=r"(b[3]),"=r"(b[4]),"=r"(b[5]),"=r"(b[6]),"=r"(b[7]),"=r"(b[8]),"=r"(b[9]),"=r"(b[10]),"=r"(b[11]),"=r"(b[12]),"=r"(b[13]),"=r"(b[14]),"=r"(b[15]),
          "=r"(b[16]),"=r"(b[17]),"=r"(b[18]),"=r"(b[19]),"=r"(b[20]),"=r"(b[21]),"=r"(b[22]),"=r"(b[23]),"=r"(b[24]),"=r"(b[25]),"=r"(b[26]),"=r"(b[27]),"=r"(b[28]),"=r"(b[29]),"=r"(b[30]),"=r"(b[31]) : "r"(base + ((col + 32) & 511)));
      acc += a[0] ^ b[31];
    } else if (MODE == 2) {
      r[0] = acc + it;
      ptx::tmem_st_x32(base + col, r);
    } else {
      uint32_t a[16];
      ptx::tmem_ld_x16_sync(base + col, a);
      acc += a[0] ^ a[15];
    }
  }
  if (MODE == 2) ptx::tmem_st_wait();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 8 * 256); cudaMalloc(&s, 4 * 1024 * 256);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(iters, d, s);
        if (mode == 1) k<1><<<148, warps * 32>>>(iters, d, s);
        if (mode == 2) k<2><<<148, warps * 32>>>(iters, d, s);
        if (mode == 3) k<3><<<148, warps * 32>>>(iters, d, s);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double per_instr = (mode == 1 ? 2.0 : 1.0) * (mode == 3 ? 0.5 : 1.0);
      const double bytes = (double)iters * warps * 4096.0 * per_instr;
      printf("warps=%2d mode=%d (%s): %lld cyc  -> %.1f B/cyc/SM, %.1f cyc per warp-instr\n", warps, mode,
             mode == 0 ? "ld.x32+wait" : mode == 1 ? "2x ld.x32, 1 wait" : mode == 2 ? "st.x32" : "ld.x16+wait", h, bytes / h,
             (double)h / (iters * (mode == 1 ? 2.0 : 1.0)));
    }
  }
  return 0;
}
