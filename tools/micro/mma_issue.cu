// Microbenchmark: cost of issuing tcgen05.mma / tcgen05.commit / mbarrier waits from ONE thread (small attention-sized shapes).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../uncertainty-vit_b200/csrc/ptx_sm100.cuh"

template <int MODE>   // 0: SS mma N=32 back-to-back ; 1: SS N=208 ; 2: TS mma N=64 ; 3: SS N=32 + commit + wait each 12 ; 4: same with 16 spinning warps
__global__ void k(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bars[2];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  if (threadIdx.x == 32) { ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::mbar_init(ptx::smem_u32(&bars[1]), 1); ptx::fence_barrier_init(); }
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - ptx::smem_u32(smem_raw)))[i] = 0;
  ptx::fence_proxy_async();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t bar = ptx::smem_u32(&bars[0]), bar2 = ptx::smem_u32(&bars[1]);
  if (warp == 1 && lane == 0) {
    const uint64_t da = ptx::make_smem_desc(base, 16, 1024), db = ptx::make_smem_desc(base + 32768, 16, 1024);
    const uint64_t dbm = ptx::make_smem_desc(base + 32768, 4096, 1024);
    const uint32_t id32 = ptx::make_idesc_bf16(128, 32, false, false), id208 = ptx::make_idesc_bf16(128, 208, false, false);
    const uint32_t idts = ptx::make_idesc_bf16(128, 64, false, true);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) ptx::umma_bf16(tmem, da + 2 * (it & 3), db + 2 * (it & 3), id32, (it & 3) ? 1u : 0u);
      if (MODE == 1) ptx::umma_bf16(tmem, da + 2 * (it & 3), db + 2 * (it & 3), id208, (it & 3) ? 1u : 0u);
      if (MODE == 2) ptx::umma_bf16_ts(tmem + 256, tmem + 8 * (it & 1), dbm + 128 * (it & 1), idts, 1u);
      if (MODE >= 3) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, da + 2 * k, db + 2 * k, id32, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 64, da + 2 * k, db + 2 * k, id32, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16_ts(tmem + 256, tmem + 8 * (k & 1), dbm + 128 * (k & 1), idts, 1u);
        ptx::umma_commit(bar);
        ptx::mbar_wait(bar, phase);
        phase ^= 1u;
      }
    }
    ptx::umma_commit(bar2);
    ptx::mbar_wait(bar2, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    bars[1] = 0xdeadbeefULL;   // release the spinners (plain flag reuse below)
  } else if (MODE == 4 && warp >= 2) {
    // spinning warps: poll a barrier that never completes until the issuer is done (emulates waiting element-wise warps)
    volatile uint64_t* flag = &bars[1];
    while (*flag != 0xdeadbeefULL) { ptx::mbar_try_wait(bar2 + 0, 1); }
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 256);
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 64, smem>>>(iters, d);
      if (mode == 1) k<1><<<148, 64, smem>>>(iters, d);
      if (mode == 2) k<2><<<148, 64, smem>>>(iters, d);
      if (mode == 3) k<3><<<148, 64, smem>>>(iters, d);
      if (mode == 4) k<4><<<148, 64 + 16 * 32, smem>>>(iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d error %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const char* names[] = {"SS mma 128x32x16 back-to-back", "SS mma 128x208x16 back-to-back", "TS mma 128x64x16 back-to-back",
                           "round: 8 SS + 4 TS + commit + wait", "round (same) with 16 spinning warps"};
    printf("mode %d (%s): %.1f cycles per iteration\n", mode, names[mode], (double)h / iters);
  }
  return 0;
}
