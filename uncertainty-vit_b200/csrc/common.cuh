// Shared device/host helpers for libb200vit (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef __CUDA_ARCH__
#define B200VIT_HOST 1
#endif

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------
// error plumbing (api.cu owns the storage)
// ----------------------------------------------------------------------------------------
void b200vit_set_error(const char* fmt, ...);
#define B200_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      b200vit_set_error(__VA_ARGS__);        \
      return -1;                             \
    }                                        \
  } while (0)
#define B200_CHECK_LAUNCH(name)                                              \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      b200vit_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                       \
    }                                                                        \
  } while (0)

int b200vit_num_sms();

// ----------------------------------------------------------------------------------------
// small device utilities
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// erf-GELU (torch.nn.GELU(), reference modeling_finetune.py:65-82) and its derivative for the bf16 GEMM epilogues.
// The epilogue warps are bound by the SFU (16 MUFU/clk/SM) before they are bound by issue slots, so Phi(x) - 1/2 is a degree-17
// odd minimax-style polynomial on [-4, 4] (input clamped): FMA pipe only. |gelu err| <= 1.9e-5 for |x| <= 4 and <= 2.9e-5 * |x|
// beyond (true tails are < 3.2e-5) — two orders of magnitude below the bf16 resolution of the tensors it feeds.
// The derivative needs exp(-x^2/2) as well: one MUFU.EX2.
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float phi_minus_half(float x) {
  const float xc = fminf(fmaxf(x, -4.0f), 4.0f);
  const float u = xc * xc;
  float p = 7.804711256e-11f;
  p = fmaf(p, u, -6.827683748e-09f);
  p = fmaf(p, u, 2.666981721e-07f);
  p = fmaf(p, u, -6.222018266e-06f);
  p = fmaf(p, u, 9.829133151e-05f);
  p = fmaf(p, u, -1.130966313e-03f);
  p = fmaf(p, u, 9.869967510e-03f);
  p = fmaf(p, u, -6.640203406e-02f);
  p = fmaf(p, u, 3.989198652e-01f);
  return p * xc;
}
__device__ __forceinline__ float gelu_erf(float x) {
  return fmaf(x, phi_minus_half(x), 0.5f * x);           // x * Phi(x)
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float e = mufu_ex2(x * (x * -0.72134752044448170f));   // exp(-x^2/2)
  return fmaf(x * 0.39894228040143268f, e, 0.5f + phi_minus_half(x));   // Phi(x) + x * phi(x)
}

// Philox4x32-10 (counter-based; Salmon et al. 2011). Same constants as cuRAND / torch.
struct Philox4 {
  uint32_t x, y, z, w;
};
template <int ROUNDS>
__host__ __device__ __forceinline__ Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}
