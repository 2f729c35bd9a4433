// Shared device helpers of the attention kernels (attention.cu: dot-product attention; wattention.cu: Wasserstein attention).
#pragma once
#include "common.cuh"

int b200vit_relbias_grad_launch(const void* ds_work, int B, int H, int N, int ld_ds, const int32_t* rel_index, float* dtable, void* stream);
int b200vit_keep_bits_launch(uint8_t* keep_bits, int BH, int N, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id,
                             const uint8_t* keep_in, void* stream);

namespace attn {

constexpr int HD = 64;            // head dim
constexpr int PITCH = HD + 8;     // smem row pitch in bf16 (144 B: conflict-free ldmatrix)
constexpr int NMAX = 208;         // 13 tiles of 16
constexpr int DSP = NMAX + 8;     // bf16 row pitch of the shared dS^T matrix (432 B = 27 x 16 B: conflict-free ldmatrix)
constexpr int FWD_WARPS = 8;
constexpr int BWD_WARPS = 13;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {   // 2^x, one MUFU; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Loads rows [0,N) x 64 bf16 of a [*, row_stride] global matrix into smem [NMAX][PITCH]; rows >= N zero-filled up to n_pad.
__device__ __forceinline__ void load_tile_rows(bf16* s, const bf16* g, long long row_stride, int N, int n_pad) {
  for (int i = threadIdx.x; i < n_pad * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    if (r < N) cp_async16(smem_u32(s + r * PITCH + c), g + (long long)r * row_stride + c);
    else *reinterpret_cast<uint4*>(s + r * PITCH + c) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Dropout stream: keep-bit of element (i, j) of head bh comes from a 16-bit lane of Philox4x32-7 (the Crush-resistant
// round count of Salmon et al.); one call yields the 8 values a thread owns in four consecutive 8-key tiles of one row.
__device__ __forceinline__ Philox4 dropout_group(uint64_t seed, uint32_t stream, uint32_t bh, uint32_t i, uint32_t quad, uint32_t group) {
  return philox4x32<7>(bh, i, quad * 8u + group, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ uint32_t dropout_u16(const Philox4& r, int idx /*0..7, compile-time*/) {
  const uint32_t w = idx < 4 ? (idx < 2 ? r.x : r.y) : (idx < 6 ? r.z : r.w);
  return (idx & 1) ? (w >> 16) : (w & 0xffffu);
}


}  // namespace attn
