// tcgen05 / TMEM attention forward for ViT token counts (N <= 208, head_dim 64) on sm_100a.
//   S = scale * q k^T + rel_pos_bias[h] ; P = softmax_j(S) ; P~ = dropout(P) ; O = P~ v        (Attention.forward, modeling_finetune.py:145-188)
//
// One CTA per (batch, head, 128-query tile), two CTAs resident per SM (256 TMEM columns each):
//   warp 4, one lane : TMA loads (Q tile, K, V as 3-D boxes of the un-permuted [B, N, 3, H, 64] QKV GEMM output — rows past N are
//                      zero-filled by the TMA unit; the log2(e)-prescaled, -inf padded bias tile streams through a 2-stage ring of
//                      [128 x 32] fp32 SWIZZLE_128B boxes), tcgen05.mma S = Q K^T (128 x n_pad x 64, accumulator in TMEM columns
//                      [0, n_pad)), later tcgen05.mma O = P~ V with the A operand read straight from TENSOR MEMORY (the bf16
//                      probabilities overlay the S columns they were computed from) and V as an MN-major smem operand.
//   warps 0..3       : one thread per query row (TMEM lane == row): no shuffles, no shared-memory round trip for the scores.
//                      pass 1: s2 = s * scale*log2e + bias, running max, s2 written back to TMEM;
//                      pass 2: p = 2^(s2 - max), row sum, Philox4x32-7 dropout (or injected mask), keep bits packed for the
//                      backward, bf16 pairs stored to TMEM columns [0, n_pad/2);
//                      epilogue: O (TMEM columns [128, 192)) * 1/((1-p) * rowsum) -> bf16 -> swizzled smem tile -> TMA store
//                      (rows past N clipped by the tensor map), log-sum-exp to global.
#include <mutex>

#include "../../include/b200vit.h"
#include "attn_common.cuh"
#include "ptx_sm100.cuh"

namespace {

using namespace attn;

constexpr int TILE_M = 128;
constexpr int BIAS_STAGES = 2;
constexpr int BIAS_STAGE_BYTES = TILE_M * 128;           // 128 rows x 32 fp32
constexpr int SM_Q = 0;                                   // 128 rows x 128 B (re-used as the bf16 O staging tile)
constexpr int SM_K = SM_Q + TILE_M * 128;                 // NMAX rows x 128 B
constexpr int SM_V = SM_K + NMAX * 128;
constexpr int SM_BIAS = SM_V + NMAX * 128;
constexpr int SM_BAR = SM_BIAS + BIAS_STAGES * BIAS_STAGE_BYTES;
constexpr int FWD100_SMEM = SM_BAR + 128 + 1024;          // + barriers + 1024-byte alignment slack
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 128;                                // O accumulator: TMEM columns [128, 192)
constexpr int SOFTMAX_WARPS = 4;
constexpr int FWD100_THREADS = (SOFTMAX_WARPS + 1) * 32;
static_assert(SM_K % 1024 == 0 && SM_V % 1024 == 0 && SM_BIAS % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
static_assert(2 * FWD100_SMEM + 2048 <= 232448, "two CTAs per SM");

struct Fwd100Params {
  float* lse;               // [B, H, N]
  uint8_t* keep_bits;       // [B, H, N, 32]
  const uint8_t* keep_in;   // [B, H, N, N] or null
  int B, H, N, n_pad, m_tiles;
  float sl2, inv_keep;
  uint32_t thresh;
  uint64_t seed;
  uint32_t stream_id;
};

// Keep bits of the 32 (or 16) keys [32c, 32c + COLS) of query row i — bit e = key 32c + e. Same Philox stream layout as the
// mma.sync kernels (dropout_group / dropout_u16 in attn_common.cuh), so b200vit_dropout_mask and the backward agree.
template <int COLS>
__device__ __forceinline__ uint32_t keep_word(const Fwd100Params& p, int bh, int i, int c) {
  uint32_t w = 0u;
  if (p.keep_in == nullptr) {
#pragma unroll
    for (int quad = 0; quad < 4; ++quad) {
      const Philox4 r = dropout_group(p.seed, p.stream_id, bh, i, quad, c);
#pragma unroll
      for (int n4 = 0; n4 < COLS / 8; ++n4) {
        w |= (dropout_u16(r, n4 * 2) >= p.thresh ? 1u : 0u) << (n4 * 8 + quad * 2);
        w |= (dropout_u16(r, n4 * 2 + 1) >= p.thresh ? 1u : 0u) << (n4 * 8 + quad * 2 + 1);
      }
    }
  } else if (i < p.N) {
    const uint8_t* src = p.keep_in + ((long long)bh * p.N + i) * p.N;
    for (int e = 0; e < COLS; ++e) {
      const int j = c * 32 + e;
      if (j < p.N && src[j]) w |= 1u << e;
    }
  }
  return w;
}

template <int COLS, bool HAS_BIAS>
__device__ __forceinline__ void pass1_chunk(const Fwd100Params& p, uint32_t taddr, const uint8_t* bias_row, int row, int c, float& mx) {
  uint32_t s[COLS];
  if constexpr (COLS == 32) ptx::tmem_ld_x32_sync(taddr, reinterpret_cast<uint32_t(&)[32]>(s));
  else ptx::tmem_ld_x16_sync(taddr, reinterpret_cast<uint32_t(&)[16]>(s));
#pragma unroll
  for (int q = 0; q < COLS / 4; ++q) {
    float v[4];
    if (HAS_BIAS) {   // bias is pre-multiplied by log2(e); padding columns hold -inf (key mask); SWIZZLE_128B: 16-byte chunk ^ (row & 7)
      const float4 b4 = *reinterpret_cast<const float4*>(bias_row + ((q ^ (row & 7)) << 4));
      v[0] = fmaf(__uint_as_float(s[4 * q]), p.sl2, b4.x); v[1] = fmaf(__uint_as_float(s[4 * q + 1]), p.sl2, b4.y);
      v[2] = fmaf(__uint_as_float(s[4 * q + 2]), p.sl2, b4.z); v[3] = fmaf(__uint_as_float(s[4 * q + 3]), p.sl2, b4.w);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (c * 32 + 4 * q + e) < p.N ? __uint_as_float(s[4 * q + e]) * p.sl2 : -INFINITY;
    }
    mx = fmaxf(mx, fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])));
#pragma unroll
    for (int e = 0; e < 4; ++e) s[4 * q + e] = __float_as_uint(v[e]);
  }
  if constexpr (COLS == 32) ptx::tmem_st_x32(taddr, reinterpret_cast<const uint32_t(&)[32]>(s));
  else ptx::tmem_st_x16(taddr, reinterpret_cast<const uint32_t(&)[16]>(s));
}

template <int COLS, bool DROP>
__device__ __forceinline__ void pass2_chunk(const Fwd100Params& p, uint32_t trow, int bh, int i, int c, float mx, float& l) {
  uint32_t s[COLS];
  if constexpr (COLS == 32) ptx::tmem_ld_x32_sync(trow + c * 32, reinterpret_cast<uint32_t(&)[32]>(s));
  else ptx::tmem_ld_x16_sync(trow + c * 32, reinterpret_cast<uint32_t(&)[16]>(s));
  float pr[COLS];
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < COLS; ++e) {
    pr[e] = ex2(__uint_as_float(s[e]) - mx);
    acc += pr[e];
  }
  l += acc;
  if (DROP) {
    const uint32_t w = keep_word<COLS>(p, bh, i, c);
#pragma unroll
    for (int e = 0; e < COLS; ++e)
      if (!((w >> e) & 1u)) pr[e] = 0.f;
    if (i < p.N) *reinterpret_cast<uint32_t*>(p.keep_bits + ((long long)bh * p.N + i) * 32 + c * 4) = w;
  }
  uint32_t pk[COLS / 2];
#pragma unroll
  for (int e = 0; e < COLS / 2; ++e) pk[e] = pack_bf16x2(pr[2 * e], pr[2 * e + 1]);
  if constexpr (COLS == 32) ptx::tmem_st_x16(trow + c * 16, reinterpret_cast<const uint32_t(&)[16]>(pk));
  else ptx::tmem_st_x8(trow + c * 16, reinterpret_cast<const uint32_t(&)[8]>(pk));
}

template <bool DROP, bool HAS_BIAS>
__global__ void __launch_bounds__(FWD100_THREADS, 2)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                      const __grid_constant__ CUtensorMap tm_bias, const __grid_constant__ CUtensorMap tm_out, const Fwd100Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar_qk = base + SM_BAR, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_p = bar_qk + 24, bar_o = bar_qk + 32;
  auto bias_full = [&](int s) { return bar_qk + 40u + 8u * s; };
  auto bias_empty = [&](int s) { return bar_qk + 40u + 8u * (BIAS_STAGES + s); };
  const uint32_t tmem_slot = bar_qk + 40u + 16u * BIAS_STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x % p.m_tiles, bh = blockIdx.x / p.m_tiles;
  const int b = bh / p.H, h = bh - b * p.H;
  const int m0 = mt * TILE_M;
  const int n_pad = p.n_pad;
  const int nchunks = (n_pad + 31) >> 5;
  const int tail_cols = n_pad - (nchunks - 1) * 32;   // 16 or 32

  if (warp == SOFTMAX_WARPS) {
    if (lane == 0) {
      ptx::mbar_init(bar_qk, 1); ptx::mbar_init(bar_v, 1); ptx::mbar_init(bar_s, 1);
      ptx::mbar_init(bar_p, SOFTMAX_WARPS * 32); ptx::mbar_init(bar_o, 1);
      for (int s = 0; s < BIAS_STAGES; ++s) { ptx::mbar_init(bias_full(s), 1); ptx::mbar_init(bias_empty(s), SOFTMAX_WARPS); }
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_kv); ptx::prefetch_tmap(&tm_out);
      if (HAS_BIAS) ptx::prefetch_tmap(&tm_bias);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == SOFTMAX_WARPS) {
    if (lane == 0) {
      // ---------------- TMA producer + MMA issuer ----------------
      ptx::mbar_arrive_expect_tx(bar_qk, TILE_M * 128 + n_pad * 128);
      ptx::tma_load_3d(base + SM_Q, &tm_q, bar_qk, h * HD, m0, b);
      ptx::tma_load_3d(base + SM_K, &tm_kv, bar_qk, (p.H + h) * HD, 0, b);
      ptx::mbar_arrive_expect_tx(bar_v, n_pad * 128);
      ptx::tma_load_3d(base + SM_V, &tm_kv, bar_v, (2 * p.H + h) * HD, 0, b);
      if (HAS_BIAS) {
        for (int c = 0; c < nchunks && c < BIAS_STAGES; ++c) {
          ptx::mbar_arrive_expect_tx(bias_full(c), BIAS_STAGE_BYTES);
          ptx::tma_load_3d(base + SM_BIAS + c * BIAS_STAGE_BYTES, &tm_bias, bias_full(c), c * 32, m0, h);
        }
      }
      ptx::mbar_wait(bar_qk, 0);
      ptx::tc_fence_after();
      const uint32_t idesc_s = ptx::make_idesc_bf16(TILE_M, n_pad, false, false);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        ptx::umma_bf16(tmem_base, ptx::make_smem_desc(base + SM_Q + k * 32, 16, 1024), ptx::make_smem_desc(base + SM_K + k * 32, 16, 1024),
                       idesc_s, k > 0 ? 1u : 0u);
      ptx::umma_commit(bar_s);
      if (HAS_BIAS) {
        for (int c = BIAS_STAGES; c < nchunks; ++c) {
          const int s = c % BIAS_STAGES, use = c / BIAS_STAGES;
          ptx::mbar_wait(bias_empty(s), (uint32_t)((use - 1) & 1));
          ptx::mbar_arrive_expect_tx(bias_full(s), BIAS_STAGE_BYTES);
          ptx::tma_load_3d(base + SM_BIAS + s * BIAS_STAGE_BYTES, &tm_bias, bias_full(s), c * 32, m0, h);
        }
      }
      // O = P~ V : A = bf16 probabilities in TMEM columns [0, n_pad/2), B = V (MN-major: key rows of 64 d-elements)
      ptx::mbar_wait(bar_p, 0);
      ptx::mbar_wait(bar_v, 0);
      ptx::tc_fence_after();
      const uint32_t idesc_o = ptx::make_idesc_bf16(TILE_M, HD, false, true);
      for (int kk = 0; kk < n_pad / 16; ++kk)
        ptx::umma_bf16_ts(tmem_base + O_COL, tmem_base + kk * 8, ptx::make_smem_desc(base + SM_V + kk * 2048, NMAX * 128, 1024), idesc_o,
                          kk > 0 ? 1u : 0u);
      ptx::umma_commit(bar_o);
    }
  } else {
    // ---------------- softmax warps: one thread per query row ----------------
    const int row = warp * 32 + lane;
    const int i = m0 + row;
    const bool active = m0 + warp * 32 < p.N;            // warps whose 32 rows are all past N only keep the barrier protocol alive
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % BIAS_STAGES;
      if (HAS_BIAS) ptx::mbar_wait(bias_full(s), (uint32_t)((c / BIAS_STAGES) & 1));
      if (active) {
        const uint8_t* bias_row = gbase + SM_BIAS + s * BIAS_STAGE_BYTES + row * 128;
        if (c + 1 < nchunks || tail_cols == 32) pass1_chunk<32, HAS_BIAS>(p, trow + c * 32, bias_row, row, c, mx);
        else pass1_chunk<16, HAS_BIAS>(p, trow + c * 32, bias_row, row, c, mx);
      }
      if (HAS_BIAS) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bias_empty(s));
      }
    }
    ptx::tmem_st_wait();
    float l = 0.f;
    if (active) {
      for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks || tail_cols == 32) pass2_chunk<32, DROP>(p, trow, bh, i, c, mx, l);
        else pass2_chunk<16, DROP>(p, trow, bh, i, c, mx, l);
      }
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    ptx::mbar_arrive(bar_p);
    // ---------------- epilogue ----------------
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();
    if (active) {
      const float inv = p.inv_keep / l;
      uint8_t* orow = gbase + SM_Q + row * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[32];
        ptx::tmem_ld_x32_sync(trow + O_COL + half * 32, o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + (((half * 4 + q) ^ (row & 7)) << 4)) = u;
        }
      }
      if (i < p.N && p.lse != nullptr) p.lse[(long long)bh * p.N + i] = (mx + log2f(l)) / LOG2E;
    }
    ptx::fence_proxy_async();
    ptx::named_bar_sync(1, SOFTMAX_WARPS * 32);
    if (warp == 0 && lane == 0) {
      ptx::tma_store_3d(&tm_out, base + SM_Q, h * HD, m0, b);
      ptx::bulk_commit();
      ptx::bulk_wait_read0();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == SOFTMAX_WARPS) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

// 3-D tiled map, SWIZZLE_128B: dims {d0 (contiguous), d1, d2}, strides in ELEMENTS of dims 1 and 2, box {b0, b1, 1}.
int make_tmap3(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
               uint32_t b0, uint32_t b1) {
  PFN_encodeTiled enc = encode_fn();
  if (enc == nullptr) {
    b200vit_set_error("attn: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -2;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * esize, s2 * esize};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200vit_set_error("attn: cuTensorMapEncodeTiled failed (%d) ptr=%p dims=%llu,%llu,%llu box=%u,%u", (int)r, ptr, (unsigned long long)d0,
                      (unsigned long long)d1, (unsigned long long)d2, b0, b1);
    return -3;
  }
  return 0;
}

template <bool DROP, bool HAS_BIAS>
cudaError_t launch_fwd100(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tb, const CUtensorMap& to, const Fwd100Params& p,
                          cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_sm100_kernel<DROP, HAS_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD100_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  attn_fwd_sm100_kernel<DROP, HAS_BIAS><<<p.B * p.H * p.m_tiles, FWD100_THREADS, FWD100_SMEM, stream>>>(tq, tkv, tb, to, p);
  return cudaGetLastError();
}

}  // namespace

extern "C" int b200vit_attn_fwd(const void* qkv, const float* bias, int64_t ld_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim,
                                float scale, float p_drop, uint64_t seed, uint32_t stream_id, const uint8_t* keep_in, void* out, float* lse,
                                uint8_t* keep_bits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv != nullptr && out != nullptr, "attn_fwd: null pointer");
  B200_CHECK_ARG(B > 0 && H > 0, "attn_fwd: bad B=%d H=%d", B, H);
  B200_CHECK_ARG(head_dim == HD, "attn_fwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "attn_fwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "attn_fwd: bad p_drop");
  B200_CHECK_ARG(p_drop == 0.f || keep_bits != nullptr, "attn_fwd: dropout needs the keep_bits buffer [B,H,N,32]");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "attn_fwd: qkv/out must be 16-byte aligned");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(bias == nullptr || (ld_bias >= n_pad && ld_bias % 4 == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
                 "attn_fwd: bias must be the padded layout of b200vit_rel_pos_bias ([H,N,ld], ld %% 4 == 0, ld >= %d, 16-byte aligned)", n_pad);
  Fwd100Params p;
  p.lse = lse; p.keep_bits = keep_bits; p.keep_in = keep_in; p.B = B; p.H = H; p.N = N; p.n_pad = n_pad; p.m_tiles = (N + TILE_M - 1) / TILE_M;
  p.sl2 = scale * LOG2E; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  p.thresh = (uint32_t)(p_drop * 65536.0f + 0.5f); p.seed = seed; p.stream_id = stream_id;
  const uint64_t row = 3ull * H * HD;
  CUtensorMap tq, tkv, tb, to;
  int rc;
  if ((rc = make_tmap3(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, row, N, B, row, row * N, HD, TILE_M))) return rc;
  if ((rc = make_tmap3(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, row, N, B, row, row * N, HD, n_pad))) return rc;
  if ((rc = make_tmap3(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, (uint64_t)H * HD, N, B, (uint64_t)H * HD, (uint64_t)H * HD * N, HD, TILE_M))) return rc;
  if (bias != nullptr) {
    if ((rc = make_tmap3(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, bias, ld_bias, N, H, ld_bias, (uint64_t)ld_bias * N, 32, TILE_M))) return rc;
  } else {
    tb = tq;
  }
  cudaError_t e;
  if (p_drop > 0.f) e = bias != nullptr ? launch_fwd100<true, true>(tq, tkv, tb, to, p, stream) : launch_fwd100<true, false>(tq, tkv, tb, to, p, stream);
  else e = bias != nullptr ? launch_fwd100<false, true>(tq, tkv, tb, to, p, stream) : launch_fwd100<false, false>(tq, tkv, tb, to, p, stream);
  if (e != cudaSuccess) { b200vit_set_error("attn_fwd: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
