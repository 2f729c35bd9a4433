// Wasserstein-distance attention of the dual-stream ("--stochastic") transformer on tcgen05 / TMEM (sm_100a), forward and backward.
// dist Attention.forward (modeling_finetune_dist.py:111-179) + wasserstein_distance_matmul (uncertainty_evaluations.py:276-294):
//   m1 = sigmoid(scale q)  m2 = sigmoid(k)  u1 = sqrt(max(sigmoid(cq), 1e-24))  u2 = sqrt(max(sigmoid(ck), 1e-24))       (cq, ck = elu(.) + 1 > 0)
//   D_ij = |m1_i|^2 + |m2_j|^2 - 2 m1_i.m2_j + sum c1_i + sum c2_j - 2 u1_i.u2_j = r_i + c_j - 2 [m1_i | u1_i].[m2_j | u2_j]     (ONE K = 128 product)
//   A = sigmoid(-D + 1e-24) + rel_pos_bias ;  P = softmax_j(A) ;  P~ = dropout(P) ;  mean = P~ v ;  cov = (P~)^2 cv
//
// Kernels
//   wattn_prep_kernel   : element-wise, HBM-bound: X1 = [m1 | u1], X2 = [m2 | u2] as bf16 [B, N, 2, H, 128] and the row norms r_i, c_j (taken from the
//                         SAME bf16-rounded values the tensor cores multiply, halved). sigmoid = 0.5 + 0.5 tanh(x/2): 1 MUFU; sqrt(sigmoid(c)) = rsqrt(1 + e^-c): 2.
//                         The forward writes this workspace, the backward reads it (no transform is ever repeated).
//   wattn_fwd_kernel    : persistent, one CTA per SM. TMA: X1 tile, X2, V, CV (3-D boxes, rows past N zero-filled), bias ring. tcgen05: S = X1 X2^T
//                         (128 x n_pad x 128) into TMEM; EIGHT element-wise warps = two threads per query row (TMEM lane quadrant = warp & 3,
//                         column half = warp >> 2) run ONE pass: t = sigmoid(-D) = 0.5 - 0.5 tanh(D/2), p = 2^(t log2e + bias - m_i) with the stabiliser
//                         m_i = max_j bias_ij + log2e >= max_j A_ij (t < 1), so no running max and no rescaling; P~ and (P~)^2 go back to TMEM as
//                         bf16 and feed two TS-MMAs (A operand from tensor memory): O_m = P~ V, O_c = (P~)^2 CV; epilogue via smem + TMA store.
//   wattn_bwd_kv_kernel : key-tile owner (TMEM lane = key row), 64-query boxes, two ping-pong groups: S^T = X2 X1^T, Gm^T = V dOm^T, Gc^T = CV dOc^T
//                         (tcgen05), element-wise phase out of TMEM, P~^T and (P~^2)^T back to TMEM as the A operands of dV += P~^T dOm,
//                         dCV += (P~^2)^T dOc; dD^T (and dA^T for the bias-table gradient) streamed to the workspace.
//   wattn_bwd_dx_kernel : per (batch, head): dD^T is loaded ONCE and used both as a K-major A operand (key side: dD^T X1) and as an MN-major A
//                         operand (query side: dD X2) — the same shared-memory bytes under two descriptors; epilogue applies
//                         dm = 2 (rho m - acc), the sigmoid / sqrt / elu chain rules and writes dq, dk, dcq, dck.
// Backward math: dP~ = dOm.v + 2 P~ (dOc.cv) ; dA = P (f dP~ - Delta), Delta_i = dOm_i.Om_i + 2 dOc_i.Oc_i ;
//   dD = -dA t (1 - t) ; rho_i = sum_j dD_ij ; kappa_j = sum_i dD_ij ; dm1 = 2 (rho m1 - dD m2) ; du1 = 2 (rho u1 - dD u2) ; (same for the key side)
//   dq = scale dm1 m1 (1 - m1) ; dcq' = du1 u1 (1 - u1^2) / 2 ; d(pre-elu) = dcq' (cq' <= 1 ? cq' : 1).
#include <cstdlib>
#include <mutex>

#include "../../include/b200vit.h"
#include "attn_common.cuh"
#include "ptx_sm100.cuh"

namespace {

using namespace attn;

constexpr int TILE_M = 128;
constexpr int XW = 2 * HD;                      // 128 columns of [m | u]
constexpr int KT_BYTES = NMAX * 128;            // one [NMAX rows x 64 bf16] SWIZZLE_128B tile
constexpr int BIAS_STAGE_BYTES = TILE_M * 128;  // [128 rows x 32 fp32]

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {      // one MUFU; |error| ~ 2^-11, far below the bf16 resolution of what it feeds
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// mbarrier wait with an out-of-line slow path (bounded: a pipeline bug traps instead of hanging the GPU)
__device__ __noinline__ void w_wait_slow(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {
        printf("b200vit: mbarrier timeout (wattn, block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void w_wait(uint32_t bar, uint32_t parity) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  w_wait_slow(bar, parity);
}

// ------------------------------------------------------------------------------------------------
// prep: X = [sigmoid(s q) | rsqrt(1 + e^-cq)], [sigmoid(k) | rsqrt(1 + e^-ck)] and the squared row norms (x log2 e)
// ------------------------------------------------------------------------------------------------
// 16 lanes per (token, q|k, head) unit: lanes 0..7 transform the 64 mean elements (8 each), lanes 8..15 the 64 cov elements. A warp owns a
// token: its two half-warps take the even / odd units of the token's 2 H, four units (four 16-byte loads per lane) in flight at a time —
// one integer division per token instead of per unit, and enough bytes in flight that the first use of a load no longer waits (48 % of the
// stall samples of the one-unit-per-iteration version, 90 us).
constexpr int PREP_BATCH = 4;
__global__ void __launch_bounds__(256) wattn_prep_kernel(const bf16* __restrict__ qkv_m, const bf16* __restrict__ qkv_c, bf16* __restrict__ X,
                                                         float* __restrict__ rn, float* __restrict__ cn, int B, int H, int N, float scale) {
  const int l16 = threadIdx.x & 15, hw = (threadIdx.x >> 4) & 1;
  const bool is_cov = l16 >= 8;
  const int upt = 2 * H;                                     // units per token
  const bf16* src_base = (is_cov ? qkv_c : qkv_m) + (l16 & 7) * 8;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int tokens = B * N;
  for (int bn = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; bn < tokens; bn += warps) {
    const int b = bn / N, n = bn - b * N;
    const bf16* src = src_base + (size_t)bn * (3 * H * HD);
    for (int w0 = hw; w0 < upt; w0 += 2 * PREP_BATCH) {
      uint4 raw[PREP_BATCH];
#pragma unroll
      for (int k4 = 0; k4 < PREP_BATCH; ++k4) {
        const int wh = w0 + 2 * k4;                          // which * H + h : also the 64-column slot inside the q | k part of the qkv row
        if (wh < upt) raw[k4] = *reinterpret_cast<const uint4*>(src + wh * HD);
      }
#pragma unroll
      for (int k4 = 0; k4 < PREP_BATCH; ++k4) {
        const int wh = w0 + 2 * k4;
        if (wh >= upt) break;                                // warp-uniform (upt is even: both half-warps run the same iterations)
        const int which = wh >= H ? 1 : 0;
        const uint32_t* pr = &raw[k4].x;
        // mean half: sigmoid(a x) = 0.5 + 0.5 tanh(a x / 2) (ONE MUFU); cov half: sqrt(sigmoid(x)) = rsqrt(1 + 2^(-x log2e)) (two)
        // (sigmoid of elu + 1 > 0 is > 1/2: the 1e-24 clamp of the reference never binds)
        const float a = is_cov ? -LOG2E : 0.5f * (which == 0 ? scale : 1.0f);
        uint4 outv;
        uint32_t* po = &outv.x;
        float nrm = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 v = unpack_bf16x2(pr[k]);
          float y0, y1;
          if (is_cov) {
            y0 = rsqrt_approx(1.0f + ex2(v.x * a));
            y1 = rsqrt_approx(1.0f + ex2(v.y * a));
          } else {
            y0 = fmaf(tanh_approx(v.x * a), 0.5f, 0.5f);
            y1 = fmaf(tanh_approx(v.y * a), 0.5f, 0.5f);
          }
          po[k] = pack_bf16x2(y0, y1);
          const float2 r = unpack_bf16x2(po[k]);
          nrm += r.x * r.x + r.y * r.y;
        }
        const size_t u = (size_t)bn * upt + wh;
        *reinterpret_cast<uint4*>(X + u * XW + l16 * 8) = outv;          // X is [B N, 2, H, 128] = [unit, 128]
        nrm += __shfl_xor_sync(0xffffffffu, nrm, 8);
        nrm += __shfl_xor_sync(0xffffffffu, nrm, 4);
        nrm += __shfl_xor_sync(0xffffffffu, nrm, 2);
        nrm += __shfl_xor_sync(0xffffffffu, nrm, 1);
        if (l16 == 0) (which == 0 ? rn : cn)[(size_t)(b * H + (wh - which * H)) * N + n] = 0.5f * nrm;   // HALF squared norms: D / 2 = rn + cn - x1.x2
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int F_SM_Q = 0;                               // X1 tile: mean half [128 x 128 B], cov half
constexpr int F_SM_K = F_SM_Q + 2 * TILE_M * 128;       // X2: mean half [NMAX x 128 B], cov half
constexpr int F_SM_V = F_SM_K + 2 * KT_BYTES;           // V ; its first 16 KB double as the bf16 O_m staging tile of the epilogue
constexpr int F_SM_CV = F_SM_V + KT_BYTES;              // CV; staging tile of O_c
constexpr int F_BIAS_STAGES = 4;
constexpr int F_SM_BIAS = F_SM_CV + KT_BYTES;
constexpr int F_SM_CN = F_SM_BIAS + F_BIAS_STAGES * BIAS_STAGE_BYTES;   // 2 x NMAX floats: key norms of the item (x log2 e)
constexpr int F_SM_LS = F_SM_CN + 2048;                 // 2 x 128 floats: partial row sums of the two column halves
constexpr int F_SM_BAR = F_SM_LS + 1024;
constexpr int F_SMEM = F_SM_BAR + 1024 + 1024;
constexpr int F_EW_WARPS = 8;
constexpr int F_THREADS = (F_EW_WARPS + 2) * 32;
constexpr int F_T_PSQ = 208, F_T_OM = 320, F_T_OC = 384;   // TMEM columns: S / P~ [0, 208), (P~)^2 [208, 312), O_m [320, 384), O_c [384, 448)
static_assert(F_SM_K % 1024 == 0 && F_SM_V % 1024 == 0 && F_SM_CV % 1024 == 0 && F_SM_BIAS % 1024 == 0 && (F_SM_K + KT_BYTES) % 1024 == 0,
              "SWIZZLE_128B tiles need 1024-byte alignment");
static_assert(F_SMEM <= 232448, "forward kernel smem");

struct WFwdParams {
  const float* rn;          // [B, H, N] half squared query norms
  const float* cn;          // [B, H, N] half squared key norms
  const float* rowmax;      // [H, N] max_j bias_ij (log2 domain)
  float* lse;               // [B, H, N] or null
  uint8_t* keep_bits;       // [B, H, N, 32]
  const uint8_t* keep_in;   // [B, H, N, N] or null
  int B, H, N, n_pad, m_tiles, items;
  float inv_keep;
  uint32_t thresh;
  uint64_t seed;
  const uint64_t* seed_dev;
  uint32_t stream_id;
};

// keep bits of keys [32c, 32c + 32) of query row i (bit e = key 32c + e) from the packed mask b200vit_keep_bits_launch wrote before this kernel
__device__ __forceinline__ uint32_t w_keep_word(const WFwdParams& p, int bh, int i, int c) {
  return i < p.N ? __ldg(reinterpret_cast<const uint32_t*>(p.keep_bits) + ((long long)bh * p.N + i) * 8 + c) : 0u;
}

// NC (16 or 32) scores of one query row -> probabilities; P~ and (P~)^2 back to TMEM as bf16 pairs. Both 16-column TMEM loads of a 32-column
// chunk are in flight before the single wait. k0 = log2e / 2 - m_i folds the sigmoid's offset and the stabiliser into one constant.
template <bool DROP, int NC>
__device__ __forceinline__ void wfwd_chunk(uint32_t t_src, uint32_t t_p, uint32_t t_psq, const uint8_t* bias_row, int row, const float* cn,
                                           float rnh, float k0, float& l, uint32_t w) {
  uint32_t s[32];
  if constexpr (NC == 32) {
    ptx::tmem_ld_x16_pair_sync(t_src, reinterpret_cast<uint32_t(&)[16]>(s[0]), t_src + 16, reinterpret_cast<uint32_t(&)[16]>(s[16]));
  } else {
    ptx::tmem_ld_x16_sync(t_src, reinterpret_cast<uint32_t(&)[16]>(s[0]));
  }
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < NC / 4; ++q) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias_row + ((q ^ (row & 7)) << 4));
    const float4 c4 = *reinterpret_cast<const float4*>(cn + 4 * q);
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float hd = (rnh + cc[e]) - __uint_as_float(s[4 * q + e]);                          // D / 2
      const float th = tanh_approx(hd);                                                        // sigmoid(-D) = 0.5 - 0.5 tanh(D / 2)
      float pv = ex2(fmaf(th, -0.5f * LOG2E, bb[e] + k0));                                     // bias padding = -inf -> 0
      acc += pv;
      if (DROP && !(w & (1u << (4 * q + e)))) pv = 0.f;
      s[4 * q + e] = __float_as_uint(pv);
    }
  }
  l += acc;
#pragma unroll
  for (int hlf = 0; hlf < NC / 16; ++hlf) {
    uint32_t pk[8], pq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float p0 = __uint_as_float(s[hlf * 16 + 2 * e]), p1 = __uint_as_float(s[hlf * 16 + 2 * e + 1]);
      pk[e] = pack_bf16x2(p0, p1);
      pq[e] = pack_bf16x2(p0 * p0, p1 * p1);
    }
    ptx::tmem_st_x8(t_p + hlf * 8, pk);
    ptx::tmem_st_x8(t_psq + hlf * 8, pq);
  }
}

template <bool DROP>
__global__ void __launch_bounds__(F_THREADS, 1)
wattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2, const __grid_constant__ CUtensorMap tm_v,
                 const __grid_constant__ CUtensorMap tm_cv, const __grid_constant__ CUtensorMap tm_bias, const __grid_constant__ CUtensorMap tm_om,
                 const __grid_constant__ CUtensorMap tm_oc, const WFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;     // 0..7 element-wise, 8 MMA, 9 TMA
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = base + F_SM_BAR;
  const uint32_t qk_full = bar0, v_full = bar0 + 8, s_full = bar0 + 16, p_full = bar0 + 24, o_full = bar0 + 32, v_free = bar0 + 48, t_free = bar0 + 56;
  auto bias_full = [&](int s) { return bar0 + 64u + 8u * s; };
  auto bias_empty = [&](int s) { return bar0 + 96u + 8u * s; };
  const uint32_t tmem_slot = bar0 + 160u;

  const int n_pad = p.n_pad;
  const int nchunks = (n_pad + 31) >> 5;
  const int tail_cols = n_pad - (nchunks - 1) * 32;   // 16 or 32
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_items = first < p.items ? (p.items - first + stride - 1) / stride : 0;
  auto item_of = [&](int it, int& b, int& h, int& m0, int& bh) {
    const int item = first + it * stride;
    const int mt = item % p.m_tiles;
    bh = item / p.m_tiles;
    b = bh / p.H; h = bh - b * p.H; m0 = mt * TILE_M;
  };

  if (warp == F_EW_WARPS) {
    if (lane == 0) {
      ptx::mbar_init(qk_full, 1); ptx::mbar_init(v_full, 1); ptx::mbar_init(s_full, 1); ptx::mbar_init(p_full, F_EW_WARPS * 32);
      ptx::mbar_init(o_full, 2); ptx::mbar_init(v_free, 1); ptx::mbar_init(t_free, F_EW_WARPS);
      for (int s = 0; s < F_BIAS_STAGES; ++s) { ptx::mbar_init(bias_full(s), 1); ptx::mbar_init(bias_empty(s), F_EW_WARPS / 2); }
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tm_x1); ptx::prefetch_tmap(&tm_x2); ptx::prefetch_tmap(&tm_v); ptx::prefetch_tmap(&tm_cv);
      ptx::prefetch_tmap(&tm_bias); ptx::prefetch_tmap(&tm_om); ptx::prefetch_tmap(&tm_oc);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == F_EW_WARPS) {
    if (lane < 2 && n_items > 0) {
      // ---------------- MMA issue: thread 0 issues S; the two independent O chains are issued SIMT, one per thread (an MMA costs ~100 cycles to
      // issue, 26 of them in a row would sit on the critical path between the softmax and the epilogue) ----------------
      const int m = lane;
      const uint64_t dQm = ptx::make_smem_desc(base + F_SM_Q, 16, 1024), dQc = ptx::make_smem_desc(base + F_SM_Q + TILE_M * 128, 16, 1024);
      const uint64_t dKm = ptx::make_smem_desc(base + F_SM_K, 16, 1024), dKc = ptx::make_smem_desc(base + F_SM_K + KT_BYTES, 16, 1024);
      const uint64_t dVx = ptx::make_smem_desc(base + (m == 0 ? F_SM_V : F_SM_CV), KT_BYTES, 1024);   // MN-major: V | CV
      const uint32_t t_o = tmem_base + (m == 0 ? F_T_OM : F_T_OC), t_a = tmem_base + (m == 0 ? 0 : F_T_PSQ);
      const uint32_t idesc_s = ptx::make_idesc_bf16(TILE_M, n_pad, false, false), idesc_o = ptx::make_idesc_bf16(TILE_M, HD, false, true);
      const int ksteps = n_pad >> 4;
      for (int it = 0; it < n_items; ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        w_wait(qk_full, ph);
        if (m == 0) {
          // S = X1m X2m^T + X1u X2u^T : the K = 128 contraction as two 64-wide swizzle atoms (4 k-steps each)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16(tmem_base, dQm + 2 * ks, dKm + 2 * ks, idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16(tmem_base, dQc + 2 * ks, dKc + 2 * ks, idesc_s, 1u);
          ptx::umma_commit(s_full);               // also: X1 / X2 have been read (the TMA thread may fetch the next item's)
        }
        __syncwarp(0x3u);
        w_wait(v_full, ph);
        w_wait(p_full, ph);
        if (it > 0) w_wait(t_free, ph ^ 1u);      // the epilogue has drained O_m / O_c of the previous item
        ptx::tc_fence_after();
        // thread 0: O_m = P~ V (A = TMEM columns [0, n_pad/2)); thread 1: O_c = (P~)^2 CV (A = the squared copy). Same accumulate flags in both lanes.
        for (int kk = 0; kk < ksteps; ++kk) ptx::umma_bf16_ts(t_o, t_a + kk * 8, dVx + 128 * kk, idesc_o, kk > 0 ? 1u : 0u);
        ptx::umma_commit(o_full);
      }
    }
  } else if (warp == F_EW_WARPS + 1) {
    if (lane == 0 && n_items > 0) {
      // ---------------- TMA: X1 tile + X2 (freed by the S MMAs), V + CV (freed by the epilogue's TMA store) ----------------
      for (int it = 0; it < n_items; ++it) {
        int b, h, m0, bh;
        item_of(it, b, h, m0, bh);
        if (it > 0) w_wait(s_full, (uint32_t)((it - 1) & 1));
        ptx::mbar_arrive_expect_tx(qk_full, 2 * TILE_M * 128 + 2 * n_pad * 128);
        ptx::tma_load_3d(base + F_SM_Q, &tm_x1, qk_full, h * XW, m0, b);
        ptx::tma_load_3d(base + F_SM_Q + TILE_M * 128, &tm_x1, qk_full, h * XW + HD, m0, b);
        ptx::tma_load_3d(base + F_SM_K, &tm_x2, qk_full, (p.H + h) * XW, 0, b);
        ptx::tma_load_3d(base + F_SM_K + KT_BYTES, &tm_x2, qk_full, (p.H + h) * XW + HD, 0, b);
        if (it > 0) w_wait(v_free, (uint32_t)((it - 1) & 1));
        ptx::mbar_arrive_expect_tx(v_full, 2 * n_pad * 128);
        ptx::tma_load_3d(base + F_SM_V, &tm_v, v_full, (2 * p.H + h) * HD, 0, b);
        ptx::tma_load_3d(base + F_SM_CV, &tm_cv, v_full, (2 * p.H + h) * HD, 0, b);
      }
    } else if (lane == 1 && n_items > 0) {
      // ---------------- TMA: bias ring, [128 queries x 32 keys] fp32 boxes, chunk c is consumed by column half c & 1 ----------------
      int gc = 0;
      for (int it = 0; it < n_items; ++it) {
        int b, h, m0, bh;
        item_of(it, b, h, m0, bh);
        for (int c = 0; c < nchunks; ++c, ++gc) {
          const int s = gc & (F_BIAS_STAGES - 1);
          if (gc >= F_BIAS_STAGES) w_wait(bias_empty(s), (uint32_t)(((gc / F_BIAS_STAGES) - 1) & 1));
          ptx::mbar_arrive_expect_tx(bias_full(s), BIAS_STAGE_BYTES);
          ptx::tma_load_3d(base + F_SM_BIAS + s * BIAS_STAGE_BYTES, &tm_bias, bias_full(s), c * 32, m0, h);
        }
      }
    }
  } else {
    // ---------------- element-wise warps: two threads per query row ----------------
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const int tid = threadIdx.x;                             // 0..255
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint64_t seed = (DROP && p.seed_dev != nullptr) ? __ldg(reinterpret_cast<const unsigned long long*>(p.seed_dev)) : p.seed;
    float* s_ls = reinterpret_cast<float*>(gbase + F_SM_LS);
    for (int it = 0; it < n_items; ++it) {
      int b, h, m0, bh;
      item_of(it, b, h, m0, bh);
      const uint32_t ph = (uint32_t)(it & 1);
      const int i = m0 + row;
      const bool active = m0 + quad * 32 < p.N;            // warps whose 32 rows are all past N only keep the barrier protocol alive
      float* s_cn = reinterpret_cast<float*>(gbase + F_SM_CN) + (it & 1) * NMAX;
      if (tid < n_pad) s_cn[tid] = tid < p.N ? __ldg(p.cn + (long long)bh * p.N + tid) : 0.f;
      const float rnh = i < p.N ? __ldg(p.rn + (long long)bh * p.N + i) : 0.f;
      const float mst = i < p.N ? __ldg(p.rowmax + (long long)h * p.N + i) + LOG2E : 0.f;
      const float k0 = 0.5f * LOG2E - mst;
      ptx::named_bar_sync(1, F_EW_WARPS * 32);
      w_wait(s_full, ph);
      ptx::tc_fence_after();
      float l = 0.f;
      for (int c = half; c < nchunks; c += 2) {
        const int gc = it * nchunks + c, s = gc & (F_BIAS_STAGES - 1);
        w_wait(bias_full(s), (uint32_t)((gc / F_BIAS_STAGES) & 1));
        if (active) {
          const uint8_t* bias_row = gbase + F_SM_BIAS + s * BIAS_STAGE_BYTES + row * 128;
          const bool full = c + 1 < nchunks || tail_cols == 32;
          uint32_t w = 0xffffffffu;
          if (DROP) w = w_keep_word(p, bh, i, c);
          if (full) wfwd_chunk<DROP, 32>(trow + c * 32, trow + c * 16, trow + F_T_PSQ + c * 16, bias_row, row, s_cn + c * 32, rnh, k0, l, w);
          else wfwd_chunk<DROP, 16>(trow + c * 32, trow + c * 16, trow + F_T_PSQ + c * 16, bias_row, row, s_cn + c * 32, rnh, k0, l, w);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bias_empty(s));
      }
      s_ls[half * TILE_M + row] = l;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
      ptx::named_bar_sync(2, F_EW_WARPS * 32);
      const float ltot = s_ls[row] + s_ls[TILE_M + row];
      // ---------------- epilogue: half 0 drains O_m, half 1 drains O_c ----------------
      w_wait(o_full, ph);
      ptx::tc_fence_after();
      if (active) {
        const float inv1 = p.inv_keep / ltot;
        const float inv = half == 0 ? inv1 : inv1 * inv1;
        uint8_t* orow = gbase + (half == 0 ? F_SM_V : F_SM_CV) + row * 128;
        const uint32_t tsrc = trow + (half == 0 ? F_T_OM : F_T_OC);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o[32];
          ptx::tmem_ld_x32_sync(tsrc + hh * 32, o);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + (((hh * 4 + q) ^ (row & 7)) << 4)) = u;
          }
        }
        if (half == 0 && i < p.N && p.lse != nullptr) p.lse[(long long)bh * p.N + i] = (mst + log2f(ltot)) / LOG2E;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(t_free);
      ptx::fence_proxy_async();
      ptx::named_bar_sync(3, F_EW_WARPS * 32);
      if (warp == 0 && lane == 0) {
        ptx::tma_store_3d(&tm_om, base + F_SM_V, h * HD, m0, b);
        ptx::tma_store_3d(&tm_oc, base + F_SM_CV, h * HD, m0, b);
        ptx::bulk_commit();
        ptx::bulk_wait_read0();
        ptx::mbar_arrive(v_free);
      }
    }
    if (warp == 0 && lane == 0) ptx::bulk_wait0();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == F_EW_WARPS) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Delta[b, h, i] = sum_d dOm[i,d] Om[i,d] + 2 sum_d dOc[i,d] Oc[i,d]  (= sum_j P~_ij dP~_ij): one warp per token row, 8 lanes per head
__global__ void __launch_bounds__(256) wattn_bwd_prep_kernel(const bf16* __restrict__ om, const bf16* __restrict__ dom, const bf16* __restrict__ oc,
                                                             const bf16* __restrict__ doc, float* __restrict__ dvec, int B, int H, int N) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * N;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(r / N), i = (int)(r - (long long)b * N);
    for (int h0 = 0; h0 < H; h0 += 4) {
      const int h = h0 + (lane >> 3);
      float acc = 0.f;
      if (h < H) {
        const long long off = (r * H + h) * HD + (lane & 7) * 8;
        const uint4 a0 = *reinterpret_cast<const uint4*>(om + off), d0 = *reinterpret_cast<const uint4*>(dom + off);
        const uint4 a1 = *reinterpret_cast<const uint4*>(oc + off), d1 = *reinterpret_cast<const uint4*>(doc + off);
        const uint32_t* pa0 = &a0.x; const uint32_t* pd0 = &d0.x; const uint32_t* pa1 = &a1.x; const uint32_t* pd1 = &d1.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 x0 = unpack_bf16x2(pa0[k]), y0 = unpack_bf16x2(pd0[k]), x1 = unpack_bf16x2(pa1[k]), y1 = unpack_bf16x2(pd1[k]);
          acc += x0.x * y0.x + x0.y * y0.y + 2.0f * (x1.x * y1.x + x1.y * y1.y);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (h < H && (lane & 7) == 0) dvec[((long long)b * H + h) * N + i] = acc;
    }
  }
}

// keep_bits [BH, N(query i), 8 words over keys] -> keep_t [BH, N(key j), 8 words over queries] (32 x 32 bit-block butterflies)
__global__ void __launch_bounds__(256) w_keep_transpose_kernel(const uint32_t* __restrict__ keep_bits, uint32_t* __restrict__ keep_t, int N) {
  __shared__ uint32_t tile[NMAX * 9];
  const long long bh = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blocks = (N + 31) >> 5;
  for (int idx = threadIdx.x; idx < N * 8; idx += blockDim.x) tile[(idx >> 3) * 9 + (idx & 7)] = keep_bits[bh * N * 8 + idx];
  __syncthreads();
  if (warp < blocks) {
    const int jb = warp;
    uint32_t mine[8];
#pragma unroll
    for (int ib = 0; ib < 8; ++ib) {
      mine[ib] = 0u;
      if (ib < blocks) {
        const int i = ib * 32 + lane;
        uint32_t x = i < N ? tile[i * 9 + jb] : 0u;
#pragma unroll
        for (int step = 0; step < 5; ++step) {
          const int s_ = 16 >> step;
          const uint32_t m = step == 0 ? 0x0000ffffu : step == 1 ? 0x00ff00ffu : step == 2 ? 0x0f0f0f0fu : step == 3 ? 0x33333333u : 0x55555555u;
          const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s_);
          x = (lane & s_) ? (((y >> s_) & m) | (x & ~m)) : ((x & m) | ((y << s_) & ~m));
        }
        mine[ib] = x;
      }
    }
    const int j = jb * 32 + lane;
    if (j < N) {
      uint4* dst = reinterpret_cast<uint4*>(keep_t + (bh * N + j) * 8);
      dst[0] = make_uint4(mine[0], mine[1], mine[2], mine[3]);
      dst[1] = make_uint4(mine[4], mine[5], mine[6], mine[7]);
    }
  }
}

__device__ __forceinline__ void w_st_global_32B(void* ptr, const uint32_t (&a)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]),
               "r"(a[6]), "r"(a[7])
               : "memory");
}

// column sums over the 32 lanes of a warp of 32 per-lane values (recursive halving, 31 shuffles); lane l ends with column l
__device__ __forceinline__ float w_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const bool up = lane & 1;
  const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// ---- key-tile kernel ----
constexpr int BOXQ = 64;                                 // queries per box
constexpr int KV_SM_X2 = 0;                              // X2 key tile: mean half [128 x 128 B], cov half
constexpr int KV_SM_V = KV_SM_X2 + 2 * TILE_M * 128;     // V key tile
constexpr int KV_SM_CV = KV_SM_V + TILE_M * 128;         // CV key tile
constexpr int KV_SLOTS = 3;
constexpr int KV_SLOT_BYTES = 4 * BOXQ * 128;            // X1 box mean half | X1 box cov half | dOm box | dOc box  (each [64 x 128 B])
constexpr int KV_SM_BOX = KV_SM_CV + TILE_M * 128;
// bias^T ring: each element-wise group owns TWO private stages of [128 keys x 16 queries] fp32 (SWIZZLE_64B, one stage per 16-query step). A
// ring shared by both groups would hand a stage back and forth between them, and a waiter that does not observe EVERY phase of an mbarrier
// cannot tell a fresh fill from a stale one by parity.
constexpr int KV_BIAS_CHUNK_BYTES = TILE_M * 64;
constexpr int KV_SM_BIAS = KV_SM_BOX + KV_SLOTS * KV_SLOT_BYTES;
constexpr int KV_SM_STATS = KV_SM_BIAS + 4 * KV_BIAS_CHUNK_BYTES;     // 2 buffers x {lse2, Delta, rn} x 256 floats
constexpr int KV_SM_BAR = KV_SM_STATS + 2 * 3 * 1024;
constexpr int KV_SMEM = KV_SM_BAR + 1024 + 1024;
constexpr int KV_EW_WARPS = 8;
constexpr int KV_THREADS = (KV_EW_WARPS + 2) * 32;
constexpr int KV_T_GROUP = 192, KV_T_ST = 0, KV_T_GM = 64, KV_T_GC = 128, KV_T_DV = 384, KV_T_DCV = 448;
static_assert(KV_SM_BOX % 1024 == 0 && KV_SM_BIAS % 1024 == 0 && KV_SMEM <= 232448, "kv kernel smem layout");

struct WKvParams {
  const float* lse;          // [B, H, N]
  const float* dvec;         // [B, H, N]
  const float* rn;           // [B, H, N] half squared norms
  const float* cn;
  const uint32_t* keep_t;    // [B, H, N(key), 8]
  const bf16* qkv_c;         // raw elu(.)+1 values (chain rule of dCV)
  bf16* dd_out;              // [B, H, N(key), ld(query)]  dD^T
  bf16* da_out;              // same layout, dA^T (bias-table gradient) or null
  int ld;
  bf16* dqkv_m;              // [B, N, 3, H, 64]
  bf16* dqkv_c;
  float* dv_bias;            // [H*64] += or null
  float* dcv_bias;
  int B, H, N, n_pad, k_tiles, items;
  float inv_keep;
};

// 16 queries of one key row out of TMEM: P, P~, dP~, dA, dD; P~^T / (P~^2)^T written back as bf16 pairs behind the read pointer
// three x16 loads in flight, one wait, ONE asm block (the destination registers are defined only after the wait)
__device__ __forceinline__ void w_tmem_ld_x16_triple_sync(uint32_t ta, uint32_t (&a)[16], uint32_t tb, uint32_t (&b)[16], uint32_t tc, uint32_t (&c)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%48];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%49];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47}, [%50];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),
        "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]),
        "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]), "=r"(c[9]), "=r"(c[10]), "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15])
      : "r"(ta), "r"(tb), "r"(tc)
      : "memory");
}

template <bool DROP>
__device__ __forceinline__ void wkv_step16(const WKvParams& p, uint32_t t_st, uint32_t t_gm, uint32_t t_gc, uint32_t t_p, uint32_t t_psq,
                                           const uint8_t* bias_row, int row, const float* s_lse, const float* s_dl, const float* s_rn, float cnj,
                                           uint32_t kw, bool valid, bf16* dd_row, bf16* da_row) {
  uint32_t x[16], gm[16], gc[16];
  w_tmem_ld_x16_triple_sync(t_st, x, t_gm, gm, t_gc, gc);
  uint32_t pk[8], pq[8], dd[8], da[8];
  if (valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_row + ((q ^ ((row >> 1) & 3)) << 4));      // SWIZZLE_64B: 16-byte chunk ^ address bits 7..8
      const float4 l4 = *reinterpret_cast<const float4*>(s_lse + 4 * q);
      const float4 d4 = *reinterpret_cast<const float4*>(s_dl + 4 * q);
      const float4 r4 = *reinterpret_cast<const float4*>(s_rn + 4 * q);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w};
      float pt[4], pt2[4], dDv[4], dAv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = 4 * q + e;
        const float th = tanh_approx((rr[e] + cnj) - __uint_as_float(x[idx]));     // tanh(D / 2); t = sigmoid(-D) = 0.5 - 0.5 th
        const float pv = ex2(fmaf(th, -0.5f * LOG2E, bb[e] + (0.5f * LOG2E - ll[e])));   // lse2 = +inf past N -> 0
        const float f = (!DROP || (kw & (1u << idx))) ? p.inv_keep : 0.f;
        pt[e] = f * pv;
        const float dpt = fmaf(2.0f * pt[e], __uint_as_float(gc[idx]), __uint_as_float(gm[idx]));
        dAv[e] = pv * fmaf(f, dpt, -dl[e]);
        dDv[e] = dAv[e] * fmaf(th * 0.25f, th, -0.25f);                          // -dA t (1 - t), t (1 - t) = (1 - th^2) / 4
        pt2[e] = pt[e] * pt[e];
      }
      pk[2 * q] = pack_bf16x2(pt[0], pt[1]); pk[2 * q + 1] = pack_bf16x2(pt[2], pt[3]);
      pq[2 * q] = pack_bf16x2(pt2[0], pt2[1]); pq[2 * q + 1] = pack_bf16x2(pt2[2], pt2[3]);
      dd[2 * q] = pack_bf16x2(dDv[0], dDv[1]); dd[2 * q + 1] = pack_bf16x2(dDv[2], dDv[3]);
      da[2 * q] = pack_bf16x2(dAv[0], dAv[1]); da[2 * q + 1] = pack_bf16x2(dAv[2], dAv[3]);
    }
    w_st_global_32B(dd_row, dd);
    if (da_row != nullptr) w_st_global_32B(da_row, da);
  } else {      // key rows past N: P = 0 (their X2 / V rows are TMA zero fill, not masked scores)
#pragma unroll
    for (int q = 0; q < 8; ++q) pk[q] = pq[q] = 0u;
  }
  ptx::tmem_st_x8(t_p, pk);
  ptx::tmem_st_x8(t_psq, pq);
}

template <bool DROP>
__global__ void __launch_bounds__(KV_THREADS, 1)
wattn_bwd_kv_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2, const __grid_constant__ CUtensorMap tm_v,
                    const __grid_constant__ CUtensorMap tm_cv, const __grid_constant__ CUtensorMap tm_dom, const __grid_constant__ CUtensorMap tm_doc,
                    const __grid_constant__ CUtensorMap tm_bias, const WKvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;       // 0..7 element-wise, 8 MMA issue (two threads), 9 TMA + bias ring
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = base + KV_SM_BAR;
  const uint32_t kv_full = bar0, kv_free = bar0 + 8, acc_full = bar0 + 16, acc_empty = bar0 + 24;
  auto s_full = [&](int g) { return bar0 + 32u + 8u * g; };
  auto p_full = [&](int g) { return bar0 + 48u + 8u * g; };
  auto bias_full = [&](int s) { return bar0 + 64u + 8u * s; };      // 4 stages: group g owns 2g, 2g + 1
  auto bias_empty = [&](int s) { return bar0 + 160u + 8u * s; };
  auto ld_full = [&](int s) { return bar0 + 96u + 8u * s; };
  auto ld_empty = [&](int s) { return bar0 + 128u + 8u * s; };
  const uint32_t tmem_slot = bar0 + 192u;

  const int n_pad = p.n_pad;
  const int nboxes = (n_pad + BOXQ - 1) / BOXQ;     // group g owns boxes g, g + 2, ...
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_items = first < p.items ? (p.items - first + stride - 1) / stride : 0;
  auto item_of = [&](int it, int& b, int& h, int& j0, int& bh) {
    const int item = first + it * stride;
    const int kt = item % p.k_tiles;
    bh = item / p.k_tiles;
    b = bh / p.H; h = bh - b * p.H; j0 = kt * TILE_M;
  };
  auto group_count = [&](int it, int bi) { const int g = bi & 1; return it * ((nboxes + 1 - g) >> 1) + (bi >> 1); };

  if (warp == KV_EW_WARPS) {
    if (lane == 0) {
      ptx::mbar_init(kv_full, 1); ptx::mbar_init(kv_free, 2); ptx::mbar_init(acc_full, 2); ptx::mbar_init(acc_empty, KV_EW_WARPS);
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(s_full(s), 2); ptx::mbar_init(p_full(s), 128); }
      for (int s = 0; s < 4; ++s) { ptx::mbar_init(bias_full(s), 1); ptx::mbar_init(bias_empty(s), 4); }
      for (int s = 0; s < KV_SLOTS; ++s) { ptx::mbar_init(ld_full(s), 1); ptx::mbar_init(ld_empty(s), 2); }
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tm_x1); ptx::prefetch_tmap(&tm_x2); ptx::prefetch_tmap(&tm_v); ptx::prefetch_tmap(&tm_cv);
      ptx::prefetch_tmap(&tm_dom); ptx::prefetch_tmap(&tm_doc); ptx::prefetch_tmap(&tm_bias);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == KV_EW_WARPS) {
    if (lane < 2 && n_items > 0) {
      // ---------------- MMA issue, SIMT from two threads: m = 0 owns S^T (8 k-steps) and dV, m = 1 owns Gm^T / Gc^T (4 + 4) and dCV ----------------
      const int m = lane;
      const uint64_t dA0 = ptx::make_smem_desc(base + (m == 0 ? KV_SM_X2 : KV_SM_V), 16, 1024);                    // X2 mean half | V
      const uint64_t dA1 = ptx::make_smem_desc(base + (m == 0 ? KV_SM_X2 + TILE_M * 128 : KV_SM_CV), 16, 1024);    // X2 cov half  | CV
      const uint32_t t_acc = tmem_base + (m == 0 ? KV_T_DV : KV_T_DCV);
      const uint32_t idesc_acc = ptx::make_idesc_bf16(TILE_M, HD, false, true);
      int cnt[2] = {0, 0};
      int gb_s = 0;
      auto issue_scores = [&](int bi) {
        const int sl = gb_s % KV_SLOTS, g = bi & 1;
        w_wait(ld_full(sl), (uint32_t)((gb_s / KV_SLOTS) & 1));
        const int wb = min(BOXQ, n_pad - bi * BOXQ);
        const uint32_t idesc = ptx::make_idesc_bf16(TILE_M, wb, false, false);
        const uint32_t slot = base + KV_SM_BOX + sl * KV_SLOT_BYTES;
        // m = 0: B = X1 box mean half, then cov half (same accumulator); m = 1: B = dOm box -> Gm, dOc box -> Gc
        const uint64_t dB0 = ptx::make_smem_desc(slot + (m == 0 ? 0 : 2 * BOXQ * 128), 16, 1024);
        const uint64_t dB1 = ptx::make_smem_desc(slot + (m == 0 ? BOXQ * 128 : 3 * BOXQ * 128), 16, 1024);
        const uint32_t d0 = tmem_base + g * KV_T_GROUP + (m == 0 ? KV_T_ST : KV_T_GM);
        const uint32_t d1 = tmem_base + g * KV_T_GROUP + (m == 0 ? KV_T_ST : KV_T_GC);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_bf16(d0, dA0 + 2 * ks, dB0 + 2 * ks, idesc, ks > 0 ? 1u : 0u);
        // The accumulate flag of a tcgen05.mma issued SIMT by two lanes must be the SAME in both lanes (measured: with 1 in lane 0 and 0 in
        // lane 1 the second lane's MMA accumulated as well — the enable-input-d predicate is evaluated warp-uniformly). Here lane 0 continues
        // S^T (accumulate) while lane 1 starts Gc^T (overwrite): issue that one k-step from divergent branches, one active lane each.
        if (m == 0) ptx::umma_bf16(d1, dA1, dB1, idesc, 1u);
        else ptx::umma_bf16(d1, dA1, dB1, idesc, 0u);
#pragma unroll
        for (int ks = 1; ks < 4; ++ks) ptx::umma_bf16(d1, dA1 + 2 * ks, dB1 + 2 * ks, idesc, 1u);
        ptx::umma_commit(s_full(g));
        if (bi == nboxes - 1) ptx::umma_commit(kv_free);
        ++gb_s;
      };
      int gb = 0;
      for (int it = 0; it < n_items; ++it) {
        w_wait(kv_full, (uint32_t)(it & 1));
        issue_scores(0);
        if (nboxes > 1) issue_scores(1);
        for (int bi = 0; bi < nboxes; ++bi, ++gb) {
          const int sl = gb % KV_SLOTS, g = bi & 1;
          w_wait(p_full(g), (uint32_t)(cnt[g] & 1));
          ++cnt[g];
          if (bi == 0 && it > 0) w_wait(acc_empty, (uint32_t)((it - 1) & 1));
          ptx::tc_fence_after();
          const int wb = min(BOXQ, n_pad - bi * BOXQ);
          const uint32_t slot = base + KV_SM_BOX + sl * KV_SLOT_BYTES;
          const uint64_t db = ptx::make_smem_desc(slot + (m == 0 ? 2 * BOXQ * 128 : 3 * BOXQ * 128), BOXQ * 128, 1024);   // dOm | dOc box, MN-major
          const uint32_t a = tmem_base + g * KV_T_GROUP + (m == 0 ? KV_T_ST : KV_T_GM);     // P~^T | (P~^2)^T written in place
          for (int kk = 0; kk < (wb >> 4); ++kk) ptx::umma_bf16_ts(t_acc, a + kk * 8, db + 128 * kk, idesc_acc, (bi > 0 || kk > 0) ? 1u : 0u);
          ptx::umma_commit(ld_empty(sl));
          if (bi == nboxes - 1) ptx::umma_commit(acc_full);
          if (bi + 2 < nboxes) issue_scores(bi + 2);
        }
      }
    }
  } else if (warp == KV_EW_WARPS + 1) {
    if (lane == 0 && n_items > 0) {
      // ---------------- TMA: key tiles (one item ahead) and the box ring (runs on across items) ----------------
      int b, h, j0, bh;
      item_of(0, b, h, j0, bh);
      auto load_keys = [&]() {
        ptx::mbar_arrive_expect_tx(kv_full, 4 * TILE_M * 128);
        ptx::tma_load_3d(base + KV_SM_X2, &tm_x2, kv_full, (p.H + h) * XW, j0, b);
        ptx::tma_load_3d(base + KV_SM_X2 + TILE_M * 128, &tm_x2, kv_full, (p.H + h) * XW + HD, j0, b);
        ptx::tma_load_3d(base + KV_SM_V, &tm_v, kv_full, (2 * p.H + h) * HD, j0, b);
        ptx::tma_load_3d(base + KV_SM_CV, &tm_cv, kv_full, (2 * p.H + h) * HD, j0, b);
      };
      load_keys();
      int gb = 0;
      for (int it = 0; it < n_items; ++it) {
        for (int bi = 0; bi < nboxes; ++bi, ++gb) {
          const int sl = gb % KV_SLOTS;
          if (gb >= KV_SLOTS) w_wait(ld_empty(sl), (uint32_t)(((gb / KV_SLOTS) - 1) & 1));
          const uint32_t slot = base + KV_SM_BOX + sl * KV_SLOT_BYTES;
          ptx::mbar_arrive_expect_tx(ld_full(sl), KV_SLOT_BYTES);
          ptx::tma_load_3d(slot, &tm_x1, ld_full(sl), h * XW, bi * BOXQ, b);
          ptx::tma_load_3d(slot + BOXQ * 128, &tm_x1, ld_full(sl), h * XW + HD, bi * BOXQ, b);
          ptx::tma_load_3d(slot + 2 * BOXQ * 128, &tm_dom, ld_full(sl), h * HD, bi * BOXQ, b);
          ptx::tma_load_3d(slot + 3 * BOXQ * 128, &tm_doc, ld_full(sl), h * HD, bi * BOXQ, b);
        }
        if (it + 1 < n_items) {
          item_of(it + 1, b, h, j0, bh);
          w_wait(kv_free, (uint32_t)(it & 1));
          load_keys();
        }
      }
    } else if ((lane == 1 || lane == 2) && n_items > 0) {
      // ---------------- TMA: bias^T, one producer thread per element-wise group, [128 keys x 16 queries] fp32 chunks in that group's step order ----------------
      const int g = lane - 1;
      int cb = 0;
      for (int it = 0; it < n_items; ++it) {
        int b, h, j0, bh;
        item_of(it, b, h, j0, bh);
        for (int bi = g; bi < nboxes; bi += 2) {
          const int wb = min(BOXQ, n_pad - bi * BOXQ);
          for (int s16 = 0; s16 * 16 < wb; ++s16, ++cb) {
            const int st = 2 * g + (cb & 1);
            if (cb >= 2) w_wait(bias_empty(st), (uint32_t)(((cb >> 1) - 1) & 1));
            ptx::mbar_arrive_expect_tx(bias_full(st), KV_BIAS_CHUNK_BYTES);
            ptx::tma_load_3d(base + KV_SM_BIAS + st * KV_BIAS_CHUNK_BYTES, &tm_bias, bias_full(st), bi * BOXQ + s16 * 16, j0, h);
          }
        }
      }
    }
  } else {
    // ---------------- element-wise warps: one thread per key row; group g = warp >> 2 owns boxes g, g + 2, ... ----------------
    const int quad = warp & 3, g = warp >> 2;
    const int row = quad * 32 + lane;
    const int tid = threadIdx.x;                      // 0..255
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tg = trow + g * KV_T_GROUP;
    float nx_lse = INFINITY, nx_d = 0.f, nx_rn = 0.f, nx_cn = 0.f;
    auto fetch_stats = [&](int it) {
      int b, h, j0, bh;
      item_of(it, b, h, j0, bh);
      const bool in = tid < p.N;
      nx_lse = in ? __ldg(p.lse + (long long)bh * p.N + tid) * LOG2E : INFINITY;
      nx_d = in ? __ldg(p.dvec + (long long)bh * p.N + tid) : 0.f;
      nx_rn = in ? __ldg(p.rn + (long long)bh * p.N + tid) : 0.f;
      nx_cn = (j0 + row) < p.N ? __ldg(p.cn + (long long)bh * p.N + j0 + row) : 0.f;
    };
    if (n_items > 0) fetch_stats(0);
    int cb = 0;                                          // running 16-query step counter of this group (bias stage / phase)
    for (int it = 0; it < n_items; ++it) {
      int b, h, j0, bh;
      item_of(it, b, h, j0, bh);
      float* s_lse = reinterpret_cast<float*>(gbase + KV_SM_STATS) + (it & 1) * 768;
      float* s_dl = s_lse + 256;
      float* s_rn = s_lse + 512;
      s_lse[tid] = nx_lse; s_dl[tid] = nx_d; s_rn[tid] = nx_rn;
      const float cnj = nx_cn;
      ptx::named_bar_sync(1, KV_EW_WARPS * 32);
      if (it + 1 < n_items) fetch_stats(it + 1);
      const int j = j0 + row;
      const bool valid = j < p.N;
      const bool active = j0 + quad * 32 < p.N;
      const long long rowoff = ((long long)bh * p.N + (valid ? j : 0)) * p.ld;
      bf16* dd_base = p.dd_out + rowoff;
      bf16* da_base = p.da_out != nullptr ? p.da_out + rowoff : nullptr;
      const uint32_t* kt_row = p.keep_t + ((long long)bh * p.N + (valid ? j : 0)) * 8;
      for (int bi = g; bi < nboxes; bi += 2) {
        const int c0 = bi * BOXQ;
        const int wb = min(BOXQ, n_pad - c0);
        w_wait(s_full(g), (uint32_t)(group_count(it, bi) & 1));
        ptx::tc_fence_after();
        uint32_t kw = 0xffffffffu;
        for (int s16 = 0; s16 * 16 < wb; ++s16, ++cb) {
          const int st = 2 * g + (cb & 1);
          const int o = s16 * 16;                        // column offset inside the box
          const int q0c = c0 + o;                        // first query of this 16-step
          if (DROP && (s16 & 1) == 0) kw = valid ? __ldg(kt_row + (q0c >> 5)) : 0xffffffffu;
          w_wait(bias_full(st), (uint32_t)((cb >> 1) & 1));
          if (active) {
            const uint8_t* bias_row = gbase + KV_SM_BIAS + st * KV_BIAS_CHUNK_BYTES + row * 64;
            wkv_step16<DROP>(p, tg + KV_T_ST + o, tg + KV_T_GM + o, tg + KV_T_GC + o, tg + KV_T_ST + (o >> 1), tg + KV_T_GM + (o >> 1), bias_row, row,
                             s_lse + q0c, s_dl + q0c, s_rn + q0c, cnj, (s16 & 1) ? (kw >> 16) : kw, valid, dd_base + q0c,
                             da_base != nullptr ? da_base + q0c : nullptr);
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bias_empty(st));
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(p_full(g));
      }
      // ---------------- epilogue: dV rows (group 0) / dCV rows (group 1, through elu') of this key tile ----------------
      w_wait(acc_full, (uint32_t)(it & 1));
      ptx::tc_fence_after();
      if (active) {
        bf16* dst = (g == 0 ? p.dqkv_m : p.dqkv_c) + ((long long)b * p.N + (valid ? j : 0)) * (3LL * p.H * HD) + 2LL * p.H * HD + h * HD;
        const bf16* raw = p.qkv_c + ((long long)b * p.N + (valid ? j : 0)) * (3LL * p.H * HD) + 2LL * p.H * HD + h * HD;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o[32];
          ptx::tmem_ld_x32_sync(trow + (g == 0 ? KV_T_DV : KV_T_DCV) + hh * 32, o);
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = valid ? __uint_as_float(o[e]) : 0.f;
          if (g == 1 && valid) {     // d(pre-activation) = dcv' * elu'(z), elu'(z) = cv' for cv' <= 1 (z <= 0), else 1
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 rw = *reinterpret_cast<const uint4*>(raw + hh * 32 + 8 * q);
              const uint32_t* pw = &rw.x;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 c2 = unpack_bf16x2(pw[k]);
                v[8 * q + 2 * k] *= fminf(c2.x, 1.0f);
                v[8 * q + 2 * k + 1] *= fminf(c2.y, 1.0f);
              }
            }
          }
          if (valid) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t u[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) u[k] = pack_bf16x2(v[16 * q + 2 * k], v[16 * q + 2 * k + 1]);
              w_st_global_32B(dst + hh * 32 + 16 * q, u);
            }
          }
          float* bias_grad = g == 0 ? p.dv_bias : p.dcv_bias;
          if (bias_grad != nullptr) {
            const float c = w_colsum32(v, lane);
            atomicAdd(bias_grad + h * HD + hh * 32 + lane, c);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == KV_EW_WARPS) ptx::tmem_dealloc(tmem_base, 512);
}

// ---- per-(batch, head) kernel: dX1 = dD X2 (query side), dX2 = dD^T X1 (key side), chain rules, dq / dk / dcq / dck ----
constexpr int DX_SM_A = 0;                               // dD^T: 4 chunks [NMAX key rows x 64 query cols]
constexpr int DX_SM_X1 = DX_SM_A + 4 * KT_BYTES;         // X1[bh]: mean half, cov half ([NMAX x 128 B] each)
constexpr int DX_SM_X2 = DX_SM_X1 + 2 * KT_BYTES;
constexpr int DX_SM_BAR = DX_SM_X2 + 2 * KT_BYTES;
constexpr int DX_SMEM = DX_SM_BAR + 1024 + 1024;
constexpr int DX_EW_WARPS = 8;
constexpr int DX_THREADS = (DX_EW_WARPS + 2) * 32;
static_assert(DX_SMEM <= 232448, "dx kernel smem layout");

struct WDxParams {
  const bf16* X;            // [B, N, 2, H, 128]
  const bf16* qkv_c;        // raw elu(.)+1 values
  bf16* dqkv_m;
  bf16* dqkv_c;
  float* dq_bias;           // [H*64] += or null
  float* dcq_bias;
  int B, H, N, n_pad, items;
  float scale;
};

__global__ void __launch_bounds__(DX_THREADS, 1)
wattn_bwd_dx_kernel(const __grid_constant__ CUtensorMap tm_dd, const __grid_constant__ CUtensorMap tm_x, const WDxParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;      // 0..7 epilogue, 8 MMA, 9 TMA
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = base + DX_SM_BAR;
  const uint32_t full = bar0, mma_done = bar0 + 8, sums_done = bar0 + 16, tmem_free = bar0 + 24;
  const uint32_t tmem_slot = bar0 + 64u;
  const int n_pad = p.n_pad;
  const int ksteps = n_pad >> 4;
  const int n_items = (int)blockIdx.x < p.items ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int tiles = (p.N + TILE_M - 1) / TILE_M;           // 1 or 2 query / key tiles
  if (warp == DX_EW_WARPS) {
    if (lane == 0) {
      ptx::mbar_init(full, 1); ptx::mbar_init(mma_done, 1); ptx::mbar_init(sums_done, DX_EW_WARPS); ptx::mbar_init(tmem_free, DX_EW_WARPS);
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tm_dd); ptx::prefetch_tmap(&tm_x);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == DX_EW_WARPS + 1) {
    if (lane == 0) {       // ---------------- TMA ----------------
      for (int it = 0; it < n_items; ++it) {
        const int bh = blockIdx.x + it * gridDim.x;
        const int b = bh / p.H, h = bh - b * p.H;
        if (it > 0) {
          w_wait(mma_done, (uint32_t)((it - 1) & 1));
          w_wait(sums_done, (uint32_t)((it - 1) & 1));
        }
        ptx::mbar_arrive_expect_tx(full, 8 * n_pad * 128);
        for (int c = 0; c < 4; ++c) ptx::tma_load_3d(base + DX_SM_A + c * KT_BYTES, &tm_dd, full, c * 64, 0, bh);
        ptx::tma_load_3d(base + DX_SM_X1, &tm_x, full, h * XW, 0, b);
        ptx::tma_load_3d(base + DX_SM_X1 + KT_BYTES, &tm_x, full, h * XW + HD, 0, b);
        ptx::tma_load_3d(base + DX_SM_X2, &tm_x, full, (p.H + h) * XW, 0, b);
        ptx::tma_load_3d(base + DX_SM_X2 + KT_BYTES, &tm_x, full, (p.H + h) * XW + HD, 0, b);
      }
    }
  } else if (warp == DX_EW_WARPS) {
    if (lane == 0) {       // ---------------- MMA ----------------
      const uint32_t idesc_q = ptx::make_idesc_bf16(TILE_M, XW, true, true), idesc_k = ptx::make_idesc_bf16(TILE_M, XW, false, true);
      const uint64_t dbx1 = ptx::make_smem_desc(base + DX_SM_X1, KT_BYTES, 1024), dbx2 = ptx::make_smem_desc(base + DX_SM_X2, KT_BYTES, 1024);
      for (int it = 0; it < n_items; ++it) {
        w_wait(full, (uint32_t)(it & 1));
        if (it > 0) w_wait(tmem_free, (uint32_t)((it - 1) & 1));
        ptx::tc_fence_after();
        for (int t = 0; t < tiles; ++t) {
          // query side: acc[i, :] = sum_j dD^T[j, i] X2[j, :]   (A = dD^T as an MN-major operand: chunks 2t, 2t+1 are the two 64-wide M halves)
          const uint64_t daq = ptx::make_smem_desc(base + DX_SM_A + 2 * t * KT_BYTES, KT_BYTES, 1024);
          for (int kk = 0; kk < ksteps; ++kk) ptx::umma_bf16(tmem_base + t * XW, daq + 128 * kk, dbx2 + 128 * kk, idesc_q, kk > 0 ? 1u : 0u);
          // key side: acc[j, :] = sum_i dD^T[j, i] X1[i, :]     (the same bytes as a K-major operand: rows 128t.. of chunk kk / 4)
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t dak = ptx::make_smem_desc(base + DX_SM_A + (kk >> 2) * KT_BYTES + t * TILE_M * 128, 16, 1024) + 2 * (kk & 3);
            ptx::umma_bf16(tmem_base + (2 + t) * XW, dak, dbx1 + 128 * kk, idesc_k, kk > 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(mma_done);
      }
    }
  } else {
    // ---------------- epilogue warps: warps 0..3 the query side, 4..7 the key side; one thread per row ----------------
    const int quad = warp & 3, side = warp >> 2;
    const int r = quad * 32 + lane;
    for (int it = 0; it < n_items; ++it) {
      const int bh = blockIdx.x + it * gridDim.x;
      const int b = bh / p.H, h = bh - b * p.H;
      w_wait(full, (uint32_t)(it & 1));
      // rho_i = sum_j dD_ij / kappa_j = sum_i dD_ij from the SAME bf16 values the tensor cores multiply
      float sums[2] = {0.f, 0.f};
      for (int t = 0; t < tiles; ++t) {
        const int n = t * TILE_M + r;
        if (n >= n_pad) continue;
        float acc = 0.f;
        if (side == 0) {
          const uint8_t* chunk = gbase + DX_SM_A + (n >> 6) * KT_BYTES;
          const int col = n & 63;
          for (int j = 0; j < n_pad; ++j) {
            const uint16_t raw = *reinterpret_cast<const uint16_t*>(chunk + j * 128 + ((((col >> 3) ^ (j & 7))) << 4) + (col & 7) * 2);
            acc += __uint_as_float((uint32_t)raw << 16);
          }
        } else {
          for (int c = 0; c < 4; ++c) {
            const uint8_t* rowp = gbase + DX_SM_A + c * KT_BYTES + n * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              // the swizzle only permutes the 16-byte chunks of a row and a sum does not care about the order: start each lane at a
              // different chunk, or the 32 rows of a warp (128 B apart) would all hit the same four banks
              const uint4 v4 = *reinterpret_cast<const uint4*>(rowp + (((q + lane) & 7) << 4));
              const uint32_t* pv = &v4.x;
#pragma unroll
              for (int k = 0; k < 4; ++k) { const float2 f2 = unpack_bf16x2(pv[k]); acc += f2.x + f2.y; }
            }
          }
        }
        sums[t] = acc;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(sums_done);
      w_wait(mma_done, (uint32_t)(it & 1));
      ptx::tc_fence_after();
      for (int t = 0; t < tiles; ++t) {
        const int n = t * TILE_M + r;
        const bool valid = n < p.N;
        const bool active = t * TILE_M + quad * 32 < p.N;
        if (!active) continue;
        const long long tok = (long long)b * p.N + (valid ? n : 0);
        const bf16* xrow = p.X + ((tok * 2 + side) * p.H + h) * XW;
        const bf16* raw = p.qkv_c + tok * (3LL * p.H * HD) + (long long)side * p.H * HD + h * HD;
        bf16* dm_dst = p.dqkv_m + tok * (3LL * p.H * HD) + (long long)side * p.H * HD + h * HD;
        bf16* dc_dst = p.dqkv_c + tok * (3LL * p.H * HD) + (long long)side * p.H * HD + h * HD;
        const float rho = sums[t];
        const float mul = side == 0 ? p.scale : 1.0f;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {       // 32-column chunks: 0, 1 = mean half, 2, 3 = cov half
          uint32_t o[32];
          ptx::tmem_ld_x32_sync(tmem_base + ((uint32_t)(quad * 32) << 16) + (side * 2 + t) * XW + cc * 32, o);
          float v[32];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 xv = *reinterpret_cast<const uint4*>(xrow + cc * 32 + 8 * q);
            const uint32_t* px = &xv.x;
            uint4 rw = make_uint4(0u, 0u, 0u, 0u);
            if (cc >= 2) rw = *reinterpret_cast<const uint4*>(raw + (cc - 2) * 32 + 8 * q);
            const uint32_t* pw = &rw.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 x2 = unpack_bf16x2(px[k]);
              const float2 c2 = unpack_bf16x2(pw[k]);
              const float xs[2] = {x2.x, x2.y}, cs[2] = {c2.x, c2.y};
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int idx = 8 * q + 2 * k + e;
                const float dx = 2.0f * (rho * xs[e] - __uint_as_float(o[idx]));
                float outv;
                if (cc < 2) outv = mul * dx * xs[e] * (1.0f - xs[e]);                                  // d sigmoid
                else outv = dx * xs[e] * (1.0f - xs[e] * xs[e]) * 0.5f * fminf(cs[e], 1.0f);           // d sqrt(sigmoid) and elu'
                v[idx] = valid ? outv : 0.f;
              }
            }
          }
          if (valid) {
            bf16* dst = (cc < 2 ? dm_dst : dc_dst) + (cc & 1) * 32;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t u[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) u[k] = pack_bf16x2(v[16 * q + 2 * k], v[16 * q + 2 * k + 1]);
              w_st_global_32B(dst + 16 * q, u);
            }
          }
          float* bias_grad = cc < 2 ? p.dq_bias : p.dcq_bias;
          if (side == 0 && bias_grad != nullptr) {
            const float c = w_colsum32(v, lane);
            atomicAdd(bias_grad + h * HD + (cc & 1) * 32 + lane, c);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_free);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == DX_EW_WARPS) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled w_encode_fn() {
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
int w_tmap3(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
            uint32_t b0, uint32_t b1, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = w_encode_fn();
  if (enc == nullptr) {
    b200vit_set_error("wattn: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -2;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * esize, s2 * esize};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200vit_set_error("wattn: cuTensorMapEncodeTiled failed (%d) ptr=%p dims=%llu,%llu,%llu box=%u,%u", (int)r, ptr, (unsigned long long)d0,
                      (unsigned long long)d1, (unsigned long long)d2, b0, b1);
    return -3;
  }
  return 0;
}

size_t w_align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct XWork {
  bf16* X;
  float* rn;
  float* cn;
};
XWork split_xwork(void* xwork, int B, int H, int N) {
  XWork w;
  uint8_t* p = static_cast<uint8_t*>(xwork);
  w.X = reinterpret_cast<bf16*>(p);
  p += w_align256((size_t)B * N * 2 * H * XW * sizeof(bf16));
  w.rn = reinterpret_cast<float*>(p);
  p += w_align256((size_t)B * H * N * sizeof(float));
  w.cn = reinterpret_cast<float*>(p);
  return w;
}

}  // namespace

extern "C" size_t b200vit_wattn_workspace_bytes(int32_t B, int32_t H, int32_t N) {
  if (B <= 0 || H <= 0 || N <= 0) return 0;
  return w_align256((size_t)B * N * 2 * H * XW * sizeof(bf16)) + 2 * w_align256((size_t)B * H * N * sizeof(float));
}

extern "C" int b200vit_wattn_fwd(const void* qkv_mean, const void* qkv_cov, const float* bias, int64_t ld_bias, const float* bias_rowmax, void* xwork,
                                 int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop, uint64_t seed,
                                 const uint64_t* seed_dev, uint32_t stream_id, const uint8_t* keep_in, void* out_mean, void* out_cov, float* lse,
                                 uint8_t* keep_bits, int32_t keep_ready, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv_mean && qkv_cov && out_mean && out_cov && xwork, "wattn_fwd: null pointer (xwork of b200vit_wattn_workspace_bytes is required)");
  B200_CHECK_ARG(B > 0 && H > 0, "wattn_fwd: bad B=%d H=%d", B, H);
  B200_CHECK_ARG(head_dim == HD, "wattn_fwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "wattn_fwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || keep_bits != nullptr), "wattn_fwd: bad p_drop / missing keep_bits");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(bias != nullptr && bias_rowmax != nullptr && ld_bias >= n_pad && ld_bias % 4 == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                 "wattn_fwd: the dual-stream attention needs the shared relative position bias (padded layout of b200vit_rel_pos_bias, ld %% 4 == 0, "
                 ">= %d, with its row maxima); the reference fails without it (modeling_finetune_dist.py:155)", n_pad);
  const uintptr_t addrs[] = {(uintptr_t)qkv_mean, (uintptr_t)qkv_cov, (uintptr_t)out_mean, (uintptr_t)out_cov, (uintptr_t)xwork};
  for (uintptr_t a : addrs) B200_CHECK_ARG((a & 15) == 0, "wattn_fwd: tensors must be 16-byte aligned");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(xwork) & 255) == 0, "wattn_fwd: xwork must be 256-byte aligned");
  B200_CHECK_ARG((long long)B * N * 2 * H < (1LL << 31), "wattn_fwd: B * N * H too large");
  const XWork xw = split_xwork(xwork, B, H, N);
  const int sms = b200vit_num_sms();
  wattn_prep_kernel<<<sms * 8, 256, 0, stream>>>(static_cast<const bf16*>(qkv_mean), static_cast<const bf16*>(qkv_cov), xw.X, xw.rn, xw.cn, B, H, N, scale);
  B200_CHECK_LAUNCH("wattn_prep");

  WFwdParams p;
  p.rn = xw.rn; p.cn = xw.cn; p.rowmax = bias_rowmax; p.lse = lse; p.keep_bits = keep_bits; p.keep_in = keep_in;
  p.B = B; p.H = H; p.N = N; p.n_pad = n_pad; p.m_tiles = (N + TILE_M - 1) / TILE_M; p.items = B * H * p.m_tiles;
  p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  p.thresh = (uint32_t)(p_drop * 65536.0f + 0.5f); p.seed = seed; p.seed_dev = seed_dev; p.stream_id = stream_id;
  const uint64_t xrow = 2ull * H * XW, qrow = 3ull * H * HD, orow = (uint64_t)H * HD;
  CUtensorMap tx1, tx2, tv, tcv, tb, tom, toc;
  int rc;
  if ((rc = w_tmap3(&tx1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xw.X, xrow, N, B, xrow, xrow * N, HD, TILE_M))) return rc;
  if ((rc = w_tmap3(&tx2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xw.X, xrow, N, B, xrow, xrow * N, HD, n_pad))) return rc;
  if ((rc = w_tmap3(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv_mean, qrow, N, B, qrow, qrow * N, HD, n_pad))) return rc;
  if ((rc = w_tmap3(&tcv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv_cov, qrow, N, B, qrow, qrow * N, HD, n_pad))) return rc;
  if ((rc = w_tmap3(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, bias, ld_bias, N, H, ld_bias, (uint64_t)ld_bias * N, 32, TILE_M))) return rc;
  if ((rc = w_tmap3(&tom, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_mean, orow, N, B, orow, orow * N, HD, TILE_M))) return rc;
  if ((rc = w_tmap3(&toc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_cov, orow, N, B, orow, orow * N, HD, TILE_M))) return rc;
  static bool configured[2] = {false, false};
  const bool drop = p_drop > 0.f;
  if (!configured[drop]) {
    cudaError_t e = drop ? cudaFuncSetAttribute(wattn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM)
                         : cudaFuncSetAttribute(wattn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("wattn_fwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured[drop] = true;
  }
  if (drop && !keep_ready) {
    if ((rc = b200vit_keep_bits_launch(keep_bits, B * H, N, p_drop, seed, seed_dev, stream_id, keep_in, stream_))) return rc;
  }
  const int ctas = p.items < sms ? p.items : sms;
  if (drop) wattn_fwd_kernel<true><<<ctas, F_THREADS, F_SMEM, stream>>>(tx1, tx2, tv, tcv, tb, tom, toc, p);
  else wattn_fwd_kernel<false><<<ctas, F_THREADS, F_SMEM, stream>>>(tx1, tx2, tv, tcv, tb, tom, toc, p);
  B200_CHECK_LAUNCH("wattn_fwd");
  return 0;
}

int b200vit_relbias_grad_launch(const void* ds_work, int B, int H, int N, int ld_ds, const int32_t* rel_index, float* dtable, void* stream);

extern "C" size_t b200vit_wattn_bwd_workspace_bytes(int32_t B, int32_t H, int32_t N, int32_t with_dtable) {
  if (B <= 0 || H <= 0 || N <= 0) return 0;
  const size_t n_pad = (size_t)(N + 15) / 16 * 16, rows = (size_t)B * H * N;
  return (with_dtable ? 2 : 1) * w_align256(rows * n_pad * sizeof(bf16)) + w_align256(rows * sizeof(float)) + w_align256(rows * 8 * sizeof(uint32_t));
}

extern "C" int b200vit_wattn_bwd(const void* qkv_mean, const void* qkv_cov, const void* xwork, const void* out_mean, const void* out_cov,
                                 const void* dout_mean, const void* dout_cov, const float* lse, const float* bias_t, int64_t ld_bias,
                                 const uint8_t* keep_bits, void* work, const int32_t* rel_index, float* dtable, float* dq_bias, float* dv_bias,
                                 float* dcq_bias, float* dcv_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop,
                                 void* dqkv_mean, void* dqkv_cov, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv_mean && qkv_cov && xwork && out_mean && out_cov && dout_mean && dout_cov && lse && dqkv_mean && dqkv_cov && work,
                 "wattn_bwd: null pointer (xwork of the forward and the workspace of b200vit_wattn_bwd_workspace_bytes are required)");
  B200_CHECK_ARG(B > 0 && H > 0 && head_dim == HD && N > 0 && N <= NMAX, "wattn_bwd: head_dim 64 and N <= %d only", NMAX);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || keep_bits != nullptr), "wattn_bwd: dropout needs keep_bits from the forward");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(bias_t != nullptr && ld_bias >= n_pad && ld_bias % 4 == 0 && (reinterpret_cast<uintptr_t>(bias_t) & 15) == 0,
                 "wattn_bwd: needs the transposed padded bias of b200vit_rel_pos_bias");
  B200_CHECK_ARG(dtable == nullptr || rel_index != nullptr, "wattn_bwd: dtable needs rel_index");
  const uintptr_t addrs[] = {(uintptr_t)qkv_mean, (uintptr_t)qkv_cov, (uintptr_t)out_mean, (uintptr_t)out_cov, (uintptr_t)dout_mean, (uintptr_t)dout_cov,
                             (uintptr_t)dqkv_mean, (uintptr_t)dqkv_cov, (uintptr_t)keep_bits};
  for (uintptr_t a : addrs) B200_CHECK_ARG((a & 15) == 0, "wattn_bwd: tensors must be 16-byte aligned");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(work) & 255) == 0 && (reinterpret_cast<uintptr_t>(xwork) & 255) == 0, "wattn_bwd: workspaces must be 256-byte aligned");
  const XWork xw = split_xwork(const_cast<void*>(xwork), B, H, N);
  const size_t rows = (size_t)B * H * N;
  uint8_t* wp = static_cast<uint8_t*>(work);
  bf16* dd = reinterpret_cast<bf16*>(wp);
  wp += w_align256(rows * n_pad * sizeof(bf16));
  float* dvec = reinterpret_cast<float*>(wp);
  wp += w_align256(rows * sizeof(float));
  uint32_t* keep_t = reinterpret_cast<uint32_t*>(wp);
  wp += w_align256(rows * 8 * sizeof(uint32_t));
  bf16* da = dtable != nullptr ? reinterpret_cast<bf16*>(wp) : nullptr;
  const int sms = b200vit_num_sms();
  const bool drop = p_drop > 0.f;

  wattn_bwd_prep_kernel<<<sms * 8, 256, 0, stream>>>(static_cast<const bf16*>(out_mean), static_cast<const bf16*>(dout_mean), static_cast<const bf16*>(out_cov),
                                                     static_cast<const bf16*>(dout_cov), dvec, B, H, N);
  B200_CHECK_LAUNCH("wattn_bwd_prep");
  if (drop) {
    w_keep_transpose_kernel<<<B * H, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(keep_bits), keep_t, N);
    B200_CHECK_LAUNCH("wattn_keep_transpose");
  }

  WKvParams kp;
  kp.lse = lse; kp.dvec = dvec; kp.rn = xw.rn; kp.cn = xw.cn; kp.keep_t = keep_t; kp.qkv_c = static_cast<const bf16*>(qkv_cov);
  kp.dd_out = dd; kp.da_out = da; kp.ld = n_pad; kp.dqkv_m = static_cast<bf16*>(dqkv_mean); kp.dqkv_c = static_cast<bf16*>(dqkv_cov);
  kp.dv_bias = dv_bias; kp.dcv_bias = dcv_bias; kp.B = B; kp.H = H; kp.N = N; kp.n_pad = n_pad; kp.k_tiles = (N + TILE_M - 1) / TILE_M;
  kp.items = B * H * kp.k_tiles; kp.inv_keep = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  const uint64_t xrow = 2ull * H * XW, qrow = 3ull * H * HD, orow = (uint64_t)H * HD;
  CUtensorMap tx1, tx2, tv, tcv, tdom, tdoc, tb, tdd, txf;
  int rc;
  if ((rc = w_tmap3(&tx1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xw.X, xrow, N, B, xrow, xrow * N, HD, BOXQ))) return rc;
  if ((rc = w_tmap3(&tx2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xw.X, xrow, N, B, xrow, xrow * N, HD, TILE_M))) return rc;
  if ((rc = w_tmap3(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv_mean, qrow, N, B, qrow, qrow * N, HD, TILE_M))) return rc;
  if ((rc = w_tmap3(&tcv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv_cov, qrow, N, B, qrow, qrow * N, HD, TILE_M))) return rc;
  if ((rc = w_tmap3(&tdom, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout_mean, orow, N, B, orow, orow * N, HD, BOXQ))) return rc;
  if ((rc = w_tmap3(&tdoc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout_cov, orow, N, B, orow, orow * N, HD, BOXQ))) return rc;
  if ((rc = w_tmap3(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, bias_t, ld_bias, N, H, ld_bias, (uint64_t)ld_bias * N, 16, TILE_M, CU_TENSOR_MAP_SWIZZLE_64B)))
    return rc;
  static bool kv_configured[2] = {false, false};
  if (!kv_configured[drop]) {
    cudaError_t e = drop ? cudaFuncSetAttribute(wattn_bwd_kv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM)
                         : cudaFuncSetAttribute(wattn_bwd_kv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("wattn_bwd (kv): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    kv_configured[drop] = true;
  }
  const int kv_ctas = kp.items < sms ? kp.items : sms;
  if (drop) wattn_bwd_kv_kernel<true><<<kv_ctas, KV_THREADS, KV_SMEM, stream>>>(tx1, tx2, tv, tcv, tdom, tdoc, tb, kp);
  else wattn_bwd_kv_kernel<false><<<kv_ctas, KV_THREADS, KV_SMEM, stream>>>(tx1, tx2, tv, tcv, tdom, tdoc, tb, kp);
  B200_CHECK_LAUNCH("wattn_bwd_kv");

  if ((rc = w_tmap3(&tdd, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dd, n_pad, N, (uint64_t)B * H, n_pad, (uint64_t)n_pad * N, 64, n_pad))) return rc;
  if ((rc = w_tmap3(&txf, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xw.X, xrow, N, B, xrow, xrow * N, HD, n_pad))) return rc;
  WDxParams dp;
  dp.X = xw.X; dp.qkv_c = static_cast<const bf16*>(qkv_cov); dp.dqkv_m = static_cast<bf16*>(dqkv_mean); dp.dqkv_c = static_cast<bf16*>(dqkv_cov);
  dp.dq_bias = dq_bias; dp.dcq_bias = dcq_bias; dp.B = B; dp.H = H; dp.N = N; dp.n_pad = n_pad; dp.items = B * H; dp.scale = scale;
  static bool dx_configured = false;
  if (!dx_configured) {
    cudaError_t e = cudaFuncSetAttribute(wattn_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DX_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("wattn_bwd (dx): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    dx_configured = true;
  }
  wattn_bwd_dx_kernel<<<dp.items < sms ? dp.items : sms, DX_THREADS, DX_SMEM, stream>>>(tdd, txf, dp);
  B200_CHECK_LAUNCH("wattn_bwd_dx");
  if (dtable != nullptr) return b200vit_relbias_grad_launch(da, B, H, N, n_pad, rel_index, dtable, stream);
  return 0;
}
