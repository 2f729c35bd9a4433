// tcgen05 / TMEM attention forward for ViT token counts (N <= 208, head_dim 64) on sm_100a.
//   S = scale * q k^T + rel_pos_bias[h] ; P = softmax_j(S) ; P~ = dropout(P) ; O = P~ v        (Attention.forward, modeling_finetune.py:145-188)
//
// One CTA per (batch, head, 128-query tile), two CTAs resident per SM (256 TMEM columns each):
//   warp 4, one lane : TMA loads (Q tile, K, V as 3-D boxes of the un-permuted [B, N, 3, H, 64] QKV GEMM output — rows past N are
//                      zero-filled by the TMA unit; the log2(e)-prescaled, -inf padded bias tile streams through a 2-stage ring of
//                      [128 x 32] fp32 SWIZZLE_128B boxes), tcgen05.mma S = Q K^T (128 x n_pad x 64, accumulator in TMEM columns
//                      [0, n_pad)), later tcgen05.mma O = P~ V with the A operand read straight from TENSOR MEMORY (the bf16
//                      probabilities overlay the S columns they were computed from) and V as an MN-major smem operand.
//   warps 0..7       : TWO threads per query row (TMEM lane == row; warps w and w + 4 share lane quadrant w): the even 32-key chunks of a
//                      row belong to one, the odd chunks to the other, the row maximum and the row sum meet through shared memory
//                      (the kernel is bound by this element-wise phase: one thread per row left the SM at 12 of 64 warps).
//                      pass 1: s2 = s * scale*log2e + bias, running max, s2 written back to TMEM;
//                      pass 2: p = 2^(s2 - max), row sum, dropout by the precomputed keep bits, bf16 pairs stored over score columns
//                      the SAME thread has already consumed (p_col below), so the two threads of a row never touch each other's columns;
//                      epilogue: O (TMEM columns [128, 192)) * 1/((1-p) * rowsum) -> bf16 -> swizzled smem tile -> TMA store
//                      (rows past N clipped by the tensor map), log-sum-exp to global.
#include <cstdlib>
#include <mutex>

#include "../../include/b200vit.h"
#include "attn_common.cuh"
#include "ptx_sm100.cuh"

namespace {

using namespace attn;

constexpr int TILE_M = 128;
constexpr int BIAS_STAGES = 2;
constexpr int BIAS_STAGE_BYTES = TILE_M * 128;           // 128 rows x 32 fp32
// per-lane shared memory of the forward kernel (a CTA holds FWD_LANES independent lanes)
constexpr int SM_Q = 0;                                   // 128 rows x 128 B
constexpr int SM_K = SM_Q + TILE_M * 128;                 // NMAX rows x 128 B
constexpr int SM_V = SM_K + NMAX * 128;                   // NMAX rows x 128 B; its first 16 KB double as the bf16 O staging tile of the epilogue
constexpr int SM_BIAS = SM_V + NMAX * 128;               // dense bias: the TMA ring; indexed bias: the head's table row (<= 4 KB)
constexpr int IDX_PITCH_W = B200VIT_ATTN_IDX_PITCH / 2;   // 105 words per index row: rows 9 banks apart, conflict-free one-thread-per-row reads
constexpr int IDX_TILE_BYTES = TILE_M * B200VIT_ATTN_IDX_PITCH * 2;   // 53 760
constexpr int FWD_LANES = 2;
constexpr int FWD_QUADS = 4;                              // TMEM lane quadrants of a 128-row tile
constexpr int FWD_EW_WARPS = 2 * FWD_QUADS;               // softmax warps per lane: two threads per query row
constexpr int FWD100_THREADS = FWD_LANES * (FWD_EW_WARPS + 2) * 32;   // + MMA warp + TMA warp per lane
// BIAS: 0 none, 1 dense fp32 [H, N, ld] through a TMA ring, 2 indexed (table row + resident uint16 index tile shared by both lanes)
__host__ __device__ constexpr int fwd_sm_bar(int bias) { return SM_BIAS + (bias == 2 ? B200VIT_ATTN_TAB_MAX * 4 : BIAS_STAGES * BIAS_STAGE_BYTES); }
__host__ __device__ constexpr int fwd_sm_xch(int bias) { return fwd_sm_bar(bias) + 1024; }   // [2 halves][128 rows] row maxima, then row sums
__host__ __device__ constexpr int fwd_lane_bytes(int bias) { return fwd_sm_xch(bias) + 2048; }
__host__ __device__ constexpr int fwd_smem(int bias) { return FWD_LANES * fwd_lane_bytes(bias) + (bias == 2 ? IDX_TILE_BYTES : 0) + 1024; }
constexpr int O_COL = 128;                                // O accumulator: TMEM columns [128, 192) of the lane's 256-column half
static_assert(SM_K % 1024 == 0 && SM_V % 1024 == 0 && SM_BIAS % 1024 == 0 && fwd_lane_bytes(1) % 1024 == 0 && fwd_lane_bytes(2) % 1024 == 0,
              "SWIZZLE_128B tiles need 1024-byte alignment");
static_assert(fwd_smem(1) <= 232448 && fwd_smem(2) <= 232448, "forward kernel smem");

struct Fwd100Params {
  float* lse;               // [B, H, N]
  uint8_t* keep_bits;       // [B, H, N, 32]
  const uint8_t* keep_in;   // [B, H, N, N] or null
  int B, H, N, n_pad, m_tiles, items;
  float sl2, inv_keep;
  uint32_t thresh;
  uint64_t seed;
  const uint64_t* seed_dev;  // when non-null the Philox key is read from device memory (a captured CUDA graph re-keys every replay)
  uint32_t stream_id;
  const uint16_t* bias_idx16;   // indexed bias (BIAS == 2): [m_tiles][128][B200VIT_ATTN_IDX_PITCH]
  const float* bias_tab;        // [H][tab_pitch]
  int tab_pitch;
};

// Keep bits of the 32 keys [32c, 32c + 32) of query row i — bit e = key 32c + e — from the packed mask b200vit_keep_bits_launch wrote
// before this kernel (Philox stream of dropout_group / dropout_u16 in attn_common.cuh, or the injected mask).
__device__ __forceinline__ uint32_t keep_word(const Fwd100Params& p, int bh, int i, int c) {
  return i < p.N ? __ldg(reinterpret_cast<const uint32_t*>(p.keep_bits) + ((long long)bh * p.N + i) * 8 + c) : 0u;
}

template <int COLS, int BIAS>
__device__ __forceinline__ void pass1_chunk(const Fwd100Params& p, uint32_t taddr, const uint8_t* bias_row, const uint32_t* idx_row, const float* tab,
                                            int row, int c, float& mx) {
  uint32_t s[COLS];
  if constexpr (COLS == 32) ptx::tmem_ld_x32_sync(taddr, reinterpret_cast<uint32_t(&)[32]>(s));
  else ptx::tmem_ld_x16_sync(taddr, reinterpret_cast<uint32_t(&)[16]>(s));
#pragma unroll
  for (int q = 0; q < COLS / 4; ++q) {
    float v[4];
    if (BIAS == 1) {   // bias is pre-multiplied by log2(e); padding columns hold -inf (key mask); SWIZZLE_128B: 16-byte chunk ^ (row & 7)
      const float4 b4 = *reinterpret_cast<const float4*>(bias_row + ((q ^ (row & 7)) << 4));
      v[0] = fmaf(__uint_as_float(s[4 * q]), p.sl2, b4.x); v[1] = fmaf(__uint_as_float(s[4 * q + 1]), p.sl2, b4.y);
      v[2] = fmaf(__uint_as_float(s[4 * q + 2]), p.sl2, b4.z); v[3] = fmaf(__uint_as_float(s[4 * q + 3]), p.sl2, b4.w);
    } else if (BIAS == 2) {   // gather from the head's table row through the resident index tile (entry nbins = -inf for the padding columns)
      const uint32_t w0 = idx_row[2 * q], w1 = idx_row[2 * q + 1];
      v[0] = fmaf(__uint_as_float(s[4 * q]), p.sl2, tab[w0 & 0xffffu]); v[1] = fmaf(__uint_as_float(s[4 * q + 1]), p.sl2, tab[w0 >> 16]);
      v[2] = fmaf(__uint_as_float(s[4 * q + 2]), p.sl2, tab[w1 & 0xffffu]); v[3] = fmaf(__uint_as_float(s[4 * q + 3]), p.sl2, tab[w1 >> 16]);
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

// 16 probabilities of one row: p = 2^(s2 - max), row sum, dropout by the keep bits `w16` (bit e = key e of this step), bf16 pairs to TMEM
template <bool DROP>
__device__ __forceinline__ void pass2_step16(uint32_t t_src, uint32_t t_dst, float mx, float& l, uint32_t w16) {
  uint32_t s[16];
  ptx::tmem_ld_x16_sync(t_src, s);
  float pr[16];
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    pr[e] = ex2(__uint_as_float(s[e]) - mx);
    acc += pr[e];
  }
  l += acc;
  if (DROP) {
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (!(w16 & (1u << e))) pr[e] = 0.f;
  }
  uint32_t pk[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) pk[e] = pack_bf16x2(pr[2 * e], pr[2 * e + 1]);
  ptx::tmem_st_x8(t_dst, pk);
}

// TMEM column (fp32 units) of the 16 bf16 pairs of key chunk c: chunks 0..3 overwrite the lower half of their own score columns, chunks
// 4..6 the upper half of the scores of chunk c - 4 — same parity, hence the same thread, which has consumed them by then. All of P lands
// in columns [0, 128), clear of the O accumulator at [128, 192) (the dead scores of chunks 4 and 5).
__host__ __device__ __forceinline__ constexpr int p_col(int c) { return c < 4 ? 32 * c : 32 * (c - 4) + 16; }

template <int COLS, bool DROP>
__device__ __forceinline__ void pass2_chunk(const Fwd100Params& p, uint64_t seed, uint32_t trow, int bh, int i, int c, float mx, float& l) {
  uint32_t w = 0xffffffffu;
  if (DROP) w = keep_word(p, bh, i, c);
  pass2_step16<DROP>(trow + c * 32, trow + p_col(c), mx, l, w);
  if constexpr (COLS == 32) pass2_step16<DROP>(trow + c * 32 + 16, trow + p_col(c) + 8, mx, l, w >> 16);
}

// Persistent: one CTA per SM holds FWD_LANES independent lanes (own smem, own 256-column TMEM half, own barriers); lane L of CTA c
// walks items (2c + L) + k * 2 * gridDim of the (batch, head, query-tile) list. Q / K of the next item are requested as soon as the
// S MMAs of the current one have read them, V once the epilogue's TMA store has drained the staging tile that aliases it, the bias
// ring simply runs on: the load / launch latency a one-shot CTA exposes per item is paid once per lane.
template <bool DROP, int BIAS>
__global__ void __launch_bounds__(FWD100_THREADS, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                      const __grid_constant__ CUtensorMap tm_bias, const __grid_constant__ CUtensorMap tm_out, const Fwd100Params p) {
  constexpr bool HAS_BIAS = BIAS == 1;          // the dense bias ring and its barriers
  constexpr int FWD_LANE_BYTES = fwd_lane_bytes(BIAS), SM_BAR = fwd_sm_bar(BIAS), SM_XCH = fwd_sm_xch(BIAS);
  extern __shared__ uint8_t smem_raw[];
  const int warp_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA warps 0..15: softmax warps (lane L = warp >> 3; TMEM lane quadrant = CTA warp index & 3, a hardware rule; chunk parity = bit 2);
  // 16, 17: MMA and TMA warp of lane 0; 18, 19: of lane 1.   `warp` = role inside the lane: 0..7 softmax, 8 MMA, 9 TMA
  const int L = warp_cta < FWD_LANES * FWD_EW_WARPS ? warp_cta / FWD_EW_WARPS : (warp_cta - FWD_LANES * FWD_EW_WARPS) >> 1;
  const int warp = warp_cta < FWD_LANES * FWD_EW_WARPS ? warp_cta % FWD_EW_WARPS : FWD_EW_WARPS + ((warp_cta - FWD_LANES * FWD_EW_WARPS) & 1);
  const uint32_t cta_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t base = cta_base + L * FWD_LANE_BYTES;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = base + SM_BAR;
  const uint32_t qk_full = bar0, v_full = bar0 + 8, s_full = bar0 + 16, p_full = bar0 + 24, o_full = bar0 + 32,
                 v_free = bar0 + 48, t_free = bar0 + 56;
  auto bias_full = [&](int s) { return bar0 + 64u + 8u * s; };
  auto bias_empty = [&](int s) { return bar0 + 80u + 8u * s; };
  const uint32_t tmem_slot = cta_base + SM_BAR + 128u;

  const int n_pad = p.n_pad;
  const int nchunks = (n_pad + 31) >> 5;
  const int tail_cols = n_pad - (nchunks - 1) * 32;   // 16 or 32
  // dense / no bias: lane L of CTA c walks items (2c + L) + k * 2 * gridDim of the (batch, head, query-tile) list. Indexed bias: a CTA keeps ONE
  // query tile (its index tile is resident) and its lanes walk the (batch, head) pairs: mt = c % m_tiles, gridDim % m_tiles == 0 (host)
  const int cta_mt = BIAS == 2 ? (int)blockIdx.x % p.m_tiles : 0;
  const int first = BIAS == 2 ? ((int)blockIdx.x / p.m_tiles) * FWD_LANES + L : (int)blockIdx.x * FWD_LANES + L;
  const int stride = BIAS == 2 ? ((int)gridDim.x / p.m_tiles) * FWD_LANES : (int)gridDim.x * FWD_LANES;
  const int total_items = BIAS == 2 ? p.B * p.H : p.items;
  const int n_items = first < total_items ? (total_items - first + stride - 1) / stride : 0;
  auto item_of = [&](int it, int& b, int& h, int& m0, int& bh) {
    const int item = first + it * stride;
    const int mt = BIAS == 2 ? cta_mt : item % p.m_tiles;
    bh = BIAS == 2 ? item : item / p.m_tiles;
    b = bh / p.H; h = bh - b * p.H; m0 = mt * TILE_M;
  };
  // indexed bias: the uint16 index tile of this CTA's query tile, loaded once (both lanes read it), and the lane's table row
  const uint32_t* idx_s = reinterpret_cast<const uint32_t*>(smem_raw + (cta_base - ptx::smem_u32(smem_raw)) + FWD_LANES * FWD_LANE_BYTES);
  float* tab_s = reinterpret_cast<float*>(gbase + SM_BIAS);
  if constexpr (BIAS == 2) {
    const uint4* src = reinterpret_cast<const uint4*>(p.bias_idx16 + (size_t)cta_mt * (IDX_TILE_BYTES / 2));
    uint4* dst = reinterpret_cast<uint4*>(smem_raw + (cta_base - ptx::smem_u32(smem_raw)) + FWD_LANES * FWD_LANE_BYTES);
    for (int i = threadIdx.x; i < IDX_TILE_BYTES / 16; i += FWD100_THREADS) dst[i] = __ldg(src + i);
  }

  if (warp == FWD_EW_WARPS) {
    if (lane == 0) {
      ptx::mbar_init(qk_full, 1); ptx::mbar_init(v_full, 1); ptx::mbar_init(s_full, 1); ptx::mbar_init(p_full, FWD_EW_WARPS * 32);
      ptx::mbar_init(o_full, 1); ptx::mbar_init(v_free, 1); ptx::mbar_init(t_free, FWD_EW_WARPS);
      for (int s = 0; s < BIAS_STAGES; ++s) { ptx::mbar_init(bias_full(s), 1); ptx::mbar_init(bias_empty(s), FWD_QUADS); }   // a chunk is read by one parity
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_kv); ptx::prefetch_tmap(&tm_out);
      if (HAS_BIAS) ptx::prefetch_tmap(&tm_bias);
    }
    __syncwarp();
    if (L == 0) {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base += L * 256;

  if (warp == FWD_EW_WARPS) {
    if (lane == 0 && n_items > 0) {
      // ---------------- MMA issue ----------------
      const uint64_t dQ = ptx::make_smem_desc(base + SM_Q, 16, 1024), dK = ptx::make_smem_desc(base + SM_K, 16, 1024);
      const uint64_t dV = ptx::make_smem_desc(base + SM_V, NMAX * 128, 1024);                   // MN-major: key rows of 64 d-elements
      const uint32_t idesc_s = ptx::make_idesc_bf16(TILE_M, n_pad, false, false), idesc_o = ptx::make_idesc_bf16(TILE_M, HD, false, true);
      const int ksteps = n_pad >> 4;
      for (int it = 0; it < n_items; ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        ptx::mbar_wait(qk_full, ph);
        if (it > 0) { ptx::mbar_wait(t_free, ph ^ 1u); ptx::tc_fence_after(); }    // the epilogue has drained O of the previous item
        ptx::umma_bf16(tmem_base, dQ, dK, idesc_s, 0u);
        ptx::umma_bf16(tmem_base, dQ + 2, dK + 2, idesc_s, 1u);
        ptx::umma_bf16(tmem_base, dQ + 4, dK + 4, idesc_s, 1u);
        ptx::umma_bf16(tmem_base, dQ + 6, dK + 6, idesc_s, 1u);
        ptx::umma_commit(s_full);                 // also tells the TMA thread that Q / K have been read (one completion per item)
        // O = P~ V : A = bf16 probabilities in TMEM columns [0, n_pad/2), B = V
        ptx::mbar_wait(v_full, ph);
        ptx::mbar_wait(p_full, ph);
        ptx::tc_fence_after();
        for (int kk = 0; kk < ksteps; ++kk)
          ptx::umma_bf16_ts(tmem_base + O_COL, tmem_base + p_col(kk >> 1) + 8 * (kk & 1), dV + 128 * kk, idesc_o, kk > 0 ? 1u : 0u);
        ptx::umma_commit(o_full);
      }
    }
  } else if (warp == FWD_EW_WARPS + 1) {
    if (lane == 0 && n_items > 0) {
      // ---------------- TMA: Q, K (freed by the S MMAs) and V (freed by the epilogue's TMA store) ----------------
      for (int it = 0; it < n_items; ++it) {
        int b, h, m0, bh;
        item_of(it, b, h, m0, bh);
        if (it > 0) ptx::mbar_wait(s_full, (uint32_t)((it - 1) & 1));      // S MMAs of the previous item complete: Q / K are free
        ptx::mbar_arrive_expect_tx(qk_full, TILE_M * 128 + n_pad * 128);
        ptx::tma_load_3d(base + SM_Q, &tm_q, qk_full, h * HD, m0, b);
        ptx::tma_load_3d(base + SM_K, &tm_kv, qk_full, (p.H + h) * HD, 0, b);
        if (it > 0) ptx::mbar_wait(v_free, (uint32_t)((it - 1) & 1));
        ptx::mbar_arrive_expect_tx(v_full, n_pad * 128);
        ptx::tma_load_3d(base + SM_V, &tm_kv, v_full, (2 * p.H + h) * HD, 0, b);
      }
    } else if (HAS_BIAS && lane == 1 && n_items > 0) {
      // ---------------- TMA: bias ring, [128 queries x 32 keys] fp32 boxes ----------------
      int gc = 0;
      for (int it = 0; it < n_items; ++it) {
        int b, h, m0, bh;
        item_of(it, b, h, m0, bh);
        for (int c = 0; c < nchunks; ++c, ++gc) {
          const int s = gc % BIAS_STAGES;
          if (gc >= BIAS_STAGES) ptx::mbar_wait(bias_empty(s), (uint32_t)(((gc / BIAS_STAGES) - 1) & 1));
          ptx::mbar_arrive_expect_tx(bias_full(s), BIAS_STAGE_BYTES);
          ptx::tma_load_3d(base + SM_BIAS + s * BIAS_STAGE_BYTES, &tm_bias, bias_full(s), c * 32, m0, h);
        }
      }
    }
  } else {
    // ---------------- softmax warps: two threads per query row ----------------
    const int quad = warp & 3, par = warp >> 2;              // TMEM lane quadrant; parity of the key chunks this thread owns
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint64_t seed = (DROP && p.seed_dev != nullptr) ? __ldg(reinterpret_cast<const unsigned long long*>(p.seed_dev)) : p.seed;
    float* xch_mx = reinterpret_cast<float*>(gbase + SM_XCH);          // [2][128]
    float* xch_l = xch_mx + 2 * TILE_M;                                 // [2][128]
    for (int it = 0; it < n_items; ++it) {
      int b, h, m0, bh;
      item_of(it, b, h, m0, bh);
      const uint32_t ph = (uint32_t)(it & 1);
      const int i = m0 + row;
      const bool active = m0 + quad * 32 < p.N;            // warps whose 32 rows are all past N only keep the barrier protocol alive
      const int gc0 = it * nchunks;                         // the bias ring counts chunks across items
      if constexpr (BIAS == 2) {
        if (it == 0) {                                       // the first item's table row (later ones are fetched during the previous item)
          const int et = (warp * 32 + lane);
          for (int k2 = et; k2 < p.tab_pitch; k2 += FWD_EW_WARPS * 32) tab_s[k2] = __ldg(p.bias_tab + (size_t)h * p.tab_pitch + k2);
          ptx::named_bar_sync(1 + L, FWD_EW_WARPS * 32);
        }
      }
      ptx::mbar_wait(s_full, ph);
      ptx::tc_fence_after();
      float mx = -INFINITY;
      for (int c = par; c < nchunks; c += 2) {
        const int gc = gc0 + c, s = gc % BIAS_STAGES;
        if (HAS_BIAS) ptx::mbar_wait(bias_full(s), (uint32_t)((gc / BIAS_STAGES) & 1));
        if (active) {
          const uint8_t* bias_row = gbase + SM_BIAS + s * BIAS_STAGE_BYTES + row * 128;
          const uint32_t* idx_row = idx_s + row * IDX_PITCH_W + c * 16;
          if (c + 1 < nchunks || tail_cols == 32) pass1_chunk<32, BIAS>(p, trow + c * 32, bias_row, idx_row, tab_s, row, c, mx);
          else pass1_chunk<16, BIAS>(p, trow + c * 32, bias_row, idx_row, tab_s, row, c, mx);
        }
        if (HAS_BIAS) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bias_empty(s));
        }
      }
      xch_mx[par * TILE_M + row] = mx;
      ptx::tmem_st_wait();
      ptx::named_bar_sync(1 + L, FWD_EW_WARPS * 32);
      mx = fmaxf(mx, xch_mx[(par ^ 1) * TILE_M + row]);
      // indexed bias: every thread of the lane is past pass 1, the table row is dead: fetch the next item's (in flight across pass 2)
      float tnext[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (BIAS == 2) {
        if (it + 1 < n_items) {
          int b2, h2, m2, bh2;
          item_of(it + 1, b2, h2, m2, bh2);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const int e = (warp * 32 + lane) + k2 * FWD_EW_WARPS * 32;
            if (e < p.tab_pitch) tnext[k2] = __ldg(p.bias_tab + (size_t)h2 * p.tab_pitch + e);
          }
        }
      }
      float l = 0.f;
      if (active) {
        for (int c = par; c < nchunks; c += 2) {
          if (c + 1 < nchunks || tail_cols == 32) pass2_chunk<32, DROP>(p, seed, trow, bh, i, c, mx, l);
          else pass2_chunk<16, DROP>(p, seed, trow, bh, i, c, mx, l);
        }
      }
      xch_l[par * TILE_M + row] = l;
      if constexpr (BIAS == 2) {
        if (it + 1 < n_items) {
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const int e = (warp * 32 + lane) + k2 * FWD_EW_WARPS * 32;
            if (e < p.tab_pitch) tab_s[e] = tnext[k2];
          }
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
      ptx::named_bar_sync(1 + L, FWD_EW_WARPS * 32);
      l += xch_l[(par ^ 1) * TILE_M + row];
      // ---------------- epilogue: this thread's 32 of the row's 64 output columns ----------------
      ptx::mbar_wait(o_full, ph);
      ptx::tc_fence_after();
      if (active) {
        const float inv = p.inv_keep / l;
        uint8_t* orow = gbase + SM_V + row * 128;           // V is dead (o_full) and the previous store has drained the tile (v_free)
        uint32_t o[32];
        ptx::tmem_ld_x32_sync(trow + O_COL + par * 32, o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + (((par * 4 + q) ^ (row & 7)) << 4)) = u;
        }
        if (par == 0 && i < p.N && p.lse != nullptr) p.lse[(long long)bh * p.N + i] = (mx + log2f(l)) / LOG2E;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(t_free);              // the next item's S MMAs may overwrite this lane's TMEM half
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1 + L, FWD_EW_WARPS * 32);
      if (warp == 0 && lane == 0) {
        ptx::tma_store_3d(&tm_out, base + SM_V, h * HD, m0, b);
        ptx::bulk_commit();
        ptx::bulk_wait_read0();
        ptx::mbar_arrive(v_free);                           // the TMA warp may now overwrite V (and with it the staging tile)
      }
    }
    if (warp == 0 && lane == 0) ptx::bulk_wait0();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == FWD_EW_WARPS && L == 0) ptx::tmem_dealloc(tmem_base, 512);
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

template <bool DROP, int BIAS>
cudaError_t launch_fwd100(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tb, const CUtensorMap& to, const Fwd100Params& p,
                          cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_sm100_kernel<DROP, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(BIAS));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int ctas = min(b200vit_num_sms(), (p.items + FWD_LANES - 1) / FWD_LANES);
  if (BIAS == 2) {                       // a CTA owns one query tile: whole groups of m_tiles CTAs
    ctas = ctas / p.m_tiles * p.m_tiles;
    if (ctas < p.m_tiles) ctas = p.m_tiles;
  }
  attn_fwd_sm100_kernel<DROP, BIAS><<<ctas, FWD100_THREADS, fwd_smem(BIAS), stream>>>(tq, tkv, tb, to, p);
  return cudaGetLastError();
}


// =================================================================================================================
// backward
// =================================================================================================================
//   P = 2^(s2 - lse2) ; dP~ = dO V^T ; dP = c M o dP~ ; dS = P o (dP - D) ; dV = P~^T dO ; dK = scale dS^T Q ; dQ = scale dS K
// Three kernels:
//   attn_bwd_prep_kernel  : D_i = sum_d dO[i,d] O[i,d] and the keep mask transposed to [key][query bits] (one pass over O, dO).
//   attn_bwd_kv_kernel    : one CTA per (batch, head, 128-KEY tile), two CTAs per SM. TMEM lane == key row, one thread per key
//                           row. Per 64-query quarter: S^T = K Q^T and dP~^T = V dO^T (tcgen05, 64 TMEM columns each), the
//                           element-wise phase straight out of TMEM (bias^T tile through a TMA ring, lse / D broadcast from smem,
//                           keep bits from registers), P~^T and dS^T written back to TMEM as bf16 and consumed in place as the
//                           A operands of dV += P~^T dO and dK += dS^T Q (accumulators in TMEM columns [128, 256)); dS^T is also
//                           streamed to the workspace for the dQ kernel and the relative-position-bias table gradient.
//   attn_bwd_dq_kernel    : one CTA per (batch, head, 128-query tile): dQ = scale * dS K with dS^T read back through TMA as an
//                           MN-major A operand, K as an MN-major B operand; q_bias gradient fused.
// per-lane shared memory (a CTA holds KV_LANES independent lanes, each the size of a classic CTA)
constexpr int KV_SM_K = 0;                          // 128 key rows x 128 B
constexpr int KV_SM_V = 16384;
constexpr int KV_SM_Q = 32768;                      // 4 slots x (32 query rows x 128 B)
constexpr int KV_SM_DO = 49152;                     // 4 slots x (32 query rows x 128 B)
constexpr int KV_SM_BIAS = 65536;                   // 2 stages x (128 key rows x 32 fp32)
constexpr int KV_SM_LSE = 98304;                    // 2 buffers x NMAX fp32 (log2 domain; +inf past N)
constexpr int KV_SM_D = KV_SM_LSE + 2048;           // 2 buffers x NMAX fp32
constexpr int KV_SM_BAR = KV_SM_D + 2048;
constexpr int KV_LANE_BYTES = KV_SM_BAR + 1024;     // 103424 (multiple of 1024)
constexpr int KV_LANES = 2;
constexpr int KV_EW_WARPS = 8;                      // element-wise warps per lane: warp w owns TMEM lane quadrant w & 3, query half w >> 2
constexpr int KV_LANE_WARPS = KV_EW_WARPS + 2;      // + TMA/MMA warp + bias-ring warp
constexpr int KV_THREADS = KV_LANES * KV_LANE_WARPS * 32;
constexpr int KV_SMEM = KV_LANES * KV_LANE_BYTES + 1024;
constexpr int KV_X = 0, KV_Y = 64, KV_DV = 128, KV_DK = 192;   // TMEM columns inside a lane's 256-column half (X_g = KV_X + 32 g)
static_assert(KV_LANE_BYTES % 1024 == 0 && KV_SMEM <= 232448, "kv kernel smem layout");

// profiling experiment: clock64 timeline of lane 0 of CTA 0 — compiled in only with -DB200VIT_KV_TRACE (B200VIT_EXTRA_NVCC_FLAGS, see
// tools/micro/kv_trace.py); a normal build has no trace code on the MMA-issuing thread that paces the kernel
#ifdef B200VIT_KV_TRACE
__device__ long long g_kv_trace[3 * 512 * 4];
__device__ __forceinline__ void kv_trace(int debug, int L, int role, int& n, int code, int it, int bi) {   // fire-and-forget stores, no atomics
  if (blockIdx.x == 0 && L == 0 && n < 512) {
    long long* e = g_kv_trace + (role * 512 + n) * 4;
    e[0] = code; e[1] = it; e[2] = bi; e[3] = clock64();
    ++n;
  }
}
#define KV_TRACE(...) kv_trace(__VA_ARGS__)
#else
#define KV_TRACE(...) ((void)0)
#endif

// 32-byte global store (sm_100: 256-bit vector stores): one LSU instruction per 32 bytes of a thread-private row segment
__device__ __forceinline__ void st_global_32B(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6,
                                              uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6),
               "r"(a7)
               : "memory");
}

struct BwdKvParams {
  const float* lse;          // [B, H, N]
  const float* dvec;         // [B, H, N]   D_i
  const uint32_t* keep_t;    // [B, H, N(key), 8]  bit (i % 32) of word i / 32 = keep(i, j)
  bf16* ds_out;              // [B, H, N(key), ld_ds(query)]
  int ld_ds;
  bf16* dqkv;                // [B, N, 3, H, 64]
  float* dv_bias;            // [H*64] += or null
  int B, H, N, n_pad, k_tiles, items;
  float scale, sl2, inv_keep;
  int debug;   // unused in normal builds (trace builds: -DB200VIT_KV_TRACE)
};

// column sums over the 32 lanes of a warp of 32 per-lane values (recursive halving, 31 shuffles); lane l ends with column l
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
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

// column sums over the 32 lanes of a warp of 64 per-lane values: recursive halving (62 shuffles); lane l ends with columns 2l, 2l+1
__device__ __forceinline__ void warp_colsum64(float (&v)[64], int lane, float& c0, float& c1) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 32], keep = up ? v[i + 32] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 2;
    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 1;
    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  c0 = v[0]; c1 = v[1];
}

// D[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d] : one warp per token row, 8 lanes per head, 4 heads per pass.
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ dvec,
                                                            int B, int H, int N) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * N;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(r / N), i = (int)(r - (long long)b * N);
    for (int h0 = 0; h0 < H; h0 += 4) {
      const int h = h0 + (lane >> 3);
      float acc = 0.f;
      if (h < H) {
        const long long off = (r * H + h) * HD + (lane & 7) * 8;
        const uint4 ov = *reinterpret_cast<const uint4*>(out + off), dv = *reinterpret_cast<const uint4*>(dout + off);
        const uint32_t* op = &ov.x; const uint32_t* dp = &dv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = unpack_bf16x2(op[k]), d = unpack_bf16x2(dp[k]);
          acc += a.x * d.x + a.y * d.y;
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (h < H && (lane & 7) == 0) dvec[((long long)b * H + h) * N + i] = acc;
    }
  }
}

// keep_bits [BH, N(query i), 8 words over keys] -> keep_t [BH, N(key j), 8 words over queries]. One CTA per (batch, head): the 6 KB
// bit matrix is staged in shared memory (9-word pitch: conflict-free column reads), warp jb transposes the 32x32 bit blocks of its
// 32 keys by ballot and writes whole 32-byte rows.
__global__ void __launch_bounds__(256) keep_transpose_kernel(const uint32_t* __restrict__ keep_bits, uint32_t* __restrict__ keep_t, int N) {
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
        uint32_t x = i < N ? tile[i * 9 + jb] : 0u;      // lane = query row, bit = key
        // 32x32 bit-matrix transpose: five butterfly steps (swap the off-diagonal s x s blocks between lanes l and l ^ s)
#pragma unroll
        for (int step = 0; step < 5; ++step) {
          const int s_ = 16 >> step;
          const uint32_t m = step == 0 ? 0x0000ffffu : step == 1 ? 0x00ff00ffu : step == 2 ? 0x0f0f0f0fu : step == 3 ? 0x33333333u : 0x55555555u;
          const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s_);
          x = (lane & s_) ? (((y >> s_) & m) | (x & ~m)) : ((x & m) | ((y << s_) & ~m));
        }
        mine[ib] = x;                                     // lane = key, bit = query
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

// 16 queries of one key row: P, dropout, dS out of TMEM; P~^T / dS^T written back as bf16 pairs (the A operands of the dV / dK MMAs)
// at columns the same warp has already consumed, dS^T also to the workspace row.
template <bool DROP, bool HAS_BIAS>
__device__ __forceinline__ void bwd_step16(const BwdKvParams& p, uint32_t tx, uint32_t ty, uint32_t tpx, uint32_t tpy, const uint8_t* bias_row, int q0,
                                           int row, const float* sl, const float* sd, uint32_t kw, bool valid, bf16* ds_row) {
  uint32_t x[16], y[16];
  ptx::tmem_ld_x16_pair_sync(tx, x, ty, y);
  uint32_t px[8], dx[8];
  if (valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (HAS_BIAS) b4 = *reinterpret_cast<const float4*>(bias_row + (((q0 + q) ^ (row & 7)) << 4));
      const float4 l4 = *reinterpret_cast<const float4*>(sl + 4 * q);
      const float4 d4 = *reinterpret_cast<const float4*>(sd + 4 * q);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
      float pt[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = 4 * q + e;
        const float pv = ex2(fmaf(__uint_as_float(x[idx]), p.sl2, bb[e]) - ll[e]);
        if (DROP) {   // P~ = f P, dP = f dP~ with f = keep / (1 - p)
          const float f = (kw & (1u << idx)) ? p.inv_keep : 0.f;
          pt[e] = f * pv;
          ds[e] = pv * fmaf(f, __uint_as_float(y[idx]), -dd[e]);
        } else {
          pt[e] = pv;
          ds[e] = pv * (__uint_as_float(y[idx]) - dd[e]);
        }
      }
      px[2 * q] = pack_bf16x2(pt[0], pt[1]); px[2 * q + 1] = pack_bf16x2(pt[2], pt[3]);
      dx[2 * q] = pack_bf16x2(ds[0], ds[1]); dx[2 * q + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    st_global_32B(ds_row, dx[0], dx[1], dx[2], dx[3], dx[4], dx[5], dx[6], dx[7]);
  } else {      // key rows past N: P = dS = 0 (their K / V rows are TMA zero fill, not -inf scores)
#pragma unroll
    for (int q = 0; q < 8; ++q) px[q] = dx[q] = 0u;
  }
  ptx::tmem_st_x8(tpx, px);
  ptx::tmem_st_x8(tpy, dx);
}

// mbarrier wait whose slow path (spin loop + timeout trap) is ONE out-of-line function: the kv kernel waits at ~20 sites and the inlined
// slow paths cost it instruction-cache misses ("no instruction" stalls); measured -14 us per launch. (The same change in the GEMM kernel
// halves ITS throughput - a call in a setmaxnreg kernel - and slows the forward kernel, so it stays local to this kernel.)
__device__ __noinline__ void wait_slow(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {
        printf("b200vit: mbarrier timeout (kv kernel, block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void wait_ool(uint32_t bar, uint32_t parity) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  wait_slow(bar, parity);
}

// Persistent: one CTA per SM holds KV_LANES independent lanes; lane L of CTA c walks items (2c + L) + k * 2 * gridDim of the
// (batch, head, key-tile) list. Inside a lane the 8 element-wise warps form two groups that ping-pong over the 32-query boxes of an
// item (group g owns TMEM columns X_g / Y_g): while group g computes box b, the tensor core runs dV/dK += of group 1-g's box and the
// score MMAs of its next one, so the MMA round trip hides behind the other group's arithmetic. K / V of the NEXT item are requested
// as soon as the last score MMA of the current item has read them and the Q / dO box ring runs on across item boundaries.
template <bool DROP, bool HAS_BIAS>
__global__ void __launch_bounds__(KV_THREADS, 1)
attn_bwd_kv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_bias, const BwdKvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA warps 0..15: element-wise warps (lane L = warp >> 3; TMEM lane quadrant = CTA warp index & 3, a hardware rule), 16..19: the
  // TMA+MMA warp and the bias-ring warp of lane 0, then of lane 1.  `warp` = role inside the lane: 0..7 element-wise, 8 TMA+MMA, 9 bias ring
  const int L = warp_cta < KV_LANES * KV_EW_WARPS ? warp_cta / KV_EW_WARPS : (warp_cta - KV_LANES * KV_EW_WARPS) >> 1;
  const int warp = warp_cta < KV_LANES * KV_EW_WARPS ? warp_cta % KV_EW_WARPS : KV_EW_WARPS + ((warp_cta - KV_LANES * KV_EW_WARPS) & 1);
  const uint32_t cta_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t base = cta_base + L * KV_LANE_BYTES;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = base + KV_SM_BAR;
  const uint32_t kv_full = bar0, kv_free = bar0 + 8, acc_full = bar0 + 16, acc_empty = bar0 + 24;
  auto s_full = [&](int g) { return bar0 + 32u + 8u * g; };
  auto p_full = [&](int g) { return bar0 + 48u + 8u * g; };
  auto bias_full = [&](int s) { return bar0 + 64u + 8u * s; };
  auto bias_empty = [&](int s) { return bar0 + 80u + 8u * s; };
  auto ld_full = [&](int s) { return bar0 + 96u + 8u * s; };     // 4 Q/dO box slots
  auto ld_empty = [&](int s) { return bar0 + 128u + 8u * s; };
  const uint32_t tmem_slot = cta_base + KV_SM_BAR + 192u;         // lane 0's barrier page holds the CTA-wide TMEM slot

  const int n_pad = p.n_pad;
  const int nboxes = (n_pad + 31) >> 5;        // 32-query boxes per item; group g owns boxes bi = g, g + 2, ...
  const int first = blockIdx.x * KV_LANES + L, stride = gridDim.x * KV_LANES;
  const int n_items = first < p.items ? (p.items - first + stride - 1) / stride : 0;
  auto item_of = [&](int it, int& b, int& h, int& j0, int& bh) {
    const int item = first + it * stride;
    const int kt = item % p.k_tiles;
    bh = item / p.k_tiles;
    b = bh / p.H; h = bh - b * p.H; j0 = kt * TILE_M;
  };
  // boxes of group g per item, and the running per-group box counter (barrier phases)
  auto group_count = [&](int it, int bi) { const int g = bi & 1; return it * ((nboxes + 1 - g) >> 1) + (bi >> 1); };

  if (warp == KV_EW_WARPS) {
    if (lane == 0) {
      ptx::mbar_init(kv_full, 1); ptx::mbar_init(kv_free, 2); ptx::mbar_init(acc_full, 2); ptx::mbar_init(acc_empty, KV_EW_WARPS);
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(s_full(s), 2); ptx::mbar_init(p_full(s), 128); ptx::mbar_init(bias_full(s), 1); ptx::mbar_init(bias_empty(s), 4);
      }
      for (int s = 0; s < 4; ++s) { ptx::mbar_init(ld_full(s), 1); ptx::mbar_init(ld_empty(s), 2); }
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_kv); ptx::prefetch_tmap(&tm_do);
      if (HAS_BIAS) ptx::prefetch_tmap(&tm_bias);
    }
    __syncwarp();
    if (L == 0) {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base += L * 256;

  if (warp == KV_EW_WARPS) {
    if (lane < 2 && n_items > 0) {
      // ---------------- MMA issue only: this code paces the whole lane, so it is lean and issued SIMT from TWO threads ----------------
      // thread m = 0 owns the S^T / dV chain (X = K Q^T, dV += P~^T dO), thread m = 1 the dP^T / dK chain (Y = V dO^T, dK += dS^T Q):
      // one warp instruction issues both MMAs; every completion barrier therefore expects two commits.
      // descriptor arithmetic: +2 in the encoded start address = +32 bytes (one K-major k-step), +128 = +2048 bytes (one MN-major k-step),
      // +256 = one 4 KB box slot
      const int m = lane;
      const uint64_t dA = ptx::make_smem_desc(base + (m == 0 ? KV_SM_K : KV_SM_V), 16, 1024);
      const uint64_t dBk = ptx::make_smem_desc(base + (m == 0 ? KV_SM_Q : KV_SM_DO), 16, 1024);      // score B operand, K-major
      const uint64_t dBm = ptx::make_smem_desc(base + (m == 0 ? KV_SM_DO : KV_SM_Q), 4096, 1024);    // accumulate B operand, MN-major
      const uint32_t t_score = tmem_base + (m == 0 ? KV_X : KV_Y), t_acc = tmem_base + (m == 0 ? KV_DV : KV_DK);
      const uint32_t idesc_acc = ptx::make_idesc_bf16(TILE_M, HD, false, true);
      const uint32_t idesc_s32 = ptx::make_idesc_bf16(TILE_M, 32, false, false);
      const uint32_t idesc_tail = ptx::make_idesc_bf16(TILE_M, n_pad - (nboxes - 1) * 32, false, false);
      const int tail_ksteps = (n_pad - (nboxes - 1) * 32) >> 4;
      int trn = 0;
      int cnt[2] = {0, 0};      // boxes accumulated so far per group
      int gb_s = 0;             // running box index of the NEXT score issue
      auto issue_scores = [&](int bi) {   // X_g = K Q_box^T | Y_g = V dO_box^T ; the last box of an item releases K / V
        const int sl = gb_s & 3, g = bi & 1;
        wait_ool(ld_full(sl), (uint32_t)((gb_s >> 2) & 1));      // TMA data: no tcgen05 fence needed
        if (m == 0) KV_TRACE(p.debug, L, 2, trn, 20, 0, bi);
        const uint32_t idesc = bi == nboxes - 1 ? idesc_tail : idesc_s32;
        const uint64_t db = dBk + (uint64_t)(sl * 256);
        const uint32_t d = t_score + g * 32;
        ptx::umma_bf16(d, dA, db, idesc, 0u);
        ptx::umma_bf16(d, dA + 2, db + 2, idesc, 1u);
        ptx::umma_bf16(d, dA + 4, db + 4, idesc, 1u);
        ptx::umma_bf16(d, dA + 6, db + 6, idesc, 1u);
        if (m == 0) KV_TRACE(p.debug, L, 2, trn, 21, 0, bi);
        ptx::umma_commit(s_full(g));
        if (bi == nboxes - 1) ptx::umma_commit(kv_free);
        ++gb_s;
      };
      int gb = 0;
      for (int it = 0; it < n_items; ++it) {
        wait_ool(kv_full, (uint32_t)(it & 1));
        issue_scores(0);
        if (nboxes > 1) issue_scores(1);
        for (int bi = 0; bi < nboxes; ++bi, ++gb) {
          const int sl = gb & 3, g = bi & 1;
          wait_ool(p_full(g), (uint32_t)(cnt[g] & 1));
          ++cnt[g];
          if (m == 0) KV_TRACE(p.debug, L, 2, trn, 10, it, bi);
          if (bi == 0 && it > 0) wait_ool(acc_empty, (uint32_t)((it - 1) & 1));
          ptx::tc_fence_after();
          if (m == 0) KV_TRACE(p.debug, L, 2, trn, 13, it, bi);
          const uint64_t db = dBm + (uint64_t)(sl * 256);
          const uint32_t a = t_score + g * 32;     // bf16 pairs written in place by the element-wise warps
          ptx::umma_bf16_ts(t_acc, a, db, idesc_acc, bi > 0 ? 1u : 0u);
          if (bi < nboxes - 1 || tail_ksteps > 1) ptx::umma_bf16_ts(t_acc, a + 8, db + 128, idesc_acc, 1u);
          if (m == 0) KV_TRACE(p.debug, L, 2, trn, 14, it, bi);
          ptx::umma_commit(ld_empty(sl));
          if (bi == nboxes - 1) ptx::umma_commit(acc_full);
          if (m == 0) KV_TRACE(p.debug, L, 2, trn, 11, it, bi);
          if (bi + 2 < nboxes) issue_scores(bi + 2);   // same group's next box: runs after the MMAs above (issue order)
          if (m == 0) KV_TRACE(p.debug, L, 2, trn, 12, it, bi);
        }
      }
    }
  } else if (warp == KV_EW_WARPS + 1) {
    if (lane == 0 && n_items > 0) {
      // ---------------- TMA: K / V tiles (one item ahead) and the Q / dO box ring (4 slots, runs on across items) ----------------
      int b, h, j0, bh;
      item_of(0, b, h, j0, bh);
      ptx::mbar_arrive_expect_tx(kv_full, 2 * 16384);
      ptx::tma_load_3d(base + KV_SM_K, &tm_kv, kv_full, (p.H + h) * HD, j0, b);
      ptx::tma_load_3d(base + KV_SM_V, &tm_kv, kv_full, (2 * p.H + h) * HD, j0, b);
      int gb = 0;
      for (int it = 0; it < n_items; ++it) {
        for (int bi = 0; bi < nboxes; ++bi, ++gb) {
          const int sl = gb & 3;
          if (gb >= 4) wait_ool(ld_empty(sl), (uint32_t)(((gb >> 2) - 1) & 1));
          ptx::mbar_arrive_expect_tx(ld_full(sl), 2 * 4096);
          ptx::tma_load_3d(base + KV_SM_Q + sl * 4096, &tm_q, ld_full(sl), h * HD, bi * 32, b);
          ptx::tma_load_3d(base + KV_SM_DO + sl * 4096, &tm_do, ld_full(sl), h * HD, bi * 32, b);
        }
        if (it + 1 < n_items) {     // the boxes above were requested for item `it`; now K / V of item it + 1 once item it has released them
          item_of(it + 1, b, h, j0, bh);
          wait_ool(kv_free, (uint32_t)(it & 1));
          ptx::mbar_arrive_expect_tx(kv_full, 2 * 16384);
          ptx::tma_load_3d(base + KV_SM_K, &tm_kv, kv_full, (p.H + h) * HD, j0, b);
          ptx::tma_load_3d(base + KV_SM_V, &tm_kv, kv_full, (2 * p.H + h) * HD, j0, b);
        }
      }
    } else if (HAS_BIAS && lane == 1 && n_items > 0) {
      // ---------------- TMA: bias^T ring, [128 keys x 32 queries] fp32 boxes in box order ----------------
      int gb = 0;
      for (int it = 0; it < n_items; ++it) {
        int b, h, j0, bh;
        item_of(it, b, h, j0, bh);
        for (int bi = 0; bi < nboxes; ++bi, ++gb) {
          const int st = gb & 1;
          if (gb >= 2) wait_ool(bias_empty(st), (uint32_t)(((gb >> 1) - 1) & 1));
          ptx::mbar_arrive_expect_tx(bias_full(st), BIAS_STAGE_BYTES);
          ptx::tma_load_3d(base + KV_SM_BIAS + st * BIAS_STAGE_BYTES, &tm_bias, bias_full(st), bi * 32, j0, h);
        }
      }
    }
  } else {
    // ---------------- element-wise warps: one thread per key row; group g = warp >> 2 owns boxes g, g + 2, ... ----------------
    const int quad = warp & 3, g = warp >> 2;
    const int row = quad * 32 + lane;
    const int tid = warp * 32 + lane;                       // 0..255 inside the lane
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t cx = trow + KV_X + g * 32, cy = trow + KV_Y + g * 32;
    int trn = 0;
    // lse / D of the NEXT item are fetched into registers while the current item computes (no global-load latency at the item boundary)
    float nx_lse = INFINITY, nx_d = 0.f;
    auto fetch_row_stats = [&](int it) {
      int b, h, j0, bh;
      item_of(it, b, h, j0, bh);
      nx_lse = tid < p.N ? __ldg(p.lse + (long long)bh * p.N + tid) * LOG2E : INFINITY;
      nx_d = tid < p.N ? __ldg(p.dvec + (long long)bh * p.N + tid) : 0.f;
    };
    if (n_items > 0 && tid < n_pad) fetch_row_stats(0);
    for (int it = 0; it < n_items; ++it) {
      int b, h, j0, bh;
      item_of(it, b, h, j0, bh);
      float* s_lse = reinterpret_cast<float*>(gbase + KV_SM_LSE) + (it & 1) * NMAX;
      float* s_d = reinterpret_cast<float*>(gbase + KV_SM_D) + (it & 1) * NMAX;
      if (tid < n_pad) {
        s_lse[tid] = nx_lse;
        s_d[tid] = nx_d;
      }
      ptx::named_bar_sync(1 + L, KV_EW_WARPS * 32);
      if (it + 1 < n_items && tid < n_pad) fetch_row_stats(it + 1);
      if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 5 + 100 * g, it, 0);
      const int j = j0 + row;
      const bool valid = j < p.N;
      const bool active = j0 + quad * 32 < p.N;
      bf16* ds_base = p.ds_out + ((long long)bh * p.N + (valid ? j : 0)) * p.ld_ds;
      const uint32_t* kt_row = p.keep_t + ((long long)bh * p.N + (valid ? j : 0)) * 8;
      // dV / dK row of this thread (written by the epilogue): touch it now so that the address translation and the L2 line are warm
      // by then (the first store of an item otherwise stalls ~2500 cycles; clock64 trace in tools/micro/kv_trace.py)
      bf16* dst = p.dqkv + ((long long)b * p.N + (valid ? j : 0)) * (3LL * p.H * HD) + (long long)(g == 0 ? 2 : 1) * p.H * HD + h * HD;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(dst));
      for (int bi = g; bi < nboxes; bi += 2) {
        const int c0 = bi * 32;
        const int gb = it * nboxes + bi, st = gb & 1;
        uint32_t kw = 0xffffffffu;
        if (DROP && valid) kw = __ldg(kt_row + bi);
        if (HAS_BIAS) wait_ool(bias_full(st), (uint32_t)((gb >> 1) & 1));
        wait_ool(s_full(g), (uint32_t)(group_count(it, bi) & 1));
        ptx::tc_fence_after();
        if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 1 + 100 * g, it, bi);
        if (active) {
          const uint8_t* bias_row = gbase + KV_SM_BIAS + st * BIAS_STAGE_BYTES + row * 128;
          bwd_step16<DROP, HAS_BIAS>(p, cx, cy, cx, cy, bias_row, 0, row, s_lse + c0, s_d + c0, kw, valid, ds_base + c0);
          if (n_pad - c0 >= 32)
            bwd_step16<DROP, HAS_BIAS>(p, cx + 16, cy + 16, cx + 8, cy + 8, bias_row, 4, row, s_lse + c0 + 16, s_d + c0 + 16, kw >> 16, valid,
                                       ds_base + c0 + 16);
        }
        if (HAS_BIAS) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bias_empty(st));
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(p_full(g));
        if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 2 + 100 * g, it, bi);
      }
      // ---------------- epilogue: dV rows (group 0) / dK rows (group 1) of this key tile ----------------
      wait_ool(acc_full, (uint32_t)(it & 1));
      ptx::tc_fence_after();
      if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 3 + 100 * g, it, 0);
      if (active) {
        const float mul = g == 0 ? 1.0f : p.scale;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          ptx::tmem_ld_x32_sync(trow + (g == 0 ? KV_DV : KV_DK) + half * 32, o);
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(o[e]) * mul;
          if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 6 + 100 * g, it, half);
          if (valid) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
              st_global_32B(dst + half * 32 + 16 * q, pack_bf16x2(v[16 * q], v[16 * q + 1]), pack_bf16x2(v[16 * q + 2], v[16 * q + 3]),
                            pack_bf16x2(v[16 * q + 4], v[16 * q + 5]), pack_bf16x2(v[16 * q + 6], v[16 * q + 7]),
                            pack_bf16x2(v[16 * q + 8], v[16 * q + 9]), pack_bf16x2(v[16 * q + 10], v[16 * q + 11]),
                            pack_bf16x2(v[16 * q + 12], v[16 * q + 13]), pack_bf16x2(v[16 * q + 14], v[16 * q + 15]));
          }
          if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 7 + 100 * g, it, half);
          if (g == 0 && p.dv_bias != nullptr) {   // v_bias gradient: rows past N are exactly zero
            const float c = warp_colsum32(v, lane);
            atomicAdd(p.dv_bias + h * HD + half * 32 + lane, c);
          }
          if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 8 + 100 * g, it, half);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
      if (lane == 0 && quad == 0) KV_TRACE(p.debug, L, g, trn, 4 + 100 * g, it, 0);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == KV_EW_WARPS && L == 0) ptx::tmem_dealloc(tmem_base, 512);
}

// ---------------- dQ = scale * dS K ----------------
// Persistent (one CTA per SM): a TMA thread runs one item ahead through a 2-buffer operand ring, an MMA thread alternates between two
// 64-column TMEM accumulators, four epilogue warps drain accumulator a while the MMAs of the next item fill accumulator 1 - a.
constexpr int DQ_BUF_BYTES = 3 * NMAX * 128;            // 2 chunks (64 queries each) of dS^T [NMAX key rows x 128 B] (MN-major A) + K (MN-major B)
constexpr int DQ_SM_O = 2 * DQ_BUF_BYTES;               // 128 x 128 B staging
constexpr int DQ_SM_BAR = DQ_SM_O + TILE_M * 128;
constexpr int DQ_SMEM = DQ_SM_BAR + 128 + 1024;
constexpr int DQ_THREADS = 192;
static_assert(DQ_BUF_BYTES % 1024 == 0 && DQ_SM_O % 1024 == 0 && DQ_SMEM <= 232448, "dq kernel smem layout");

struct BwdDqParams {
  float* dq_bias;   // [H*64] += or null
  int B, H, N, n_pad, m_tiles, items;
  float scale;
};

__global__ void __launch_bounds__(DQ_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_ds, const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_dq,
                   const BwdDqParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = base + DQ_SM_BAR;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 16u + 8u * s; };
  auto done = [&](int a) { return bar0 + 32u + 8u * a; };
  auto tfree = [&](int a) { return bar0 + 48u + 8u * a; };
  const uint32_t tmem_slot = bar0 + 64u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pad = p.n_pad;
  const int n_items = (int)blockIdx.x < p.items ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); ptx::mbar_init(done(s), 1); ptx::mbar_init(tfree(s), 4); }
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tm_ds); ptx::prefetch_tmap(&tm_kv); ptx::prefetch_tmap(&tm_dq);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (warp == 4) {
    if (lane == 0) {       // ---------------- TMA ----------------
      for (int it = 0; it < n_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int mt = item % p.m_tiles, bh = item / p.m_tiles;
        const int b = bh / p.H, h = bh - b * p.H, m0 = mt * TILE_M, s = it & 1;
        if (it >= 2) ptx::mbar_wait(empty(s), (uint32_t)(((it >> 1) - 1) & 1));
        const uint32_t buf = base + s * DQ_BUF_BYTES;
        ptx::mbar_arrive_expect_tx(full(s), 3 * n_pad * 128);
        ptx::tma_load_3d(buf, &tm_ds, full(s), m0, 0, bh);
        ptx::tma_load_3d(buf + NMAX * 128, &tm_ds, full(s), m0 + 64, 0, bh);
        ptx::tma_load_3d(buf + 2 * NMAX * 128, &tm_kv, full(s), (p.H + h) * HD, 0, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {       // ---------------- MMA ----------------
      const uint32_t idesc = ptx::make_idesc_bf16(TILE_M, HD, true, true);
      for (int it = 0; it < n_items; ++it) {
        const int s = it & 1;
        const uint32_t buf = base + s * DQ_BUF_BYTES;
        const uint64_t da = ptx::make_smem_desc(buf, NMAX * 128, 1024), db = ptx::make_smem_desc(buf + 2 * NMAX * 128, NMAX * 128, 1024);
        ptx::mbar_wait(full(s), (uint32_t)((it >> 1) & 1));
        if (it >= 2) ptx::mbar_wait(tfree(s), (uint32_t)(((it >> 1) - 1) & 1));
        ptx::tc_fence_after();
        for (int kk = 0; kk < n_pad / 16; ++kk) ptx::umma_bf16(tmem_base + s * 64, da + 128 * kk, db + 128 * kk, idesc, kk > 0 ? 1u : 0u);
        ptx::umma_commit(empty(s));
        ptx::umma_commit(done(s));
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int row = warp * 32 + lane;
    for (int it = 0; it < n_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int mt = item % p.m_tiles, bh = item / p.m_tiles;
      const int b = bh / p.H, h = bh - b * p.H, m0 = mt * TILE_M, s = it & 1;
      const bool active = m0 + warp * 32 < p.N;
      ptx::mbar_wait(done(s), (uint32_t)((it >> 1) & 1));
      ptx::tc_fence_after();
      float v[64];
      if (active) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          ptx::tmem_ld_x32_sync(tmem_base + ((uint32_t)(warp * 32) << 16) + s * 64 + half * 32, o);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[half * 32 + e] = __uint_as_float(o[e]) * p.scale;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tfree(s));
      if (it > 0) {                         // the TMA store of the previous item must have read the staging tile
        if (warp == 0 && lane == 0) ptx::bulk_wait_read0();
        ptx::named_bar_sync(1, 128);
      }
      if (active) {
        uint8_t* orow = gbase + DQ_SM_O + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(orow + ((q ^ (row & 7)) << 4)) = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                                                                 pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        if (p.dq_bias != nullptr) {   // q_bias gradient: query rows past N are exactly zero (their dS columns are zero)
          float c0, c1;
          warp_colsum64(v, lane, c0, c1);
          atomicAdd(p.dq_bias + h * HD + 2 * lane, c0);
          atomicAdd(p.dq_bias + h * HD + 2 * lane + 1, c1);
        }
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(2, 128);
      if (warp == 0 && lane == 0) {
        ptx::tma_store_3d(&tm_dq, base + DQ_SM_O, h * HD, m0, b);
        ptx::bulk_commit();
      }
    }
    if (warp == 0 && lane == 0) ptx::bulk_wait0();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, 128);
}

template <bool DROP, bool HAS_BIAS>
cudaError_t launch_bwd_kv(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tdo, const CUtensorMap& tb, const BwdKvParams& p,
                          cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kv_kernel<DROP, HAS_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int ctas = min(b200vit_num_sms(), (p.items + KV_LANES - 1) / KV_LANES);
  attn_bwd_kv_kernel<DROP, HAS_BIAS><<<ctas, KV_THREADS, KV_SMEM, stream>>>(tq, tkv, tdo, tb, p);
  return cudaGetLastError();
}

}  // namespace

int b200vit_relbias_grad_launch(const void* ds_work, int B, int H, int N, int ld_ds, const int32_t* rel_index, float* dtable, void* stream);

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

#ifdef B200VIT_KV_TRACE
extern "C" int b200vit_debug_kv_trace(long long* host_out, int max_events) {   // tools only (not in the public header)
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_out, g_kv_trace, sizeof(long long) * 3 * 512 * 4);
  return 3 * 512;
}
#endif

extern "C" size_t b200vit_attn_bwd_workspace_bytes(int32_t B, int32_t H, int32_t N) {
  const size_t n_pad = (size_t)(N + 15) / 16 * 16;
  const size_t rows = (size_t)B * H * N;
  return align256(rows * n_pad * sizeof(bf16)) + align256(rows * sizeof(float)) + align256(rows * 8 * sizeof(uint32_t));
}

extern "C" int b200vit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias_t, int64_t ld_bias,
                                const uint8_t* keep_bits, void* work, int32_t ld_ds, const int32_t* rel_index, float* dtable,
                                float* dq_bias, float* dv_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop,
                                void* dqkv, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv && out && dout && lse && dqkv && work, "attn_bwd: null pointer (the workspace of b200vit_attn_bwd_workspace_bytes is required)");
  B200_CHECK_ARG(B > 0 && H > 0, "attn_bwd: bad B=%d H=%d", B, H);
  B200_CHECK_ARG(head_dim == HD, "attn_bwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "attn_bwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "attn_bwd: bad p_drop");
  B200_CHECK_ARG(p_drop == 0.f || keep_bits != nullptr, "attn_bwd: dropout needs keep_bits from the forward");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(ld_ds == n_pad, "attn_bwd: ld_ds must be N rounded up to 16 (%d)", n_pad);
  B200_CHECK_ARG(bias_t == nullptr || (ld_bias >= n_pad && ld_bias % 4 == 0 && (reinterpret_cast<uintptr_t>(bias_t) & 15) == 0),
                 "attn_bwd: bias_t must be the transposed padded layout of b200vit_rel_pos_bias ([H,N,ld], ld %% 4 == 0, >= %d, 16-byte aligned)", n_pad);
  B200_CHECK_ARG(dtable == nullptr || rel_index != nullptr, "attn_bwd: dtable needs rel_index");
  const uintptr_t addrs[] = {(uintptr_t)qkv, (uintptr_t)out, (uintptr_t)dout, (uintptr_t)dqkv, (uintptr_t)work, (uintptr_t)keep_bits};
  for (uintptr_t a : addrs) B200_CHECK_ARG((a & 15) == 0, "attn_bwd: tensors must be 16-byte aligned");
  const size_t rows = (size_t)B * H * N;
  bf16* ds = static_cast<bf16*>(work);
  float* dvec = reinterpret_cast<float*>(static_cast<uint8_t*>(work) + align256(rows * n_pad * sizeof(bf16)));
  uint32_t* keep_t = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(dvec) + align256(rows * sizeof(float)));
  const int sms = b200vit_num_sms();
  const bool drop = p_drop > 0.f;

  attn_bwd_prep_kernel<<<sms * 8, 256, 0, stream>>>(static_cast<const bf16*>(out), static_cast<const bf16*>(dout), dvec, B, H, N);
  B200_CHECK_LAUNCH("attn_bwd_prep");
  if (drop) {
    keep_transpose_kernel<<<B * H, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(keep_bits), keep_t, N);
    B200_CHECK_LAUNCH("keep_transpose");
  }

  BwdKvParams p;
  p.lse = lse; p.dvec = dvec; p.keep_t = keep_t; p.ds_out = ds; p.ld_ds = ld_ds; p.dqkv = static_cast<bf16*>(dqkv); p.dv_bias = dv_bias;
  p.B = B; p.H = H; p.N = N; p.n_pad = n_pad; p.k_tiles = (N + TILE_M - 1) / TILE_M; p.items = B * H * p.k_tiles;
  p.scale = scale; p.sl2 = scale * LOG2E; p.inv_keep = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  p.debug = 0;
  const uint64_t row = 3ull * H * HD, orow = (uint64_t)H * HD;
  CUtensorMap tq, tkv, tdo, tb, tds, tkfull, tdq;
  int rc;
  if ((rc = make_tmap3(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, row, N, B, row, row * N, HD, 32))) return rc;
  if ((rc = make_tmap3(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, row, N, B, row, row * N, HD, TILE_M))) return rc;
  if ((rc = make_tmap3(&tdo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout, orow, N, B, orow, orow * N, HD, 32))) return rc;
  if (bias_t != nullptr) {
    if ((rc = make_tmap3(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, bias_t, ld_bias, N, H, ld_bias, (uint64_t)ld_bias * N, 32, TILE_M))) return rc;
  } else {
    tb = tq;
  }
  cudaError_t e;
  if (drop) e = bias_t != nullptr ? launch_bwd_kv<true, true>(tq, tkv, tdo, tb, p, stream) : launch_bwd_kv<true, false>(tq, tkv, tdo, tb, p, stream);
  else e = bias_t != nullptr ? launch_bwd_kv<false, true>(tq, tkv, tdo, tb, p, stream) : launch_bwd_kv<false, false>(tq, tkv, tdo, tb, p, stream);
  if (e != cudaSuccess) { b200vit_set_error("attn_bwd (kv): launch failed: %s", cudaGetErrorString(e)); return (int)e; }

  // dQ = scale * dS K
  if ((rc = make_tmap3(&tds, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ds, ld_ds, N, (uint64_t)B * H, ld_ds, (uint64_t)ld_ds * N, 64, n_pad))) return rc;
  if ((rc = make_tmap3(&tkfull, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, row, N, B, row, row * N, HD, n_pad))) return rc;
  if ((rc = make_tmap3(&tdq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dqkv, row, N, B, row, row * N, HD, TILE_M))) return rc;
  BwdDqParams dp;
  dp.dq_bias = dq_bias; dp.B = B; dp.H = H; dp.N = N; dp.n_pad = n_pad; dp.m_tiles = (N + TILE_M - 1) / TILE_M; dp.scale = scale;
  dp.items = B * H * dp.m_tiles;
  static bool dq_configured = false;
  if (!dq_configured) {
    e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("attn_bwd (dq): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    dq_configured = true;
  }
  attn_bwd_dq_kernel<<<dp.items < sms ? dp.items : sms, DQ_THREADS, DQ_SMEM, stream>>>(tds, tkfull, tdq, dp);
  B200_CHECK_LAUNCH("attn_bwd_dq");
  if (dtable != nullptr) return b200vit_relbias_grad_launch(ds, B, H, N, ld_ds, rel_index, dtable, stream);
  return 0;
}

extern "C" int b200vit_attn_fwd(const void* qkv, const float* bias, int64_t ld_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim,
                                float scale, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id, const uint8_t* keep_in, void* out,
                                float* lse, uint8_t* keep_bits, int32_t keep_ready, const uint16_t* bias_idx16, const float* bias_tab, int32_t nbins,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(qkv != nullptr && out != nullptr, "attn_fwd: null pointer");
  const bool indexed = bias_idx16 != nullptr;
  B200_CHECK_ARG(!indexed || (bias_tab != nullptr && nbins > 0 && nbins < B200VIT_ATTN_TAB_MAX && (reinterpret_cast<uintptr_t>(bias_idx16) & 15) == 0),
                 "attn_fwd: the indexed bias needs bias_tab, 0 < nbins < %d and a 16-byte aligned index tile", B200VIT_ATTN_TAB_MAX);
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
  p.lse = lse; p.keep_bits = keep_bits; p.keep_in = keep_in; p.B = B; p.H = H; p.N = N; p.n_pad = n_pad; p.m_tiles = (N + TILE_M - 1) / TILE_M; p.items = B * H * p.m_tiles;
  p.sl2 = scale * LOG2E; p.inv_keep = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  p.thresh = (uint32_t)(p_drop * 65536.0f + 0.5f); p.seed = seed; p.seed_dev = seed_dev; p.stream_id = stream_id;
  p.bias_idx16 = bias_idx16; p.bias_tab = bias_tab; p.tab_pitch = indexed ? ((nbins + 1 + 3) & ~3) : 0;
  const uint64_t row = 3ull * H * HD;
  CUtensorMap tq, tkv, tb, to;
  int rc;
  if ((rc = make_tmap3(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, row, N, B, row, row * N, HD, TILE_M))) return rc;
  if ((rc = make_tmap3(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, row, N, B, row, row * N, HD, n_pad))) return rc;
  if ((rc = make_tmap3(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, (uint64_t)H * HD, N, B, (uint64_t)H * HD, (uint64_t)H * HD * N, HD, TILE_M))) return rc;
  if (bias != nullptr && !indexed) {
    if ((rc = make_tmap3(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, bias, ld_bias, N, H, ld_bias, (uint64_t)ld_bias * N, 32, TILE_M))) return rc;
  } else {
    tb = tq;
  }
  cudaError_t e;
  if (p_drop > 0.f && !keep_ready) {
    if ((rc = b200vit_keep_bits_launch(keep_bits, B * H, N, p_drop, seed, seed_dev, stream_id, keep_in, stream_))) return rc;
  }
  const int mode = indexed ? 2 : (bias != nullptr ? 1 : 0);
  if (p_drop > 0.f) e = mode == 2 ? launch_fwd100<true, 2>(tq, tkv, tb, to, p, stream) : mode == 1 ? launch_fwd100<true, 1>(tq, tkv, tb, to, p, stream) : launch_fwd100<true, 0>(tq, tkv, tb, to, p, stream);
  else e = mode == 2 ? launch_fwd100<false, 2>(tq, tkv, tb, to, p, stream) : mode == 1 ? launch_fwd100<false, 1>(tq, tkv, tb, to, p, stream) : launch_fwd100<false, 0>(tq, tkv, tb, to, p, stream);
  if (e != cudaSuccess) { b200vit_set_error("attn_fwd: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
