// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> shared memory (SWIZZLE_128B) -> tcgen05.mma.cta_group::2
// (a CTA pair on one TPC computes a 256x256 tile: 256x256x16 UMMA, each CTA stages 128 rows of A and 128 rows of B, so the
// L2->SMEM operand traffic per flop is 2/3 of the single-CTA 128x256 tile), fp32 accumulators in TMEM (double-buffered),
// tcgen05.ld epilogue transposed through shared memory for coalesced I/O with fused bias / GELU / layer-scale + drop-path +
// residual / dGELU / ELU+1 / fp32 / split-K atomic accumulation.
//
//   D[M,N] = A[M,K] * B[N,K]^T
//
// Replaces every nn.Linear / F.linear of the reference hot path and their autograd backward:
//   QKV, proj (modeling_finetune.py:149,186), fc1/fc2 (modeling_finetune.py:76-81), lm_head (modeling_cyclical.py:219-225),
//   patch-embed conv-as-GEMM (modeling_finetune.py:324), head (modeling_finetune.py:522).
//
// Warp roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle (together one
// warpgroup that shrinks to 40 registers), warps 4..19 = epilogue (four warpgroups grown to 104 registers; warp w owns TMEM lanes
// 32*(w%4).. and the 64-column quarter (w-4)/4 of the 256-column accumulator) — 4 epilogue warps per scheduler hide the
// TMEM/LDS/MUFU latencies that 2 per scheduler could not.
#include <mutex>

#include "../../include/b200vit.h"
#include "ptx_sm100.cuh"

namespace {

constexpr int BLOCK_M = 128;   // rows per CTA; a CTA pair (cta_group::2) owns a 256 x 256 output tile
constexpr int PAIR_M = 256;
constexpr int BLOCK_N = 256;   // tile width; each CTA of the pair stages HALF_N rows of B
constexpr int HALF_N = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;
// Two launch shapes, chosen per epilogue (measured, profiles/r1_gemm_shapes.log):
//   math-heavy epilogues (GELU, dGELU): 16 epilogue warps (4 per scheduler, 64 columns each), 5-stage ring, setmaxnreg 40 / 104
//   I/O epilogues (bf16, residual, fp32, split-K atomics, ELU+1): 8 epilogue warps (128 columns each), 6-stage ring
template <int MODE>
struct Cfg {
  static constexpr bool kWide = (MODE == B200VIT_EPI_GELU || MODE == B200VIT_EPI_DGELU || MODE == B200VIT_EPI_ELU1);   // dGELU: 16 warps hide the aux-load latency (8: 190 us, 16: 153 us)
  static constexpr int STAGES = kWide ? 5 : 6;
  static constexpr int NUM_EPI_WARPS = kWide ? 16 : 8;
  static constexpr int EPI_COLS = BLOCK_N / (NUM_EPI_WARPS / 4);
  static constexpr int NUM_THREADS = (4 + NUM_EPI_WARPS) * 32;
  static constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 32 * 32 * 4;
  static constexpr int SMEM_BYTES = STAGES * (BLOCK_M * BLOCK_K * 2 + HALF_N * BLOCK_K * 2) + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
  static constexpr int REGS_PRODUCER = 40;    // setmaxnreg budget of warpgroup 0 (TMA / MMA / TMEM-alloc / idle warps)
  static constexpr int REGS_EPILOGUE = kWide ? 104 : 208;   // 128*40 + 512*104 <= 640*96 ; 128*40 + 256*208 <= 384*168
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};
constexpr int FIRST_EPI_WARP = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_STAGE_BYTES = HALF_N * BLOCK_K * 2;   // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512

struct EpiParams {
  int mode;
  const float* bias;
  const float* colscale;
  const float* rowscale;
  int rows_per_scale;
  const float* residual;
  long long ld_residual;
  const bf16* aux;
  long long ld_aux;
  float* out_f32;
  long long ld_f32;
  bf16* out_bf16;
  long long ld_bf16;
  bf16* out2_bf16;
  long long ld2_bf16;
  float alpha;
  float* colsum;
  int debug_skip_io;   // profiling only: drain TMEM but skip the epilogue's global loads/stores
};

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, num_kb, split_k, kb_per_split;
  EpiParams epi;
};

__device__ __forceinline__ void store_bf16x4(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// Per-row operand of the epilogue that lives in global memory (residual fp32x4 / aux bf16x4), prefetched one 32-column
// chunk ahead so that its latency never sits on the TMEM-drain critical path.
struct EpiPrefetch {
  float4 v;   // residual (RESIDUAL) or the 4 aux values converted to fp32 (DGELU)
};

template <int MODE>
__device__ __forceinline__ EpiPrefetch epi_prefetch(const EpiParams& e, int m, int n, bool ok) {
  EpiPrefetch pf;
  pf.v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!ok) return pf;
  if (MODE == B200VIT_EPI_RESIDUAL) {
    pf.v = *reinterpret_cast<const float4*>(e.residual + (long long)m * e.ld_residual + n);
  } else if (MODE == B200VIT_EPI_DGELU) {
    const uint2 a = *reinterpret_cast<const uint2*>(e.aux + (long long)m * e.ld_aux + n);
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y);
    pf.v = make_float4(a0.x, a0.y, a1.x, a1.y);
  }
  return pf;
}

// Fused epilogue on 4 consecutive accumulator columns [n, n+4) of row m (coalesced layout: 8 lanes cover 32 columns of a row,
// a warp instruction covers 4 rows x 128 B). bias4 / cs4 / rs / pf were loaded ahead of time. Returns the value written
// (for the fused column sums).
template <int MODE>
__device__ __forceinline__ float4 epilogue4(const EpiParams& e, int m, int n, float4 v, const float4& bias4, const float4& cs4, float rs,
                                            const EpiPrefetch& pf) {
  if (MODE == B200VIT_EPI_F32_ATOMIC) {
    ptx::red_add_v4(e.out_f32 + (long long)m * e.ld_f32 + n, v.x * e.alpha, v.y * e.alpha, v.z * e.alpha, v.w * e.alpha);
    return v;
  }
  v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
  const bool e_has_out2 = e.out2_bf16 != nullptr;
  switch (MODE) {
    case B200VIT_EPI_BF16: {
      v.x *= cs4.x; v.y *= cs4.y; v.z *= cs4.z; v.w *= cs4.w;
      store_bf16x4(e.out_bf16 + (long long)m * e.ld_bf16 + n, v);
      break;
    }
    case B200VIT_EPI_GELU: {
      // out2 (training forward only) = gelu'(t) = Phi(t) + t phi(t): the backward GEMM epilogue then is one multiply per element
      // instead of re-evaluating the polynomial + exp (the dGELU GEMM was epilogue-bound at 0.6 PFLOP/s).
      float o[4], d[4];
      const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float q = phi_minus_half(t[i]);
        o[i] = fmaf(t[i], q, 0.5f * t[i]);                                            // t * Phi(t)
        d[i] = 0.f;
        if (e_has_out2) {
          const float ex = mufu_ex2(t[i] * (t[i] * -0.72134752044448170f));            // exp(-t^2/2)
          d[i] = fmaf(t[i] * 0.39894228040143268f, ex, 0.5f + q);
        }
      }
      if (e_has_out2) store_bf16x4(e.out2_bf16 + (long long)m * e.ld2_bf16 + n, make_float4(d[0], d[1], d[2], d[3]));
      v = make_float4(o[0], o[1], o[2], o[3]);
      store_bf16x4(e.out_bf16 + (long long)m * e.ld_bf16 + n, v);
      break;
    }
    case B200VIT_EPI_RESIDUAL: {
      if (e.out2_bf16 != nullptr) store_bf16x4(e.out2_bf16 + (long long)m * e.ld2_bf16 + n, v);
      v = make_float4(fmaf(rs * cs4.x, v.x, pf.v.x), fmaf(rs * cs4.y, v.y, pf.v.y), fmaf(rs * cs4.z, v.z, pf.v.z), fmaf(rs * cs4.w, v.w, pf.v.w));
      *reinterpret_cast<float4*>(e.out_f32 + (long long)m * e.ld_f32 + n) = v;
      break;
    }
    case B200VIT_EPI_DGELU: {
      v = make_float4(v.x * pf.v.x, v.y * pf.v.y, v.z * pf.v.z, v.w * pf.v.w);     // aux = gelu'(pre) saved by the forward epilogue
      store_bf16x4(e.out_bf16 + (long long)m * e.ld_bf16 + n, v);
      break;
    }
    case B200VIT_EPI_ELU1: {
      v = make_float4(v.x > 0.f ? v.x + 1.0f : __expf(v.x), v.y > 0.f ? v.y + 1.0f : __expf(v.y), v.z > 0.f ? v.z + 1.0f : __expf(v.z),
                      v.w > 0.f ? v.w + 1.0f : __expf(v.w));   // elu(x) + 1
      store_bf16x4(e.out_bf16 + (long long)m * e.ld_bf16 + n, v);
      break;
    }
    case B200VIT_EPI_F32:
    default:
      *reinterpret_cast<float4*>(e.out_f32 + (long long)m * e.ld_f32 + n) = v;
      break;
  }
  return v;
}

template <bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__(Cfg<MODE>::NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  constexpr int STAGES = Cfg<MODE>::STAGES;
  constexpr int NUM_EPI_WARPS = Cfg<MODE>::NUM_EPI_WARPS;
  constexpr int EPI_COLS = Cfg<MODE>::EPI_COLS;
  constexpr int REGS_PRODUCER = Cfg<MODE>::REGS_PRODUCER;
  constexpr int REGS_EPILOGUE = Cfg<MODE>::REGS_EPILOGUE;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned stage bases
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  const uint32_t stage_base = bar_base + 256u;    // epilogue transpose staging: NUM_EPI_WARPS x 32 x 32 fp32
  auto smem_a = [&](int s) { return smem_base + s * STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * STAGE_BYTES + A_STAGE_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();   // 0 = leader of the pair (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);      // leader: its own arrive.expect_tx; bytes from BOTH CTAs' TMA loads
      ptx::mbar_init(empty_bar(s), 1);     // per CTA: multicast tcgen05.commit of the leader
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);                     // per CTA: multicast commit
      ptx::mbar_init(tempty_bar(a), 2 * NUM_EPI_WARPS);    // leader: epilogue warps of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_units = num_tiles * p.split_k;

  if (warp < FIRST_EPI_WARP) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
  if (warp == 0) {
    // ===================== TMA producer (one lane per CTA; both CTAs credit the leader's full barrier) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = pair; u < num_units; u += num_pairs) {
        const int tile = u / p.split_k, split = u - tile * p.split_k;
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int m0 = m_blk * PAIR_M + (int)cta_rank * BLOCK_M;
        const int n0 = n_blk * BLOCK_N + (int)cta_rank * HALF_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              ptx::tma_load_2d_2sm(smem_a(stage) + c * (BLOCK_K * 128), &tmap_a, full_bar(stage), m0 + c * 64, kb * BLOCK_K);
          } else {
            ptx::tma_load_2d_2sm(smem_a(stage), &tmap_a, full_bar(stage), kb * BLOCK_K, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < HALF_N / 64; ++c)
              ptx::tma_load_2d_2sm(smem_b(stage) + c * (BLOCK_K * 128), &tmap_b, full_bar(stage), n0 + c * 64, kb * BLOCK_K);
          } else {
            ptx::tma_load_2d_2sm(smem_b(stage), &tmap_b, full_bar(stage), kb * BLOCK_K, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane of the LEADER CTA) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(PAIR_M, BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = pair; u < num_units; u += num_pairs) {
        const int tile = u / p.split_k, split = u - tile * p.split_k;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: 32 B per UMMA_K inside the 128-byte swizzle row; MN-major: 16 k-rows x 128 B = 2048 B
            const uint64_t adesc = A_MN ? ptx::make_smem_desc(smem_a(stage) + k * 2048, BLOCK_K * 128, 1024)
                                        : ptx::make_smem_desc(smem_a(stage) + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? ptx::make_smem_desc(smem_b(stage) + k * 2048, BLOCK_K * 128, 1024)
                                        : ptx::make_smem_desc(smem_b(stage) + k * 32, 16, 1024);
            ptx::umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit_2sm(empty_bar(stage), 3);  // frees this smem slot in BOTH CTAs once the MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit_2sm(tfull_bar(acc), 3);      // accumulators complete -> both CTAs' epilogues
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= FIRST_EPI_WARP) {
    // ===================== epilogue warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPILOGUE));
    const int ew = warp - FIRST_EPI_WARP;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;                // 64-column quarter of the accumulator
    float* stage = reinterpret_cast<float*>(smem_raw + (stage_base - ptx::smem_u32(smem_raw))) + ew * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int cg = lane & 7, sub = lane >> 3;      // coalesced view: 4-column group and row-within-4 of this lane
    constexpr int CHUNKS = EPI_COLS / 32;          // 2 or 4 chunks of 32 columns per warp
    const bool io = !p.epi.debug_skip_io;
    // ---- everything that does not depend on the accumulator is fetched ahead of time: for chunk c + 1 while chunk c is being processed, and
    // for chunk 0 of the NEXT tile while the last chunk of this one is (when the epilogue paces the kernel — the K = 768 shapes — the
    // accumulator is already waiting, and a load issued at the top of a tile would sit on the critical path once per tile)
    auto unit_bases = [&](int u, int& m_base, int& n_base) {
      const int tile = u / p.split_k;
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      m_base = m_blk * PAIR_M + (int)cta_rank * BLOCK_M + quarter * 32;
      n_base = n_blk * BLOCK_N + half * EPI_COLS + cg * 4;
    };
    auto load_cols = [&](int n, float4& b4, float4& c4) {
      b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      c4 = make_float4(1.f, 1.f, 1.f, 1.f);
      if (MODE != B200VIT_EPI_F32_ATOMIC && n < p.N && io) {
        if (p.epi.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.epi.bias + n));
        if ((MODE == B200VIT_EPI_BF16 || MODE == B200VIT_EPI_RESIDUAL) && p.epi.colscale != nullptr)
          c4 = __ldg(reinterpret_cast<const float4*>(p.epi.colscale + n));
      }
    };
    auto row_scale = [&](int m) {
      return (MODE == B200VIT_EPI_RESIDUAL && p.epi.rowscale != nullptr && m < p.M) ? __ldg(p.epi.rowscale + m / p.epi.rows_per_scale) : 1.0f;
    };
    float4 bias4, cs4;
    float rs[8];
    EpiPrefetch pf[8];
    if (pair < num_units) {
      int m_base, n_base;
      unit_bases(pair, m_base, n_base);
      load_cols(n_base, bias4, cs4);
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int m = m_base + rr * 4 + sub;
        rs[rr] = row_scale(m);
        pf[rr] = epi_prefetch<MODE>(p.epi, m, n_base, io && m < p.M && n_base < p.N);
      }
    }
    for (int u = pair; u < num_units; u += num_pairs) {
      int m_base, n_base;
      unit_bases(u, m_base, n_base);
      const bool has_next = u + num_pairs < num_units;
      int m_next = 0, n_next = 0;
      if (has_next) unit_bases(u + num_pairs, m_next, n_next);
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N + half * EPI_COLS;
      bool wrapped = false;                  // chunk 0 of the next tile has been requested
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int n = n_base + c * 32;
        if (n - cg * 4 >= p.N) break;      // warp-uniform: this 32-column chunk lies outside the matrix
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + c * 32, r);
        ptx::tmem_ld_wait();
        // transpose through shared memory: lane = row on the TMEM side, lane = (row % 4, 4-column group) on the global side
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(stage + lane * 32 + ((g ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
        __syncwarp();
        float4 bias_n, cs_n;
        EpiPrefetch nxt[8];
        float rs_n[8];
        const bool more = c + 1 < CHUNKS && n + 32 - cg * 4 < p.N;       // another chunk of this tile follows
        const bool wrap = !more && has_next;                               // otherwise: chunk 0 of the next tile
        wrapped = wrapped || wrap;
        const int n_pre = more ? n + 32 : n_next, m_pre = more ? m_base : m_next;
        load_cols((more || wrap) ? n_pre : p.N, bias_n, cs_n);
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int m = m_pre + rr * 4 + sub;
          nxt[rr] = epi_prefetch<MODE>(p.epi, m, n_pre, (more || wrap) && io && m < p.M && n_pre < p.N);
          rs_n[rr] = wrap ? row_scale(m) : rs[rr];
        }
        const bool col_ok = n < p.N;       // N % 8 == 0: a 4-column group is entirely valid or not
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int row = rr * 4 + sub;
          const float4 v = *reinterpret_cast<const float4*>(stage + row * 32 + ((cg ^ (row & 7)) << 2));
          const int m = m_base + row;
          if (col_ok && m < p.M && io) {
            const float4 o = epilogue4<MODE>(p.epi, m, n, v, bias4, cs4, rs[rr], pf[rr]);
            csum.x += o.x; csum.y += o.y; csum.z += o.z; csum.w += o.w;
          }
        }
        if (p.epi.colsum != nullptr) {     // fused bias gradient: column sums of the values written by this tile
#pragma unroll
          for (int o = 8; o < 32; o <<= 1) {
            csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o); csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
            csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o); csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
          }
          if (sub == 0 && col_ok) ptx::red_add_v4(p.epi.colsum + n, csum.x, csum.y, csum.z, csum.w);
        }
        bias4 = bias_n;
        cs4 = cs_n;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) { pf[rr] = nxt[rr]; rs[rr] = rs_n[rr]; }
        __syncwarp();
      }
      if (has_next && !wrapped) {            // this warp's columns lie outside the matrix in this tile: nothing was processed, fetch directly
        load_cols(n_next, bias4, cs4);
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int m = m_next + rr * 4 + sub;
          rs[rr] = row_scale(m);
          pf[rr] = epi_prefetch<MODE>(p.epi, m, n_next, io && m < p.M && n_next < p.N);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster_relaxed(tempty_bar(acc), 0);   // the leader's MMA thread waits for both CTAs' epilogues
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();     // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows `ld` elements apart; box = box_inner x box_outer.
int make_tmap(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
              uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    b200vit_set_error("gemm: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -2;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(bf16)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200vit_set_error("gemm: cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu ld=%llu", (int)r, ptr,
                      (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
    return -3;
  }
  return 0;
}

// split-K factor for the fp32-atomic (wgrad) epilogue. Cost model in units of one k-block (512 tensor cycles), fitted to
// profiles/r1_gemm_shapes.log: every wave of work units costs its k-blocks plus a small turnaround; only the LAST epilogue is
// exposed (earlier ones overlap the next unit's mainloop through the double-buffered TMEM accumulator).
int pick_split_k(int tiles, int num_kb, int pairs) {
  int best = 1;
  double best_cost = 1e30;
  const int max_split = num_kb / 8 > 0 ? num_kb / 8 : 1;
  for (int s = 1; s <= max_split && s <= 64; ++s) {
    const int kbs = (num_kb + s - 1) / s;
    const int waves = (tiles * s + pairs - 1) / pairs;
    const double cost = (double)waves * (kbs + 4) + 16.0;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

template <bool A_MN, bool B_MN, int MODE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t stream) {
  static bool configured = false;  // benign race: the attribute call is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<A_MN, B_MN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MODE>::SMEM_BYTES);
    if (e != cudaSuccess) {
      b200vit_set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg<MODE>::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg<MODE>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<A_MN, B_MN, MODE>, ta, tb, p);
  if (e != cudaSuccess) {
    b200vit_set_error("gemm_bf16: cluster launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

template <bool A_MN, bool B_MN>
int launch_mode(int mode, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t stream) {
  switch (mode) {
    case B200VIT_EPI_BF16: return launch<A_MN, B_MN, B200VIT_EPI_BF16>(ta, tb, p, grid, stream);
    case B200VIT_EPI_GELU: return launch<A_MN, B_MN, B200VIT_EPI_GELU>(ta, tb, p, grid, stream);
    case B200VIT_EPI_RESIDUAL: return launch<A_MN, B_MN, B200VIT_EPI_RESIDUAL>(ta, tb, p, grid, stream);
    case B200VIT_EPI_DGELU: return launch<A_MN, B_MN, B200VIT_EPI_DGELU>(ta, tb, p, grid, stream);
    case B200VIT_EPI_F32: return launch<A_MN, B_MN, B200VIT_EPI_F32>(ta, tb, p, grid, stream);
    case B200VIT_EPI_F32_ATOMIC: return launch<A_MN, B_MN, B200VIT_EPI_F32_ATOMIC>(ta, tb, p, grid, stream);
    default: return launch<A_MN, B_MN, B200VIT_EPI_ELU1>(ta, tb, p, grid, stream);
  }
}

}  // namespace

extern "C" int b200vit_gemm_bf16(const b200vit_gemm_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200_CHECK_ARG(d != nullptr, "gemm: null descriptor");
  B200_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm: bad shape M=%d N=%d K=%d", d->M, d->N, d->K);
  B200_CHECK_ARG(d->N % 8 == 0, "gemm: N=%d must be a multiple of 8", d->N);
  B200_CHECK_ARG(d->A != nullptr && d->B != nullptr, "gemm: null operand");
  B200_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (16-byte TMA strides)");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->B) & 15) == 0,
                 "gemm: operands must be 16-byte aligned");
  B200_CHECK_ARG(d->lda >= (d->a_mn_major ? d->M : d->K) && d->ldb >= (d->b_mn_major ? d->N : d->K), "gemm: leading dimension too small");
  const int mode = d->epilogue;
  B200_CHECK_ARG(mode >= 0 && mode <= B200VIT_EPI_ELU1, "gemm: unknown epilogue %d", mode);
  const bool wants_bf16 = mode == B200VIT_EPI_BF16 || mode == B200VIT_EPI_GELU || mode == B200VIT_EPI_DGELU || mode == B200VIT_EPI_ELU1;
  if (wants_bf16) B200_CHECK_ARG(d->out_bf16 != nullptr && d->ld_bf16 % 8 == 0 && d->ld_bf16 >= d->N, "gemm: bad out_bf16/ld_bf16");
  if (!wants_bf16) B200_CHECK_ARG(d->out_f32 != nullptr && d->ld_f32 % 4 == 0 && d->ld_f32 >= d->N, "gemm: bad out_f32/ld_f32");
  if (mode == B200VIT_EPI_RESIDUAL)
    B200_CHECK_ARG(d->residual != nullptr && d->ld_residual % 4 == 0 && (d->rowscale == nullptr || d->rows_per_scale > 0),
                   "gemm: RESIDUAL epilogue needs residual (+ rows_per_scale with rowscale)");
  if (mode == B200VIT_EPI_DGELU) B200_CHECK_ARG(d->aux != nullptr && d->ld_aux % 8 == 0, "gemm: DGELU epilogue needs aux");
  if (d->out2_bf16 != nullptr) B200_CHECK_ARG(d->ld2_bf16 % 8 == 0 && d->ld2_bf16 >= d->N, "gemm: bad ld2_bf16");
  B200_CHECK_ARG(d->split_k <= 1 || mode == B200VIT_EPI_F32_ATOMIC, "gemm: split_k requires the F32_ATOMIC epilogue");

  const int sms = b200vit_num_sms();
  B200_CHECK_ARG(sms > 0, "gemm: no CUDA device");

  GemmParams p;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.tiles_m = (d->M + PAIR_M - 1) / PAIR_M;
  p.tiles_n = (d->N + BLOCK_N - 1) / BLOCK_N;
  p.num_kb = (d->K + BLOCK_K - 1) / BLOCK_K;
  int cap = (d->max_ctas > 0 ? (d->max_ctas < sms ? d->max_ctas : sms) : sms) / 2;   // CTA pairs
  if (cap < 1) cap = 1;
  int split = 1;
  if (mode == B200VIT_EPI_F32_ATOMIC) split = d->split_k > 1 ? d->split_k : (d->split_k == 1 ? 1 : pick_split_k(p.tiles_m * p.tiles_n, p.num_kb, cap));
  if (split > p.num_kb) split = p.num_kb;
  p.kb_per_split = (p.num_kb + split - 1) / split;
  p.split_k = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.epi.mode = mode;
  p.epi.bias = d->bias; p.epi.colscale = d->colscale; p.epi.rowscale = d->rowscale;
  p.epi.rows_per_scale = d->rows_per_scale > 0 ? d->rows_per_scale : 1;
  p.epi.residual = d->residual; p.epi.ld_residual = d->ld_residual;
  p.epi.aux = static_cast<const bf16*>(d->aux); p.epi.ld_aux = d->ld_aux;
  p.epi.out_f32 = d->out_f32; p.epi.ld_f32 = d->ld_f32;
  p.epi.out_bf16 = static_cast<bf16*>(d->out_bf16); p.epi.ld_bf16 = d->ld_bf16;
  p.epi.out2_bf16 = static_cast<bf16*>(d->out2_bf16); p.epi.ld2_bf16 = d->ld2_bf16;
  p.epi.alpha = d->alpha == 0.0f ? 1.0f : d->alpha;
  p.epi.colsum = d->colsum;
  p.epi.debug_skip_io = d->debug_flags & 1;

  CUtensorMap ta, tb;
  int rc;
  // K-major: inner = K, rows = M|N, box 64 x BLOCK;   MN-major: inner = M|N, rows = K, box 64 x BLOCK_K
  rc = d->a_mn_major ? make_tmap(&ta, d->A, d->M, d->K, d->lda, 64, BLOCK_K) : make_tmap(&ta, d->A, d->K, d->M, d->lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  rc = d->b_mn_major ? make_tmap(&tb, d->B, d->N, d->K, d->ldb, 64, BLOCK_K) : make_tmap(&tb, d->B, d->K, d->N, d->ldb, BLOCK_K, HALF_N);
  if (rc) return rc;

  const int units = p.tiles_m * p.tiles_n * p.split_k;
  const int grid = 2 * (units < cap ? units : cap);
  if (d->a_mn_major) return d->b_mn_major ? launch_mode<true, true>(mode, ta, tb, p, grid, stream) : launch_mode<true, false>(mode, ta, tb, p, grid, stream);
  return d->b_mn_major ? launch_mode<false, true>(mode, ta, tb, p, grid, stream) : launch_mode<false, false>(mode, ta, tb, p, grid, stream);
}
