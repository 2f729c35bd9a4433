// Fused flash-style multi-head attention for ViT token counts (N <= 208, head_dim 64), forward and backward.
//   S = scale * q k^T + rel_pos_bias[h] ; P = softmax_j(S) ; P~ = dropout(P) ; O = P~ v
// (Attention.forward, modeling_finetune.py:145-188). One CTA per (batch, head); Q/K/V staged in shared memory with
// cp.async, mma.sync m16n8k16 bf16 tensor-core tiles, online softmax with quad-shuffle row reductions, counter-based
// Philox dropout (or an injected keep-mask), relative-position-bias add from the L2-resident [H,N,N] tensor.
// Backward recomputes P from the saved log-sum-exp, owns one 16-key tile per warp (dK/dV in registers), accumulates dQ
// in shared memory, and scatter-adds dS straight into the relative_position_bias_table gradient (732 bins per head).
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

constexpr int HD = 64;            // head dim
constexpr int PITCH = HD + 8;     // smem row pitch in bf16 (144 B: conflict-free ldmatrix)
constexpr int NMAX = 208;         // 13 tiles of 16
constexpr int FWD_WARPS = 7;
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

// Loads rows [0,N) x 64 bf16 of a [*, row_stride] global matrix into smem [NMAX][PITCH]; rows >= N zero-filled up to n_pad.
__device__ __forceinline__ void load_tile_rows(bf16* s, const bf16* g, long long row_stride, int N, int n_pad) {
  for (int i = threadIdx.x; i < n_pad * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    if (r < N) cp_async16(smem_u32(s + r * PITCH + c), g + (long long)r * row_stride + c);
    else *reinterpret_cast<uint4*>(s + r * PITCH + c) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// keep-bit of element (i, j) of head bh: 16-bit lanes of Philox4x32-10; one call covers the 8 values a thread owns in
// four consecutive 8-key tiles for one row (see DESIGN.md "dropout stream").
__device__ __forceinline__ Philox4 dropout_group(uint64_t seed, uint32_t stream, uint32_t bh, uint32_t i, uint32_t quad, uint32_t group) {
  return philox4x32_10(bh, i, quad * 8u + group, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ uint32_t dropout_u16(const Philox4& r, int idx /*0..7*/) {
  const uint32_t w = idx < 4 ? (idx < 2 ? r.x : r.y) : (idx < 6 ? r.z : r.w);
  return (idx & 1) ? (w >> 16) : (w & 0xffffu);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct AttnFwdParams {
  const bf16* qkv;      // [B, N, 3, H, 64]
  const float* bias;    // [H, N, ld_bias] or null
  long long ld_bias;
  bf16* out;            // [B, N, H*64]
  float* lse;           // [B, H, N]  natural-log log-sum-exp of the biased, scaled scores
  uint8_t* keep_bits;   // [B, H, N, 32] packed keep mask (bit j%8 of byte j/8), written when p_drop > 0
  const uint8_t* keep_in;  // optional injected keep mask [B, H, N, N] (0/1); else Philox
  int B, H, N;
  float scale, p_drop;
  uint64_t seed;
  uint32_t stream_id;
};

__global__ void __launch_bounds__(FWD_WARPS * 32, 2) attn_fwd_kernel(const AttnFwdParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sK = sQ + NMAX * PITCH;
  bf16* sV = sK + NMAX * PITCH;
  const int bh = blockIdx.x;
  const int b = bh / p.H, h = bh - b * p.H;
  const int N = p.N;
  const int ntile = (N + 15) >> 4;
  const int n_pad = ntile * 16;
  const long long row_stride = 3LL * p.H * HD;
  const bf16* gq = p.qkv + (long long)b * N * row_stride + h * HD;
  load_tile_rows(sQ, gq, row_stride, N, n_pad);
  load_tile_rows(sK, gq + p.H * HD, row_stride, N, n_pad);
  load_tile_rows(sV, gq + 2 * p.H * HD, row_stride, N, n_pad);
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane & 3, qrow = lane >> 2;
  const bool drop = p.p_drop > 0.f;
  const float inv_keep = drop ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const uint32_t thresh = (uint32_t)(p.p_drop * 65536.0f + 0.5f);
  const float sl2 = p.scale * LOG2E;

  for (int mt = warp; mt < ntile; mt += FWD_WARPS) {
    // Q fragments for this 16-row tile: 4 k-steps
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldsm_x4(smem_u32(sQ + (mt * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const int i0 = mt * 16 + qrow, i1 = i0 + 8;

    for (int j0 = 0; j0 < n_pad; j0 += 64) {
      const int nts = min(8, (n_pad - j0) >> 3);  // valid 8-key tiles in this chunk (even)
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np * 2 < nts) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(smem_u32(sK + (j0 + np * 16 + (lane >> 4) * 8 + (lane & 7)) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
            mma16816(s[np * 2], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
            mma16816(s[np * 2 + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
          }
        }
      }
      // scale (log2 domain), bias, key masking, running max
      float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = j0 + nt * 8 + quad * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int jj = j + (e & 1);
          const int ii = (e < 2) ? i0 : i1;
          float v = s[nt][e] * sl2;
          if (p.bias != nullptr && jj < N && ii < N) v += __ldg(p.bias + ((long long)h * N + ii) * p.ld_bias + jj) * LOG2E;
          if (jj >= N || nt >= nts) v = -INFINITY;
          s[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        corr[r] = (mrow[r] == -INFINITY) ? 0.f : exp2f(mrow[r] - mx[r]);
        mrow[r] = mx[r];
        lrow[r] *= corr[r];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
      // probabilities, row sums, dropout
      uint32_t bits[2][2] = {{0u, 0u}, {0u, 0u}};  // [row][word]: bit (nt%4)*8 + quad*2 + lo of word nt/4
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        Philox4 r0, r1;
        if (drop && p.keep_in == nullptr) {
          r0 = dropout_group(p.seed, p.stream_id, bh, i0, quad, (j0 >> 5) + g);
          r1 = dropout_group(p.seed, p.stream_id, bh, i1, quad, (j0 >> 5) + g);
        }
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) {
          const int nt = g * 4 + n4;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = exp2f(s[nt][e] - mrow[e >> 1]);   // -inf -> 0
            lrow[e >> 1] += pv;
            float outv = pv;
            if (drop) {
              const int jj = j0 + nt * 8 + quad * 2 + (e & 1);
              const int ii = (e < 2) ? i0 : i1;
              bool keep;
              if (p.keep_in != nullptr) keep = (jj < N && ii < N) ? p.keep_in[(((long long)bh * N + ii) * N) + jj] != 0 : false;
              else keep = dropout_u16(e < 2 ? r0 : r1, n4 * 2 + (e & 1)) >= thresh;
              if (keep) bits[e >> 1][g] |= 1u << (n4 * 8 + quad * 2 + (e & 1));
              outv = keep ? pv * inv_keep : 0.f;
            }
            s[nt][e] = outv;
          }
        }
      }
      if (drop && p.keep_bits != nullptr) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t w = bits[r][g];
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            const int ii = r == 0 ? i0 : i1;
            if (quad == 0 && ii < N) *reinterpret_cast<uint32_t*>(p.keep_bits + ((long long)bh * N + ii) * 32 + (j0 >> 3) + g * 4) = w;
          }
      }
      // O += P~ V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk * 2 < nts) {
          const uint32_t a0 = pack_bf16x2(s[kk * 2][0], s[kk * 2][1]), a1 = pack_bf16x2(s[kk * 2][2], s[kk * 2][3]);
          const uint32_t a2 = pack_bf16x2(s[kk * 2 + 1][0], s[kk * 2 + 1][1]), a3 = pack_bf16x2(s[kk * 2 + 1][2], s[kk * 2 + 1][3]);
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(smem_u32(sV + (j0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + dp * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
            mma16816(o[dp * 2], a0, a1, a2, a3, b0, b1);
            mma16816(o[dp * 2 + 1], a0, a1, a2, a3, b2, b3);
          }
        }
      }
    }
    // finalise: row sums across the quad, normalise, store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    }
    const float inv0 = 1.0f / lrow[0], inv1 = 1.0f / lrow[1];
    bf16* orow0 = p.out + ((long long)b * N + i0) * (p.H * HD) + h * HD;
    bf16* orow1 = p.out + ((long long)b * N + i1) * (p.H * HD) + h * HD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + quad * 2;
      if (i0 < N) *reinterpret_cast<uint32_t*>(orow0 + c) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
      if (i1 < N) *reinterpret_cast<uint32_t*>(orow1 + c) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
    if (quad == 0 && p.lse != nullptr) {
      if (i0 < N) p.lse[(long long)bh * N + i0] = (mrow[0] + log2f(lrow[0])) / LOG2E;
      if (i1 < N) p.lse[(long long)bh * N + i1] = (mrow[1] + log2f(lrow[1])) / LOG2E;
    }
  }
}

// materialises the Philox keep mask as uint8 [B,H,N,N] (tests: inject the SAME mask into the CPU oracle)
__global__ void dropout_mask_kernel(uint8_t* out, int BH, int N, float p_drop, uint64_t seed, uint32_t stream_id) {
  const uint32_t thresh = (uint32_t)(p_drop * 65536.0f + 0.5f);
  const long long total = (long long)BH * N * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % N);
    const int i = (int)((idx / N) % N);
    const int bh = (int)(idx / ((long long)N * N));
    const int jt = j >> 3, quad = (j & 7) >> 1, lo = j & 1;
    const Philox4 r = dropout_group(seed, stream_id, bh, i, quad, jt >> 2);
    out[idx] = dropout_u16(r, (jt & 3) * 2 + lo) >= thresh ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct AttnBwdParams {
  const bf16* qkv;       // [B, N, 3, H, 64]
  const bf16* out;       // [B, N, H*64]   forward output
  const bf16* dout;      // [B, N, H*64]
  const float* lse;      // [B, H, N]
  const float* bias;     // [H, N, ld_bias] or null
  long long ld_bias;
  const uint8_t* keep_bits;  // [B, H, N, 32] or null (p_drop == 0)
  bf16* ds_out;          // [B, H, N(key), ld_ds(query)] bf16 dS^T for the rel-pos-bias gradient, or null
  int ld_ds;
  float* dq_bias;        // [H*64] += column sums of dQ (q_bias gradient) or null
  float* dv_bias;        // [H*64] += column sums of dV (v_bias gradient) or null
  bf16* dqkv;            // [B, N, 3, H, 64]
  int B, H, N;
  float scale, p_drop;
};

__global__ void __launch_bounds__(BWD_WARPS * 32, 1) attn_bwd_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sK = sQ + NMAX * PITCH;
  bf16* sV = sK + NMAX * PITCH;
  bf16* sdO = sV + NMAX * PITCH;
  float* sdQ = reinterpret_cast<float*>(sdO + NMAX * PITCH);   // [NMAX][64]
  float* sLse = sdQ + NMAX * HD;                                // [NMAX] (log2 domain)
  float* sD = sLse + NMAX;                                      // [NMAX]
  bf16* sStage = reinterpret_cast<bf16*>(sD + NMAX);            // [BWD_WARPS][16][24]

  const int bh = blockIdx.x;
  const int b = bh / p.H, h = bh - b * p.H;
  const int N = p.N;
  const int ntile = (N + 15) >> 4;
  const int n_pad = ntile * 16;
  const long long row_stride = 3LL * p.H * HD;
  const long long o_stride = (long long)p.H * HD;
  const bf16* gq = p.qkv + (long long)b * N * row_stride + h * HD;
  const bf16* go = p.out + (long long)b * N * o_stride + h * HD;
  const bf16* gdo = p.dout + (long long)b * N * o_stride + h * HD;
  load_tile_rows(sQ, gq, row_stride, N, n_pad);
  load_tile_rows(sK, gq + p.H * HD, row_stride, N, n_pad);
  load_tile_rows(sV, gq + 2 * p.H * HD, row_stride, N, n_pad);
  load_tile_rows(sdO, gdo, o_stride, N, n_pad);
  for (int i = threadIdx.x; i < n_pad * HD; i += blockDim.x) sdQ[i] = 0.f;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) sLse[i] = i < N ? p.lse[(long long)bh * N + i] * LOG2E : 0.f;
  cp_async_wait_all();
  __syncthreads();
  // D_i = sum_d dO[i,d] * O[i,d]  (8 lanes per row)
  for (int idx = threadIdx.x; idx < n_pad * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = (idx & 7) * 8;
    float acc = 0.f;
    if (r < N) {
      const uint4 ov = *reinterpret_cast<const uint4*>(go + (long long)r * o_stride + c);
      const uint4 dv = *reinterpret_cast<const uint4*>(sdO + r * PITCH + c);
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
    if ((idx & 7) == 0) sD[r] = acc;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane & 3, qrow = lane >> 2;
  const bool drop = p.p_drop > 0.f && p.keep_bits != nullptr;
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const float sl2 = p.scale * LOG2E;

  const bool active = warp < ntile;
  const int jt = active ? warp : 0;
  const int jA = jt * 16 + qrow, jB = jA + 8;  // the two key rows this thread owns in C fragments
  uint32_t ka[4][4], va[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    ldsm_x4(smem_u32(sK + (jt * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8), ka[ks][0], ka[ks][1], ka[ks][2], ka[ks][3]);
    ldsm_x4(smem_u32(sV + (jt * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8), va[ks][0], va[ks][1], va[ks][2], va[ks][3]);
  }
  float dv[8][4], dk[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; }
  bf16* stage = sStage + warp * 16 * 24;

  // Every warp owns one 16-key tile; in step s it visits query tile (s + jt) mod ntile, so within a step the warps touch
  // DISJOINT dQ tiles and the shared-memory accumulation needs no atomics (block barrier between steps).
  for (int step = 0; step < ntile; ++step) {
    if (active) {
      int it = step + jt;
      if (it >= ntile) it -= ntile;
      // S^T = K_j Q_i^T and dP^T = V_j dO_i^T : [16 keys x 16 queries]
      float st[2][4], dp[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) { st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f; dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b0, b1, b2, b3;
        const int off = (it * 16 + (lane >> 4) * 8 + (lane & 7)) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(sQ + off), b0, b1, b2, b3);
        mma16816(st[0], ka[ks][0], ka[ks][1], ka[ks][2], ka[ks][3], b0, b1);
        mma16816(st[1], ka[ks][0], ka[ks][1], ka[ks][2], ka[ks][3], b2, b3);
        ldsm_x4(smem_u32(sdO + off), b0, b1, b2, b3);
        mma16816(dp[0], va[ks][0], va[ks][1], va[ks][2], va[ks][3], b0, b1);
        mma16816(dp[1], va[ks][0], va[ks][1], va[ks][2], va[ks][3], b2, b3);
      }
      // elementwise: P, dropout, dS
      float pt[2][4], ds[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int ibase = it * 16 + n * 8 + quad * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = ibase + (e & 1);                        // query
          const int j = (e < 2) ? jA : jB;                      // key
          float dsv = 0.f, ptv = 0.f;
          if (i < N && j < N) {
            float sv = st[n][e] * sl2;
            if (p.bias != nullptr) sv += __ldg(p.bias + ((long long)h * N + i) * p.ld_bias + j) * LOG2E;
            const float pv = exp2f(sv - sLse[i]);
            float keepf = 1.0f;
            if (drop) {
              const uint8_t byte = p.keep_bits[((long long)bh * N + i) * 32 + (j >> 3)];
              keepf = ((byte >> (j & 7)) & 1) ? inv_keep : 0.f;
            }
            ptv = pv * keepf;
            dsv = pv * (dp[n][e] * keepf - sD[i]);
          }
          pt[n][e] = ptv;
          ds[n][e] = dsv;
        }
      }
      // A fragments (m = keys, k = queries) from the C fragments
      const uint32_t pa0 = pack_bf16x2(pt[0][0], pt[0][1]), pa1 = pack_bf16x2(pt[0][2], pt[0][3]);
      const uint32_t pa2 = pack_bf16x2(pt[1][0], pt[1][1]), pa3 = pack_bf16x2(pt[1][2], pt[1][3]);
      const uint32_t da0 = pack_bf16x2(ds[0][0], ds[0][1]), da1 = pack_bf16x2(ds[0][2], ds[0][3]);
      const uint32_t da2 = pack_bf16x2(ds[1][0], ds[1][1]), da3 = pack_bf16x2(ds[1][2], ds[1][3]);
      if (p.ds_out != nullptr) {   // dS^T for the relative-position-bias gradient (reduced over the batch by a second kernel)
        bf16* d0 = p.ds_out + ((long long)bh * N + jA) * p.ld_ds + it * 16 + quad * 2;
        bf16* d1 = p.ds_out + ((long long)bh * N + jB) * p.ld_ds + it * 16 + quad * 2;
        if (jA < N) { *reinterpret_cast<uint32_t*>(d0) = da0; *reinterpret_cast<uint32_t*>(d0 + 8) = da2; }
        if (jB < N) { *reinterpret_cast<uint32_t*>(d1) = da1; *reinterpret_cast<uint32_t*>(d1 + 8) = da3; }
      }
      // dV_j += P~^T dO_i ; dK_j += dS^T Q_i   (B = [query][d] row-major -> ldmatrix.trans)
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t b0, b1, b2, b3;
        const int off = (it * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + dpair * 16 + (lane >> 4) * 8;
        ldsm_x4_t(smem_u32(sdO + off), b0, b1, b2, b3);
        mma16816(dv[dpair * 2], pa0, pa1, pa2, pa3, b0, b1);
        mma16816(dv[dpair * 2 + 1], pa0, pa1, pa2, pa3, b2, b3);
        ldsm_x4_t(smem_u32(sQ + off), b0, b1, b2, b3);
        mma16816(dk[dpair * 2], da0, da1, da2, da3, b0, b1);
        mma16816(dk[dpair * 2 + 1], da0, da1, da2, da3, b2, b3);
      }
      // dQ_i += dS K_j : stage dS^T [key][query] in smem, reload transposed as A (m = query, k = key)
      __syncwarp();
      *reinterpret_cast<uint32_t*>(stage + qrow * 24 + quad * 2) = da0;
      *reinterpret_cast<uint32_t*>(stage + (qrow + 8) * 24 + quad * 2) = da1;
      *reinterpret_cast<uint32_t*>(stage + qrow * 24 + 8 + quad * 2) = da2;
      *reinterpret_cast<uint32_t*>(stage + (qrow + 8) * 24 + 8 + quad * 2) = da3;
      __syncwarp();
      uint32_t sa0, sa1, sa2, sa3;
      ldsm_x4_t(smem_u32(stage + ((lane >> 4) * 8 + (lane & 7)) * 24 + ((lane >> 3) & 1) * 8), sa0, sa1, sa2, sa3);
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        float dq0[4] = {0.f, 0.f, 0.f, 0.f}, dq1[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(sK + (jt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + dpair * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
        mma16816(dq0, sa0, sa1, sa2, sa3, b0, b1);
        mma16816(dq1, sa0, sa1, sa2, sa3, b2, b3);
        float2* q0 = reinterpret_cast<float2*>(sdQ + (it * 16 + qrow) * HD + dpair * 16 + quad * 2);
        float2 v;
        v = q0[0]; v.x += dq0[0]; v.y += dq0[1]; q0[0] = v;
        v = q0[4 * HD]; v.x += dq0[2]; v.y += dq0[3]; q0[4 * HD] = v;          // row + 8  (float2 units: 8 * HD / 2)
        v = q0[4]; v.x += dq1[0]; v.y += dq1[1]; q0[4] = v;                    // cols + 8
        v = q0[4 * HD + 4]; v.x += dq1[2]; v.y += dq1[3]; q0[4 * HD + 4] = v;
      }
    }
    __syncthreads();
  }
  if (active) {
    // write dK (scaled) and dV for this key tile
    bf16* gdk = p.dqkv + (long long)b * N * row_stride + p.H * HD + h * HD;
    bf16* gdv = gdk + p.H * HD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + quad * 2;
      if (jA < N) {
        *reinterpret_cast<uint32_t*>(gdk + (long long)jA * row_stride + c) = pack_bf16x2(dk[dt][0] * p.scale, dk[dt][1] * p.scale);
        *reinterpret_cast<uint32_t*>(gdv + (long long)jA * row_stride + c) = pack_bf16x2(dv[dt][0], dv[dt][1]);
      }
      if (jB < N) {
        *reinterpret_cast<uint32_t*>(gdk + (long long)jB * row_stride + c) = pack_bf16x2(dk[dt][2] * p.scale, dk[dt][3] * p.scale);
        *reinterpret_cast<uint32_t*>(gdv + (long long)jB * row_stride + c) = pack_bf16x2(dv[dt][2], dv[dt][3]);
      }
      if (p.dv_bias != nullptr) {   // v_bias gradient: column sums of dV over this tile's keys (rows >= N are exactly zero)
        float s0 = dv[dt][0] + dv[dt][2], s1 = dv[dt][1] + dv[dt][3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        if (qrow == 0) { atomicAdd(p.dv_bias + h * HD + c, s0); atomicAdd(p.dv_bias + h * HD + c + 1, s1); }
      }
    }
  }
  __syncthreads();
  // dQ (scaled) -> global ; q_bias gradient = column sums (blockDim % 16 == 0, so a thread's column group is fixed)
  bf16* gdq = p.dqkv + (long long)b * N * row_stride + h * HD;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = threadIdx.x; idx < N * 16; idx += blockDim.x) {
    const int r = idx >> 4, c = (idx & 15) * 4;
    float4 v = *reinterpret_cast<const float4*>(sdQ + r * HD + c);
    v.x *= p.scale; v.y *= p.scale; v.z *= p.scale; v.w *= p.scale;
    cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(gdq + (long long)r * row_stride + c) = u;
  }
  if (p.dq_bias != nullptr) {
    cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
    cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
    if ((threadIdx.x & 31) < 16) {
      float* dst = p.dq_bias + h * HD + (threadIdx.x & 15) * 4;
      atomicAdd(dst, cs.x); atomicAdd(dst + 1, cs.y); atomicAdd(dst + 2, cs.z); atomicAdd(dst + 3, cs.w);
    }
  }
}

// Relative-position-bias table gradient: dtable[index[i, j], h] += sum_b dS[b, h, i, j]  (dS^T stored [B, H, j, ld] by attn_bwd)
__global__ void __launch_bounds__(256) relbias_grad_kernel(const bf16* __restrict__ ds, int B, int H, int N, int ld,
                                                           const int* __restrict__ rel_index, float* __restrict__ dtable) {
  const int half = ld >> 1;                       // pairs of queries
  const long long total = (long long)H * N * half;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int ip = (int)(t % half);
    const int j = (int)((t / half) % N);
    const int h = (int)(t / ((long long)half * N));
    const int i = ip * 2;
    if (i >= N) continue;
    float s0 = 0.f, s1 = 0.f;
    const bf16* src = ds + ((long long)h * N + j) * ld + i;
    const long long bstride = (long long)H * N * ld;
    for (int b = 0; b < B; ++b) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + b * bstride));
      s0 += v.x; s1 += v.y;
    }
    atomicAdd(dtable + (long long)rel_index[i * N + j] * H + h, s0);
    if (i + 1 < N) atomicAdd(dtable + (long long)rel_index[(i + 1) * N + j] * H + h, s1);
  }
}

constexpr size_t FWD_SMEM = 3 * NMAX * PITCH * sizeof(bf16);
constexpr size_t BWD_SMEM = 4 * NMAX * PITCH * sizeof(bf16) + NMAX * HD * sizeof(float) + 2 * NMAX * sizeof(float) + BWD_WARPS * 16 * 24 * sizeof(bf16);

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_attn_fwd(const void* qkv, const float* bias, int64_t ld_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim,
                                float scale, float p_drop, uint64_t seed, uint32_t stream_id, const uint8_t* keep_in, void* out, float* lse,
                                uint8_t* keep_bits, void* stream) {
  B200_CHECK_ARG(qkv != nullptr && out != nullptr, "attn_fwd: null pointer");
  B200_CHECK_ARG(head_dim == HD, "attn_fwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "attn_fwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "attn_fwd: bad p_drop");
  B200_CHECK_ARG(p_drop == 0.f || keep_bits != nullptr, "attn_fwd: dropout needs the keep_bits buffer [B,H,N,32]");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("attn_fwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  AttnFwdParams p;
  p.qkv = static_cast<const bf16*>(qkv); p.bias = bias; p.ld_bias = ld_bias; p.out = static_cast<bf16*>(out); p.lse = lse;
  p.keep_bits = keep_bits; p.keep_in = keep_in; p.B = B; p.H = H; p.N = N; p.scale = scale; p.p_drop = p_drop; p.seed = seed; p.stream_id = stream_id;
  attn_fwd_kernel<<<B * H, FWD_WARPS * 32, FWD_SMEM, STREAM>>>(p);
  B200_CHECK_LAUNCH("attn_fwd");
  return 0;
}

extern "C" int b200vit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias, int64_t ld_bias,
                                const uint8_t* keep_bits, void* ds_work, int32_t ld_ds, const int32_t* rel_index, float* dtable,
                                float* dq_bias, float* dv_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop,
                                void* dqkv, void* stream) {
  B200_CHECK_ARG(qkv && out && dout && lse && dqkv, "attn_bwd: null pointer");
  B200_CHECK_ARG(head_dim == HD, "attn_bwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "attn_bwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop == 0.f || keep_bits != nullptr, "attn_bwd: dropout needs keep_bits from the forward");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(dtable == nullptr || (rel_index != nullptr && ds_work != nullptr && ld_ds >= n_pad && ld_ds % 2 == 0),
                 "attn_bwd: dtable needs rel_index and a bf16 workspace [B,H,N,ld_ds] with even ld_ds >= %d", n_pad);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("attn_bwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  AttnBwdParams p;
  p.qkv = static_cast<const bf16*>(qkv); p.out = static_cast<const bf16*>(out); p.dout = static_cast<const bf16*>(dout); p.lse = lse;
  p.bias = bias; p.ld_bias = ld_bias; p.keep_bits = keep_bits;
  p.ds_out = dtable != nullptr ? static_cast<bf16*>(ds_work) : nullptr; p.ld_ds = ld_ds; p.dq_bias = dq_bias; p.dv_bias = dv_bias;
  p.dqkv = static_cast<bf16*>(dqkv); p.B = B; p.H = H; p.N = N; p.scale = scale; p.p_drop = p_drop;
  attn_bwd_kernel<<<B * H, BWD_WARPS * 32, BWD_SMEM, STREAM>>>(p);
  B200_CHECK_LAUNCH("attn_bwd");
  if (dtable != nullptr) {
    const int sms = b200vit_num_sms();
    relbias_grad_kernel<<<sms * 8, 256, 0, STREAM>>>(static_cast<const bf16*>(ds_work), B, H, N, ld_ds, rel_index, dtable);
    B200_CHECK_LAUNCH("relbias_grad");
  }
  return 0;
}

extern "C" int b200vit_dropout_mask(uint8_t* out, int32_t BH, int32_t N, float p_drop, uint64_t seed, uint32_t stream_id, void* stream) {
  B200_CHECK_ARG(out != nullptr && BH > 0 && N > 0, "dropout_mask: bad arguments");
  const int sms = b200vit_num_sms();
  dropout_mask_kernel<<<sms * 8, 256, 0, STREAM>>>(out, BH, N, p_drop, seed, stream_id);
  B200_CHECK_LAUNCH("dropout_mask");
  return 0;
}
