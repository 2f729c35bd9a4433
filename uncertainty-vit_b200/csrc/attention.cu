// Fused flash-style multi-head attention for ViT token counts (N <= 208, head_dim 64), forward and backward.
//   S = scale * q k^T + rel_pos_bias[h] ; P = softmax_j(S) ; P~ = dropout(P) ; O = P~ v
// (Attention.forward, modeling_finetune.py:145-188). One CTA per (batch, head); Q/K/V staged in shared memory with
// cp.async, mma.sync m16n8k16 bf16 tensor-core tiles, online softmax with quad-shuffle row reductions, counter-based
// Philox dropout (or an injected keep-mask), relative-position-bias add from an L2-resident, log2(e)-prescaled, -inf padded
// [H, N, ld] tensor (the key mask is baked into the padding).
// Backward: phase 1 — every warp owns one 16-key tile (dK/dV accumulate in registers), recomputes P from the saved
// log-sum-exp and parks dS^T as bf16 in shared memory; phase 2 — every warp owns one 16-query tile and forms dQ = dS K from
// that shared dS^T (no atomics anywhere). dS^T is also streamed out coalesced for the relative-position-bias table gradient,
// which a second kernel reduces over the batch and scatter-adds through the reference's relative_position_index.
#include "../../include/b200vit.h"
#include "attn_common.cuh"

namespace {

using namespace attn;

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct AttnFwdParams {
  const bf16* qkv;      // [B, N, 3, H, 64]
  const float* bias;    // [H, N, ld_bias] * log2(e), columns [N, n_pad) = -inf ; or null
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

struct FwdRowState {
  float o[8][4];
  float m[2], l[2];
};

// Bias values of one chunk of up to 4 8-key tiles for the two query rows of this thread.
struct BiasRegs {
  float2 r0[4], r1[4];
};
template <int NTS>
__device__ __forceinline__ void load_bias(BiasRegs& b, const float* brow0, const float* brow1, int j0, int quad) {
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
    b.r0[nt] = __ldg(reinterpret_cast<const float2*>(brow0 + j0 + nt * 8 + quad * 2));
    b.r1[nt] = __ldg(reinterpret_cast<const float2*>(brow1 + j0 + nt * 8 + quad * 2));
  }
}

// One chunk of NTS (<= 4) 8-key tiles starting at key j0 for the 16-query tile of this warp. `cur` holds this chunk's bias
// (fetched while the previous chunk was being processed).
template <int NTS, bool DROP, bool HAS_BIAS>
__device__ __forceinline__ void fwd_chunk(const AttnFwdParams& p, const bf16* sK, const bf16* sV, const uint32_t (&qa)[4][4], FwdRowState& st,
                                          int j0, int i0, int i1, const BiasRegs& cur, int bh, int lane, float sl2, uint32_t thresh) {
  const int quad = lane & 3;
  const int N = p.N;
  float s[NTS][4];
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
  for (int np = 0; np < NTS / 2; ++np) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(smem_u32(sK + (j0 + np * 16 + (lane >> 4) * 8 + (lane & 7)) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
      mma16816(s[np * 2], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
      mma16816(s[np * 2 + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
    }
  }
  float mx0 = st.m[0], mx1 = st.m[1];
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
    if (HAS_BIAS) {   // bias is pre-multiplied by log2(e); padding columns hold -inf (key mask)
      s[nt][0] = fmaf(s[nt][0], sl2, cur.r0[nt].x); s[nt][1] = fmaf(s[nt][1], sl2, cur.r0[nt].y);
      s[nt][2] = fmaf(s[nt][2], sl2, cur.r1[nt].x); s[nt][3] = fmaf(s[nt][3], sl2, cur.r1[nt].y);
    } else {
      const int j = j0 + nt * 8 + quad * 2;
      s[nt][0] = j < N ? s[nt][0] * sl2 : -INFINITY; s[nt][1] = j + 1 < N ? s[nt][1] * sl2 : -INFINITY;
      s[nt][2] = j < N ? s[nt][2] * sl2 : -INFINITY; s[nt][3] = j + 1 < N ? s[nt][3] * sl2 : -INFINITY;
    }
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float c0 = ex2(st.m[0] - mx0), c1 = ex2(st.m[1] - mx1);   // first chunk: ex2(-inf) = 0
  st.m[0] = mx0; st.m[1] = mx1;
  st.l[0] *= c0; st.l[1] *= c1;
#pragma unroll
  for (int i = 0; i < 8; ++i) { st.o[i][0] *= c0; st.o[i][1] *= c0; st.o[i][2] *= c1; st.o[i][3] *= c1; }
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
    s[nt][0] = ex2(s[nt][0] - mx0); s[nt][1] = ex2(s[nt][1] - mx0);
    s[nt][2] = ex2(s[nt][2] - mx1); s[nt][3] = ex2(s[nt][3] - mx1);
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
  st.l[0] += l0; st.l[1] += l1;
  if (DROP) {   // NTS <= 4: one 32-key Philox group per chunk (j0 is a multiple of 32)
    uint32_t w0 = 0u, w1 = 0u;   // keep bits of rows i0 / i1 for keys j0 .. j0+31 (this thread: 2 bits per 8-key tile)
    if (p.keep_in == nullptr) {
      const Philox4 r0 = dropout_group(p.seed, p.stream_id, bh, i0, quad, j0 >> 5);
      const Philox4 r1 = dropout_group(p.seed, p.stream_id, bh, i1, quad, j0 >> 5);
#pragma unroll
      for (int n4 = 0; n4 < NTS; ++n4) {
        const int sh = n4 * 8 + quad * 2;
        w0 |= (dropout_u16(r0, n4 * 2) >= thresh ? 1u : 0u) << sh;
        w0 |= (dropout_u16(r0, n4 * 2 + 1) >= thresh ? 2u : 0u) << sh;
        w1 |= (dropout_u16(r1, n4 * 2) >= thresh ? 1u : 0u) << sh;
        w1 |= (dropout_u16(r1, n4 * 2 + 1) >= thresh ? 2u : 0u) << sh;
      }
    } else {
#pragma unroll
      for (int n4 = 0; n4 < NTS; ++n4) {
        const int j = j0 + n4 * 8 + quad * 2;
        const int sh = n4 * 8 + quad * 2;
        if (i0 < N && j < N && p.keep_in[((long long)bh * N + i0) * N + j]) w0 |= 1u << sh;
        if (i0 < N && j + 1 < N && p.keep_in[((long long)bh * N + i0) * N + j + 1]) w0 |= 2u << sh;
        if (i1 < N && j < N && p.keep_in[((long long)bh * N + i1) * N + j]) w1 |= 1u << sh;
        if (i1 < N && j + 1 < N && p.keep_in[((long long)bh * N + i1) * N + j + 1]) w1 |= 2u << sh;
      }
    }
#pragma unroll
    for (int n4 = 0; n4 < NTS; ++n4) {
      const int sh = n4 * 8 + quad * 2;
      if (!((w0 >> sh) & 1u)) s[n4][0] = 0.f;
      if (!((w0 >> sh) & 2u)) s[n4][1] = 0.f;
      if (!((w1 >> sh) & 1u)) s[n4][2] = 0.f;
      if (!((w1 >> sh) & 2u)) s[n4][3] = 0.f;
    }
    w0 |= __shfl_xor_sync(0xffffffffu, w0, 1); w0 |= __shfl_xor_sync(0xffffffffu, w0, 2);
    w1 |= __shfl_xor_sync(0xffffffffu, w1, 1); w1 |= __shfl_xor_sync(0xffffffffu, w1, 2);
    if (quad == 0) {
      if (i0 < N) *reinterpret_cast<uint32_t*>(p.keep_bits + ((long long)bh * N + i0) * 32 + (j0 >> 3)) = w0;
      if (i1 < N) *reinterpret_cast<uint32_t*>(p.keep_bits + ((long long)bh * N + i1) * 32 + (j0 >> 3)) = w1;
    }
  }
  // O += P~ V   (the 1/(1-p) rescale of the kept probabilities is folded into the final normalisation)
#pragma unroll
  for (int kk = 0; kk < NTS / 2; ++kk) {
    const uint32_t a0 = pack_bf16x2(s[kk * 2][0], s[kk * 2][1]), a1 = pack_bf16x2(s[kk * 2][2], s[kk * 2][3]);
    const uint32_t a2 = pack_bf16x2(s[kk * 2 + 1][0], s[kk * 2 + 1][1]), a3 = pack_bf16x2(s[kk * 2 + 1][2], s[kk * 2 + 1][3]);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sV + (j0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + dp * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
      mma16816(st.o[dp * 2], a0, a1, a2, a3, b0, b1);
      mma16816(st.o[dp * 2 + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

template <bool DROP, bool HAS_BIAS>
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
  const float inv_keep = DROP ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const uint32_t thresh = (uint32_t)(p.p_drop * 65536.0f + 0.5f);
  const float sl2 = p.scale * LOG2E;

  for (int mt = warp; mt < ntile; mt += FWD_WARPS) {
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldsm_x4(smem_u32(sQ + (mt * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    FwdRowState st;
#pragma unroll
    for (int i = 0; i < 8; ++i) st.o[i][0] = st.o[i][1] = st.o[i][2] = st.o[i][3] = 0.f;
    st.m[0] = st.m[1] = -INFINITY;
    st.l[0] = st.l[1] = 0.f;
    const int i0 = mt * 16 + qrow, i1 = i0 + 8;
    const float* brow0 = HAS_BIAS ? p.bias + ((long long)h * N + min(i0, N - 1)) * p.ld_bias : nullptr;
    const float* brow1 = HAS_BIAS ? p.bias + ((long long)h * N + min(i1, N - 1)) * p.ld_bias : nullptr;
    // 32-key chunks; the bias of chunk c+1 is requested before chunk c is processed (software pipeline over the L2 latency)
    BiasRegs bcur, bnext;
    const int nfull = n_pad >> 5;                 // full 32-key chunks; a 16-key tail remains when n_pad % 32 != 0
    const bool tail = (n_pad & 31) != 0;
    if (HAS_BIAS) {
      if (nfull > 0) load_bias<4>(bcur, brow0, brow1, 0, quad);
      else load_bias<2>(bcur, brow0, brow1, 0, quad);
    }
#pragma unroll 1
    for (int c = 0; c < nfull; ++c) {
      if (HAS_BIAS) {
        if (c + 1 < nfull) load_bias<4>(bnext, brow0, brow1, (c + 1) * 32, quad);
        else if (tail) load_bias<2>(bnext, brow0, brow1, (c + 1) * 32, quad);
      }
      fwd_chunk<4, DROP, HAS_BIAS>(p, sK, sV, qa, st, c * 32, i0, i1, bcur, bh, lane, sl2, thresh);
      bcur = bnext;
    }
    if (tail) fwd_chunk<2, DROP, HAS_BIAS>(p, sK, sV, qa, st, nfull * 32, i0, i1, bcur, bh, lane, sl2, thresh);
    // finalise: row sums across the quad, normalise (and apply the dropout rescale), store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      st.l[r] += __shfl_xor_sync(0xffffffffu, st.l[r], 1);
      st.l[r] += __shfl_xor_sync(0xffffffffu, st.l[r], 2);
    }
    const float inv0 = inv_keep / st.l[0], inv1 = inv_keep / st.l[1];
    bf16* orow0 = p.out + ((long long)b * N + i0) * (p.H * HD) + h * HD;
    bf16* orow1 = p.out + ((long long)b * N + i1) * (p.H * HD) + h * HD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + quad * 2;
      if (i0 < N) *reinterpret_cast<uint32_t*>(orow0 + c) = pack_bf16x2(st.o[dt][0] * inv0, st.o[dt][1] * inv0);
      if (i1 < N) *reinterpret_cast<uint32_t*>(orow1 + c) = pack_bf16x2(st.o[dt][2] * inv1, st.o[dt][3] * inv1);
    }
    if (quad == 0 && p.lse != nullptr) {
      if (i0 < N) p.lse[(long long)bh * N + i0] = (st.m[0] + log2f(st.l[0])) / LOG2E;
      if (i1 < N) p.lse[(long long)bh * N + i1] = (st.m[1] + log2f(st.l[1])) / LOG2E;
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
    const int v = (jt & 3) * 2 + lo;
    const uint32_t w = v < 4 ? (v < 2 ? r.x : r.y) : (v < 6 ? r.z : r.w);
    out[idx] = ((v & 1) ? (w >> 16) : (w & 0xffffu)) >= thresh ? 1 : 0;
  }
}

// RelativePositionBias.forward (modeling_finetune.py:359-364) in the layout the attention kernels read:
//   out_fwd[h, i, j] = scale * table[index[i, j], h] for j < N, -inf for N <= j < ld     (forward: key mask baked in)
//   out_bwd[h, j, i] = scale * table[index[i, j], h] for i < N, 0 for N <= i < ld        (backward: transposed)
__global__ void rel_pos_bias_kernel(const float* __restrict__ table, const int* __restrict__ index, int N, int H, int ld, float scale,
                                    float* __restrict__ out_fwd, float* __restrict__ out_bwd) {
  const int total = H * N * ld;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int c = t % ld, r = (t / ld) % N, h = t / (ld * N);
    if (out_fwd != nullptr) out_fwd[t] = c < N ? scale * __ldg(table + (long long)index[r * N + c] * H + h) : -INFINITY;
    if (out_bwd != nullptr) out_bwd[t] = c < N ? scale * __ldg(table + (long long)index[c * N + r] * H + h) : 0.f;
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
  const float* bias_t;   // [H, N(key j), ld_bias(query i)] * log2(e), transposed ; or null
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
  bf16* sdS = sdO + NMAX * PITCH;                               // [NMAX(key j)][DSP(query i)]  dS^T
  float* sLse = reinterpret_cast<float*>(sdS + NMAX * DSP);     // [NMAX] (log2 domain)
  float* sD = sLse + NMAX;                                      // [NMAX]

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

  // ================= phase 1: this warp owns key tile jt; loop over query tiles =================
  if (active) {
    const int jt = warp;
    const int jA = jt * 16 + qrow, jB = jA + 8;  // the two key rows this thread owns in C fragments
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; }
    const float* btA = p.bias_t != nullptr ? p.bias_t + ((long long)h * N + min(jA, N - 1)) * p.ld_bias : nullptr;
    const float* btB = p.bias_t != nullptr ? p.bias_t + ((long long)h * N + min(jB, N - 1)) * p.ld_bias : nullptr;
    const uint8_t* kb_base = drop ? p.keep_bits + (long long)bh * N * 32 + jt * 2 : nullptr;

    // operands of the element-wise phase (bias^T, packed keep bits) are fetched ONE STEP AHEAD: their L2 latency is covered by a
    // whole step of tensor + element-wise work instead of sitting on the critical path of a 13-warp CTA
    float2 nbA[2], nbB[2];
    uint32_t nkw[2][2];
    auto fetch = [&](int it_, float2 (&fa)[2], float2 (&fb)[2], uint32_t (&fk)[2][2]) {
      const int ia_ = it_ * 16 + quad * 2;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        fa[n] = btA != nullptr ? __ldg(reinterpret_cast<const float2*>(btA + ia_ + n * 8)) : make_float2(0.f, 0.f);
        fb[n] = btB != nullptr ? __ldg(reinterpret_cast<const float2*>(btB + ia_ + n * 8)) : make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = ia_ + n * 8 + e;
          fk[n][e] = (drop && i < N) ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(kb_base + (long long)i * 32)) : 0xffffu;
        }
      }
    };
    fetch(0, nbA, nbB, nkw);
#pragma unroll 1
    for (int it = 0; it < ntile; ++it) {
      const int ia = it * 16 + quad * 2;   // queries ia, ia+1 (n-tile 0) and ia+8, ia+9 (n-tile 1)
      float2 bA[2] = {nbA[0], nbA[1]}, bB[2] = {nbB[0], nbB[1]};
      uint32_t kw[2][2] = {{nkw[0][0], nkw[0][1]}, {nkw[1][0], nkw[1][1]}};
      if (it + 1 < ntile) fetch(it + 1, nbA, nbB, nkw);
      // S^T = K_j Q_i^T and dP^T = V_j dO_i^T : [16 keys x 16 queries]
      float st[2][4], dp[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) { st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f; dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
        const int aoff = (jt * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8;
        const int boff = (it * 16 + (lane >> 4) * 8 + (lane & 7)) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(sK + aoff), a0, a1, a2, a3);
        ldsm_x4(smem_u32(sQ + boff), b0, b1, b2, b3);
        mma16816(st[0], a0, a1, a2, a3, b0, b1);
        mma16816(st[1], a0, a1, a2, a3, b2, b3);
        ldsm_x4(smem_u32(sV + aoff), a0, a1, a2, a3);
        ldsm_x4(smem_u32(sdO + boff), b0, b1, b2, b3);
        mma16816(dp[0], a0, a1, a2, a3, b0, b1);
        mma16816(dp[1], a0, a1, a2, a3, b2, b3);
      }
      // elementwise: P, dropout, dS.  C layout: rows = keys (jA: e<2, jB: e>=2), cols = queries ia + n*8 + (e&1)
      float pt[2][4], ds[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = ia + n * 8 + (e & 1);
          const int j = (e < 2) ? jA : jB;
          const float bias = (e < 2) ? ((e & 1) ? bA[n].y : bA[n].x) : ((e & 1) ? bB[n].y : bB[n].x);
          float ptv = 0.f, dsv = 0.f;
          if (i < N && j < N) {
            const float pv = ex2(fmaf(st[n][e], sl2, bias) - sLse[i]);
            const float keepf = ((kw[n][e & 1] >> (qrow + (e < 2 ? 0 : 8))) & 1u) ? inv_keep : 0.f;
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
      // park dS^T [key][query] for phase 2 (dQ) and for the bias-table gradient
      *reinterpret_cast<uint32_t*>(sdS + jA * DSP + ia) = da0;
      *reinterpret_cast<uint32_t*>(sdS + jB * DSP + ia) = da1;
      *reinterpret_cast<uint32_t*>(sdS + jA * DSP + ia + 8) = da2;
      *reinterpret_cast<uint32_t*>(sdS + jB * DSP + ia + 8) = da3;
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
    }
    // write dK (scaled) and dV for this key tile ; v_bias gradient = column sums of dV (rows >= N are exactly zero)
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
      if (p.dv_bias != nullptr) {
        float s0 = dv[dt][0] + dv[dt][2], s1 = dv[dt][1] + dv[dt][3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        if (qrow == 0) { atomicAdd(p.dv_bias + h * HD + c, s0); atomicAdd(p.dv_bias + h * HD + c + 1, s1); }
      }
    }
  }
  __syncthreads();

  // ================= phase 2: this warp owns query tile it; dQ_i = sum_j dS_ij K_j from the shared dS^T =================
  if (active) {
    const int it = warp;
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll 1
    for (int jt = 0; jt < ntile; ++jt) {
      uint32_t a0, a1, a2, a3;   // A[m = query][k = key] = transposed read of dS^T[key][query]
      ldsm_x4_t(smem_u32(sdS + (jt * 16 + (lane >> 4) * 8 + (lane & 7)) * DSP + it * 16 + ((lane >> 3) & 1) * 8), a0, a1, a2, a3);
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(sK + (jt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + dpair * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
        mma16816(dq[dpair * 2], a0, a1, a2, a3, b0, b1);
        mma16816(dq[dpair * 2 + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    bf16* gdq = p.dqkv + (long long)b * N * row_stride + h * HD;
    const int iA = it * 16 + qrow, iB = iA + 8;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + quad * 2;
      const float v0 = dq[dt][0] * p.scale, v1 = dq[dt][1] * p.scale, v2 = dq[dt][2] * p.scale, v3 = dq[dt][3] * p.scale;
      if (iA < N) *reinterpret_cast<uint32_t*>(gdq + (long long)iA * row_stride + c) = pack_bf16x2(v0, v1);
      if (iB < N) *reinterpret_cast<uint32_t*>(gdq + (long long)iB * row_stride + c) = pack_bf16x2(v2, v3);
      if (p.dq_bias != nullptr) {   // q_bias gradient (rows >= N of dS are exactly zero)
        float s0 = v0 + v2, s1 = v1 + v3;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        if (qrow == 0) { atomicAdd(p.dq_bias + h * HD + c, s0); atomicAdd(p.dq_bias + h * HD + c + 1, s1); }
      }
    }
  }
  // dS^T rows -> global, coalesced 16-byte chunks (for the relative-position-bias table gradient)
  if (p.ds_out != nullptr) {
    const int chunks = n_pad >> 3;
    for (int idx = threadIdx.x; idx < N * chunks; idx += blockDim.x) {
      const int j = idx / chunks, c = (idx - j * chunks) * 8;
      *reinterpret_cast<uint4*>(p.ds_out + ((long long)bh * N + j) * p.ld_ds + c) = *reinterpret_cast<const uint4*>(sdS + j * DSP + c);
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
    int b = 0;
    for (; b + 8 <= B; b += 8) {     // 8 independent loads in flight
      uint32_t w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = *reinterpret_cast<const uint32_t*>(src + (long long)(b + u) * bstride);
#pragma unroll
      for (int u = 0; u < 8; ++u) { const float2 v = unpack_bf16x2(w[u]); s0 += v.x; s1 += v.y; }
    }
    for (; b < B; ++b) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + (long long)b * bstride));
      s0 += v.x; s1 += v.y;
    }
    atomicAdd(dtable + (long long)rel_index[i * N + j] * H + h, s0);
    if (i + 1 < N) atomicAdd(dtable + (long long)rel_index[(i + 1) * N + j] * H + h, s1);
  }
}

constexpr size_t FWD_SMEM = 3 * NMAX * PITCH * sizeof(bf16);
constexpr size_t BWD_SMEM = 4 * NMAX * PITCH * sizeof(bf16) + NMAX * DSP * sizeof(bf16) + 2 * NMAX * sizeof(float);

template <bool DROP, bool HAS_BIAS>
cudaError_t launch_fwd(const AttnFwdParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<DROP, HAS_BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  attn_fwd_kernel<DROP, HAS_BIAS><<<p.B * p.H, FWD_WARPS * 32, FWD_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_attn_fwd_mma(const void* qkv, const float* bias, int64_t ld_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim,
                                float scale, float p_drop, uint64_t seed, uint32_t stream_id, const uint8_t* keep_in, void* out, float* lse,
                                uint8_t* keep_bits, void* stream) {
  B200_CHECK_ARG(qkv != nullptr && out != nullptr, "attn_fwd: null pointer");
  B200_CHECK_ARG(head_dim == HD, "attn_fwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "attn_fwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "attn_fwd: bad p_drop");
  B200_CHECK_ARG(p_drop == 0.f || keep_bits != nullptr, "attn_fwd: dropout needs the keep_bits buffer [B,H,N,32]");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(bias == nullptr || (ld_bias >= n_pad && ld_bias % 2 == 0 && (reinterpret_cast<uintptr_t>(bias) & 7) == 0),
                 "attn_fwd: bias must be the padded layout of b200vit_rel_pos_bias ([H,N,ld], ld even >= %d, 8-byte aligned)", n_pad);
  AttnFwdParams p;
  p.qkv = static_cast<const bf16*>(qkv); p.bias = bias; p.ld_bias = ld_bias; p.out = static_cast<bf16*>(out); p.lse = lse;
  p.keep_bits = keep_bits; p.keep_in = keep_in; p.B = B; p.H = H; p.N = N; p.scale = scale; p.p_drop = p_drop; p.seed = seed; p.stream_id = stream_id;
  cudaError_t e;
  if (p_drop > 0.f) e = bias != nullptr ? launch_fwd<true, true>(p, STREAM) : launch_fwd<true, false>(p, STREAM);
  else e = bias != nullptr ? launch_fwd<false, true>(p, STREAM) : launch_fwd<false, false>(p, STREAM);
  if (e != cudaSuccess) { b200vit_set_error("attn_fwd: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int b200vit_attn_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias_t, int64_t ld_bias,
                                const uint8_t* keep_bits, void* ds_work, int32_t ld_ds, const int32_t* rel_index, float* dtable,
                                float* dq_bias, float* dv_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop,
                                void* dqkv, void* stream) {
  B200_CHECK_ARG(qkv && out && dout && lse && dqkv, "attn_bwd: null pointer");
  B200_CHECK_ARG(head_dim == HD, "attn_bwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "attn_bwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop == 0.f || keep_bits != nullptr, "attn_bwd: dropout needs keep_bits from the forward");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(bias_t == nullptr || (ld_bias >= n_pad && ld_bias % 2 == 0 && (reinterpret_cast<uintptr_t>(bias_t) & 7) == 0),
                 "attn_bwd: bias_t must be the transposed padded layout of b200vit_rel_pos_bias ([H,N,ld], ld even >= %d)", n_pad);
  B200_CHECK_ARG(dtable == nullptr || (rel_index != nullptr && ds_work != nullptr && ld_ds >= n_pad && ld_ds % 8 == 0 &&
                                       (reinterpret_cast<uintptr_t>(ds_work) & 15) == 0),
                 "attn_bwd: dtable needs rel_index and a 16-byte aligned bf16 workspace [B,H,N,ld_ds] with ld_ds %% 8 == 0, >= %d", n_pad);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("attn_bwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  AttnBwdParams p;
  p.qkv = static_cast<const bf16*>(qkv); p.out = static_cast<const bf16*>(out); p.dout = static_cast<const bf16*>(dout); p.lse = lse;
  p.bias_t = bias_t; p.ld_bias = ld_bias; p.keep_bits = keep_bits;
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

// internal (not part of the public ABI): shared by the dual-stream backward in wattention.cu
int b200vit_relbias_grad_launch(const void* ds_work, int B, int H, int N, int ld_ds, const int32_t* rel_index, float* dtable, void* stream) {
  const int sms = b200vit_num_sms();
  relbias_grad_kernel<<<sms * 8, 256, 0, STREAM>>>(static_cast<const bf16*>(ds_work), B, H, N, ld_ds, rel_index, dtable);
  B200_CHECK_LAUNCH("relbias_grad");
  return 0;
}

extern "C" int b200vit_dropout_mask(uint8_t* out, int32_t BH, int32_t N, float p_drop, uint64_t seed, uint32_t stream_id, void* stream) {
  B200_CHECK_ARG(out != nullptr && BH > 0 && N > 0, "dropout_mask: bad arguments");
  const int sms = b200vit_num_sms();
  dropout_mask_kernel<<<sms * 8, 256, 0, STREAM>>>(out, BH, N, p_drop, seed, stream_id);
  B200_CHECK_LAUNCH("dropout_mask");
  return 0;
}

extern "C" int b200vit_rel_pos_bias(const float* table, const int32_t* index, int32_t N, int32_t H, int32_t ld, float scale, float* out_fwd,
                                    float* out_bwd_t, void* stream) {
  B200_CHECK_ARG(table != nullptr && index != nullptr && (out_fwd != nullptr || out_bwd_t != nullptr) && N > 0 && H > 0 && ld >= N,
                 "rel_pos_bias: bad arguments");
  const int sms = b200vit_num_sms();
  rel_pos_bias_kernel<<<sms * 4, 256, 0, STREAM>>>(table, index, N, H, ld, scale, out_fwd, out_bwd_t);
  B200_CHECK_LAUNCH("rel_pos_bias");
  return 0;
}
