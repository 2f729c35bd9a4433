// Wasserstein-distance attention of the dual-stream ("--stochastic") transformer, forward.
// dist Attention.forward (modeling_finetune_dist.py:111-179) + wasserstein_distance_matmul (uncertainty_evaluations.py:276-294):
//   m1 = sigmoid(scale*q)  m2 = sigmoid(k)  u1 = sqrt(max(sigmoid(cq),1e-24))  u2 = sqrt(max(sigmoid(ck),1e-24))
//   D_ij = |m1_i|^2 + |m2_j|^2 - 2 m1_i.m2_j + sum c1_i + sum c2_j - 2 u1_i.u2_j  =  r_i + c_j - 2 [m1_i|u1_i].[m2_j|u2_j]   (ONE K=128 product)
//   A = sigmoid(-D + 1e-24) + rel_pos_bias ; P = softmax_j(A) ; P~ = dropout(P) ; mean = P~ v ; cov = (P~)^2 cv
// One CTA per (batch, head): the sigmoid / sqrt transforms are applied once while staging [m|u] into shared memory; the row/column
// norms are taken from the SAME bf16-rounded operands the tensor cores multiply, so D keeps its cancellation structure.
#include "../../include/b200vit.h"
#include "attn_common.cuh"

namespace {

using namespace attn;

constexpr int XP = 2 * HD + 8;    // bf16 row pitch of the [m | u] operand (272 B = 17 x 16 B: conflict-free ldmatrix)
constexpr int WF_WARPS = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct WAttnFwdParams {
  const bf16* qkv_m;    // [B, N, 3, H, 64]  mean stream (q, k, v)
  const bf16* qkv_c;    // [B, N, 3, H, 64]  covariance stream AFTER elu()+1 (cq, ck, cv)
  const float* bias;    // [H, N, ld_bias] * log2(e), -inf padded (required: the reference adds rel_pos_bias unconditionally, :155)
  long long ld_bias;
  bf16* out_m;          // [B, N, H*64]
  bf16* out_c;          // [B, N, H*64]
  float* lse;           // [B, H, N]
  uint8_t* keep_bits;
  const uint8_t* keep_in;
  int B, H, N;
  float scale, p_drop;
  uint64_t seed;
  const uint64_t* seed_dev;   // when non-null: the Philox key is read from device memory (CUDA-graph replays)
  uint32_t stream_id;
};

// Stages rows of [sigmoid(s*mean) | sqrt(sigmoid(cov))] (bf16) into smem and their squared norms (from the rounded values).
__device__ __forceinline__ void stage_mu(bf16* sX, float* sNorm, const bf16* gmean, const bf16* gcov, long long row_stride, float mean_scale,
                                         int N, int n_pad) {
  for (int idx = threadIdx.x; idx < n_pad * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = (idx & 7) * 8;
    float nrm = 0.f;
    uint4 om = make_uint4(0u, 0u, 0u, 0u), ou = make_uint4(0u, 0u, 0u, 0u);
    if (r < N) {
      const uint4 qm = *reinterpret_cast<const uint4*>(gmean + (long long)r * row_stride + c);
      const uint4 qc = *reinterpret_cast<const uint4*>(gcov + (long long)r * row_stride + c);
      const uint32_t* pm = &qm.x; const uint32_t* pc = &qc.x;
      uint32_t* wm = &om.x; uint32_t* wu = &ou.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = unpack_bf16x2(pm[k]), b = unpack_bf16x2(pc[k]);
        wm[k] = pack_bf16x2(sigmoidf_(a.x * mean_scale), sigmoidf_(a.y * mean_scale));
        wu[k] = pack_bf16x2(sqrtf(fmaxf(sigmoidf_(b.x), 1e-24f)), sqrtf(fmaxf(sigmoidf_(b.y), 1e-24f)));
        const float2 ra = unpack_bf16x2(wm[k]), rb = unpack_bf16x2(wu[k]);
        nrm += ra.x * ra.x + ra.y * ra.y + rb.x * rb.x + rb.y * rb.y;
      }
    }
    *reinterpret_cast<uint4*>(sX + r * XP + c) = om;
    *reinterpret_cast<uint4*>(sX + r * XP + HD + c) = ou;
    nrm += __shfl_xor_sync(0xffffffffu, nrm, 1);
    nrm += __shfl_xor_sync(0xffffffffu, nrm, 2);
    nrm += __shfl_xor_sync(0xffffffffu, nrm, 4);
    if ((idx & 7) == 0) sNorm[r] = nrm;
  }
}

struct WRowState {
  float om[8][4], oc[8][4];
  float m[2], l[2];
};

template <int NTS, bool DROP>
__device__ __forceinline__ void wfwd_chunk(const WAttnFwdParams& p, const bf16* sX2, const float* sCn, const bf16* sV, const bf16* sCV,
                                           const uint32_t (&qa)[8][4], WRowState& st, float rn0, float rn1, int j0, int i0, int i1,
                                           const float* brow0, const float* brow1, int bh, int lane, uint32_t thresh) {
  const int quad = lane & 3;
  const int N = p.N;
  float2 bv0[NTS], bv1[NTS];
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
    bv0[nt] = __ldg(reinterpret_cast<const float2*>(brow0 + j0 + nt * 8 + quad * 2));
    bv1[nt] = __ldg(reinterpret_cast<const float2*>(brow1 + j0 + nt * 8 + quad * 2));
  }
  float s[NTS][4];
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
  for (int np = 0; np < NTS / 2; ++np) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(smem_u32(sX2 + (j0 + np * 16 + (lane >> 4) * 8 + (lane & 7)) * XP + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
      mma16816(s[np * 2], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
      mma16816(s[np * 2 + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
    }
  }
  float mx0 = st.m[0], mx1 = st.m[1];
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
    const float2 cn = *reinterpret_cast<const float2*>(sCn + j0 + nt * 8 + quad * 2);
    // D = r_i + c_j - 2 s ; t = sigmoid(-D) = 1 / (1 + e^D) ; logit (log2 domain) = t * log2e + bias * log2e
    const float d0 = rn0 + cn.x - 2.f * s[nt][0], d1 = rn0 + cn.y - 2.f * s[nt][1];
    const float d2 = rn1 + cn.x - 2.f * s[nt][2], d3 = rn1 + cn.y - 2.f * s[nt][3];
    s[nt][0] = fmaf(__fdividef(1.0f, 1.0f + ex2(d0 * LOG2E)), LOG2E, bv0[nt].x);
    s[nt][1] = fmaf(__fdividef(1.0f, 1.0f + ex2(d1 * LOG2E)), LOG2E, bv0[nt].y);
    s[nt][2] = fmaf(__fdividef(1.0f, 1.0f + ex2(d2 * LOG2E)), LOG2E, bv1[nt].x);
    s[nt][3] = fmaf(__fdividef(1.0f, 1.0f + ex2(d3 * LOG2E)), LOG2E, bv1[nt].y);
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float c0 = ex2(st.m[0] - mx0), c1 = ex2(st.m[1] - mx1);
  st.m[0] = mx0; st.m[1] = mx1;
  st.l[0] *= c0; st.l[1] *= c1;
  const float c0s = c0 * c0, c1s = c1 * c1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    st.om[i][0] *= c0; st.om[i][1] *= c0; st.om[i][2] *= c1; st.om[i][3] *= c1;
    st.oc[i][0] *= c0s; st.oc[i][1] *= c0s; st.oc[i][2] *= c1s; st.oc[i][3] *= c1s;
  }
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
    s[nt][0] = ex2(s[nt][0] - mx0); s[nt][1] = ex2(s[nt][1] - mx0);
    s[nt][2] = ex2(s[nt][2] - mx1); s[nt][3] = ex2(s[nt][3] - mx1);
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
  st.l[0] += l0; st.l[1] += l1;
  if (DROP) {
    // NTS <= 4: one 32-key group per chunk (j0 is a multiple of 32)
    uint32_t w0 = 0u, w1 = 0u;
    if (p.keep_in == nullptr) {
      const uint64_t seed = p.seed_dev != nullptr ? __ldg(reinterpret_cast<const unsigned long long*>(p.seed_dev)) : p.seed;
      const Philox4 r0 = dropout_group(seed, p.stream_id, bh, i0, quad, j0 >> 5);
      const Philox4 r1 = dropout_group(seed, p.stream_id, bh, i1, quad, j0 >> 5);
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
  // mean += P~ V ; cov += (P~)^2 CV
#pragma unroll
  for (int kk = 0; kk < NTS / 2; ++kk) {
    const uint32_t a0 = pack_bf16x2(s[kk * 2][0], s[kk * 2][1]), a1 = pack_bf16x2(s[kk * 2][2], s[kk * 2][3]);
    const uint32_t a2 = pack_bf16x2(s[kk * 2 + 1][0], s[kk * 2 + 1][1]), a3 = pack_bf16x2(s[kk * 2 + 1][2], s[kk * 2 + 1][3]);
    const uint32_t q0 = pack_bf16x2(s[kk * 2][0] * s[kk * 2][0], s[kk * 2][1] * s[kk * 2][1]);
    const uint32_t q1 = pack_bf16x2(s[kk * 2][2] * s[kk * 2][2], s[kk * 2][3] * s[kk * 2][3]);
    const uint32_t q2 = pack_bf16x2(s[kk * 2 + 1][0] * s[kk * 2 + 1][0], s[kk * 2 + 1][1] * s[kk * 2 + 1][1]);
    const uint32_t q3 = pack_bf16x2(s[kk * 2 + 1][2] * s[kk * 2 + 1][2], s[kk * 2 + 1][3] * s[kk * 2 + 1][3]);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b0, b1, b2, b3;
      const int off = (j0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + dp * 16 + (lane >> 4) * 8;
      ldsm_x4_t(smem_u32(sV + off), b0, b1, b2, b3);
      mma16816(st.om[dp * 2], a0, a1, a2, a3, b0, b1);
      mma16816(st.om[dp * 2 + 1], a0, a1, a2, a3, b2, b3);
      ldsm_x4_t(smem_u32(sCV + off), b0, b1, b2, b3);
      mma16816(st.oc[dp * 2], q0, q1, q2, q3, b0, b1);
      mma16816(st.oc[dp * 2 + 1], q0, q1, q2, q3, b2, b3);
    }
  }
}

template <bool DROP>
__global__ void __launch_bounds__(WF_WARPS * 32, 1) wattn_fwd_kernel(const WAttnFwdParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sX1 = reinterpret_cast<bf16*>(smem);
  bf16* sX2 = sX1 + NMAX * XP;
  bf16* sV = sX2 + NMAX * XP;
  bf16* sCV = sV + NMAX * PITCH;
  float* sRn = reinterpret_cast<float*>(sCV + NMAX * PITCH);
  float* sCn = sRn + NMAX;
  const int bh = blockIdx.x;
  const int b = bh / p.H, h = bh - b * p.H;
  const int N = p.N;
  const int ntile = (N + 15) >> 4;
  const int n_pad = ntile * 16;
  const long long row_stride = 3LL * p.H * HD;
  const bf16* gm = p.qkv_m + (long long)b * N * row_stride + h * HD;
  const bf16* gc = p.qkv_c + (long long)b * N * row_stride + h * HD;
  load_tile_rows(sV, gm + 2 * p.H * HD, row_stride, N, n_pad);
  load_tile_rows(sCV, gc + 2 * p.H * HD, row_stride, N, n_pad);
  stage_mu(sX1, sRn, gm, gc, row_stride, p.scale, N, n_pad);                      // q is scaled BEFORE the sigmoid, cq is not (:135-136)
  stage_mu(sX2, sCn, gm + p.H * HD, gc + p.H * HD, row_stride, 1.0f, N, n_pad);
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane & 3, qrow = lane >> 2;
  const float inv_keep = DROP ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const uint32_t thresh = (uint32_t)(p.p_drop * 65536.0f + 0.5f);

  for (int mt = warp; mt < ntile; mt += WF_WARPS) {
    uint32_t qa[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      ldsm_x4(smem_u32(sX1 + (mt * 16 + (lane & 15)) * XP + ks * 16 + (lane >> 4) * 8), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    WRowState st;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      st.om[i][0] = st.om[i][1] = st.om[i][2] = st.om[i][3] = 0.f;
      st.oc[i][0] = st.oc[i][1] = st.oc[i][2] = st.oc[i][3] = 0.f;
    }
    st.m[0] = st.m[1] = -INFINITY;
    st.l[0] = st.l[1] = 0.f;
    const int i0 = mt * 16 + qrow, i1 = i0 + 8;
    const float rn0 = sRn[i0], rn1 = sRn[i1];
    const float* brow0 = p.bias + ((long long)h * N + min(i0, N - 1)) * p.ld_bias;
    const float* brow1 = p.bias + ((long long)h * N + min(i1, N - 1)) * p.ld_bias;
    int j0 = 0;
    for (; j0 + 32 <= n_pad; j0 += 32) wfwd_chunk<4, DROP>(p, sX2, sCn, sV, sCV, qa, st, rn0, rn1, j0, i0, i1, brow0, brow1, bh, lane, thresh);
    if (j0 < n_pad) wfwd_chunk<2, DROP>(p, sX2, sCn, sV, sCV, qa, st, rn0, rn1, j0, i0, i1, brow0, brow1, bh, lane, thresh);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      st.l[r] += __shfl_xor_sync(0xffffffffu, st.l[r], 1);
      st.l[r] += __shfl_xor_sync(0xffffffffu, st.l[r], 2);
    }
    const float inv0 = inv_keep / st.l[0], inv1 = inv_keep / st.l[1];
    const float sq0 = inv0 * inv0, sq1 = inv1 * inv1;          // cov = (dropout(P))^2 @ cv : square AFTER the 1/(1-p) rescale (:158-162)
    const long long orow0 = ((long long)b * N + i0) * (p.H * HD) + h * HD, orow1 = ((long long)b * N + i1) * (p.H * HD) + h * HD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + quad * 2;
      if (i0 < N) {
        *reinterpret_cast<uint32_t*>(p.out_m + orow0 + c) = pack_bf16x2(st.om[dt][0] * inv0, st.om[dt][1] * inv0);
        *reinterpret_cast<uint32_t*>(p.out_c + orow0 + c) = pack_bf16x2(st.oc[dt][0] * sq0, st.oc[dt][1] * sq0);
      }
      if (i1 < N) {
        *reinterpret_cast<uint32_t*>(p.out_m + orow1 + c) = pack_bf16x2(st.om[dt][2] * inv1, st.om[dt][3] * inv1);
        *reinterpret_cast<uint32_t*>(p.out_c + orow1 + c) = pack_bf16x2(st.oc[dt][2] * sq1, st.oc[dt][3] * sq1);
      }
    }
    if (quad == 0 && p.lse != nullptr) {
      if (i0 < N) p.lse[(long long)bh * N + i0] = (st.m[0] + log2f(st.l[0])) / LOG2E;
      if (i1 < N) p.lse[(long long)bh * N + i1] = (st.m[1] + log2f(st.l[1])) / LOG2E;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, kernel 1: key-tile owners. dV, dCV, dK, dCK, kappa; streams dD^T (for kernel 2) and dA^T (bias-table gradient).
//   dP~ = dOm v^T + 2 P~ (dOc cv^T) ; dP = dP~ M/(1-p) ; dA = P (dP - Delta_i), Delta_i = dOm_i.Om_i + 2 dOc_i.Oc_i
//   dD = -dA t (1 - t), t = sigmoid(-D) ; kappa_j = sum_i dD_ij ; dX2_j = sum_i dD_ij [m1_i | u1_i]
//   dm2 = 2 (kappa m2 - dX2_m) ; dc2 = kappa - dX2_u / u2 ; dk = dm2 m2 (1-m2) ; dck(pre-ELU) = dc2 c2 (1-c2) elu'
// ------------------------------------------------------------------------------------------------
constexpr int WB1_WARPS = 7;

struct WAttnBwdParams {
  const bf16* qkv_m; const bf16* qkv_c;     // forward inputs (cov after elu+1)
  const bf16* out_m; const bf16* out_c;     // forward outputs
  const bf16* dout_m; const bf16* dout_c;
  const float* lse;
  const float* bias_t; long long ld_bias;   // transposed padded bias * log2e
  const uint8_t* keep_bits;
  bf16* dD;                                 // [B, H, N(key), ld_ds(query)] bf16 workspace: dD^T
  bf16* dA;                                 // same layout: dA^T (bias gradient) or null
  int ld_ds;
  bf16* dqkv_m; bf16* dqkv_c;               // outputs (cov: gradient w.r.t. the PRE-activation of elu+1)
  float* dq_bias; float* dv_bias; float* dcq_bias; float* dcv_bias;   // [H*64] += or null
  int B, H, N;
  float scale, p_drop;
};

__device__ __forceinline__ float elu1_grad(float y) { return y > 1.0f ? 1.0f : y; }   // d/dz (elu(z)+1) expressed through y = elu(z)+1

__global__ void __launch_bounds__(WB1_WARPS * 32, 1) wattn_bwd_kv_kernel(const WAttnBwdParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sX1 = reinterpret_cast<bf16*>(smem);
  bf16* sdOm = sX1 + NMAX * XP;
  bf16* sdOc = sdOm + NMAX * PITCH;
  bf16* sV = sdOc + NMAX * PITCH;
  bf16* sCV = sV + NMAX * PITCH;
  bf16* sX2w = sCV + NMAX * PITCH;                               // [WB1_WARPS][16][XP]
  float* sRn = reinterpret_cast<float*>(sX2w + WB1_WARPS * 16 * XP);
  float* sLse = sRn + NMAX;
  float* sDelta = sLse + NMAX;
  float* sCnW = sDelta + NMAX;                                   // [WB1_WARPS][16]

  const int bh = blockIdx.x;
  const int b = bh / p.H, h = bh - b * p.H;
  const int N = p.N;
  const int ntile = (N + 15) >> 4;
  const int n_pad = ntile * 16;
  const long long row_stride = 3LL * p.H * HD;
  const long long o_stride = (long long)p.H * HD;
  const bf16* gm = p.qkv_m + (long long)b * N * row_stride + h * HD;
  const bf16* gc = p.qkv_c + (long long)b * N * row_stride + h * HD;
  const long long obase = (long long)b * N * o_stride + h * HD;
  load_tile_rows(sV, gm + 2 * p.H * HD, row_stride, N, n_pad);
  load_tile_rows(sCV, gc + 2 * p.H * HD, row_stride, N, n_pad);
  load_tile_rows(sdOm, p.dout_m + obase, o_stride, N, n_pad);
  load_tile_rows(sdOc, p.dout_c + obase, o_stride, N, n_pad);
  stage_mu(sX1, sRn, gm, gc, row_stride, p.scale, N, n_pad);
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) sLse[i] = i < N ? p.lse[(long long)bh * N + i] * LOG2E : 0.f;
  cp_async_wait_all();
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_pad * 8; idx += blockDim.x) {   // Delta_i = dOm.Om + 2 dOc.Oc
    const int r = idx >> 3, c = (idx & 7) * 8;
    float acc = 0.f;
    if (r < N) {
      const uint4 om = *reinterpret_cast<const uint4*>(p.out_m + obase + (long long)r * o_stride + c);
      const uint4 oc = *reinterpret_cast<const uint4*>(p.out_c + obase + (long long)r * o_stride + c);
      const uint4 dm = *reinterpret_cast<const uint4*>(sdOm + r * PITCH + c);
      const uint4 dc = *reinterpret_cast<const uint4*>(sdOc + r * PITCH + c);
      const uint32_t* a = &om.x; const uint32_t* bq = &oc.x; const uint32_t* d1 = &dm.x; const uint32_t* d2 = &dc.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = unpack_bf16x2(a[k]), y = unpack_bf16x2(bq[k]), u = unpack_bf16x2(d1[k]), v = unpack_bf16x2(d2[k]);
        acc += x.x * u.x + x.y * u.y + 2.f * (y.x * v.x + y.y * v.y);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if ((idx & 7) == 0) sDelta[r] = acc;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane & 3, qrow = lane >> 2;
  const bool drop = p.p_drop > 0.f && p.keep_bits != nullptr;
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  bf16* myX2 = sX2w + warp * 16 * XP;
  float* myCn = sCnW + warp * 16;

#pragma unroll 1
  for (int jt = warp; jt < ntile; jt += WB1_WARPS) {
    // ---- stage this key tile's [m2 | u2] (16 x 128) and its norms
    __syncwarp();
    for (int idx = lane; idx < 16 * 8; idx += 32) {
      const int r = idx >> 3, c = (idx & 7) * 8;
      const int j = jt * 16 + r;
      float nrm = 0.f;
      uint4 om = make_uint4(0u, 0u, 0u, 0u), ou = make_uint4(0u, 0u, 0u, 0u);
      if (j < N) {
        const uint4 km = *reinterpret_cast<const uint4*>(gm + p.H * HD + (long long)j * row_stride + c);
        const uint4 kc = *reinterpret_cast<const uint4*>(gc + p.H * HD + (long long)j * row_stride + c);
        const uint32_t* pm = &km.x; const uint32_t* pc = &kc.x;
        uint32_t* wm = &om.x; uint32_t* wu = &ou.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = unpack_bf16x2(pm[k]), bb = unpack_bf16x2(pc[k]);
          wm[k] = pack_bf16x2(sigmoidf_(a.x), sigmoidf_(a.y));
          wu[k] = pack_bf16x2(sqrtf(fmaxf(sigmoidf_(bb.x), 1e-24f)), sqrtf(fmaxf(sigmoidf_(bb.y), 1e-24f)));
          const float2 ra = unpack_bf16x2(wm[k]), rb = unpack_bf16x2(wu[k]);
          nrm += ra.x * ra.x + ra.y * ra.y + rb.x * rb.x + rb.y * rb.y;
        }
      }
      *reinterpret_cast<uint4*>(myX2 + r * XP + c) = om;
      *reinterpret_cast<uint4*>(myX2 + r * XP + HD + c) = ou;
      nrm += __shfl_xor_sync(0xffffffffu, nrm, 1);
      nrm += __shfl_xor_sync(0xffffffffu, nrm, 2);
      nrm += __shfl_xor_sync(0xffffffffu, nrm, 4);
      if ((idx & 7) == 0) myCn[r] = nrm;
    }
    __syncwarp();
    uint32_t ka2[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      ldsm_x4(smem_u32(myX2 + (lane & 15) * XP + ks * 16 + (lane >> 4) * 8), ka2[ks][0], ka2[ks][1], ka2[ks][2], ka2[ks][3]);
    const int jA = jt * 16 + qrow, jB = jA + 8;
    const float cnA = myCn[qrow], cnB = myCn[qrow + 8];
    float dv[8][4], dcv[8][4], dx2[16][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; dcv[i][0] = dcv[i][1] = dcv[i][2] = dcv[i][3] = 0.f; }
#pragma unroll
    for (int i = 0; i < 16; ++i) dx2[i][0] = dx2[i][1] = dx2[i][2] = dx2[i][3] = 0.f;
    float kapA = 0.f, kapB = 0.f;
    const float* btA = p.bias_t + ((long long)h * N + min(jA, N - 1)) * p.ld_bias;
    const float* btB = p.bias_t + ((long long)h * N + min(jB, N - 1)) * p.ld_bias;
    const uint8_t* kb_base = drop ? p.keep_bits + (long long)bh * N * 32 + jt * 2 : nullptr;

#pragma unroll 1
    for (int it = 0; it < ntile; ++it) {
      const int ia = it * 16 + quad * 2;
      float2 bA[2], bB[2];
      uint32_t kw[2][2];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        bA[n] = __ldg(reinterpret_cast<const float2*>(btA + ia + n * 8));
        bB[n] = __ldg(reinterpret_cast<const float2*>(btB + ia + n * 8));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = ia + n * 8 + e;
          kw[n][e] = (drop && i < N) ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(kb_base + (long long)i * 32)) : 0xffffu;
        }
      }
      float st[2][4], g1[2][4], g2[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) st[n][e] = g1[n][e] = g2[n][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(sX1 + (it * 16 + (lane >> 4) * 8 + (lane & 7)) * XP + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
        mma16816(st[0], ka2[ks][0], ka2[ks][1], ka2[ks][2], ka2[ks][3], b0, b1);
        mma16816(st[1], ka2[ks][0], ka2[ks][1], ka2[ks][2], ka2[ks][3], b2, b3);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
        const int aoff = (jt * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8;
        const int boff = (it * 16 + (lane >> 4) * 8 + (lane & 7)) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(sV + aoff), a0, a1, a2, a3);
        ldsm_x4(smem_u32(sdOm + boff), b0, b1, b2, b3);
        mma16816(g1[0], a0, a1, a2, a3, b0, b1);
        mma16816(g1[1], a0, a1, a2, a3, b2, b3);
        ldsm_x4(smem_u32(sCV + aoff), a0, a1, a2, a3);
        ldsm_x4(smem_u32(sdOc + boff), b0, b1, b2, b3);
        mma16816(g2[0], a0, a1, a2, a3, b0, b1);
        mma16816(g2[1], a0, a1, a2, a3, b2, b3);
      }
      float pt[2][4], dd[2][4], da[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = ia + n * 8 + (e & 1);
          const int j = (e < 2) ? jA : jB;
          const float bias = (e < 2) ? ((e & 1) ? bA[n].y : bA[n].x) : ((e & 1) ? bB[n].y : bB[n].x);
          float ptv = 0.f, ddv = 0.f, dav = 0.f;
          if (i < N && j < N) {
            const float D = sRn[i] + ((e < 2) ? cnA : cnB) - 2.f * st[n][e];
            const float t = __fdividef(1.0f, 1.0f + ex2(D * LOG2E));
            const float pv = ex2(fmaf(t, LOG2E, bias) - sLse[i]);
            const float keepf = ((kw[n][e & 1] >> (qrow + (e < 2 ? 0 : 8))) & 1u) ? inv_keep : 0.f;
            ptv = pv * keepf;
            const float dP = (g1[n][e] + 2.f * ptv * g2[n][e]) * keepf;
            dav = pv * (dP - sDelta[i]);
            ddv = -dav * t * (1.0f - t);
          }
          pt[n][e] = ptv; dd[n][e] = ddv; da[n][e] = dav;
          if (e < 2) kapA += ddv; else kapB += ddv;
        }
      }
      const uint32_t pa0 = pack_bf16x2(pt[0][0], pt[0][1]), pa1 = pack_bf16x2(pt[0][2], pt[0][3]);
      const uint32_t pa2 = pack_bf16x2(pt[1][0], pt[1][1]), pa3 = pack_bf16x2(pt[1][2], pt[1][3]);
      const uint32_t qa0 = pack_bf16x2(pt[0][0] * pt[0][0], pt[0][1] * pt[0][1]), qa1 = pack_bf16x2(pt[0][2] * pt[0][2], pt[0][3] * pt[0][3]);
      const uint32_t qa2 = pack_bf16x2(pt[1][0] * pt[1][0], pt[1][1] * pt[1][1]), qa3 = pack_bf16x2(pt[1][2] * pt[1][2], pt[1][3] * pt[1][3]);
      const uint32_t da0 = pack_bf16x2(dd[0][0], dd[0][1]), da1 = pack_bf16x2(dd[0][2], dd[0][3]);
      const uint32_t da2 = pack_bf16x2(dd[1][0], dd[1][1]), da3 = pack_bf16x2(dd[1][2], dd[1][3]);
      {  // dD^T (kernel 2) and dA^T (bias-table gradient) to global, [bh][key][query]
        bf16* d0 = p.dD + ((long long)bh * N + jA) * p.ld_ds + ia;
        bf16* d1 = p.dD + ((long long)bh * N + jB) * p.ld_ds + ia;
        if (jA < N) { *reinterpret_cast<uint32_t*>(d0) = da0; *reinterpret_cast<uint32_t*>(d0 + 8) = da2; }
        if (jB < N) { *reinterpret_cast<uint32_t*>(d1) = da1; *reinterpret_cast<uint32_t*>(d1 + 8) = da3; }
        if (p.dA != nullptr) {
          bf16* a0p = p.dA + ((long long)bh * N + jA) * p.ld_ds + ia;
          bf16* a1p = p.dA + ((long long)bh * N + jB) * p.ld_ds + ia;
          if (jA < N) { *reinterpret_cast<uint32_t*>(a0p) = pack_bf16x2(da[0][0], da[0][1]); *reinterpret_cast<uint32_t*>(a0p + 8) = pack_bf16x2(da[1][0], da[1][1]); }
          if (jB < N) { *reinterpret_cast<uint32_t*>(a1p) = pack_bf16x2(da[0][2], da[0][3]); *reinterpret_cast<uint32_t*>(a1p + 8) = pack_bf16x2(da[1][2], da[1][3]); }
        }
      }
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t b0, b1, b2, b3;
        const int off = (it * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + dpair * 16 + (lane >> 4) * 8;
        ldsm_x4_t(smem_u32(sdOm + off), b0, b1, b2, b3);
        mma16816(dv[dpair * 2], pa0, pa1, pa2, pa3, b0, b1);
        mma16816(dv[dpair * 2 + 1], pa0, pa1, pa2, pa3, b2, b3);
        ldsm_x4_t(smem_u32(sdOc + off), b0, b1, b2, b3);
        mma16816(dcv[dpair * 2], qa0, qa1, qa2, qa3, b0, b1);
        mma16816(dcv[dpair * 2 + 1], qa0, qa1, qa2, qa3, b2, b3);
      }
#pragma unroll
      for (int cp = 0; cp < 8; ++cp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(sX1 + (it * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * XP + cp * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
        mma16816(dx2[cp * 2], da0, da1, da2, da3, b0, b1);
        mma16816(dx2[cp * 2 + 1], da0, da1, da2, da3, b2, b3);
      }
    }
    // ---- tile epilogue
    kapA += __shfl_xor_sync(0xffffffffu, kapA, 1); kapA += __shfl_xor_sync(0xffffffffu, kapA, 2);
    kapB += __shfl_xor_sync(0xffffffffu, kapB, 1); kapB += __shfl_xor_sync(0xffffffffu, kapB, 2);
    bf16* gdkm = p.dqkv_m + (long long)b * N * row_stride + p.H * HD + h * HD;
    bf16* gdvm = gdkm + p.H * HD;
    bf16* gdkc = p.dqkv_c + (long long)b * N * row_stride + p.H * HD + h * HD;
    bf16* gdvc = gdkc + p.H * HD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + quad * 2;
      // values of this thread: rows jA (x,y = [0],[1]) and jB ([2],[3])
      const float2 cvA = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sCV + jA * PITCH + c));
      const float2 cvB = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sCV + jB * PITCH + c));
      const float cv0 = dcv[dt][0] * elu1_grad(cvA.x), cv1 = dcv[dt][1] * elu1_grad(cvA.y);
      const float cv2 = dcv[dt][2] * elu1_grad(cvB.x), cv3 = dcv[dt][3] * elu1_grad(cvB.y);
      const float2 m2A = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myX2 + qrow * XP + c));
      const float2 m2B = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myX2 + (qrow + 8) * XP + c));
      const float2 u2A = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myX2 + qrow * XP + HD + c));
      const float2 u2B = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myX2 + (qrow + 8) * XP + HD + c));
      auto dk_of = [](float kap, float m2, float acc) { return 2.f * (kap * m2 - acc) * m2 * (1.f - m2); };
      auto dck_of = [](float kap, float u2, float acc, float ckv) {
        const float c2 = u2 * u2;
        return (kap - acc / u2) * c2 * (1.f - c2) * (ckv > 1.0f ? 1.0f : ckv);
      };
      if (jA < N) {
        const float2 ck = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(gc + p.H * HD + (long long)jA * row_stride + c));
        *reinterpret_cast<uint32_t*>(gdvm + (long long)jA * row_stride + c) = pack_bf16x2(dv[dt][0], dv[dt][1]);
        *reinterpret_cast<uint32_t*>(gdvc + (long long)jA * row_stride + c) = pack_bf16x2(cv0, cv1);
        *reinterpret_cast<uint32_t*>(gdkm + (long long)jA * row_stride + c) = pack_bf16x2(dk_of(kapA, m2A.x, dx2[dt][0]), dk_of(kapA, m2A.y, dx2[dt][1]));
        *reinterpret_cast<uint32_t*>(gdkc + (long long)jA * row_stride + c) =
            pack_bf16x2(dck_of(kapA, u2A.x, dx2[8 + dt][0], ck.x), dck_of(kapA, u2A.y, dx2[8 + dt][1], ck.y));
      }
      if (jB < N) {
        const float2 ck = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(gc + p.H * HD + (long long)jB * row_stride + c));
        *reinterpret_cast<uint32_t*>(gdvm + (long long)jB * row_stride + c) = pack_bf16x2(dv[dt][2], dv[dt][3]);
        *reinterpret_cast<uint32_t*>(gdvc + (long long)jB * row_stride + c) = pack_bf16x2(cv2, cv3);
        *reinterpret_cast<uint32_t*>(gdkm + (long long)jB * row_stride + c) = pack_bf16x2(dk_of(kapB, m2B.x, dx2[dt][2]), dk_of(kapB, m2B.y, dx2[dt][3]));
        *reinterpret_cast<uint32_t*>(gdkc + (long long)jB * row_stride + c) =
            pack_bf16x2(dck_of(kapB, u2B.x, dx2[8 + dt][2], ck.x), dck_of(kapB, u2B.y, dx2[8 + dt][3], ck.y));
      }
      if (p.dv_bias != nullptr || p.dcv_bias != nullptr) {     // rows >= N contribute exact zeros
        float s0 = dv[dt][0] + dv[dt][2], s1 = dv[dt][1] + dv[dt][3];
        float z0 = (jA < N ? cv0 : 0.f) + (jB < N ? cv2 : 0.f), z1 = (jA < N ? cv1 : 0.f) + (jB < N ? cv3 : 0.f);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          z0 += __shfl_xor_sync(0xffffffffu, z0, o); z1 += __shfl_xor_sync(0xffffffffu, z1, o);
        }
        if (qrow == 0) {
          if (p.dv_bias != nullptr) { atomicAdd(p.dv_bias + h * HD + c, s0); atomicAdd(p.dv_bias + h * HD + c + 1, s1); }
          if (p.dcv_bias != nullptr) { atomicAdd(p.dcv_bias + h * HD + c, z0); atomicAdd(p.dcv_bias + h * HD + c + 1, z1); }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, kernel 2: query-tile owners. dX1_i = sum_j dD_ij [m2_j | u2_j] ; rho_i = sum_j dD_ij
//   dm1 = 2 (rho m1 - dX1_m) ; dc1 = rho - dX1_u / u1 ; dq = scale dm1 m1 (1-m1) ; dcq(pre-ELU) = dc1 c1 (1-c1) elu'
// ------------------------------------------------------------------------------------------------
constexpr int WB2_WARPS = 13;

__global__ void __launch_bounds__(WB2_WARPS * 32, 1) wattn_bwd_q_kernel(const WAttnBwdParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sdD = reinterpret_cast<bf16*>(smem);                    // [NMAX(key)][DSP(query)]
  bf16* sX2 = sdD + NMAX * DSP;
  float* sRho = reinterpret_cast<float*>(sX2 + NMAX * XP);
  float* sTmp = sRho + NMAX;
  const int bh = blockIdx.x;
  const int b = bh / p.H, h = bh - b * p.H;
  const int N = p.N;
  const int ntile = (N + 15) >> 4;
  const int n_pad = ntile * 16;
  const long long row_stride = 3LL * p.H * HD;
  const bf16* gm = p.qkv_m + (long long)b * N * row_stride + h * HD;
  const bf16* gc = p.qkv_c + (long long)b * N * row_stride + h * HD;
  {
    const int chunks = n_pad >> 3;
    for (int idx = threadIdx.x; idx < n_pad * chunks; idx += blockDim.x) {
      const int j = idx / chunks, c = (idx - j * chunks) * 8;
      if (j < N) cp_async16(smem_u32(sdD + j * DSP + c), p.dD + ((long long)bh * N + j) * p.ld_ds + c);
      else *reinterpret_cast<uint4*>(sdD + j * DSP + c) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  stage_mu(sX2, sTmp, gm + p.H * HD, gc + p.H * HD, row_stride, 1.0f, N, n_pad);
  cp_async_wait_all();
  __syncthreads();
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += __bfloat162float(sdD[j * DSP + i]);
    sRho[i] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane & 3, qrow = lane >> 2;
  if (warp >= ntile) return;
  const int it = warp;
  float dx1[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) dx1[i][0] = dx1[i][1] = dx1[i][2] = dx1[i][3] = 0.f;
#pragma unroll 1
  for (int jt = 0; jt < ntile; ++jt) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4_t(smem_u32(sdD + (jt * 16 + (lane >> 4) * 8 + (lane & 7)) * DSP + it * 16 + ((lane >> 3) & 1) * 8), a0, a1, a2, a3);
#pragma unroll
    for (int cp = 0; cp < 8; ++cp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sX2 + (jt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * XP + cp * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
      mma16816(dx1[cp * 2], a0, a1, a2, a3, b0, b1);
      mma16816(dx1[cp * 2 + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  const int iA = it * 16 + qrow, iB = iA + 8;
  const float rhoA = sRho[iA], rhoB = sRho[iB];
  bf16* gdqm = p.dqkv_m + (long long)b * N * row_stride + h * HD;
  bf16* gdqc = p.dqkv_c + (long long)b * N * row_stride + h * HD;
  auto dq_of = [&](float rho, float qv, float acc) {
    const float m1 = sigmoidf_(qv * p.scale);
    return p.scale * 2.f * (rho * m1 - acc) * m1 * (1.f - m1);
  };
  auto dcq_of = [](float rho, float cqv, float acc) {
    const float c1 = sigmoidf_(cqv);
    const float u1 = sqrtf(fmaxf(c1, 1e-24f));
    return (rho - acc / u1) * c1 * (1.f - c1) * (cqv > 1.0f ? 1.0f : cqv);
  };
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int c = dt * 8 + quad * 2;
    float m0 = 0.f, m1v = 0.f, m2 = 0.f, m3 = 0.f, c0 = 0.f, c1v = 0.f, c2 = 0.f, c3 = 0.f;
    if (iA < N) {
      const float2 q = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(gm + (long long)iA * row_stride + c));
      const float2 cq = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(gc + (long long)iA * row_stride + c));
      m0 = dq_of(rhoA, q.x, dx1[dt][0]); m1v = dq_of(rhoA, q.y, dx1[dt][1]);
      c0 = dcq_of(rhoA, cq.x, dx1[8 + dt][0]); c1v = dcq_of(rhoA, cq.y, dx1[8 + dt][1]);
      *reinterpret_cast<uint32_t*>(gdqm + (long long)iA * row_stride + c) = pack_bf16x2(m0, m1v);
      *reinterpret_cast<uint32_t*>(gdqc + (long long)iA * row_stride + c) = pack_bf16x2(c0, c1v);
    }
    if (iB < N) {
      const float2 q = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(gm + (long long)iB * row_stride + c));
      const float2 cq = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(gc + (long long)iB * row_stride + c));
      m2 = dq_of(rhoB, q.x, dx1[dt][2]); m3 = dq_of(rhoB, q.y, dx1[dt][3]);
      c2 = dcq_of(rhoB, cq.x, dx1[8 + dt][2]); c3 = dcq_of(rhoB, cq.y, dx1[8 + dt][3]);
      *reinterpret_cast<uint32_t*>(gdqm + (long long)iB * row_stride + c) = pack_bf16x2(m2, m3);
      *reinterpret_cast<uint32_t*>(gdqc + (long long)iB * row_stride + c) = pack_bf16x2(c2, c3);
    }
    if (p.dq_bias != nullptr || p.dcq_bias != nullptr) {
      float s0 = m0 + m2, s1 = m1v + m3, z0 = c0 + c2, z1 = c1v + c3;
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        z0 += __shfl_xor_sync(0xffffffffu, z0, o); z1 += __shfl_xor_sync(0xffffffffu, z1, o);
      }
      if (qrow == 0) {
        if (p.dq_bias != nullptr) { atomicAdd(p.dq_bias + h * HD + c, s0); atomicAdd(p.dq_bias + h * HD + c + 1, s1); }
        if (p.dcq_bias != nullptr) { atomicAdd(p.dcq_bias + h * HD + c, z0); atomicAdd(p.dcq_bias + h * HD + c + 1, z1); }
      }
    }
  }
}

constexpr size_t WB1_SMEM = NMAX * XP * sizeof(bf16) + 4 * NMAX * PITCH * sizeof(bf16) + WB1_WARPS * 16 * XP * sizeof(bf16) +
                            3 * NMAX * sizeof(float) + WB1_WARPS * 16 * sizeof(float);
constexpr size_t WB2_SMEM = NMAX * DSP * sizeof(bf16) + NMAX * XP * sizeof(bf16) + 2 * NMAX * sizeof(float);

constexpr size_t WFWD_SMEM = 2 * NMAX * XP * sizeof(bf16) + 2 * NMAX * PITCH * sizeof(bf16) + 2 * NMAX * sizeof(float);

template <bool DROP>
cudaError_t launch_wfwd(const WAttnFwdParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wattn_fwd_kernel<DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WFWD_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  wattn_fwd_kernel<DROP><<<p.B * p.H, WF_WARPS * 32, WFWD_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

extern "C" int b200vit_wattn_fwd(const void* qkv_mean, const void* qkv_cov, const float* bias, int64_t ld_bias, int32_t B, int32_t H, int32_t N,
                                 int32_t head_dim, float scale, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id,
                                 const uint8_t* keep_in, void* out_mean, void* out_cov, float* lse, uint8_t* keep_bits, void* stream) {
  B200_CHECK_ARG(qkv_mean && qkv_cov && out_mean && out_cov, "wattn_fwd: null pointer");
  B200_CHECK_ARG(head_dim == HD, "wattn_fwd: head_dim %d unsupported (64 only)", head_dim);
  B200_CHECK_ARG(N > 0 && N <= NMAX, "wattn_fwd: N=%d unsupported (1..%d)", N, NMAX);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || keep_bits != nullptr), "wattn_fwd: bad p_drop / missing keep_bits");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(bias != nullptr && ld_bias >= n_pad && ld_bias % 2 == 0 && (reinterpret_cast<uintptr_t>(bias) & 7) == 0,
                 "wattn_fwd: the dual-stream attention needs the shared relative position bias (padded layout, ld even >= %d); "
                 "the reference fails without it (modeling_finetune_dist.py:155)", n_pad);
  WAttnFwdParams p;
  p.qkv_m = static_cast<const bf16*>(qkv_mean); p.qkv_c = static_cast<const bf16*>(qkv_cov); p.bias = bias; p.ld_bias = ld_bias;
  p.out_m = static_cast<bf16*>(out_mean); p.out_c = static_cast<bf16*>(out_cov); p.lse = lse; p.keep_bits = keep_bits; p.keep_in = keep_in;
  p.B = B; p.H = H; p.N = N; p.scale = scale; p.p_drop = p_drop; p.seed = seed; p.seed_dev = seed_dev; p.stream_id = stream_id;
  cudaError_t e = p_drop > 0.f ? launch_wfwd<true>(p, static_cast<cudaStream_t>(stream)) : launch_wfwd<false>(p, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { b200vit_set_error("wattn_fwd: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int b200vit_wattn_bwd(const void* qkv_mean, const void* qkv_cov, const void* out_mean, const void* out_cov, const void* dout_mean,
                                 const void* dout_cov, const float* lse, const float* bias_t, int64_t ld_bias, const uint8_t* keep_bits,
                                 void* work_dD, void* work_dA, int32_t ld_ds, const int32_t* rel_index, float* dtable, float* dq_bias,
                                 float* dv_bias, float* dcq_bias, float* dcv_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim,
                                 float scale, float p_drop, void* dqkv_mean, void* dqkv_cov, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(qkv_mean && qkv_cov && out_mean && out_cov && dout_mean && dout_cov && lse && dqkv_mean && dqkv_cov && work_dD, "wattn_bwd: null pointer");
  B200_CHECK_ARG(head_dim == HD && N > 0 && N <= NMAX, "wattn_bwd: head_dim 64 and N <= %d only", NMAX);
  B200_CHECK_ARG(p_drop == 0.f || keep_bits != nullptr, "wattn_bwd: dropout needs keep_bits from the forward");
  const int n_pad = (N + 15) / 16 * 16;
  B200_CHECK_ARG(bias_t != nullptr && ld_bias >= n_pad && ld_bias % 2 == 0, "wattn_bwd: needs the transposed padded bias");
  B200_CHECK_ARG(ld_ds >= n_pad && ld_ds % 8 == 0 && (reinterpret_cast<uintptr_t>(work_dD) & 15) == 0, "wattn_bwd: workspace [B,H,N,ld_ds], ld_ds %% 8 == 0, >= %d", n_pad);
  B200_CHECK_ARG(dtable == nullptr || (rel_index != nullptr && work_dA != nullptr), "wattn_bwd: dtable needs rel_index and the dA workspace");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wattn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WB1_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wattn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WB2_SMEM);
    if (e != cudaSuccess) { b200vit_set_error("wattn_bwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  WAttnBwdParams p;
  p.qkv_m = static_cast<const bf16*>(qkv_mean); p.qkv_c = static_cast<const bf16*>(qkv_cov);
  p.out_m = static_cast<const bf16*>(out_mean); p.out_c = static_cast<const bf16*>(out_cov);
  p.dout_m = static_cast<const bf16*>(dout_mean); p.dout_c = static_cast<const bf16*>(dout_cov);
  p.lse = lse; p.bias_t = bias_t; p.ld_bias = ld_bias; p.keep_bits = keep_bits;
  p.dD = static_cast<bf16*>(work_dD); p.dA = dtable != nullptr ? static_cast<bf16*>(work_dA) : nullptr; p.ld_ds = ld_ds;
  p.dqkv_m = static_cast<bf16*>(dqkv_mean); p.dqkv_c = static_cast<bf16*>(dqkv_cov);
  p.dq_bias = dq_bias; p.dv_bias = dv_bias; p.dcq_bias = dcq_bias; p.dcv_bias = dcv_bias;
  p.B = B; p.H = H; p.N = N; p.scale = scale; p.p_drop = p_drop;
  wattn_bwd_kv_kernel<<<B * H, WB1_WARPS * 32, WB1_SMEM, st>>>(p);
  B200_CHECK_LAUNCH("wattn_bwd_kv");
  wattn_bwd_q_kernel<<<B * H, WB2_WARPS * 32, WB2_SMEM, st>>>(p);
  B200_CHECK_LAUNCH("wattn_bwd_q");
  if (dtable != nullptr) return b200vit_relbias_grad_launch(work_dA, B, H, N, ld_ds, rel_index, dtable, stream);
  return 0;
}
