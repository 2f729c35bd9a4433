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
                                 int32_t head_dim, float scale, float p_drop, uint64_t seed, uint32_t stream_id, const uint8_t* keep_in,
                                 void* out_mean, void* out_cov, float* lse, uint8_t* keep_bits, void* stream) {
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
  p.B = B; p.H = H; p.N = N; p.scale = scale; p.p_drop = p_drop; p.seed = seed; p.stream_id = stream_id;
  cudaError_t e = p_drop > 0.f ? launch_wfwd<true>(p, static_cast<cudaStream_t>(stream)) : launch_wfwd<false>(p, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { b200vit_set_error("wattn_fwd: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
