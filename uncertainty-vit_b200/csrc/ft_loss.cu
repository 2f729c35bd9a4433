// Fine-tune criterion on the device (train_class_batch, engine_for_finetuning_dist.py:286-304; engine_for_finetuning.py:30-43):
//   loss = SoftTargetCrossEntropy / LabelSmoothingCrossEntropy(logits, targets)  [ + WassersteinLossFineTuning(anchor, positive, negative) ]
// WassersteinLossFineTuning.forward (distloss.py:39-70): every input through a sigmoid; d(a,b) = |m_a - m_b|^2 + |sqrt(c_a) - sqrt(c_b)|^2;
// pos / neg / pos-vs-neg distance vectors each divided by their max-abs; triplet = -log sigmoid(neg - pos + 1e-24) / max * lambda_ft, summed;
// margin = clamp(pos - pvn, 0) / max * lambda_pvn, summed. Gradients w.r.t. the ANCHOR features only (the positive / negative forwards come
// from a deep copy of the model, :293-296) and through all four max normalisers, as autograd does in the reference.
// Three launches: rows (one warp per sample), finalize (one CTA), feature gradients (element-wise).
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sqc(float s) { return sqrtf(fmaxf(s, 1e-24f)); }

// work layout (floats): ce[B] | pos[B] | neg[B] | pvn[B] | cpos[B] | cneg[B] | stats[8]
__global__ void __launch_bounds__(256) ft_rows_kernel(const float* __restrict__ logits, long long ldl, const float* __restrict__ targets, int B, int K,
                                                      const float* __restrict__ fm, const float* __restrict__ fc, const float* __restrict__ pm,
                                                      const float* __restrict__ pc, const float* __restrict__ nm, const float* __restrict__ nc,
                                                      int C, float grad_scale, float* __restrict__ work, float* __restrict__ dlogits, long long ldd,
                                                      bf16* __restrict__ dlogits_bf16, long long ldd16, int Kp) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* z = logits + (long long)b * ldl;
  const float* t = targets + (long long)b * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[k]);
  mx = warp_max(mx);
  float se = 0.f, st = 0.f, stz = 0.f;
  for (int k = lane; k < K; k += 32) {
    se += __expf(z[k] - mx);
    st += t[k];
    stz += t[k] * (z[k] - mx);
  }
  se = warp_sum(se); st = warp_sum(st); stz = warp_sum(stz);
  const float lse = logf(se);
  if (lane == 0) work[b] = st * lse - stz;                      // -sum_k t_k log_softmax(z)_k
  const float inv = 1.0f / se, gs = grad_scale / B;
  for (int k = lane; k < Kp; k += 32) {
    const float g = k < K ? (__expf(z[k] - mx) * inv * st - t[k]) * gs : 0.f;
    if (dlogits != nullptr && k < K) dlogits[(long long)b * ldd + k] = g;
    if (dlogits_bf16 != nullptr) dlogits_bf16[(long long)b * ldd16 + k] = __float2bfloat16(g);
  }
  if (fm == nullptr) return;
  float dp = 0.f, dn = 0.f, dv = 0.f;
  for (int c = lane; c < C; c += 32) {
    const long long i = (long long)b * C + c;
    const float am = sigm(fm[i]), ac = sqc(sigm(fc[i]));
    const float qm = sigm(pm[i]), qc = sqc(sigm(pc[i]));
    const float rm = sigm(nm[i]), rc = sqc(sigm(nc[i]));
    dp += (am - qm) * (am - qm) + (ac - qc) * (ac - qc);
    dn += (am - rm) * (am - rm) + (ac - rc) * (ac - rc);
    dv += (qm - rm) * (qm - rm) + (qc - rc) * (qc - rc);
  }
  dp = warp_sum(dp); dn = warp_sum(dn); dv = warp_sum(dv);
  if (lane == 0) { work[B + b] = dp; work[2 * B + b] = dn; work[3 * B + b] = dv; }
}

__global__ void __launch_bounds__(1024) ft_finalize_kernel(float* __restrict__ work, int B, int has_triplet, float lam_ft, float lam_pvn,
                                                           float grad_scale, float* __restrict__ loss_out) {
  __shared__ float sh_v[32];
  __shared__ int sh_i[32];
  __shared__ float bcast[8];
  __shared__ int bidx[8];
  const int tid = threadIdx.x;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) sh_v[tid >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh_v[w];
    __syncthreads();
    return r;
  };
  auto block_argmax = [&](float v, int idx, int slot) {      // max of |.| with lowest-index tie-break -> bcast[slot], bidx[slot]
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { sh_v[tid >> 5] = v; sh_i[tid >> 5] = idx; }
    __syncthreads();
    if (tid == 0) {
      float bv = sh_v[0]; int bi = sh_i[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (sh_v[w] > bv || (sh_v[w] == bv && sh_i[w] < bi)) { bv = sh_v[w]; bi = sh_i[w]; }
      bcast[slot] = bv; bidx[slot] = bi;
    }
    __syncthreads();
  };
  float s = 0.f;
  for (int b = tid; b < B; b += blockDim.x) s += work[b];
  const float ce = block_sum(s) / B;
  float wl = 0.f;
  if (has_triplet) {
    const float* pos = work + B; const float* neg = work + 2 * B; const float* pvn = work + 3 * B;
    float* cpos = work + 4 * B; float* cneg = work + 5 * B;
    float v[3] = {-1.f, -1.f, -1.f}; int ix[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int b = tid; b < B; b += blockDim.x) {
      const float a[3] = {fabsf(pos[b]), fabsf(neg[b]), fabsf(pvn[b])};
      for (int j = 0; j < 3; ++j) if (a[j] > v[j]) { v[j] = a[j]; ix[j] = b; }
    }
    for (int j = 0; j < 3; ++j) block_argmax(v[j], ix[j], j);
    const float P = bcast[0], Nn = bcast[1], V = bcast[2];
    const int ip = bidx[0], in_ = bidx[1];
    // t_b = softplus(-(w - u + 1e-24)), g_b = clamp(u - x, 0)
    float tv = -1.f, gv = -1.f; int ti = 0x7fffffff, gi = 0x7fffffff;
    float st = 0.f, sg = 0.f;
    for (int b = tid; b < B; b += blockDim.x) {
      const float u = pos[b] / P, w = neg[b] / Nn, x = pvn[b] / V;
      const float t = -logf(sigm(w - u + 1e-24f));
      const float g = fmaxf(u - x, 0.f);
      st += t; sg += g;
      if (fabsf(t) > tv) { tv = fabsf(t); ti = b; }
      if (fabsf(g) > gv) { gv = fabsf(g); gi = b; }
    }
    block_argmax(tv, ti, 3);
    const float Tm = bcast[3]; const int it = bidx[3];
    block_argmax(gv, gi, 4);
    const float Gm = bcast[4]; const int igi = bidx[4];
    const float St = block_sum(st), Sg = block_sum(sg);
    wl = lam_ft * St / Tm + lam_pvn * Sg / Gm;
    // dL/du_b, dL/dw_b and their projections through u = pos / P, w = neg / Nn
    float pu = 0.f, pw = 0.f;
    for (int b = tid; b < B; b += blockDim.x) {
      const float u = pos[b] / P, w = neg[b] / Nn, x = pvn[b] / V;
      const float sgm = sigm(-(w - u + 1e-24f));
      const float dLt = lam_ft * (1.0f / Tm - (b == it ? St / (Tm * Tm) : 0.f));
      const float dLg = lam_pvn * (1.0f / Gm - (b == igi ? Sg / (Gm * Gm) : 0.f));
      const float du = dLt * sgm + (u - x > 0.f ? dLg : 0.f);
      const float dw = -dLt * sgm;
      cpos[b] = du / P * grad_scale;
      cneg[b] = dw / Nn * grad_scale;
      pu += du * pos[b];
      pw += dw * neg[b];
    }
    const float Pu = block_sum(pu), Pw = block_sum(pw);
    if (tid == 0) {
      cpos[ip] -= Pu / (P * P) * grad_scale;
      cneg[in_] -= Pw / (Nn * Nn) * grad_scale;
    }
  }
  if (tid == 0) { loss_out[0] = ce + wl; loss_out[1] = ce; loss_out[2] = wl; }
}

__global__ void __launch_bounds__(256) ft_bwd_kernel(const float* __restrict__ fm, const float* __restrict__ fc, const float* __restrict__ pm,
                                                     const float* __restrict__ pc, const float* __restrict__ nm, const float* __restrict__ nc,
                                                     const float* __restrict__ cpos, const float* __restrict__ cneg, int B, int C,
                                                     float* __restrict__ dfm, float* __restrict__ dfc) {
  const long long total = (long long)B * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int b = (int)(i / C);
    const float kp = cpos[b], kn = cneg[b];
    const float sa = sigm(fm[i]), sb = sigm(fc[i]);
    const float ua = sqc(sb);
    const float qm = sigm(pm[i]), qc = sqc(sigm(pc[i])), rm = sigm(nm[i]), rc = sqc(sigm(nc[i]));
    dfm[i] = (kp * 2.0f * (sa - qm) + kn * 2.0f * (sa - rm)) * sa * (1.0f - sa);
    dfc[i] = (kp * 2.0f * (ua - qc) + kn * 2.0f * (ua - rc)) * (sb > 1e-24f ? 0.5f / ua : 0.f) * sb * (1.0f - sb);
  }
}

}  // namespace

extern "C" size_t b200vit_finetune_loss_workspace_floats(int32_t B) { return (size_t)6 * (size_t)(B > 0 ? B : 0) + 8; }

extern "C" int b200vit_finetune_loss(const float* logits, int64_t ld_logits, const float* targets, int32_t B, int32_t K, const float* mean_feat,
                                     const float* cov_feat, const float* pos_mean, const float* pos_cov, const float* neg_mean,
                                     const float* neg_cov, int32_t C, float lambda_finetuning, float lambda_pvn, float grad_scale, float* work,
                                     float* dlogits, int64_t ld_dlogits, void* dlogits_bf16, int64_t ld_dlogits_bf16, int32_t K_padded,
                                     float* d_mean_feat, float* d_cov_feat, float* loss_out, void* stream) {
  B200_CHECK_ARG(logits != nullptr && targets != nullptr && work != nullptr && loss_out != nullptr, "finetune_loss: null pointer");
  B200_CHECK_ARG(B > 0 && K > 0 && ld_logits >= K, "finetune_loss: bad shape B=%d K=%d", B, K);
  const bool trip = mean_feat != nullptr;
  B200_CHECK_ARG(!trip || (cov_feat && pos_mean && pos_cov && neg_mean && neg_cov && d_mean_feat && d_cov_feat && C > 0),
                 "finetune_loss: the triplet term needs all six feature tensors and both gradient outputs");
  B200_CHECK_ARG(dlogits_bf16 == nullptr || (K_padded >= K && ld_dlogits_bf16 >= K_padded), "finetune_loss: bad padded logits-gradient shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Kp = dlogits_bf16 != nullptr ? K_padded : K;
  ft_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, ld_logits, targets, B, K, mean_feat, cov_feat, pos_mean, pos_cov, neg_mean, neg_cov, C, grad_scale,
                                              work, dlogits, ld_dlogits, static_cast<bf16*>(dlogits_bf16), ld_dlogits_bf16, Kp);
  B200_CHECK_LAUNCH("finetune_loss_rows");
  ft_finalize_kernel<<<1, 1024, 0, st>>>(work, B, trip ? 1 : 0, lambda_finetuning, lambda_pvn, grad_scale, loss_out);
  B200_CHECK_LAUNCH("finetune_loss_finalize");
  if (trip) {
    long long blocks = ((long long)B * C + 255) / 256;
    const long long cap = (long long)b200vit_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    ft_bwd_kernel<<<(int)blocks, 256, 0, st>>>(mean_feat, cov_feat, pos_mean, pos_cov, neg_mean, neg_cov, work + 4 * B, work + 5 * B, B, C, d_mean_feat,
                                               d_cov_feat);
    B200_CHECK_LAUNCH("finetune_loss_bwd");
  }
  return 0;
}
