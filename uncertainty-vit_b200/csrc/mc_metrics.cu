// MC-sample uncertainty reduction (evaluate_MC_dropout, uncertainty_evaluations.py:77-85; ECELoss :110-202; NLL :270-272):
// one CTA per evaluated image reduces its S x K logits to the mean logits, softmax confidence / prediction, top-1/top-5
// hits, NLL, and the sample-based extensions (predictive entropy, variance, mutual information); the 15-bin calibration
// histogram is accumulated with one atomic per image. A single-CTA finalize turns histogram + row stats into the summary.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

constexpr int MC_THREADS = 256;
constexpr int MC_KPT = 16;  // classes per thread -> K <= 4096

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  v = sh[0];
#pragma unroll
  for (int i = 1; i < MC_THREADS / 32; ++i) v = is_max ? fmaxf(v, sh[i]) : v + sh[i];
  return v;
}

__global__ void __launch_bounds__(MC_THREADS) mc_reduce_kernel(const float* __restrict__ logits, const int* __restrict__ labels, int S, int N,
                                                               int K, int n_bins, float* __restrict__ mean_logits,
                                                               float* __restrict__ row_stats, float* __restrict__ hist) {
  __shared__ float sh[MC_THREADS / 32];
  __shared__ int sh_arg[MC_THREADS / 32];
  __shared__ float sh_argv[MC_THREADS / 32];
  const int n = blockIdx.x;
  const int tid = threadIdx.x;
  float zsum[MC_KPT], psum[MC_KPT], p2sum[MC_KPT];
#pragma unroll
  for (int i = 0; i < MC_KPT; ++i) zsum[i] = psum[i] = p2sum[i] = 0.f;
  float ent_s = 0.f;  // sum_s H(p_s) partial
  for (int s = 0; s < S; ++s) {
    const float* z = logits + ((long long)s * N + n) * K;
    float zv[MC_KPT];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MC_KPT; ++i) {
      const int k = tid + i * MC_THREADS;
      zv[i] = k < K ? z[k] : -INFINITY;
      mx = fmaxf(mx, zv[i]);
      if (k < K) zsum[i] += zv[i];
    }
    mx = block_reduce(mx, true, sh);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < MC_KPT; ++i) se += (tid + i * MC_THREADS < K) ? __expf(zv[i] - mx) : 0.f;
    se = block_reduce(se, false, sh);
    const float lse = logf(se);
    const float inv = 1.0f / se;
#pragma unroll
    for (int i = 0; i < MC_KPT; ++i) {
      if (tid + i * MC_THREADS < K) {
        const float pr = __expf(zv[i] - mx) * inv;
        psum[i] += pr;
        p2sum[i] += pr * pr;
        ent_s -= pr * (zv[i] - mx - lse);
      }
    }
  }
  const float invS = 1.0f / S;
  // mean logits, their softmax
  float mx = -INFINITY;
  int arg = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < MC_KPT; ++i) {
    const int k = tid + i * MC_THREADS;
    if (k < K) {
      zsum[i] *= invS;
      mean_logits[(long long)n * K + k] = zsum[i];
      if (zsum[i] > mx) { mx = zsum[i]; arg = k; }   // first maximum within the thread (k ascending)
    }
  }
  // block arg-max with lowest-index tie-break (np.argmax / torch.max semantics)
  float bv = mx; int bi = arg;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  __syncthreads();
  if ((tid & 31) == 0) { sh_argv[tid >> 5] = bv; sh_arg[tid >> 5] = bi; }
  __syncthreads();
  bv = sh_argv[0]; bi = sh_arg[0];
#pragma unroll
  for (int i = 1; i < MC_THREADS / 32; ++i)
    if (sh_argv[i] > bv || (sh_argv[i] == bv && sh_arg[i] < bi)) { bv = sh_argv[i]; bi = sh_arg[i]; }
  const float zmax = bv;
  const int pred = bi;
  const int label = labels[n];
  __syncthreads();  // mean_logits row written above by this CTA
  const float zl = __ldcg(mean_logits + (long long)n * K + ((label >= 0 && label < K) ? label : 0));
  float se = 0.f, rank = 0.f, hbar = 0.f, var = 0.f;
#pragma unroll
  for (int i = 0; i < MC_KPT; ++i) {
    const int k = tid + i * MC_THREADS;
    if (k < K) {
      se += __expf(zsum[i] - zmax);
      rank += zsum[i] > zl ? 1.f : 0.f;
      const float pb = psum[i] * invS;
      hbar -= pb > 0.f ? pb * logf(pb) : 0.f;
      var += p2sum[i] * invS - pb * pb;
    }
  }
  se = block_reduce(se, false, sh);
  rank = block_reduce(rank, false, sh);
  hbar = block_reduce(hbar, false, sh);
  var = block_reduce(var, false, sh);
  ent_s = block_reduce(ent_s, false, sh);
  if (tid == 0) {
    const float conf = 1.0f / se;                 // max softmax probability of the mean logits
    const float c1 = pred == label ? 1.f : 0.f;
    const float c5 = rank < 5.f ? 1.f : 0.f;
    const float nll = -(zl - zmax - logf(se));
    float* rs = row_stats + (long long)n * 8;
    rs[0] = conf; rs[1] = (float)pred; rs[2] = c1; rs[3] = c5; rs[4] = nll; rs[5] = hbar; rs[6] = var; rs[7] = hbar - ent_s * invS;
    // bins (lo, hi] on np.linspace(0, 1, n_bins + 1)  (uncertainty_evaluations.py:112-117,173-176)
    const double step = 1.0 / n_bins;
    const double c = (double)conf;
    for (int b = 0; b < n_bins; ++b) {
      const double lo = b * step, hi = (b + 1 == n_bins) ? 1.0 : (b + 1) * step;
      if (c > lo && c <= hi) {
        atomicAdd(hist + b * 3, 1.0f);
        atomicAdd(hist + b * 3 + 1, conf);
        atomicAdd(hist + b * 3 + 2, c1);
        break;
      }
    }
  }
}

__global__ void __launch_bounds__(1024) mc_finalize_kernel(const float* __restrict__ row_stats, const float* __restrict__ hist, int N, int n_bins,
                                                           float* __restrict__ summary) {
  __shared__ double sh[32][6];
  double a[6] = {0, 0, 0, 0, 0, 0};  // c1, c5, nll, ent, var, mi
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* rs = row_stats + (long long)n * 8;
    a[0] += rs[2]; a[1] += rs[3]; a[2] += rs[4]; a[3] += rs[5]; a[4] += rs[6]; a[5] += rs[7];
  }
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
  if ((threadIdx.x & 31) == 0)
    for (int j = 0; j < 6; ++j) sh[threadIdx.x >> 5][j] = a[j];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[6] = {0, 0, 0, 0, 0, 0};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
      for (int j = 0; j < 6; ++j) t[j] += sh[w][j];
    double ece = 0.0, ece_ref = 0.0;
    const double acc0 = N > 0 ? row_stats[2] : 0.0, acc1v = N > 1 ? row_stats[8 + 2] : 0.0;
    for (int b = 0; b < n_bins; ++b) {
      const double cnt = hist[b * 3];
      if (cnt > 0) {
        const double prop = cnt / N, conf = hist[b * 3 + 1] / cnt, acc = hist[b * 3 + 2] / cnt;
        ece += prop * fabs(conf - acc);
        // what the reference computes (uint8 fancy-indexing quirk, see oracle/vit_oracle.py::ece)
        ece_ref += prop * fabs(conf - ((N - cnt) * acc0 + cnt * acc1v) / N);
      }
    }
    summary[0] = (float)(100.0 * t[0] / N); summary[1] = (float)(100.0 * t[1] / N);
    summary[2] = (float)ece; summary[3] = (float)ece_ref; summary[4] = (float)(t[2] / N);
    summary[5] = (float)(t[3] / N); summary[6] = (float)(t[4] / N); summary[7] = (float)(t[5] / N);
  }
}

}  // namespace

extern "C" int b200vit_mc_reduce(const float* logits, const int32_t* labels, int32_t S, int32_t N, int32_t K, int32_t n_bins,
                                 float* mean_logits, float* row_stats, float* hist, void* stream) {
  B200_CHECK_ARG(logits && labels && mean_logits && row_stats && hist, "mc_reduce: null pointer");
  B200_CHECK_ARG(S > 0 && N > 0 && K > 0 && K <= MC_THREADS * MC_KPT && n_bins > 0, "mc_reduce: bad shape S=%d N=%d K=%d (K <= %d)", S, N, K, MC_THREADS * MC_KPT);
  mc_reduce_kernel<<<N, MC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(logits, labels, S, N, K, n_bins, mean_logits, row_stats, hist);
  B200_CHECK_LAUNCH("mc_reduce");
  return 0;
}

extern "C" int b200vit_mc_finalize(const float* row_stats, const float* hist, int32_t N, int32_t n_bins, float* summary, void* stream) {
  B200_CHECK_ARG(row_stats && hist && summary && N > 0 && n_bins > 0, "mc_finalize: bad arguments");
  mc_finalize_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(row_stats, hist, N, n_bins, summary);
  B200_CHECK_LAUNCH("mc_finalize");
  return 0;
}
