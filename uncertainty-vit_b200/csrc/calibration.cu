// Class-wise calibration metrics of the evaluation printout (evaluate_MC_dropout, uncertainty_evaluations.py:82-85; evaluate,
// engine_for_finetuning.py:195-213):
//   TACE  : TACELoss.loss (uncertainty_evaluations.py:241-261) — softmax probabilities below `threshold` set to 0, then per class n_bins
//           ADAPTIVE bins whose lower edges are the order statistics sorted_p[i * (N / n_bins)] (compute_bin_boundaries, :119-132),
//           sum_bins prop * |mean conf - mean acc| (compute_bins, :159-186), averaged over the classes.
//           Like ECELoss, compute_bins indexes a numpy array with a uint8 array (:173-184), so what the reference prints uses
//           bin_acc = ((N - n_b) * acc[0] + n_b * acc[1]) / N; both the documented value and that one are returned.
//   AUROC : torchmetrics AUROC(task="multiclass") (:49,85; macro average of the one-vs-rest areas, exact thresholds, ties by the trapezoid =
//           Mann-Whitney U with half credit for ties; a class without positives or without negatives scores 0). torchmetrics is not in
//           the image: restated from its published algorithm, parity unpinned against the package itself.
// One CTA per class sorts that class's N probabilities (bitonic network, in shared memory when they fit), derives the bin edges, bins every
// sample, and ranks the positives by binary search in the sorted array.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

constexpr int CAL_THREADS = 1024;
constexpr int MAX_BINS = 64;
constexpr int POS_CAP = 4096;
constexpr int SMEM_SORT_MAX = 32768;   // floats sorted in shared memory (128 KB)

__device__ __forceinline__ float prob_of(float z, float mx, float inv, int is_prob) {
  return is_prob ? z : __fmul_rn(__expf(__fsub_rn(z, mx)), inv);
}

// {max, 1 / sum exp(z - max)} per row; one warp per row
__global__ void __launch_bounds__(256) cal_row_stats_kernel(const float* __restrict__ logits, int N, int K, float2* __restrict__ rs) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* z = logits + (long long)n * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[k]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += __expf(__fsub_rn(z[k], mx));
  s = warp_sum(s);
  if (lane == 0) rs[n] = make_float2(mx, 1.0f / s);
}

// probabilities, transposed: pt[k][n] (n < N), +inf in the padding n in [N, Npad). 32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) cal_transpose_kernel(const float* __restrict__ logits, const float2* __restrict__ rs, int N, int K, int Npad,
                                                            int is_prob, float* __restrict__ pt) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int n = n0 + j, k = k0 + tx;
    float v = INFINITY;
    if (n < N && k < K) {
      const float2 st = is_prob ? make_float2(0.f, 1.f) : rs[n];
      v = prob_of(logits[(long long)n * K + k], st.x, st.y, is_prob);
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int k = k0 + j, n = n0 + tx;
    if (k < K && n < Npad) pt[(long long)k * Npad + n] = tile[tx][j];
  }
}

__global__ void __launch_bounds__(CAL_THREADS) cal_class_kernel(const float* __restrict__ logits, const float2* __restrict__ rs,
                                                                const int* __restrict__ labels, int N, int K, int Npad, int is_prob,
                                                                float threshold, int n_bins, float* __restrict__ pt, int use_smem,
                                                                double* __restrict__ per_class /* [K][3] */) {
  extern __shared__ float dyn[];
  __shared__ float lower[MAX_BINS + 1];
  __shared__ double cnt[MAX_BINS], sconf[MAX_BINS], sacc[MAX_BINS];
  __shared__ float pos[POS_CAP];
  __shared__ int npos_sh;
  __shared__ double red[CAL_THREADS / 32];
  const int k = blockIdx.x, tid = threadIdx.x;
  float* g = pt + (long long)k * Npad;
  float* v = use_smem ? dyn : g;
  if (use_smem)
    for (int i = tid; i < Npad; i += CAL_THREADS) v[i] = g[i];
  if (tid < MAX_BINS) { cnt[tid] = 0.0; sconf[tid] = 0.0; sacc[tid] = 0.0; }
  if (tid == 0) npos_sh = 0;
  __syncthreads();
  // ---- bitonic sort, ascending (+inf padding ends up behind the data)
  for (int kk = 2; kk <= Npad; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < Npad; i += CAL_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const float a = v[i], b = v[l];
          const bool asc = (i & kk) == 0;
          if ((a > b) == asc) { v[i] = b; v[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  // ---- adaptive bin edges (compute_bin_boundaries, :119-132): lower_i = thresholded sorted[i * bin_n], upper_last = 1.0
  const int bin_n = N / n_bins;
  if (tid <= n_bins) {
    float e = 1.0f;
    if (tid < n_bins) {
      e = v[tid * bin_n];
      if (e < threshold) e = 0.f;
    }
    lower[tid] = e;
  }
  __syncthreads();
  // ---- bin every sample of this class column (compute_bins, :159-186) and collect the positives for the AUROC
  for (int n = tid; n < N; n += CAL_THREADS) {
    const float2 st = is_prob ? make_float2(0.f, 1.f) : rs[n];
    const float p = prob_of(logits[(long long)n * K + k], st.x, st.y, is_prob);
    const float ptv = p < threshold ? 0.f : p;
    const bool is_pos = labels[n] == k;
    for (int i = 0; i < n_bins; ++i) {
      if (ptv > lower[i] && ptv <= lower[i + 1]) {
        atomicAdd(&cnt[i], 1.0);
        atomicAdd(&sconf[i], (double)ptv);
        if (is_pos) atomicAdd(&sacc[i], 1.0);
        break;
      }
    }
    if (is_pos) {
      const int idx = atomicAdd(&npos_sh, 1);
      if (idx < POS_CAP) pos[idx] = p;
    }
  }
  __syncthreads();
  const int npos = npos_sh;
  // ---- AUROC of class k vs rest: U = sum over positives of (#negatives below) + 0.5 (#negatives tied)
  double u = 0.0;
  auto rank_one = [&](float s) {
    int lo = 0, hi = N;                  // lower_bound in v[0, N)
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] < s) lo = mid + 1; else hi = mid; }
    const int less_all = lo;
    hi = N;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] <= s) lo = mid + 1; else hi = mid; }
    return make_int2(less_all, lo - less_all);   // {#all < s, #all == s}
  };
  if (npos <= POS_CAP) {
    for (int a = tid; a < npos; a += CAL_THREADS) {
      const float s = pos[a];
      const int2 all = rank_one(s);
      int pl = 0, pe = 0;
      for (int b = 0; b < npos; ++b) { pl += pos[b] < s; pe += pos[b] == s; }
      u += (double)(all.x - pl) + 0.5 * (double)(all.y - pe);
    }
  } else {
    // more positives than the shared list holds: rank against the label array directly
    for (int n = tid; n < N; n += CAL_THREADS) {
      if (labels[n] != k) continue;
      const float2 st = is_prob ? make_float2(0.f, 1.f) : rs[n];
      const float s = prob_of(logits[(long long)n * K + k], st.x, st.y, is_prob);
      const int2 all = rank_one(s);
      int pl = 0, pe = 0;
      for (int m = 0; m < N; ++m) {
        if (labels[m] != k) continue;
        const float2 sm = is_prob ? make_float2(0.f, 1.f) : rs[m];
        const float q = prob_of(logits[(long long)m * K + k], sm.x, sm.y, is_prob);
        pl += q < s; pe += q == s;
      }
      u += (double)(all.x - pl) + 0.5 * (double)(all.y - pe);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
  if ((tid & 31) == 0) red[tid >> 5] = u;
  __syncthreads();
  if (tid == 0) {
    double ut = 0.0;
    for (int w = 0; w < CAL_THREADS / 32; ++w) ut += red[w];
    const double nneg = (double)(N - npos);
    const double auc = (npos > 0 && nneg > 0) ? ut / ((double)npos * nneg) : 0.0;
    const double a0 = N > 0 && labels[0] == k ? 1.0 : 0.0, a1 = N > 1 && labels[1] == k ? 1.0 : 0.0;
    double tace = 0.0, tace_ref = 0.0;
    for (int i = 0; i < n_bins; ++i) {
      if (cnt[i] > 0.0) {
        const double prop = cnt[i] / N, conf = sconf[i] / cnt[i];
        tace += prop * fabs(conf - sacc[i] / cnt[i]);
        tace_ref += prop * fabs(conf - ((N - cnt[i]) * a0 + cnt[i] * a1) / N);
      }
    }
    per_class[k * 3] = tace; per_class[k * 3 + 1] = tace_ref; per_class[k * 3 + 2] = auc;
  }
}

__global__ void __launch_bounds__(1024) cal_finalize_kernel(const double* __restrict__ per_class, int K, float* __restrict__ out) {
  __shared__ double sh[32][3];
  double a[3] = {0, 0, 0};
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    for (int j = 0; j < 3; ++j) a[j] += per_class[k * 3 + j];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
  if ((threadIdx.x & 31) == 0)
    for (int j = 0; j < 3; ++j) sh[threadIdx.x >> 5][j] = a[j];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[3] = {0, 0, 0};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
      for (int j = 0; j < 3; ++j) t[j] += sh[w][j];
    out[0] = (float)(t[0] / K); out[1] = (float)(t[1] / K); out[2] = (float)(t[2] / K);
  }
}

int next_pow2(int n) {
  int p = 32;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace

extern "C" size_t b200vit_tace_auroc_workspace_bytes(int32_t N, int32_t K) {
  if (N <= 0 || K <= 0) return 0;
  const size_t npad = (size_t)next_pow2(N);
  return (size_t)N * sizeof(float2) + (size_t)K * npad * sizeof(float) + (size_t)K * 3 * sizeof(double) + 256;
}

extern "C" int b200vit_tace_auroc(const float* logits, int32_t is_prob, const int32_t* labels, int32_t N, int32_t K, float threshold,
                                  int32_t n_bins, void* work, float* out, void* stream) {
  B200_CHECK_ARG(logits != nullptr && labels != nullptr && work != nullptr && out != nullptr, "tace_auroc: null pointer");
  B200_CHECK_ARG(N > 0 && K > 0 && n_bins > 0 && n_bins <= MAX_BINS && N >= n_bins, "tace_auroc: bad shape N=%d K=%d n_bins=%d (N >= n_bins, n_bins <= %d)", N, K,
                 n_bins, MAX_BINS);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(work) & 255) == 0, "tace_auroc: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int npad = next_pow2(N);
  char* w = static_cast<char*>(work);
  double* per_class = reinterpret_cast<double*>(w);
  size_t off = ((size_t)K * 3 * sizeof(double) + 255) / 256 * 256;
  float2* rs = reinterpret_cast<float2*>(w + off);
  off += ((size_t)N * sizeof(float2) + 255) / 256 * 256;
  float* pt = reinterpret_cast<float*>(w + off);
  if (!is_prob) {
    cal_row_stats_kernel<<<(N + 7) / 8, 256, 0, st>>>(logits, N, K, rs);
    B200_CHECK_LAUNCH("tace_row_stats");
  }
  cal_transpose_kernel<<<dim3((npad + 31) / 32, (K + 31) / 32), 256, 0, st>>>(logits, rs, N, K, npad, is_prob, pt);
  B200_CHECK_LAUNCH("tace_transpose");
  const int use_smem = npad <= SMEM_SORT_MAX;
  const size_t dyn = use_smem ? (size_t)npad * sizeof(float) : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(cal_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SORT_MAX * (int)sizeof(float));
    attr_set = true;
  }
  cal_class_kernel<<<K, CAL_THREADS, dyn, st>>>(logits, rs, labels, N, K, npad, is_prob, threshold, n_bins, pt, use_smem, per_class);
  B200_CHECK_LAUNCH("tace_class");
  cal_finalize_kernel<<<1, 1024, 0, st>>>(per_class, K, out);
  B200_CHECK_LAUNCH("tace_finalize");
  return 0;
}
