// Mixup / CutMix of a fine-tune batch on the device, "batch" mode (one lambda / one box per batch):
// timm.data.mixup.Mixup._mix_batch + mixup_target (timm 0.3.2, the version the reference pins in requirements.txt:3), called by
// engine_for_finetuning.py:87-88 / engine_for_finetuning_dist.py:356-357 with the Mixup built at run_class_finetuning.py:339-347.
//
//   mixup :  x[b] <- lam * x[b] + (1 - lam) * x[B-1-b]                      (x.mul_(lam).add_(x.flip(0).mul_(1 - lam)))
//   cutmix:  x[b][:, yl:yh, xl:xh] <- x[B-1-b][:, yl:yh, xl:xh]
//   target:  t[b] = lam * smooth_one_hot(y[b]) + (1 - lam) * smooth_one_hot(y[B-1-b])
//
// HBM-bound, in place: the pair (b, B-1-b) is handled by one thread so both originals are read before either is overwritten
// (8 B/element of traffic for mixup; cutmix touches only the box, one warp per box row). fp32 with the reference's roundings — two rounded products and one
// rounded sum, lam and 1 - lam rounded to fp32 first, no FMA contraction — so images and targets are bit-identical to the torch ops.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float mix2(float a, float c, float lam, float om) { return __fadd_rn(__fmul_rn(a, lam), __fmul_rn(c, om)); }

// one thread per float4 of the lower half of the batch, two independent pairs of 16-byte loads in flight per thread; chw4 = C*H*W / 4
__global__ void __launch_bounds__(256) mixup_pairs_kernel(float* __restrict__ x, int B, long long chw4, float lam, float om) {
  const long long total = (long long)(B / 2) * chw4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto mix4 = [&](const float4& a, const float4& c) {
    return make_float4(mix2(a.x, c.x, lam, om), mix2(a.y, c.y, lam, om), mix2(a.z, c.z, lam, om), mix2(a.w, c.w, lam, om));
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const long long j = i + stride;
    const long long b0 = i / chw4, e0 = i - b0 * chw4;
    float4* pa0 = reinterpret_cast<float4*>(x) + b0 * chw4 + e0;
    float4* pc0 = reinterpret_cast<float4*>(x) + (long long)(B - 1 - b0) * chw4 + e0;
    const bool two = j < total;
    const long long b1 = two ? j / chw4 : b0, e1 = two ? j - b1 * chw4 : e0;
    float4* pa1 = reinterpret_cast<float4*>(x) + b1 * chw4 + e1;
    float4* pc1 = reinterpret_cast<float4*>(x) + (long long)(B - 1 - b1) * chw4 + e1;
    const float4 a0 = *pa0, c0 = *pc0;
    float4 a1 = a0, c1 = c0;
    if (two) { a1 = *pa1; c1 = *pc1; }
    *pa0 = mix4(a0, c0);
    *pc0 = mix4(c0, a0);
    if (two) { *pa1 = mix4(a1, c1); *pc1 = mix4(c1, a1); }
  }
}

// odd batch: the middle image pairs with itself
__global__ void __launch_bounds__(256) mixup_self_kernel(float* __restrict__ x, long long n, float lam, float om) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = mix2(x[i], x[i], lam, om);
}

// cutmix: swap the box between the images of a pair; one warp per box row (pair, channel, y), lanes along x (coalesced, no per-element division)
__global__ void __launch_bounds__(256) cutmix_pairs_kernel(float* __restrict__ x, int B, int Cc, int H, int W, int yl, int yh, int xl, int xh) {
  const int bh = yh - yl;
  const long long rows = (long long)(B / 2) * Cc * bh;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const long long b = r / ((long long)Cc * bh);
    const int rem = (int)(r - b * Cc * bh);
    const int c = rem / bh, y = yl + rem % bh;
    const long long off = ((long long)c * H + y) * W;
    float* pa = x + b * (long long)Cc * H * W + off;
    float* pc = x + (long long)(B - 1 - b) * Cc * H * W + off;
    for (int xx = xl + lane; xx < xh; xx += 64) {
      const int x2 = xx + 32;
      const float a0 = pa[xx], c0 = pc[xx];
      float a1 = 0.f, c1 = 0.f;
      if (x2 < xh) { a1 = pa[x2]; c1 = pc[x2]; }
      pa[xx] = c0;
      pc[xx] = a0;
      if (x2 < xh) { pa[x2] = c1; pc[x2] = a1; }
    }
  }
}

__global__ void __launch_bounds__(256) mixup_target_kernel(const long long* __restrict__ labels, int B, int K, float lam, float om, float on,
                                                           float off, float* __restrict__ out) {
  const long long total = (long long)B * K;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int b = (int)(i / K), k = (int)(i - (long long)b * K);
    const float y1 = labels[b] == k ? on : off, y2 = labels[B - 1 - b] == k ? on : off;
    out[i] = mix2(y1, y2, lam, om);
  }
}

int grid_for(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)b200vit_num_sms() * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_mixup_batch(float* x, int32_t B, int32_t Cc, int32_t H, int32_t W, float lam, float one_minus_lam, int32_t use_cutmix,
                                   int32_t yl, int32_t yh, int32_t xl, int32_t xh, const int64_t* labels, int32_t K, float on_value,
                                   float off_value, float* soft_targets, void* stream) {
  B200_CHECK_ARG(B > 0 && Cc > 0 && H > 0 && W > 0, "mixup_batch: bad image shape");
  B200_CHECK_ARG((labels == nullptr) == (soft_targets == nullptr) && (labels == nullptr || K > 0), "mixup_batch: labels and soft_targets go together");
  if (x != nullptr && lam != 1.0f) {
    const long long chw = (long long)Cc * H * W;
    if (use_cutmix) {
      B200_CHECK_ARG(0 <= yl && yl <= yh && yh <= H && 0 <= xl && xl <= xh && xh <= W, "mixup_batch: box [%d:%d, %d:%d] outside %d x %d", yl, yh, xl, xh, H, W);
      const long long n = (long long)(B / 2) * Cc * (yh - yl) * (xh - xl);
      if (n > 0) {
        cutmix_pairs_kernel<<<grid_for((long long)(B / 2) * Cc * (yh - yl) * 32), 256, 0, STREAM>>>(x, B, Cc, H, W, yl, yh, xl, xh);
        B200_CHECK_LAUNCH("cutmix_pairs");
      }
    } else {
      B200_CHECK_ARG(chw % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "mixup_batch: C*H*W must be a multiple of 4 and x 16-byte aligned");
      if (B / 2 > 0) {
        mixup_pairs_kernel<<<grid_for(((long long)(B / 2) * (chw / 4) + 1) / 2), 256, 0, STREAM>>>(x, B, chw / 4, lam, one_minus_lam);
        B200_CHECK_LAUNCH("mixup_pairs");
      }
      if (B & 1) {
        mixup_self_kernel<<<grid_for(chw), 256, 0, STREAM>>>(x + (long long)(B / 2) * chw, chw, lam, one_minus_lam);
        B200_CHECK_LAUNCH("mixup_self");
      }
    }
  }
  if (labels != nullptr) {
    mixup_target_kernel<<<grid_for((long long)B * K), 256, 0, STREAM>>>(reinterpret_cast<const long long*>(labels), B, K, lam, one_minus_lam,
                                                                       on_value, off_value, soft_targets);
    B200_CHECK_LAUNCH("mixup_target");
  }
  return 0;
}
