// patch_transform of the pre-training data pipeline on the device: transforms.ToTensor() + transforms.Normalize(mean, std)
// (datasets.py:80-85, DataAugmentationForBEiT.__call__ :110-118). The host loader then hands over uint8 pixels (19 MB per 128-image batch
// instead of 77 MB of fp32) and the conversion rides on the copy stream.
//   ToTensor : HWC uint8 -> CHW, .to(float32).div(255)           Normalize: tensor.sub_(mean[c]).div_(std[c])
// Three rounded fp32 operations per element, in that order, so the result equals torchvision's bit for bit.
// HBM-bound: 1 byte read + 4 bytes written per element; one thread converts 4 consecutive pixels of a row (all channels) through an
// exact 256-entry table per channel.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

struct NormParams {
  float mean[4], stdv[4];
};

__device__ __forceinline__ float norm1(unsigned v, float m, float s) { return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), m), s); }

// A channel has only 256 possible inputs: each CTA first builds the exact 256-entry table per channel (the three rounded operations, once per
// value) in shared memory, so the per-element work is one table read instead of two IEEE divisions and the kernel is bandwidth-bound.
template <int C>
__global__ void __launch_bounds__(256) normalize_hwc_kernel(const uint8_t* __restrict__ src, long long npix4, int hw, NormParams p,
                                                            float* __restrict__ out) {
  __shared__ float lut[C][256];
#pragma unroll
  for (int c = 0; c < C; ++c) lut[c][threadIdx.x] = norm1(threadIdx.x, p.mean[c], p.stdv[c]);
  __syncthreads();
  // thread -> 4 consecutive pixels (hw % 4 == 0, so they belong to one image): 4*C contiguous bytes in, one float4 per channel plane out
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix4; i += stride) {
    const long long pix = i * 4;
    const long long b = pix / hw;
    const int q = (int)(pix - b * hw);
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + pix * C);   // pix % 4 == 0 -> byte offset % 4 == 0
    uint32_t w[C];
#pragma unroll
    for (int k = 0; k < C; ++k) w[k] = __ldg(s32 + k);
    float v[4][C];
#pragma unroll
    for (int j = 0; j < 4 * C; ++j) v[j / C][j % C] = lut[j % C][(w[j >> 2] >> (8 * (j & 3))) & 0xFFu];
#pragma unroll
    for (int c = 0; c < C; ++c)
      *reinterpret_cast<float4*>(out + (b * C + c) * hw + q) = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
  }
}

__global__ void __launch_bounds__(256) normalize_generic_kernel(const uint8_t* __restrict__ src, int hwc, long long total, int C, int hw,
                                                                NormParams p, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {   // i indexes the CHW output
    const long long b = i / ((long long)C * hw);
    const int r = (int)(i - b * C * hw);
    const int c = r / hw, q = r - c * hw;
    const long long s = hwc ? (b * hw + q) * C + c : i;
    out[i] = norm1(src[s], p.mean[c], p.stdv[c]);
  }
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_normalize_u8(const uint8_t* src, int32_t hwc, int32_t B, int32_t C, int32_t H, int32_t W, const float* mean_host,
                                    const float* std_host, float* out, void* stream) {
  B200_CHECK_ARG(src != nullptr && out != nullptr && mean_host != nullptr && std_host != nullptr, "normalize_u8: null pointer");
  B200_CHECK_ARG(B > 0 && C > 0 && C <= 4 && H > 0 && W > 0, "normalize_u8: bad shape (C <= 4)");
  NormParams p;
  for (int c = 0; c < 4; ++c) {
    p.mean[c] = c < C ? mean_host[c] : 0.f;
    p.stdv[c] = c < C ? std_host[c] : 1.f;
    B200_CHECK_ARG(p.stdv[c] != 0.f, "normalize_u8: std[%d] is zero", c);
  }
  const int hw = H * W;
  const long long total = (long long)B * C * hw;
  const long long cap = (long long)b200vit_num_sms() * 16;
  auto grid = [&](long long n) { long long g = (n + 255) / 256; return (int)(g < 1 ? 1 : (g > cap ? cap : g)); };
  const bool fast = hwc && C == 3 && hw % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (fast) {
    const long long npix4 = (long long)B * hw / 4;
    normalize_hwc_kernel<3><<<grid(npix4), 256, 0, STREAM>>>(src, npix4, hw, p, out);
  } else {
    normalize_generic_kernel<<<grid(total), 256, 0, STREAM>>>(src, hwc, total, C, hw, p, out);
  }
  B200_CHECK_LAUNCH("normalize_u8");
  return 0;
}
