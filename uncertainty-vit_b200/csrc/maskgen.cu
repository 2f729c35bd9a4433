// Block-wise mask generator of the pre-training input pipeline on the device.
//
// Reference: MaskingGenerator.__call__ / _mask (masking_generator.py:29-92), called once per image by
// DataAugmentationForBEiT.__call__ (datasets.py:104-118); the boolean gather lm_head(x[:, 1:][mask]) that consumes the mask
// (modeling_cyclical.py:221-225) fixes the order of the masked-row list: row-major over (image, patch).
//
// The generator is a sequential accept/reject loop per image (a few dozen rectangle proposals), trivially parallel over the
// images of a batch: one thread per image, the H x W grid held as one 32-bit word per grid row so a proposal costs h popc's.
// All arithmetic is IEEE double with explicit roundings (no FMA contraction), i.e. what CPython computes for
//   target_area = a + (b - a) * u;  aspect = exp(lo + (hi - lo) * u);  h = round(sqrt(area * aspect));  w = round(sqrt(area / aspect))
// so that with an INJECTED stream of uniforms the masks equal the reference generator's bit for bit (tests). Without one the
// uniforms come from Philox4x32-10 keyed on (seed; global image number, draw index): 53-bit doubles built like random.random().
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

struct MaskGenParams {
  int B, height, width, tokens;
  int num_masking, min_patches, max_patches;
  double log_lo, log_hi;
  unsigned long long seed, first_image;
  const double* uniforms;
  int per_image;
};

struct UniformStream {
  const double* inj;
  int n, cur;
  unsigned long long seed, image;
  Philox4 cache;
  bool overflow;
  __device__ double next() {
    if (inj != nullptr) {
      if (cur >= n) {
        overflow = true;
        return 0.0;
      }
      return inj[cur++];
    }
    if ((cur & 1) == 0)
      cache = philox4x32<10>((uint32_t)image, (uint32_t)(image >> 32), (uint32_t)(cur >> 1), 0x4D41534Bu /* 'MASK' */, (uint32_t)seed,
                             (uint32_t)(seed >> 32));
    const uint32_t a = (cur & 1) ? cache.z : cache.x, b = (cur & 1) ? cache.w : cache.y;
    ++cur;
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
  }
  // random.uniform(a, b) = a + (b - a) * random()
  __device__ double uniform(double a, double b) { return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), next())); }
  // integer in [0, n]; the injected-stream convention of the tests: floor(u * (n + 1))
  __device__ int randint(int n) {
    const int v = (int)__dmul_rn(next(), (double)(n + 1));
    return v < n ? v : n;
  }
};

__global__ void __launch_bounds__(32) block_masks_kernel(MaskGenParams p, uint8_t* __restrict__ mask, int32_t* __restrict__ count) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  uint32_t rows[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) rows[i] = 0u;
  UniformStream rng;
  rng.inj = p.uniforms ? p.uniforms + (size_t)b * p.per_image : nullptr;
  rng.n = p.per_image;
  rng.cur = 0;
  rng.seed = p.seed;
  rng.image = p.first_image + (unsigned long long)b;
  rng.overflow = false;

  int masked = 0;
  while (masked < p.num_masking && !rng.overflow) {                          // __call__, masking_generator.py:80-92
    const int max_mask = min(p.num_masking - masked, p.max_patches);
    int delta = 0;
    for (int attempt = 0; attempt < 10 && !rng.overflow; ++attempt) {        // _mask, :56-78
      const double area = rng.uniform((double)p.min_patches, (double)max_mask);
      const double aspect = exp(rng.uniform(p.log_lo, p.log_hi));
      const int h = (int)rint(sqrt(__dmul_rn(area, aspect)));
      const int w = (int)rint(sqrt(__ddiv_rn(area, aspect)));
      if (w < p.width && h < p.height) {
        const int top = rng.randint(p.height - h);
        const int left = rng.randint(p.width - w);
        if (rng.overflow) break;
        const uint32_t span = (w <= 0 ? 0u : (0xFFFFFFFFu >> (32 - w))) << left;
        int already = 0;
        for (int i = 0; i < h; ++i) already += __popc(rows[top + i] & span);
        const int fresh = h * w - already;
        if (fresh > 0 && fresh <= max_mask) {
          for (int i = 0; i < h; ++i) rows[top + i] |= span;
          delta = fresh;
        }
        if (delta > 0) break;
      }
    }
    if (delta == 0) break;
    masked += delta;
  }
  uint8_t* m = mask + (size_t)b * p.height * p.width;
  for (int i = 0; i < p.height; ++i)
    for (int j = 0; j < p.width; ++j) m[i * p.width + j] = (uint8_t)((rows[i] >> j) & 1u);
  count[b] = rng.overflow ? -1 : masked;
}

// Masked-row list in the reference's boolean-gather order: rows[k] = b * tokens + 1 + patch for the k-th set mask bit, row-major over
// (b, patch). One CTA: chunked exclusive scan of the per-image counts, then each thread writes the rows of its image.
constexpr int ROWS_THREADS = 256;
__global__ void __launch_bounds__(ROWS_THREADS) mask_rows_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ count, int B, int np,
                                                                int tokens, int32_t* __restrict__ rows) {
  __shared__ int sh[ROWS_THREADS];
  __shared__ int base_sh, bad_sh;
  if (threadIdx.x == 0) {
    base_sh = 0;
    bad_sh = 0;
  }
  __syncthreads();
  for (int c = 0; c < B; c += ROWS_THREADS) {
    const int b = c + threadIdx.x;
    const int n = b < B ? count[b] : 0;
    if (n < 0) bad_sh = 1;
    sh[threadIdx.x] = n > 0 ? n : 0;
    __syncthreads();
    for (int o = 1; o < ROWS_THREADS; o <<= 1) {                             // Hillis-Steele inclusive scan
      const int v = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += v;
      __syncthreads();
    }
    const int base = base_sh;
    if (b < B && n > 0 && rows != nullptr) {
      int k = base + sh[threadIdx.x] - n;
      const uint8_t* m = mask + (size_t)b * np;
      for (int q = 0; q < np; ++q)
        if (m[q]) rows[k++] = b * tokens + 1 + q;
    }
    __syncthreads();
    if (threadIdx.x == ROWS_THREADS - 1) base_sh = base + sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) count[B] = bad_sh ? -1 : base_sh;
}

// mask_dropout_prob of the pre-training loop (engine_for_cyclical.py:62-66): mask &= bernoulli(1 - p), per patch. One CTA per image;
// keep_in (uint8, optional) injects the Bernoulli draws, otherwise Philox4x32-10 keyed on (seed; image, patch).
__global__ void __launch_bounds__(256) mask_dropout_kernel(uint8_t* __restrict__ mask, int32_t* __restrict__ count, int np, float p_drop,
                                                           uint32_t k0, uint32_t k1, unsigned long long first_image,
                                                           const uint8_t* __restrict__ keep_in) {
  __shared__ int sh[8];
  const int b = blockIdx.x;
  int n = 0;
  for (int q = threadIdx.x; q < np; q += blockDim.x) {
    const size_t i = (size_t)b * np + q;
    bool keep;
    if (keep_in != nullptr) {
      keep = keep_in[i] != 0;
    } else {
      const unsigned long long img = first_image + (unsigned long long)b;
      const Philox4 r = philox4x32_10((uint32_t)q, 0x3c6ef372u, (uint32_t)img, (uint32_t)(img >> 32), k0, k1);
      keep = ((float)(r.x >> 8) * (1.0f / 16777216.0f)) >= p_drop;       // bernoulli(1 - p)
    }
    const uint8_t m = (mask[i] != 0 && keep) ? 1 : 0;
    mask[i] = m;
    n += m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    count[b] = t;
  }
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_block_masks(uint8_t* mask, int32_t* count, int32_t* rows, int32_t B, int32_t height, int32_t width, int32_t tokens,
                                   int32_t num_masking_patches, int32_t min_num_patches, int32_t max_num_patches, double log_aspect_lo,
                                   double log_aspect_hi, uint64_t seed, uint64_t first_image, const double* uniforms,
                                   int32_t uniforms_per_image, void* stream) {
  B200_CHECK_ARG(mask != nullptr && count != nullptr && B > 0, "block_masks: null output or empty batch");
  B200_CHECK_ARG(height > 0 && height <= 32 && width > 0 && width <= 32, "block_masks: grid %d x %d (each side must be in [1, 32])", height, width);
  B200_CHECK_ARG(tokens >= height * width + 1, "block_masks: tokens %d < patches + cls", tokens);
  B200_CHECK_ARG(num_masking_patches >= 0 && num_masking_patches <= height * width && min_num_patches >= 0 && max_num_patches >= 0,
                 "block_masks: bad patch counts");
  B200_CHECK_ARG(log_aspect_lo <= log_aspect_hi, "block_masks: aspect range");
  B200_CHECK_ARG(uniforms == nullptr || uniforms_per_image > 0, "block_masks: injected uniforms need uniforms_per_image > 0");
  MaskGenParams p;
  p.B = B; p.height = height; p.width = width; p.tokens = tokens;
  p.num_masking = num_masking_patches; p.min_patches = min_num_patches; p.max_patches = max_num_patches;
  p.log_lo = log_aspect_lo; p.log_hi = log_aspect_hi;
  p.seed = seed; p.first_image = first_image;
  p.uniforms = uniforms; p.per_image = uniforms_per_image;
  block_masks_kernel<<<(B + 31) / 32, 32, 0, STREAM>>>(p, mask, count);
  B200_CHECK_LAUNCH("block_masks");
  mask_rows_kernel<<<1, ROWS_THREADS, 0, STREAM>>>(mask, count, B, height * width, tokens, rows);
  B200_CHECK_LAUNCH("mask_rows");
  return 0;
}

extern "C" int b200vit_mask_dropout(uint8_t* mask, int32_t* count, int32_t* rows, int32_t B, int32_t num_patches, int32_t tokens, float p_drop,
                                    uint64_t seed, uint64_t first_image, const uint8_t* keep_in, void* stream) {
  B200_CHECK_ARG(mask != nullptr && count != nullptr && B > 0 && num_patches > 0, "mask_dropout: null pointer or empty batch");
  B200_CHECK_ARG(tokens >= num_patches + 1, "mask_dropout: tokens %d < patches + cls", tokens);
  B200_CHECK_ARG(p_drop >= 0.f && p_drop <= 1.f, "mask_dropout: p_drop %f outside [0, 1]", (double)p_drop);
  mask_dropout_kernel<<<B, 256, 0, STREAM>>>(mask, count, num_patches, p_drop, (uint32_t)seed, (uint32_t)(seed >> 32), first_image, keep_in);
  B200_CHECK_LAUNCH("mask_dropout");
  mask_rows_kernel<<<1, ROWS_THREADS, 0, STREAM>>>(mask, count, B, num_patches, tokens, rows);
  B200_CHECK_LAUNCH("mask_rows");
  return 0;
}
