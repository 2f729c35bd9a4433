// data2vec step, optional pieces of engine_for_cyclical.train_one_epoch (all HBM-bound, 128-bit vectorised):
//   channel_stats  : per (teacher layer, image, channel) mean / variance over the patch tokens, folded with the optional batch statistics
//                    into one {shift, scale} pair per (image, channel): F.batch_norm / F.instance_norm of the targets (:94-104, :112-115)
//   column_std     : z0 = sqrt(var(outputs, dim=0) + 1e-6) of the student rows, the var_w0 hinge and its gradient coefficients (:130-137)
//   gaussian_sample: z = mean + sqrt(max(cov, 0)) * eps, eps ~ N(0,1) from Philox4x32-10 / Box-Muller or injected; the optional
//                    reparameterised draw in front of `head` (modeling_finetune_dist.py:314-325, commented out in the reference: default off)
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

constexpr int MAX_LAYERS = 24;
struct LayerPtrs {
  const float* p[MAX_LAYERS];
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// grid (C/128, samples, layers), 256 threads = 8 warps x 32 lanes; lane owns 4 channels, warps stride over the rows of the sample.
// Two passes (mean, then centred squares): the second read of the 600 KB slab comes from L2.
__global__ void __launch_bounds__(256) chan_stats_kernel(LayerPtrs layers, long long ld, int sample_rows, int row0, int nrows, int C,
                                                         float2* __restrict__ out, int samples) {
  __shared__ float4 sh[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * 4;
  const int b = blockIdx.y, l = blockIdx.z;
  const float* base = layers.p[l] + ((long long)b * sample_rows + row0) * ld + c;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = warp; t < nrows; t += 8) {
    const float4 v = ld4(base + (long long)t * ld);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  sh[warp][lane] = s;
  __syncthreads();
  float4 m = sh[0][lane];
#pragma unroll
  for (int w = 1; w < 8; ++w) { m.x += sh[w][lane].x; m.y += sh[w][lane].y; m.z += sh[w][lane].z; m.w += sh[w][lane].w; }
  const float inv = 1.0f / nrows;
  m.x *= inv; m.y *= inv; m.z *= inv; m.w *= inv;
  __syncthreads();
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = warp; t < nrows; t += 8) {
    const float4 v = ld4(base + (long long)t * ld);
    const float dx = v.x - m.x, dy = v.y - m.y, dz = v.z - m.z, dw = v.w - m.w;
    q.x += dx * dx; q.y += dy * dy; q.z += dz * dz; q.w += dw * dw;
  }
  sh[warp][lane] = q;
  __syncthreads();
  if (warp == 0) {
    float4 v = sh[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) { v.x += sh[w][lane].x; v.y += sh[w][lane].y; v.z += sh[w][lane].z; v.w += sh[w][lane].w; }
    float2* o = out + ((long long)l * samples + b) * C + c;
    o[0] = make_float2(m.x, v.x * inv); o[1] = make_float2(m.y, v.y * inv);      // biased variance, as batch_norm / instance_norm normalise
    o[2] = make_float2(m.z, v.z * inv); o[3] = make_float2(m.w, v.w * inv);
  }
}

// {mean_bc, var_bc} -> {shift, scale}, in place. One thread per (layer, channel), loop over the images.
//   batch norm only : shift = mu_c,    scale = r_c = rsqrt(var_c + eps)                       (statistics over all images and tokens)
//   instance norm   : shift = mean_bc, scale = r_c * rsqrt(r_c^2 var_bc + eps)                (r_c = 1 without the batch norm in front)
__global__ void __launch_bounds__(128) chan_affine_kernel(float2* __restrict__ st, int L, int samples, int C, int batch_norm, int instance_norm,
                                                          float eps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * C) return;
  const int l = idx / C, c = idx % C;
  float2* p = st + (long long)l * samples * C + c;
  float mu = 0.f, rc = 1.f;
  if (batch_norm) {
    double sm = 0.0;
    for (int b = 0; b < samples; ++b) sm += p[(long long)b * C].x;
    mu = (float)(sm / samples);
    double sv = 0.0;
    for (int b = 0; b < samples; ++b) {
      const float2 v = p[(long long)b * C];
      sv += (double)v.y + ((double)v.x - mu) * ((double)v.x - mu);
    }
    rc = rsqrtf((float)(sv / samples) + eps);
  }
  for (int b = 0; b < samples; ++b) {
    const float2 v = p[(long long)b * C];
    float2 o;
    if (instance_norm) {
      o.x = v.x;
      o.y = rc * rsqrtf(rc * rc * v.y + eps);
    } else {
      o.x = mu;
      o.y = rc;
    }
    p[(long long)b * C] = o;
  }
}

// ---------------- z0 / var_w0 hinge ----------------
// partial {n, mean, M2} of each column over a chunk of the valid rows; grid (C/128, chunks)
__global__ void __launch_bounds__(256) colstat_partial_kernel(const float* __restrict__ y, int R, int C, const int* __restrict__ n_valid,
                                                              int chunks, float* __restrict__ part /* [chunks][3][C] */) {
  __shared__ float4 sh[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * 4;
  const int Rv = n_valid != nullptr ? min(*n_valid, R) : R;
  const int per = (Rv + chunks - 1) / chunks;
  const int r0 = min(blockIdx.y * per, Rv), r1 = min(r0 + per, Rv);
  const int n = r1 - r0;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0 + warp; r < r1; r += 8) {
    const float4 v = ld4(y + (long long)r * C + c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  sh[warp][lane] = s;
  __syncthreads();
  float4 m = sh[0][lane];
#pragma unroll
  for (int w = 1; w < 8; ++w) { m.x += sh[w][lane].x; m.y += sh[w][lane].y; m.z += sh[w][lane].z; m.w += sh[w][lane].w; }
  const float inv = n > 0 ? 1.0f / n : 0.f;
  m.x *= inv; m.y *= inv; m.z *= inv; m.w *= inv;
  __syncthreads();
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0 + warp; r < r1; r += 8) {
    const float4 v = ld4(y + (long long)r * C + c);
    const float dx = v.x - m.x, dy = v.y - m.y, dz = v.z - m.z, dw = v.w - m.w;
    q.x += dx * dx; q.y += dy * dy; q.z += dz * dz; q.w += dw * dw;
  }
  sh[warp][lane] = q;
  __syncthreads();
  if (warp == 0) {
    float4 v = sh[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) { v.x += sh[w][lane].x; v.y += sh[w][lane].y; v.z += sh[w][lane].z; v.w += sh[w][lane].w; }
    float* o = part + (long long)blockIdx.y * 3 * C;
    const float nf = (float)n;
    *reinterpret_cast<float4*>(o + c) = make_float4(nf, nf, nf, nf);
    *reinterpret_cast<float4*>(o + C + c) = m;
    *reinterpret_cast<float4*>(o + 2 * C + c) = v;
  }
}

// one CTA: combine the chunk statistics per column (Chan et al.), z0 = sqrt(unbiased var + eps) (torch.var default), hinge
// std_loss0 = sum relu(margin - z0) / C, and {mean_c, k_c} with k_c = d(k_scale * var_w0 * std_loss0)/dz0_c / ((n-1) z0_c)
__global__ void __launch_bounds__(1024) colstat_finalize_kernel(const float* __restrict__ part, int chunks, int C, float eps, float margin,
                                                                float k_scale, float* __restrict__ z0, float* __restrict__ hinge_out,
                                                                float2* __restrict__ col_hinge) {
  __shared__ float sh[32];
  float hsum = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double n = 0.0, mean = 0.0, m2 = 0.0;
    float nbs[32], mbs[32], qbs[32];                 // all partials of the column in flight at once (chunks <= 32)
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float* p = part + (long long)k * 3 * C;
      const bool in = k < chunks;
      nbs[k] = in ? p[c] : 0.f; mbs[k] = in ? p[C + c] : 0.f; qbs[k] = in ? p[2 * C + c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const double nb = nbs[k], mb = mbs[k], qb = qbs[k];
      if (nb <= 0.0) continue;
      const double d = mb - mean, nt = n + nb;
      mean += d * nb / nt;
      m2 += qb + d * d * n * nb / nt;
      n = nt;
    }
    const float var = n > 1.0 ? (float)(m2 / (n - 1.0)) : 0.f;
    const float z = sqrtf(var + eps);
    if (z0 != nullptr) z0[c] = z;
    const float h = fmaxf(margin - z, 0.f);
    hsum += h;
    if (col_hinge != nullptr) col_hinge[c] = make_float2((float)mean, (h > 0.f && n > 1.0) ? -k_scale / ((float)C * (float)(n - 1.0) * z) : 0.f);
  }
  hsum = warp_sum(hsum);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = hsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    if (hinge_out != nullptr) hinge_out[0] = t / C;
  }
}

// ---------------- reparameterised Gaussian sample ----------------
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

__global__ void __launch_bounds__(256) gaussian_sample_kernel(const float* __restrict__ mean, const float* __restrict__ cov,
                                                              const float* __restrict__ eps_in, long long n4, uint32_t k0, uint32_t k1,
                                                              uint32_t stream_id, float* __restrict__ eps_out, float* __restrict__ out_f32,
                                                              bf16* __restrict__ out_bf16) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float e[4];
    if (eps_in != nullptr) {
      const float4 v = ld4(eps_in + i * 4);
      e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
    } else {
      const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), stream_id, 0x6a09e667u, k0, k1);
      // Box-Muller: two uniform pairs -> four standard normals
      const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
      float s0, c0, s1, c1;
      sincospif(2.0f * u01(r.y), &s0, &c0);
      sincospif(2.0f * u01(r.w), &s1, &c1);
      e[0] = r0 * c0; e[1] = r0 * s0; e[2] = r1 * c1; e[3] = r1 * s1;
    }
    const float4 m = ld4(mean + i * 4), c = ld4(cov + i * 4);
    const float4 z = make_float4(fmaf(sqrtf(fmaxf(c.x, 0.f)), e[0], m.x), fmaf(sqrtf(fmaxf(c.y, 0.f)), e[1], m.y),
                                 fmaf(sqrtf(fmaxf(c.z, 0.f)), e[2], m.z), fmaf(sqrtf(fmaxf(c.w, 0.f)), e[3], m.w));
    if (eps_out != nullptr) *reinterpret_cast<float4*>(eps_out + i * 4) = make_float4(e[0], e[1], e[2], e[3]);
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + i * 4) = z;
    if (out_bf16 != nullptr) {
      uint2 u;
      u.x = pack_bf16x2(z.x, z.y);
      u.y = pack_bf16x2(z.z, z.w);
      *reinterpret_cast<uint2*>(out_bf16 + i * 4) = u;
    }
  }
}

__global__ void __launch_bounds__(256) gaussian_sample_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ cov,
                                                                  const float* __restrict__ eps, long long n, float* __restrict__ dmean,
                                                                  float* __restrict__ dcov) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = dz[i];
    if (dmean != nullptr) dmean[i] += g;
    if (dcov != nullptr) {
      const float c = cov[i];
      dcov[i] += c > 0.f ? g * eps[i] * 0.5f * rsqrtf(c) : 0.f;      // d sqrt(max(c,0)) / dc
    }
  }
}

// ModelEmaV2._update walks the WHOLE state dict (engine_for_cyclical.py:182-185), so the teacher's int64 relative_position_index buffer
// goes through `d * e + (1 - d) * m` too: int64 * python float -> fp32 products, fp32 sum, truncated back to int64 by copy_.
__global__ void __launch_bounds__(256) ema_index_kernel(int* __restrict__ e, const int* __restrict__ m, int n, float d, float od) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) e[i] = (int)__fadd_rn(__fmul_rn(d, (float)e[i]), __fmul_rn(od, (float)m[i]));
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_channel_stats(const float* const* layers_host, int32_t num_layers, int64_t ld_layer, int32_t samples, int32_t sample_rows,
                                     int32_t row0, int32_t nrows, int32_t C, int32_t batch_norm, int32_t instance_norm, float eps,
                                     float* affine_out, void* stream) {
  B200_CHECK_ARG(layers_host != nullptr && num_layers > 0 && num_layers <= MAX_LAYERS, "channel_stats: 1..%d layers", MAX_LAYERS);
  B200_CHECK_ARG(affine_out != nullptr && samples > 0 && nrows > 0 && row0 >= 0 && row0 + nrows <= sample_rows, "channel_stats: bad row window");
  B200_CHECK_ARG(C % 128 == 0 && ld_layer >= C && ld_layer % 4 == 0, "channel_stats: C=%d must be a multiple of 128", C);
  B200_CHECK_ARG(batch_norm || instance_norm, "channel_stats: neither batch_norm nor instance_norm requested");
  B200_CHECK_ARG(samples <= 65535 && num_layers <= 65535, "channel_stats: grid limits");
  LayerPtrs lp;
  for (int i = 0; i < num_layers; ++i) {
    B200_CHECK_ARG(layers_host[i] != nullptr, "channel_stats: layer %d is null", i);
    lp.p[i] = layers_host[i];
  }
  float2* st = reinterpret_cast<float2*>(affine_out);
  chan_stats_kernel<<<dim3(C / 128, samples, num_layers), 256, 0, STREAM>>>(lp, ld_layer, sample_rows, row0, nrows, C, st, samples);
  B200_CHECK_LAUNCH("channel_stats");
  chan_affine_kernel<<<(num_layers * C + 127) / 128, 128, 0, STREAM>>>(st, num_layers, samples, C, batch_norm, instance_norm, eps);
  B200_CHECK_LAUNCH("channel_affine");
  return 0;
}

constexpr int COLSTAT_CHUNKS = 32;
extern "C" size_t b200vit_column_std_workspace_bytes(int32_t C) { return (size_t)COLSTAT_CHUNKS * 3 * (size_t)C * sizeof(float); }

extern "C" int b200vit_column_std(const float* y, int32_t R, int32_t C, const int32_t* n_valid_dev, float eps, float margin, float k_scale,
                                  float* work, float* z0, float* hinge_out, float* col_hinge, void* stream) {
  B200_CHECK_ARG(y != nullptr && work != nullptr && R > 0 && C % 128 == 0, "column_std: bad arguments (C=%d must be a multiple of 128)", C);
  const int chunks = COLSTAT_CHUNKS;
  colstat_partial_kernel<<<dim3(C / 128, chunks), 256, 0, STREAM>>>(y, R, C, n_valid_dev, chunks, work);
  B200_CHECK_LAUNCH("column_std_partial");
  colstat_finalize_kernel<<<1, 1024, 0, STREAM>>>(work, chunks, C, eps, margin, k_scale, z0, hinge_out, reinterpret_cast<float2*>(col_hinge));
  B200_CHECK_LAUNCH("column_std_finalize");
  return 0;
}

extern "C" int b200vit_gaussian_sample(const float* mean, const float* cov, const float* eps_in, int64_t n, uint64_t seed, uint32_t stream_id,
                                       float* eps_out, float* out_f32, void* out_bf16, void* stream) {
  B200_CHECK_ARG(mean != nullptr && cov != nullptr && n >= 0 && n % 4 == 0, "gaussian_sample: n must be a multiple of 4");
  B200_CHECK_ARG(out_f32 != nullptr || out_bf16 != nullptr, "gaussian_sample: no output");
  if (n == 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)b200vit_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  gaussian_sample_kernel<<<(int)blocks, 256, 0, STREAM>>>(mean, cov, eps_in, n / 4, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, eps_out,
                                                          out_f32, static_cast<bf16*>(out_bf16));
  B200_CHECK_LAUNCH("gaussian_sample");
  return 0;
}

extern "C" int b200vit_gaussian_sample_bwd(const float* dz, const float* cov, const float* eps, int64_t n, float* dmean, float* dcov,
                                           void* stream) {
  B200_CHECK_ARG(dz != nullptr && n >= 0 && (dcov == nullptr || (cov != nullptr && eps != nullptr)), "gaussian_sample_bwd: null pointer");
  if (n == 0) return 0;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)b200vit_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  gaussian_sample_bwd_kernel<<<(int)blocks, 256, 0, STREAM>>>(dz, cov, eps, n, dmean, dcov);
  B200_CHECK_LAUNCH("gaussian_sample_bwd");
  return 0;
}

extern "C" int b200vit_ema_index_update(int32_t* ema_index, const int32_t* model_index, int32_t n, double decay, void* stream) {
  B200_CHECK_ARG(ema_index != nullptr && model_index != nullptr && n >= 0, "ema_index_update: null pointer");
  if (n == 0) return 0;
  ema_index_kernel<<<(n + 255) / 256, 256, 0, STREAM>>>(ema_index, model_index, n, (float)decay, (float)(1.0 - decay));
  B200_CHECK_LAUNCH("ema_index_update");
  return 0;
}
