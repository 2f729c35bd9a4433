// data2vec step kernels (HBM-bound, 128-bit vectorised):
//   d2v_target_loss : per masked row, LayerNorm (eps 1e-5, no affine) of the EMA teacher's top-K block outputs, their mean,
//                     optional post LayerNorm, smooth-L1(beta) / MSE against the student row, dLoss/dy   (engine_for_cyclical.py:90-150)
//   ema_update      : e = d*e + (1-d)*m over a flat fp32 arena (+ bf16 shadow of e)                     (engine_for_cyclical.py:182-185)
//   sumsq / adamw   : grad-norm, clip, AdamW, bf16 weight shadow and EMA in one pass over the arenas    (utils.py:364-390)
//   wasserstein_loss: WassersteinLoss fwd/bwd (distloss.py:13-30)
#include "../../include/b200vit.h"
#include "common.cuh"

namespace {

constexpr int MAXV = 8;  // float4 per lane -> C <= 1024
constexpr int MAX_LAYERS = 24;

struct LayerPtrs {
  const float* p[MAX_LAYERS];
};
struct AffinePtrs {
  const float2* p[MAX_LAYERS];   // per layer: [samples, C] {shift, scale} of the instance / batch norm (b200vit_channel_stats), or all NULL
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// no-affine LayerNorm of NV float4 per lane, in place; F.layer_norm(x.float(), (C,)) eps=1e-5
template <int NV>
__device__ __forceinline__ void ln_inplace(float4 (&v)[NV], int C, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) { v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd; }
}

// One warp per masked row. layers[l] is a [B, T, C] fp32 residual stream (cls row included): source row = row_index[r].
template <int NV>
__global__ void __launch_bounds__(256) d2v_target_loss_kernel(LayerPtrs layers, int num_layers, long long ld_layer,
                                                              const int* __restrict__ row_index, const float* __restrict__ y,
                                                              int R, int C, int ln_each, int ln_post, float beta, int l2_loss,
                                                              float grad_scale, float* __restrict__ targets,
                                                              bf16* __restrict__ dy_bf16, float* __restrict__ dy_f32,
                                                              float* __restrict__ row_loss, const int* __restrict__ n_valid,
                                                              AffinePtrs aff, int rows_per_sample, int compact_tokens,
                                                              const float2* __restrict__ col_hinge) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  if (n_valid != nullptr) {
    // R is the capacity of a padded row list: rows >= *n_valid carry no loss and no gradient, the mean runs over *n_valid rows
    const int rv = *n_valid;
    if (r >= rv) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (targets != nullptr) *reinterpret_cast<float4*>(targets + (long long)r * C + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dy_bf16 != nullptr) *reinterpret_cast<uint2*>(dy_bf16 + (long long)r * C + c) = make_uint2(0u, 0u);
        if (dy_f32 != nullptr) *reinterpret_cast<float4*>(dy_f32 + (long long)r * C + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (lane == 0 && row_loss != nullptr) row_loss[r] = 0.f;
      return;
    }
    grad_scale *= (float)R / (float)rv;
  }
  long long src = row_index != nullptr ? row_index[r] : r;
  // compact_tokens = T: the layers are compact [B, T-1, C] patch-row tensors (no cls row) while row_index holds residual-stream rows b*T+1+p
  if (compact_tokens > 0) src = src - src / compact_tokens - 1;
  const long long sample = rows_per_sample > 0 ? src / rows_per_sample : 0;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < num_layers; ++l) {
    const float* row = layers.p[l] + src * ld_layer;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = ld4_stream(row + (i * 32 + lane) * 4);
    if (aff.p[l] != nullptr) {
      // F.batch_norm / F.instance_norm over the patch tokens of each channel (engine_for_cyclical.py:94-104,112-115) as one affine map
      const float2* a = aff.p[l] + sample * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 a01 = *reinterpret_cast<const float4*>(a + (i * 32 + lane) * 4);
        const float4 a23 = *reinterpret_cast<const float4*>(a + (i * 32 + lane) * 4 + 2);
        v[i].x = (v[i].x - a01.x) * a01.y; v[i].y = (v[i].y - a01.z) * a01.w;
        v[i].z = (v[i].z - a23.x) * a23.y; v[i].w = (v[i].w - a23.z) * a23.w;
      }
    }
    if (ln_each) ln_inplace<NV>(v, C, 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) { acc[i].x += v[i].x; acc[i].y += v[i].y; acc[i].z += v[i].z; acc[i].w += v[i].w; }
  }
  const float inv = 1.0f / num_layers;
#pragma unroll
  for (int i = 0; i < NV; ++i) { acc[i].x *= inv; acc[i].y *= inv; acc[i].z *= inv; acc[i].w *= inv; }
  if (ln_post) ln_inplace<NV>(acc, C, 1e-5f);
  float loss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (targets != nullptr) *reinterpret_cast<float4*>(targets + (long long)r * C + c) = acc[i];
    if (y == nullptr) continue;
    const float4 yv = ld4(y + (long long)r * C + c);
    const float e[4] = {yv.x - acc[i].x, yv.y - acc[i].y, yv.z - acc[i].z, yv.w - acc[i].w};
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (l2_loss) {
        loss += e[j] * e[j];
        g[j] = 2.0f * e[j];
      } else {
        const float a = fabsf(e[j]);
        // F.smooth_l1_loss: 0.5*e^2/beta if |e| < beta else |e| - 0.5*beta
        loss += a < beta ? 0.5f * e[j] * e[j] / beta : a - 0.5f * beta;
        g[j] = a < beta ? e[j] / beta : (e[j] > 0.f ? 1.0f : -1.0f);
      }
      g[j] *= grad_scale;
    }
    if (col_hinge != nullptr) {
      // var_w0 hinge on the column standard deviation of the student output (engine_for_cyclical.py:130-137):
      // d/dy[r,c] = k_c * (y[r,c] - mean_c), {mean_c, k_c} prepared by b200vit_column_std
      const float4 h01 = *reinterpret_cast<const float4*>(col_hinge + c);
      const float4 h23 = *reinterpret_cast<const float4*>(col_hinge + c + 2);
      g[0] += h01.y * (yv.x - h01.x); g[1] += h01.w * (yv.y - h01.z);
      g[2] += h23.y * (yv.z - h23.x); g[3] += h23.w * (yv.w - h23.z);
    }
    if (dy_bf16 != nullptr) {
      uint2 u;
      u.x = pack_bf16x2(g[0], g[1]);
      u.y = pack_bf16x2(g[2], g[3]);
      *reinterpret_cast<uint2*>(dy_bf16 + (long long)r * C + c) = u;
    }
    if (dy_f32 != nullptr) *reinterpret_cast<float4*>(dy_f32 + (long long)r * C + c) = make_float4(g[0], g[1], g[2], g[3]);
  }
  loss = warp_sum(loss);
  if (lane == 0 && row_loss != nullptr) row_loss[r] = loss;
}

// deterministic single-CTA sum: out[0] = scale * sum(v[0..n)); with n_valid, scale is rescaled by n / *n_valid (mean over the valid rows)
__global__ void __launch_bounds__(1024) reduce_sum_kernel(const float* __restrict__ v, int n, float scale, float* __restrict__ out,
                                                          const int* __restrict__ n_valid, const float* __restrict__ add_term, float add_weight,
                                                          float mult) {
  __shared__ double sh[32];
  if (n_valid != nullptr) scale *= (float)n / (float)*n_valid;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      float r = (float)(s * (double)scale);
      if (add_term != nullptr) r += add_weight * add_term[0];
      out[0] = r * mult;
    }
  }
}

// e = d*e + (1-d)*m ; optional bf16 shadow of e
// bit-exact restatement of the reference lambda `cur_decay * e + (1. - cur_decay) * m` on fp32 tensors: both python scalars
// are rounded to fp32, two rounded products, one rounded add (no FMA contraction).
__device__ __forceinline__ float ema_mix(float d, float od, float e, float m) { return __fadd_rn(__fmul_rn(d, e), __fmul_rn(od, m)); }

__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ e, const float* __restrict__ m, long long n4, float d, float od,
                                                  bf16* __restrict__ shadow) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ld4(e + i * 4);
    const float4 b = ld4_stream(m + i * 4);
    // same expression as the reference lambda: cur_decay * e + (1. - cur_decay) * m
    const float4 o = make_float4(ema_mix(d, od, a.x, b.x), ema_mix(d, od, a.y, b.y), ema_mix(d, od, a.z, b.z), ema_mix(d, od, a.w, b.w));
    *reinterpret_cast<float4*>(e + i * 4) = o;
    if (shadow != nullptr) {
      uint2 u;
      u.x = pack_bf16x2(o.x, o.y);
      u.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(shadow + i * 4) = u;
    }
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n4, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ld4(g + i * 4);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  __shared__ float sh[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

// Fused clip + AdamW (+ bf16 weight shadow, + EMA teacher update and its bf16 shadow) over flat arenas.
// Per 1024-element chunk: hp[chunk] = {lr_scale, wd_scale}; lr = base_lr * lr_scale, weight_decay = base_wd * wd_scale
// (param groups of optim_factory.py:58-97: wd_scale 0 for 1-D / bias / skip-list params, lr_scale = layer decay; the
// per-step schedule values of engine_for_cyclical.py:47-53 arrive as base_lr / base_wd).
// torch.optim.AdamW update order: p *= 1 - lr*wd ; m,v EMA ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n4, const float2* __restrict__ hp,
                                                    float base_lr, float base_wd, float beta1, float beta2, float eps, float bc1, float sqrt_bc2,
                                                    const float* __restrict__ gnorm_sq, float max_norm, float grad_div,
                                                    bf16* __restrict__ p_shadow, float* __restrict__ ema, float ema_decay, float od,
                                                    bf16* __restrict__ ema_shadow) {
  float coef = 1.0f / grad_div;
  if (gnorm_sq != nullptr && max_norm > 0.f) {
    const float norm = sqrtf(__ldg(gnorm_sq)) / grad_div;
    coef *= fminf(1.0f, max_norm / (norm + 1e-6f));  // torch.nn.utils.clip_grad_norm_
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float2 h = __ldg(hp + (i >> 8));  // 1024 elements = 256 float4 per chunk
    const float lr = base_lr * h.x, wd = base_wd * h.y;
    float4 pv = ld4(p + i * 4);
    const float4 gv = ld4_stream(g + i * 4);
    float4 mv = ld4(m + i * 4), vv = ld4(v + i * 4);
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * coef;
      pp[j] *= 1.0f - lr * wd;
      mp[j] = beta1 * mp[j] + (1.0f - beta1) * gj;
      vp[j] = beta2 * vp[j] + (1.0f - beta2) * gj * gj;
      const float denom = sqrtf(vp[j]) / sqrt_bc2 + eps;
      pp[j] -= (lr / bc1) * (mp[j] / denom);
    }
    *reinterpret_cast<float4*>(p + i * 4) = pv;
    *reinterpret_cast<float4*>(m + i * 4) = mv;
    *reinterpret_cast<float4*>(v + i * 4) = vv;
    if (p_shadow != nullptr) {
      uint2 u;
      u.x = pack_bf16x2(pv.x, pv.y);
      u.y = pack_bf16x2(pv.z, pv.w);
      *reinterpret_cast<uint2*>(p_shadow + i * 4) = u;
    }
    if (ema != nullptr) {
      const float4 ev = ld4(ema + i * 4);
      const float4 o = make_float4(ema_mix(ema_decay, od, ev.x, pv.x), ema_mix(ema_decay, od, ev.y, pv.y), ema_mix(ema_decay, od, ev.z, pv.z),
                                   ema_mix(ema_decay, od, ev.w, pv.w));
      *reinterpret_cast<float4*>(ema + i * 4) = o;
      if (ema_shadow != nullptr) {
        uint2 u;
        u.x = pack_bf16x2(o.x, o.y);
        u.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(ema_shadow + i * 4) = u;
      }
    }
  }
}

// ---------------- WassersteinLoss (distloss.py:13-30) ----------------
// pass 1: w_r = sum (s(a)-s(g))^2 + sum (sqrt(max(s(b),1e-24)) - sqrt(max(s(h),1e-24)))^2 ; one warp per row
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) wloss_rowdist_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g,
                                                            const float* __restrict__ h, int R, int C, float* __restrict__ w) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 av = ld4(a + (long long)r * C + c), bv = ld4(b + (long long)r * C + c);
    const float4 gv = ld4(g + (long long)r * C + c), hv = ld4(h + (long long)r * C + c);
    const float* ap = &av.x; const float* bp = &bv.x; const float* gp = &gv.x; const float* hp = &hv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dm = sigm(ap[j]) - sigm(gp[j]);
      const float dc = sqrtf(fmaxf(sigm(bp[j]), 1e-24f)) - sqrtf(fmaxf(sigm(hp[j]), 1e-24f));
      s += dm * dm + dc * dc;
    }
  }
  s = warp_sum(s);
  if (lane == 0) w[r] = s;
}

// pass 2 (single CTA): wmax = max|w| ; l_r = -log(sigmoid(-w_r/wmax + 1e-24)) ; lmax = max|l| ; loss = lam * sum l_r / lmax
// and the row coefficients of dloss/dw_r INCLUDING the two max-normalisers (autograd differentiates through torch.max):
//   stats[0]=loss, stats[1]=wmax, stats[2]=lmax, stats[3]=argmax_w, stats[4]=argmax_l ; coef[r] = dloss/dw_r
__global__ void __launch_bounds__(1024) wloss_finalize_kernel(const float* __restrict__ w, int R_cap, float lam, float* __restrict__ stats,
                                                              float* __restrict__ coef, const int* __restrict__ n_valid) {
  // padded row list: only the first *n_valid rows take part in the normalisers and the sum; the rest get a zero coefficient
  const int R = n_valid != nullptr ? min(*n_valid, R_cap) : R_cap;
  for (int r = R + threadIdx.x; r < R_cap; r += blockDim.x) coef[r] = 0.f;
  __shared__ float sh_v[32];
  __shared__ int sh_i[32];
  __shared__ float bc[4];
  __shared__ int bi[2];
  auto block_argmax = [&](float v, int idx) {
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sh_v[threadIdx.x >> 5] = v; sh_i[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
      v = threadIdx.x < (blockDim.x >> 5) ? sh_v[threadIdx.x] : -1.f;
      idx = threadIdx.x < (blockDim.x >> 5) ? sh_i[threadIdx.x] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
      }
      if (threadIdx.x == 0) { sh_v[0] = v; sh_i[0] = idx; }
    }
    __syncthreads();
  };
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh_v[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      v = threadIdx.x < (blockDim.x >> 5) ? sh_v[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) sh_v[0] = v;
    }
    __syncthreads();
    const float r = sh_v[0];
    __syncthreads();
    return r;
  };
  float v = -1.f; int idx = 0x7fffffff;
  for (int r = threadIdx.x; r < R; r += blockDim.x) { const float a = fabsf(w[r]); if (a > v) { v = a; idx = r; } }
  block_argmax(v, idx);
  if (threadIdx.x == 0) { bc[0] = sh_v[0]; bi[0] = sh_i[0]; }
  __syncthreads();
  const float wmax = bc[0];
  const int iw = bi[0];
  // l_r and its max
  v = -1.f; idx = 0x7fffffff;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float z = -w[r] / wmax + 1e-24f;
    const float l = -logf(sigm(z));
    if (fabsf(l) > v) { v = fabsf(l); idx = r; }
  }
  block_argmax(v, idx);
  if (threadIdx.x == 0) { bc[1] = sh_v[0]; bi[1] = sh_i[0]; }
  __syncthreads();
  const float lmax = bc[1];
  const int il = bi[1];
  float sl = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) sl += -logf(sigm(-w[r] / wmax + 1e-24f));
  const float sum_l = block_sum(sl);
  // loss = lam * S / lmax with S = sum l_r, lmax = l_il (l >= 0).  dloss/dl_r = lam * (1/lmax - [r==il] * S / lmax^2)
  // l_r = softplus(-z_r), z_r = -u_r + 1e-24, u_r = w_r / wmax (w >= 0 -> wmax = w_iw): dl/du_r = sigmoid(-z_r) = 1 - sigmoid(z_r)
  // du_r/dw_k = [r==k]/wmax - [k==iw] * w_r / wmax^2
  float part = 0.f;  // sum_r dloss/du_r * w_r
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float z = -w[r] / wmax + 1e-24f;
    const float dl = lam * (1.0f / lmax - (r == il ? sum_l / (lmax * lmax) : 0.f));
    const float du = dl * (1.0f - sigm(z));
    coef[r] = du / wmax;
    part += du * w[r];
  }
  const float tot = block_sum(part);
  if (threadIdx.x == 0) {
    coef[iw] -= tot / (wmax * wmax);
    stats[0] = lam * sum_l / lmax; stats[1] = wmax; stats[2] = lmax; stats[3] = (float)iw; stats[4] = (float)il;
  }
}

// pass 3: da, db (grads of the student mean / cov outputs) = coef[r] * dw_r/d(.) * grad_scale, ACCUMULATED into da/db
__global__ void __launch_bounds__(256) wloss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g,
                                                        const float* __restrict__ h, const float* __restrict__ coef, int R, int C,
                                                        float grad_scale, float* __restrict__ da, float* __restrict__ db) {
  const long long total = (long long)R * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int r = (int)(i / C);
    const float k = coef[r] * grad_scale;
    const float sa = sigm(a[i]), sg = sigm(g[i]), sb = sigm(b[i]), shh = sigm(h[i]);
    da[i] += k * 2.0f * (sa - sg) * sa * (1.0f - sa);
    const float ub = sqrtf(fmaxf(sb, 1e-24f)), uh = sqrtf(fmaxf(shh, 1e-24f));
    // d sqrt(clamp(sb))/dsb = 0.5/ub where sb > 1e-24 (sigmoid of a finite float always is)
    db[i] += k * 2.0f * (ub - uh) * (sb > 1e-24f ? 0.5f / ub : 0.f) * sb * (1.0f - sb);
  }
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_d2v_target_loss_ex(const b200vit_d2v_desc* d, void* stream) {
  B200_CHECK_ARG(d != nullptr, "d2v_target_loss: null descriptor");
  const int R = d->R, C = d->C, num_layers = d->num_layers;
  B200_CHECK_ARG(d->layers != nullptr && num_layers > 0 && num_layers <= MAX_LAYERS, "d2v_target_loss: 1..%d layers", MAX_LAYERS);
  B200_CHECK_ARG(C % 128 == 0 && C <= 128 * MAXV, "d2v_target_loss: C=%d must be a multiple of 128, <= %d", C, 128 * MAXV);
  B200_CHECK_ARG(d->loss_out == nullptr || (d->row_loss != nullptr && d->y != nullptr), "d2v_target_loss: loss_out needs y and a row_loss workspace of R floats");
  B200_CHECK_ARG(d->affine == nullptr || d->rows_per_sample > 0, "d2v_target_loss: channel affine maps need rows_per_sample");
  B200_CHECK_ARG(d->col_hinge == nullptr || d->y != nullptr, "d2v_target_loss: the column-std hinge needs y");
  if (R == 0) return 0;
  B200_CHECK_ARG(R > 0, "d2v_target_loss: negative row count");
  LayerPtrs lp;
  AffinePtrs ap;
  for (int i = 0; i < MAX_LAYERS; ++i) ap.p[i] = nullptr;
  for (int i = 0; i < num_layers; ++i) {
    B200_CHECK_ARG(d->layers[i] != nullptr, "d2v_target_loss: layer %d is null", i);
    lp.p[i] = d->layers[i];
    if (d->affine != nullptr) ap.p[i] = reinterpret_cast<const float2*>(d->affine[i]);
  }
  const int grid = (R + 7) / 8;
#define TL(NV)                                                                                                                        \
  d2v_target_loss_kernel<NV><<<grid, 256, 0, STREAM>>>(lp, num_layers, d->ld_layer, d->row_index, d->y, R, C, d->ln_each, d->ln_post, \
                                                       d->beta, d->l2_loss, d->grad_scale, d->targets, static_cast<bf16*>(d->dy_bf16), \
                                                       d->dy_f32, d->row_loss, d->n_valid_dev, ap, d->rows_per_sample, d->compact_tokens, \
                                                       reinterpret_cast<const float2*>(d->col_hinge))
  switch (C / 128) {
    case 1: TL(1); break; case 2: TL(2); break; case 3: TL(3); break; case 4: TL(4); break;
    case 5: TL(5); break; case 6: TL(6); break; case 7: TL(7); break; default: TL(8); break;
  }
#undef TL
  B200_CHECK_LAUNCH("d2v_target_loss");
  if (d->loss_out != nullptr) {
    reduce_sum_kernel<<<1, 1024, 0, STREAM>>>(d->row_loss, R, 1.0f / ((float)R * (float)C), d->loss_out, d->n_valid_dev, d->loss_add,
                                              d->loss_add_weight, d->loss_mult);
    B200_CHECK_LAUNCH("d2v_loss_reduce");
  }
  return 0;
}

extern "C" int b200vit_d2v_target_loss(const float* const* layers_host, int32_t num_layers, int64_t ld_layer, const int32_t* row_index,
                                       const float* y, int32_t R, int32_t C, int32_t ln_each, int32_t ln_post, float beta, int32_t l2_loss,
                                       float grad_scale, float* targets, void* dy_bf16, float* dy_f32, float* row_loss, float* loss_out,
                                       const int32_t* n_valid_dev, void* stream) {
  B200_CHECK_ARG(row_index != nullptr, "d2v_target_loss: null row_index");
  b200vit_d2v_desc d = {};
  d.layers = layers_host; d.num_layers = num_layers; d.ld_layer = ld_layer; d.row_index = row_index; d.y = y; d.R = R; d.C = C;
  d.ln_each = ln_each; d.ln_post = ln_post; d.beta = beta; d.l2_loss = l2_loss; d.grad_scale = grad_scale; d.targets = targets;
  d.dy_bf16 = dy_bf16; d.dy_f32 = dy_f32; d.row_loss = row_loss; d.loss_out = loss_out; d.n_valid_dev = n_valid_dev; d.loss_mult = 1.0f;
  return b200vit_d2v_target_loss_ex(&d, stream);
}

// out = wa * *a + wb * *b  (device scalars; b may be NULL): total loss of the --stochastic step, (loss_cyc + loss_stochastic) * loss_scale
__global__ void scalar_fma_kernel(float* out, const float* a, float wa, const float* b, float wb) {
  out[0] = wa * a[0] + (b != nullptr ? wb * b[0] : 0.f);
}
extern "C" int b200vit_scalar_fma(float* out, const float* a, float wa, const float* b, float wb, void* stream) {
  B200_CHECK_ARG(out != nullptr && a != nullptr, "scalar_fma: null pointer");
  scalar_fma_kernel<<<1, 1, 0, STREAM>>>(out, a, wa, b, wb);
  B200_CHECK_LAUNCH("scalar_fma");
  return 0;
}

extern "C" int b200vit_ema_update(float* ema, const float* model, int64_t n, double decay, void* ema_bf16, void* stream) {
  B200_CHECK_ARG(ema != nullptr && model != nullptr && n >= 0 && n % 4 == 0, "ema_update: n=%lld must be a multiple of 4", (long long)n);
  B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(ema) | reinterpret_cast<uintptr_t>(model)) & 15) == 0, "ema_update: arenas must be 16-byte aligned");
  if (n == 0) return 0;
  const int sms = b200vit_num_sms();
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  ema_kernel<<<(int)blocks, 256, 0, STREAM>>>(ema, model, n / 4, (float)decay, (float)(1.0 - decay), static_cast<bf16*>(ema_bf16));
  B200_CHECK_LAUNCH("ema_update");
  return 0;
}

extern "C" int b200vit_sumsq(const float* g, int64_t n, float* out_accum, void* stream) {
  B200_CHECK_ARG(g != nullptr && out_accum != nullptr && n >= 0 && n % 4 == 0, "sumsq: n must be a multiple of 4");
  if (n == 0) return 0;
  const int sms = b200vit_num_sms();
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  sumsq_kernel<<<(int)blocks, 256, 0, STREAM>>>(g, n / 4, out_accum);
  B200_CHECK_LAUNCH("sumsq");
  return 0;
}

extern "C" int b200vit_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* hp_lr_wd, float lr, float weight_decay, float beta1, float beta2,
                                  float eps, int32_t step, const float* gnorm_sq, float max_norm, float grad_div, void* p_bf16, float* ema,
                                  double ema_decay, void* ema_bf16, void* stream) {
  B200_CHECK_ARG(p != nullptr && g != nullptr && m != nullptr && v != nullptr && hp_lr_wd != nullptr, "adamw_step: null pointer");
  B200_CHECK_ARG(n >= 0 && n % 1024 == 0, "adamw_step: arena length must be a multiple of 1024 (one {lr,wd} pair per 1024 elements)");
  B200_CHECK_ARG(step >= 1, "adamw_step: step counts from 1");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  const int sms = b200vit_num_sms();
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  adamw_kernel<<<(int)blocks, 256, 0, STREAM>>>(p, g, m, v, n / 4, reinterpret_cast<const float2*>(hp_lr_wd), lr, weight_decay, beta1, beta2, eps, (float)bc1,
                                                (float)sqrt(bc2), gnorm_sq, max_norm, grad_div > 0.f ? grad_div : 1.0f,
                                                static_cast<bf16*>(p_bf16), ema, (float)ema_decay, (float)(1.0 - ema_decay), static_cast<bf16*>(ema_bf16));
  B200_CHECK_LAUNCH("adamw_step");
  return 0;
}

extern "C" int b200vit_wasserstein_loss(const float* mean_out, const float* cov_out, const float* pos_mean, const float* pos_cov, int32_t R,
                                        int32_t C, float lam, float grad_scale, float* work /* 2R+8 floats */, float* d_mean_out,
                                        float* d_cov_out, float* loss_out, const int32_t* n_valid_dev, void* stream) {
  B200_CHECK_ARG(mean_out && cov_out && pos_mean && pos_cov && work && loss_out, "wasserstein_loss: null pointer");
  B200_CHECK_ARG(R > 0 && C % 4 == 0, "wasserstein_loss: bad shape R=%d C=%d", R, C);
  float* w = work;
  float* coef = work + R;
  float* stats = work + 2 * R;
  wloss_rowdist_kernel<<<(R + 7) / 8, 256, 0, STREAM>>>(mean_out, cov_out, pos_mean, pos_cov, R, C, w);
  B200_CHECK_LAUNCH("wloss_rowdist");
  wloss_finalize_kernel<<<1, 1024, 0, STREAM>>>(w, R, lam, stats, coef, n_valid_dev);
  B200_CHECK_LAUNCH("wloss_finalize");
  cudaMemcpyAsync(loss_out, stats, sizeof(float), cudaMemcpyDeviceToDevice, STREAM);
  if (d_mean_out != nullptr && d_cov_out != nullptr) {
    const int sms = b200vit_num_sms();
    long long blocks = ((long long)R * C + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    wloss_bwd_kernel<<<(int)blocks, 256, 0, STREAM>>>(mean_out, cov_out, pos_mean, pos_cov, coef, R, C, grad_scale, d_mean_out, d_cov_out);
    B200_CHECK_LAUNCH("wloss_bwd");
  }
  return 0;
}
