// Bandwidth-bound row kernels of the block forward/backward: LayerNorm fwd/bwd (with optional row gather/scatter),
// layer-scale + drop-path backward with its column reductions, bf16 column sums, fp32->bf16 casts, patch im2col,
// token assembly (cls concat + mask-token blend) and its backward, relative-position-bias gather, mean pooling.
// All loads/stores are 128-bit; one warp owns one row; column reductions accumulate per CTA and finish with one
// atomicAdd per column per CTA.
#include "../../include/b200vit.h"
#include "common.cuh"
#include "ptx_sm100.cuh"
#include <type_traits>
#include <stdlib.h>

namespace {

constexpr int LN_MAXV = 8;  // float4 per lane: C <= 32 * 4 * 8 = 1024

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_bf16x4(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float4 ld_bf16x4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta   (nn.LayerNorm eps=1e-6, modeling_finetune.py:270,280)
// two-pass variance in registers (matches ATen's fp32 result to ~1e-7)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long long ldx, const int* __restrict__ row_index,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                     int rows, int C, bf16* __restrict__ y, float* __restrict__ y32,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long src = row_index != nullptr ? row_index[warp] : warp;
  const float* xr = x + src * ldx;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = ld4(xr + (i * 32 + lane) * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    if (mean_out != nullptr) mean_out[warp] = mean;
    if (rstd_out != nullptr) rstd_out[warp] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gamma != nullptr) g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    if (beta != nullptr) b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (y != nullptr) st_bf16x4(y + (long long)warp * C + c, o);
    if (y32 != nullptr) st4(y32 + (long long)warp * C + c, o);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward. dx[src_row] += rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));
// dgamma += sum dy*xhat ; dbeta += sum dy.   Each warp loops over rows; per-CTA smem reduction; atomics at the end.
// ------------------------------------------------------------------------------------------------
// Optional second stage on the SAME rows, fused (the next op of the backward pass reads the dx this kernel has just produced): the backward
// of x_out = x_in + rowscale * gamma2 * t (scale_residual_bwd below): dt = rowscale gamma2 dx (bf16), dgamma2 += sum rowscale t dx,
// dbias2 += sum dt. Saves the 77 MB re-read of dx and one launch per LayerNorm of a block.
struct SrbStage {
  const bf16* t;            // null: stage off
  const float* rowscale;
  int rows_per_scale;
  const float* gamma2;
  bf16* dt;
  float* dgamma2;
  float* dbias2;
};

template <int NV, typename DY, bool SRB>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const DY* __restrict__ dy, const float* __restrict__ x, long long ldx,
                                                     const int* __restrict__ row_index, const float* __restrict__ gamma,
                                                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                     int rows, int C, float* __restrict__ dx, long long lddx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, const SrbStage srb) {
  extern __shared__ float red[];  // [2][C] (+ [2][C] for the fused stage)
  for (int i = threadIdx.x; i < (SRB ? 4 : 2) * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 g[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
    g[i] = gamma != nullptr ? __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 ag2[SRB ? NV : 1], ab2[SRB ? NV : 1], g2[SRB ? NV : 1];
  if constexpr (SRB) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      ag2[i] = ab2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      g2[i] = srb.gamma2 != nullptr ? __ldg(reinterpret_cast<const float4*>(srb.gamma2 + (i * 32 + lane) * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
  }
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const long long src = row_index != nullptr ? row_index[r] : r;
    const float* xr = x + src * ldx;
    const float mean = mean_in[r], rstd = rstd_in[r];
    float4 xh[NV], d[NV];
    float s1 = 0.f, s2 = 0.f;
    bool live = false;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 xv = ld4(xr + c);
      float4 dv;
      if constexpr (sizeof(DY) == 2) dv = ld_bf16x4(reinterpret_cast<const bf16*>(dy) + (long long)r * C + c);
      else dv = ld4(reinterpret_cast<const float*>(dy) + (long long)r * C + c);
      live |= (dv.x != 0.f) | (dv.y != 0.f) | (dv.z != 0.f) | (dv.w != 0.f);
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
      ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
      d[i] = make_float4(dv.x * g[i].x, dv.y * g[i].y, dv.z * g[i].z, dv.w * g[i].w);
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
    }
    // A row whose dy is all zero adds nothing. It is skipped rather than added as zeros because dx[src] is updated with a plain
    // read-modify-write: the padding entries of a fixed-capacity row list (b200vit_d2v_target_loss, n_valid_dev) may repeat a row
    // number that a live row of another warp is updating at the same time.
    if (!SRB && !__any_sync(0xffffffffu, live)) continue;      // (the fused stage must write dt for every row; it is never used with a row list)
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    float* dxr = dx + src * lddx;
    float rs = 1.0f;
    if constexpr (SRB) rs = srb.rowscale != nullptr ? __ldg(srb.rowscale + r / srb.rows_per_scale) : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o = ld4(dxr + c);
      o.x += rstd * (d[i].x - s1 - xh[i].x * s2);
      o.y += rstd * (d[i].y - s1 - xh[i].y * s2);
      o.z += rstd * (d[i].z - s1 - xh[i].z * s2);
      o.w += rstd * (d[i].w - s1 - xh[i].w * s2);
      st4(dxr + c, o);
      if constexpr (SRB) {
        const float4 tv = ld_bf16x4(srb.t + (long long)r * C + c);
        const float4 q = make_float4(rs * g2[i].x * o.x, rs * g2[i].y * o.y, rs * g2[i].z * o.z, rs * g2[i].w * o.w);
        st_bf16x4(srb.dt + (long long)r * C + c, q);
        ag2[i].x += rs * tv.x * o.x; ag2[i].y += rs * tv.y * o.y; ag2[i].z += rs * tv.z * o.z; ag2[i].w += rs * tv.w * o.w;
        ab2[i].x += q.x; ab2[i].y += q.y; ab2[i].z += q.z; ab2[i].w += q.w;
      }
    }
  }
  if constexpr (SRB) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      atomicAdd(&red[2 * C + c], ag2[i].x); atomicAdd(&red[2 * C + c + 1], ag2[i].y); atomicAdd(&red[2 * C + c + 2], ag2[i].z); atomicAdd(&red[2 * C + c + 3], ag2[i].w);
      atomicAdd(&red[3 * C + c], ab2[i].x); atomicAdd(&red[3 * C + c + 1], ab2[i].y); atomicAdd(&red[3 * C + c + 2], ab2[i].z); atomicAdd(&red[3 * C + c + 3], ab2[i].w);
    }
  }
  if (dgamma != nullptr || dbeta != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      atomicAdd(&red[c], ag[i].x); atomicAdd(&red[c + 1], ag[i].y); atomicAdd(&red[c + 2], ag[i].z); atomicAdd(&red[c + 3], ag[i].w);
      atomicAdd(&red[C + c], ab[i].x); atomicAdd(&red[C + c + 1], ab[i].y); atomicAdd(&red[C + c + 2], ab[i].z); atomicAdd(&red[C + c + 3], ab[i].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      if (dgamma != nullptr) atomicAdd(dgamma + i, red[i]);
      if (dbeta != nullptr) atomicAdd(dbeta + i, red[C + i]);
    }
  }
  if constexpr (SRB) {
    if (dgamma == nullptr && dbeta == nullptr) __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      if (srb.dgamma2 != nullptr) atomicAdd(srb.dgamma2 + i, red[2 * C + i]);
      if (srb.dbias2 != nullptr) atomicAdd(srb.dbias2 + i, red[3 * C + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, column-owner layout (dense rows, no row list): a thread owns 4 columns of LNC_RB rows per iteration, blockDim =
// (C/4, LNC_TY) — each y-slice of C/4 threads works on its own rows. The column sums (dgamma, dbeta and, fused, dgamma2 / dbias2) are 4 + 4
// (+ 4 + 4) registers per thread instead of 2 (4) x C/32 per lane of the warp-per-row kernel above (202 / 228 registers, one CTA of 8 warps
// per SM, 4.35 TB/s); the two row sums per row cross the slice through shared memory (one named barrier pair per LNC_RB rows). The dx / t
// loads of the second phase are issued before the reduction, so they are in flight across it. SRB: the scale-residual backward of the
// branch that consumes this dx (scale_residual_bwd_kernel below) runs in the same pass.
// ------------------------------------------------------------------------------------------------
constexpr int LNC_RB = 2, LNC_TY_MAX = 8, LNC_THREADS = 768;   // blockDim = (C/4, LNC_THREADS / (C/4)): one CTA of ~768 threads per SM
template <typename DY, bool SRB>
struct LncRaw {                      // the global loads of one iteration, issued one iteration ahead (software pipeline)
  typename std::conditional<sizeof(DY) == 2, uint2, float4>::type dy[LNC_RB];
  float4 x[LNC_RB], dx[LNC_RB];
  uint2 t[SRB ? LNC_RB : 1];
  float mean[LNC_RB], rstd[LNC_RB], rsc[SRB ? LNC_RB : 1];
};

template <typename DY, bool SRB>
__global__ void __launch_bounds__(LNC_THREADS, 1) ln_bwd_cols_kernel(const DY* __restrict__ dy, const float* __restrict__ x, long long ldx,
                                                                   const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                                   const float* __restrict__ rstd_in, int rows, int C, float* __restrict__ dx,
                                                                   long long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   const SrbStage srb) {
  extern __shared__ float sm[];
  const int TX = blockDim.x, TY = blockDim.y, tx = threadIdx.x, ty = threadIdx.y, nw = TX >> 5, w = tx >> 5, lane = tx & 31;
  float* part = sm + ty * (8 * 2 * LNC_RB + 2 * LNC_RB);   // [8 warps][2 RB] partial row sums of this slice
  float* fin = part + 8 * 2 * LNC_RB;                     // [2 RB]
  float* colred = sm + LNC_TY_MAX * (8 * 2 * LNC_RB + 2 * LNC_RB);   // [4][C] cross-slice column sums
  const int c = tx * 4;
  const float invC = 1.0f / (float)C;
  const float4 g = gamma != nullptr ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 g2 = make_float4(1.f, 1.f, 1.f, 1.f);
  if constexpr (SRB) { if (srb.gamma2 != nullptr) g2 = __ldg(reinterpret_cast<const float4*>(srb.gamma2 + c)); }
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, ag2 = ag, ab2 = ag;
  const int ngroups = (rows + TY * LNC_RB - 1) / (TY * LNC_RB);
  using Raw = LncRaw<DY, SRB>;
  auto load = [&](int grp, Raw& q) {
    const int r0 = (grp * TY + ty) * LNC_RB;
#pragma unroll
    for (int u = 0; u < LNC_RB; ++u) {
      const int r = r0 + u;
      const bool ok = grp < ngroups && r < rows;
      if constexpr (sizeof(DY) == 2) q.dy[u] = ok ? *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy) + (long long)r * C + c) : make_uint2(0u, 0u);
      else q.dy[u] = ok ? ld4(reinterpret_cast<const float*>(dy) + (long long)r * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      q.x[u] = ok ? ld4(x + (long long)r * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      q.dx[u] = ok ? ld4(dx + (long long)r * lddx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      q.mean[u] = ok ? __ldg(mean_in + r) : 0.f;
      q.rstd[u] = ok ? __ldg(rstd_in + r) : 0.f;
      if constexpr (SRB) {
        q.t[u] = ok ? __ldg(reinterpret_cast<const uint2*>(srb.t + (long long)r * C + c)) : make_uint2(0u, 0u);
        q.rsc[u] = ok ? (srb.rowscale != nullptr ? __ldg(srb.rowscale + r / srb.rows_per_scale) : 1.0f) : 0.0f;
      }
    }
  };
  auto process = [&](int grp, const Raw& cur) {
    const int r0 = (grp * TY + ty) * LNC_RB;
    float4 xh[LNC_RB], d[LNC_RB];
    float p1[LNC_RB], p2[LNC_RB];
#pragma unroll
    for (int u = 0; u < LNC_RB; ++u) {
      float4 dv;
      if constexpr (sizeof(DY) == 2) {
        const float2 a0 = unpack_bf16x2(cur.dy[u].x), a1 = unpack_bf16x2(cur.dy[u].y);
        dv = make_float4(a0.x, a0.y, a1.x, a1.y);
      } else {
        dv = cur.dy[u];
      }
      const float mean = cur.mean[u], rstd = cur.rstd[u];
      const float4 xv = cur.x[u];
      xh[u] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      ag.x += dv.x * xh[u].x; ag.y += dv.y * xh[u].y; ag.z += dv.z * xh[u].z; ag.w += dv.w * xh[u].w;
      ab.x += dv.x; ab.y += dv.y; ab.z += dv.z; ab.w += dv.w;
      d[u] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
      p1[u] = d[u].x + d[u].y + d[u].z + d[u].w;
      p2[u] = d[u].x * xh[u].x + d[u].y * xh[u].y + d[u].z * xh[u].z + d[u].w * xh[u].w;
    }
#pragma unroll
    for (int u = 0; u < LNC_RB; ++u) {
      p1[u] = warp_sum(p1[u]);
      p2[u] = warp_sum(p2[u]);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < LNC_RB; ++u) { part[w * 2 * LNC_RB + 2 * u] = p1[u]; part[w * 2 * LNC_RB + 2 * u + 1] = p2[u]; }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + ty), "r"(TX) : "memory");
    if (tx < 2 * LNC_RB) {
      float a = 0.f;
      for (int k = 0; k < nw; ++k) a += part[k * 2 * LNC_RB + tx];
      fin[tx] = a * invC;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + ty), "r"(TX) : "memory");
#pragma unroll
    for (int u = 0; u < LNC_RB; ++u) {
      const int r = r0 + u;
      if (r >= rows) continue;
      const float s1 = fin[2 * u], s2 = fin[2 * u + 1], rstd = cur.rstd[u];
      float4 o = cur.dx[u];
      o.x += rstd * (d[u].x - s1 - xh[u].x * s2);
      o.y += rstd * (d[u].y - s1 - xh[u].y * s2);
      o.z += rstd * (d[u].z - s1 - xh[u].z * s2);
      o.w += rstd * (d[u].w - s1 - xh[u].w * s2);
      st4(dx + (long long)r * lddx + c, o);
      if constexpr (SRB) {
        const float2 t0 = unpack_bf16x2(cur.t[u].x), t1 = unpack_bf16x2(cur.t[u].y);
        const float rs = cur.rsc[u];
        const float4 q = make_float4(rs * g2.x * o.x, rs * g2.y * o.y, rs * g2.z * o.z, rs * g2.w * o.w);
        st_bf16x4(srb.dt + (long long)r * C + c, q);
        ag2.x += rs * t0.x * o.x; ag2.y += rs * t0.y * o.y; ag2.z += rs * t1.x * o.z; ag2.w += rs * t1.y * o.w;
        ab2.x += q.x; ab2.y += q.y; ab2.z += q.z; ab2.w += q.w;
      }
    }
  };
  // two register buffers in ping-pong (no copies: a move out of a register with a load in flight would wait for it): the loads of group
  // k + 1 are in flight across the reduction, barriers and stores of group k
  Raw qa, qb;
  const int stride = gridDim.x;
  load(blockIdx.x, qa);
  for (int grp = blockIdx.x; grp < ngroups; grp += 2 * stride) {
    load(grp + stride, qb);
    process(grp, qa);
    load(grp + 2 * stride, qa);
    if (grp + stride < ngroups) process(grp + stride, qb);
  }
  // column sums: slices add into shared memory one after the other, then ONE vector atomic per 4 columns and quantity per CTA
  for (int s = 0; s < TY; ++s) {
    if (ty == s) {
      float4* cr = reinterpret_cast<float4*>(colred);
      const int q4 = C >> 2;
      auto put = [&](int k, const float4& v) {
        float4 cur = s == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : cr[k * q4 + tx];
        cr[k * q4 + tx] = make_float4(cur.x + v.x, cur.y + v.y, cur.z + v.z, cur.w + v.w);
      };
      put(0, ag); put(1, ab);
      if constexpr (SRB) { put(2, ag2); put(3, ab2); }
    }
    __syncthreads();
  }
  if (ty == 0) {
    const float4* cr = reinterpret_cast<const float4*>(colred);
    const int q4 = C >> 2;
    auto red4 = [](float* p, const float4& v) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    };
    if (dgamma != nullptr) red4(dgamma + c, cr[tx]);
    if (dbeta != nullptr) red4(dbeta + c, cr[q4 + tx]);
    if constexpr (SRB) {
      if (srb.dgamma2 != nullptr) red4(srb.dgamma2 + c, cr[2 * q4 + tx]);
      if (srb.dbias2 != nullptr) red4(srb.dbias2 + c, cr[3 * q4 + tx]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward (+ fused scale-residual backward) with the operand streams staged through shared memory by bulk async copies.
// The register-staged kernels above top out at 4.4-4.7 TB/s: the bytes a thread can keep in flight are the registers it can spare
// (measured: software pipelining, 384 / 768-thread CTAs, ping-pong buffers all land between 57 and 82 us for 271-349 MB). Here one producer
// lane streams LNB_G-row chunks of dy (bf16), x, dx (fp32), t (bf16) and the rows' mean / rstd into a ring of shared-memory stages with
// cp.async.bulk + mbarrier (up to ~190 KB in flight per SM, no registers involved), and two consumer groups of C/4 threads (column-owner
// layout: 4 columns x LNB_G rows per thread, column sums in 16 registers) take alternate stages. Rows are contiguous (ldx == lddx == C) and
// rows % LNB_G == 0 on this path; everything else uses the kernels above.
// ------------------------------------------------------------------------------------------------
constexpr int LNB_G = 4;                 // rows per stage
constexpr int LNB_GROUPS = 2;            // consumer groups
constexpr int LNB_MAX_STAGES = 6;
struct LnbParams {
  const bf16* dy; const float* x; const float* gamma; const float* mean; const float* rstd;
  float* dx; float* dgamma; float* dbeta;
  SrbStage srb;
  int rows, C, nstages, stage_bytes;
};

template <bool SRB>
__global__ void __launch_bounds__(LNB_GROUPS * 256 + 32, 1) ln_bwd_bulk_kernel(const LnbParams p) {
  extern __shared__ __align__(128) uint8_t lnb_smem[];
  const int C = p.C, TX = C >> 2;
  // stage layout: dy [G][C] bf16 | x [G][C] f32 | dx [G][C] f32 | t [G][C] bf16 (SRB) | mean [G] | rstd [G]
  const int off_x = LNB_G * C * 2, off_dx = off_x + LNB_G * C * 4, off_t = off_dx + LNB_G * C * 4;
  const int off_ms = off_t + (SRB ? LNB_G * C * 2 : 0);
  uint8_t* stages = lnb_smem;
  float* red = reinterpret_cast<float*>(lnb_smem + (size_t)p.nstages * p.stage_bytes);      // [GROUPS][8 warps][2 G] + [GROUPS][2 G] + [4][C]
  const uint32_t bar0 = ptx::smem_u32(red + LNB_GROUPS * (8 * 2 * LNB_G + 2 * LNB_G) + 4 * C);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (LNB_MAX_STAGES + s); };
  const int tid = threadIdx.x;
  const int nchunks = p.rows / LNB_G;
  const int first = blockIdx.x, stride = gridDim.x;
  const int my = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;
  if (tid == 0) {
    for (int s = 0; s < p.nstages; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int ctid = tid - 32;                      // consumer thread index (tid >= 32)
  if (tid < 32) {
    if (tid == 0) {
      // ---------------- producer ----------------
      const uint32_t tx_bytes = (uint32_t)(LNB_G * C * (SRB ? 12 : 10) + 2 * LNB_G * 4);
      for (int i = 0; i < my; ++i) {
        const int s = i % p.nstages;
        if (i >= p.nstages) ptx::mbar_wait(empty(s), (uint32_t)(((i / p.nstages) - 1) & 1));
        const long long r0 = (long long)(first + i * stride) * LNB_G;
        const uint32_t dst = ptx::smem_u32(stages + (size_t)s * p.stage_bytes);
        ptx::mbar_arrive_expect_tx(full(s), tx_bytes);
        ptx::bulk_load_1d(dst, p.dy + r0 * C, LNB_G * C * 2, full(s));
        ptx::bulk_load_1d(dst + off_x, p.x + r0 * C, LNB_G * C * 4, full(s));
        ptx::bulk_load_1d(dst + off_dx, p.dx + r0 * C, LNB_G * C * 4, full(s));
        if (SRB) ptx::bulk_load_1d(dst + off_t, p.srb.t + r0 * C, LNB_G * C * 2, full(s));
        ptx::bulk_load_1d(dst + off_ms, p.mean + r0, LNB_G * 4, full(s));
        ptx::bulk_load_1d(dst + off_ms + LNB_G * 4, p.rstd + r0, LNB_G * 4, full(s));
      }
    }
    return;
  }
  // ---------------- consumers ----------------
  const int grp = ctid / TX, tx = ctid - grp * TX;
  if (grp >= LNB_GROUPS) return;                  // blockDim is 32 + GROUPS * TX exactly; defensive
  const int w = tx >> 5, lane = tx & 31, nw = TX >> 5;
  float* part = red + grp * (8 * 2 * LNB_G + 2 * LNB_G);
  float* fin = part + 8 * 2 * LNB_G;
  float* colred = red + LNB_GROUPS * (8 * 2 * LNB_G + 2 * LNB_G);
  const int c = tx * 4;
  const float invC = 1.0f / (float)C;
  const float4 g = p.gamma != nullptr ? __ldg(reinterpret_cast<const float4*>(p.gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 g2 = make_float4(1.f, 1.f, 1.f, 1.f);
  if constexpr (SRB) { if (p.srb.gamma2 != nullptr) g2 = __ldg(reinterpret_cast<const float4*>(p.srb.gamma2 + c)); }
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, ag2 = ag, ab2 = ag;
  for (int i = grp; i < my; i += LNB_GROUPS) {
    const int s = i % p.nstages;
    const long long r0 = (long long)(first + i * stride) * LNB_G;
    const uint8_t* st = stages + (size_t)s * p.stage_bytes;
    ptx::mbar_wait(full(s), (uint32_t)((i / p.nstages) & 1));
    float4 xh[LNB_G], d[LNB_G];
    float p1[LNB_G], p2[LNB_G], rstd[LNB_G];
#pragma unroll
    for (int u = 0; u < LNB_G; ++u) {
      const uint2 dr = *reinterpret_cast<const uint2*>(st + (u * C + c) * 2);
      const float4 xv = *reinterpret_cast<const float4*>(st + off_x + (u * C + c) * 4);
      const float mean = *reinterpret_cast<const float*>(st + off_ms + u * 4);
      rstd[u] = *reinterpret_cast<const float*>(st + off_ms + (LNB_G + u) * 4);
      const float2 a0 = unpack_bf16x2(dr.x), a1 = unpack_bf16x2(dr.y);
      const float4 dv = make_float4(a0.x, a0.y, a1.x, a1.y);
      xh[u] = make_float4((xv.x - mean) * rstd[u], (xv.y - mean) * rstd[u], (xv.z - mean) * rstd[u], (xv.w - mean) * rstd[u]);
      ag.x += dv.x * xh[u].x; ag.y += dv.y * xh[u].y; ag.z += dv.z * xh[u].z; ag.w += dv.w * xh[u].w;
      ab.x += dv.x; ab.y += dv.y; ab.z += dv.z; ab.w += dv.w;
      d[u] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
      p1[u] = d[u].x + d[u].y + d[u].z + d[u].w;
      p2[u] = d[u].x * xh[u].x + d[u].y * xh[u].y + d[u].z * xh[u].z + d[u].w * xh[u].w;
    }
#pragma unroll
    for (int u = 0; u < LNB_G; ++u) {
      p1[u] = warp_sum(p1[u]);
      p2[u] = warp_sum(p2[u]);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < LNB_G; ++u) { part[w * 2 * LNB_G + 2 * u] = p1[u]; part[w * 2 * LNB_G + 2 * u + 1] = p2[u]; }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(TX) : "memory");
    if (tx < 2 * LNB_G) {
      float a = 0.f;
      for (int k = 0; k < nw; ++k) a += part[k * 2 * LNB_G + tx];
      fin[tx] = a * invC;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(TX) : "memory");
    float rsc[LNB_G];
    if constexpr (SRB) {
#pragma unroll
      for (int u = 0; u < LNB_G; ++u)
        rsc[u] = p.srb.rowscale != nullptr ? __ldg(p.srb.rowscale + ((int)r0 + u) / p.srb.rows_per_scale) : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < LNB_G; ++u) {
      const float s1 = fin[2 * u], s2 = fin[2 * u + 1];
      float4 o = *reinterpret_cast<const float4*>(st + off_dx + (u * C + c) * 4);
      o.x += rstd[u] * (d[u].x - s1 - xh[u].x * s2);
      o.y += rstd[u] * (d[u].y - s1 - xh[u].y * s2);
      o.z += rstd[u] * (d[u].z - s1 - xh[u].z * s2);
      o.w += rstd[u] * (d[u].w - s1 - xh[u].w * s2);
      st4(p.dx + (r0 + u) * C + c, o);
      if constexpr (SRB) {
        const uint2 tr = *reinterpret_cast<const uint2*>(st + off_t + (u * C + c) * 2);
        const float2 t0 = unpack_bf16x2(tr.x), t1 = unpack_bf16x2(tr.y);
        const float rs = rsc[u];
        const float4 q = make_float4(rs * g2.x * o.x, rs * g2.y * o.y, rs * g2.z * o.z, rs * g2.w * o.w);
        st_bf16x4(p.srb.dt + (r0 + u) * C + c, q);
        ag2.x += rs * t0.x * o.x; ag2.y += rs * t0.y * o.y; ag2.z += rs * t1.x * o.z; ag2.w += rs * t1.y * o.w;
        ab2.x += q.x; ab2.y += q.y; ab2.z += q.z; ab2.w += q.w;
      }
    }
    // every thread of the group has read the stage (fin was read after the second barrier, the stage just above): hand it back
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(TX) : "memory");
    if (tx == 0) ptx::mbar_arrive(empty(s));
  }
  // column sums: the groups add into shared memory one after the other (all consumers, barrier 8), then one vector atomic per 4 columns
  float4* cr = reinterpret_cast<float4*>(colred);
  for (int sgrp = 0; sgrp < LNB_GROUPS; ++sgrp) {
    if (grp == sgrp) {
      auto put = [&](int k, const float4& v) {
        const float4 cur = sgrp == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : cr[k * TX + tx];
        cr[k * TX + tx] = make_float4(cur.x + v.x, cur.y + v.y, cur.z + v.z, cur.w + v.w);
      };
      put(0, ag); put(1, ab);
      if constexpr (SRB) { put(2, ag2); put(3, ab2); }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(8), "r"(LNB_GROUPS * TX) : "memory");
  }
  if (grp == 0) {
    auto red4 = [](float* q, const float4& v) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    };
    if (p.dgamma != nullptr) red4(p.dgamma + c, cr[tx]);
    if (p.dbeta != nullptr) red4(p.dbeta + c, cr[TX + tx]);
    if constexpr (SRB) {
      if (p.srb.dgamma2 != nullptr) red4(p.srb.dgamma2 + c, cr[2 * TX + tx]);
      if (p.srb.dbias2 != nullptr) red4(p.srb.dbias2 + c, cr[3 * TX + tx]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward of  x_out = x_in + rowscale[b] * gamma * t   (Block.forward, modeling_finetune.py:296-298):
//   dt = rowscale*gamma*dx (bf16) ; dgamma += sum rowscale*t*dx ; dbias += sum dt      (dbias = grad of the branch's
//   last Linear bias, i.e. column sum of dt)
// ------------------------------------------------------------------------------------------------
// Column-reducing row kernels: blockDim = (column groups, ROWL row lanes), ONE CTA per SM-slot, RB_UNROLL rows in flight per thread
// (enough bytes in flight for HBM), the row lanes are combined in shared memory and each CTA issues ONE atomic per column —
// same-address atomics from many small CTAs were the bottleneck of the first version.
constexpr int RB_UNROLL = 4;

constexpr int SRB_UNROLL = 2;
__global__ void scale_residual_bwd_kernel(const float* __restrict__ dx, long long lddx, const bf16* __restrict__ t,
                                          const float* __restrict__ rowscale, int rows_per_scale, const float* __restrict__ gamma, int rows, int C,
                                          bf16* __restrict__ dt, float* __restrict__ dgamma, float* __restrict__ dbias) {
  extern __shared__ float red[];   // [2][C]
  const int c = threadIdx.x * 4;
  const int rowl = blockDim.y;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 2 * C; i += blockDim.x * blockDim.y) red[i] = 0.f;
  __syncthreads();
  const float4 gm = gamma != nullptr ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
  const int step = gridDim.x * rowl * SRB_UNROLL;
  // software pipeline: the loads of the next SRB_UNROLL rows are in flight while the current ones are scaled and stored (a thread that
  // loads, computes and stores in turn keeps its bytes in flight only half of the time: 3.5 TB/s before, measured)
  struct Raw {
    float4 d[SRB_UNROLL];
    uint2 t[SRB_UNROLL];
    float rs[SRB_UNROLL];
  };
  auto load = [&](int r0, Raw& q) {
#pragma unroll
    for (int u = 0; u < SRB_UNROLL; ++u) {
      const int r = r0 + u;
      const bool ok = r < rows;
      q.d[u] = ok ? ld4(dx + (long long)r * lddx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      q.t[u] = (ok && t != nullptr) ? __ldg(reinterpret_cast<const uint2*>(t + (long long)r * C + c)) : make_uint2(0u, 0u);
      q.rs[u] = ok ? (rowscale != nullptr ? __ldg(rowscale + r / rows_per_scale) : 1.0f) : 0.0f;
    }
  };
  auto process = [&](int r0, const Raw& cur) {
#pragma unroll
    for (int u = 0; u < SRB_UNROLL; ++u) {
      const int r = r0 + u;
      if (r < rows) {
        const float rs = cur.rs[u];
        const float4 d = cur.d[u];
        const float2 t0 = unpack_bf16x2(cur.t[u].x), t1 = unpack_bf16x2(cur.t[u].y);
        const float4 o = make_float4(rs * gm.x * d.x, rs * gm.y * d.y, rs * gm.z * d.z, rs * gm.w * d.w);
        st_bf16x4(dt + (long long)r * C + c, o);
        ag.x += rs * t0.x * d.x; ag.y += rs * t0.y * d.y; ag.z += rs * t1.x * d.z; ag.w += rs * t1.y * d.w;
        ab.x += o.x; ab.y += o.y; ab.z += o.z; ab.w += o.w;
      }
    }
  };
  Raw qa, qb;                               // ping-pong (no register copies of in-flight loads)
  int r0 = (blockIdx.x * rowl + threadIdx.y) * SRB_UNROLL;
  load(r0, qa);
  for (; r0 < rows; r0 += 2 * step) {
    load(r0 + step, qb);
    process(r0, qa);
    load(r0 + 2 * step, qa);
    process(r0 + step, qb);
  }
  atomicAdd(&red[c], ag.x); atomicAdd(&red[c + 1], ag.y); atomicAdd(&red[c + 2], ag.z); atomicAdd(&red[c + 3], ag.w);
  atomicAdd(&red[C + c], ab.x); atomicAdd(&red[C + c + 1], ab.y); atomicAdd(&red[C + c + 2], ab.z); atomicAdd(&red[C + c + 3], ab.w);
  __syncthreads();
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < C; i += blockDim.x * blockDim.y) {
    if (dgamma != nullptr) atomicAdd(dgamma + i, red[i]);
    if (dbias != nullptr) atomicAdd(dbias + i, red[C + i]);
  }
}

// out[c] += sum_r x[r, c]   (bias gradients, e.g. fc1.bias from dpre). One thread owns one 8-column group (16-byte loads).
__global__ void colsum_bf16_kernel(const bf16* __restrict__ x, long long ldx, int rows, int C, float* __restrict__ out) {
  extern __shared__ float red[];   // [C]
  const int c = threadIdx.x * 8;
  const int rowl = blockDim.y;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < C; i += blockDim.x * blockDim.y) red[i] = 0.f;
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int step = gridDim.x * rowl * RB_UNROLL;
  for (int r0 = (blockIdx.x * rowl + threadIdx.y) * RB_UNROLL; r0 < rows; r0 += step) {
    uint4 v[RB_UNROLL];
#pragma unroll
    for (int u = 0; u < RB_UNROLL; ++u)
      if (r0 + u < rows) v[u] = *reinterpret_cast<const uint4*>(x + (long long)(r0 + u) * ldx + c);
#pragma unroll
    for (int u = 0; u < RB_UNROLL; ++u)
      if (r0 + u < rows) {
        const uint32_t* w = &v[u].x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(w[k]);
          acc[2 * k] += f.x; acc[2 * k + 1] += f.y;
        }
      }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) atomicAdd(&red[c + k], acc[k]);
  __syncthreads();
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < C; i += blockDim.x * blockDim.y) atomicAdd(out + i, red[i]);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n4, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) st_bf16x4(dst + i * 4, ld4(src + i * 4));
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) dst[i] = __float2bfloat16(src[i]);
}

// images [B, Cin, H, W] fp32 -> patch matrix [B*gh*gw, Cin*P*P] bf16, k = (c, py, px) as Conv2d weight.flatten(1)
// (PatchEmbed, modeling_finetune.py:319-325). One thread moves 4 consecutive pixels of a patch row.
__global__ void im2col_kernel(const float* __restrict__ img, int B, int Cin, int H, int W, int P, bf16* __restrict__ out) {
  const int gh = H / P, gw = W / P;
  const int K = Cin * P * P;
  const long long total4 = (long long)B * gh * gw * K / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const long long e = i * 4;
    const int k = (int)(e % K);
    const long long row = e / K;
    const int px = k % P, py = (k / P) % P, c = k / (P * P);
    const int gx = (int)(row % gw), gy = (int)((row / gw) % gh);
    const int b = (int)(row / ((long long)gw * gh));
    const float* s = img + (((long long)b * Cin + c) * H + gy * P + py) * W + gx * P + px;
    st_bf16x4(out + e, ld4(s));
  }
}

// x[b, 0] = cls ; x[b, 1+p] = pe[b,p] * (1 - w) + mask_token * w   (+ pos_embed)   (modeling_cyclical.py:175-194)
__global__ void assemble_tokens_kernel(const float* __restrict__ pe, const float* __restrict__ cls, const float* __restrict__ mask_token,
                                       const uint8_t* __restrict__ mask, const float* __restrict__ pos, int B, int np, int C,
                                       float* __restrict__ x) {
  const int T = np + 1;
  const int c4 = C >> 2;
  const long long total = (long long)B * T * c4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % c4) * 4;
    const long long row = i / c4;
    const int t = (int)(row % T);
    const int b = (int)(row / T);
    float4 v;
    if (t == 0) {
      v = ld4(cls + c);
    } else {
      const long long prow = (long long)b * np + (t - 1);
      v = ld4(pe + prow * C + c);
      if (mask != nullptr) {
        const float w = mask[prow] ? 1.0f : 0.0f;
        const float4 m = ld4(mask_token + c);
        v = make_float4(v.x * (1.f - w) + m.x * w, v.y * (1.f - w) + m.y * w, v.z * (1.f - w) + m.z * w, v.w * (1.f - w) + m.w * w);
      }
    }
    if (pos != nullptr) {
      const float4 p = ld4(pos + (long long)t * C + c);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    st4(x + row * C + c, v);
  }
}

// backward of assemble_tokens: dpe = dx[:,1:] * (1-w) (bf16) ; dmask_token += sum dx*w ; dcls += sum_b dx[:,0] ; dpos += sum_b dx
// grid.x strides over batch entries; thread = 4 channels.
__global__ void assemble_tokens_bwd_kernel(const float* __restrict__ dx, const uint8_t* __restrict__ mask, int B, int np, int C,
                                           bf16* __restrict__ dpe, float* __restrict__ dcls, float* __restrict__ dmask_token,
                                           float* __restrict__ dpos) {
  const int T = np + 1;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= C) return;
  float4 am = make_float4(0.f, 0.f, 0.f, 0.f), ac = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float4 d0 = ld4(dx + ((long long)b * T) * C + c);
    ac.x += d0.x; ac.y += d0.y; ac.z += d0.z; ac.w += d0.w;
    if (dpos != nullptr) { atomicAdd(dpos + c, d0.x); atomicAdd(dpos + c + 1, d0.y); atomicAdd(dpos + c + 2, d0.z); atomicAdd(dpos + c + 3, d0.w); }
    for (int p = 0; p < np; ++p) {
      const long long prow = (long long)b * np + p;
      const float4 d = ld4(dx + ((long long)b * T + 1 + p) * C + c);
      const float w = (mask != nullptr && mask[prow]) ? 1.0f : 0.0f;
      st_bf16x4(dpe + prow * C + c, make_float4(d.x * (1.f - w), d.y * (1.f - w), d.z * (1.f - w), d.w * (1.f - w)));
      am.x += d.x * w; am.y += d.y * w; am.z += d.z * w; am.w += d.w * w;
      if (dpos != nullptr) {
        float* q = dpos + (long long)(1 + p) * C + c;
        atomicAdd(q, d.x); atomicAdd(q + 1, d.y); atomicAdd(q + 2, d.z); atomicAdd(q + 3, d.w);
      }
    }
  }
  if (dcls != nullptr) { atomicAdd(dcls + c, ac.x); atomicAdd(dcls + c + 1, ac.y); atomicAdd(dcls + c + 2, ac.z); atomicAdd(dcls + c + 3, ac.w); }
  if (dmask_token != nullptr && mask != nullptr) {
    atomicAdd(dmask_token + c, am.x); atomicAdd(dmask_token + c + 1, am.y); atomicAdd(dmask_token + c + 2, am.z); atomicAdd(dmask_token + c + 3, am.w);
  }
}

// out[b, c] = mean_{t=1..T-1} x[b, t, c]   (VisionTransformer.forward_features mean pooling, modeling_finetune.py:512-514)
__global__ void meanpool_kernel(const float* __restrict__ x, int B, int T, int C, float* __restrict__ out) {
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const int b = blockIdx.x;
  if (c >= C) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 1; t < T; ++t) {
    const float4 v = ld4(x + ((long long)b * T + t) * C + c);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const float inv = 1.0f / (T - 1);
  st4(out + (long long)b * C + c, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
}

// backward of meanpool: dx[b, 0, :] += 0 ; dx[b, t >= 1, :] += dpool[b, :] / (T - 1)
__global__ void meanpool_bwd_kernel(const float* __restrict__ dpool, int B, int T, int C, float* __restrict__ dx) {
  const int c4 = C >> 2;
  const long long total = (long long)B * (T - 1) * c4;
  const float inv = 1.0f / (T - 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const long long row = i / c4;
    const int t = (int)(row % (T - 1)) + 1;
    const int b = (int)(row / (T - 1));
    const float4 d = ld4(dpool + (long long)b * C + c);
    float* q = dx + ((long long)b * T + t) * C + c;
    float4 v = ld4(q);
    v.x += d.x * inv; v.y += d.y * inv; v.z += d.z * inv; v.w += d.w * inv;
    st4(q, v);
  }
}

int grid_for(long long work_items, int threads, int sms, int per_sm) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = (long long)sms * per_sm;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int b200vit_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_index, const float* gamma, const float* beta,
                                     float eps, int32_t rows, int32_t C, void* y_bf16, float* y_f32, float* mean, float* rstd,
                                     void* stream) {
  B200_CHECK_ARG(x != nullptr && rows >= 0 && C > 0 && C % 128 == 0 && C <= 128 * LN_MAXV, "layernorm_fwd: C=%d must be a multiple of 128, <= %d", C, 128 * LN_MAXV);
  B200_CHECK_ARG(y_bf16 != nullptr || y_f32 != nullptr, "layernorm_fwd: no output");
  if (rows == 0) return 0;
  const int grid = (rows + 7) / 8;
#define LN_FWD(NV) ln_fwd_kernel<NV><<<grid, 256, 0, STREAM>>>(x, ldx, row_index, gamma, beta, eps, rows, C, static_cast<bf16*>(y_bf16), y_f32, mean, rstd)
  switch (C / 128) {
    case 1: LN_FWD(1); break; case 2: LN_FWD(2); break; case 3: LN_FWD(3); break; case 4: LN_FWD(4); break;
    case 5: LN_FWD(5); break; case 6: LN_FWD(6); break; case 7: LN_FWD(7); break; default: LN_FWD(8); break;
  }
#undef LN_FWD
  B200_CHECK_LAUNCH("layernorm_fwd");
  return 0;
}

// B200VIT_LN_BWD_LAYOUT (same-box A/B): 0 = bulk-staged where it applies (default), 1 = register-staged column-owner kernel for dense rows,
// 2 = warp-per-row kernel everywhere
static int b200vit_ln_bwd_layout() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200VIT_LN_BWD_LAYOUT"); v = e != nullptr ? atoi(e) : 0; }
  return v;
}

static int launch_ln_bwd(const void* dy, int32_t dy_is_f32, const float* x, int64_t ldx, const int32_t* row_index, const float* gamma,
                         const float* mean, const float* rstd, int32_t rows, int32_t C, float* dx, int64_t lddx, float* dgamma, float* dbeta,
                         const SrbStage& srb, void* stream) {
  const int sms = b200vit_num_sms();
  const bool fused = srb.t != nullptr;
  const bool aligned = ((reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta) | reinterpret_cast<uintptr_t>(srb.dgamma2) |
                         reinterpret_cast<uintptr_t>(srb.dbias2)) & 15) == 0;          // red.global.add.v4 needs 16-byte aligned gradient rows
  const bool bulk_ok = row_index == nullptr && aligned && !dy_is_f32 && ldx == C && lddx == C && rows % LNB_G == 0 && C % 128 == 0 && C / 4 <= 256 &&
                       ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(mean) |
                         reinterpret_cast<uintptr_t>(rstd) | reinterpret_cast<uintptr_t>(srb.t)) & 15) == 0 && b200vit_ln_bwd_layout() == 0;
  if (bulk_ok) {
    LnbParams lp;
    lp.dy = static_cast<const bf16*>(dy); lp.x = x; lp.gamma = gamma; lp.mean = mean; lp.rstd = rstd; lp.dx = dx; lp.dgamma = dgamma; lp.dbeta = dbeta;
    lp.srb = srb; lp.rows = rows; lp.C = C;
    lp.stage_bytes = LNB_G * C * (fused ? 12 : 10) + 128;           // + mean / rstd (2 x 16 B), padded to keep every stage 128-byte aligned
    const size_t tail = (LNB_GROUPS * (8 * 2 * LNB_G + 2 * LNB_G) + 4 * C) * sizeof(float) + 2 * LNB_MAX_STAGES * 8 + 16;
    int nst = (int)((220 * 1024 - tail) / lp.stage_bytes);
    if (nst > LNB_MAX_STAGES) nst = LNB_MAX_STAGES;
    nst &= ~1;       // even: a stage then always belongs to the same consumer group, which therefore observes every phase of its full barrier
    lp.nstages = nst;
    const size_t smem = (size_t)nst * lp.stage_bytes + tail;
    const int nchunks = rows / LNB_G;
    const int bgrid = nchunks < sms ? nchunks : sms;
    const int threads = 32 + LNB_GROUPS * (C / 4);
    static bool cfg_done[2] = {false, false};
    if (!cfg_done[fused]) {
      cudaError_t e = fused ? cudaFuncSetAttribute(ln_bwd_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)
                            : cudaFuncSetAttribute(ln_bwd_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
      if (e != cudaSuccess) { b200vit_set_error("layernorm_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
      cfg_done[fused] = true;
    }
    if (nst >= 2) {
      if (fused) ln_bwd_bulk_kernel<true><<<bgrid, threads, smem, STREAM>>>(lp);
      else ln_bwd_bulk_kernel<false><<<bgrid, threads, smem, STREAM>>>(lp);
      B200_CHECK_LAUNCH("layernorm_bwd (bulk-staged)");
      return 0;
    }
  }
  if (row_index == nullptr && aligned && b200vit_ln_bwd_layout() != 2) {              // dense rows: column-owner kernel
    const int tx = C / 4;
    int ty = LNC_THREADS / tx;
    if (ty > LNC_TY_MAX) ty = LNC_TY_MAX;
    const int ngroups = (rows + ty * LNC_RB - 1) / (ty * LNC_RB);
    const int cgrid = ngroups < sms ? ngroups : sms;
    const size_t csmem = (LNC_TY_MAX * (8 * 2 * LNC_RB + 2 * LNC_RB) + 4 * C) * sizeof(float);
    const dim3 blk(tx, ty);
    if (fused) {
      if (dy_is_f32) ln_bwd_cols_kernel<float, true><<<cgrid, blk, csmem, STREAM>>>(static_cast<const float*>(dy), x, ldx, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb);
      else ln_bwd_cols_kernel<bf16, true><<<cgrid, blk, csmem, STREAM>>>(static_cast<const bf16*>(dy), x, ldx, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb);
    } else {
      if (dy_is_f32) ln_bwd_cols_kernel<float, false><<<cgrid, blk, csmem, STREAM>>>(static_cast<const float*>(dy), x, ldx, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb);
      else ln_bwd_cols_kernel<bf16, false><<<cgrid, blk, csmem, STREAM>>>(static_cast<const bf16*>(dy), x, ldx, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb);
    }
    B200_CHECK_LAUNCH("layernorm_bwd (column-owner)");
    return 0;
  }
  int grid = (rows + 7) / 8;
  if (grid > sms * 2) grid = sms * 2;
  const size_t smem = (fused ? 4 : 2) * C * sizeof(float);
#define LN_BWD(NV)                                                                                                                             \
  if (fused) {                                                                                                                                 \
    if (dy_is_f32) ln_bwd_kernel<NV, float, true><<<grid, 256, smem, STREAM>>>(static_cast<const float*>(dy), x, ldx, row_index, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb); \
    else ln_bwd_kernel<NV, bf16, true><<<grid, 256, smem, STREAM>>>(static_cast<const bf16*>(dy), x, ldx, row_index, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb);          \
  } else {                                                                                                                                     \
    if (dy_is_f32) ln_bwd_kernel<NV, float, false><<<grid, 256, smem, STREAM>>>(static_cast<const float*>(dy), x, ldx, row_index, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb); \
    else ln_bwd_kernel<NV, bf16, false><<<grid, 256, smem, STREAM>>>(static_cast<const bf16*>(dy), x, ldx, row_index, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, srb);          \
  }
  switch (C / 128) {
    case 1: LN_BWD(1); break; case 2: LN_BWD(2); break; case 3: LN_BWD(3); break; case 4: LN_BWD(4); break;
    case 5: LN_BWD(5); break; case 6: LN_BWD(6); break; case 7: LN_BWD(7); break; default: LN_BWD(8); break;
  }
#undef LN_BWD
  B200_CHECK_LAUNCH("layernorm_bwd");
  return 0;
}

extern "C" int b200vit_layernorm_bwd(const void* dy, int32_t dy_is_f32, const float* x, int64_t ldx, const int32_t* row_index,
                                     const float* gamma, const float* mean, const float* rstd, int32_t rows, int32_t C, float* dx,
                                     int64_t lddx, float* dgamma, float* dbeta, void* stream) {
  B200_CHECK_ARG(dy != nullptr && x != nullptr && mean != nullptr && rstd != nullptr && dx != nullptr, "layernorm_bwd: null pointer");
  B200_CHECK_ARG(C > 0 && C % 128 == 0 && C <= 128 * LN_MAXV, "layernorm_bwd: C=%d must be a multiple of 128, <= %d", C, 128 * LN_MAXV);
  if (rows == 0) return 0;
  SrbStage off = {};
  return launch_ln_bwd(dy, dy_is_f32, x, ldx, row_index, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, off, stream);
}

extern "C" int b200vit_layernorm_bwd_scale_residual(const void* dy, int32_t dy_is_f32, const float* x, int64_t ldx, const float* gamma,
                                                    const float* mean, const float* rstd, int32_t rows, int32_t C, float* dx, int64_t lddx,
                                                    float* dgamma, float* dbeta, const void* t_bf16, const float* rowscale,
                                                    int32_t rows_per_scale, const float* gamma2, void* dt_bf16, float* dgamma2, float* dbias2,
                                                    void* stream) {
  B200_CHECK_ARG(dy != nullptr && x != nullptr && mean != nullptr && rstd != nullptr && dx != nullptr, "layernorm_bwd_scale_residual: null pointer");
  B200_CHECK_ARG(t_bf16 != nullptr && dt_bf16 != nullptr, "layernorm_bwd_scale_residual: the fused stage needs t and dt");
  B200_CHECK_ARG(C > 0 && C % 128 == 0 && C <= 128 * LN_MAXV, "layernorm_bwd_scale_residual: C=%d must be a multiple of 128, <= %d", C, 128 * LN_MAXV);
  B200_CHECK_ARG(rowscale == nullptr || rows_per_scale > 0, "layernorm_bwd_scale_residual: rows_per_scale must be > 0");
  if (rows == 0) return 0;
  SrbStage st;
  st.t = static_cast<const bf16*>(t_bf16); st.rowscale = rowscale; st.rows_per_scale = rows_per_scale > 0 ? rows_per_scale : 1; st.gamma2 = gamma2;
  st.dt = static_cast<bf16*>(dt_bf16); st.dgamma2 = dgamma2; st.dbias2 = dbias2;
  return launch_ln_bwd(dy, dy_is_f32, x, ldx, nullptr, gamma, mean, rstd, rows, C, dx, lddx, dgamma, dbeta, st, stream);
}

extern "C" int b200vit_scale_residual_bwd(const float* dx, int64_t lddx, const void* t_bf16, const float* rowscale, int32_t rows_per_scale,
                                          const float* gamma, int32_t rows, int32_t C, void* dt_bf16, float* dgamma, float* dbias,
                                          void* stream) {
  B200_CHECK_ARG(dx != nullptr && dt_bf16 != nullptr, "scale_residual_bwd: null pointer");
  B200_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 4096, "scale_residual_bwd: C=%d must be a multiple of 4, <= 4096", C);
  B200_CHECK_ARG(rowscale == nullptr || rows_per_scale > 0, "scale_residual_bwd: rows_per_scale must be > 0");
  if (rows == 0) return 0;
  const int sms = b200vit_num_sms();
  const int tx = C / 4;
  int ty = 512 / tx;                       // ~512 threads per CTA, 3 CTAs per SM, two iterations of SRB_UNROLL rows in flight per thread
  if (ty < 1) ty = 1;
  if (ty > 8) ty = 8;
  int grid = (rows + ty * SRB_UNROLL - 1) / (ty * SRB_UNROLL);
  if (grid > sms * 3) grid = sms * 3;
  scale_residual_bwd_kernel<<<grid, dim3(tx, ty), 2 * C * sizeof(float), STREAM>>>(dx, lddx, static_cast<const bf16*>(t_bf16), rowscale, rows_per_scale > 0 ? rows_per_scale : 1,
                                                          gamma, rows, C, static_cast<bf16*>(dt_bf16), dgamma, dbias);
  B200_CHECK_LAUNCH("scale_residual_bwd");
  return 0;
}

extern "C" int b200vit_colsum_bf16(const void* x, int64_t ldx, int32_t rows, int32_t C, float* out, void* stream) {
  B200_CHECK_ARG(x != nullptr && out != nullptr && C > 0 && C % 8 == 0 && C <= 8192 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                 "colsum_bf16: C and ldx must be multiples of 8 (16-byte rows), C <= 8192 (C=%d)", C);
  if (rows == 0) return 0;
  const int sms = b200vit_num_sms();
  const int tx = C / 8;
  int ty = 768 / tx;
  if (ty < 1) ty = 1;
  if (ty > 8) ty = 8;
  int grid = (rows + ty * RB_UNROLL - 1) / (ty * RB_UNROLL);
  if (grid > sms * 2) grid = sms * 2;
  colsum_bf16_kernel<<<grid, dim3(tx, ty), C * sizeof(float), STREAM>>>(static_cast<const bf16*>(x), ldx, rows, C, out);
  B200_CHECK_LAUNCH("colsum_bf16");
  return 0;
}

extern "C" int b200vit_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  B200_CHECK_ARG(src != nullptr && dst != nullptr && n >= 0, "cast: bad arguments");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "cast: unaligned pointers");
  if (n == 0) return 0;
  const int sms = b200vit_num_sms();
  cast_bf16_kernel<<<grid_for(n / 4 + 1, 256, sms, 8), 256, 0, STREAM>>>(src, static_cast<bf16*>(dst), n / 4, n);
  B200_CHECK_LAUNCH("cast_f32_to_bf16");
  return 0;
}

extern "C" int b200vit_im2col_patches(const float* img, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t P, void* out_bf16, void* stream) {
  B200_CHECK_ARG(img != nullptr && out_bf16 != nullptr && B > 0 && P % 4 == 0 && H % P == 0 && W % P == 0 && W % 4 == 0, "im2col: bad arguments");
  const int sms = b200vit_num_sms();
  const long long total4 = (long long)B * Cin * H * W / 4;
  im2col_kernel<<<grid_for(total4, 256, sms, 8), 256, 0, STREAM>>>(img, B, Cin, H, W, P, static_cast<bf16*>(out_bf16));
  B200_CHECK_LAUNCH("im2col_patches");
  return 0;
}

extern "C" int b200vit_assemble_tokens(const float* pe, const float* cls, const float* mask_token, const uint8_t* mask, const float* pos_embed,
                                       int32_t B, int32_t np, int32_t C, float* x, void* stream) {
  B200_CHECK_ARG(pe != nullptr && cls != nullptr && x != nullptr && C % 4 == 0, "assemble_tokens: bad arguments");
  B200_CHECK_ARG(mask == nullptr || mask_token != nullptr, "assemble_tokens: mask without mask_token");
  const int sms = b200vit_num_sms();
  assemble_tokens_kernel<<<grid_for((long long)B * (np + 1) * (C / 4), 256, sms, 8), 256, 0, STREAM>>>(pe, cls, mask_token, mask, pos_embed, B, np, C, x);
  B200_CHECK_LAUNCH("assemble_tokens");
  return 0;
}

extern "C" int b200vit_assemble_tokens_bwd(const float* dx, const uint8_t* mask, int32_t B, int32_t np, int32_t C, void* dpe_bf16, float* dcls,
                                           float* dmask_token, float* dpos_embed, void* stream) {
  B200_CHECK_ARG(dx != nullptr && dpe_bf16 != nullptr && C % 4 == 0, "assemble_tokens_bwd: bad arguments");
  const int threads = 64;
  dim3 grid(B, (C / 4 + threads - 1) / threads);
  assemble_tokens_bwd_kernel<<<grid, threads, 0, STREAM>>>(dx, mask, B, np, C, static_cast<bf16*>(dpe_bf16), dcls, dmask_token, dpos_embed);
  B200_CHECK_LAUNCH("assemble_tokens_bwd");
  return 0;
}

extern "C" int b200vit_meanpool_tokens(const float* x, int32_t B, int32_t T, int32_t C, float* out, void* stream) {
  B200_CHECK_ARG(x != nullptr && out != nullptr && T > 1 && C % 4 == 0, "meanpool_tokens: bad arguments");
  const int threads = 64;
  dim3 grid(B, (C / 4 + threads - 1) / threads);
  meanpool_kernel<<<grid, threads, 0, STREAM>>>(x, B, T, C, out);
  B200_CHECK_LAUNCH("meanpool_tokens");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Stochastic depth keep/keep_prob per (layer, draw, sample): timm drop_path semantics (modeling_finetune.py:51-62),
// uniform numbers from Philox4x32-10 keyed on (seed; layer, draw, sample) instead of torch's global generator.
// ------------------------------------------------------------------------------------------------
struct DropPathProbs { float p[64]; };
__global__ void drop_path_scales_kernel(DropPathProbs probs, int L, int draws, int B, unsigned long long seed, float* __restrict__ out) {
  const int total = L * draws * B;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i % B, d = (i / B) % draws, l = i / (B * draws);
    const float p = probs.p[l];
    const Philox4 r = philox4x32_10((uint32_t)b, (uint32_t)d, (uint32_t)l, 0xD509A7u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u = (r.x >> 8) * (1.0f / 16777216.0f);    // [0,1)
    out[i] = (p <= 0.f) ? 1.0f : (u >= p ? 1.0f / (1.0f - p) : 0.0f);
  }
}

extern "C" int b200vit_meanpool_tokens_bwd(const float* dpool, int32_t B, int32_t T, int32_t C, float* dx, void* stream) {
  B200_CHECK_ARG(dpool != nullptr && dx != nullptr && T > 1 && C % 4 == 0, "meanpool_tokens_bwd: bad arguments");
  const int sms = b200vit_num_sms();
  meanpool_bwd_kernel<<<grid_for((long long)B * (T - 1) * (C / 4), 256, sms, 8), 256, 0, STREAM>>>(dpool, B, T, C, dx);
  B200_CHECK_LAUNCH("meanpool_tokens_bwd");
  return 0;
}

extern "C" int b200vit_drop_path_scales(const float* probs_host, int32_t L, int32_t draws, int32_t B, uint64_t seed, float* out, void* stream) {
  B200_CHECK_ARG(probs_host != nullptr && out != nullptr && L > 0 && L <= 64 && draws > 0 && B > 0, "drop_path_scales: bad arguments (L <= 64)");
  DropPathProbs pr;
  for (int i = 0; i < L; ++i) {
    B200_CHECK_ARG(probs_host[i] >= 0.f && probs_host[i] < 1.f, "drop_path_scales: p[%d] out of range", i);
    pr.p[i] = probs_host[i];
  }
  const int total = L * draws * B;
  drop_path_scales_kernel<<<(total + 255) / 256, 256, 0, STREAM>>>(pr, L, draws, B, seed, out);
  B200_CHECK_LAUNCH("drop_path_scales");
  return 0;
}
