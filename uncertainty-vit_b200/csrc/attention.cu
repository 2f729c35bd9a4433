// Helpers of the attention path that are not tensor-core work (the tcgen05 forward / backward kernels live in attention_sm100.cu):
//   dropout_mask_kernel  : materialises the Philox keep mask of the forward (parity tests inject the SAME mask into the CPU oracle)
//   rel_pos_bias_kernel  : RelativePositionBias.forward (modeling_finetune.py:359-364) in the padded, log2(e)-prescaled layouts the
//                          attention kernels read (key mask baked into the -inf padding; transposed copy for the backward)
//   relbias_grad_kernel  : dtable[index[i, j], h] += sum_b dS[b, h, i, j] from the dS^T workspace of the backward, scatter-added through
//                          the reference's relative_position_index (modeling_finetune.py:339-353)
#include "../../include/b200vit.h"
#include "attn_common.cuh"

namespace {

using namespace attn;

// materialises the Philox keep mask as uint8 [B,H,N,N] (tests: inject the SAME mask into the CPU oracle)
__global__ void dropout_mask_kernel(uint8_t* out, int BH, int N, float p_drop, uint64_t seed, uint32_t stream_id) {
  const uint32_t thresh = (uint32_t)(p_drop * 65536.0f + 0.5f);
  const long long total = (long long)BH * N * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % N);
    const int i = (int)((idx / N) % N);
    const int bh = (int)(idx / ((long long)N * N));
    const int jt = j >> 3, quad = (j & 7) >> 1, lo = j & 1;
    const Philox4 r = dropout_group(seed, stream_id, bh, i, quad, jt >> 2);
    const int v = (jt & 3) * 2 + lo;
    const uint32_t w = v < 4 ? (v < 2 ? r.x : r.y) : (v < 6 ? r.z : r.w);
    out[idx] = ((v & 1) ? (w >> 16) : (w & 0xffffu)) >= thresh ? 1 : 0;
  }
}

// Packed keep mask of one attention layer, [B*H, N, 8 words] (bit e of word c = key 32c + e of that query row), from the Philox stream of
// dropout_mask_kernel or from an injected uint8 mask. One thread per (row, 32-key word). Generating the mask here — a 20 us, fully
// occupied kernel — instead of inside the attention forward takes ~11 instructions per decision off the one-thread-per-row softmax loop,
// which is latency-bound (the forward with dropout was 129 us against 87 us without).
__global__ void __launch_bounds__(256) keep_bits_kernel(uint32_t* __restrict__ keep_bits, int BH, int N, uint32_t thresh, uint64_t seed,
                                                        const uint64_t* __restrict__ seed_dev, uint32_t stream_id, const uint8_t* __restrict__ keep_in) {
  const long long total = (long long)BH * N * 8;
  const uint64_t sd = seed_dev != nullptr ? *seed_dev : seed;
  const int words = (N + 31) >> 5;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 7);
    const long long row = idx >> 3;              // bh * N + i
    uint32_t w = 0u;
    if (c < words) {
      const int i = (int)(row % N);
      const int bh = (int)(row / N);
      if (keep_in == nullptr) {
        // word n4 of a Philox draw holds the two 16-bit lanes of keys (n4 * 8 + quad * 2, + 1): both compared at once (vset2), the two
        // result bits (0 and 16) folded into adjacent bit positions
        const uint32_t th2 = thresh | (thresh << 16);
#pragma unroll
        for (int quad = 0; quad < 4; ++quad) {
          const Philox4 r = dropout_group(sd, stream_id, bh, i, quad, c);
          const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int n4 = 0; n4 < 4; ++n4) {
            const uint32_t ge = __vcmpgeu2(rw[n4], th2) & 0x00010001u;      // bit 0: low lane kept, bit 16: high lane kept
            w |= ((ge | (ge >> 15)) & 3u) << (n4 * 8 + quad * 2);
          }
        }
      } else {
        const uint8_t* src = keep_in + row * N;
        for (int e = 0; e < 32; ++e) {
          const int j = c * 32 + e;
          if (j < N && src[j]) w |= 1u << e;
        }
      }
    }
    keep_bits[idx] = w;
  }
}

// RelativePositionBias.forward (modeling_finetune.py:359-364) in the layout the attention kernels read:
//   out_fwd[h, i, j] = scale * table[index[i, j], h] for j < N, -inf for N <= j < ld     (forward: key mask baked in)
//   out_bwd[h, j, i] = scale * table[index[i, j], h] for i < N, 0 for N <= i < ld        (backward: transposed)
__global__ void rel_pos_bias_kernel(const float* __restrict__ table, const int* __restrict__ index, int N, int H, int ld, float scale,
                                    float* __restrict__ out_fwd, float* __restrict__ out_bwd) {
  const int total = H * N * ld;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int c = t % ld, r = (t / ld) % N, h = t / (ld * N);
    if (out_fwd != nullptr) out_fwd[t] = c < N ? scale * __ldg(table + (long long)index[r * N + c] * H + h) : -INFINITY;
    if (out_bwd != nullptr) out_bwd[t] = c < N ? scale * __ldg(table + (long long)index[c * N + r] * H + h) : 0.f;
  }
}

// Indexed form of the bias for the forward kernel (attention_sm100.cu, BIAS == 2): tab[h][k] = scale * table[k][h] with tab[h][nbins] = -inf
// (the key mask), and the relative_position_index as uint16 tiles [query tile][128 rows][IDX_PITCH] (row pitch 105 words: conflict-free for
// one-thread-per-row reads; columns >= N point at the -inf entry). 52 KB per query tile stay resident in shared memory for the whole
// kernel instead of 106 KB of fp32 bias per (batch, head, query tile) streaming through a TMA ring.
__global__ void rel_pos_index_tiles_kernel(const float* __restrict__ table, const int* __restrict__ index, int N, int H, int nbins, float scale,
                                           int tab_pitch, float* __restrict__ tab, int m_tiles, uint16_t* __restrict__ idx16) {
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int t = t0; t < H * tab_pitch; t += stride) {
    const int k = t % tab_pitch, h = t / tab_pitch;
    tab[t] = k < nbins ? scale * __ldg(table + (long long)k * H + h) : (k == nbins ? -INFINITY : 0.f);
  }
  const int per_tile = 128 * B200VIT_ATTN_IDX_PITCH;
  for (int t = t0; t < m_tiles * per_tile; t += stride) {
    const int j = t % B200VIT_ATTN_IDX_PITCH, r = (t / B200VIT_ATTN_IDX_PITCH) % 128, mt = t / per_tile;
    const int i = mt * 128 + r;
    idx16[t] = (uint16_t)(j >= N ? nbins : (i < N ? index[i * N + j] : 0));
  }
}

// rowmax[h, i] = max_j scale * table[index[i, j], h]: the softmax stabiliser of the single-pass Wasserstein attention forward
__global__ void rel_pos_bias_rowmax_kernel(const float* __restrict__ table, const int* __restrict__ index, int N, int H, float scale,
                                           float* __restrict__ rowmax) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * N) return;
  const int i = t % N, h = t / N;
  float m = -INFINITY;
  for (int j = 0; j < N; ++j) m = fmaxf(m, scale * __ldg(table + (long long)index[i * N + j] * H + h));
  rowmax[t] = m;
}

// Relative-position-bias table gradient: dtable[index[i, j], h] += sum_b dS[b, h, i, j]  (dS^T stored [B, H, j, ld] by attn_bwd).
// One pass over the 126 MB workspace of a ViT-B layer: a thread owns 8 consecutive queries of one (h, key) row (16-byte loads, 8 batches in
// flight), RB_SPLIT threads share the batch dimension and meet in shared memory, so the scatter issues one atomic per (h, i, j).
constexpr int RB_ITEMS = 64, RB_SPLIT = 4;
__global__ void __launch_bounds__(RB_ITEMS * RB_SPLIT) relbias_grad_kernel(const bf16* __restrict__ ds, int B, int H, int N, int ld,
                                                                          const int* __restrict__ rel_index, float* __restrict__ dtable) {
  __shared__ float part[RB_SPLIT][RB_ITEMS][8];
  const int vecs = ld >> 3;                       // 16-byte vectors per (h, key) row
  const long long total = (long long)H * N * vecs;
  const long long t = (long long)blockIdx.x * RB_ITEMS + threadIdx.x;
  const int q = threadIdx.y;
  const int v = (int)(t % vecs);
  const int j = (int)((t / vecs) % N);
  const int h = (int)(t / ((long long)vecs * N));
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (t < total && v * 8 < N) {
    const int per = (B + RB_SPLIT - 1) / RB_SPLIT;
    const int b0 = q * per, b1 = min(B, b0 + per);
    const long long bstride = (long long)H * N * ld;
    const bf16* src = ds + ((long long)h * N + j) * ld + v * 8;
    int b = b0;
    for (; b + 8 <= b1; b += 8) {
      uint4 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = __ldg(reinterpret_cast<const uint4*>(src + (long long)(b + u) * bstride));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float2 f0 = unpack_bf16x2(w[u].x), f1 = unpack_bf16x2(w[u].y), f2 = unpack_bf16x2(w[u].z), f3 = unpack_bf16x2(w[u].w);
        acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y; acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
      }
    }
    for (; b < b1; ++b) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(src + (long long)b * bstride));
      const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z), f3 = unpack_bf16x2(w.w);
      acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y; acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[q][threadIdx.x][e] = acc[e];
  __syncthreads();
  // 512 (item, query) sums per CTA, two per thread
  for (int k = threadIdx.y * RB_ITEMS + threadIdx.x; k < RB_ITEMS * 8; k += RB_ITEMS * RB_SPLIT) {
    const int it = k >> 3, e = k & 7;
    const long long tt = (long long)blockIdx.x * RB_ITEMS + it;
    if (tt >= total) continue;
    const int vv = (int)(tt % vecs), jj = (int)((tt / vecs) % N), hh = (int)(tt / ((long long)vecs * N));
    const int i = vv * 8 + e;
    if (i >= N) continue;
    float s = 0.f;
#pragma unroll
    for (int qq = 0; qq < RB_SPLIT; ++qq) s += part[qq][it][e];
    atomicAdd(dtable + (long long)rel_index[i * N + jj] * H + hh, s);
  }
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// internal (not part of the public ABI): shared by the dual-stream backward in wattention.cu
int b200vit_relbias_grad_launch(const void* ds_work, int B, int H, int N, int ld_ds, const int32_t* rel_index, float* dtable, void* stream) {
  B200_CHECK_ARG(ld_ds % 8 == 0 && (reinterpret_cast<uintptr_t>(ds_work) & 15) == 0, "relbias_grad: dS^T rows must be 16-byte aligned (ld %% 8 == 0)");
  const long long items = (long long)H * N * (ld_ds / 8);
  relbias_grad_kernel<<<(unsigned)((items + RB_ITEMS - 1) / RB_ITEMS), dim3(RB_ITEMS, RB_SPLIT), 0, STREAM>>>(static_cast<const bf16*>(ds_work), B, H, N, ld_ds,
                                                                                                              rel_index, dtable);
  B200_CHECK_LAUNCH("relbias_grad");
  return 0;
}

// internal: shared by the dot-product and the Wasserstein attention forwards
int b200vit_keep_bits_launch(uint8_t* keep_bits, int BH, int N, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id,
                             const uint8_t* keep_in, void* stream) {
  const int sms = b200vit_num_sms();
  keep_bits_kernel<<<sms * 8, 256, 0, STREAM>>>(reinterpret_cast<uint32_t*>(keep_bits), BH, N, (uint32_t)(p_drop * 65536.0f + 0.5f), seed, seed_dev,
                                                stream_id, keep_in);
  B200_CHECK_LAUNCH("keep_bits");
  return 0;
}

extern "C" int b200vit_keep_bits(uint8_t* keep_bits, int32_t BH, int32_t N, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id,
                                 const uint8_t* keep_in, void* stream) {
  B200_CHECK_ARG(keep_bits != nullptr && BH > 0 && N > 0 && N <= NMAX && p_drop >= 0.f && p_drop < 1.f, "keep_bits: bad arguments");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(keep_bits) & 3) == 0, "keep_bits: the mask must be 4-byte aligned");
  return b200vit_keep_bits_launch(keep_bits, BH, N, p_drop, seed, seed_dev, stream_id, keep_in, stream);
}

extern "C" int b200vit_dropout_mask(uint8_t* out, int32_t BH, int32_t N, float p_drop, uint64_t seed, uint32_t stream_id, void* stream) {
  B200_CHECK_ARG(out != nullptr && BH > 0 && N > 0, "dropout_mask: bad arguments");
  const int sms = b200vit_num_sms();
  dropout_mask_kernel<<<sms * 8, 256, 0, STREAM>>>(out, BH, N, p_drop, seed, stream_id);
  B200_CHECK_LAUNCH("dropout_mask");
  return 0;
}

extern "C" int b200vit_rel_pos_index_tiles(const float* table, const int32_t* index, int32_t N, int32_t H, int32_t nbins, float scale, float* tab_out,
                                           uint16_t* idx16_out, void* stream) {
  B200_CHECK_ARG(table != nullptr && index != nullptr && tab_out != nullptr && idx16_out != nullptr && N > 0 && N <= 208 && H > 0 && nbins > 0 &&
                     nbins < B200VIT_ATTN_TAB_MAX, "rel_pos_index_tiles: bad arguments (N <= 208, nbins < %d)", B200VIT_ATTN_TAB_MAX);
  const int tab_pitch = (nbins + 1 + 3) & ~3;
  const int m_tiles = (N + 127) / 128;
  rel_pos_index_tiles_kernel<<<64, 256, 0, STREAM>>>(table, index, N, H, nbins, scale, tab_pitch, tab_out, m_tiles, idx16_out);
  B200_CHECK_LAUNCH("rel_pos_index_tiles");
  return 0;
}

extern "C" int b200vit_rel_pos_bias(const float* table, const int32_t* index, int32_t N, int32_t H, int32_t ld, float scale, float* out_fwd,
                                    float* out_bwd_t, float* rowmax_fwd, void* stream) {
  B200_CHECK_ARG(table != nullptr && index != nullptr && (out_fwd != nullptr || out_bwd_t != nullptr || rowmax_fwd != nullptr) && N > 0 && H > 0 &&
                     ld >= N, "rel_pos_bias: bad arguments");
  const int sms = b200vit_num_sms();
  if (out_fwd != nullptr || out_bwd_t != nullptr) {
    rel_pos_bias_kernel<<<sms * 4, 256, 0, STREAM>>>(table, index, N, H, ld, scale, out_fwd, out_bwd_t);
    B200_CHECK_LAUNCH("rel_pos_bias");
  }
  if (rowmax_fwd != nullptr) {
    rel_pos_bias_rowmax_kernel<<<(H * N + 127) / 128, 128, 0, STREAM>>>(table, index, N, H, scale, rowmax_fwd);
    B200_CHECK_LAUNCH("rel_pos_bias_rowmax");
  }
  return 0;
}
