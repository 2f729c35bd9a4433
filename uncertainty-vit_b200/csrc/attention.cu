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

// Relative-position-bias table gradient: dtable[index[i, j], h] += sum_b dS[b, h, i, j]  (dS^T stored [B, H, j, ld] by attn_bwd)
__global__ void __launch_bounds__(256) relbias_grad_kernel(const bf16* __restrict__ ds, int B, int H, int N, int ld,
                                                           const int* __restrict__ rel_index, float* __restrict__ dtable) {
  const int half = ld >> 1;                       // pairs of queries
  const long long total = (long long)H * N * half;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int ip = (int)(t % half);
    const int j = (int)((t / half) % N);
    const int h = (int)(t / ((long long)half * N));
    const int i = ip * 2;
    if (i >= N) continue;
    float s0 = 0.f, s1 = 0.f;
    const bf16* src = ds + ((long long)h * N + j) * ld + i;
    const long long bstride = (long long)H * N * ld;
    int b = 0;
    for (; b + 8 <= B; b += 8) {     // 8 independent loads in flight
      uint32_t w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = *reinterpret_cast<const uint32_t*>(src + (long long)(b + u) * bstride);
#pragma unroll
      for (int u = 0; u < 8; ++u) { const float2 v = unpack_bf16x2(w[u]); s0 += v.x; s1 += v.y; }
    }
    for (; b < B; ++b) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + (long long)b * bstride));
      s0 += v.x; s1 += v.y;
    }
    atomicAdd(dtable + (long long)rel_index[i * N + j] * H + h, s0);
    if (i + 1 < N) atomicAdd(dtable + (long long)rel_index[(i + 1) * N + j] * H + h, s1);
  }
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

// internal (not part of the public ABI): shared by the dual-stream backward in wattention.cu
int b200vit_relbias_grad_launch(const void* ds_work, int B, int H, int N, int ld_ds, const int32_t* rel_index, float* dtable, void* stream) {
  const int sms = b200vit_num_sms();
  relbias_grad_kernel<<<sms * 8, 256, 0, STREAM>>>(static_cast<const bf16*>(ds_work), B, H, N, ld_ds, rel_index, dtable);
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
