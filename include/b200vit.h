/*
 * libb200vit — C ABI of the B200-native (sm_100a) ViT-B/16 / ViT-L/16 data2vec + uncertainty hot path.
 *
 * Every entry point replaces a group of ATen library calls that the reference
 * (fx-erick/uncertainty-vit, pure PyTorch) issues on its hot path; the reference call site is cited
 * beside each declaration as file:line into the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - the CALLER owns all memory (including workspaces); kernels never allocate or synchronise;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, <0 for a bad argument, >0 = cudaError_t of a failed launch;
 *     b200vit_last_error() returns a thread-local message for the last non-zero return;
 *   - bf16 tensors are raw uint16 storage (__nv_bfloat16), row-major, leading dimensions in ELEMENTS.
 */
#ifndef B200VIT_H_
#define B200VIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VIT_ABI_VERSION 8   /* 2: seed_dev in attn_fwd / wattn_fwd, caller-owned attention-backward workspace;
                                   3: n_valid_dev in d2v_target_loss / wasserstein_loss (padded row lists), block_masks, mixup_batch, normalize_u8;
                                   4: d2v_target_loss_ex, channel_stats, column_std, mask_dropout, gaussian_sample, tace_auroc, finetune_loss;
                                   5: tcgen05 Wasserstein attention (wattn_fwd / wattn_bwd take the transformed-operand workspace and the bias
                                      row maxima), rel_pos_bias rowmax output;
                                   6: keep_bits (the dropout mask as its own kernel), keep_ready in attn_fwd / wattn_fwd, set_sm_limit;
                                   7: layernorm_bwd_scale_residual;
                                   8: indexed relative-position bias in attn_fwd (rel_pos_index_tiles) */

const char* b200vit_last_error(void);
int b200vit_abi_version(void);
/* number of SMs of the current device (148 on B200); <0 when no CUDA device is usable */
int b200vit_device_sm_count(void);
/* Caps the number of SMs the library's persistent kernels size their grids for (0 = no cap; rounded down to an even number); returns the
 * previous cap. The data-parallel engine lowers it while the NCCL gradient all-reduce of the upper layers runs beside the backward pass of
 * the lower ones (replaces DDP's bucketed overlap, run_cyclical.py:515-519): a persistent kernel with a static tile schedule that cannot get
 * all its CTAs resident at once would otherwise run its last CTAs as a second wave. Process-wide, not thread-safe by design (one process
 * per GPU). b200vit_device_sm_count() reports the capped value. */
int b200vit_set_sm_limit(int32_t n);

/* ------------------------------------------------------------------------------------------------
 * GEMM (tcgen05.mma, TMEM accumulators, TMA-fed, persistent, warp-specialised).
 *   D[M,N] = A[M,K] * B[N,K]^T  with bf16 operands and fp32 accumulation.
 * Replaces F.linear / nn.Linear in Attention.forward (modeling_finetune.py:149,186),
 * Mlp.forward (modeling_finetune.py:76-81), lm_head (modeling_cyclical.py:219-225),
 * PatchEmbed conv-as-GEMM (modeling_finetune.py:324), head (modeling_finetune.py:522)
 * and the autograd backward of each (dgrad: a/b *_mn_major; wgrad: both mn-major + split_k).
 * ---------------------------------------------------------------------------------------------- */
enum b200vit_epilogue {
  B200VIT_EPI_BF16 = 0,       /* out_bf16 = (acc + bias[n]) * colscale[n]           (bias/colscale optional) */
  B200VIT_EPI_GELU = 1,       /* t = acc + bias; out_bf16 = gelu_erf(t); out2_bf16 = gelu_erf'(t) (optional)  */
  B200VIT_EPI_RESIDUAL = 2,   /* t = acc + bias; out2_bf16 = t (optional);
                                 out_f32 = residual + rowscale[m / rows_per_scale] * colscale[n] * t         */
  B200VIT_EPI_DGELU = 3,      /* out_bf16 = acc * aux[m,n], aux = the gelu_erf'(t) saved by EPI_GELU's out2  */
  B200VIT_EPI_F32 = 4,        /* out_f32 = acc + bias                                                        */
  B200VIT_EPI_F32_ATOMIC = 5, /* out_f32 += alpha * acc  (red.global.add.v4; required with split_k)          */
  B200VIT_EPI_ELU1 = 6        /* out_bf16 = elu(acc + bias) + 1   (cov-stream QKV, modeling_finetune_dist.py:127)    */
};

typedef struct b200vit_gemm_desc {
  int32_t M, N, K;
  const void* A;       /* a_mn_major == 0: [M, K] row-major (lda >= K);  1: stored as [K, M] row-major (lda >= M) */
  int64_t lda;
  int32_t a_mn_major;
  const void* B;       /* b_mn_major == 0: [N, K] row-major (nn.Linear weight);  1: stored as [K, N] row-major    */
  int64_t ldb;
  int32_t b_mn_major;
  int32_t epilogue;    /* enum b200vit_epilogue */
  const float* bias;   /* [N] or NULL */
  const float* colscale; /* [N] or NULL */
  const float* rowscale; /* [ceil(M / rows_per_scale)] or NULL (drop-path keep/keep_prob per sample) */
  int32_t rows_per_scale;
  const float* residual; /* fp32 [M, ld_residual] */
  int64_t ld_residual;
  const void* aux;     /* bf16 [M, ld_aux] */
  int64_t ld_aux;
  float* out_f32;
  int64_t ld_f32;
  void* out_bf16;
  int64_t ld_bf16;
  void* out2_bf16;
  int64_t ld2_bf16;
  float alpha;
  int32_t split_k;     /* F32_ATOMIC only: 0 = pick for wave efficiency, 1 = none, >1 = explicit number of K splits */
  int32_t max_ctas;    /* 0: one CTA per SM */
  int32_t debug_flags; /* 0 in production; bit 0 = skip the epilogue's global I/O (roofline experiments only) */
  float* colsum;       /* optional fp32 [N]: += column sums of the values written (fused bias gradient, e.g. fc1.bias from dGELU) */
} b200vit_gemm_desc;

int b200vit_gemm_bf16(const b200vit_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention on tcgen05 / TMEM (scores and probabilities never leave tensor memory between the two GEMMs; TMA 3-D boxes of the
 * un-permuted QKV GEMM output; one thread per score row, no shuffles; Philox4x32-7 dropout).
 * Replaces Attention.forward lines 155-185 (modeling_finetune.py): q*scale, q k^T, + rel_pos_bias, softmax,
 * attn_drop, attn @ v, transpose/reshape — and its autograd backward.
 *   qkv   : bf16 [B, N, 3, H, 64] (the QKV GEMM output, no permute copy), 16-byte aligned
 *   bias  : fp32 [H, N, ld_bias] = log2(e) * rel_pos_bias, -inf in columns [N, ld_bias) (b200vit_rel_pos_bias out_fwd), or NULL;
 *           ld_bias % 4 == 0, 16-byte aligned (it is read through a TMA tensor map)
 *   out : bf16 [B, N, H*64]      lse : fp32 [B, H, N]
 *   keep_bits : packed dropout keep mask [B, H, N, 32] bytes (bit j%8 of byte j/8), written by fwd, read by bwd
 *   keep_in   : optional injected keep mask uint8 [B, H, N, N]; NULL = Philox4x32-7 keyed on (seed, stream_id)
 *   seed_dev  : optional DEVICE pointer to the 64-bit key; when non-NULL it replaces `seed` (a captured CUDA graph of the training
 *               step stays valid across steps: the host rewrites *seed_dev before every replay)
 * N <= 208, head_dim == 64.
 * ---------------------------------------------------------------------------------------------- */
int b200vit_attn_fwd(const void* qkv, const float* bias, int64_t ld_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim,
                     float scale, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id, const uint8_t* keep_in, void* out,
                     float* lse, uint8_t* keep_bits, int32_t keep_ready, const uint16_t* bias_idx16, const float* bias_tab, int32_t nbins,
                     void* stream);
/* The relative-position bias in INDEXED form (RelativePositionBias.forward gathers table[index], modeling_finetune.py:359-364): when bias_idx16
 * != NULL b200vit_attn_fwd ignores `bias` and gathers log2(e) * table[index[i, j], h] itself — the uint16 index tile of a 128-query tile
 * (52 KB) stays resident in shared memory for the whole kernel and the head's table row (3 KB) is reloaded per item, instead of 106 KB of fp32
 * bias per (batch, head, query tile) streaming through a TMA ring (25 % of the forward's stall samples were waits for that ring).
 *   tab_out   : fp32 [H, tab_pitch], tab_pitch = (nbins + 1) rounded up to 4; entry nbins = -inf (key mask)
 *   idx16_out : uint16 [ceil(N / 128), 128, B200VIT_ATTN_IDX_PITCH]; columns >= N hold nbins
 * N <= 208, nbins < B200VIT_ATTN_TAB_MAX. Values are bit-identical to the dense bias of b200vit_rel_pos_bias. */
#define B200VIT_ATTN_IDX_PITCH 210
#define B200VIT_ATTN_TAB_MAX 1024
int b200vit_rel_pos_index_tiles(const float* table, const int32_t* index, int32_t N, int32_t H, int32_t nbins, float scale, float* tab_out,
                                uint16_t* idx16_out, void* stream);
/* The packed keep mask of one attention layer, [B*H, N, 32] bytes (bit j%8 of byte j/8 of row i = keep(i, j)), from Philox4x32-7 keyed on
 * (seed or *seed_dev, stream_id) or from an injected uint8 [B*H, N, N] mask. b200vit_attn_fwd / b200vit_wattn_fwd run it themselves unless
 * they are called with keep_ready != 0 — then keep_bits must already hold the mask: it depends only on the key, so the engine draws the masks
 * of all layers of a step on a side stream, next to the EMA-teacher forward (the draw costs ~40 us per ViT-B layer, integer-multiply bound). */
int b200vit_keep_bits(uint8_t* keep_bits, int32_t BH, int32_t N, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id,
                      const uint8_t* keep_in, void* stream);
/* Bytes of the caller-owned workspace of b200vit_attn_bwd: dS^T bf16 [B, H, N, ld_ds] | D fp32 [B, H, N] | transposed keep bits. */
size_t b200vit_attn_bwd_workspace_bytes(int32_t B, int32_t H, int32_t N);
/* bias_t: the TRANSPOSED padded bias (b200vit_rel_pos_bias out_bwd_t) or NULL.
 * dqkv: bf16 [B, N, 3, H, 64] (fully overwritten). work: 256-byte aligned workspace of b200vit_attn_bwd_workspace_bytes() bytes;
 * ld_ds = N rounded up to 16. The key-tile kernel leaves dS^T as bf16 [B, H, N(key), ld_ds(query)] at the start of `work`; the dQ
 * kernel reads it back, and when dtable != NULL (the gradient of relative_position_bias_table [num_bins, H], +=) a reduction kernel
 * sums it over the batch and scatter-adds through rel_index int32 [N, N] (the reference's relative_position_index,
 * modeling_finetune.py:339-353). dq_bias / dv_bias (optional, +=, [H*64]) are the q_bias / v_bias gradients (column sums of
 * dQ / dV; modeling_finetune.py:148). */
int b200vit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* bias_t, int64_t ld_bias,
                     const uint8_t* keep_bits, void* work, int32_t ld_ds, const int32_t* rel_index, float* dtable,
                     float* dq_bias, float* dv_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop,
                     void* dqkv, void* stream);
/* Wasserstein-distance attention of the dual-stream (--stochastic) model on tcgen05 / TMEM: dist Attention.forward
 * (modeling_finetune_dist.py:111-179) with wasserstein_distance_matmul (uncertainty_evaluations.py:276-294). qkv_mean / qkv_cov: bf16
 * [B, N, 3, H, 64]; qkv_cov holds elu(.)+1 already (GEMM epilogue B200VIT_EPI_ELU1). bias (required) and bias_rowmax ([H, N] = max_j bias) in the
 * layouts of b200vit_rel_pos_bias (log2(e)-scaled). out_mean = P~ v, out_cov = (P~)^2 cv, both bf16 [B, N, H*64].
 * xwork: caller-owned, 256-byte aligned, b200vit_wattn_workspace_bytes(B, H, N) bytes: the forward stores the transformed operands
 * [sigmoid(scale q) | sqrt(sigmoid(cq))], [sigmoid(k) | sqrt(sigmoid(ck))] (bf16 [B, N, 2, H, 128]) and their row norms there; the backward of the
 * same forward must be given the same (unmodified) buffer. */
size_t b200vit_wattn_workspace_bytes(int32_t B, int32_t H, int32_t N);
int b200vit_wattn_fwd(const void* qkv_mean, const void* qkv_cov, const float* bias, int64_t ld_bias, const float* bias_rowmax, void* xwork,
                      int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                      uint32_t stream_id, const uint8_t* keep_in, void* out_mean, void* out_cov, float* lse, uint8_t* keep_bits, int32_t keep_ready,
                      void* stream);
/* Backward of b200vit_wattn_fwd (prep: Delta_i and the transposed keep bits; key-tile kernel -> dV, dCV and dD^T; per-(batch, head) kernel ->
 * dQ, dCQ, dK, dCK from dD^T and the transformed operands). dqkv_cov is the gradient w.r.t. the PRE-activation of elu(.)+1, i.e. ready for the
 * QKV wgrad / dgrad GEMMs. work: caller-owned, 256-byte aligned, b200vit_wattn_bwd_workspace_bytes(B, H, N, dtable != NULL) bytes (dD^T, Delta,
 * transposed keep bits and, for the bias-table gradient, dA^T). Bias gradients (optional, +=, [H*64]): dq_bias / dv_bias (q_bias, v_bias) and
 * dcq_bias / dcv_bias (cov_q_bias, cov_v_bias). */
size_t b200vit_wattn_bwd_workspace_bytes(int32_t B, int32_t H, int32_t N, int32_t with_dtable);
int b200vit_wattn_bwd(const void* qkv_mean, const void* qkv_cov, const void* xwork, const void* out_mean, const void* out_cov,
                      const void* dout_mean, const void* dout_cov, const float* lse, const float* bias_t, int64_t ld_bias,
                      const uint8_t* keep_bits, void* work, const int32_t* rel_index, float* dtable, float* dq_bias, float* dv_bias,
                      float* dcq_bias, float* dcv_bias, int32_t B, int32_t H, int32_t N, int32_t head_dim, float scale, float p_drop,
                      void* dqkv_mean, void* dqkv_cov, void* stream);
/* The Philox keep mask of b200vit_attn_fwd as uint8 [BH, N, N] (parity tests inject it into the CPU oracle). */
int b200vit_dropout_mask(uint8_t* out, int32_t BH, int32_t N, float p_drop, uint64_t seed, uint32_t stream_id, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound). C must be a multiple of 128 (<= 1024) for the LayerNorm kernels.
 * ---------------------------------------------------------------------------------------------- */
/* nn.LayerNorm (norm1/norm2/norm/fc_norm, modeling_finetune.py:270,280,420-421). Source row r is x[row_index[r]] when
 * row_index != NULL (final-norm + masked-row gather of modeling_cyclical.py:219-224). Outputs are compact [rows, C]. */
int b200vit_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_index, const float* gamma, const float* beta, float eps,
                          int32_t rows, int32_t C, void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream);
/* dx[row_index[r]] += LN-backward(dy[r]) ; dgamma += ; dbeta +=
 * The dx update is a plain read-modify-write: row_index entries of rows with a non-zero dy must be distinct. Rows whose dy is entirely
 * zero are skipped, so the padding of a fixed-capacity row list (see b200vit_d2v_target_loss, n_valid_dev) may repeat any valid row. */
int b200vit_layernorm_bwd(const void* dy, int32_t dy_is_f32, const float* x, int64_t ldx, const int32_t* row_index, const float* gamma,
                          const float* mean, const float* rstd, int32_t rows, int32_t C, float* dx, int64_t lddx, float* dgamma,
                          float* dbeta, void* stream);
/* b200vit_layernorm_bwd (all rows, no row list) followed, in the same pass over the rows, by b200vit_scale_residual_bwd of the residual
 * branch BELOW this LayerNorm in the backward order (the attention branch after norm2's backward; the previous block's MLP branch after
 * norm1's): dx += LN-backward(dy); dt = rowscale * gamma2 * dx; dgamma2 += sum rowscale * t * dx; dbias2 += sum dt. The 77 MB fp32 gradient
 * stream is read once instead of twice per LayerNorm. With bf16 dy, contiguous rows (ldx == lddx == C), rows % 4 == 0 and 16-byte aligned
 * pointers both LayerNorm-backward entries stage their operand rows through shared memory with cp.async.bulk (5.6-5.7 TB/s at M = 25 216,
 * C = 768); other shapes run register-staged kernels with the same results. */
int b200vit_layernorm_bwd_scale_residual(const void* dy, int32_t dy_is_f32, const float* x, int64_t ldx, const float* gamma, const float* mean,
                                         const float* rstd, int32_t rows, int32_t C, float* dx, int64_t lddx, float* dgamma, float* dbeta,
                                         const void* t_bf16, const float* rowscale, int32_t rows_per_scale, const float* gamma2, void* dt_bf16,
                                         float* dgamma2, float* dbias2, void* stream);
/* backward of x + drop_path(gamma * t) (Block.forward, modeling_finetune.py:296-298):
 * dt = rowscale[r / rows_per_scale] * gamma * dx (bf16); dgamma += sum rowscale * t * dx; dbias += sum dt */
int b200vit_scale_residual_bwd(const float* dx, int64_t lddx, const void* t_bf16, const float* rowscale, int32_t rows_per_scale,
                               const float* gamma, int32_t rows, int32_t C, void* dt_bf16, float* dgamma, float* dbias, void* stream);
int b200vit_colsum_bf16(const void* x, int64_t ldx, int32_t rows, int32_t C, float* out_accum, void* stream);
int b200vit_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* PatchEmbed (modeling_finetune.py:319-325) as im2col + GEMM: [B,Cin,H,W] fp32 -> [B*(H/P)*(W/P), Cin*P*P] bf16 */
int b200vit_im2col_patches(const float* img, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t P, void* out_bf16, void* stream);
/* cls concat + mask-token blend (+ pos_embed) (modeling_cyclical.py:175-194); mask uint8 [B*np] or NULL */
int b200vit_assemble_tokens(const float* pe, const float* cls, const float* mask_token, const uint8_t* mask, const float* pos_embed,
                            int32_t B, int32_t np, int32_t C, float* x, void* stream);
int b200vit_assemble_tokens_bwd(const float* dx, const uint8_t* mask, int32_t B, int32_t np, int32_t C, void* dpe_bf16, float* dcls,
                                float* dmask_token, float* dpos_embed, void* stream);
/* DropPath (timm drop_path, modeling_finetune.py:51-62): out[l, d, b] = keep / (1 - p_l), keep ~ Bernoulli(1 - p_l) from
 * Philox4x32-10 keyed on (seed; l, d, b). probs_host: HOST array of L drop probabilities (linspace(0, rate, L), :401). */
int b200vit_drop_path_scales(const float* probs_host, int32_t L, int32_t draws, int32_t B, uint64_t seed, float* out, void* stream);
/* Block-wise mask generator of the pre-training input pipeline (MaskingGenerator.__call__ / _mask, masking_generator.py:29-92; one call
 * per image in DataAugmentationForBEiT.__call__, datasets.py:104-118) for a whole batch on the device, plus the masked-row list the
 * student's head gathers (modeling_cyclical.py:221-225: row-major over (image, patch)).
 *   mask  [B, height*width] uint8 {0,1};   count [B+1] int32: masked patches per image, count[B] = their sum R (-1: injected stream exhausted)
 *   rows  [>= B*num_masking_patches] int32 or NULL: rows[k] = b*tokens + 1 + patch of the k-th masked patch (tokens = patches + cls)
 *   log_aspect_lo/hi = log(min_aspect), log(max_aspect) as the reference computes them (:42-43); height, width <= 32.
 * uniforms == NULL: draws come from Philox4x32-10 keyed on (seed; first_image + b, draw index) as 53-bit doubles.
 * uniforms != NULL: [B, uniforms_per_image] doubles in [0,1) consumed in call order (uniform(a,b) = a + (b-a)*u, randint(0,n) =
 * min(n, floor(u*(n+1)))): with the same stream fed to the reference generator the masks are bit-identical. */
int b200vit_block_masks(uint8_t* mask, int32_t* count, int32_t* rows, int32_t B, int32_t height, int32_t width, int32_t tokens,
                        int32_t num_masking_patches, int32_t min_num_patches, int32_t max_num_patches, double log_aspect_lo,
                        double log_aspect_hi, uint64_t seed, uint64_t first_image, const double* uniforms, int32_t uniforms_per_image,
                        void* stream);
/* RelativePositionBias.forward (modeling_finetune.py:359-364) in the padded layouts the attention kernels read:
 *   out_fwd  [H, N, ld]: scale * table[index[i,j], h] for j < N, -inf for N <= j < ld  (key mask baked into the padding)
 *   out_bwd_t[H, N, ld]: the transpose (row = key j, column = query i), 0 in the padding.   Either may be NULL.
 *   rowmax_fwd[H, N]  : max_j scale * table[index[i,j], h] (optional): the stabiliser of the single-pass Wasserstein attention forward.
 * The attention kernels expect scale = log2(e) and ld = N rounded up to 16. */
int b200vit_rel_pos_bias(const float* table, const int32_t* index, int32_t N, int32_t H, int32_t ld, float scale, float* out_fwd,
                         float* out_bwd_t, float* rowmax_fwd, void* stream);
/* x[:, 1:].mean(1) (modeling_finetune.py:512-514) */
int b200vit_meanpool_tokens(const float* x, int32_t B, int32_t T, int32_t C, float* out, void* stream);
/* backward of the mean pooling: dx[b, t>=1, :] += dpool[b, :] / (T-1) */
int b200vit_meanpool_tokens_bwd(const float* dpool, int32_t B, int32_t T, int32_t C, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * data2vec step kernels
 * ---------------------------------------------------------------------------------------------- */
/* Target builder + loss (engine_for_cyclical.py:90-150), masked rows only: per row r (source row row_index[r] of each
 * [*, ld_layer] fp32 teacher layer): t = LN?( mean_l LN?(layer_l[row]) ), eps 1e-5 no affine; loss = smooth_l1(y, t, beta)
 * (or MSE) mean over R*C; dy = dloss/dy * grad_scale. layers_host is a HOST array of device pointers.
 * Any of targets / dy_bf16 / dy_f32 / y may be NULL. row_loss: workspace of R floats; loss_out: device scalar.
 * n_valid_dev (device int32, or NULL): the row list is PADDED to a fixed capacity R and only its first *n_valid_dev rows are masked
 * patches (block-wise masking yields a slightly different count every batch, masking_generator.py:80-92; a fixed R keeps every launch
 * shape of the step constant). Padded rows get zero loss, zero dy and zero targets; the mean and grad_scale are rescaled by
 * R / *n_valid_dev, i.e. grad_scale is given as if all R rows were valid. Padded row_index entries must still be valid rows. */
int b200vit_d2v_target_loss(const float* const* layers_host, int32_t num_layers, int64_t ld_layer, const int32_t* row_index,
                            const float* y, int32_t R, int32_t C, int32_t ln_each, int32_t ln_post, float beta, int32_t l2_loss,
                            float grad_scale, float* targets, void* dy_bf16, float* dy_f32, float* row_loss, float* loss_out,
                            const int32_t* n_valid_dev, void* stream);
/* The same kernel through a descriptor, with the optional pieces of engine_for_cyclical.train_one_epoch:
 *   affine / rows_per_sample : instance / batch norm of the teacher layers (:94-104) or of the averaged target (:112-115) as one
 *                              {shift, scale} pair per (image, channel) from b200vit_channel_stats: v = (v - shift) * scale before the
 *                              per-row LayerNorm; image = source row / rows_per_sample. affine: HOST array of num_layers device pointers.
 *   compact_tokens = T       : `layers` are compact [B, T-1, C] patch-row tensors (no cls row) while row_index still holds residual-stream
 *                              rows b*T + 1 + p (second pass of post_target_instance_norm over the averaged target of ALL patch rows)
 *   row_index == NULL        : rows 0..R-1 (first pass of post_target_instance_norm: targets for every patch row)
 *   col_hinge                : device float2 [C] {mean_c, k_c} from b200vit_column_std: dy[r,c] += k_c * (y[r,c] - mean_c) (var_w0 hinge, :136-137)
 *   loss_out = loss_mult * (mean loss + loss_add_weight * *loss_add)   (loss_add: device scalar or NULL; `loss * loss_scale`, :160-163) */
typedef struct b200vit_d2v_desc {
  const float* const* layers;
  int32_t num_layers;
  int64_t ld_layer;
  const int32_t* row_index;
  const float* y;
  int32_t R, C;
  int32_t ln_each, ln_post;
  float beta;
  int32_t l2_loss;
  float grad_scale;
  float* targets;
  void* dy_bf16;
  float* dy_f32;
  float* row_loss;
  float* loss_out;
  const int32_t* n_valid_dev;
  const float* const* affine;
  int32_t rows_per_sample;
  int32_t compact_tokens;
  const float* col_hinge;
  const float* loss_add;
  float loss_add_weight;
  float loss_mult;
} b200vit_d2v_desc;
int b200vit_d2v_target_loss_ex(const b200vit_d2v_desc* desc, void* stream);
/* Statistics of target_batch_norm / target_instance_norm / post_target_instance_norm (engine_for_cyclical.py:94-104,112-115: F.batch_norm(training)
 * and F.instance_norm over the patch tokens of each channel, biased variance, eps 1e-5). For every layer l, image b and channel c the mean and
 * variance over rows [row0, row0 + nrows) of the image's `sample_rows` rows (row0 = 1 skips the cls row of a residual stream) are folded into
 * affine_out float2 [num_layers, samples, C] = {shift, scale}: batch norm alone {mu_c, rsqrt(var_c + eps)}; instance norm
 * {mean_bc, r_c * rsqrt(r_c^2 var_bc + eps)} with r_c the batch-norm scale in front of it (1 without). */
int b200vit_channel_stats(const float* const* layers_host, int32_t num_layers, int64_t ld_layer, int32_t samples, int32_t sample_rows,
                          int32_t row0, int32_t nrows, int32_t C, int32_t batch_norm, int32_t instance_norm, float eps, float* affine_out,
                          void* stream);
/* z0 = sqrt(outputs.var(dim=0) + eps) over the (valid) student rows and the var_w0 hinge std_loss0 = sum relu(margin - z0) / C
 * (engine_for_cyclical.py:130-139). y fp32 [R, C]; z0 [C] (optional); hinge_out device scalar (optional); col_hinge float2 [C] (optional) =
 * {mean_c, k_c}, k_c = -k_scale / (C (n-1) z0_c) where the hinge is active: the dy term of k_scale * std_loss0 (k_scale = var_w0 * loss_scale).
 * work: b200vit_column_std_workspace_bytes(C) bytes. */
size_t b200vit_column_std_workspace_bytes(int32_t C);
int b200vit_column_std(const float* y, int32_t R, int32_t C, const int32_t* n_valid_dev, float eps, float margin, float k_scale, float* work,
                       float* z0, float* hinge_out, float* col_hinge, void* stream);
/* out = wa * *a + wb * *b on device scalars (b may be NULL): (loss_cyc + loss_stochastic) * loss_scale, engine_for_cyclical.py:160-163 */
int b200vit_scalar_fma(float* out, const float* a, float wa, const float* b, float wb, void* stream);
/* mask_dropout_prob (engine_for_cyclical.py:62-66): mask = logical_and(bernoulli(1 - p_drop), mask) per patch, in place, then the per-image
 * counts (count [B+1], count[B] = total) and the masked-row list as b200vit_block_masks builds them. keep_in (uint8 [B*num_patches], optional)
 * injects the Bernoulli draws; otherwise Philox4x32-10 keyed on (seed; first_image + b, patch). */
int b200vit_mask_dropout(uint8_t* mask, int32_t* count, int32_t* rows, int32_t B, int32_t num_patches, int32_t tokens, float p_drop, uint64_t seed,
                         uint64_t first_image, const uint8_t* keep_in, void* stream);
/* EMA teacher update e = d*e + (1-d)*m (engine_for_cyclical.py:182-185, timm ModelEmaV2._update) + bf16 shadow */
int b200vit_ema_update(float* ema, const float* model, int64_t n, double decay, void* ema_bf16, void* stream);
/* The same update applied to the teacher's INTEGER relative_position_index buffer, as ModelEmaV2._update does by walking the whole state
 * dict: e = trunc(fp32(d) * fp32(e) + fp32(1 - d) * fp32(m)). For many decays (0.9; about half of the values on the 0.999 -> 0.9998 anneal)
 * some entries come back one lower, i.e. the reference's teacher reads a drifting bias index; reproduced for parity. */
int b200vit_ema_index_update(int32_t* ema_index, const int32_t* model_index, int32_t n, double decay, void* stream);
/* out_accum += sum g^2 (clip_grad_norm_, utils.py:374-377) */
int b200vit_sumsq(const float* g, int64_t n, float* out_accum, void* stream);
/* clip + torch.optim.AdamW + bf16 weight shadow + EMA over flat arenas (utils.py:364-390, optim_factory.py:58-97).
 * hp_lr_wd: device float2 {lr_scale, wd_scale} per 1024-element chunk (param groups); lr / weight_decay: this step's
 * schedule values (engine_for_cyclical.py:47-53). grad_div: divisor applied to g first (loss scale x world size). */
int b200vit_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* hp_lr_wd, float lr, float weight_decay,
                       float beta1, float beta2,
                       float eps, int32_t step, const float* gnorm_sq, float max_norm, float grad_div, void* p_bf16, float* ema,
                       double ema_decay, void* ema_bf16, void* stream);
/* WassersteinLoss.forward + backward (distloss.py:13-30,73-79). work: 2R+8 floats. d_* are accumulated (+=).
 * n_valid_dev as in d2v_target_loss: rows >= *n_valid_dev of a padded list take no part in the normalisers, the sum or the gradient. */
int b200vit_wasserstein_loss(const float* mean_out, const float* cov_out, const float* pos_mean, const float* pos_cov, int32_t R,
                             int32_t C, float lam, float grad_scale, float* work, float* d_mean_out, float* d_cov_out,
                             float* loss_out, const int32_t* n_valid_dev, void* stream);

/* Reparameterised Gaussian draw z = mean + sqrt(max(cov, 0)) * eps in front of the classifier head of the dual-stream model
 * (modeling_finetune_dist.py:314-325: Normal(mean, sqrt(cov)).sample(), commented out in the reference — optional here, default off).
 * eps_in != NULL injects the noise; otherwise eps ~ N(0,1) from Philox4x32-10 keyed on (seed; element / 4, stream_id) + Box-Muller.
 * eps_out (optional) receives the noise used (the backward needs it). n % 4 == 0. */
int b200vit_gaussian_sample(const float* mean, const float* cov, const float* eps_in, int64_t n, uint64_t seed, uint32_t stream_id,
                            float* eps_out, float* out_f32, void* out_bf16, void* stream);
/* dmean += dz ; dcov += dz * eps / (2 sqrt(cov)) where cov > 0 (either may be NULL) */
int b200vit_gaussian_sample_bwd(const float* dz, const float* cov, const float* eps, int64_t n, float* dmean, float* dcov, void* stream);
/* Fine-tune criterion (train_class_batch, engine_for_finetuning_dist.py:286-304): soft-target / label-smoothing cross-entropy of the logits
 * (mean over the batch) + optionally WassersteinLossFineTuning (distloss.py:39-70) of the anchor features against the positive / negative
 * features, with gradients w.r.t. logits and anchor features (through all max normalisers, like autograd). logits fp32 [B, ld_logits];
 * targets fp32 [B, K]; features fp32 [B, C] (mean_feat == NULL: cross-entropy only). dlogits fp32 [B, ld_dlogits] and / or dlogits_bf16
 * [B, ld_dlogits_bf16] with columns [K, K_padded) zeroed (the operand of the head's backward GEMMs). loss_out[3] = {total, ce, wloss}.
 * work: b200vit_finetune_loss_workspace_floats(B) floats. */
size_t b200vit_finetune_loss_workspace_floats(int32_t B);
int b200vit_finetune_loss(const float* logits, int64_t ld_logits, const float* targets, int32_t B, int32_t K, const float* mean_feat,
                          const float* cov_feat, const float* pos_mean, const float* pos_cov, const float* neg_mean, const float* neg_cov,
                          int32_t C, float lambda_finetuning, float lambda_pvn, float grad_scale, float* work, float* dlogits,
                          int64_t ld_dlogits, void* dlogits_bf16, int64_t ld_dlogits_bf16, int32_t K_padded, float* d_mean_feat,
                          float* d_cov_feat, float* loss_out, void* stream);

/* patch_transform of the data pipeline on the device: transforms.ToTensor() + transforms.Normalize(mean, std) (datasets.py:80-85) from uint8
 * pixels, so only 1 byte per element crosses PCIe. src uint8 [B,H,W,C] (hwc != 0: PIL / numpy layout) or [B,C,H,W]; C <= 4; mean_host / std_host:
 * HOST arrays of C floats. out fp32 [B,C,H,W] = ((src / 255) - mean[c]) / std[c], three rounded fp32 operations (bit-identical to torchvision). */
int b200vit_normalize_u8(const uint8_t* src, int32_t hwc, int32_t B, int32_t C, int32_t H, int32_t W, const float* mean_host,
                         const float* std_host, float* out, void* stream);
/* Mixup / CutMix of a fine-tune batch in place, "batch" mode (timm 0.3.2 Mixup._mix_batch + mixup_target, built at
 * run_class_finetuning.py:339-347 and applied at engine_for_finetuning.py:87-88): image b is paired with image B-1-b.
 *   use_cutmix == 0: x[b] = lam * x[b] + one_minus_lam * x[B-1-b] (two rounded products + one rounded sum, as the torch ops)
 *   use_cutmix != 0: x[b][:, yl:yh, xl:xh] = x[B-1-b][:, yl:yh, xl:xh]                (lam is then the box-corrected lambda)
 *   soft_targets[b, k] = lam * v(y[b], k) + one_minus_lam * v(y[B-1-b], k), v = on_value if k is the label else off_value
 * x fp32 [B, C, H, W] (NULL: targets only); labels int64 [B] and soft_targets fp32 [B, K] (both NULL: images only); lam == 1 leaves x alone. */
int b200vit_mixup_batch(float* x, int32_t B, int32_t C, int32_t H, int32_t W, float lam, float one_minus_lam, int32_t use_cutmix,
                        int32_t yl, int32_t yh, int32_t xl, int32_t xh, const int64_t* labels, int32_t K, float on_value, float off_value,
                        float* soft_targets, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MC-sample uncertainty reduction (uncertainty_evaluations.py:77-85,110-202,270-272)
 *   logits fp32 [S, N, K]; labels int32 [N].
 *   mean_logits [N,K]; row_stats [N, 8] = {conf, pred, correct@1, correct@5, nll, entropy(pbar), variance, mutual_info};
 *   hist [n_bins, 3] = {count, sum conf, sum correct} (+=, zero it first);
 * b200vit_mc_finalize: summary[8] = {acc1 %, acc5 %, ECE, ECE(reference indexing quirk), NLL, mean entropy, mean variance, mean MI}
 * ---------------------------------------------------------------------------------------------- */
int b200vit_mc_reduce(const float* logits, const int32_t* labels, int32_t S, int32_t N, int32_t K, int32_t n_bins, float* mean_logits,
                      float* row_stats, float* hist, void* stream);
int b200vit_mc_finalize(const float* row_stats, const float* hist, int32_t N, int32_t n_bins, float* summary, void* stream);

/* Class-wise calibration of the evaluation printout: TACELoss.loss (uncertainty_evaluations.py:241-261; threshold 0.01, 30 adaptive bins
 * per class from the sorted class probabilities, :119-132,159-186) and torchmetrics' multiclass AUROC (:49,85; macro one-vs-rest, exact).
 * logits fp32 [N, K] (is_prob != 0: already probabilities), labels int32 [N]; N >= n_bins, n_bins <= 64.
 * out[3] = {TACE, TACE as the reference computes it (uint8 fancy-index quirk of compute_bins, :173-184), AUROC}.
 * work: 256-byte aligned, b200vit_tace_auroc_workspace_bytes(N, K) bytes. */
size_t b200vit_tace_auroc_workspace_bytes(int32_t N, int32_t K);
int b200vit_tace_auroc(const float* logits, int32_t is_prob, const int32_t* labels, int32_t N, int32_t K, float threshold, int32_t n_bins,
                       void* work, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VIT_H_ */
