"""Tensor-level wrappers over the C ABI. PyTorch only supplies device memory and the CUDA stream; every call below ends in a
hand-written sm_100a kernel of libb200vit.so. Tensors must be CUDA tensors; there is no CPU path."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_DGELU, EPI_ELU1, EPI_F32, EPI_F32_ATOMIC, EPI_GELU, EPI_RESIDUAL, GemmDesc)

LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)
GEMM_TIMING = None  # bench.py sets this to a list to collect (start_event, end_event, algorithmic_flops) per GEMM launch


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


_LAST_DEV = None   # device of the most recent tensor argument: kernels are enqueued on THAT device's current stream


def _p(t: Optional[torch.Tensor]):
    global _LAST_DEV
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.B200VitError("libb200vit ops need CUDA tensors (no CPU fallback)")
    _LAST_DEV = t.device
    return t.data_ptr()


def _stream():
    """Current stream of the device the call's tensors live on (not of torch's current device: an engine on cuda:1 must not launch on
    cuda:0's stream). Every wrapper evaluates its tensor arguments (_p) before this."""
    global _RESTORE_DEV
    dev = _LAST_DEV
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        _RESTORE_DEV = torch.cuda.current_device()
        torch.cuda.set_device(dev)          # the launch itself needs the tensors' device current; check() switches back
    return torch.cuda.current_stream(dev).cuda_stream


_RESTORE_DEV = None


def check(rc: int, what: str = "") -> None:
    global _RESTORE_DEV
    if _RESTORE_DEV is not None:
        torch.cuda.set_device(_RESTORE_DEV)
        _RESTORE_DEV = None
    _lib.check(rc, what)


def sm_count() -> int:
    return _lib.lib().b200vit_device_sm_count()


def set_sm_limit(n: int) -> int:
    """Caps the SM count the persistent kernels size their grids for (0 = none); returns the previous cap."""
    return _lib.lib().b200vit_set_sm_limit(int(n))


# ----------------------------------------------------------------------------------------------------------------
# GEMM
# ----------------------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, *, a_mn: bool = False, b_mn: bool = False,
         epilogue: int = EPI_BF16, bias=None, colscale=None, rowscale=None, rows_per_scale: int = 0, residual=None,
         aux=None, out_f32=None, out_bf16=None, out2_bf16=None, alpha: float = 1.0, split_k: int = 0, max_ctas: int = 0,
         lda: Optional[int] = None, ldb: Optional[int] = None, colsum=None, debug_flags: int = 0) -> None:
    """D[M,N] = A[M,K] B[N,K]^T (bf16 in, fp32 accumulate). a_mn/b_mn: operand stored as [K, M] / [K, N] row-major."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = _p(a), (lda if lda is not None else a.stride(0)), int(a_mn)
    d.B, d.ldb, d.b_mn_major = _p(b), (ldb if ldb is not None else b.stride(0)), int(b_mn)
    d.epilogue = epilogue
    d.bias, d.colscale, d.rowscale, d.rows_per_scale = _p(bias), _p(colscale), _p(rowscale), rows_per_scale
    if residual is not None:
        d.residual, d.ld_residual = _p(residual), residual.stride(0)
    if aux is not None:
        d.aux, d.ld_aux = _p(aux), aux.stride(0)
    if out_f32 is not None:
        d.out_f32, d.ld_f32 = _p(out_f32), out_f32.stride(0)
    if out_bf16 is not None:
        d.out_bf16, d.ld_bf16 = _p(out_bf16), out_bf16.stride(0)
    if out2_bf16 is not None:
        d.out2_bf16, d.ld2_bf16 = _p(out2_bf16), out2_bf16.stride(0)
    d.alpha, d.split_k, d.max_ctas = alpha, split_k, max_ctas
    d.colsum = _p(colsum)
    d.debug_flags = debug_flags
    if GEMM_TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.lib().b200vit_gemm_bf16(C.byref(d), _stream()), "gemm_bf16")
        e1.record()
        GEMM_TIMING.append((e0, e1, 2.0 * M * N * K))
    else:
        check(_lib.lib().b200vit_gemm_bf16(C.byref(d), _stream()), "gemm_bf16")
    _count()


def linear_fwd(x_bf16, w_bf16, **kw):
    """x[M,K] @ w[N,K]^T"""
    M, K = x_bf16.shape
    N = w_bf16.shape[0]
    gemm(x_bf16, w_bf16, M, N, K, **kw)


def linear_dgrad(dy_bf16, w_bf16, **kw):
    """dx[M,K] = dy[M,N] @ w[N,K]   (w is the MN-major B operand)"""
    M, N = dy_bf16.shape
    K = w_bf16.shape[1]
    gemm(dy_bf16, w_bf16, M, K, N, b_mn=True, **kw)


def linear_wgrad(dy_bf16, x_bf16, dw_f32, alpha: float = 1.0):
    """dw[N,K] += dy[M,N]^T @ x[M,K]   (both operands MN-major, split-K with fp32 atomics)"""
    M, N = dy_bf16.shape
    K = x_bf16.shape[1]
    gemm(dy_bf16, x_bf16, N, K, M, a_mn=True, b_mn=True, epilogue=EPI_F32_ATOMIC, out_f32=dw_f32, alpha=alpha)


# ----------------------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------------------
def attn_fwd(qkv, bias, B, H, N, scale, p_drop=0.0, seed=0, stream_id=0, keep_in=None, out=None, lse=None, keep_bits=None, seed_dev=None,
             keep_ready=False):
    """bias: the padded dense bias of rel_pos_bias / pad_attn_bias, or None. When it carries the indexed form (.tab / .idx16, rel_pos_bias with
    want_index_tiles) the kernel gathers from the table instead of streaming the dense tensor — same values."""
    ld_bias = bias.stride(1) if bias is not None else 0
    idx16 = getattr(bias, "idx16", None) if bias is not None else None
    tab = getattr(bias, "tab", None) if idx16 is not None else None
    check(_lib.lib().b200vit_attn_fwd(_p(qkv), _p(bias), ld_bias, B, H, N, 64, scale, p_drop, seed, _p(seed_dev), stream_id, _p(keep_in), _p(out),
                                      _p(lse), _p(keep_bits), int(keep_ready), _p(idx16), _p(tab), int(bias.nbins) if idx16 is not None else 0,
                                      _stream()), "attn_fwd")
    _count(2 if (p_drop > 0 and not keep_ready) else 1)           # + the packed keep-mask kernel


def keep_bits(out, BH, N, p_drop, seed=0, stream_id=0, keep_in=None, seed_dev=None):
    """Packed dropout keep mask of one attention layer into `out` (uint8 [BH, N, 32]): Philox stream (seed / *seed_dev, stream_id) or injected mask."""
    check(_lib.lib().b200vit_keep_bits(_p(out), BH, N, p_drop, seed, _p(seed_dev), stream_id, _p(keep_in), _stream()), "keep_bits")
    _count()


def attn_bwd_workspace(B, H, N, device):
    """Workspace of b200vit_attn_bwd: dS^T bf16 [B,H,N,ld] | D fp32 [B,H,N] | transposed keep bits u32 [B,H,N,8]."""
    return torch.empty(int(_lib.lib().b200vit_attn_bwd_workspace_bytes(B, H, N)), dtype=torch.uint8, device=device)


def attn_bwd(qkv, out, dout, lse, bias_t, keep_bits, rel_index, dtable, B, H, N, scale, p_drop, dqkv, ds_work=None, dq_bias=None, dv_bias=None):
    bias = bias_t
    ld_bias = bias.stride(1) if bias is not None else 0
    ld_ds = attn_ld(N)
    if ds_work is None:
        ds_work = attn_bwd_workspace(B, H, N, qkv.device)
    check(_lib.lib().b200vit_attn_bwd(_p(qkv), _p(out), _p(dout), _p(lse), _p(bias), ld_bias, _p(keep_bits), _p(ds_work), ld_ds, _p(rel_index),
                                      _p(dtable), _p(dq_bias), _p(dv_bias), B, H, N, 64, scale, p_drop, _p(dqkv), _stream()), "attn_bwd")
    _count((4 if p_drop > 0 else 3) + (1 if dtable is not None else 0))


def _aligned_bytes(nbytes: int, device) -> torch.Tensor:
    """uint8 buffer whose data pointer is 256-byte aligned (torch's CUDA allocator hands out 512-byte aligned blocks)."""
    t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    assert t.data_ptr() % 256 == 0
    return t


def wattn_workspace(B, H, N, device) -> torch.Tensor:
    """Transformed-operand workspace of the Wasserstein attention (written by wattn_fwd, read by wattn_bwd of the same forward)."""
    return _aligned_bytes(_lib.lib().b200vit_wattn_workspace_bytes(B, H, N), device)


def wattn_bwd_workspace(B, H, N, with_dtable: bool, device) -> torch.Tensor:
    return _aligned_bytes(_lib.lib().b200vit_wattn_bwd_workspace_bytes(B, H, N, int(with_dtable)), device)


def wattn_fwd(qkv_mean, qkv_cov, bias, B, H, N, scale, p_drop=0.0, seed=0, stream_id=0, keep_in=None, out_mean=None, out_cov=None, lse=None,
              keep_bits=None, seed_dev=None, xwork=None, bias_rowmax=None, keep_ready=False):
    """bias: padded fwd layout of rel_pos_bias(); its row maxima ride along as `bias.rowmax` (or pass bias_rowmax)."""
    if bias_rowmax is None:
        bias_rowmax = getattr(bias, "rowmax", None)
    if bias_rowmax is None:
        raise _lib.B200VitError("wattn_fwd: the padded bias must come from ops.rel_pos_bias / ops.pad_attn_bias (row maxima attached)")
    if xwork is None:
        xwork = wattn_workspace(B, H, N, qkv_mean.device)
    check(_lib.lib().b200vit_wattn_fwd(_p(qkv_mean), _p(qkv_cov), _p(bias), bias.stride(1), _p(bias_rowmax), _p(xwork), B, H, N, 64, scale, p_drop,
                                       seed, _p(seed_dev), stream_id, _p(keep_in), _p(out_mean), _p(out_cov), _p(lse), _p(keep_bits), int(keep_ready),
                                       _stream()), "wattn_fwd")
    _count(3 if (p_drop > 0 and not keep_ready) else 2)
    return xwork


def wattn_bwd(qkv_mean, qkv_cov, xwork, out_mean, out_cov, dout_mean, dout_cov, lse, bias_t, keep_bits, rel_index, dtable, B, H, N, scale, p_drop,
              dqkv_mean, dqkv_cov, work=None, dq_bias=None, dv_bias=None, dcq_bias=None, dcv_bias=None):
    if work is None:
        work = wattn_bwd_workspace(B, H, N, dtable is not None, qkv_mean.device)
    check(_lib.lib().b200vit_wattn_bwd(_p(qkv_mean), _p(qkv_cov), _p(xwork), _p(out_mean), _p(out_cov), _p(dout_mean), _p(dout_cov), _p(lse),
                                       _p(bias_t), bias_t.stride(1), _p(keep_bits), _p(work), _p(rel_index), _p(dtable), _p(dq_bias), _p(dv_bias),
                                       _p(dcq_bias), _p(dcv_bias), B, H, N, 64, scale, p_drop, _p(dqkv_mean), _p(dqkv_cov), _stream()), "wattn_bwd")
    _count(3 + (1 if p_drop > 0 else 0) + (1 if dtable is not None else 0))


def dropout_mask(BH, N, p_drop, seed, stream_id, device) -> torch.Tensor:
    out = torch.empty(BH, N, N, dtype=torch.uint8, device=device)
    check(_lib.lib().b200vit_dropout_mask(_p(out), BH, N, p_drop, seed, stream_id, _stream()), "dropout_mask")
    _count()
    return out


# ----------------------------------------------------------------------------------------------------------------
# row kernels
# ----------------------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, rows, C_, y_bf16=None, y_f32=None, mean=None, rstd=None, row_index=None, ldx=None):
    check(_lib.lib().b200vit_layernorm_fwd(_p(x), ldx if ldx is not None else C_, _p(row_index), _p(gamma), _p(beta), eps, rows, C_,
                                           _p(y_bf16), _p(y_f32), _p(mean), _p(rstd), _stream()), "layernorm_fwd")
    _count()


def layernorm_bwd(dy, x, gamma, mean, rstd, rows, C_, dx, dgamma=None, dbeta=None, row_index=None, ldx=None, lddx=None):
    check(_lib.lib().b200vit_layernorm_bwd(_p(dy), int(dy.dtype == torch.float32), _p(x), ldx if ldx is not None else C_, _p(row_index),
                                           _p(gamma), _p(mean), _p(rstd), rows, C_, _p(dx), lddx if lddx is not None else C_,
                                           _p(dgamma), _p(dbeta), _stream()), "layernorm_bwd")
    _count()


def layernorm_bwd_scale_residual(dy, x, gamma, mean, rstd, rows, C_, dx, dgamma, dbeta, t_bf16, rowscale, rows_per_scale, gamma2, dt_bf16,
                                 dgamma2=None, dbias2=None):
    """layernorm_bwd on all rows fused with the scale_residual_bwd that reads the dx it produces: one pass over the gradient stream, operand rows
    staged through shared memory by bulk async copies (61 us against 48 + 41 us for the two kernels at M = 25 216, C = 768).
    B200VIT_FUSED_LN=0 launches the two kernels separately (same-box A/B measurements, tools/ab.sh)."""
    if os.environ.get("B200VIT_FUSED_LN", "1") == "0":
        layernorm_bwd(dy, x, gamma, mean, rstd, rows, C_, dx, dgamma, dbeta)
        scale_residual_bwd(dx, t_bf16, rowscale, rows_per_scale, gamma2, rows, C_, dt_bf16, dgamma2, dbias2)
        return
    check(_lib.lib().b200vit_layernorm_bwd_scale_residual(_p(dy), int(dy.dtype == torch.float32), _p(x), C_, _p(gamma), _p(mean), _p(rstd), rows, C_,
                                                          _p(dx), C_, _p(dgamma), _p(dbeta), _p(t_bf16), _p(rowscale), rows_per_scale, _p(gamma2),
                                                          _p(dt_bf16), _p(dgamma2), _p(dbias2), _stream()), "layernorm_bwd_scale_residual")
    _count()


def scale_residual_bwd(dx, t_bf16, rowscale, rows_per_scale, gamma, rows, C_, dt_bf16, dgamma=None, dbias=None):
    check(_lib.lib().b200vit_scale_residual_bwd(_p(dx), C_, _p(t_bf16), _p(rowscale), rows_per_scale, _p(gamma), rows, C_, _p(dt_bf16),
                                                _p(dgamma), _p(dbias), _stream()), "scale_residual_bwd")
    _count()


def colsum_bf16(x, rows, C_, out, ldx=None):
    check(_lib.lib().b200vit_colsum_bf16(_p(x), ldx if ldx is not None else x.stride(0), rows, C_, _p(out), _stream()), "colsum_bf16")
    _count()


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    check(_lib.lib().b200vit_cast_f32_to_bf16(_p(src), _p(dst), src.numel(), _stream()), "cast_f32_to_bf16")
    _count()
    return dst


def im2col(img, P, out):
    B, Cin, H, W = img.shape
    check(_lib.lib().b200vit_im2col_patches(_p(img), B, Cin, H, W, P, _p(out), _stream()), "im2col_patches")
    _count()


def assemble_tokens(pe, cls, mask_token, mask_u8, pos, B, np_, C_, x):
    check(_lib.lib().b200vit_assemble_tokens(_p(pe), _p(cls), _p(mask_token), _p(mask_u8), _p(pos), B, np_, C_, _p(x), _stream()),
          "assemble_tokens")
    _count()


def assemble_tokens_bwd(dx, mask_u8, B, np_, C_, dpe_bf16, dcls, dmask_token, dpos=None):
    check(_lib.lib().b200vit_assemble_tokens_bwd(_p(dx), _p(mask_u8), B, np_, C_, _p(dpe_bf16), _p(dcls), _p(dmask_token), _p(dpos),
                                                 _stream()), "assemble_tokens_bwd")
    _count()


def drop_path_scales(probs, draws, B, seed, device, out=None) -> torch.Tensor:
    """[L, draws, B] fp32 keep/(1-p) factors (device Philox; no torch RNG involved)."""
    L = len(probs)
    if out is None:
        out = torch.empty(L, draws, B, dtype=torch.float32, device=device)
    arr = (C.c_float * L)(*[float(p) for p in probs])
    check(_lib.lib().b200vit_drop_path_scales(arr, L, draws, B, seed, _p(out), _stream()), "drop_path_scales")
    _count()
    return out


def block_masks(B, height, width, tokens, num_masking_patches, min_num_patches, max_num_patches, log_aspect_lo, log_aspect_hi, seed,
                first_image, device, uniforms: Optional[torch.Tensor] = None, want_rows: bool = True, rows: Optional[torch.Tensor] = None):
    """Device block-wise masks for B images (masking_generator.py:29-92) -> (mask uint8 [B, height*width], count int32 [B+1] with
    count[B] = total masked rows R, rows int32 [B*num_masking_patches] of which the first R are valid, or None)."""
    mask = torch.empty(B, height * width, dtype=torch.uint8, device=device)
    count = torch.empty(B + 1, dtype=torch.int32, device=device)
    if rows is None:
        rows = torch.empty(max(B * num_masking_patches, 1), dtype=torch.int32, device=device) if want_rows else None
    elif rows.dtype != torch.int32 or rows.numel() < B * num_masking_patches or not rows.is_contiguous():
        raise _lib.B200VitError("block_masks: rows must be a contiguous int32 buffer of at least B * num_masking_patches entries")
    per_image = 0
    if uniforms is not None:
        if uniforms.dtype != torch.float64 or uniforms.dim() != 2 or uniforms.shape[0] != B or not uniforms.is_contiguous():
            raise _lib.B200VitError("block_masks: injected uniforms must be a contiguous float64 [B, n] tensor")
        per_image = uniforms.shape[1]
    check(_lib.lib().b200vit_block_masks(_p(mask), _p(count), _p(rows), B, height, width, tokens, num_masking_patches, min_num_patches,
                                         max_num_patches, float(log_aspect_lo), float(log_aspect_hi), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                         int(first_image) & 0xFFFFFFFFFFFFFFFF, _p(uniforms), per_image, _stream()), "block_masks")
    _count(2)
    return mask, count, rows


def normalize_u8(src: torch.Tensor, mean, std, hwc: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ToTensor + Normalize (datasets.py:80-85) of uint8 pixels on the device: src [B,H,W,C] (hwc) or [B,C,H,W] -> fp32 [B,C,H,W]."""
    if src.dtype != torch.uint8 or src.dim() != 4 or not src.is_contiguous():
        raise _lib.B200VitError("normalize_u8: src must be a contiguous uint8 4-D tensor")
    if hwc:
        B, H, W, Cc = src.shape
    else:
        B, Cc, H, W = src.shape
    if len(mean) != Cc or len(std) != Cc:
        raise _lib.B200VitError("normalize_u8: mean / std need one value per channel")
    if out is None:
        out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=src.device)
    m = (C.c_float * Cc)(*[float(v) for v in mean])
    s = (C.c_float * Cc)(*[float(v) for v in std])
    check(_lib.lib().b200vit_normalize_u8(_p(src), int(hwc), B, Cc, H, W, m, s, _p(out), _stream()), "normalize_u8")
    _count()
    return out


def mixup_batch(x: Optional[torch.Tensor], lam: float, use_cutmix: bool = False, box=(0, 0, 0, 0), labels: Optional[torch.Tensor] = None,
                num_classes: int = 0, on_value: float = 1.0, off_value: float = 0.0) -> Optional[torch.Tensor]:
    """In-place Mixup / CutMix of x fp32 [B,C,H,W] against x.flip(0) (timm Mixup, batch mode) and the mixed soft targets [B,K] (or None)."""
    import numpy as np
    soft = None
    if labels is not None:
        if labels.dtype != torch.int64 or not labels.is_contiguous():
            raise _lib.B200VitError("mixup_batch: labels must be contiguous int64")
        soft = torch.empty(labels.shape[0], num_classes, dtype=torch.float32, device=labels.device)
    if x is not None and (x.dtype != torch.float32 or x.dim() != 4 or not x.is_contiguous()):
        raise _lib.B200VitError("mixup_batch: x must be a contiguous fp32 [B,C,H,W] tensor")
    B, Cc, H, W = x.shape if x is not None else (labels.shape[0], 1, 1, 1)
    yl, yh, xl, xh = (int(v) for v in box)
    check(_lib.lib().b200vit_mixup_batch(_p(x), B, Cc, H, W, float(np.float32(lam)), float(np.float32(1.0 - lam)), int(use_cutmix), yl, yh, xl, xh,
                                         _p(labels), num_classes, float(np.float32(on_value)), float(np.float32(off_value)), _p(soft), _stream()),
          "mixup_batch")
    _count((1 if x is not None and lam != 1.0 else 0) + (1 if labels is not None else 0))
    return soft


LOG2E = 1.4426950408889634


def attn_ld(N: int) -> int:
    return (N + 15) // 16 * 16


ATTN_IDX_PITCH, ATTN_TAB_MAX = 210, 1024         # B200VIT_ATTN_IDX_PITCH / B200VIT_ATTN_TAB_MAX of include/b200vit.h
INDEXED_BIAS = os.environ.get("B200VIT_INDEXED_BIAS", "1") != "0"      # A/B switch: 0 = the forward streams the dense fp32 bias


def rel_pos_bias(table, index_i32, N, H, want_bwd: bool = True, want_rowmax: bool = False, want_index_tiles: bool = False):
    """(bias_fwd [H,N,ld], bias_bwd_t [H,N,ld] or None) in the padded, log2(e)-scaled layout of the attention kernels.
    want_index_tiles: also the INDEXED form of the same bias for b200vit_attn_fwd (attributes .tab / .idx16 / .nbins of bias_fwd)."""
    ld = attn_ld(N)
    fwd = torch.empty(H, N, ld, dtype=torch.float32, device=table.device)
    bwd = torch.empty(H, N, ld, dtype=torch.float32, device=table.device) if want_bwd else None
    rowmax = torch.empty(H, N, dtype=torch.float32, device=table.device) if want_rowmax else None
    check(_lib.lib().b200vit_rel_pos_bias(_p(table), _p(index_i32), N, H, ld, LOG2E, _p(fwd), _p(bwd), _p(rowmax), _stream()), "rel_pos_bias")
    _count(2 if want_rowmax else 1)
    if rowmax is not None:
        fwd.rowmax = rowmax            # stabiliser of the single-pass Wasserstein attention forward (rides along with the padded bias)
    nbins = int(table.shape[0])
    if want_index_tiles and INDEXED_BIAS and N <= 208 and nbins < ATTN_TAB_MAX:
        tab = torch.empty(H, (nbins + 4) // 4 * 4, dtype=torch.float32, device=table.device)
        idx16 = torch.empty((N + 127) // 128, 128, ATTN_IDX_PITCH, dtype=torch.int16, device=table.device)
        check(_lib.lib().b200vit_rel_pos_index_tiles(_p(table), _p(index_i32), N, H, nbins, LOG2E, _p(tab), _p(idx16), _stream()), "rel_pos_index_tiles")
        _count()
        fwd.tab, fwd.idx16, fwd.nbins = tab, idx16, nbins
    return fwd, bwd


def pad_attn_bias(bias: torch.Tensor):
    """Generic [H,N,N] additive bias -> the kernels' padded layouts (torch ops; tests / non-table biases only)."""
    H, N, _ = bias.shape
    ld = attn_ld(N)
    fwd = torch.full((H, N, ld), float("-inf"), dtype=torch.float32, device=bias.device)
    fwd[:, :, :N] = bias * LOG2E
    bwd = torch.zeros((H, N, ld), dtype=torch.float32, device=bias.device)
    bwd[:, :, :N] = bias.transpose(1, 2) * LOG2E
    fwd = fwd.contiguous()
    fwd.rowmax = (bias * LOG2E).amax(-1).contiguous()
    return fwd, bwd.contiguous()


def meanpool_tokens(x, B, T, C_, out):
    check(_lib.lib().b200vit_meanpool_tokens(_p(x), B, T, C_, _p(out), _stream()), "meanpool_tokens")
    _count()


def meanpool_tokens_bwd(dpool, B, T, C_, dx):
    check(_lib.lib().b200vit_meanpool_tokens_bwd(_p(dpool), B, T, C_, _p(dx), _stream()), "meanpool_tokens_bwd")
    _count()


# ----------------------------------------------------------------------------------------------------------------
# data2vec step
# ----------------------------------------------------------------------------------------------------------------
def d2v_target_loss(layers: Sequence[torch.Tensor], ld_layer, row_index, y, R, C_, ln_each=True, ln_post=True, beta=2.0, l2_loss=False,
                    grad_scale=1.0, targets=None, dy_bf16=None, dy_f32=None, row_loss=None, loss_out=None, n_valid=None):
    """n_valid: device int32 [1] when row_index is padded to the capacity R (rows >= n_valid: zero loss / dy / targets)."""
    arr = (C.c_void_p * len(layers))(*[_p(t) for t in layers])
    check(_lib.lib().b200vit_d2v_target_loss(arr, len(layers), ld_layer, _p(row_index), _p(y), R, C_, int(ln_each), int(ln_post), beta,
                                             int(l2_loss), grad_scale, _p(targets), _p(dy_bf16), _p(dy_f32), _p(row_loss), _p(loss_out),
                                             _p(n_valid), _stream()), "d2v_target_loss")
    _count(2 if loss_out is not None else 1)


def d2v_target_loss_ex(layers: Sequence[torch.Tensor], ld_layer, row_index, y, R, C_, ln_each=True, ln_post=True, beta=2.0, l2_loss=False,
                       grad_scale=1.0, targets=None, dy_bf16=None, dy_f32=None, row_loss=None, loss_out=None, n_valid=None, affine=None,
                       rows_per_sample=0, compact_tokens=0, col_hinge=None, loss_add=None, loss_add_weight=0.0, loss_mult=1.0):
    """b200vit_d2v_target_loss_ex: the target builder + loss with the optional instance / batch-norm maps (`affine`: one float2 [samples, C]
    tensor per layer from channel_stats), compact layers, the var_w0 hinge and the loss_scale multiplier (see include/b200vit.h)."""
    d = _lib.D2VDesc()
    arr = (C.c_void_p * len(layers))(*[_p(t) for t in layers])
    d.layers, d.num_layers, d.ld_layer = arr, len(layers), ld_layer
    d.row_index, d.y, d.R, d.C = _p(row_index), _p(y), R, C_
    d.ln_each, d.ln_post, d.beta, d.l2_loss, d.grad_scale = int(ln_each), int(ln_post), beta, int(l2_loss), grad_scale
    d.targets, d.dy_bf16, d.dy_f32, d.row_loss, d.loss_out, d.n_valid_dev = _p(targets), _p(dy_bf16), _p(dy_f32), _p(row_loss), _p(loss_out), _p(n_valid)
    aff = None
    if affine is not None:
        aff = (C.c_void_p * len(layers))(*[_p(t) for t in affine])
        d.affine = aff
    d.rows_per_sample, d.compact_tokens = rows_per_sample, compact_tokens
    d.col_hinge, d.loss_add, d.loss_add_weight, d.loss_mult = _p(col_hinge), _p(loss_add), loss_add_weight, loss_mult
    check(_lib.lib().b200vit_d2v_target_loss_ex(C.byref(d), _stream()), "d2v_target_loss_ex")
    _count(2 if loss_out is not None else 1)


def channel_stats(layers: Sequence[torch.Tensor], ld_layer, samples, sample_rows, row0, nrows, C_, batch_norm, instance_norm, eps=1e-5, out=None):
    """{shift, scale} maps of target_batch_norm / target_instance_norm (engine_for_cyclical.py:94-104): float2 [len(layers), samples, C]."""
    if out is None:
        out = torch.empty(len(layers), samples, C_, 2, dtype=torch.float32, device=layers[0].device)
    arr = (C.c_void_p * len(layers))(*[_p(t) for t in layers])
    check(_lib.lib().b200vit_channel_stats(arr, len(layers), ld_layer, samples, sample_rows, row0, nrows, C_, int(batch_norm), int(instance_norm), eps,
                                           _p(out), _stream()), "channel_stats")
    _count(2)
    return out


def column_std(y, R, C_, n_valid=None, eps=1e-6, margin=0.5, k_scale=0.0, want_hinge_grad=False, work=None, z0=None, hinge=None, col_hinge=None):
    """z0 = sqrt(y.var(0) + eps), std_loss0 = sum relu(margin - z0) / C and (optionally) the {mean_c, k_c} gradient map of k_scale * std_loss0."""
    dev = y.device
    if work is None:
        work = torch.empty(int(_lib.lib().b200vit_column_std_workspace_bytes(C_)) // 4, dtype=torch.float32, device=dev)
    z0 = torch.empty(C_, dtype=torch.float32, device=dev) if z0 is None else z0
    hinge = torch.empty(1, dtype=torch.float32, device=dev) if hinge is None else hinge
    if want_hinge_grad and col_hinge is None:
        col_hinge = torch.empty(C_, 2, dtype=torch.float32, device=dev)
    check(_lib.lib().b200vit_column_std(_p(y), R, C_, _p(n_valid), eps, margin, k_scale, _p(work), _p(z0), _p(hinge), _p(col_hinge), _stream()),
          "column_std")
    _count(2)
    return z0, hinge, col_hinge


def scalar_fma(out, a, wa, b=None, wb=0.0):
    check(_lib.lib().b200vit_scalar_fma(_p(out), _p(a), wa, _p(b), wb, _stream()), "scalar_fma")
    _count()


def mask_dropout(mask_u8, B, num_patches, tokens, p_drop, seed=0, first_image=0, keep_in=None, rows=None, count=None):
    """mask &= bernoulli(1 - p) in place (engine_for_cyclical.py:62-66) -> (count int32 [B+1], rows int32 [capacity])."""
    dev = mask_u8.device
    if count is None:
        count = torch.empty(B + 1, dtype=torch.int32, device=dev)
    if rows is None:
        rows = torch.zeros(B * num_patches, dtype=torch.int32, device=dev)
    check(_lib.lib().b200vit_mask_dropout(_p(mask_u8), _p(count), _p(rows), B, num_patches, tokens, p_drop, int(seed) & 0xFFFFFFFFFFFFFFFF,
                                          int(first_image) & 0xFFFFFFFFFFFFFFFF, _p(keep_in), _stream()), "mask_dropout")
    _count(2)
    return count, rows


def gaussian_sample(mean, cov, eps_in=None, seed=0, stream_id=0, want_eps=False, out_f32=None, out_bf16=None):
    """z = mean + sqrt(max(cov, 0)) * eps (Philox / Box-Muller or injected eps) -> (z fp32 or None, z bf16 or None, eps or None)."""
    n = mean.numel()
    if out_f32 is None and out_bf16 is None:
        out_f32 = torch.empty_like(mean)
    eps_out = torch.empty_like(mean) if want_eps else None
    check(_lib.lib().b200vit_gaussian_sample(_p(mean), _p(cov), _p(eps_in), n, int(seed) & 0xFFFFFFFFFFFFFFFF, stream_id, _p(eps_out), _p(out_f32),
                                             _p(out_bf16), _stream()), "gaussian_sample")
    _count()
    return out_f32, out_bf16, eps_out


def gaussian_sample_bwd(dz, cov, eps, dmean=None, dcov=None):
    check(_lib.lib().b200vit_gaussian_sample_bwd(_p(dz), _p(cov), _p(eps), dz.numel(), _p(dmean), _p(dcov), _stream()), "gaussian_sample_bwd")
    _count()


def finetune_loss(logits, targets, K, feats=None, lam_ft=1e-4, lam_pvn=1e-4, grad_scale=1.0, dlogits=None, dlogits_bf16=None, work=None,
                  loss_out=None):
    """Soft-target cross-entropy (+ WassersteinLossFineTuning when feats = (mean, cov, pos_mean, pos_cov, neg_mean, neg_cov), fp32 [B, C]).
    logits fp32 [B, >= K] (row stride taken from the tensor); returns (loss_out[3] = {total, ce, wloss}, d_mean_feat, d_cov_feat)."""
    B = logits.shape[0]
    dev = logits.device
    if work is None:
        work = torch.empty(int(_lib.lib().b200vit_finetune_loss_workspace_floats(B)), dtype=torch.float32, device=dev)
    if loss_out is None:
        loss_out = torch.empty(3, dtype=torch.float32, device=dev)
    dfm = dfc = None
    f = [None] * 6
    Cf = 0
    if feats is not None:
        f = [t.contiguous() for t in feats]
        Cf = f[0].shape[1]
        dfm, dfc = torch.empty_like(f[0]), torch.empty_like(f[1])
    check(_lib.lib().b200vit_finetune_loss(_p(logits), logits.stride(0), _p(targets), B, K, *[_p(t) for t in f], Cf, lam_ft, lam_pvn, grad_scale,
                                           _p(work), _p(dlogits), dlogits.stride(0) if dlogits is not None else 0, _p(dlogits_bf16),
                                           dlogits_bf16.stride(0) if dlogits_bf16 is not None else 0,
                                           dlogits_bf16.shape[1] if dlogits_bf16 is not None else 0, _p(dfm), _p(dfc), _p(loss_out), _stream()),
          "finetune_loss")
    _count(3 if feats is not None else 2)
    return loss_out, dfm, dfc


def tace_auroc(logits, labels_i32, is_prob=False, threshold=0.01, n_bins=30):
    """(TACE, TACE as the reference computes it, macro one-vs-rest AUROC) of fp32 [N, K] logits / probabilities: device tensor [3]."""
    N, K = logits.shape
    dev = logits.device
    work = torch.empty(int(_lib.lib().b200vit_tace_auroc_workspace_bytes(N, K)) + 256, dtype=torch.uint8, device=dev)
    off = (-work.data_ptr()) % 256
    out = torch.empty(3, dtype=torch.float32, device=dev)
    check(_lib.lib().b200vit_tace_auroc(_p(logits), int(is_prob), _p(labels_i32), N, K, threshold, n_bins, work.data_ptr() + off, _p(out),
                                        _stream()), "tace_auroc")
    _count(3 if is_prob else 4)
    return out


def ema_update(ema, model, decay, ema_bf16=None):
    check(_lib.lib().b200vit_ema_update(_p(ema), _p(model), ema.numel(), decay, _p(ema_bf16), _stream()), "ema_update")
    _count()


def ema_index_update(ema_index_i32, model_index_i32, decay):
    check(_lib.lib().b200vit_ema_index_update(_p(ema_index_i32), _p(model_index_i32), ema_index_i32.numel(), decay, _stream()), "ema_index_update")
    _count()


def sumsq(g, out_accum):
    check(_lib.lib().b200vit_sumsq(_p(g), g.numel(), _p(out_accum), _stream()), "sumsq")
    _count()


def adamw_step(p, g, m, v, hp, step, lr, weight_decay, beta1=0.9, beta2=0.999, eps=1e-8, gnorm_sq=None, max_norm=0.0, grad_div=1.0, p_bf16=None,
               ema=None, ema_decay=0.0, ema_bf16=None):
    check(_lib.lib().b200vit_adamw_step(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(hp), lr, weight_decay, beta1, beta2, eps, step, _p(gnorm_sq), max_norm,
                                        grad_div, _p(p_bf16), _p(ema), ema_decay, _p(ema_bf16), _stream()), "adamw_step")
    _count()


def wasserstein_loss(mean_out, cov_out, pos_mean, pos_cov, lam, grad_scale, work, d_mean, d_cov, loss_out, n_valid=None):
    R, C_ = mean_out.shape
    check(_lib.lib().b200vit_wasserstein_loss(_p(mean_out), _p(cov_out), _p(pos_mean), _p(pos_cov), R, C_, lam, grad_scale, _p(work),
                                              _p(d_mean), _p(d_cov), _p(loss_out), _p(n_valid), _stream()), "wasserstein_loss")
    _count(3 if d_mean is not None else 2)


def mc_reduce(logits, labels_i32, n_bins=15, finalize=True):
    """logits fp32 [S, N, K] -> (mean_logits [N, K], row_stats [N, 8], hist [n_bins, 3], summary [8] or None when finalize=False: the
    multi-rank evaluation reduces its own images, gathers row_stats / sums hist across ranks and calls mc_finalize once)."""
    S, N, K = logits.shape
    dev = logits.device
    mean_logits = torch.empty(N, K, dtype=torch.float32, device=dev)
    row_stats = torch.empty(N, 8, dtype=torch.float32, device=dev)
    hist = torch.zeros(n_bins, 3, dtype=torch.float32, device=dev)
    check(_lib.lib().b200vit_mc_reduce(_p(logits), _p(labels_i32), S, N, K, n_bins, _p(mean_logits), _p(row_stats), _p(hist), _stream()),
          "mc_reduce")
    _count()
    summary = mc_finalize(row_stats, hist, N, n_bins) if finalize else None
    return mean_logits, row_stats, hist, summary


def mc_finalize(row_stats, hist, N, n_bins=15):
    summary = torch.empty(8, dtype=torch.float32, device=row_stats.device)
    check(_lib.lib().b200vit_mc_finalize(_p(row_stats), _p(hist), N, n_bins, _p(summary), _stream()), "mc_finalize")
    _count()
    return summary
