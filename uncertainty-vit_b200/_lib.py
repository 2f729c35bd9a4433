"""ctypes binding of libb200vit.so (the C ABI declared in include/b200vit.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200VIT_LIB") or os.path.join(HERE, "libb200vit.so")     # B200VIT_LIB: an alternative build (kernel A/B runs)

EPI_BF16, EPI_GELU, EPI_RESIDUAL, EPI_DGELU, EPI_F32, EPI_F32_ATOMIC, EPI_ELU1 = range(7)

vp, i32, i64, f32, u64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64, C.c_uint32


class GemmDesc(C.Structure):
    _fields_ = [("M", i32), ("N", i32), ("K", i32),
                ("A", vp), ("lda", i64), ("a_mn_major", i32),
                ("B", vp), ("ldb", i64), ("b_mn_major", i32),
                ("epilogue", i32),
                ("bias", vp), ("colscale", vp), ("rowscale", vp), ("rows_per_scale", i32),
                ("residual", vp), ("ld_residual", i64),
                ("aux", vp), ("ld_aux", i64),
                ("out_f32", vp), ("ld_f32", i64),
                ("out_bf16", vp), ("ld_bf16", i64),
                ("out2_bf16", vp), ("ld2_bf16", i64),
                ("alpha", f32), ("split_k", i32), ("max_ctas", i32), ("debug_flags", i32), ("colsum", vp)]


class D2VDesc(C.Structure):
    _fields_ = [("layers", C.POINTER(vp)), ("num_layers", i32), ("ld_layer", i64), ("row_index", vp), ("y", vp), ("R", i32), ("C", i32),
                ("ln_each", i32), ("ln_post", i32), ("beta", f32), ("l2_loss", i32), ("grad_scale", f32), ("targets", vp), ("dy_bf16", vp),
                ("dy_f32", vp), ("row_loss", vp), ("loss_out", vp), ("n_valid_dev", vp), ("affine", C.POINTER(vp)), ("rows_per_sample", i32),
                ("compact_tokens", i32), ("col_hinge", vp), ("loss_add", vp), ("loss_add_weight", f32), ("loss_mult", f32)]


_PROTOS = {
    "b200vit_last_error": (C.c_char_p, []),
    "b200vit_abi_version": (i32, []),
    "b200vit_device_sm_count": (i32, []),
    "b200vit_set_sm_limit": (i32, [i32]),
    "b200vit_gemm_bf16": (i32, [C.POINTER(GemmDesc), vp]),
    "b200vit_attn_fwd": (i32, [vp, vp, i64, i32, i32, i32, i32, f32, f32, u64, vp, u32, vp, vp, vp, vp, i32, vp, vp, i32, vp]),
    "b200vit_rel_pos_index_tiles": (i32, [vp, vp, i32, i32, i32, f32, vp, vp, vp]),
    "b200vit_keep_bits": (i32, [vp, i32, i32, f32, u64, vp, u32, vp, vp]),
    "b200vit_attn_bwd": (i32, [vp, vp, vp, vp, vp, i64, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, vp, vp]),
    "b200vit_attn_bwd_workspace_bytes": (C.c_size_t, [i32, i32, i32]),
    "b200vit_wattn_workspace_bytes": (C.c_size_t, [i32, i32, i32]),
    "b200vit_wattn_fwd": (i32, [vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, f32, f32, u64, vp, u32, vp, vp, vp, vp, vp, i32, vp]),
    "b200vit_wattn_bwd_workspace_bytes": (C.c_size_t, [i32, i32, i32, i32]),
    "b200vit_wattn_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, vp, vp, vp]),
    "b200vit_dropout_mask": (i32, [vp, i32, i32, f32, u64, u32, vp]),
    "b200vit_layernorm_fwd": (i32, [vp, i64, vp, vp, vp, f32, i32, i32, vp, vp, vp, vp, vp]),
    "b200vit_layernorm_bwd": (i32, [vp, i32, vp, i64, vp, vp, vp, vp, i32, i32, vp, i64, vp, vp, vp]),
    "b200vit_layernorm_bwd_scale_residual": (i32, [vp, i32, vp, i64, vp, vp, vp, i32, i32, vp, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    "b200vit_scale_residual_bwd": (i32, [vp, i64, vp, vp, i32, vp, i32, i32, vp, vp, vp, vp]),
    "b200vit_colsum_bf16": (i32, [vp, i64, i32, i32, vp, vp]),
    "b200vit_cast_f32_to_bf16": (i32, [vp, vp, i64, vp]),
    "b200vit_im2col_patches": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "b200vit_assemble_tokens": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    "b200vit_assemble_tokens_bwd": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "b200vit_drop_path_scales": (i32, [C.POINTER(f32), i32, i32, i32, u64, vp, vp]),
    "b200vit_mixup_batch": (i32, [vp, i32, i32, i32, i32, f32, f32, i32, i32, i32, i32, i32, vp, i32, f32, f32, vp, vp]),
    "b200vit_normalize_u8": (i32, [vp, i32, i32, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), vp, vp]),
    "b200vit_block_masks": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, C.c_double, C.c_double, u64, u64, vp, i32, vp]),
    "b200vit_rel_pos_bias": (i32, [vp, vp, i32, i32, i32, f32, vp, vp, vp, vp]),
    "b200vit_meanpool_tokens": (i32, [vp, i32, i32, i32, vp, vp]),
    "b200vit_meanpool_tokens_bwd": (i32, [vp, i32, i32, i32, vp, vp]),
    "b200vit_d2v_target_loss": (i32, [C.POINTER(vp), i32, i64, vp, vp, i32, i32, i32, i32, f32, i32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "b200vit_d2v_target_loss_ex": (i32, [C.POINTER(D2VDesc), vp]),
    "b200vit_channel_stats": (i32, [C.POINTER(vp), i32, i64, i32, i32, i32, i32, i32, i32, i32, f32, vp, vp]),
    "b200vit_column_std_workspace_bytes": (C.c_size_t, [i32]),
    "b200vit_column_std": (i32, [vp, i32, i32, vp, f32, f32, f32, vp, vp, vp, vp, vp]),
    "b200vit_scalar_fma": (i32, [vp, vp, f32, vp, f32, vp]),
    "b200vit_mask_dropout": (i32, [vp, vp, vp, i32, i32, i32, f32, u64, u64, vp, vp]),
    "b200vit_gaussian_sample": (i32, [vp, vp, vp, i64, u64, u32, vp, vp, vp, vp]),
    "b200vit_gaussian_sample_bwd": (i32, [vp, vp, vp, i64, vp, vp, vp]),
    "b200vit_finetune_loss_workspace_floats": (C.c_size_t, [i32]),
    "b200vit_finetune_loss": (i32, [vp, i64, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, vp, vp, i64, vp, i64, i32, vp, vp, vp, vp]),
    "b200vit_tace_auroc_workspace_bytes": (C.c_size_t, [i32, i32]),
    "b200vit_tace_auroc": (i32, [vp, i32, vp, i32, i32, f32, i32, vp, vp, vp]),
    "b200vit_ema_update": (i32, [vp, vp, i64, C.c_double, vp, vp]),
    "b200vit_ema_index_update": (i32, [vp, vp, i32, C.c_double, vp]),
    "b200vit_sumsq": (i32, [vp, i64, vp, vp]),
    "b200vit_adamw_step": (i32, [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, i32, vp, f32, f32, vp, vp, C.c_double, vp, vp]),
    "b200vit_wasserstein_loss": (i32, [vp, vp, vp, vp, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp]),
    "b200vit_mc_reduce": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "b200vit_mc_finalize": (i32, [vp, vp, i32, i32, vp, vp]),
}

_lib = None


class B200VitError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libb200vit.so (built in-tree by build.py / __graft_entry__.build()). Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200VitError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU or PyTorch fallback for the B200 hot path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.b200vit_abi_version() != 8:
            raise B200VitError("libb200vit ABI version mismatch")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b200vit_last_error().decode("utf-8", "replace")
        raise B200VitError(f"libb200vit {what} failed (rc={rc}): {msg}")


def exported_symbols():
    return sorted(_PROTOS)
