"""Forward / backward schedules of the ViT hot path over the C-ABI kernels (no autograd, no torch math).

`vit_forward` / `vit_backward` take a ParamSource (fp32 master tensors + bf16 shadows, addressed by the REFERENCE's state-dict
names) so the same schedule serves the nn.Module boundary (modeling.py, through torch.autograd.Function) and
the flat-arena training engine (engine.py).

Math restated from: Block.forward (modeling_finetune.py:290-299), Attention.forward (:145-188), Mlp.forward (:75-82),
VisionTransformerForCyclicalTraining.forward_features/forward (modeling_cyclical.py:170-225),
VisionTransformer.forward_features/forward (modeling_finetune.py:476-523).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import EPI_BF16, EPI_DGELU, EPI_F32, EPI_GELU, EPI_RESIDUAL, B200VitError


@dataclass
class VitConfig:
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: float = 4.0
    num_classes: int = 1000
    ln_eps: float = 1e-6
    kind: str = "cyclical"          # "cyclical" | "finetune"
    dist: bool = False
    drop_path_rate: float = 0.0
    attn_drop_rate: float = 0.0
    has_gamma: bool = True
    use_abs_pos_emb: bool = False
    sample_head: bool = False       # dual-stream classifier: logits = head(mean + sqrt(max(cov, 0)) * eps) instead of head(mean); the draw the
                                    # reference sketches and leaves commented out (modeling_finetune_dist.py:314-325). Default off.

    @property
    def grid(self):
        return self.img_size // self.patch_size

    @property
    def num_patches(self):
        return self.grid * self.grid

    @property
    def tokens(self):
        return self.num_patches + 1

    @property
    def hidden(self):
        return int(self.embed_dim * self.mlp_ratio)

    @property
    def drop_path_probs(self) -> List[float]:
        return [float(x) for x in torch.linspace(0, self.drop_path_rate, self.depth)]   # modeling_finetune.py:401


class ParamSource:
    """fp32 master parameters and their bf16 GEMM shadows by reference state-dict name."""

    def f32(self, name: str) -> Optional[torch.Tensor]:
        raise NotImplementedError

    def bf16(self, name: str) -> torch.Tensor:
        raise NotImplementedError

    def qkv_bias(self, prefix: str, cov: bool = False) -> torch.Tensor:
        """cat(q_bias, 0, v_bias) fp32 [3C] (modeling_finetune.py:148)."""
        raise NotImplementedError

    def rel_index_i32(self) -> torch.Tensor:
        raise NotImplementedError

    def head_padded(self):
        """(head.weight bf16 [Kp, C], head.bias fp32 [Kp]) with Kp = num_classes rounded up to 8, padding rows zero."""
        raise NotImplementedError


@dataclass
class Noise:
    """Randomness of one training forward. Defaults = device Philox streams; any field may be injected (parity tests)."""
    seed: int = 0
    seed_dev: Optional[torch.Tensor] = None             # int64 [1] on the device: overrides `seed` for attention dropout (CUDA-graph replays)
    drop_path_scale: Optional[torch.Tensor] = None      # fp32 [L, draws, B] = keep / (1 - p_l)
    attn_keep: Optional[List[torch.Tensor]] = None      # per layer uint8 [B, H, N, N]
    head_eps: Optional[torch.Tensor] = None             # fp32 [B, C]: injected N(0,1) noise of the reparameterised head sample (cfg.sample_head)
    keep_bits_all: Optional[torch.Tensor] = None        # uint8 [L, B, H, N, 32]: packed attention-dropout masks of ALL layers, drawn ahead of the
    keep_join: Optional[object] = None                  # forward on a side stream (draw_keep_bits): one event per layer, block i waits for event i
    drop_path_active: bool = True                       # applies only to training forwards
    attn_drop_active: Optional[bool] = None             # None: follow `train`; True: dropout even in eval (MC-dropout, enable_dropout())


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


# ------------------------------------------------------------------------------------------------------------------
# one block
# ------------------------------------------------------------------------------------------------------------------
def block_forward(ps: ParamSource, cfg: VitConfig, i: int, x_in: torch.Tensor, B: int, bias: Optional[torch.Tensor], *, save: bool,
                  dp_scale: Optional[torch.Tensor], p_attn: float, seed: int, keep_in: Optional[torch.Tensor],
                  x_mid: Optional[torch.Tensor] = None, x_out: Optional[torch.Tensor] = None, seed_dev: Optional[torch.Tensor] = None,
                  keep_pre: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """x_in: fp32 [B*T, C] residual stream. Returns the saved tensors (x_out under 'x_out')."""
    T, C, H, Hd = cfg.tokens, cfg.embed_dim, cfg.num_heads, cfg.hidden
    M = B * T
    dev = x_in.device
    p = f"blocks.{i}."
    bf = torch.bfloat16
    h1 = _empty((M, C), bf, dev)
    mean1 = _empty((M,), torch.float32, dev)
    rstd1 = _empty((M,), torch.float32, dev)
    ops.layernorm_fwd(x_in, ps.f32(p + "norm1.weight"), ps.f32(p + "norm1.bias"), cfg.ln_eps, M, C, y_bf16=h1, mean=mean1, rstd=rstd1)
    qkv = _empty((M, 3 * C), bf, dev)
    ops.gemm(h1, ps.bf16(p + "attn.qkv.weight"), M, 3 * C, C, epilogue=EPI_BF16, bias=ps.qkv_bias(p), out_bf16=qkv)
    attn_out = _empty((M, C), bf, dev)
    lse = _empty((B, H, T), torch.float32, dev) if save else None
    keep_bits = (keep_pre if keep_pre is not None else torch.empty((B, H, T, 32), dtype=torch.uint8, device=dev)) if p_attn > 0 else None
    ops.attn_fwd(qkv, bias, B, H, T, (C // H) ** -0.5, p_attn, seed, i, keep_in, attn_out, lse, keep_bits, seed_dev=seed_dev,
                 keep_ready=keep_pre is not None and p_attn > 0)
    if x_mid is None:
        x_mid = _empty((M, C), torch.float32, dev)
    t1 = _empty((M, C), bf, dev) if save else None
    dp1 = dp_scale[0] if dp_scale is not None else None
    dp2 = dp_scale[1] if dp_scale is not None else None
    g1 = ps.f32(p + "gamma_1") if cfg.has_gamma else None
    g2 = ps.f32(p + "gamma_2") if cfg.has_gamma else None
    ops.gemm(attn_out, ps.bf16(p + "attn.proj.weight"), M, C, C, epilogue=EPI_RESIDUAL, bias=ps.f32(p + "attn.proj.bias"), colscale=g1,
             rowscale=dp1, rows_per_scale=T, residual=x_in, out_f32=x_mid, out2_bf16=t1)
    h2 = _empty((M, C), bf, dev)
    mean2 = _empty((M,), torch.float32, dev)
    rstd2 = _empty((M,), torch.float32, dev)
    ops.layernorm_fwd(x_mid, ps.f32(p + "norm2.weight"), ps.f32(p + "norm2.bias"), cfg.ln_eps, M, C, y_bf16=h2, mean=mean2, rstd=rstd2)
    act = _empty((M, Hd), bf, dev)
    pre = _empty((M, Hd), bf, dev) if save else None      # receives gelu'(fc1 output): the dGELU epilogue of the backward is one multiply
    ops.gemm(h2, ps.bf16(p + "mlp.fc1.weight"), M, Hd, C, epilogue=EPI_GELU, bias=ps.f32(p + "mlp.fc1.bias"), out_bf16=act, out2_bf16=pre)
    if x_out is None:
        x_out = _empty((M, C), torch.float32, dev)
    t2 = _empty((M, C), bf, dev) if save else None
    ops.gemm(act, ps.bf16(p + "mlp.fc2.weight"), M, C, Hd, epilogue=EPI_RESIDUAL, bias=ps.f32(p + "mlp.fc2.bias"), colscale=g2,
             rowscale=dp2, rows_per_scale=T, residual=x_mid, out_f32=x_out, out2_bf16=t2)
    if not save:
        return {"x_out": x_out, "x_mid": x_mid}
    return dict(x_in=x_in, h1=h1, mean1=mean1, rstd1=rstd1, qkv=qkv, attn_out=attn_out, lse=lse, keep_bits=keep_bits, t1=t1, x_mid=x_mid,
                h2=h2, mean2=mean2, rstd2=rstd2, act=act, pre=pre, t2=t2, x_out=x_out, dp1=dp1, dp2=dp2, p_attn=p_attn)


def block_backward(ps: ParamSource, cfg: VitConfig, i: int, s: Dict[str, torch.Tensor], dx: torch.Tensor, B: int,
                   bias: Optional[torch.Tensor], grads: Dict[str, torch.Tensor], dtable: Optional[torch.Tensor], ws: Dict[str, torch.Tensor],
                   first_srb_done: bool = False, below: Optional[Dict[str, torch.Tensor]] = None):
    """dx: fp32 [B*T, C] gradient of the block output; updated IN PLACE to the gradient of the block input.
    Each LayerNorm backward is fused with the scale-residual backward that consumes the dx it produces (one pass over the fp32 gradient
    stream instead of two): norm2's with this block's attention branch, norm1's with the MLP branch of the block BELOW (`below` = that
    block's saved tensors; its block_backward is then called with first_srb_done=True)."""
    T, C, H, Hd = cfg.tokens, cfg.embed_dim, cfg.num_heads, cfg.hidden
    M = B * T
    p = f"blocks.{i}."
    g = lambda n: grads.get(p + n)
    dt, dpre, dh, dqkv = ws["dt"], ws["dpre"], ws["dh"], ws["dqkv"]
    g2 = ps.f32(p + "gamma_2") if cfg.has_gamma else None
    g1 = ps.f32(p + "gamma_1") if cfg.has_gamma else None
    # ---- MLP branch: x_out = x_mid + dp2 * gamma_2 * (fc2(gelu(fc1(LN2 x_mid))))
    if not first_srb_done:
        ops.scale_residual_bwd(dx, s["t2"], s["dp2"], T, g2, M, C, dt, g("gamma_2") if cfg.has_gamma else None, g("mlp.fc2.bias"))
    ops.linear_wgrad(dt, s["act"], g("mlp.fc2.weight"))
    # fc1.bias gradient = column sums of dpre: fused into the dGELU epilogue (one launch and one 155 MB read less per block)
    ops.gemm(dt, ps.bf16(p + "mlp.fc2.weight"), M, Hd, C, b_mn=True, epilogue=EPI_DGELU, aux=s["pre"], out_bf16=dpre, colsum=g("mlp.fc1.bias"))
    ops.linear_wgrad(dpre, s["h2"], g("mlp.fc1.weight"))
    ops.gemm(dpre, ps.bf16(p + "mlp.fc1.weight"), M, C, Hd, b_mn=True, epilogue=EPI_BF16, out_bf16=dh)
    # ---- norm2 backward + attention branch: x_mid = x_in + dp1 * gamma_1 * proj(attn(LN1 x_in))
    ops.layernorm_bwd_scale_residual(dh, s["x_mid"], ps.f32(p + "norm2.weight"), s["mean2"], s["rstd2"], M, C, dx, g("norm2.weight"), g("norm2.bias"),
                                     s["t1"], s["dp1"], T, g1, dt, g("gamma_1") if cfg.has_gamma else None, g("attn.proj.bias"))
    ops.linear_wgrad(dt, s["attn_out"], g("attn.proj.weight"))
    ops.gemm(dt, ps.bf16(p + "attn.proj.weight"), M, C, C, b_mn=True, epilogue=EPI_BF16, out_bf16=dh)
    ops.attn_bwd(s["qkv"], s["attn_out"], dh, s["lse"], bias, s["keep_bits"], ps.rel_index_i32() if dtable is not None else None, dtable,
                 B, H, T, (C // H) ** -0.5, s["p_attn"], dqkv, ds_work=ws.get("attn_ws"), dq_bias=g("attn.q_bias"), dv_bias=g("attn.v_bias"))
    ops.linear_wgrad(dqkv, s["h1"], g("attn.qkv.weight"))
    ops.gemm(dqkv, ps.bf16(p + "attn.qkv.weight"), M, C, 3 * C, b_mn=True, epilogue=EPI_BF16, out_bf16=dh)
    if below is None:
        ops.layernorm_bwd(dh, s["x_in"], ps.f32(p + "norm1.weight"), s["mean1"], s["rstd1"], M, C, dx, g("norm1.weight"), g("norm1.bias"))
    else:
        q = f"blocks.{i - 1}."
        ops.layernorm_bwd_scale_residual(dh, s["x_in"], ps.f32(p + "norm1.weight"), s["mean1"], s["rstd1"], M, C, dx, g("norm1.weight"), g("norm1.bias"),
                                         below["t2"], below["dp2"], T, ps.f32(q + "gamma_2") if cfg.has_gamma else None, dt,
                                         grads.get(q + "gamma_2") if cfg.has_gamma else None, grads.get(q + "mlp.fc2.bias"))


def _backward_blocks(block_fn, cfg: VitConfig, saved, after_block, *args):
    """Runs block_fn(i, saved[i], first_srb_done, below) for i = depth-1 .. 0, chaining the LayerNorm / scale-residual fusion across blocks and
    releasing each block's activations as soon as the block below no longer needs them."""
    for i in reversed(range(cfg.depth)):
        below = saved[i - 1] if i > 0 else None
        block_fn(i, saved[i], i < cfg.depth - 1, below)
        saved[i] = None
        if after_block is not None:
            after_block(i)


def backward_workspace(cfg: VitConfig, B: int, dev) -> Dict[str, torch.Tensor]:
    M, C, Hd = B * cfg.tokens, cfg.embed_dim, cfg.hidden
    bf = torch.bfloat16
    T = cfg.tokens
    return dict(dt=_empty((M, C), bf, dev), dpre=_empty((M, Hd), bf, dev), dh=_empty((M, C), bf, dev), dqkv=_empty((M, 3 * C), bf, dev),
                ds=torch.zeros((B, cfg.num_heads, T, (T + 15) // 16 * 16), dtype=bf, device=dev),
                attn_ws=ops.attn_bwd_workspace(B, cfg.num_heads, T, dev))


# ------------------------------------------------------------------------------------------------------------------
# stem / bias
# ------------------------------------------------------------------------------------------------------------------
def patches_bf16(cfg: VitConfig, images: torch.Tensor) -> torch.Tensor:
    B = images.shape[0]
    K = cfg.in_chans * cfg.patch_size * cfg.patch_size
    out = _empty((B * cfg.num_patches, K), torch.bfloat16, images.device)
    ops.im2col(images, cfg.patch_size, out)
    return out


def stem_forward(ps: ParamSource, cfg: VitConfig, patches: torch.Tensor, B: int, mask_u8: Optional[torch.Tensor], x: Optional[torch.Tensor] = None,
                 prefix: str = "") -> torch.Tensor:
    C, npat = cfg.embed_dim, cfg.num_patches
    dev = patches.device
    pe = _empty((B * npat, C), torch.float32, dev)
    ops.gemm(patches, ps.bf16(prefix + "patch_embed.proj.weight"), B * npat, C, patches.shape[1], epilogue=EPI_F32,
             bias=ps.f32(prefix + "patch_embed.proj.bias"), out_f32=pe)
    if x is None:
        x = _empty((B * cfg.tokens, C), torch.float32, dev)
    ops.assemble_tokens(pe, ps.f32(prefix + "cls_token"), ps.f32(prefix + "mask_token") if mask_u8 is not None else None, mask_u8,
                        ps.f32("pos_embed") if cfg.use_abs_pos_emb else None, B, npat, C, x)
    return x


def stem_backward(ps: ParamSource, cfg: VitConfig, patches: torch.Tensor, dx: torch.Tensor, B: int, mask_u8, grads, prefix: str = ""):
    C, npat = cfg.embed_dim, cfg.num_patches
    dpe = _empty((B * npat, C), torch.bfloat16, dx.device)
    ops.assemble_tokens_bwd(dx, mask_u8, B, npat, C, dpe, grads.get(prefix + "cls_token"), grads.get(prefix + "mask_token"),
                            grads.get("pos_embed") if cfg.use_abs_pos_emb else None)
    ops.linear_wgrad(dpe, patches, grads[prefix + "patch_embed.proj.weight"].view(C, -1))
    ops.colsum_bf16(dpe, B * npat, C, grads[prefix + "patch_embed.proj.bias"])


def rel_bias(ps: ParamSource, cfg: VitConfig, dev, want_bwd: bool = True):
    """(bias_fwd, bias_bwd_t) padded log2(e)-scaled relative position bias, or (None, None). The dual-stream forward also gets the row maxima
    (bias_fwd.rowmax), the stabiliser of its single-pass softmax."""
    table = ps.f32("rel_pos_bias.relative_position_bias_table")
    if table is None:
        return None, None
    return ops.rel_pos_bias(table, ps.rel_index_i32(), cfg.tokens, cfg.num_heads, want_bwd, want_rowmax=cfg.dist, want_index_tiles=not cfg.dist)


# ------------------------------------------------------------------------------------------------------------------
# whole network
# ------------------------------------------------------------------------------------------------------------------
_MASK_STREAMS: Dict[int, torch.cuda.Stream] = {}


def _mask_stream(dev) -> torch.cuda.Stream:
    st = _MASK_STREAMS.get(dev.index)
    if st is None:
        st = _MASK_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return st


def draw_keep_bits(cfg: VitConfig, B: int, noise: Noise, dev, side_stream=None) -> None:
    """Draws the packed attention-dropout masks of all layers of one forward on a side stream (they depend only on the Philox key), so that
    the integer-multiply-bound draw (~40 us per ViT-B layer at B=128) runs beside whatever the main stream does — the EMA-teacher forward
    in the data2vec step, the previous layers of the same forward in MC-sample inference. One event per layer: block i of vit_forward /
    dist_forward waits for the mask of layer i only. No-op with injected masks."""
    if cfg.attn_drop_rate <= 0.0 or noise.attn_keep is not None or noise.attn_drop_active is False:
        return
    side_stream = side_stream if side_stream is not None else _mask_stream(dev)
    H, T = cfg.num_heads, cfg.tokens
    kb = torch.empty((cfg.depth, B, H, T, 32), dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream(dev)
    side_stream.wait_stream(cur)
    events = []
    with torch.cuda.stream(side_stream):
        for l in range(cfg.depth):
            ops.keep_bits(kb[l], B * H, T, cfg.attn_drop_rate, seed=noise.seed, stream_id=l, seed_dev=noise.seed_dev)
            events.append(side_stream.record_event())
    kb.record_stream(side_stream)
    noise.keep_bits_all, noise.keep_join = kb, events


def _keep_for_block(noise: Noise, kb_all, i: int, dev):
    if kb_all is None:
        return None
    if noise.keep_join is not None:
        torch.cuda.current_stream(dev).wait_event(noise.keep_join[i])
    return kb_all[i]


def make_drop_path_scales(cfg: VitConfig, B: int, noise: Noise, dev, draws: int = 2) -> Optional[torch.Tensor]:
    if not noise.drop_path_active or cfg.drop_path_rate <= 0.0:
        return None
    if noise.drop_path_scale is not None:
        return noise.drop_path_scale
    return ops.drop_path_scales(cfg.drop_path_probs, draws, B, noise.seed, dev)


def vit_forward(ps: ParamSource, cfg: VitConfig, images: torch.Tensor, *, mask_u8: Optional[torch.Tensor] = None,
                row_index: Optional[torch.Tensor] = None, mode: str = "masked", train: bool = False, save: bool = False,
                noise: Optional[Noise] = None, collect: Optional[List[int]] = None, collect_what: str = "end",
                patches: Optional[torch.Tensor] = None):
    """Deterministic (single-stream) ViT.
    mode: 'masked'  -> lm_head(norm(x)[:,1:][mask])            [R, C]   (row_index = flat rows b*T+1+p of the masked patches)
          'all'     -> lm_head(norm(x)[:,1:])                  [B, np, C]
          'layers'  -> residual streams of the blocks in `collect` (fp32 [B, T, C], cls row included), no head
          'logits'  -> head(fc_norm(mean_{t>=1} x))            [B, K]   (fine-tune model)
          'features'-> fc_norm(mean_{t>=1} x)                  [B, C]
    Returns (output, ctx); ctx holds what vit_backward needs when save=True."""
    if images.dtype != torch.float32 or not images.is_contiguous():
        images = images.float().contiguous()
    B = images.shape[0]
    T, C = cfg.tokens, cfg.embed_dim
    dev = images.device
    noise = noise or Noise()
    if patches is None:
        patches = patches_bf16(cfg, images)
    x = stem_forward(ps, cfg, patches, B, mask_u8)
    bias, bias_t = rel_bias(ps, cfg, dev, want_bwd=save)
    dps = make_drop_path_scales(cfg, B, noise, dev) if train else None
    attn_drop_on = train if noise.attn_drop_active is None else noise.attn_drop_active
    p_attn = cfg.attn_drop_rate if attn_drop_on else 0.0
    saved = []
    layers: Dict[int, torch.Tensor] = {}
    collect = collect or []
    if p_attn > 0 and noise.keep_bits_all is None:
        draw_keep_bits(cfg, B, noise, dev)                 # not pre-drawn by the engine: layer i+1's mask is drawn while layer i computes
    kb_all = noise.keep_bits_all if p_attn > 0 else None
    for i in range(cfg.depth):
        keep_in = noise.attn_keep[i] if (noise.attn_keep is not None and p_attn > 0) else None
        s = block_forward(ps, cfg, i, x, B, bias, save=save, dp_scale=dps[i] if dps is not None else None, p_attn=p_attn, seed=noise.seed,
                          keep_in=keep_in, seed_dev=noise.seed_dev, keep_pre=_keep_for_block(noise, kb_all, i, dev))
        if i in collect:
            if collect_what == "fc":      # fc_feature = x_out - x_mid (modeling_cyclical.py:203-205); rarely used
                layers[i] = (s["x_out"] - s["x_mid"]).view(B, T, C)
            else:
                layers[i] = s["x_out"].view(B, T, C)
        x = s["x_out"]
        if save:
            saved.append(s)
    ctx = dict(B=B, saved=saved, patches=patches, mask_u8=mask_u8, row_index=row_index, mode=mode, bias=bias_t, x_final=x) if save else None
    if mode == "layers":
        return layers, ctx
    if mode in ("masked", "all"):
        if mode == "all":
            row_index = all_patch_rows(B, T, dev)
        R = row_index.numel()
        hn = _empty((R, C), torch.bfloat16, dev)
        mean = _empty((R,), torch.float32, dev)
        rstd = _empty((R,), torch.float32, dev)
        out = _empty((R, C), torch.float32, dev)
        if R > 0:
            ops.layernorm_fwd(x, ps.f32("norm.weight"), ps.f32("norm.bias"), cfg.ln_eps, R, C, y_bf16=hn, mean=mean, rstd=rstd, row_index=row_index)
            ops.gemm(hn, ps.bf16("lm_head.weight"), R, C, C, epilogue=EPI_F32, bias=ps.f32("lm_head.bias"), out_f32=out)
        if save:
            ctx.update(hn=hn, hmean=mean, hrstd=rstd, row_index=row_index)
        return (out.view(B, T - 1, C) if mode == "all" else out), ctx
    if mode in ("logits", "features"):
        pooled = _empty((B, C), torch.float32, dev)
        ops.meanpool_tokens(x, B, T, C, pooled)
        feat = _empty((B, C), torch.float32, dev)
        fb = _empty((B, C), torch.bfloat16, dev)
        fmean = _empty((B,), torch.float32, dev)
        frstd = _empty((B,), torch.float32, dev)
        ops.layernorm_fwd(pooled, ps.f32("fc_norm.weight"), ps.f32("fc_norm.bias"), cfg.ln_eps, B, C, y_bf16=fb, y_f32=feat, mean=fmean, rstd=frstd)
        if save:
            ctx.update(pooled=pooled, fb=fb, fmean=fmean, frstd=frstd)
        if mode == "features":
            return feat, ctx
        K = cfg.num_classes
        w, hb = ps.head_padded()           # [Kp, C] bf16 / [Kp] fp32, rows >= K zero (GEMM needs N % 8 == 0)
        Kp = w.shape[0]
        logits = _empty((B, Kp), torch.float32, dev)
        ops.gemm(fb, w, B, Kp, C, epilogue=EPI_F32, bias=hb, out_f32=logits)
        return logits[:, :K], ctx
    raise B200VitError(f"unknown forward mode {mode!r}")


_ROWS_CACHE: Dict[tuple, torch.Tensor] = {}


def all_patch_rows(B: int, T: int, dev) -> torch.Tensor:
    key = (B, T, str(dev))
    t = _ROWS_CACHE.get(key)
    if t is None:
        r = torch.arange(B * T, dtype=torch.int32).view(B, T)[:, 1:].reshape(-1).contiguous()
        t = r.to(dev)
        _ROWS_CACHE[key] = t
    return t


def vit_backward(ps: ParamSource, cfg: VitConfig, ctx, dout: torch.Tensor, grads: Dict[str, torch.Tensor], dout_is_bf16_rows: bool = False,
                 after_block=None):
    """Accumulates parameter gradients into `grads` (fp32 tensors by reference name, caller zero-initialises).
    dout: gradient of the forward output ('masked'/'all' modes): fp32 or bf16 [R, C].
    after_block(i): called once block i's backward has been enqueued (the gradients of blocks i.. and of the head are final then)."""
    B = ctx["B"]
    T, C = cfg.tokens, cfg.embed_dim
    M = B * T
    dev = dout.device
    mode = ctx["mode"]
    if mode not in ("masked", "all"):
        raise B200VitError(f"backward is implemented for the data2vec modes ('masked', 'all'), not {mode!r}")
    row_index = ctx["row_index"]
    R = row_index.numel()
    dy = dout.reshape(R, C)
    if dy.dtype != torch.bfloat16:
        dy = ops.cast_bf16(dy.contiguous())
    dx = torch.zeros((M, C), dtype=torch.float32, device=dev)
    if R > 0:
        ops.linear_wgrad(dy, ctx["hn"], grads["lm_head.weight"])
        ops.colsum_bf16(dy, R, C, grads["lm_head.bias"])
        dhn = _empty((R, C), torch.bfloat16, dev)
        ops.gemm(dy, ps.bf16("lm_head.weight"), R, C, C, b_mn=True, epilogue=EPI_BF16, out_bf16=dhn)
        ops.layernorm_bwd(dhn, ctx["x_final"], ps.f32("norm.weight"), ctx["hmean"], ctx["hrstd"], R, C, dx, grads["norm.weight"], grads["norm.bias"],
                          row_index=row_index)
    ws = backward_workspace(cfg, B, dev)
    dtable = grads.get("rel_pos_bias.relative_position_bias_table")
    _backward_blocks(lambda i, sv, done, below: block_backward(ps, cfg, i, sv, dx, B, ctx["bias"], grads, dtable, ws, done, below),
                     cfg, ctx["saved"], after_block)
    stem_backward(ps, cfg, ctx["patches"], dx, B, ctx["mask_u8"], grads)


# ------------------------------------------------------------------------------------------------------------------
# dual-stream ("--stochastic") network: mean and covariance streams stacked along the row dimension ([2*B*T, C]: rows
# [0, M) = mean, [M, 2M) = covariance). The streams share norm1/norm2, qkv.weight (quirk §A.2-1), the MLP and gamma_1/2,
# so LayerNorm, fc1 and fc2 run ONCE over the stacked rows; QKV / proj differ only in bias / epilogue / weight.
# dist Block.forward (modeling_finetune_dist.py:41-59), dist Attention.forward (:111-179),
# DistVisionTransformerForCyclicalTraining.forward (modeling_cyclical_dist.py:108-165), DistVisionTransformer.forward (:280-326)
# ------------------------------------------------------------------------------------------------------------------
def dist_block_forward(ps: ParamSource, cfg: VitConfig, i: int, x_in: torch.Tensor, B: int, bias: torch.Tensor, *, save: bool,
                       dp_scale: Optional[torch.Tensor], p_attn: float, seed: int, keep_in: Optional[torch.Tensor],
                       seed_dev: Optional[torch.Tensor] = None, xwork_scratch: Optional[torch.Tensor] = None,
                       keep_pre: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    T, C, H, Hd = cfg.tokens, cfg.embed_dim, cfg.num_heads, cfg.hidden
    M = B * T
    M2 = 2 * M
    dev = x_in.device
    p = f"blocks.{i}."
    bf = torch.bfloat16
    h1 = _empty((M2, C), bf, dev)
    mean1 = _empty((M2,), torch.float32, dev)
    rstd1 = _empty((M2,), torch.float32, dev)
    ops.layernorm_fwd(x_in, ps.f32(p + "norm1.weight"), ps.f32(p + "norm1.bias"), cfg.ln_eps, M2, C, y_bf16=h1, mean=mean1, rstd=rstd1)
    qkv = _empty((M2, 3 * C), bf, dev)
    w = ps.bf16(p + "attn.qkv.weight")
    ops.gemm(h1[:M], w, M, 3 * C, C, epilogue=EPI_BF16, bias=ps.qkv_bias(p), out_bf16=qkv[:M])
    ops.gemm(h1[M:], w, M, 3 * C, C, epilogue=ops.EPI_ELU1, bias=ps.qkv_bias(p, cov=True), out_bf16=qkv[M:])     # elu(.)+1 (:127)
    att = _empty((M2, C), bf, dev)
    lse = _empty((B, H, T), torch.float32, dev) if save else None
    keep_bits = (keep_pre if keep_pre is not None else torch.empty((B, H, T, 32), dtype=torch.uint8, device=dev)) if p_attn > 0 else None
    # the transformed operands [sigmoid(q) | sqrt(sigmoid(cq))], [sigmoid(k) | sqrt(sigmoid(ck))]: kept for the backward when saving, else one
    # scratch buffer shared by all blocks of the forward
    xwork = ops.wattn_workspace(B, H, T, dev) if (save or xwork_scratch is None) else xwork_scratch
    ops.wattn_fwd(qkv[:M], qkv[M:], bias, B, H, T, (C // H) ** -0.5, p_attn, seed, i, keep_in, att[:M], att[M:], lse, keep_bits, seed_dev=seed_dev,
                  xwork=xwork, keep_ready=keep_pre is not None and p_attn > 0)
    x_mid = _empty((M2, C), torch.float32, dev)
    t1 = _empty((M2, C), bf, dev) if save else None
    g1 = ps.f32(p + "gamma_1") if cfg.has_gamma else None
    g2 = ps.f32(p + "gamma_2") if cfg.has_gamma else None
    dpv = (lambda d: dp_scale[d]) if dp_scale is not None else (lambda d: None)
    ops.gemm(att[:M], ps.bf16(p + "attn.proj.weight"), M, C, C, epilogue=EPI_RESIDUAL, bias=ps.f32(p + "attn.proj.bias"), colscale=g1,
             rowscale=dpv(0), rows_per_scale=T, residual=x_in[:M], out_f32=x_mid[:M], out2_bf16=t1[:M] if save else None)
    ops.gemm(att[M:], ps.bf16(p + "attn.cov_proj.weight"), M, C, C, epilogue=EPI_RESIDUAL, bias=ps.f32(p + "attn.cov_proj.bias"), colscale=g1,
             rowscale=dpv(2), rows_per_scale=T, residual=x_in[M:], out_f32=x_mid[M:], out2_bf16=t1[M:] if save else None)
    h2 = _empty((M2, C), bf, dev)
    mean2 = _empty((M2,), torch.float32, dev)
    rstd2 = _empty((M2,), torch.float32, dev)
    ops.layernorm_fwd(x_mid, ps.f32(p + "norm2.weight"), ps.f32(p + "norm2.bias"), cfg.ln_eps, M2, C, y_bf16=h2, mean=mean2, rstd=rstd2)
    act = _empty((M2, Hd), bf, dev)
    pre = _empty((M2, Hd), bf, dev) if save else None
    ops.gemm(h2, ps.bf16(p + "mlp.fc1.weight"), M2, Hd, C, epilogue=EPI_GELU, bias=ps.f32(p + "mlp.fc1.bias"), out_bf16=act, out2_bf16=pre)
    x_out = _empty((M2, C), torch.float32, dev)
    t2 = _empty((M2, C), bf, dev) if save else None
    dp_mlp = torch.cat((dp_scale[1], dp_scale[3])).contiguous() if dp_scale is not None else None      # draws 1 (mean) and 3 (cov)
    ops.gemm(act, ps.bf16(p + "mlp.fc2.weight"), M2, C, Hd, epilogue=EPI_RESIDUAL, bias=ps.f32(p + "mlp.fc2.bias"), colscale=g2,
             rowscale=dp_mlp, rows_per_scale=T, residual=x_mid, out_f32=x_out, out2_bf16=t2)
    if not save:
        return {"x_out": x_out, "x_mid": x_mid}
    return dict(x_in=x_in, h1=h1, mean1=mean1, rstd1=rstd1, qkv=qkv, att=att, lse=lse, keep_bits=keep_bits, xwork=xwork, t1=t1, x_mid=x_mid, h2=h2,
                mean2=mean2, rstd2=rstd2, act=act, pre=pre, t2=t2, x_out=x_out, dp_attn_m=dpv(0), dp_attn_c=dpv(2), dp_mlp=dp_mlp, p_attn=p_attn)


def dist_forward(ps: ParamSource, cfg: VitConfig, images: torch.Tensor, *, mask_u8: Optional[torch.Tensor] = None,
                 row_index: Optional[torch.Tensor] = None, mode: str = "masked", train: bool = False, save: bool = False,
                 noise: Optional[Noise] = None, collect: Optional[List[int]] = None):
    """Dual-stream ViT. Modes as vit_forward; outputs are (mean, cov) pairs:
       'masked' / 'all' -> (lm_head(norm(xm)[rows]), cov_lm_head(norm(xc)[rows])); 'layers' -> ({i: xm_i}, {i: xc_i});
       'logits' -> (fc_norm(mean-pool xm), fc_norm(mean-pool xc), head(mean feature))  (modeling_finetune_dist.py:311-326)."""
    if images.dtype != torch.float32 or not images.is_contiguous():
        images = images.float().contiguous()
    B = images.shape[0]
    T, C = cfg.tokens, cfg.embed_dim
    M = B * T
    dev = images.device
    noise = noise or Noise()
    patches = patches_bf16(cfg, images)
    x = _empty((2 * M, C), torch.float32, dev)
    stem_forward(ps, cfg, patches, B, mask_u8, x=x[:M])
    stem_forward(ps, cfg, patches, B, mask_u8, x=x[M:], prefix="cov_")
    bias, bias_t = rel_bias(ps, cfg, dev, want_bwd=save)
    if bias is None:
        raise B200VitError("the dual-stream model needs use_shared_rel_pos_bias=True (the reference adds rel_pos_bias unconditionally)")
    dps = make_drop_path_scales(cfg, B, noise, dev, draws=4) if train else None
    attn_drop_on = train if noise.attn_drop_active is None else noise.attn_drop_active
    p_attn = cfg.attn_drop_rate if attn_drop_on else 0.0
    saved, lm, lc = [], {}, {}
    collect = collect or []
    scratch = None if save else ops.wattn_workspace(B, cfg.num_heads, T, dev)
    if p_attn > 0 and noise.keep_bits_all is None:
        draw_keep_bits(cfg, B, noise, dev)
    kb_all = noise.keep_bits_all if p_attn > 0 else None
    for i in range(cfg.depth):
        keep_in = noise.attn_keep[i] if (noise.attn_keep is not None and p_attn > 0) else None
        s = dist_block_forward(ps, cfg, i, x, B, bias, save=save, dp_scale=dps[i] if dps is not None else None, p_attn=p_attn,
                               seed=noise.seed, keep_in=keep_in, seed_dev=noise.seed_dev, xwork_scratch=scratch,
                               keep_pre=_keep_for_block(noise, kb_all, i, dev))
        x = s["x_out"]
        if i in collect:
            lm[i] = x[:M].view(B, T, C)
            lc[i] = x[M:].view(B, T, C)
        if save:
            saved.append(s)
    ctx = dict(B=B, saved=saved, patches=patches, mask_u8=mask_u8, row_index=row_index, mode=mode, bias=bias_t, x_final=x) if save else None
    if mode == "layers":
        return (lm, lc), ctx
    if mode in ("masked", "all"):
        if mode == "all":
            row_index = all_patch_rows(B, T, dev)
        R = row_index.numel()
        rows2 = torch.cat((row_index, row_index + M)).contiguous()
        hn = _empty((2 * R, C), torch.bfloat16, dev)
        mean = _empty((2 * R,), torch.float32, dev)
        rstd = _empty((2 * R,), torch.float32, dev)
        om = _empty((R, C), torch.float32, dev)
        oc = _empty((R, C), torch.float32, dev)
        if R > 0:
            ops.layernorm_fwd(x, ps.f32("norm.weight"), ps.f32("norm.bias"), cfg.ln_eps, 2 * R, C, y_bf16=hn, mean=mean, rstd=rstd, row_index=rows2)
            ops.gemm(hn[:R], ps.bf16("lm_head.weight"), R, C, C, epilogue=EPI_F32, bias=ps.f32("lm_head.bias"), out_f32=om)
            ops.gemm(hn[R:], ps.bf16("cov_lm_head.weight"), R, C, C, epilogue=EPI_F32, bias=ps.f32("cov_lm_head.bias"), out_f32=oc)
        if save:
            ctx.update(hn=hn, hmean=mean, hrstd=rstd, row_index=row_index, rows2=rows2)
        if mode == "all":
            return (om.view(B, T - 1, C), oc.view(B, T - 1, C)), ctx
        return (om, oc), ctx
    if mode == "logits":
        pooled = _empty((2 * B, C), torch.float32, dev)
        ops.meanpool_tokens(x, 2 * B, T, C, pooled)
        feat = _empty((2 * B, C), torch.float32, dev)
        fb = _empty((2 * B, C), torch.bfloat16, dev)
        fmean = _empty((2 * B,), torch.float32, dev)
        frstd = _empty((2 * B,), torch.float32, dev)
        ops.layernorm_fwd(pooled, ps.f32("fc_norm.weight"), ps.f32("fc_norm.bias"), cfg.ln_eps, 2 * B, C, y_bf16=fb, y_f32=feat, mean=fmean, rstd=frstd)
        head_in = fb[:B]
        eps = None
        if cfg.sample_head or noise.head_eps is not None:
            head_in = _empty((B, C), torch.bfloat16, dev)
            _, _, eps = ops.gaussian_sample(feat[:B], feat[B:], eps_in=noise.head_eps, seed=noise.seed, stream_id=0x48454144, want_eps=save,
                                            out_bf16=head_in)
        if save:
            ctx.update(pooled=pooled, fb=head_in, fmean=fmean, frstd=frstd, head_eps=eps, feat_cov=feat[B:] if eps is not None else None)
        w, hb = ps.head_padded()
        Kp = w.shape[0]
        logits = _empty((B, Kp), torch.float32, dev)
        ops.gemm(head_in, w, B, Kp, C, epilogue=EPI_F32, bias=hb, out_f32=logits)
        return (feat[:B], feat[B:], logits[:, :cfg.num_classes]), ctx
    raise B200VitError(f"unknown forward mode {mode!r}")


def dist_block_backward(ps: ParamSource, cfg: VitConfig, i: int, s: Dict[str, torch.Tensor], dx: torch.Tensor, B: int, bias_t: torch.Tensor,
                        grads: Dict[str, torch.Tensor], dtable: Optional[torch.Tensor], ws: Dict[str, torch.Tensor],
                        first_srb_done: bool = False, below: Optional[Dict[str, torch.Tensor]] = None):
    """dx: fp32 [2*B*T, C] (mean rows then cov rows), updated IN PLACE. Shared weights (qkv, fc1, fc2, norms, gammas) receive the sum of
    both streams' gradients simply because the GEMMs / reductions run over the stacked rows. LayerNorm backward fused with the following
    scale-residual backward as in block_backward (norm2: one launch per stream, because proj / cov_proj have their own bias gradients)."""
    T, C, H, Hd = cfg.tokens, cfg.embed_dim, cfg.num_heads, cfg.hidden
    M = B * T
    M2 = 2 * M
    p = f"blocks.{i}."
    g = lambda n: grads.get(p + n)
    dt, dpre, dh, dqkv = ws["dt"], ws["dpre"], ws["dh"], ws["dqkv"]
    g1 = ps.f32(p + "gamma_1") if cfg.has_gamma else None
    g2 = ps.f32(p + "gamma_2") if cfg.has_gamma else None
    # ---- shared MLP over the stacked rows
    if not first_srb_done:
        ops.scale_residual_bwd(dx, s["t2"], s["dp_mlp"], T, g2, M2, C, dt, g("gamma_2") if cfg.has_gamma else None, g("mlp.fc2.bias"))
    ops.linear_wgrad(dt, s["act"], g("mlp.fc2.weight"))
    ops.gemm(dt, ps.bf16(p + "mlp.fc2.weight"), M2, Hd, C, b_mn=True, epilogue=EPI_DGELU, aux=s["pre"], out_bf16=dpre, colsum=g("mlp.fc1.bias"))
    ops.linear_wgrad(dpre, s["h2"], g("mlp.fc1.weight"))
    ops.gemm(dpre, ps.bf16(p + "mlp.fc1.weight"), M2, C, Hd, b_mn=True, epilogue=EPI_BF16, out_bf16=dh)
    # ---- norm2 backward + attention branch: mean rows through proj, cov rows through cov_proj (shared gamma_1)
    gg1 = g("gamma_1") if cfg.has_gamma else None
    n2w, gn2w, gn2b = ps.f32(p + "norm2.weight"), g("norm2.weight"), g("norm2.bias")
    ops.layernorm_bwd_scale_residual(dh[:M], s["x_mid"][:M], n2w, s["mean2"][:M], s["rstd2"][:M], M, C, dx[:M], gn2w, gn2b,
                                     s["t1"][:M], s["dp_attn_m"], T, g1, dt[:M], gg1, g("attn.proj.bias"))
    ops.layernorm_bwd_scale_residual(dh[M:], s["x_mid"][M:], n2w, s["mean2"][M:], s["rstd2"][M:], M, C, dx[M:], gn2w, gn2b,
                                     s["t1"][M:], s["dp_attn_c"], T, g1, dt[M:], gg1, g("attn.cov_proj.bias"))
    att = s["att"]
    ops.linear_wgrad(dt[:M], att[:M], g("attn.proj.weight"))
    ops.linear_wgrad(dt[M:], att[M:], g("attn.cov_proj.weight"))
    ops.gemm(dt[:M], ps.bf16(p + "attn.proj.weight"), M, C, C, b_mn=True, epilogue=EPI_BF16, out_bf16=dh[:M])
    ops.gemm(dt[M:], ps.bf16(p + "attn.cov_proj.weight"), M, C, C, b_mn=True, epilogue=EPI_BF16, out_bf16=dh[M:])
    qkv = s["qkv"]
    ops.wattn_bwd(qkv[:M], qkv[M:], s["xwork"], att[:M], att[M:], dh[:M], dh[M:], s["lse"], bias_t, s["keep_bits"],
                  ps.rel_index_i32() if dtable is not None else None, dtable, B, H, T, (C // H) ** -0.5, s["p_attn"], dqkv[:M], dqkv[M:],
                  work=ws["wattn"], dq_bias=g("attn.q_bias"), dv_bias=g("attn.v_bias"), dcq_bias=g("attn.cov_q_bias"), dcv_bias=g("attn.cov_v_bias"))
    ops.linear_wgrad(dqkv, s["h1"], g("attn.qkv.weight"))        # both streams multiply by qkv.weight (cov_qkv.weight stays unused, §A.2-1)
    ops.gemm(dqkv, ps.bf16(p + "attn.qkv.weight"), M2, C, 3 * C, b_mn=True, epilogue=EPI_BF16, out_bf16=dh)
    if below is None:
        ops.layernorm_bwd(dh, s["x_in"], ps.f32(p + "norm1.weight"), s["mean1"], s["rstd1"], M2, C, dx, g("norm1.weight"), g("norm1.bias"))
    else:
        q = f"blocks.{i - 1}."
        ops.layernorm_bwd_scale_residual(dh, s["x_in"], ps.f32(p + "norm1.weight"), s["mean1"], s["rstd1"], M2, C, dx, g("norm1.weight"), g("norm1.bias"),
                                         below["t2"], below["dp_mlp"], T, ps.f32(q + "gamma_2") if cfg.has_gamma else None, dt,
                                         grads.get(q + "gamma_2") if cfg.has_gamma else None, grads.get(q + "mlp.fc2.bias"))


def dist_backward(ps: ParamSource, cfg: VitConfig, ctx, dout_m: torch.Tensor, dout_c: torch.Tensor, grads: Dict[str, torch.Tensor], after_block=None):
    """Backward of dist_forward in the 'masked' / 'all' modes: gradients of (mean output, cov output) -> parameter gradients."""
    B = ctx["B"]
    T, C = cfg.tokens, cfg.embed_dim
    M = B * T
    dev = dout_m.device
    if ctx["mode"] not in ("masked", "all"):
        raise B200VitError("dual-stream backward is implemented for the data2vec modes ('masked', 'all')")
    row_index, rows2 = ctx["row_index"], ctx["rows2"]
    R = row_index.numel()
    dx = torch.zeros((2 * M, C), dtype=torch.float32, device=dev)
    if R > 0:
        dym = ops.cast_bf16(dout_m.reshape(R, C).float().contiguous())
        dyc = ops.cast_bf16(dout_c.reshape(R, C).float().contiguous())
        hn = ctx["hn"]
        ops.linear_wgrad(dym, hn[:R], grads["lm_head.weight"])
        ops.linear_wgrad(dyc, hn[R:], grads["cov_lm_head.weight"])
        ops.colsum_bf16(dym, R, C, grads["lm_head.bias"])
        ops.colsum_bf16(dyc, R, C, grads["cov_lm_head.bias"])
        dhn = _empty((2 * R, C), torch.bfloat16, dev)
        ops.gemm(dym, ps.bf16("lm_head.weight"), R, C, C, b_mn=True, epilogue=EPI_BF16, out_bf16=dhn[:R])
        ops.gemm(dyc, ps.bf16("cov_lm_head.weight"), R, C, C, b_mn=True, epilogue=EPI_BF16, out_bf16=dhn[R:])
        ops.layernorm_bwd(dhn, ctx["x_final"], ps.f32("norm.weight"), ctx["hmean"], ctx["hrstd"], 2 * R, C, dx, grads["norm.weight"],
                          grads["norm.bias"], row_index=rows2)
    M2, Hd = 2 * M, cfg.hidden
    bf = torch.bfloat16
    ld = (T + 15) // 16 * 16
    dtable = grads.get("rel_pos_bias.relative_position_bias_table")
    ws = dict(dt=_empty((M2, C), bf, dev), dpre=_empty((M2, Hd), bf, dev), dh=_empty((M2, C), bf, dev), dqkv=_empty((M2, 3 * C), bf, dev),
              wattn=ops.wattn_bwd_workspace(B, cfg.num_heads, T, dtable is not None, dev))
    _backward_blocks(lambda i, sv, done, below: dist_block_backward(ps, cfg, i, sv, dx, B, ctx["bias"], grads, dtable, ws, done, below),
                     cfg, ctx["saved"], after_block)
    stem_backward(ps, cfg, ctx["patches"], dx[:M], B, ctx["mask_u8"], grads)
    stem_backward(ps, cfg, ctx["patches"], dx[M:], B, ctx["mask_u8"], grads, prefix="cov_")



# ------------------------------------------------------------------------------------------------------------------
# classifier (fine-tune) backward: logits = head(fc_norm(mean_{t>=1} x_L))   (modeling_finetune.py:512-523, modeling_finetune_dist.py:311-326)
# ------------------------------------------------------------------------------------------------------------------
def _head_backward(ps: ParamSource, cfg: VitConfig, ctx, dlogits: torch.Tensor, dfeat_extra: Optional[torch.Tensor], grads, streams: int) -> torch.Tensor:
    """Returns dx [streams*B*T, C] fp32 (gradient of the final residual stream). dlogits: fp32 [B, K], or bf16 [B, Kp] already zero-padded to
    the GEMM's N (b200vit_finetune_loss writes that directly). dfeat_extra: optional fp32 [streams*B, C] gradient flowing directly into the
    fc_norm outputs (the W-loss of the dual-stream fine-tune step); it is consumed (accumulated into)."""
    B = ctx["B"]
    T, C, K = cfg.tokens, cfg.embed_dim, cfg.num_classes
    dev = dlogits.device
    SB = streams * B
    w, _ = ps.head_padded()
    Kp = w.shape[0]
    if dlogits.dtype == torch.bfloat16 and dlogits.shape[1] == Kp:
        dl16 = dlogits
    else:
        dl = torch.zeros((B, Kp), dtype=torch.float32, device=dev)
        dl[:, :K].copy_(dlogits)
        dl16 = ops.cast_bf16(dl)
    fb = ctx["fb"]
    gw, gb = grads.get("head.weight__padded"), grads.get("head.bias__padded")
    if gw is not None and gb is not None:            # engine arenas reserve the padded classifier rows: accumulate in place
        ops.linear_wgrad(dl16, fb[:B], gw)
        ops.colsum_bf16(dl16, B, Kp, gb)
    else:
        tmp = torch.zeros((Kp, C), dtype=torch.float32, device=dev)
        tmpb = torch.zeros((Kp,), dtype=torch.float32, device=dev)
        ops.linear_wgrad(dl16, fb[:B], tmp)
        ops.colsum_bf16(dl16, B, Kp, tmpb)
        grads["head.weight"].add_(tmp[:K])
        grads["head.bias"].add_(tmpb[:K])
    dfeat = dfeat_extra if dfeat_extra is not None else torch.zeros((SB, C), dtype=torch.float32, device=dev)
    dhead = torch.empty((B, C), dtype=torch.float32, device=dev)
    ops.gemm(dl16, w, B, C, Kp, b_mn=True, epilogue=EPI_F32, out_f32=dhead)
    dfeat[:B].add_(dhead)
    if ctx.get("head_eps") is not None:      # head input = mean + sqrt(max(cov, 0)) * eps: the cov feature receives dhead * eps / (2 sqrt(cov))
        ops.gaussian_sample_bwd(dhead, ctx["feat_cov"].contiguous(), ctx["head_eps"], dcov=dfeat[B:])
    dpool = torch.zeros((SB, C), dtype=torch.float32, device=dev)
    ops.layernorm_bwd(dfeat.contiguous(), ctx["pooled"], ps.f32("fc_norm.weight"), ctx["fmean"], ctx["frstd"], SB, C, dpool, grads["fc_norm.weight"],
                      grads["fc_norm.bias"])
    dx = torch.zeros((SB * T, C), dtype=torch.float32, device=dev)
    ops.meanpool_tokens_bwd(dpool, SB, T, C, dx)
    return dx


def vit_backward_logits(ps: ParamSource, cfg: VitConfig, ctx, dlogits: torch.Tensor, grads: Dict[str, torch.Tensor]):
    B = ctx["B"]
    dx = _head_backward(ps, cfg, ctx, dlogits, None, grads, 1)
    ws = backward_workspace(cfg, B, dlogits.device)
    dtable = grads.get("rel_pos_bias.relative_position_bias_table")
    _backward_blocks(lambda i, sv, done, below: block_backward(ps, cfg, i, sv, dx, B, ctx["bias"], grads, dtable, ws, done, below),
                     cfg, ctx["saved"], None)
    stem_backward(ps, cfg, ctx["patches"], dx, B, ctx["mask_u8"], grads)


def dist_backward_logits(ps: ParamSource, cfg: VitConfig, ctx, dmean_feat, dcov_feat, dlogits, grads: Dict[str, torch.Tensor]):
    B = ctx["B"]
    T, C = cfg.tokens, cfg.embed_dim
    M = B * T
    dev = dlogits.device
    extra = torch.zeros((2 * B, C), dtype=torch.float32, device=dev)
    if dmean_feat is not None:
        extra[:B].copy_(dmean_feat)
    if dcov_feat is not None:
        extra[B:].copy_(dcov_feat)
    dx = _head_backward(ps, cfg, ctx, dlogits, extra, grads, 2)
    M2, Hd = 2 * M, cfg.hidden
    bf = torch.bfloat16
    ld = (T + 15) // 16 * 16
    dtable = grads.get("rel_pos_bias.relative_position_bias_table")
    ws = dict(dt=_empty((M2, C), bf, dev), dpre=_empty((M2, Hd), bf, dev), dh=_empty((M2, C), bf, dev), dqkv=_empty((M2, 3 * C), bf, dev),
              wattn=ops.wattn_bwd_workspace(B, cfg.num_heads, T, dtable is not None, dev))
    _backward_blocks(lambda i, sv, done, below: dist_block_backward(ps, cfg, i, sv, dx, B, ctx["bias"], grads, dtable, ws, done, below),
                     cfg, ctx["saved"], None)
    stem_backward(ps, cfg, ctx["patches"], dx[:M], B, None, grads)
    stem_backward(ps, cfg, ctx["patches"], dx[M:], B, None, grads, prefix="cov_")
