"""nn.Module boundary: the reference's model classes, parameter names, constructor kwargs, forward signatures and model-registry
names, with every forward/backward executed by the sm_100a kernels of libb200vit.so (core.py schedules).

Drop-in for `create_model('beit_base_patch16_224' | 'beit_large_patch16_224' | 'dist_beit_base_patch16_224', ...)` as called by
run_cyclical.py:289-302 (import this module instead of `modeling_cyclical`) and run_class_finetuning.py:348-372 (instead of
`modeling_finetune`); SURVEY.md §8(b). Parameters are ordinary fp32 nn.Parameters named exactly as in the reference
(§A.4), so checkpoints, optim_factory.get_parameter_groups, ModelEmaV2 and DDP keep working.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import core, ops
from ._lib import B200VitError
from .core import Noise, VitConfig


# ------------------------------------------------------------------------------------------------------------------
# parameter containers (same attribute / parameter names as the reference modules)
# ------------------------------------------------------------------------------------------------------------------
class DropPath(nn.Module):
    """modeling_finetune.py:51-62 (class name matters: enable_dropout() must NOT re-enable it, uncertainty_evaluations.py:35-39)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def extra_repr(self):
        return "p={}".format(self.drop_prob)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):
    def __init__(self, dim, num_heads, attn_drop=0.0, proj_drop=0.0, dist=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if dist:
            self.cov_qkv = nn.Linear(dim, dim * 3, bias=False)     # allocated but unused, as in the reference (§A.2-1)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        if dist:
            self.cov_q_bias = nn.Parameter(torch.zeros(dim))
            self.cov_v_bias = nn.Parameter(torch.zeros(dim))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        if dist:
            self.cov_proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, drop, attn_drop, drop_path, init_values, norm_layer, dist=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads, attn_drop, drop, dist)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), drop)
        if init_values is not None and init_values > 0:
            self.gamma_1 = nn.Parameter(init_values * torch.ones(dim))
            self.gamma_2 = nn.Parameter(init_values * torch.ones(dim))
        else:
            self.gamma_1, self.gamma_2 = None, None


class PatchEmbed(nn.Module):
    """modeling_finetune.py:304-325 (attributes patch_size / patch_shape / num_patches are read by the runners)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.patch_shape = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)


def relative_position_index(wh: int, ww: int) -> torch.Tensor:
    """Integer index buffer of RelativePositionBias (modeling_finetune.py:339-353)."""
    nrd = (2 * wh - 1) * (2 * ww - 1) + 3
    ys = torch.arange(wh).repeat_interleave(ww)
    xs = torch.arange(ww).repeat(wh)
    idx = torch.zeros((wh * ww + 1,) * 2, dtype=torch.int64)
    idx[1:, 1:] = (ys[:, None] - ys[None, :] + wh - 1) * (2 * ww - 1) + (xs[:, None] - xs[None, :] + ww - 1)
    idx[0, :] = nrd - 3
    idx[:, 0] = nrd - 2
    idx[0, 0] = nrd - 1
    return idx


class RelativePositionBias(nn.Module):
    def __init__(self, window_size, num_heads):
        super().__init__()
        self.window_size = window_size
        self.num_relative_distance = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) + 3
        self.relative_position_bias_table = nn.Parameter(torch.zeros(self.num_relative_distance, num_heads))
        self.register_buffer("relative_position_index", relative_position_index(*window_size))

    def forward(self):
        T, H = self.relative_position_index.shape[0], self.relative_position_bias_table.shape[1]
        fwd, _ = ops.rel_pos_bias(self.relative_position_bias_table.detach(), self.relative_position_index.to(torch.int32), T, H, False)
        return fwd[:, :, :T] / ops.LOG2E            # [H, T, T] in natural units, as the reference module returns


# ------------------------------------------------------------------------------------------------------------------
# ParamSource over an nn.Module (bf16 shadows refreshed when a parameter's version counter or storage changes)
# ------------------------------------------------------------------------------------------------------------------
class ModuleParams(core.ParamSource):
    def __init__(self, module: nn.Module):
        self.module = module
        self._cache: Dict[str, tuple] = {}
        self._named: Optional[Dict[str, torch.Tensor]] = None

    def _lookup(self, name: str) -> Optional[torch.Tensor]:
        if self._named is None:
            self._named = dict(self.module.named_parameters())
            self._named.update(dict(self.module.named_buffers()))
        return self._named.get(name)

    def invalidate(self):
        self._named = None
        self._cache.clear()

    def f32(self, name):
        t = self._lookup(name)
        return None if t is None else t.detach()

    def _cached(self, key, src_tensors, build):
        sig = (getattr(self.module, "_weights_version", 0),) + tuple((t._version, t.data_ptr()) for t in src_tensors)
        ent = self._cache.get(key)
        if ent is None or ent[0] != sig:
            ent = (sig, build())
            self._cache[key] = ent
        return ent[1]

    def bf16(self, name):
        t = self._lookup(name)
        return self._cached("bf16:" + name, [t], lambda: ops.cast_bf16(t.detach().reshape(t.shape[0], -1).contiguous()))

    def qkv_bias(self, prefix, cov=False):
        q = self._lookup(prefix + ("attn.cov_q_bias" if cov else "attn.q_bias"))
        v = self._lookup(prefix + ("attn.cov_v_bias" if cov else "attn.v_bias"))

        def build():
            out = torch.zeros(3 * q.numel(), dtype=torch.float32, device=q.device)
            out[: q.numel()].copy_(q.detach())
            out[2 * q.numel():].copy_(v.detach())
            return out
        return self._cached("qkvb:" + prefix + str(cov), [q, v], build)

    def rel_index_i32(self):
        t = self._lookup("rel_pos_bias.relative_position_index")
        return self._cached("relidx", [t], lambda: t.to(torch.int32).contiguous())

    def head_padded(self):
        w, b = self._lookup("head.weight"), self._lookup("head.bias")

        def build():
            K, C = w.shape
            Kp = (K + 7) // 8 * 8
            wp = torch.zeros(Kp, C, dtype=torch.float32, device=w.device)
            wp[:K].copy_(w.detach())
            bp = torch.zeros(Kp, dtype=torch.float32, device=w.device)
            bp[:K].copy_(b.detach())
            return ops.cast_bf16(wp), bp
        return self._cached("head", [w, b], build)


# ------------------------------------------------------------------------------------------------------------------
# autograd bridge: one Function for the whole network (forward saves activations in a python ctx; backward runs
# core.vit_backward and hands the per-parameter gradients back to autograd in named_parameters() order)
# ------------------------------------------------------------------------------------------------------------------
class _VitFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, images, mask_u8, row_index, mode, noise, names, *params):
        out, saved = core.vit_forward(model._ps, model.cfg, images, mask_u8=mask_u8, row_index=row_index, mode=mode, train=model.training,
                                      save=True, noise=noise)
        ctx.model, ctx.saved, ctx.names, ctx.shapes = model, saved, names, [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        model = ctx.model
        dev = dout.device
        grads = {n: torch.zeros(s, dtype=torch.float32, device=dev) for n, s in zip(ctx.names, ctx.shapes)}
        if ctx.saved["mode"] == "logits":
            core.vit_backward_logits(model._ps, model.cfg, ctx.saved, dout.float().contiguous(), grads)
        else:
            core.vit_backward(model._ps, model.cfg, ctx.saved, dout.contiguous(), grads)
        ctx.saved = None
        unused = model._unused_param_names()
        return (None, None, None, None, None, None, None) + tuple(None if n in unused else grads[n] for n in ctx.names)


def _rows_from_mask(mask_flat_bool: torch.Tensor, T: int) -> torch.Tensor:
    """Flat row numbers (b*T + 1 + p) of the masked patches in row-major (b, p) order — the order of the reference's
    boolean gather x.reshape(-1, C)[mask] (modeling_cyclical.py:222-224). torch.nonzero synchronises, as the reference does."""
    idx = torch.nonzero(mask_flat_bool, as_tuple=False).flatten()
    npat = T - 1
    return (idx // npat * T + 1 + idx % npat).to(torch.int32)


class _VitBase(nn.Module):
    cfg: VitConfig

    def _finish_init(self):
        self._ps = ModuleParams(self)
        self._weights_version = 0        # bumped by the fused engines: their kernels update the parameter arena behind autograd's version counters
        self._seed_calls = 0
        self._injected: Optional[Noise] = None

    def _apply(self, fn, *a, **k):          # .to(device) / .cuda() replace parameter storage
        r = super()._apply(fn, *a, **k)
        if hasattr(self, "_ps"):
            self._ps.invalidate()
        return r

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_ps":
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._ps = ModuleParams(new)
        return new

    def get_num_layers(self):
        return len(self.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def inject_noise(self, drop_path_keep: Optional[List[torch.Tensor]] = None, attn_keep: Optional[List[torch.Tensor]] = None):
        """Parity hook: next training forward uses these instead of the Philox streams.
        drop_path_keep: per layer float [draws, B] in {0,1}; attn_keep: per layer uint8 [B, H, N, N]."""
        dev = self.cls_token.device
        n = Noise()
        if drop_path_keep is not None:
            probs = self.cfg.drop_path_probs
            n.drop_path_scale = torch.stack([k.float() / (1.0 - p) for k, p in zip(drop_path_keep, probs)]).to(dev).contiguous()
        if attn_keep is not None:
            n.attn_keep = [k.to(dev).to(torch.uint8).contiguous() for k in attn_keep]
        self._injected = n

    def _noise(self) -> Noise:
        if self._injected is not None:
            n, self._injected = self._injected, None
        else:
            n = Noise()
        self._seed_calls += 1
        n.seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * self._seed_calls) & 0xFFFFFFFFFFFFFFFF
        blk = self.blocks[0]
        n.drop_path_active = self.training
        n.attn_drop_active = blk.attn.attn_drop.training and blk.attn.attn_drop.p > 0       # enable_dropout() semantics
        return n

    def _unused_param_names(self):
        return set()

    def _check_input(self, x):
        if not x.is_cuda:
            raise B200VitError("this model runs only on a CUDA (B200) device: there is no CPU fallback")
        H = self.patch_embed.img_size[0]
        assert x.shape[2] == H and x.shape[3] == self.patch_embed.img_size[1], \
            f"Input image size ({x.shape[2]}*{x.shape[3]}) doesn't match model ({H}*{self.patch_embed.img_size[1]})."

    def _run(self, x, mask_u8, row_index, mode, collect=None, collect_what="end"):
        noise = self._noise()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad and mode in ("masked", "all", "logits"):
            named = list(self.named_parameters())
            names = [n for n, _ in named]
            return _VitFunction.apply(self, x, mask_u8, row_index, mode, noise, names, *[p for _, p in named])
        out, _ = core.vit_forward(self._ps, self.cfg, x, mask_u8=mask_u8, row_index=row_index, mode=mode, train=self.training, save=False,
                                  noise=noise, collect=collect, collect_what=collect_what)
        return out


def _trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _reject_unsupported(**flags):
    on = [k for k, v in flags.items() if v]
    if on:
        raise NotImplementedError(f"ablation flags {on} are outside the B200 hot path (SURVEY.md §2: out of scope); no silent fallback")


class VisionTransformerForCyclicalTraining(_VitBase):
    """data2vec student / teacher network (modeling_cyclical.py:33-225)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=None, init_values=None, attn_head_dim=None,
                 use_abs_pos_emb=True, use_rel_pos_bias=False, use_shared_rel_pos_bias=False, init_std=0.02, gp_layer=False,
                 gumbel_softmax=False, sinkformer=False, h_sto_trans=False, stosa=False, **unused):
        super().__init__()
        _reject_unsupported(gp_layer=gp_layer, gumbel_softmax=gumbel_softmax, sinkformer=sinkformer, h_sto_trans=h_sto_trans, stosa=stosa,
                            use_rel_pos_bias=use_rel_pos_bias, drop_rate=drop_rate > 0, attn_head_dim=attn_head_dim is not None,
                            qk_scale=qk_scale is not None, no_qkv_bias=not qkv_bias)
        if embed_dim // num_heads != 64:
            raise NotImplementedError("head_dim must be 64 (ViT-B/16, ViT-L/16)")
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim)) if use_abs_pos_emb else None
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.rel_pos_bias = RelativePositionBias(self.patch_embed.patch_shape, num_heads) if use_shared_rel_pos_bias else None
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, drop_rate, attn_drop_rate, dpr[i], init_values, norm_layer)
                                     for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.init_std = init_std
        self.lm_head = nn.Linear(embed_dim, embed_dim)
        tn = lambda t: _trunc_normal_(t, std=init_std, a=-init_std, b=init_std)          # modeling_cyclical.py:23-24
        if self.pos_embed is not None:
            tn(self.pos_embed)
        tn(self.cls_token)
        tn(self.mask_token)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                tn(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        for layer_id, layer in enumerate(self.blocks):                                  # fix_init_weight, modeling_cyclical.py:141-147
            layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
            layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
        eps = self.norm.eps
        self.cfg = VitConfig(img_size=self.patch_embed.img_size[0], patch_size=self.patch_embed.patch_size[0], in_chans=in_chans,
                             embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, ln_eps=eps, kind="cyclical",
                             drop_path_rate=drop_path_rate, attn_drop_rate=attn_drop_rate, has_gamma=self.blocks[0].gamma_1 is not None,
                             use_abs_pos_emb=use_abs_pos_emb)
        self._finish_init()

    def forward(self, x, bool_masked_pos, return_all_tokens=False, layer_results=None):
        self._check_input(x)
        B = x.shape[0]
        T = self.cfg.tokens
        mask_u8 = None
        mflat = None
        if bool_masked_pos is not None:
            mflat = bool_masked_pos.reshape(B, -1).to(x.device) != 0
            mask_u8 = mflat.to(torch.uint8).reshape(-1).contiguous()
        if layer_results:
            which = layer_results if layer_results in ("end", "fc") else None
            if which is None:
                return []                                       # any other truthy string yields an empty list (§A.2-7)
            with torch.no_grad():
                layers = self._run(x, mask_u8, None, "layers", collect=list(range(self.cfg.depth)), collect_what=which)
            return [layers[i][:, 1:] for i in range(self.cfg.depth)]
        if return_all_tokens:
            return self._run(x, mask_u8, None, "all")
        rows = _rows_from_mask(mflat.reshape(-1), T)
        return self._run(x, mask_u8, rows, "masked")


class VisionTransformer(_VitBase):
    """Fine-tune / inference classifier (modeling_finetune.py:367-523), mean-pool head."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm,
                 init_values=None, use_abs_pos_emb=True, use_rel_pos_bias=False, use_shared_rel_pos_bias=False, use_mean_pooling=True,
                 init_scale=0.001, linear_classifier=False, has_masking=False, learn_layer_weights=False, layernorm_before_combine=False,
                 gp_layer=False, het_layer=False, sinkformer=False, gumbel_softmax=False, h_sto_trans=False, sngp=False, **unused):
        super().__init__()
        _reject_unsupported(gp_layer=gp_layer, het_layer=het_layer, sinkformer=sinkformer, gumbel_softmax=gumbel_softmax, h_sto_trans=h_sto_trans,
                            sngp=sngp, use_rel_pos_bias=use_rel_pos_bias, learn_layer_weights=learn_layer_weights, drop_rate=drop_rate > 0,
                            no_mean_pooling=not use_mean_pooling, qk_scale=qk_scale is not None, no_qkv_bias=not qkv_bias,
                            linear_classifier=linear_classifier)
        if embed_dim // num_heads != 64:
            raise NotImplementedError("head_dim must be 64 (ViT-B/16, ViT-L/16)")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        if has_masking:
            self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim)) if use_abs_pos_emb else None
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.rel_pos_bias = RelativePositionBias(self.patch_embed.patch_shape, num_heads) if use_shared_rel_pos_bias else None
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.use_rel_pos_bias = use_rel_pos_bias
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, drop_rate, attn_drop_rate, dpr[i], init_values, norm_layer)
                                     for i in range(depth)])
        self.use_mean_pooling = use_mean_pooling
        self.norm = nn.Identity()
        self.fc_norm = norm_layer(embed_dim)
        self.het_layer = het_layer
        self.head = nn.Linear(embed_dim, num_classes)
        if self.pos_embed is not None:
            _trunc_normal_(self.pos_embed, std=0.02)
        _trunc_normal_(self.cls_token, std=0.02)
        if has_masking:
            _trunc_normal_(self.mask_token, std=0.02)
        for m in self.modules():                                                          # _init_weights, modeling_finetune.py:451-460
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        for layer_id, layer in enumerate(self.blocks):
            layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
            layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
        _trunc_normal_(self.head.weight, std=0.02)
        self.head.weight.data.mul_(init_scale)
        self.head.bias.data.mul_(init_scale)
        self.cfg = VitConfig(img_size=self.patch_embed.img_size[0], patch_size=self.patch_embed.patch_size[0], in_chans=in_chans,
                             embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, num_classes=num_classes,
                             ln_eps=self.fc_norm.eps, kind="finetune", drop_path_rate=drop_path_rate, attn_drop_rate=attn_drop_rate,
                             has_gamma=self.blocks[0].gamma_1 is not None, use_abs_pos_emb=use_abs_pos_emb)
        self._finish_init()

    def get_classifier(self):
        return self.head

    def forward_features(self, x, bool_masked_pos=None):
        self._check_input(x)
        with torch.no_grad():
            return self._run(x, None, None, "features")

    def forward(self, x, bool_masked_pos=None):
        self._check_input(x)
        return self._run(x, None, None, "logits")


# ------------------------------------------------------------------------------------------------------------------
# registry (timm's when present, else a built-in one with the same create_model call shape)
# ------------------------------------------------------------------------------------------------------------------
_MODEL_REGISTRY: Dict[str, callable] = {}


def register_model(fn):
    _MODEL_REGISTRY[fn.__name__] = fn
    try:  # pragma: no cover - timm is absent in the build image
        from timm.models.registry import register_model as timm_register
        return timm_register(fn)
    except Exception:
        return fn


def create_model(model_name, pretrained=False, **kwargs):
    """timm.models.create_model call shape (run_cyclical.py:289-302): drops None kwargs, injects pretrained_cfg*."""
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    kwargs.setdefault("pretrained_cfg", None)
    kwargs.setdefault("pretrained_cfg_overlay", None)
    if model_name not in _MODEL_REGISTRY:
        raise KeyError(f"unknown model {model_name!r}; registered: {sorted(_MODEL_REGISTRY)}")
    return _MODEL_REGISTRY[model_name](pretrained=pretrained, **kwargs)


def _cfg(url="", **kwargs):
    return {"url": url, "num_classes": 1000, "input_size": (3, 224, 224), "pool_size": None, "crop_pct": 0.9, "interpolation": "bicubic",
            "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5), **kwargs}


_FINETUNE_ONLY = ("num_classes", "use_mean_pooling", "init_scale", "linear_classifier", "has_masking", "learn_layer_weights",
                  "layernorm_before_combine", "het_layer", "sngp", "drop_block_rate")


def _build(pretrained, kwargs, **arch):
    for k in ("pretrained_cfg", "pretrained_cfg_overlay"):
        kwargs.pop(k, None)
    stochastic = kwargs.pop("stochastic", False)
    finetune = any(k in kwargs for k in ("use_mean_pooling", "init_scale", "linear_classifier")) or kwargs.pop("finetune", False)
    if stochastic:
        from . import modeling_dist
        cls = modeling_dist.DistVisionTransformer if finetune else modeling_dist.DistVisionTransformerForCyclicalTraining
    else:
        cls = VisionTransformer if finetune else VisionTransformerForCyclicalTraining
    if not finetune:
        for k in _FINETUNE_ONLY:
            kwargs.pop(k, None)
    kwargs.pop("drop_block_rate", None)
    model = cls(patch_size=16, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **arch, **kwargs)
    model.default_cfg = _cfg()
    if pretrained:
        ckpt = torch.load(kwargs["init_ckpt"], map_location="cpu")
        model.load_state_dict(ckpt["model"])
    return model


@register_model
def beit_base_patch16_224(pretrained=False, **kwargs):
    """modeling_cyclical.py:282-300 / modeling_finetune.py:1221-1229. Fine-tune kwargs (use_mean_pooling / init_scale / ...) select
    the classifier, otherwise the data2vec network; stochastic=True routes to the dual-stream classes."""
    return _build(pretrained, kwargs, embed_dim=768, depth=12, num_heads=12)


@register_model
def beit_large_patch16_224(pretrained=False, **kwargs):
    """modeling_cyclical.py:326-343 / modeling_finetune.py:1250-1257."""
    return _build(pretrained, kwargs, embed_dim=1024, depth=24, num_heads=16)


@register_model
def dist_beit_base_patch16_224(pretrained=False, **kwargs):
    """modeling_cyclical.py:303-323 / modeling_finetune.py:1231-1239."""
    kwargs["stochastic"] = True
    return _build(pretrained, kwargs, embed_dim=768, depth=12, num_heads=12)


@register_model
def dist_beit_large_patch16_224(pretrained=False, **kwargs):
    kwargs["stochastic"] = True
    return _build(pretrained, kwargs, embed_dim=1024, depth=24, num_heads=16)
