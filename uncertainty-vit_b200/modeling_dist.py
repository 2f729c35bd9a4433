"""Dual-stream ("--stochastic") models: DistVisionTransformerForCyclicalTraining (modeling_cyclical_dist.py:14-165) and
DistVisionTransformer (modeling_finetune_dist.py:181-326) with the reference's parameter names, over the CUDA schedules of core.py.
Every forward returns the (mean, cov[, logits]) tuples the reference engines unpack (engine_for_cyclical.py:70,126;
engine_for_finetuning_dist.py:288)."""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn

from . import core
from ._lib import B200VitError
from .core import VitConfig
from .modeling import (Block, PatchEmbed, RelativePositionBias, _reject_unsupported, _rows_from_mask, _trunc_normal_, _VitBase)


def _dist_common_init(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, drop_rate, attn_drop_rate, drop_path_rate,
                      norm_layer, init_values, use_shared_rel_pos_bias, masked: bool):
    self.num_features = self.embed_dim = embed_dim
    self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
    self.cov_patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
    self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
    self.cov_cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
    if masked:
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.cov_mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
    self.pos_drop = nn.Dropout(p=drop_rate)
    self.cov_pos_drop = nn.Dropout(p=drop_rate)
    self.rel_pos_bias = RelativePositionBias(self.patch_embed.patch_shape, num_heads) if use_shared_rel_pos_bias else None
    dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
    self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, drop_rate, attn_drop_rate, dpr[i], init_values, norm_layer, dist=True)
                                 for i in range(depth)])


class _DistFunction(torch.autograd.Function):
    """Whole dual-stream network as one autograd node: outputs (mean, cov); backward = core.dist_backward."""

    @staticmethod
    def forward(ctx, model, images, mask_u8, row_index, mode, noise, names, *params):
        (om, oc), saved = core.dist_forward(model._ps, model.cfg, images, mask_u8=mask_u8, row_index=row_index, mode=mode, train=model.training,
                                            save=True, noise=noise)
        ctx.model, ctx.saved, ctx.names, ctx.shapes = model, saved, names, [p.shape for p in params]
        return om, oc

    @staticmethod
    def backward(ctx, dom, doc):
        model = ctx.model
        dev = dom.device
        grads = {n: torch.zeros(s, dtype=torch.float32, device=dev) for n, s in zip(ctx.names, ctx.shapes)}
        core.dist_backward(model._ps, model.cfg, ctx.saved, dom.contiguous(), doc.contiguous(), grads)
        ctx.saved = None
        unused = model._unused_param_names()
        return (None,) * 7 + tuple(None if n in unused else grads[n] for n in ctx.names)


class _DistLogitsFunction(torch.autograd.Function):
    """Dual-stream classifier: outputs (mean_feat, cov_feat, logits) — the triple train_class_batch unpacks
    (engine_for_finetuning_dist.py:288); all three receive gradients (CE on logits + WassersteinLossFineTuning on the features)."""

    @staticmethod
    def forward(ctx, model, images, noise, names, *params):
        (fm, fc, logits), saved = core.dist_forward(model._ps, model.cfg, images, mode="logits", train=model.training, save=True, noise=noise)
        ctx.model, ctx.saved, ctx.names, ctx.shapes = model, saved, names, [p.shape for p in params]
        return fm.clone(), fc.clone(), logits.clone()

    @staticmethod
    def backward(ctx, dfm, dfc, dlogits):
        model = ctx.model
        dev = dlogits.device if dlogits is not None else dfm.device
        grads = {n: torch.zeros(s, dtype=torch.float32, device=dev) for n, s in zip(ctx.names, ctx.shapes)}
        if dlogits is None:
            dlogits = torch.zeros((ctx.saved["B"], model.cfg.num_classes), dtype=torch.float32, device=dev)
        core.dist_backward_logits(model._ps, model.cfg, ctx.saved, dfm, dfc, dlogits.float().contiguous(), grads)
        ctx.saved = None
        unused = model._unused_param_names()
        return (None,) * 4 + tuple(None if n in unused else grads[n] for n in ctx.names)


class _DistBase(_VitBase):
    def _unused_param_names(self):
        # cov_qkv.weight is allocated, saved and EMA-ed but never used: the cov stream multiplies by qkv.weight (§A.2-1)
        return {n for n, _ in self.named_parameters() if n.endswith("attn.cov_qkv.weight")}

    def _run_dist(self, x, mask_u8, row_index, mode, collect=None):
        noise = self._noise()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            named = list(self.named_parameters())
            if mode == "logits":
                return _DistLogitsFunction.apply(self, x, noise, [n for n, _ in named], *[p for _, p in named])
            if mode not in ("masked", "all"):
                raise NotImplementedError(f"no backward for mode {mode!r}")
            return _DistFunction.apply(self, x, mask_u8, row_index, mode, noise, [n for n, _ in named], *[p for _, p in named])
        out, _ = core.dist_forward(self._ps, self.cfg, x, mask_u8=mask_u8, row_index=row_index, mode=mode, train=self.training, save=False,
                                   noise=noise, collect=collect)
        return out


class DistVisionTransformerForCyclicalTraining(_DistBase):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=None, init_values=None, attn_head_dim=None, use_abs_pos_emb=True,
                 use_rel_pos_bias=False, use_shared_rel_pos_bias=False, init_std=0.02, gp_layer=False, gumbel_softmax=False, sinkformer=False,
                 h_sto_trans=False, **unused):
        super().__init__()
        _reject_unsupported(gp_layer=gp_layer, gumbel_softmax=gumbel_softmax, sinkformer=sinkformer, h_sto_trans=h_sto_trans,
                            use_rel_pos_bias=use_rel_pos_bias, drop_rate=drop_rate > 0, attn_head_dim=attn_head_dim is not None,
                            qk_scale=qk_scale is not None, no_qkv_bias=not qkv_bias, no_shared_rel_pos_bias=not use_shared_rel_pos_bias)
        if embed_dim // num_heads != 64:
            raise NotImplementedError("head_dim must be 64 (ViT-B/16, ViT-L/16)")
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        _dist_common_init(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, drop_rate, attn_drop_rate, drop_path_rate,
                          norm_layer, init_values, use_shared_rel_pos_bias, masked=True)
        self.pos_embed = None                       # the dual-stream model has no positional embedding at all (§A.2-4)
        self.norm = norm_layer(embed_dim)
        self.init_std = init_std
        self.lm_head = nn.Linear(embed_dim, embed_dim)
        self.cov_lm_head = nn.Linear(embed_dim, embed_dim)
        # modeling_cyclical_dist.py:65-71,84-93: timm trunc_normal_(std=.02) with its default ABSOLUTE bounds +-2 (effectively untruncated,
        # unlike the deterministic cyclical model's +-1 sigma), applied to nn.Linear / LayerNorm only: both patch-embedding convolutions
        # keep nn.Conv2d's default init; mask tokens stay zero (:37-38)
        _trunc_normal_(self.cls_token, std=0.02)
        _trunc_normal_(self.cov_cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        for layer_id, layer in enumerate(self.blocks):
            for w in (layer.attn.proj.weight, layer.attn.cov_proj.weight, layer.mlp.fc2.weight):
                w.data.div_(math.sqrt(2.0 * (layer_id + 1)))
        self.cfg = VitConfig(img_size=self.patch_embed.img_size[0], patch_size=self.patch_embed.patch_size[0], in_chans=in_chans, embed_dim=embed_dim,
                             depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, ln_eps=self.norm.eps, kind="cyclical", dist=True,
                             drop_path_rate=drop_path_rate, attn_drop_rate=attn_drop_rate, has_gamma=self.blocks[0].gamma_1 is not None)
        self._finish_init()

    def forward(self, x, bool_masked_pos, return_all_tokens=False, layer_results=None):
        self._check_input(x)
        B, T = x.shape[0], self.cfg.tokens
        mask_u8 = mflat = None
        if bool_masked_pos is not None:
            mflat = bool_masked_pos.reshape(B, -1).to(x.device) != 0
            mask_u8 = mflat.to(torch.uint8).reshape(-1).contiguous()
        if layer_results:
            if layer_results != "end":              # only 'end' is collected by the dual-stream reference (:139-142)
                return [], []
            with torch.no_grad():
                lm, lc = self._run_dist(x, mask_u8, None, "layers", collect=list(range(self.cfg.depth)))
            return [lm[i][:, 1:] for i in range(self.cfg.depth)], [lc[i][:, 1:] for i in range(self.cfg.depth)]
        if return_all_tokens:
            return self._run_dist(x, mask_u8, None, "all")
        return self._run_dist(x, mask_u8, _rows_from_mask(mflat.reshape(-1), T), "masked")


class DistVisionTransformer(_DistBase):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm, init_values=None,
                 use_abs_pos_emb=True, use_rel_pos_bias=False, use_shared_rel_pos_bias=False, use_mean_pooling=True, init_scale=0.001,
                 linear_classifier=False, has_masking=False, learn_layer_weights=False, layernorm_before_combine=False, gp_layer=False,
                 het_layer=False, sinkformer=False, gumbel_softmax=False, h_sto_trans=False, sngp=False, sample_head=False, **unused):
        super().__init__()
        _reject_unsupported(gp_layer=gp_layer, het_layer=het_layer, sinkformer=sinkformer, gumbel_softmax=gumbel_softmax, h_sto_trans=h_sto_trans,
                            sngp=sngp, use_rel_pos_bias=use_rel_pos_bias, learn_layer_weights=learn_layer_weights, drop_rate=drop_rate > 0,
                            no_mean_pooling=not use_mean_pooling, qk_scale=qk_scale is not None, no_qkv_bias=not qkv_bias,
                            linear_classifier=linear_classifier, no_shared_rel_pos_bias=not use_shared_rel_pos_bias)
        if embed_dim // num_heads != 64:
            raise NotImplementedError("head_dim must be 64 (ViT-B/16, ViT-L/16)")
        self.num_classes = num_classes
        _dist_common_init(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, drop_rate, attn_drop_rate, drop_path_rate,
                          norm_layer, init_values, use_shared_rel_pos_bias, masked=False)
        self.use_rel_pos_bias = use_rel_pos_bias
        self.use_mean_pooling = use_mean_pooling
        self.norm = nn.Identity()
        self.fc_norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.cov_lm_head = nn.Identity()
        _trunc_normal_(self.cls_token, std=0.02)
        _trunc_normal_(self.cov_cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        for layer_id, layer in enumerate(self.blocks):
            for w in (layer.attn.proj.weight, layer.attn.cov_proj.weight, layer.mlp.fc2.weight):
                w.data.div_(math.sqrt(2.0 * (layer_id + 1)))              # no init_scale on the head here (§A.2-5)
        self.cfg = VitConfig(img_size=self.patch_embed.img_size[0], patch_size=self.patch_embed.patch_size[0], in_chans=in_chans, embed_dim=embed_dim,
                             depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, num_classes=num_classes, ln_eps=self.fc_norm.eps, kind="finetune",
                             dist=True, drop_path_rate=drop_path_rate, attn_drop_rate=attn_drop_rate, has_gamma=self.blocks[0].gamma_1 is not None,
                             sample_head=bool(sample_head))
        self._finish_init()

    def get_classifier(self):
        return self.head

    def forward(self, x, bool_masked_pos=None):
        """Returns (mean_feat [B,C], cov_feat [B,C], logits [B,K]); bool_masked_pos is ignored, as in the reference (§A.2-4)."""
        self._check_input(x)
        return self._run_dist(x, None, None, "logits")
