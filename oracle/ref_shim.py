"""Import shim for the REAL reference classes — TEST INFRASTRUCTURE ONLY, build container only.

`/root/reference` (fx-erick/uncertainty-vit) imports packages that are not in this image (timm, torchmetrics,
tensorboardX, deepspeed, dall_e, imageio) and two modules that are not in its own tree. This file installs minimal
stand-ins in `sys.modules` so that `modeling_finetune`, `modeling_cyclical`, `modeling_finetune_dist`,
`modeling_cyclical_dist`, `distloss`, `uncertainty_evaluations` import and run on CPU. It is used by
tools/make_golden.py to pin oracle/vit_oracle.py; nothing on the GPU box may call it (the reference does not travel).

The stand-ins restate timm's published behaviour:
  drop_path(x, p, training): per-sample mask floor(keep + U[0,1)) scaled by 1/keep  (timm.models.layers.drop)
  trunc_normal_: torch.nn.init.trunc_normal_
  ModelEmaV2: deepcopy + eval, _update zips state_dict values (timm.utils.model_ema)
"""
from __future__ import annotations

import copy
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("B200VIT_REFERENCE", "/root/reference")

# injected randomness: the golden generator sets these hooks
DROP_PATH_HOOK = None     # callable(x, p) -> keep-mask tensor [B]  (else torch.rand as timm)
DROPOUT_HOOK = None       # callable(x, p) -> keep-mask tensor like x (else F.dropout)

_REGISTRY = {}


def _drop_path(x, drop_prob: float = 0.0, training: bool = False, scale_by_keep: bool = True):
    if drop_prob == 0.0 or drop_prob is None or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    if DROP_PATH_HOOK is not None:
        mask = DROP_PATH_HOOK(x, drop_prob).to(x.dtype).reshape(shape)
    else:
        mask = torch.floor(keep_prob + torch.rand(shape, dtype=x.dtype, device=x.device))
    return x.div(keep_prob) * mask


class _InjectedDropout(nn.Module):
    """Stands in for nn.Dropout inside the reference modules (class name must start with 'Dropout',
    uncertainty_evaluations.py:35-39)."""

    def __init__(self, p=0.0, inplace=False):
        super().__init__()
        self.p = p

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        if DROPOUT_HOOK is not None:
            return x * DROPOUT_HOOK(x, self.p).to(x.dtype) / (1.0 - self.p)
        return torch.nn.functional.dropout(x, self.p, True)


_InjectedDropout.__name__ = "Dropout"


def _register_model(fn):
    _REGISTRY[fn.__name__] = fn      # last import wins, like timm's registry
    return fn


def _create_model(name, pretrained=False, **kwargs):
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    kwargs.setdefault("pretrained_cfg", None)
    kwargs.setdefault("pretrained_cfg_overlay", None)
    return _REGISTRY[name](pretrained=pretrained, **kwargs)


class _ModelEmaV2(nn.Module):
    def __init__(self, model, decay=0.9999, device=None):
        super().__init__()
        self.module = copy.deepcopy(model)
        self.module.eval()
        self.decay = decay

    def _update(self, model, update_fn):
        with torch.no_grad():
            for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
                e.copy_(update_fn(e, m))


def _accuracy(output, target, topk=(1,)):
    maxk = min(max(topk), output.size(1))
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[:min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / target.size(0) for k in topk]


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_b200vit_shim", False):
        return
    if not hasattr(np, "int"):
        np.int = int  # masking_generator.py:80 uses the removed alias
    to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    dummy = lambda *a, **k: None

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    timm = _mod("timm", _b200vit_shim=True)
    _mod("timm.models", create_model=_create_model)
    _mod("timm.models.layers", drop_path=_drop_path, to_2tuple=to_2tuple, trunc_normal_=nn.init.trunc_normal_)
    _mod("timm.models.registry", register_model=_register_model)
    _mod("timm.utils", accuracy=_accuracy, ModelEma=_ModelEmaV2, ModelEmaV2=_ModelEmaV2, get_state_dict=dummy)
    _mod("timm.data", Mixup=_Dummy, create_transform=dummy)
    _mod("timm.data.mixup", Mixup=_Dummy)
    _mod("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225),
         IMAGENET_INCEPTION_MEAN=(0.5, 0.5, 0.5), IMAGENET_INCEPTION_STD=(0.5, 0.5, 0.5))
    _mod("timm.loss", LabelSmoothingCrossEntropy=_Dummy, SoftTargetCrossEntropy=_Dummy)
    _mod("timm.optim")
    for sub, cls in (("adafactor", "Adafactor"), ("adahessian", "Adahessian"), ("adamp", "AdamP"), ("lookahead", "Lookahead"),
                     ("nadam", "Nadam"), ("nvnovograd", "NvNovoGrad"), ("radam", "RAdam"), ("rmsprop_tf", "RMSpropTF"),
                     ("sgdp", "SGDP")):
        _mod("timm.optim." + sub, **{cls: _Dummy})
    _mod("dall_e", load_model=dummy)
    _mod("dall_e.utils", map_pixels=dummy)
    _mod("tensorboardX", SummaryWriter=_Dummy)
    _mod("imageio", imread=dummy)
    _mod("torchmetrics", AUROC=_Dummy)
    _mod("cifar_semi")
    _mod("visualize_embeddings", visualize_embedding=dummy)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference builds its dropouts as nn.Dropout at construction time -> swap the class while importing/building
    nn.Dropout = _InjectedDropout


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_ROOT) and os.path.isfile(os.path.join(REFERENCE_ROOT, "modeling_finetune.py"))
