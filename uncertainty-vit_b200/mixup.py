"""Device Mixup / CutMix with timm's interface (timm.data.mixup.Mixup, timm 0.3.2 — the version the reference pins; built at
run_class_finetuning.py:339-347, applied to every fine-tune batch at engine_for_finetuning.py:87-88).

timm is not part of this image, so the parameter logic below restates its published algorithm: per batch, with probability `prob`, draw
lambda ~ Beta(alpha, alpha) for mixup or cutmix (cutmix with probability `switch_prob` when both are on); cutmix cuts a box of area ratio
1 - lambda around a uniform centre, clipped to the image, and corrects lambda to the clipped area. The draws use numpy's generator interface
in timm's call order (rand, rand, beta, randint, randint), so a seeded `np.random` reproduces a timm run. The image mixing and the smoothed,
mixed one-hot targets run in one C-ABI call (`b200vit_mixup_batch`), in place, with the torch ops' roundings. Only mode='batch' (the reference's
default and README recipe) is implemented."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops


def rand_bbox(img_shape, lam, margin=0.0, rng=np.random):
    """timm rand_bbox: box of area ratio (1 - lam) around a uniform centre, clipped to the image."""
    ratio = np.sqrt(1 - lam)
    img_h, img_w = img_shape[-2:]
    cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
    margin_y, margin_x = int(margin * cut_h), int(margin * cut_w)
    cy = rng.randint(0 + margin_y, img_h - margin_y)
    cx = rng.randint(0 + margin_x, img_w - margin_x)
    yl = np.clip(cy - cut_h // 2, 0, img_h)
    yh = np.clip(cy + cut_h // 2, 0, img_h)
    xl = np.clip(cx - cut_w // 2, 0, img_w)
    xh = np.clip(cx + cut_w // 2, 0, img_w)
    return int(yl), int(yh), int(xl), int(xh)


def rand_bbox_minmax(img_shape, minmax, rng=np.random):
    """timm rand_bbox_minmax: side lengths uniform in [min, max) of the image side, corner uniform."""
    assert len(minmax) == 2
    img_h, img_w = img_shape[-2:]
    cut_h = rng.randint(int(img_h * minmax[0]), int(img_h * minmax[1]))
    cut_w = rng.randint(int(img_w * minmax[0]), int(img_w * minmax[1]))
    yl = rng.randint(0, img_h - cut_h)
    xl = rng.randint(0, img_w - cut_w)
    return int(yl), int(yl + cut_h), int(xl), int(xl + cut_w)


def cutmix_bbox_and_lam(img_shape, lam, ratio_minmax=None, correct_lam=True, rng=np.random):
    if ratio_minmax is not None:
        yl, yu, xl, xu = rand_bbox_minmax(img_shape, ratio_minmax, rng=rng)
    else:
        yl, yu, xl, xu = rand_bbox(img_shape, lam, rng=rng)
    if correct_lam or ratio_minmax is not None:
        bbox_area = (yu - yl) * (xu - xl)
        lam = 1.0 - bbox_area / float(img_shape[-2] * img_shape[-1])
    return (yl, yu, xl, xu), lam


class Mixup:
    def __init__(self, mixup_alpha=1.0, cutmix_alpha=0.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode="batch", correct_lam=True,
                 label_smoothing=0.1, num_classes=1000, rng=None):
        self.mixup_alpha = mixup_alpha
        self.cutmix_alpha = cutmix_alpha
        self.cutmix_minmax = cutmix_minmax
        if self.cutmix_minmax is not None:
            assert len(self.cutmix_minmax) == 2
            self.cutmix_alpha = 1.0          # force cutmix alpha == 1.0 when minmax active to keep logic simple & safe
        self.mix_prob = prob
        self.switch_prob = switch_prob
        self.label_smoothing = label_smoothing
        self.num_classes = num_classes
        if mode != "batch":
            raise NotImplementedError(f"Mixup mode {mode!r}: only 'batch' (the reference's default) runs on the B200 path")
        self.mode = mode
        self.correct_lam = correct_lam
        self.mixup_enabled = True
        self.rng = rng if rng is not None else np.random

    def _params_per_batch(self) -> Tuple[float, bool]:
        lam, use_cutmix = 1.0, False
        if self.mixup_enabled and self.rng.rand() < self.mix_prob:
            if self.mixup_alpha > 0.0 and self.cutmix_alpha > 0.0:
                use_cutmix = self.rng.rand() < self.switch_prob
                lam_mix = self.rng.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else self.rng.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.mixup_alpha > 0.0:
                lam_mix = self.rng.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.cutmix_alpha > 0.0:
                use_cutmix = True
                lam_mix = self.rng.beta(self.cutmix_alpha, self.cutmix_alpha)
            else:
                assert False, "One of mixup_alpha > 0., cutmix_alpha > 0., cutmix_minmax not None should be true."
            lam = float(lam_mix)
        return lam, bool(use_cutmix)

    def draw(self, img_shape):
        """(lam, use_cutmix, box) for one batch: the host-side part of Mixup._mix_batch."""
        lam, use_cutmix = self._params_per_batch()
        box = (0, 0, 0, 0)
        if lam != 1.0 and use_cutmix:
            box, lam = cutmix_bbox_and_lam(img_shape, lam, ratio_minmax=self.cutmix_minmax, correct_lam=self.correct_lam, rng=self.rng)
        return lam, use_cutmix, box

    def __call__(self, x: torch.Tensor, target: torch.Tensor):
        """x fp32 [B,C,H,W] on the device, mixed IN PLACE (as timm does); target int64 [B] -> soft targets fp32 [B, num_classes]."""
        assert len(x) % 2 == 0, "Batch size should be even when using this"
        lam, use_cutmix, box = self.draw(x.shape)
        off_value = self.label_smoothing / self.num_classes
        on_value = 1.0 - self.label_smoothing + off_value
        soft = ops.mixup_batch(x, lam, use_cutmix, box, target.to(torch.int64).contiguous(), self.num_classes, on_value, off_value)
        return x, soft
