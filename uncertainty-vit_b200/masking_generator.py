"""Device-side mirror of the reference's block-wise `MaskingGenerator` (masking_generator.py:29-92).

Same constructor, `__repr__`, `get_shape()` and `__call__()` as the reference class, plus `batch(B)`, which draws the masks of a
whole batch in one launch of `b200vit_block_masks` (one thread per image) and also returns the masked-row list the student head gathers,
so neither the per-image Python loops of the reference nor a host `nonzero` sit in front of the step. The proposal distribution and
the accept/reject rules are the reference's; the random stream is Philox keyed on (seed, running image number) instead of Python's
global `random`, so individual masks differ from a reference run with the same seed unless the uniforms are injected (`uniforms=`)."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops


class MaskingGenerator:
    def __init__(self, input_size, num_masking_patches, min_num_patches=4, max_num_patches=None, min_aspect=0.3, max_aspect=None,
                 *, seed: int = 0, device=None):
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.height, self.width = input_size
        self.num_patches = self.height * self.width
        self.num_masking_patches = num_masking_patches
        self.min_num_patches = min_num_patches
        self.max_num_patches = num_masking_patches if max_num_patches is None else max_num_patches
        max_aspect = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(max_aspect))
        self.seed = int(seed)
        self.device = torch.device(device if device is not None else "cuda")
        self.images_drawn = 0                      # running image number: the Philox counter, so no two calls repeat a mask

    def __repr__(self):
        return "Generator(%d, %d -> [%d ~ %d], max = %d, %.3f ~ %.3f)" % (
            self.height, self.width, self.min_num_patches, self.max_num_patches, self.num_masking_patches,
            self.log_aspect_ratio[0], self.log_aspect_ratio[1])

    def get_shape(self):
        return self.height, self.width

    def batch(self, B: int, uniforms: Optional[torch.Tensor] = None, want_rows: bool = True,
              rows: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """(mask uint8 [B, H*W], count int32 [B+1] (count[B] = R), rows int32 [B*num_masking_patches], first R valid) on the device, no sync.
        rows[k] = b*(H*W+1) + 1 + patch: the flat residual-stream row of the k-th masked patch (modeling_cyclical.py:221-225 order).
        A caller-provided `rows` buffer keeps whatever it held beyond the first R entries (the engine pre-fills it with row 0)."""
        out = ops.block_masks(B, self.height, self.width, self.num_patches + 1, self.num_masking_patches, self.min_num_patches,
                              self.max_num_patches, self.log_aspect_ratio[0], self.log_aspect_ratio[1], self.seed, self.images_drawn,
                              self.device, uniforms=uniforms, want_rows=want_rows, rows=rows)
        if uniforms is None:
            self.images_drawn += B
        return out

    def state_dict(self):
        return {"seed": self.seed, "images_drawn": self.images_drawn}

    def load_state_dict(self, sd):
        self.seed, self.images_drawn = int(sd["seed"]), int(sd["images_drawn"])

    def __call__(self) -> np.ndarray:
        """One [H, W] integer mask on the host, as the reference returns it (a one-image batch; use batch() in a training loop)."""
        mask, _, _ = self.batch(1, want_rows=False)
        return mask.view(self.height, self.width).cpu().numpy().astype(np.int64)
