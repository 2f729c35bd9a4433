"""data2vec pre-training step engine (the hot loop of engine_for_cyclical.train_one_epoch, :45-186) over flat device arenas.

One step = EMA-teacher forward (eval, unmasked) -> target builder + smooth-L1 (one kernel, masked rows only) -> student
forward/backward -> [NCCL all-reduce of the flat gradient arena] -> grad-norm -> fused clip + AdamW + bf16 shadows + EMA.
No host synchronisation inside the step; the loss is read back once at the end (the reference's loss.item(), :164).

The model's nn.Parameters are re-pointed at views of the arenas, so state_dict()/checkpoints/other code see live values.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import core, ops
from ._lib import B200VitError
from .core import Noise, VitConfig

CHUNK = 1024


def get_num_layer_for_vit(var_name: str, num_max_layer: int) -> int:
    """optim_factory.py:33-44."""
    if var_name in ("cls_token", "mask_token", "pos_embed"):
        return 0
    if var_name.startswith("patch_embed"):
        return 0
    if var_name.startswith("rel_pos_bias"):
        return num_max_layer - 1
    if var_name.startswith("blocks"):
        return int(var_name.split(".")[1]) + 1
    return num_max_layer - 1


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0, warmup_steps=-1):
    """utils.py:408-425."""
    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_steps > 0:
        warmup_iters = warmup_steps
    if warmup_epochs > 0 or warmup_steps > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = np.array([final_value + 0.5 * (base_value - final_value) * (1 + math.cos(math.pi * i / (len(iters)))) for i in iters])
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


def sync_initial_parameters(arena: torch.Tensor, world_size: int, process_group=None) -> None:
    """Every rank seeds torch with seed + rank before it builds the model (run_cyclical.py:316-322), so the ranks start from DIFFERENT
    weights; the reference makes them identical through the DistributedDataParallel constructor's broadcast from rank 0 (:515-519).
    The engines replace DDP, so they do the same on the flat fp32 arena (one collective)."""
    if world_size > 1:
        torch.distributed.broadcast(arena, 0, group=process_group)


class ArenaParams(core.ParamSource):
    """ParamSource over one fp32 arena + its bf16 shadow arena."""

    def __init__(self, layout: Dict[str, tuple], f32: torch.Tensor, bf16: torch.Tensor, rel_index: torch.Tensor, num_classes_pad=None):
        self.layout, self.a32, self.a16, self._rel = layout, f32, bf16, rel_index
        self._v32: Dict[str, torch.Tensor] = {}
        self._v16: Dict[str, torch.Tensor] = {}

    def f32(self, name):
        v = self._v32.get(name)
        if v is None:
            ent = self.layout.get(name)
            if ent is None:
                return None
            off, shape = ent
            v = self.a32[off: off + int(np.prod(shape))].view(shape)
            self._v32[name] = v
        return v

    def bf16(self, name):
        v = self._v16.get(name)
        if v is None:
            off, shape = self.layout[name]
            v = self.a16[off: off + int(np.prod(shape))].view(shape[0], -1)
            self._v16[name] = v
        return v

    def qkv_bias(self, prefix, cov=False):
        return self.f32(prefix + ("attn.__cov_qkv_bias" if cov else "attn.__qkv_bias"))

    def rel_index_i32(self):
        return self._rel

    def head_padded(self):
        """(head.weight bf16 [Kp, C], head.bias fp32 [Kp]) straight from the arenas: the engine reserves Kp = K rounded up to 8 rows for
        the classifier, rows >= K are zero and stay zero (zero gradient, zero weight)."""
        (ow, (K, C)), (ob, _) = self.layout["head.weight"], self.layout["head.bias"]
        kp = (K + 7) // 8 * 8
        return self.a16[ow: ow + kp * C].view(kp, C), self.a32[ob: ob + kp]


class D2VEngine:
    """Owns arenas + optimizer state for a data2vec student and its EMA teacher."""

    def __init__(self, model, *, lr=2e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, clip_grad=3.0, ema_decay=0.9998,
                 ema_decay_init=0.999, ema_start_at=0, target_layers: Sequence[int] = (6, 7, 8, 9, 10, 11), l1_beta=2.0, l2_loss=False,
                 target_layer_norm_last=True, post_target_layer_norm=True, layer_decay: Optional[float] = None, loss_scale=-1.0,
                 skip_weight_decay: Iterable[str] = ("pos_embed", "cls_token"), world_size=1, process_group=None, seed=0,
                 lambda_pretraining: float = 1e-5, use_graph: bool = True, with_ema: bool = True, target_batch_norm=False,
                 target_instance_norm=False, post_target_instance_norm=False, var_w0: float = 0.0, var_margin0: float = 0.5,
                 start_lr_decay_at_step: int = -1, mask_dropout_prob: float = -1.0, track_z0: bool = True, overlap_allreduce: Optional[bool] = None,
                 allreduce_cut_block: Optional[int] = None, allreduce_sm_reserve: int = 0):
        """Keyword names follow engine_for_cyclical.train_one_epoch (:24-32) / run_cyclical.py's flags. world_size > 1: rank 0's parameters
        are broadcast at construction (what the DistributedDataParallel constructor does in run_cyclical.py:515-519), then only gradients
        are all-reduced."""
        self.model = model
        self.cfg: VitConfig = model.cfg
        dev = model.cls_token.device
        if dev.type != "cuda":
            raise B200VitError("D2VEngine needs the model on a CUDA (B200) device")
        self.dev = dev
        self._copy_stream = None
        self._side_stream = None
        self._teacher_stream = None
        self._teacher_forked = False
        self.use_graph = use_graph
        self.max_graphs = 2
        self.wloss_dev = None
        self._graphs = {}
        self._eager_steps = 0
        self.lr, self.wd, self.betas, self.eps, self.clip = lr, weight_decay, betas, eps, clip_grad
        self.ema_decay, self.ema_decay_init, self.ema_start_at = ema_decay, ema_decay_init, ema_start_at
        self.target_layers = list(target_layers)
        self.l1_beta, self.l2_loss, self.ln_each, self.ln_post = l1_beta, l2_loss, target_layer_norm_last, post_target_layer_norm
        self.loss_scale = loss_scale
        self.bn_targets, self.in_targets, self.post_in_targets = bool(target_batch_norm), bool(target_instance_norm), bool(post_target_instance_norm)
        self.var_w0, self.var_margin0, self.track_z0 = float(var_w0), float(var_margin0), bool(track_z0)
        self.start_lr_decay_at_step = int(start_lr_decay_at_step)
        self.mask_dropout_prob = float(mask_dropout_prob)
        self.z0_dev = self.std_loss0_dev = None
        self.lam = lambda_pretraining
        self.world_size, self.pg = world_size, process_group
        # Gradient all-reduce overlapped with the backward pass (what DDP's buckets do in the reference, run_cyclical.py:515-519): the arena is
        # laid out in named_parameters() order, so once block `cut`'s backward has run, everything from blocks.<cut> to the head is final. That
        # upper part is all-reduced on NCCL's stream while the lower blocks' backward continues (inside the captured graph too: the fork and
        # the join are graph edges). `allreduce_sm_reserve` > 0 additionally sizes the persistent kernels for that many fewer SMs during the
        # window. Measured at N = 2 (profiles/r2_allreduce_overlap_n2.md): 29.65 ms without overlap, 29.44 ms with cut = 2 / no reserve; an
        # SM reserve of 16-32 never paid (the lower blocks' GEMMs lose more than NCCL gains), so the default is 0. The gain is small because
        # NCCL's CTAs cannot co-reside with the ~200 KB-smem persistent GEMM / attention CTAs: they mostly run in the gaps.
        import os as _os
        if overlap_allreduce is None:
            overlap_allreduce = _os.environ.get("B200VIT_AR_OVERLAP", "1") != "0"
        self.overlap_ar = bool(overlap_allreduce) and world_size > 1
        self.ar_cut = allreduce_cut_block if allreduce_cut_block is not None else int(_os.environ.get("B200VIT_AR_CUT", "2"))
        self.ar_reserve = int(_os.environ.get("B200VIT_AR_SM_RESERVE", str(allreduce_sm_reserve)))
        self._ar_done_in_step = False
        self.seed = seed
        self.it = 0
        self.layer_decay, self.skip_weight_decay = layer_decay, tuple(skip_weight_decay)
        # (mean, std, hwc): set to accept uint8 pixel batches from the loader; ToTensor + Normalize (datasets.py:80-85) then run on the device
        self.pixel_norm = None
        # ---- layout: every segment starts on a CHUNK boundary; (q_bias | 0 | v_bias) of a block form ONE segment so that the
        # QKV GEMM bias cat(q_bias, zeros, v_bias) (modeling_finetune.py:148) is a plain arena view
        named = dict(model.named_parameters())
        L = self.cfg.depth + 2
        layout: Dict[str, tuple] = {}
        chunks_hp: List[tuple] = []
        off = 0
        skip = set(skip_weight_decay)

        def add(name, shape, lr_scale, wd_scale, reserve=0):
            nonlocal off
            n = max(int(np.prod(shape)), reserve)
            layout[name] = (off, tuple(shape))
            nch = (n + CHUNK - 1) // CHUNK
            chunks_hp.extend([(lr_scale, wd_scale)] * nch)
            off += nch * CHUNK

        done = set()
        for name, p in named.items():
            if name in done:
                continue
            layer_id = get_num_layer_for_vit(name, L)
            lr_scale = 1.0 if layer_decay is None else layer_decay ** (L - 1 - layer_id)
            no_decay = p.dim() == 1 or name.endswith(".bias") or name in skip            # optim_factory.py:66-67
            if name.endswith("attn.q_bias") or name.endswith("attn.cov_q_bias"):
                cov = "cov_" if name.endswith("cov_q_bias") else ""
                prefix = name[: -len(cov + "q_bias")]
                C = p.numel()
                add(prefix + f"__{cov}qkv_bias", (3 * C,), lr_scale, 0.0)
                base = layout[prefix + f"__{cov}qkv_bias"][0]
                layout[prefix + cov + "q_bias"] = (base, (C,))
                layout[prefix + cov + "v_bias"] = (base + 2 * C, (C,))
                done.update({prefix + cov + "q_bias", prefix + cov + "v_bias"})
                continue
            if name.endswith("attn.cov_qkv.weight"):
                # never receives a gradient in the reference (the cov stream multiplies by qkv.weight, §A.2-1): torch's AdamW skips
                # grad-less parameters entirely (no update, no weight decay) -> lr_scale = wd_scale = 0
                add(name, p.shape, 0.0, 0.0)
                continue
            reserve = 0
            if name in ("head.weight", "head.bias"):      # classifier rows padded to a multiple of 8 (GEMM N); the padding stays zero
                kp = (p.shape[0] + 7) // 8 * 8
                reserve = kp * (p.numel() // p.shape[0])
            add(name, p.shape, lr_scale, 0.0 if no_decay else 1.0, reserve)
        self.layout, self.n = layout, off
        f32 = lambda: torch.zeros(off, dtype=torch.float32, device=dev)
        self.p32, self.g32, self.m32, self.v32 = f32(), f32(), f32(), f32()
        self.e32 = f32() if with_ema else None
        self.p16 = torch.zeros(off, dtype=torch.bfloat16, device=dev)
        self.e16 = torch.zeros(off, dtype=torch.bfloat16, device=dev) if with_ema else None
        self.hp = torch.tensor(chunks_hp, dtype=torch.float32, device=dev).contiguous()
        with torch.no_grad():
            for name, p in named.items():
                o, shape = layout[name]
                view = self.p32[o: o + p.numel()].view(shape)
                view.copy_(p.detach())
                p.data = view                     # the module now aliases the arena
        sync_initial_parameters(self.p32, world_size, process_group)      # before the bf16 shadows and the EMA copy are derived
        ops.cast_bf16(self.p32, self.p16)
        if with_ema:
            self.e32.copy_(self.p32)              # ModelEmaV2: deepcopy of the student at construction (run_cyclical.py:503)
            ops.cast_bf16(self.e32, self.e16)
        rel = model.rel_pos_bias.relative_position_index.to(torch.int32).contiguous() if model.rel_pos_bias is not None else None
        self.student = ArenaParams(layout, self.p32, self.p16, rel)
        # the teacher owns a COPY of the integer index buffer: ModelEmaV2._update walks the whole state dict, so the reference's EMA passes
        # relative_position_index through `d*e + (1-d)*m` in fp32 and truncates it back (b200vit_ema_index_update) — entries can drift
        self.rel_e = rel.clone() if (with_ema and rel is not None) else None
        self.teacher = ArenaParams(layout, self.e32, self.e16, self.rel_e) if with_ema else None
        self.grads = {name: self.g32[o: o + int(np.prod(s))].view(s) for name, (o, s) in layout.items()}
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.opt_step = 0
        self.cur_decay = ema_decay          # the reference's loop variable (engine_for_cyclical.py:44): train_one_epoch() resets it per epoch
        self.ema_updated = True
        if hasattr(model, "_ps"):
            model._ps.invalidate()

    # ------------------------------------------------------------------------------------------------------------
    def state_dict(self, epoch: Optional[int] = None) -> Dict[str, object]:
        """Checkpoint in the layout utils.save_model writes (utils.py:462-485): see checkpoint.engine_state_dict."""
        from . import checkpoint
        return checkpoint.engine_state_dict(self, epoch)

    def load_state_dict(self, ckpt: Dict[str, object], load_optimizer: bool = True) -> None:
        from . import checkpoint
        checkpoint.load_engine_state_dict(self, ckpt, load_optimizer=load_optimizer)

    def ema_state_dict(self) -> Dict[str, torch.Tensor]:
        return {name: self.e32[o: o + int(np.prod(s))].view(s) for name, (o, s) in self.layout.items() if "__" not in name}

    @staticmethod
    def rows_from_host_mask(mask: np.ndarray, T: int) -> np.ndarray:
        """mask: [B, np] {0,1} on the HOST -> flat stream rows b*T+1+p in the reference's boolean-gather order."""
        b, p = np.nonzero(mask.reshape(mask.shape[0], -1))
        return (b * T + 1 + p).astype(np.int32)

    def decay_at(self, it: int) -> float:
        """engine_for_cyclical.py:55-56: the annealed value while it < ema_start_at; afterwards the loop variable keeps whatever it held
        (the last annealed value for the rest of that epoch, `decay` from the next epoch on, 0 after a skipped update, :182-185)."""
        if it < self.ema_start_at:
            return self.ema_decay_init + it * (self.ema_decay - self.ema_decay_init) / self.ema_start_at
        return self.cur_decay

    def step(self, images: torch.Tensor, mask_u8: torch.Tensor, rows: torch.Tensor, *, lr: Optional[float] = None,
             weight_decay: Optional[float] = None, noise: Optional[Noise] = None, graph: Optional[bool] = None,
             n_valid: Optional[torch.Tensor] = None, mask_keep: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on device-resident inputs. images fp32 [B,3,H,W]; mask_u8 uint8 [B*np]; rows int32 [R].
        n_valid (device int32 [1]): `rows` is padded to a fixed capacity and only its first n_valid entries are masked patches (the
        padding must hold valid row numbers, e.g. 0); the step then has the same launch shapes whatever the block-wise generator drew.
        Returns the device scalar loss (no sync)."""
        cfg = self.cfg
        B = images.shape[0]
        C, T = cfg.embed_dim, cfg.tokens
        R = rows.numel()
        lr = self.lr if lr is None else lr
        wd = self.wd if weight_decay is None else weight_decay
        self.cur_decay = self.decay_at(self.it)
        if self.mask_dropout_prob > 0:
            mask_u8, rows, n_valid = self.apply_mask_dropout(mask_u8, B, mask_keep)
            R = rows.numel()
        injected = noise
        if noise is None:
            noise = Noise(seed=(self.seed * 0x9E3779B97F4A7C15 + self.it + 1) & 0xFFFFFFFFFFFFFFFF)
        if graph is None:
            graph = self.use_graph
        if R == 0:
            raise B200VitError("D2VEngine.step: empty masked-row list (the loss is a mean over the masked patches)")
        if graph and injected is None and ops.GEMM_TIMING is None:
            self._fwd_bwd_graphed(images, mask_u8, rows, noise.seed, n_valid)
        else:
            self._fwd_bwd(images, mask_u8, rows, noise, n_valid)
            self._eager_steps += 1
        return self._optimizer_step(lr, wd)

    def apply_mask_dropout(self, mask_u8: torch.Tensor, B: int, keep: Optional[torch.Tensor] = None):
        """mask_dropout_prob (engine_for_cyclical.py:62-66): mask &= bernoulli(1 - p) on the device (or an injected keep tensor), the masked-row
        list rebuilt at capacity (padding = row 0) with the true count in device memory. Returns (mask, rows, n_valid)."""
        npat = self.cfg.num_patches
        m = mask_u8.clone()
        seed = (self.seed * 0x9E3779B97F4A7C15 + 0x51ED270B * (self.it + 1)) & 0xFFFFFFFFFFFFFFFF
        cap = int(B * npat)
        count, rows = ops.mask_dropout(m, B, npat, self.cfg.tokens, self.mask_dropout_prob, seed=seed, first_image=self.it * B,
                                       keep_in=keep.reshape(-1).to(torch.uint8).contiguous() if keep is not None else None,
                                       rows=torch.zeros(cap, dtype=torch.int32, device=self.dev))
        return m, rows, count[B:B + 1]

    def _targets_and_loss(self, layers, rows, y, R, n_valid, *, targets=None, dy_bf16=None, dy_f32=None, loss_mult=1.0, with_loss=True,
                          mean_stream=True):
        """Target builder + loss of engine_for_cyclical.py:90-150 for one stream. `layers`: the teacher's block outputs [B*T, C] fp32.
        mean_stream=False (the cov targets of --stochastic, :74-86) only ever gets the per-layer LayerNorm, the mean and the post LayerNorm."""
        cfg = self.cfg
        C, T = cfg.embed_dim, cfg.tokens
        B = layers[0].shape[0] // T
        dev = self.dev
        ls = self.loss_scale if self.loss_scale != -1 else 1.0
        bn, inorm, post_in = (self.bn_targets, self.in_targets, self.post_in_targets) if mean_stream else (False, False, False)
        affine = None
        if bn or inorm:
            st = ops.channel_stats(layers, C, B, T, 1, T - 1, C, bn, inorm)
            affine = [st[i] for i in range(len(layers))]
        col_hinge = loss_add = None
        row_loss = torch.empty((R,), dtype=torch.float32, device=dev) if with_loss else None
        if with_loss and y is not None and (self.track_z0 or self.var_w0 > 0):
            if self.z0_dev is None:
                self.z0_dev = torch.zeros(C, dtype=torch.float32, device=dev)
                self.std_loss0_dev = torch.zeros(1, dtype=torch.float32, device=dev)
            _, _, col_hinge = ops.column_std(y, R, C, n_valid, 1e-6, self.var_margin0, self.var_w0 * ls, want_hinge_grad=self.var_w0 > 0,
                                             z0=self.z0_dev, hinge=self.std_loss0_dev)
            if self.var_w0 > 0:
                loss_add = self.std_loss0_dev
        common = dict(beta=self.l1_beta, l2_loss=self.l2_loss if mean_stream else False, n_valid=n_valid)
        if not post_in:
            ops.d2v_target_loss_ex(layers, C, rows, y, R, C, self.ln_each, self.ln_post, grad_scale=ls / (R * C) if with_loss else 1.0,
                                   targets=targets, dy_bf16=dy_bf16, dy_f32=dy_f32, row_loss=row_loss, loss_out=self.loss_dev if with_loss else None,
                                   affine=affine, rows_per_sample=T if affine is not None else 0, col_hinge=col_hinge, loss_add=loss_add,
                                   loss_add_weight=self.var_w0, loss_mult=loss_mult, **common)
            return
        # post_target_instance_norm (:112-115) normalises the AVERAGED target over the patch tokens of each (image, channel): it needs
        # the target of every patch row first -> pass 1 over all B*(T-1) rows, statistics, pass 2 over the masked rows of that tensor
        NP = B * (T - 1)
        full = torch.empty((NP, C), dtype=torch.float32, device=dev)
        ops.d2v_target_loss_ex(layers, C, core.all_patch_rows(B, T, dev), None, NP, C, self.ln_each, False, targets=full, affine=affine,
                               rows_per_sample=T if affine is not None else 0)
        st2 = ops.channel_stats([full], C, B, T - 1, 0, T - 1, C, False, True)
        ops.d2v_target_loss_ex([full], C, rows, y, R, C, False, self.ln_post, grad_scale=ls / (R * C) if with_loss else 1.0, targets=targets,
                               dy_bf16=dy_bf16, dy_f32=dy_f32, row_loss=row_loss, loss_out=self.loss_dev if with_loss else None, affine=[st2[0]],
                               rows_per_sample=T - 1, compact_tokens=T, col_hinge=col_hinge, loss_add=loss_add, loss_add_weight=self.var_w0,
                               loss_mult=loss_mult, **common)

    def _draw_masks(self, B, noise):
        """The student's attention-dropout masks of this step, drawn on a side stream while the teacher forward runs (core.draw_keep_bits)."""
        import os as _os
        if _os.environ.get("B200VIT_SIDE_MASKS", "1") == "0":       # A/B switch: the attention forwards then draw their own masks in-stream
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.dev)
        core.draw_keep_bits(self.cfg, B, noise, self.dev, self._side_stream)

    def _teacher_stream_ctx(self):
        """Context that issues the teacher forward on a second stream (B200VIT_TEACHER_STREAM=0: on the step's own stream)."""
        import contextlib
        if os.environ.get("B200VIT_TEACHER_STREAM", "1") == "0":
            self._teacher_forked = False
            return contextlib.nullcontext()
        if self._teacher_stream is None:
            self._teacher_stream = torch.cuda.Stream(device=self.dev)
        self._teacher_stream.wait_stream(torch.cuda.current_stream(self.dev))
        self._teacher_forked = True
        return torch.cuda.stream(self._teacher_stream)

    def _teacher_join(self):
        if self._teacher_forked:
            torch.cuda.current_stream(self.dev).wait_stream(self._teacher_stream)
            self._teacher_forked = False

    def _overlapped_backward(self, run_backward):
        """Runs `run_backward(after_block)` with the gradient all-reduce of the arena's upper part (blocks.<cut> .. head) issued as soon as it is
        final, on NCCL's stream, and the lower part right after the backward; see __init__. Without data parallelism it is a plain call."""
        if not self.overlap_ar:
            run_backward(None)
            return
        cut = min(max(self.ar_cut, 1), self.cfg.depth - 1)
        off = self.layout[f"blocks.{cut}.gamma_1"][0] if f"blocks.{cut}.gamma_1" in self.layout else self.layout[f"blocks.{cut}.norm1.weight"][0]
        upper, lower = self.g32[off:], self.g32[:off]
        state = {}

        def after_block(i):
            if i == cut:
                state["work"] = torch.distributed.all_reduce(upper, group=self.pg, async_op=True)
                if self.ar_reserve > 0:
                    state["prev"] = ops.set_sm_limit(max(2, ops.sm_count() - self.ar_reserve))
        try:
            run_backward(after_block)
        finally:
            if "prev" in state:
                ops.set_sm_limit(state["prev"])
        state["work"].wait()
        torch.distributed.all_reduce(lower, group=self.pg)
        self._ar_done_in_step = True

    def _fwd_bwd(self, images, mask_u8, rows, noise, n_valid=None):
        """Teacher forward, student forward, targets + loss, student backward into the gradient arena (everything but the optimiser).
        With n_valid the padded rows of `rows` get dy = 0 from the loss kernel, so they add nothing to any gradient."""
        cfg = self.cfg
        if cfg.dist:
            return self._fwd_bwd_dist(images, mask_u8, rows, noise, n_valid)
        B = images.shape[0]
        C, T = cfg.embed_dim, cfg.tokens
        R = rows.numel()
        patches = core.patches_bf16(cfg, images)
        self._draw_masks(B, noise)
        # teacher (EMA weights, eval mode, unmasked): engine_for_cyclical.py:68-88. It shares nothing with the student forward but the
        # patches, so it runs on its own stream: its kernels fill the SMs the student's persistent GEMMs leave idle in their last wave.
        with self._teacher_stream_ctx():
            self.g32.zero_()          # 345 MB of fp32 gradients: cleared beside the forward instead of in front of the backward
            layers, _ = core.vit_forward(self.teacher, cfg, images, mode="layers", train=False, save=False, collect=self.target_layers,
                                         patches=patches)
        # student: :124-128
        out, ctx = core.vit_forward(self.student, cfg, images, mask_u8=mask_u8, row_index=rows, mode="masked", train=True, save=True, noise=noise,
                                    patches=patches)
        self._teacher_join()
        # targets + loss + dLoss/dy: :90-150 (masked rows only; LayerNorm is per row)
        dy = torch.empty((R, C), dtype=torch.bfloat16, device=self.dev)
        ls = self.loss_scale if self.loss_scale != -1 else 1.0
        self._targets_and_loss([layers[i].view(B * T, C) for i in self.target_layers], rows, out, R, n_valid, dy_bf16=dy, loss_mult=ls)
        del layers
        self._overlapped_backward(lambda cb: core.vit_backward(self.student, cfg, ctx, dy, self.grads, after_block=cb))

    def _fwd_bwd_graphed(self, images, mask_u8, rows, seed: int, n_valid=None):
        """The same launch sequence replayed from a CUDA graph: ~400 stream-ordered, allocation-free launches whose Python issue time
        (~55 us each) exceeds the run time of the small row kernels. Inputs are copied into static buffers; the per-step randomness
        enters through device memory (drop-path factors, the Philox key of attention dropout), so the captured graph stays valid across
        steps. The first two steps of a shape run eagerly (lazy kernel attributes, allocator warm-up).

        The masked-row count of block-wise masking changes from batch to batch (masking_generator.py:80-92 stops a few patches short of
        the target in a third of the images), so the graph is captured for a row CAPACITY — the count rounded up to a multiple of
        4 x batch, or the length of a caller-padded list — and the true count travels in device memory (`n_valid`): the loss kernel
        zeroes dy of the padding rows, which therefore contribute nothing downstream. One graph serves every batch of a shape."""
        cfg = self.cfg
        R = int(rows.numel())
        if n_valid is None:
            bucket = 4 * int(images.shape[0])
            cap = (R + bucket - 1) // bucket * bucket
        else:
            cap = R
        key = (tuple(images.shape), cap)
        g = self._graphs.get(key)
        if g is None and (self._eager_steps < 2 or len(self._graphs) >= self.max_graphs):
            # warm-up, or a workload that keeps changing shape: launch eagerly instead of capturing one graph (and one private
            # activation pool) per shape
            self._fwd_bwd(images, mask_u8, rows, Noise(seed=seed), n_valid)
            self._eager_steps += 1
            return
        if g is None:
            st = dict(images=torch.empty_like(images), mask=torch.empty_like(mask_u8),
                      rows=torch.zeros(cap, dtype=torch.int32, device=self.dev), nvalid=torch.zeros(1, dtype=torch.int32, device=self.dev),
                      seed=torch.zeros(1, dtype=torch.int64, device=self.dev),
                      dps=torch.empty(cfg.depth, 4 if cfg.dist else 2, images.shape[0], dtype=torch.float32, device=self.dev))
            st["nvalid"].fill_(cap)
            noise = Noise(seed=0, seed_dev=st["seed"], drop_path_scale=st["dps"] if cfg.drop_path_rate > 0 else None)
            graph = torch.cuda.CUDAGraph()
            launches0 = ops.LAUNCHES
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(graph):
                self._fwd_bwd(st["images"], st["mask"], st["rows"], noise, st["nvalid"])
            g = dict(graph=graph, st=st, launches=ops.LAUNCHES - launches0)
            ops.LAUNCHES = launches0
            self._graphs[key] = g
        st = g["st"]
        st["images"].copy_(images, non_blocking=True)
        st["mask"].copy_(mask_u8, non_blocking=True)
        st["rows"][:R].copy_(rows, non_blocking=True)       # entries beyond R keep earlier (valid) row numbers: they are padding
        if n_valid is None:
            st["nvalid"].fill_(R)
        else:
            st["nvalid"].copy_(n_valid, non_blocking=True)
        st["seed"].fill_(seed - (1 << 64) if seed >= (1 << 63) else seed)
        if cfg.drop_path_rate > 0:
            ops.drop_path_scales(cfg.drop_path_probs, 4 if cfg.dist else 2, images.shape[0], seed, self.dev, out=st["dps"])
        g["graph"].replay()
        ops.LAUNCHES += g["launches"]
        self._ar_done_in_step = self.overlap_ar        # the captured graph contains both gradient all-reduces

    def _fwd_bwd_dist(self, images, mask_u8, rows, noise, n_valid=None):
        """--stochastic step (engine_for_cyclical.py:69-86,125-126,152-158): dual-stream teacher/student, targets for both streams,
        smooth-L1 on the mean stream + WassersteinLoss(lambda) on (mean, cov) outputs vs (mean, cov) targets (everything but the optimiser)."""
        cfg = self.cfg
        B = images.shape[0]
        C, T = cfg.embed_dim, cfg.tokens
        M = B * T
        R = rows.numel()
        dev = self.dev
        self._draw_masks(B, noise)
        with self._teacher_stream_ctx():
            self.g32.zero_()
            (lm, lc), _ = core.dist_forward(self.teacher, cfg, images, mode="layers", train=False, save=False, collect=self.target_layers)
        (om, oc), ctx = core.dist_forward(self.student, cfg, images, mask_u8=mask_u8, row_index=rows, mode="masked", train=True, save=True, noise=noise)
        self._teacher_join()
        ls = self.loss_scale if self.loss_scale != -1 else 1.0
        tgt_m = torch.empty((R, C), dtype=torch.float32, device=dev)
        tgt_c = torch.empty((R, C), dtype=torch.float32, device=dev)
        d_m = torch.empty((R, C), dtype=torch.float32, device=dev)
        d_c = torch.zeros((R, C), dtype=torch.float32, device=dev)
        self._targets_and_loss([lm[i].view(M, C) for i in self.target_layers], rows, om, R, n_valid, targets=tgt_m, dy_f32=d_m)
        self._targets_and_loss([lc[i].view(M, C) for i in self.target_layers], rows, None, R, n_valid, targets=tgt_c, with_loss=False,
                               mean_stream=False)
        del lm, lc
        work = torch.empty((2 * R + 8,), dtype=torch.float32, device=dev)
        if self.wloss_dev is None:
            self.wloss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        ops.wasserstein_loss(om, oc, tgt_m, tgt_c, self.lam, ls, work, d_m, d_c, self.wloss_dev, n_valid=n_valid)
        ops.scalar_fma(self.loss_dev, self.loss_dev, ls, self.wloss_dev, ls)      # loss = (loss_cyc + std_loss0*var_w0 + loss_stochastic) * loss_scale  (:160-163)
        self._overlapped_backward(lambda cb: core.dist_backward(self.student, cfg, ctx, d_m, d_c, self.grads, after_block=cb))

    def _optimizer_step(self, lr, wd):
        if self.world_size > 1 and not self._ar_done_in_step:
            torch.distributed.all_reduce(self.g32, group=self.pg)       # DDP gradient mean = sum / world (folded into grad_div)
        self._ar_done_in_step = False
        self.gnorm_sq.zero_()
        ops.sumsq(self.g32, self.gnorm_sq)
        self.opt_step += 1
        # engine_for_cyclical.py:182-185: no EMA update when the decay is 1 or past start_lr_decay_at_step; the loop variable then drops to 0
        do_ema = self.e32 is not None and self.cur_decay != 1 and (self.start_lr_decay_at_step == -1 or self.it <= self.start_lr_decay_at_step)
        ops.adamw_step(self.p32, self.g32, self.m32, self.v32, self.hp, self.opt_step, lr, wd, self.betas[0], self.betas[1], self.eps,
                       gnorm_sq=self.gnorm_sq, max_norm=self.clip if self.clip else 0.0, grad_div=float(self.world_size), p_bf16=self.p16,
                       ema=self.e32 if do_ema else None, ema_decay=self.cur_decay, ema_bf16=self.e16 if do_ema else None)
        if do_ema and self.rel_e is not None:
            ops.ema_index_update(self.rel_e, self.student.rel_index_i32(), self.cur_decay)
        self.ema_updated = do_ema
        if self.e32 is not None and not do_ema:
            self.cur_decay = 0
        self.it += 1
        # the module aliases the arena, but its bf16 weight shadows / captured eval graphs are keyed on autograd version counters, which a
        # raw-pointer kernel update does not touch: tell the module its weights moved
        self.model._weights_version = getattr(self.model, "_weights_version", 0) + 1
        return self.loss_dev

    def grad_norm(self) -> torch.Tensor:
        return torch.sqrt(self.gnorm_sq) / self.world_size

    def _images_to_device(self, images_pinned: torch.Tensor) -> torch.Tensor:
        """Host->device copy of one batch on the current (copy) stream. fp32 batches as the reference loader yields them, or uint8 pixels
        ([B,H,W,3] / [B,3,H,W]) when `pixel_norm` is set: a quarter of the PCIe bytes, normalised by b200vit_normalize_u8."""
        if images_pinned.dtype == torch.uint8:
            if self.pixel_norm is None:
                raise ValueError("uint8 image batch: set engine.pixel_norm = (mean, std, hwc) first")
            mean, std, hwc = self.pixel_norm
            return ops.normalize_u8(images_pinned.to(self.dev, non_blocking=True), mean, std, hwc)
        return images_pinned.to(self.dev, non_blocking=True)

    def stage_host(self, images_pinned: torch.Tensor, mask_host: np.ndarray):
        """Enqueues the host->device copies of ONE batch on the engine's copy stream (so the next batch travels while the current
        step computes) and returns the staged batch for step_staged(). Inputs as the data loader yields them
        (engine_for_cyclical.py:58-60): pinned fp32 images + integer mask."""
        B = images_pinned.shape[0]
        m = np.ascontiguousarray(mask_host.reshape(B, -1).astype(np.uint8))
        rows = torch.from_numpy(self.rows_from_host_mask(m, self.cfg.tokens)).pin_memory()
        mask_pinned = torch.from_numpy(m.reshape(-1)).pin_memory()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        with torch.cuda.stream(self._copy_stream):
            images = self._images_to_device(images_pinned)
            mask_u8 = mask_pinned.to(self.dev, non_blocking=True)
            rows_d = rows.to(self.dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return images, mask_u8, rows_d, ev, (mask_pinned, rows)

    def stage_device_masks(self, images_pinned: torch.Tensor, generator):
        """stage_host() for a pipeline whose masks are drawn on the device (masking_generator.MaskingGenerator.batch): only the images
        cross PCIe. The block masks, the masked-row list (padded to batch x num_masking_patches rows) and the masked-row count stay in
        device memory — nothing is read back, and the step keeps one launch shape (see _fwd_bwd_graphed)."""
        B = images_pinned.shape[0]
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        if (generator.height * generator.width + 1) != self.cfg.tokens:
            raise ValueError(f"mask grid {generator.get_shape()} does not match the model's {self.cfg.tokens - 1} patches")
        with torch.cuda.stream(self._copy_stream):
            rows = torch.zeros(B * generator.num_masking_patches, dtype=torch.int32, device=self.dev)   # padding = row 0 (a cls row)
            mask_u8, count, rows = generator.batch(B, rows=rows)
            images = self._images_to_device(images_pinned)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return images, mask_u8.view(-1), rows, ev, (count,), count[B:B + 1]

    def launch_staged(self, staged, **kw) -> torch.Tensor:
        """Enqueues one step on a batch returned by stage_host() and returns the DEVICE loss scalar without synchronising: the caller can
        stage the next batch while the step runs and read the loss afterwards."""
        images, mask_u8, rows, ev, _keepalive, *rest = staged
        n_valid = rest[0] if rest else None          # device masks: rows is padded, the true count lives on the device
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ev)
        for t in (images, mask_u8, rows) + ((n_valid,) if n_valid is not None else ()):
            t.record_stream(cur)
        return self.step(images, mask_u8, rows, n_valid=n_valid, **kw)

    def read_loss_async(self, loss_dev: torch.Tensor):
        """Enqueues the device->host copy of a step's loss scalar behind the step and returns a handle whose .wait() yields the float. Reading
        the loss of step i after step i + 1 has been enqueued keeps the GPU's queue non-empty across the host's per-step work (the reference's
        `loss.item()` right after the step, engine_for_cyclical.py:164, idles the device for the length of that work)."""
        ring = self.__dict__.setdefault("_loss_ring", [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(4)])
        k = self.__dict__.get("_loss_ring_pos", 0)
        self._loss_ring_pos = (k + 1) % len(ring)
        host = ring[k]
        host.copy_(loss_dev.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))

        class _Pending:
            def wait(_self) -> float:
                ev.synchronize()
                return float(host[0])
        return _Pending()

    def step_staged(self, staged, **kw) -> float:
        """One step on a batch returned by stage_host(); reads the loss back (engine_for_cyclical.py:164)."""
        return float(self.launch_staged(staged, **kw).item())

    def step_host(self, images_pinned: torch.Tensor, mask_host: np.ndarray, **kw) -> float:
        """The reference-facing call: HOST batch in (pinned images + integer mask as the data loader yields them,
        engine_for_cyclical.py:58-60), loss value out (:164). Host->device copies and the loss read-back are inside."""
        return self.step_staged(self.stage_host(images_pinned, mask_host), **kw)


class FinetuneEngine(D2VEngine):
    """Fused fine-tune TRAIN step (run_class_finetuning.py / engine_for_finetuning(_dist).py) over the same flat arenas as the pre-training
    engine: classifier forward + backward on the CUDA schedules, the reference's layer-decay parameter groups as per-chunk lr scales of the
    fused clip + AdamW kernel (optim_factory.py:33-97, LayerDecayValueAssigner), no EMA teacher.
      det   : loss = SoftTargetCrossEntropy / LabelSmoothingCrossEntropy(model(x), targets)
      dist  : loss = CE(logits) + WassersteinLossFineTuning(anchor, positive, negative features); the positive / negative forwards run in
              EVAL mode (no drop-path, no dropout) and without a gradient path, like the eval-mode deep copy of
              engine_for_finetuning_dist.py:293-296.
    The criterion and its gradients are one C-ABI call (b200vit_finetune_loss); forward + loss + backward replay from a CUDA graph."""

    def __init__(self, model, *, lr=5e-4, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, clip_grad=None, layer_decay: Optional[float] = 0.65,
                 lambda_finetuning=1e-4, lambda_pvn=1e-4, smoothing: float = 0.1, world_size=1, process_group=None, seed=0, use_graph: bool = True):
        if model.cfg.kind != "finetune":
            raise B200VitError("FinetuneEngine needs a classifier (VisionTransformer / DistVisionTransformer)")
        super().__init__(model, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps, clip_grad=clip_grad, ema_decay=1.0, ema_decay_init=1.0,
                         ema_start_at=0, layer_decay=layer_decay, world_size=world_size, process_group=process_group, seed=seed,
                         use_graph=use_graph, with_ema=False)
        self.lam_ft, self.lam_pvn = lambda_finetuning, lambda_pvn
        self.smoothing = float(smoothing)           # run_class_finetuning.py:619-624: LabelSmoothingCrossEntropy(args.smoothing) for index labels
        K, C = model.cfg.num_classes, model.cfg.embed_dim
        kp = (K + 7) // 8 * 8
        ow, ob = self.layout["head.weight"][0], self.layout["head.bias"][0]
        # gradient views over the classifier rows INCLUDING the zero padding up to a multiple of 8 (the GEMM's N): core._head_backward
        # accumulates straight into them
        self.grads["head.weight__padded"] = self.g32[ow: ow + kp * C].view(kp, C)
        self.grads["head.bias__padded"] = self.g32[ob: ob + kp]
        self.loss3 = torch.zeros(3, dtype=torch.float32, device=self.dev)
        self._ft_graphs = {}
        self.last_logits = None

    def soft_targets(self, targets: torch.Tensor) -> torch.Tensor:
        """[B, K] fp32 targets: soft targets pass through; index labels become (1 - s) one-hot + s / K (== LabelSmoothingCrossEntropy)."""
        K = self.cfg.num_classes
        if targets.dim() == 1:
            s_ = self.smoothing
            return ops.mixup_batch(None, 1.0, labels=targets.to(self.dev).long().contiguous(), num_classes=K, on_value=1.0 - s_ + s_ / K,
                                   off_value=s_ / K)
        return targets.to(self.dev).float().contiguous()

    def _fwd_bwd_ft(self, images, soft, pos_images, neg_images, noise):
        cfg = self.cfg
        K = cfg.num_classes
        B = images.shape[0]
        kp = (K + 7) // 8 * 8
        dl16 = torch.empty((B, kp), dtype=torch.bfloat16, device=self.dev)
        if cfg.dist:
            trip = pos_images is not None and neg_images is not None
            if trip:
                # positive and negative images: ONE eval-mode forward of 2B images (no batch statistics anywhere: identical to two forwards;
                # engine_for_finetuning_dist.py:293-296 runs them through a deepcopy in eval mode), on the second stream next to the anchor's
                # training forward, which it shares nothing with
                quiet = Noise(drop_path_active=False, attn_drop_active=False)
                both = torch.cat((pos_images, neg_images), 0)
                with self._teacher_stream_ctx():
                    self.g32.zero_()               # the gradient arena (1.6 GB for ViT-L) is cleared beside the forwards
                    (bm, bc, _), _ = core.dist_forward(self.student, cfg, both, mode="logits", train=False, save=False, noise=quiet)
            (fm, fc, logits), ctx = core.dist_forward(self.student, cfg, images, mode="logits", train=True, save=True, noise=noise)
            feats = None
            if trip:
                self._teacher_join()
                feats = (fm, fc, bm[:B], bc[:B], bm[B:], bc[B:])
            _, dfm, dfc = ops.finetune_loss(logits, soft, K, feats=feats, lam_ft=self.lam_ft, lam_pvn=self.lam_pvn, dlogits_bf16=dl16,
                                            loss_out=self.loss3)
            if not trip:
                self.g32.zero_()
            core.dist_backward_logits(self.student, cfg, ctx, dfm, dfc, dl16, self.grads)
        else:
            logits, ctx = core.vit_forward(self.student, cfg, images, mode="logits", train=True, save=True, noise=noise)
            ops.finetune_loss(logits, soft, K, dlogits_bf16=dl16, loss_out=self.loss3)
            self.g32.zero_()
            core.vit_backward_logits(self.student, cfg, ctx, dl16, self.grads)
        return logits

    def _fwd_bwd_ft_graphed(self, images, soft, pos_images, neg_images, seed):
        cfg = self.cfg
        trip = pos_images is not None and neg_images is not None
        key = (tuple(images.shape), trip)
        g = self._ft_graphs.get(key)
        if g is None and (self._eager_steps < 2 or len(self._ft_graphs) >= self.max_graphs):
            self._eager_steps += 1
            return self._fwd_bwd_ft(images, soft, pos_images, neg_images, Noise(seed=seed))
        if g is None:
            st = dict(images=torch.empty_like(images), soft=torch.empty_like(soft), seed=torch.zeros(1, dtype=torch.int64, device=self.dev),
                      dps=torch.empty(cfg.depth, 4 if cfg.dist else 2, images.shape[0], dtype=torch.float32, device=self.dev),
                      pos=torch.empty_like(images) if trip else None, neg=torch.empty_like(images) if trip else None)
            noise = Noise(seed=0, seed_dev=st["seed"], drop_path_scale=st["dps"] if cfg.drop_path_rate > 0 else None)
            graph = torch.cuda.CUDAGraph()
            launches0 = ops.LAUNCHES
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(graph):
                logits = self._fwd_bwd_ft(st["images"], st["soft"], st["pos"], st["neg"], noise)
            g = dict(graph=graph, st=st, launches=ops.LAUNCHES - launches0, logits=logits)
            ops.LAUNCHES = launches0
            self._ft_graphs[key] = g
        st = g["st"]
        st["images"].copy_(images, non_blocking=True)
        st["soft"].copy_(soft, non_blocking=True)
        if trip:
            st["pos"].copy_(pos_images, non_blocking=True)
            st["neg"].copy_(neg_images, non_blocking=True)
        st["seed"].fill_(seed - (1 << 64) if seed >= (1 << 63) else seed)
        if cfg.drop_path_rate > 0:
            ops.drop_path_scales(cfg.drop_path_probs, 4 if cfg.dist else 2, images.shape[0], seed, self.dev, out=st["dps"])
        g["graph"].replay()
        ops.LAUNCHES += g["launches"]
        return g["logits"]

    def step(self, images: torch.Tensor, targets: torch.Tensor, pos_images: Optional[torch.Tensor] = None, neg_images: Optional[torch.Tensor] = None,
             *, lr: Optional[float] = None, weight_decay: Optional[float] = None, noise: Optional[Noise] = None, graph: Optional[bool] = None) -> torch.Tensor:
        """images fp32 [B,3,H,W]; targets: soft targets [B,K] (Mixup / CutMix) or class indices [B]. Returns the device scalar loss."""
        lr = self.lr if lr is None else lr
        wd = self.wd if weight_decay is None else weight_decay
        self.cur_decay = 1.0
        seed = (self.seed * 0x9E3779B97F4A7C15 + self.it + 1) & 0xFFFFFFFFFFFFFFFF
        soft = self.soft_targets(targets)
        images = images.float().contiguous()
        if graph is None:
            graph = self.use_graph
        if graph and noise is None and ops.GEMM_TIMING is None:
            logits = self._fwd_bwd_ft_graphed(images, soft, pos_images, neg_images, seed)
        else:
            logits = self._fwd_bwd_ft(images, soft, pos_images, neg_images, noise if noise is not None else Noise(seed=seed))
            self._eager_steps += 1
        self.loss_dev = self.loss3[0:1]
        self.last_logits = logits               # class_acc of the training log (engine_for_finetuning.py:132-133)
        return self._optimizer_step(lr, wd)


def finetune_one_epoch(engine: "FinetuneEngine", data_loader: Iterable, epoch: int = 0, start_steps: int = 0, lr_schedule_values=None,
                       wd_schedule_values=None, mixup_fn=None, num_training_steps_per_epoch: Optional[int] = None, update_freq: int = 1,
                       print_freq: int = 10, log=print) -> Dict[str, float]:
    """Loop of engine_for_finetuning.train_one_epoch (:46-170) / engine_for_finetuning_dist.dist_train_one_epoch (:312-440) over the fused
    FinetuneEngine. The loader yields (samples, targets) or, for the --stochastic triplet set, (samples, pos_samples, neg_samples, labels)
    (dist_datasets.py:143-148). Per step: schedule lr (x the group's layer-decay scale inside the fused AdamW) and weight decay, host->device
    copy, device Mixup / CutMix when `mixup_fn` is given (mixup.Mixup), one fused step, loss read back, non-finite loss aborts.
    Gradient accumulation (update_freq > 1) is not implemented: the README recipes use --update_freq 1."""
    if update_freq != 1:
        raise NotImplementedError("update_freq > 1 (gradient accumulation) is outside the B200 hot path; the README recipes use 1")
    engine.model.train(True)
    dev = engine.dev
    total = acc_sum = 0.0
    n = n_acc = 0
    lr = engine.lr
    for data_iter_step, batch in enumerate(data_loader):
        step = data_iter_step
        if num_training_steps_per_epoch is not None and step >= num_training_steps_per_epoch:
            continue
        it = start_steps + step
        lr = float(lr_schedule_values[it]) if lr_schedule_values is not None else None
        wd = float(wd_schedule_values[it]) if wd_schedule_values is not None else None
        if len(batch) == 4:
            samples, pos, neg, labels = batch
            pos, neg = pos.to(dev, non_blocking=True).float(), neg.to(dev, non_blocking=True).float()
        else:
            (samples, labels), pos, neg = batch, None, None
        samples = samples.to(dev, non_blocking=True).float().contiguous()
        labels = labels.to(dev, non_blocking=True)
        targets = labels
        if mixup_fn is not None:
            samples, targets = mixup_fn(samples, labels)
        engine.it = it
        loss = float(engine.step(samples, targets, pos, neg, lr=lr, weight_decay=wd).item())
        if not math.isfinite(loss):
            raise FloatingPointError(f"Loss is {loss}, stopping training")
        total += loss
        n += 1
        if mixup_fn is None and labels.dim() == 1:
            acc_sum += float((engine.last_logits.argmax(-1) == labels).float().mean().item())
            n_acc += 1
        if data_iter_step % print_freq == 0:
            log(f"Epoch: [{epoch}] step {data_iter_step} loss {loss:.4f} lr {lr if lr is not None else engine.lr:.6f}")
    hp = engine.hp[:, 0]
    cur = lr if lr is not None else engine.lr
    stepped = hp[hp > 0]
    return {"loss": total / max(n, 1), "class_acc": acc_sum / n_acc if n_acc else None, "lr": cur * float(hp.max().item()),
            "min_lr": cur * float(stepped.min().item()) if stepped.numel() else cur, "grad_norm": float(engine.grad_norm().item())}


@torch.no_grad()
def evaluate(data_loader: Iterable, model, device=None, num_classes: Optional[int] = None, dist_criterion=None) -> Dict[str, float]:
    """engine_for_finetuning.evaluate (:175-222) / engine_for_finetuning_dist.dist_evaluate (:442-494; pass
    dist_criterion=(lambda_finetuning, lambda_pvn) and a loader of (images, pos, neg, labels): the loss then includes
    WassersteinLossFineTuning of the three forwards): deterministic evaluation with the per-batch metrics of the reference,
    averaged over the batches weighted by batch size (loss: plain mean over batches, as MetricLogger.update(loss=...) does): cross-entropy,
    acc@1 / acc@5 (percent), 15-bin ECE, TACE (threshold 0.01, 30 adaptive bins per class), NLL and macro one-vs-rest AUROC, all reduced on
    the device (b200vit_mc_reduce with one sample, b200vit_tace_auroc, b200vit_finetune_loss). TACE needs at least 30 images per batch
    (its bin edges are order statistics of the batch); smaller batches report NaN for it, where the reference would index out of range."""
    dev = device if device is not None else next(model.parameters()).device
    sums = {"acc1": 0.0, "acc5": 0.0, "ECE": 0.0, "ECE_as_reference_computes_it": 0.0, "NLL": 0.0, "TACE": 0.0, "TACE_as_reference_computes_it": 0.0,
            "AUROC": 0.0}
    loss_sum, nb, ntot = 0.0, 0, 0
    was_training = model.training
    model.eval()
    K = model.cfg.num_classes
    for batch in data_loader:
        images, target = batch[0].to(dev).float().contiguous(), batch[-1].to(dev)
        out = model(images)
        logits = (out[-1] if isinstance(out, (tuple, list)) else out).float().contiguous()
        labels32 = target.to(torch.int32)
        _, _, _, summary = ops.mc_reduce(logits.unsqueeze(0).contiguous(), labels32)
        a1, a5, ece, ece_ref, nll = summary.tolist()[:5]
        b = images.shape[0]
        if b >= 30:
            tace, tace_ref, auroc = ops.tace_auroc(logits, labels32).tolist()
        else:
            tace, tace_ref, auroc = float("nan"), float("nan"), ops.tace_auroc(logits, labels32, n_bins=max(1, min(30, b))).tolist()[2]
        onehot = ops.mixup_batch(None, 1.0, labels=target.long().contiguous(), num_classes=K, on_value=1.0, off_value=0.0)   # nn.CrossEntropyLoss
        feats = None
        lam = (1e-4, 1e-4)
        if dist_criterion is not None and len(batch) == 4 and isinstance(out, (tuple, list)):
            pm, pc, _ = model(batch[1].to(dev).float().contiguous())
            nm, nc, _ = model(batch[2].to(dev).float().contiguous())
            feats = tuple(t.float().contiguous() for t in (out[0], out[1], pm, pc, nm, nc))
            lam = dist_criterion
        loss3, _, _ = ops.finetune_loss(logits, onehot, K, feats=feats, lam_ft=lam[0], lam_pvn=lam[1])
        loss_sum += float(loss3[0].item())
        nb += 1
        ntot += b
        for k, v in zip(sums, (a1, a5, ece, ece_ref, nll, tace, tace_ref, auroc)):
            sums[k] += v * b
    model.train(was_training)
    out = {k: v / max(ntot, 1) for k, v in sums.items()}
    out["loss"] = loss_sum / max(nb, 1)
    return out


def train_one_epoch(engine: D2VEngine, data_loader: Iterable, epoch: int = 0, start_steps: int = 0, lr_schedule_values=None,
                    wd_schedule_values=None, print_freq: int = 10, log=print, mask_generator=None) -> Dict[str, float]:
    """Loop of engine_for_cyclical.train_one_epoch (:45-225) over the fused engine: per-step lr/wd from the schedule tables,
    EMA-decay anneal, non-finite loss aborts (:166-168). With `mask_generator` (masking_generator.MaskingGenerator) the masks of each
    batch are drawn on the device and whatever mask the loader yields is ignored (the loader may then yield bare image batches)."""
    total, n = 0.0, 0
    engine.cur_decay = engine.ema_decay                                     # engine_for_cyclical.py:44

    def staged_batches():
        for batch, _ in data_loader:
            samples, bool_masked_pos = batch if isinstance(batch, (tuple, list)) else (batch, None)
            if not samples.is_pinned():
                samples = samples.pin_memory()
            if mask_generator is not None:
                yield engine.stage_device_masks(samples, mask_generator)
            else:
                yield engine.stage_host(samples, np.asarray(bool_masked_pos))

    it_batches = staged_batches()
    nxt = next(it_batches, None)
    step = 0
    pending = None                                                         # (handle, step number) of the loss not yet read back

    def settle(p):
        nonlocal total, n
        loss = p[0].wait()
        if not math.isfinite(loss):
            raise FloatingPointError(f"Loss is {loss}, stopping training")      # engine_for_cyclical.py:166-168 (one step later than the reference)
        total += loss
        n += 1
        if p[1] % print_freq == 0:
            log(f"Epoch: [{epoch}] step {p[1]} loss {loss:.4f} ema_decay {engine.cur_decay:.6f}")

    while nxt is not None:
        cur = nxt
        it = start_steps + step
        lr = float(lr_schedule_values[it]) if lr_schedule_values is not None else None
        wd = float(wd_schedule_values[it]) if wd_schedule_values is not None else None
        engine.it = it
        loss_dev = engine.launch_staged(cur, lr=lr, weight_decay=wd)      # enqueue step `step` ...
        handle = engine.read_loss_async(loss_dev)                          # ... and the 4-byte read-back of its loss behind it,
        nxt = next(it_batches, None)                                       # stage batch step+1 (host work + H2D) while it runs,
        if pending is not None:
            settle(pending)                                                # and only now wait for the loss of the PREVIOUS step (:164)
        pending = (handle, step)
        step += 1
    if pending is not None:
        settle(pending)
    return {"loss": total / max(n, 1), "cur_decay": engine.cur_decay}
