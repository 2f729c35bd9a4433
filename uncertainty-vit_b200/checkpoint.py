"""Checkpoint I/O for the fused engines and the pre-train -> fine-tune key remapping.

Reference behaviour restated here (host-side, PyTorch tensors only):
  utils.save_model / auto_load_model (utils.py:462-545)     -> save_model / auto_load_model over an engine
  utils.load_state_dict (utils.py:315-361)                    -> load_state_dict (same key filtering and messages)
  optim_factory.get_parameter_groups (optim_factory.py:58-97) -> parameter_groups (group order = torch's param indices)
  run_class_finetuning.py:391-540                             -> prepare_finetune_checkpoint

A checkpoint written by an engine has the reference's layout — {'model', 'optimizer', 'epoch', 'model_ema', ...} with the optimizer entry
in torch.optim.AdamW's state_dict format over the reference's parameter groups — so a run can move between the reference runner and the
fused engine in either direction. The engine keeps all state in flat arenas; the dictionaries below are views or copies of them."""
from __future__ import annotations

import glob
import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from .engine import get_num_layer_for_vit


# ---------------------------------------------------------------------------------------------------------------------------------
# parameter groups
# ---------------------------------------------------------------------------------------------------------------------------------
def parameter_groups(named_shapes: Iterable, weight_decay: float, skip_list: Iterable[str] = (), num_layers: Optional[int] = None,
                     layer_scales: Optional[Sequence[float]] = None, frozen: Iterable[str] = ()) -> List[dict]:
    """optim_factory.get_parameter_groups (:58-97) on (name, shape) pairs: groups appear in the order their first parameter does,
    'decay' / 'no_decay' (1-D, *.bias, skip list) and, with layer decay, 'layer_<id>_<decay|no_decay>' with lr_scale = layer_scales[id].
    num_layers = len(layer_scales) = depth + 2 (LayerDecayValueAssigner). Returns [{'name', 'weight_decay', 'lr_scale', 'params': [names]}]."""
    skip, frozen = set(skip_list), set(frozen)
    groups: Dict[str, dict] = {}
    for name, shape in named_shapes:
        if name in frozen:
            continue
        if len(shape) == 1 or name.endswith(".bias") or name in skip:
            gname, wd = "no_decay", 0.0
        else:
            gname, wd = "decay", weight_decay
        layer_id = None
        if num_layers is not None:
            layer_id = get_num_layer_for_vit(name, num_layers)
            gname = "layer_%d_%s" % (layer_id, gname)
        if gname not in groups:
            scale = layer_scales[layer_id] if layer_scales is not None else 1.0
            groups[gname] = {"name": gname, "weight_decay": wd, "lr_scale": scale, "params": []}
        groups[gname]["params"].append(name)
    return list(groups.values())


def engine_parameter_groups(engine) -> List[dict]:
    named = [(n, tuple(p.shape)) for n, p in engine.model.named_parameters() if p.requires_grad]
    if engine.layer_decay is None:
        return parameter_groups(named, engine.wd, engine.skip_weight_decay)
    L = engine.cfg.depth + 2
    scales = [engine.layer_decay ** (L - 1 - i) for i in range(L)]          # run_class_finetuning.py:569-573
    return parameter_groups(named, engine.wd, engine.skip_weight_decay, L, scales)


# ---------------------------------------------------------------------------------------------------------------------------------
# engine <-> reference checkpoint layout
# ---------------------------------------------------------------------------------------------------------------------------------
def _arena_view(engine, arena: torch.Tensor, name: str) -> torch.Tensor:
    o, shape = engine.layout[name]
    return arena[o: o + int(np.prod(shape))].view(shape)


def _never_stepped(engine, name: str) -> bool:
    """Parameters torch's AdamW never creates state for: those that never receive a gradient (dist cov_qkv.weight, engine.py layout)."""
    o, _ = engine.layout[name]
    lr_scale, wd_scale = engine.hp[o // 1024].tolist()
    return lr_scale == 0.0 and wd_scale == 0.0


def optimizer_state_dict(engine) -> dict:
    """torch.optim.AdamW.state_dict() of the optimizer the reference would have built with optim_factory.create_optimizer: parameters
    numbered group by group, {'step', 'exp_avg', 'exp_avg_sq'} per stepped parameter, one param_group per decay / layer group."""
    groups = engine_parameter_groups(engine)
    state, param_groups, idx = {}, [], 0
    for g in groups:
        ids = []
        for name in g["params"]:
            if engine.opt_step > 0 and not _never_stepped(engine, name):
                state[idx] = {"step": torch.tensor(float(engine.opt_step)), "exp_avg": _arena_view(engine, engine.m32, name).detach().clone().cpu(),
                              "exp_avg_sq": _arena_view(engine, engine.v32, name).detach().clone().cpu()}
            ids.append(idx)
            idx += 1
        param_groups.append({"lr": engine.lr * g["lr_scale"], "betas": tuple(engine.betas), "eps": engine.eps, "weight_decay": g["weight_decay"],
                             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                             "lr_scale": g["lr_scale"], "params": ids})
    return {"state": state, "param_groups": param_groups}


def load_optimizer_state_dict(engine, sd: dict) -> None:
    groups = engine_parameter_groups(engine)
    names = [n for g in groups for n in g["params"]]
    n_saved = sum(len(g["params"]) for g in sd["param_groups"])
    if n_saved != len(names) or [len(g["params"]) for g in sd["param_groups"]] != [len(g["params"]) for g in groups]:
        raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")   # torch's message
    flat_ids = [i for g in sd["param_groups"] for i in g["params"]]
    engine.m32.zero_()
    engine.v32.zero_()
    step = 0
    for saved_id, name in zip(flat_ids, names):
        st = sd["state"].get(saved_id)
        if st is None:
            continue
        _arena_view(engine, engine.m32, name).copy_(st["exp_avg"].to(engine.dev, torch.float32))
        _arena_view(engine, engine.v32, name).copy_(st["exp_avg_sq"].to(engine.dev, torch.float32))
        step = max(step, int(float(st["step"])))
    engine.opt_step = step


def engine_state_dict(engine, epoch: Optional[int] = None) -> dict:
    """What utils.save_model stores (utils.py:466-478) + the engine's own counters (step number for the EMA / lr schedules, the noise seed,
    the image counter of a device mask generator is kept by the caller)."""
    out = {"model": {k: v.detach().clone().cpu() for k, v in engine.model.state_dict().items()},
           "optimizer": optimizer_state_dict(engine), "epoch": epoch,
           "engine": {"it": engine.it, "opt_step": engine.opt_step, "seed": engine.seed}}
    if engine.e32 is not None:
        ema = {k: v.detach().clone().cpu() for k, v in engine.ema_state_dict().items()}
        for k, v in out["model"].items():           # buffers (relative_position_index) are not in the arena: ModelEmaV2 deep-copies them
            ema.setdefault(k, v.clone())
        if getattr(engine, "rel_e", None) is not None:
            # the teacher's OWN index buffer: ModelEmaV2._update runs it through the EMA arithmetic as well (see b200vit_ema_index_update)
            ema["rel_pos_bias.relative_position_index"] = engine.rel_e.detach().to(torch.int64).cpu()
        out["model_ema"] = {k: ema[k] for k in out["model"]}
    return out


def load_engine_state_dict(engine, ckpt: dict, load_optimizer: bool = True) -> None:
    """auto_load_model's torch.amp branch (utils.py:505-520): model weights (strict), then optimizer + epoch + EMA when present."""
    from . import ops
    engine.model.load_state_dict(ckpt["model"])                   # parameters alias the fp32 arena: copied in place
    ops.cast_bf16(engine.p32, engine.p16)
    if engine.e32 is not None:
        ema = ckpt.get("model_ema")
        if ema is None:
            engine.e32.copy_(engine.p32)
        else:
            for name in engine.ema_state_dict():
                _arena_view(engine, engine.e32, name).copy_(ema[name].to(engine.dev, torch.float32))
            idx = ema.get("rel_pos_bias.relative_position_index")
            if idx is not None and getattr(engine, "rel_e", None) is not None:
                engine.rel_e.copy_(idx.to(engine.dev, torch.int32))
        ops.cast_bf16(engine.e32, engine.e16)
    if load_optimizer and "optimizer" in ckpt:
        load_optimizer_state_dict(engine, ckpt["optimizer"])
    meta = ckpt.get("engine")
    if meta is not None:
        engine.it = int(meta["it"])       # (the noise seed is per rank, args.seed + rank, and stays what the constructor was given)
        if load_optimizer:
            engine.opt_step = int(meta["opt_step"])
    engine.model._weights_version = getattr(engine.model, "_weights_version", 0) + 1
    if hasattr(engine.model, "_ps"):
        engine.model._ps.invalidate()


def save_model(output_dir: str, epoch, engine, args=None, is_main_process: bool = True) -> Optional[str]:
    """utils.save_model (utils.py:462-485), torch.amp branch: <output_dir>/checkpoint-<epoch>.pth written by the main process."""
    if not is_main_process:
        return None
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, "checkpoint-%s.pth" % str(epoch))
    to_save = engine_state_dict(engine, epoch)
    if args is not None:
        to_save["args"] = args
    torch.save(to_save, path)
    return path


def latest_checkpoint(output_dir: str) -> Optional[str]:
    """The scan of auto_load_model (utils.py:491-503): the highest all-digit <n> among checkpoint-<n>.pth."""
    latest = -1
    for ckpt in glob.glob(os.path.join(glob.escape(output_dir), "checkpoint-*.pth")):
        t = ckpt.split("-")[-1].split(".")[0]
        if t.isdigit():
            latest = max(int(t), latest)
    return os.path.join(output_dir, "checkpoint-%d.pth" % latest) if latest >= 0 else None


def auto_load_model(output_dir: str, engine, resume: str = "", auto_resume: bool = True, reset_resume: bool = False) -> int:
    """utils.auto_load_model (utils.py:488-520). Returns the epoch to start from (checkpoint epoch + 1, or 0 when nothing was loaded)."""
    if auto_resume and len(resume) == 0:
        resume = latest_checkpoint(output_dir) or ""
    if not resume:
        return 0
    ckpt = torch.load(resume, map_location="cpu", weights_only=False)
    with_opt = "optimizer" in ckpt and "epoch" in ckpt and not reset_resume
    load_engine_state_dict(engine, ckpt, load_optimizer=with_opt)
    return int(ckpt["epoch"]) + 1 if with_opt and ckpt["epoch"] is not None else 0


# ---------------------------------------------------------------------------------------------------------------------------------
# utils.load_state_dict
# ---------------------------------------------------------------------------------------------------------------------------------
def load_state_dict(model: torch.nn.Module, state_dict: dict, prefix: str = "", ignore_missing: str = "relative_position_index", log=print):
    """utils.load_state_dict (utils.py:315-361): non-strict, module-by-module load under `prefix`, missing keys that contain one of the
    '|'-separated `ignore_missing` fragments are not reported. Returns (missing, unexpected, errors) after printing the reference's lines."""
    missing_keys: List[str] = []
    unexpected_keys: List[str] = []
    error_msgs: List[str] = []
    metadata = getattr(state_dict, "_metadata", None)
    state_dict = state_dict.copy()
    if metadata is not None:
        state_dict._metadata = metadata

    def load(module, pre=""):
        local_metadata = {} if metadata is None else metadata.get(pre[:-1], {})
        module._load_from_state_dict(state_dict, pre, local_metadata, True, missing_keys, unexpected_keys, error_msgs)
        for name, child in module._modules.items():
            if child is not None:
                load(child, pre + name + ".")

    load(model, prefix)
    fragments = ignore_missing.split("|")
    missing = [k for k in missing_keys if not any(f in k for f in fragments)]
    if missing:
        log("Weights of {} not initialized from pretrained model: {}".format(model.__class__.__name__, missing))
    if unexpected_keys:
        log("Weights from pretrained model not used in {}: {}".format(model.__class__.__name__, unexpected_keys))
    if error_msgs:
        log("\n".join(error_msgs))
    if hasattr(model, "_ps"):
        model._ps.invalidate()
    return missing, unexpected_keys, error_msgs


# ---------------------------------------------------------------------------------------------------------------------------------
# pre-train -> fine-tune remapping
# ---------------------------------------------------------------------------------------------------------------------------------
def _geometric_positions(src_size: int, dst_size: int):
    """run_class_finetuning.py:451-476: source sample positions on a geometric progression whose half-width matches dst_size // 2."""
    left, right = 1.01, 1.5
    while right - left > 1e-6:
        q = (left + right) / 2.0
        gp = 1.0 * (1.0 - q ** (src_size // 2)) / (1.0 - q)
        if gp > dst_size // 2:
            right = q
        else:
            left = q
    dis, cur = [], 1
    for i in range(src_size // 2):
        dis.append(cur)
        cur += q ** (i + 1)
    r_ids = [-d for d in reversed(dis)]
    x = r_ids + [0] + dis
    t = dst_size // 2.0
    dx = np.arange(-t, t + 0.1, 1.0)
    return np.asarray(x, dtype=np.float64), dx


def interpolate_rel_pos_bias_table(table: torch.Tensor, dst_num_pos: int, dst_patch_shape) -> torch.Tensor:
    """run_class_finetuning.py:433-494: resize a [(2s-1)^2 + extra, heads] relative-position table to the target window with a cubic
    spline over geometrically spaced source positions; the `extra` (cls) rows are kept.
    PARITY UNPINNED for this branch: the reference calls scipy.interpolate.interp2d(kind='cubic'), which SciPy >= 1.14 (installed: 1.18)
    has removed, so the reference cannot execute it here; RectBivariateSpline(kx=ky=3, s=0) is SciPy's documented replacement on regular
    grids. The identity case (same window, the hot-path configs) never reaches the spline."""
    src_num_pos, heads = table.shape
    if dst_patch_shape[0] != dst_patch_shape[1]:
        raise NotImplementedError()
    extra = dst_num_pos - (dst_patch_shape[0] * 2 - 1) * (dst_patch_shape[1] * 2 - 1)
    src_size = int((src_num_pos - extra) ** 0.5)
    dst_size = int((dst_num_pos - extra) ** 0.5)
    if src_size == dst_size:
        return table
    from scipy.interpolate import RectBivariateSpline
    extra_tokens = table[-extra:, :]
    body = table[:-extra, :]
    x, dx = _geometric_positions(src_size, dst_size)
    cols = []
    for h in range(heads):
        z = body[:, h].view(src_size, src_size).float().numpy().astype(np.float64)
        f = RectBivariateSpline(x, x, z, kx=3, ky=3, s=0)          # z[i, j]: i indexes y (rows), j indexes x — symmetric grids here
        cols.append(torch.from_numpy(f(dx, dx)).float().contiguous().view(-1, 1).to(table.device))
    return torch.cat((torch.cat(cols, dim=-1), extra_tokens), dim=0)


def interpolate_pos_embed(pos_embed: torch.Tensor, num_patches: int, num_extra_tokens: int) -> torch.Tensor:
    """run_class_finetuning.py:497-517: bicubic resize of the patch position embeddings, class / dist tokens unchanged."""
    emb = pos_embed.shape[-1]
    orig = int((pos_embed.shape[-2] - num_extra_tokens) ** 0.5)
    new = int(num_patches ** 0.5)
    if orig == new:
        return pos_embed
    extra = pos_embed[:, :num_extra_tokens]
    tok = pos_embed[:, num_extra_tokens:].reshape(-1, orig, orig, emb).permute(0, 3, 1, 2)
    tok = torch.nn.functional.interpolate(tok, size=(new, new), mode="bicubic", align_corners=False)
    return torch.cat((extra, tok.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)


def prepare_finetune_checkpoint(checkpoint: dict, model: torch.nn.Module, model_key: str = "model|module", dual_finetune: bool = False,
                                reinit_final_norm: bool = False, log=print) -> dict:
    """The checkpoint surgery run_class_finetuning.py does before utils.load_state_dict (:400-517), as a function:
    pick checkpoint[model_key], drop a head of another shape (and, on request, the final norms), expand a shared relative-position table
    to per-block tables when the model uses those, drop relative_position_index buffers, resize relative-position tables and pos_embed
    to the model's window. Returns the state dict to hand to load_state_dict(model, ..., prefix=args.model_prefix)."""
    ckpt_model = None
    for key in model_key.split("|"):
        if key in checkpoint:
            ckpt_model = checkpoint[key]
            break
    if ckpt_model is None:
        ckpt_model = checkpoint
    ckpt_model = dict(ckpt_model)
    state = model.state_dict()
    if not dual_finetune:
        for k in ("head.weight", "head.bias"):
            if k in ckpt_model and k in state and ckpt_model[k].shape != state[k].shape:
                log(f"Removing key {k} from pretrained checkpoint")
                del ckpt_model[k]
        if reinit_final_norm:
            for k in ("norm.weight", "norm.bias", "fc_norm.weight", "fc_norm.bias"):
                if k in ckpt_model:
                    log(f"Removing key {k} from pretrained checkpoint")
                    del ckpt_model[k]
    if getattr(model, "use_rel_pos_bias", False) and "rel_pos_bias.relative_position_bias_table" in ckpt_model:
        log("Expand the shared relative position embedding to each transformer block. ")
        shared = ckpt_model.pop("rel_pos_bias.relative_position_bias_table")
        for i in range(model.get_num_layers()):
            ckpt_model["blocks.%d.attn.relative_position_bias_table" % i] = shared.clone()
    for key in list(ckpt_model.keys()):
        if "relative_position_index" in key:
            ckpt_model.pop(key)
        if "relative_position_bias_table" in key and key in state:
            dst_num_pos = state[key].shape[0]
            new = interpolate_rel_pos_bias_table(ckpt_model[key], dst_num_pos, model.patch_embed.patch_shape)
            if new is not ckpt_model[key]:
                log("Position interpolate for %s to %d positions" % (key, dst_num_pos))
            ckpt_model[key] = new
    if "pos_embed" in ckpt_model and getattr(model, "pos_embed", None) is not None:
        num_patches = model.patch_embed.num_patches
        ckpt_model["pos_embed"] = interpolate_pos_embed(ckpt_model["pos_embed"], num_patches, model.pos_embed.shape[-2] - num_patches)
    return ckpt_model
