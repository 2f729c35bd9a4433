"""CPU oracle — TEST INFRASTRUCTURE ONLY (never imported by the product package).

A plain fp32 PyTorch-on-CPU restatement of the reference's hot path (fx-erick/uncertainty-vit), written
functionally over a state-dict that uses the reference's parameter names, with every source of
randomness (drop-path keeps, dropout masks) INJECTED so that the CUDA path can be compared on identical noise.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.

Parity pin: tools/make_golden.py imports the real reference classes (through oracle/ref_shim.py, in the build
container where /root/reference exists), loads the same seeded state-dict, injects the same masks, and stores the
reference outputs under tests/golden/; tests/test_oracle_golden.py checks this restatement against them.

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------
# architecture description
# --------------------------------------------------------------------------------------------------------------
@dataclass
class Arch:
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: float = 4.0
    num_classes: int = 1000
    ln_eps: float = 1e-6          # partial(nn.LayerNorm, eps=1e-6): modeling_cyclical.py:293, modeling_finetune.py:1227
    dist: bool = False            # dual-stream (--stochastic) model
    kind: str = "cyclical"        # "cyclical" (data2vec student/teacher) | "finetune"

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def tokens(self) -> int:
        return self.num_patches + 1

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)


VIT_B = dict(embed_dim=768, depth=12, num_heads=12)
VIT_L = dict(embed_dim=1024, depth=24, num_heads=16)
TINY = dict(img_size=64, embed_dim=128, depth=2, num_heads=2, num_classes=10)


# --------------------------------------------------------------------------------------------------------------
# relative position index  (modeling_finetune.py:339-353) — integer, bit-exact
# --------------------------------------------------------------------------------------------------------------
def relative_position_index(wh: int, ww: int) -> torch.Tensor:
    nrd = (2 * wh - 1) * (2 * ww - 1) + 3
    idx = torch.zeros((wh * ww + 1, wh * ww + 1), dtype=torch.int64)
    ys, xs = np.meshgrid(np.arange(wh), np.arange(ww), indexing="ij")
    ys = torch.from_numpy(ys.reshape(-1))
    xs = torch.from_numpy(xs.reshape(-1))
    dy = ys[:, None] - ys[None, :] + (wh - 1)
    dx = xs[:, None] - xs[None, :] + (ww - 1)
    idx[1:, 1:] = dy * (2 * ww - 1) + dx
    idx[0, :] = nrd - 3
    idx[:, 0] = nrd - 2
    idx[0, 0] = nrd - 1
    return idx


def rel_pos_bias(table: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """RelativePositionBias.forward (modeling_finetune.py:359-364): table[732,H] -> [H,N,N]."""
    n = index.shape[0]
    return table[index.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1).contiguous()


# --------------------------------------------------------------------------------------------------------------
# seeded "lively" state-dict used by the parity tests (NOT the training init: gamma=1e-4 would hide the blocks)
# --------------------------------------------------------------------------------------------------------------
def state_names(arch: Arch) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, role) in the reference's state_dict order."""
    C, Hd, nh = arch.embed_dim, arch.hidden, arch.num_heads
    P = arch.patch_size
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    add = lambda n, s, r: out.append((n, tuple(s), r))
    add("cls_token", (1, 1, C), "token")
    if arch.dist:
        add("cov_cls_token", (1, 1, C), "token")
    if arch.kind == "cyclical":
        add("mask_token", (1, 1, C), "token")
        if arch.dist:
            add("cov_mask_token", (1, 1, C), "token")
    add("patch_embed.proj.weight", (C, arch.in_chans, P, P), "weight")
    add("patch_embed.proj.bias", (C,), "bias")
    if arch.dist:
        add("cov_patch_embed.proj.weight", (C, arch.in_chans, P, P), "weight")
        add("cov_patch_embed.proj.bias", (C,), "bias")
    nrd = (2 * arch.grid - 1) ** 2 + 3
    add("rel_pos_bias.relative_position_bias_table", (nrd, nh), "table")
    for i in range(arch.depth):
        p = f"blocks.{i}."
        add(p + "gamma_1", (C,), "gamma")
        add(p + "gamma_2", (C,), "gamma")
        add(p + "norm1.weight", (C,), "ln_w")
        add(p + "norm1.bias", (C,), "bias")
        add(p + "attn.q_bias", (C,), "bias")
        add(p + "attn.v_bias", (C,), "bias")
        if arch.dist:
            add(p + "attn.cov_q_bias", (C,), "bias")
            add(p + "attn.cov_v_bias", (C,), "bias")
        add(p + "attn.qkv.weight", (3 * C, C), "weight")
        if arch.dist:
            add(p + "attn.cov_qkv.weight", (3 * C, C), "weight")   # allocated but unused (quirk A.2-1)
        add(p + "attn.proj.weight", (C, C), "weight")
        add(p + "attn.proj.bias", (C,), "bias")
        if arch.dist:
            add(p + "attn.cov_proj.weight", (C, C), "weight")
            add(p + "attn.cov_proj.bias", (C,), "bias")
        add(p + "norm2.weight", (C,), "ln_w")
        add(p + "norm2.bias", (C,), "bias")
        add(p + "mlp.fc1.weight", (Hd, C), "weight")
        add(p + "mlp.fc1.bias", (Hd,), "bias")
        add(p + "mlp.fc2.weight", (C, Hd), "weight")
        add(p + "mlp.fc2.bias", (C,), "bias")
    if arch.kind == "cyclical":
        add("norm.weight", (C,), "ln_w")
        add("norm.bias", (C,), "bias")
        add("lm_head.weight", (C, C), "weight")
        add("lm_head.bias", (C,), "bias")
        if arch.dist:
            add("cov_lm_head.weight", (C, C), "weight")
            add("cov_lm_head.bias", (C,), "bias")
    else:
        add("fc_norm.weight", (C,), "ln_w")
        add("fc_norm.bias", (C,), "bias")
        add("head.weight", (arch.num_classes, C), "weight")
        add("head.bias", (arch.num_classes,), "bias")
    return out


def make_state(arch: Arch, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic test weights (torch CPU generator): every branch carries signal."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, role in state_names(arch):
        if role == "weight":
            fan_in = int(np.prod(shape[1:]))
            t = torch.randn(shape, generator=g) * (0.7 / math.sqrt(fan_in))
        elif role == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif role == "gamma":
            t = 0.25 + 0.5 * torch.rand(shape, generator=g)
        elif role == "ln_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif role == "table":
            t = torch.randn(shape, generator=g) * 0.3
        else:  # token
            t = torch.randn(shape, generator=g) * 0.5
        sd[name] = t.float()
    sd["rel_pos_bias.relative_position_index"] = relative_position_index(arch.grid, arch.grid)
    return sd


# --------------------------------------------------------------------------------------------------------------
# injected randomness
# --------------------------------------------------------------------------------------------------------------
@dataclass
class Noise:
    """Injected stochastic-depth keeps and attention-dropout masks.

    drop_path_keep[l] : float tensor [n_draws, B] in {0,1}; n_draws = 2 (det Block, modeling_finetune.py:296-297)
                        or 4 (dist Block, modeling_finetune_dist.py:51-55), in call order.
    drop_path_prob[l] : p_l = linspace(0, dpr, L)[l]  (modeling_finetune.py:401)
    attn_keep[l]      : float tensor [B,H,N,N] in {0,1} (nn.Dropout(attn_drop) mask, modeling_finetune.py:183)
    attn_drop         : p of that dropout
    """
    drop_path_keep: Optional[List[torch.Tensor]] = None
    drop_path_prob: Optional[List[float]] = None
    attn_keep: Optional[List[torch.Tensor]] = None
    attn_drop: float = 0.0


def _dp(x: torch.Tensor, noise: Optional[Noise], layer: int, draw: int) -> torch.Tensor:
    """timm drop_path (modeling_finetune.py:51-62): x / keep_prob * mask[b], per sample."""
    if noise is None or noise.drop_path_keep is None:
        return x
    p = noise.drop_path_prob[layer]
    if p == 0.0:
        return x
    keep = noise.drop_path_keep[layer][draw].to(x.dtype).view(-1, *([1] * (x.dim() - 1)))
    return x / (1.0 - p) * keep


def _attn_dropout(attn: torch.Tensor, noise: Optional[Noise], layer: int) -> torch.Tensor:
    if noise is None or noise.attn_keep is None or noise.attn_drop == 0.0:
        return attn
    return attn * noise.attn_keep[layer].to(attn.dtype) / (1.0 - noise.attn_drop)


# --------------------------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------------------------
def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def mlp(sd, p, x):
    """Mlp.forward (modeling_finetune.py:75-82): fc2(GELU_erf(fc1 x))."""
    h = F.linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.gelu(h)
    return F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def attention(sd, p, x, bias, nh, noise, layer):
    """Attention.forward (modeling_finetune.py:145-188)."""
    B, N, C = x.shape
    qb = torch.cat((sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.v_bias"]), sd[p + "attn.v_bias"]))
    qkv = F.linear(x, sd[p + "attn.qkv.weight"], qb).reshape(B, N, 3, nh, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (q.shape[-1] ** -0.5)
    attn = q @ k.transpose(-2, -1)
    if bias is not None:
        attn = attn + bias
    attn = attn.softmax(dim=-1)
    attn = _attn_dropout(attn, noise, layer)
    x = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(x, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])


def block(sd, i, x, bias, arch: Arch, noise: Optional[Noise]):
    """Block.forward (modeling_finetune.py:290-299) -> (x, fc_feature)."""
    p = f"blocks.{i}."
    a = attention(sd, p, _ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], arch.ln_eps), bias, arch.num_heads,
                  noise, i)
    x = x + _dp(sd[p + "gamma_1"] * a, noise, i, 0)
    f = _dp(sd[p + "gamma_2"] * mlp(sd, p, _ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], arch.ln_eps)),
            noise, i, 1)
    return x + f, f


def wasserstein_distance_matmul(m1, c1, m2, c2):
    """uncertainty_evaluations.py:276-294."""
    m1, m2, c1, c2 = torch.sigmoid(m1), torch.sigmoid(m2), torch.sigmoid(c1), torch.sigmoid(c2)
    ret = -2 * m1 @ m2.transpose(-1, -2) + (m1 ** 2).sum(-1, keepdim=True) + (m2 ** 2).sum(-1, keepdim=True).transpose(-1, -2)
    u1 = torch.sqrt(torch.clamp(c1, min=1e-24))
    u2 = torch.sqrt(torch.clamp(c2, min=1e-24))
    cov = -2 * u1 @ u2.transpose(-1, -2) + c1.sum(-1, keepdim=True) + c2.sum(-1, keepdim=True).transpose(-1, -2)
    return ret + cov


def dist_attention(sd, p, x, cx, bias, nh, noise, layer):
    """dist Attention.forward (modeling_finetune_dist.py:111-179)."""
    B, N, C = x.shape
    z = torch.zeros_like(sd[p + "attn.v_bias"])
    qb = torch.cat((sd[p + "attn.q_bias"], z, sd[p + "attn.v_bias"]))
    cqb = torch.cat((sd[p + "attn.cov_q_bias"], z, sd[p + "attn.cov_v_bias"]))
    W = sd[p + "attn.qkv.weight"]                      # the cov stream re-uses the MEAN weight (:127)
    qkv = F.linear(x, W, qb).reshape(B, N, 3, nh, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    cqkv = (F.elu(F.linear(cx, W, cqb)) + 1).reshape(B, N, 3, nh, -1).permute(2, 0, 3, 1, 4)
    cq, ck, cv = cqkv[0], cqkv[1], cqkv[2]
    q = q * (q.shape[-1] ** -0.5)
    attn = torch.sigmoid(-wasserstein_distance_matmul(q, cq, k, ck) + 1e-24)
    attn = attn + bias
    attn = attn.softmax(dim=-1)
    attn = _attn_dropout(attn, noise, layer)
    om = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    oc = ((attn ** 2) @ cv).transpose(1, 2).reshape(B, N, -1)
    return (F.linear(om, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]),
            F.linear(oc, sd[p + "attn.cov_proj.weight"], sd[p + "attn.cov_proj.bias"]))


def dist_block(sd, i, xm, xc, bias, arch: Arch, noise: Optional[Noise]):
    """dist Block.forward (modeling_finetune_dist.py:41-59); 4 independent drop-path draws in this order."""
    p = f"blocks.{i}."
    n1 = lambda t: _ln(t, sd[p + "norm1.weight"], sd[p + "norm1.bias"], arch.ln_eps)
    n2 = lambda t: _ln(t, sd[p + "norm2.weight"], sd[p + "norm2.bias"], arch.ln_eps)
    m, c = dist_attention(sd, p, n1(xm), n1(xc), bias, arch.num_heads, noise, i)
    xm = xm + _dp(sd[p + "gamma_1"] * m, noise, i, 0)
    fm = _dp(sd[p + "gamma_2"] * mlp(sd, p, n2(xm)), noise, i, 1)
    xc = xc + _dp(sd[p + "gamma_1"] * c, noise, i, 2)
    fc = _dp(sd[p + "gamma_2"] * mlp(sd, p, n2(xc)), noise, i, 3)
    return xm + fm, xc + fc


# --------------------------------------------------------------------------------------------------------------
# model wrappers
# --------------------------------------------------------------------------------------------------------------
def patch_embed(sd, prefix, x, arch: Arch):
    """PatchEmbed.forward (modeling_finetune.py:319-325)."""
    y = F.conv2d(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"], stride=arch.patch_size)
    return y.flatten(2).transpose(1, 2)


def _stem(sd, x, mask, arch: Arch, cov: bool):
    pre = "cov_" if cov else ""
    t = patch_embed(sd, pre + "patch_embed.", x, arch)
    B = t.shape[0]
    if mask is not None:
        w = mask.reshape(B, -1, 1).to(t.dtype)
        t = t * (1 - w) + sd[pre + "mask_token"].expand(B, t.shape[1], -1) * w   # modeling_cyclical.py:179-182
    return torch.cat((sd[pre + "cls_token"].expand(B, -1, -1), t), dim=1)


def cyclical_forward(sd, arch: Arch, x, bool_masked_pos, return_all_tokens=False, layer_results=None,
                     noise: Optional[Noise] = None):
    """VisionTransformerForCyclicalTraining.forward (modeling_cyclical.py:170-225) and the dist variant
    (modeling_cyclical_dist.py:108-165)."""
    bias = rel_pos_bias(sd["rel_pos_bias.relative_position_bias_table"], sd["rel_pos_bias.relative_position_index"])
    if not arch.dist:
        t = _stem(sd, x, bool_masked_pos, arch, False)
        z = []
        for i in range(arch.depth):
            t, f = block(sd, i, t, bias, arch, noise)
            if layer_results == "end":
                z.append(t)
            elif layer_results == "fc":
                z.append(f)
        if layer_results:
            return [u[:, 1:] for u in z]
        t = _ln(t, sd["norm.weight"], sd["norm.bias"], arch.ln_eps)[:, 1:]
        if return_all_tokens:
            return F.linear(t, sd["lm_head.weight"], sd["lm_head.bias"])
        m = bool_masked_pos.flatten().bool()
        return F.linear(t.reshape(-1, t.shape[-1])[m], sd["lm_head.weight"], sd["lm_head.bias"])
    tm = _stem(sd, x, bool_masked_pos, arch, False)
    tc = _stem(sd, x, bool_masked_pos, arch, True)
    zm, zc = [], []
    for i in range(arch.depth):
        tm, tc = dist_block(sd, i, tm, tc, bias, arch, noise)
        if layer_results == "end":
            zm.append(tm)
            zc.append(tc)
    if layer_results:
        return [u[:, 1:] for u in zm], [u[:, 1:] for u in zc]
    tm = _ln(tm, sd["norm.weight"], sd["norm.bias"], arch.ln_eps)[:, 1:]
    tc = _ln(tc, sd["norm.weight"], sd["norm.bias"], arch.ln_eps)[:, 1:]
    if not return_all_tokens:
        m = bool_masked_pos.flatten().bool()
        tm = tm.reshape(-1, tm.shape[-1])[m]
        tc = tc.reshape(-1, tc.shape[-1])[m]
    return (F.linear(tm, sd["lm_head.weight"], sd["lm_head.bias"]),
            F.linear(tc, sd["cov_lm_head.weight"], sd["cov_lm_head.bias"]))


def finetune_forward(sd, arch: Arch, x, noise: Optional[Noise] = None):
    """VisionTransformer.forward (modeling_finetune.py:476-523, mean-pool head) /
    DistVisionTransformer.forward (modeling_finetune_dist.py:280-326)."""
    bias = rel_pos_bias(sd["rel_pos_bias.relative_position_bias_table"], sd["rel_pos_bias.relative_position_index"])
    fcn = lambda t: _ln(t, sd["fc_norm.weight"], sd["fc_norm.bias"], arch.ln_eps)
    if not arch.dist:
        t = _stem(sd, x, None, arch, False)
        for i in range(arch.depth):
            t, _ = block(sd, i, t, bias, arch, noise)
        return F.linear(fcn(t[:, 1:].mean(1)), sd["head.weight"], sd["head.bias"])
    tm = _stem(sd, x, None, arch, False)
    tc = _stem(sd, x, None, arch, True)
    for i in range(arch.depth):
        tm, tc = dist_block(sd, i, tm, tc, bias, arch, noise)
    fm, fc = fcn(tm[:, 1:].mean(1)), fcn(tc[:, 1:].mean(1))
    return fm, fc, F.linear(fm, sd["head.weight"], sd["head.bias"])


# --------------------------------------------------------------------------------------------------------------
# data2vec target builder, losses, EMA, optimiser   (engine_for_cyclical.py)
# --------------------------------------------------------------------------------------------------------------
def build_targets(layer_outputs: Sequence[torch.Tensor], target_layers: Sequence[int], bool_masked_pos,
                  target_layer_norm_last=True, target_batch_norm=False, target_instance_norm=False,
                  post_target_instance_norm=False, post_target_layer_norm=False):
    """engine_for_cyclical.py:90-122. layer_outputs: list of [B,T,C] (cls already dropped)."""
    fsz = layer_outputs[0].size(-1)
    vals = [layer_outputs[i] for i in target_layers]
    if target_instance_norm or target_batch_norm:
        vals = [v.permute(0, 2, 1) for v in vals]
    if target_batch_norm:
        vals = [F.batch_norm(v.float(), running_mean=None, running_var=None, training=True) for v in vals]
    if target_instance_norm:
        vals = [F.instance_norm(v.float()) for v in vals]
    if target_instance_norm or target_batch_norm:
        vals = [v.permute(0, 2, 1) for v in vals]
    if target_layer_norm_last:
        vals = [F.layer_norm(v.float(), (fsz,)) for v in vals]
    t = sum(vals) / len(target_layers)
    if post_target_instance_norm:
        t = F.instance_norm(t.permute(0, 2, 1).float()).permute(0, 2, 1)
    if post_target_layer_norm:
        t = F.layer_norm(t.float(), (fsz,))
    return t.reshape(-1, fsz)[bool_masked_pos.flatten().bool()]


def d2v_loss(outputs, targets, l1_beta=2.0, l2_loss=False):
    """engine_for_cyclical.py:145-150 (+ logged z0, :130-133)."""
    z0 = torch.sqrt(outputs.reshape(-1, outputs.size(-1)).var(dim=0) + 1e-6)
    loss = F.mse_loss(outputs, targets) if l2_loss else F.smooth_l1_loss(outputs, targets, beta=l1_beta)
    return loss, z0


def wasserstein_distance(m1, c1, m2, c2):
    """distloss.py:73-79."""
    ret = ((m1 - m2) ** 2).sum(-1)
    u1 = torch.sqrt(torch.clamp(c1, min=1e-24))
    u2 = torch.sqrt(torch.clamp(c2, min=1e-24))
    return ret + ((u1 - u2) ** 2).sum(-1)


def wasserstein_loss(mean_out, cov_out, pos_mean, pos_cov, lam=1e-5):
    """WassersteinLoss.forward (distloss.py:13-30)."""
    a, b, g, h = (torch.sigmoid(t) for t in (mean_out, cov_out, pos_mean, pos_cov))
    w = wasserstein_distance(a, b, g, h)
    w = w / torch.max(torch.abs(w))
    loss = -torch.log(torch.sigmoid(-w + 1e-24))
    loss = loss / torch.max(torch.abs(loss))
    return loss.sum() * lam


def wasserstein_loss_finetune(mo, co, pm, pc, nm, nc, lam_ft=1e-4, lam_pvn=1e-4):
    """WassersteinLossFineTuning.forward (distloss.py:39-70)."""
    mo, co, pm, pc, nm, nc = (torch.sigmoid(t) for t in (mo, co, pm, pc, nm, nc))
    pos = wasserstein_distance(mo, co, pm, pc)
    neg = wasserstein_distance(mo, co, nm, nc)
    pvn = wasserstein_distance(pm, pc, nm, nc)
    pos = pos / pos.abs().max()
    neg = neg / neg.abs().max()
    pvn = pvn / pvn.abs().max()
    loss = -torch.log(torch.sigmoid(neg - pos + 1e-24))
    loss = (loss / loss.abs().max() * lam_ft).sum()
    pl = torch.clamp(pos - pvn, 0)
    pl = (pl / pl.abs().max() * lam_pvn).sum()
    return loss + pl


def ema_decay_at(it: int, decay_init: float, decay: float, ema_start_at: int) -> float:
    """engine_for_cyclical.py:55-56."""
    if it < ema_start_at:
        return decay_init + it * (decay - decay_init) / ema_start_at
    return decay


def ema_update(ema: Dict[str, torch.Tensor], model: Dict[str, torch.Tensor], d: float) -> None:
    """engine_for_cyclical.py:182-185 + timm ModelEmaV2._update: e.copy_(d*e + (1-d)*m) over the state_dict, integer buffers included
    (int64 * python float -> fp32 arithmetic, truncated by copy_: relative_position_index entries can come back one lower)."""
    with torch.no_grad():
        for k in ema:
            ema[k].copy_(d * ema[k] + (1.0 - d) * model[k])


def clip_grad_norm(grads: Sequence[torch.Tensor], max_norm: float) -> Tuple[torch.Tensor, float]:
    """torch.nn.utils.clip_grad_norm_ semantics (utils.py:374-377): coef = min(1, max_norm/(norm+1e-6))."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = float(torch.clamp(max_norm / (total + 1e-6), max=1.0))
    return total, coef


def adamw_step(p, g, m, v, step: int, lr: float, wd: float, beta1=0.9, beta2=0.999, eps=1e-8) -> None:
    """torch.optim.AdamW single-tensor update (optim_factory.py:152-153 -> optim.AdamW)."""
    with torch.no_grad():
        p.mul_(1 - lr * wd)
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def get_num_layer_for_vit(name: str, num_max_layer: int) -> int:
    """optim_factory.py:33-44."""
    if name in ("cls_token", "mask_token", "pos_embed"):
        return 0
    if name.startswith("patch_embed"):
        return 0
    if name.startswith("rel_pos_bias"):
        return num_max_layer - 1
    if name.startswith("blocks"):
        return int(name.split(".")[1]) + 1
    return num_max_layer - 1


def is_no_decay(name: str, shape, skip=("pos_embed", "cls_token")) -> bool:
    """optim_factory.py:66-67: 1-D params, *.bias and the skip list get weight_decay 0."""
    return len(shape) == 1 or name.endswith(".bias") or name in skip


# --------------------------------------------------------------------------------------------------------------
# masking generator (masking_generator.py:29-92) with an injected python `random.Random`
# --------------------------------------------------------------------------------------------------------------
def blockwise_mask(rng, height=14, width=14, num_masking_patches=120, min_num_patches=16, max_num_patches=None,
                   min_aspect=0.3, max_aspect=None) -> np.ndarray:
    max_num_patches = num_masking_patches if max_num_patches is None else max_num_patches
    max_aspect = max_aspect or 1 / min_aspect
    log_ar = (math.log(min_aspect), math.log(max_aspect))
    mask = np.zeros((height, width), dtype=int)
    count = 0
    while count < num_masking_patches:
        max_mask = min(num_masking_patches - count, max_num_patches)
        delta = 0
        for _ in range(10):
            target_area = rng.uniform(min_num_patches, max_mask)
            ar = math.exp(rng.uniform(*log_ar))
            h = int(round(math.sqrt(target_area * ar)))
            w = int(round(math.sqrt(target_area / ar)))
            if w < width and h < height:
                top = rng.randint(0, height - h)
                left = rng.randint(0, width - w)
                num_masked = mask[top:top + h, left:left + w].sum()
                if 0 < h * w - num_masked <= max_mask:
                    for i in range(top, top + h):
                        for j in range(left, left + w):
                            if mask[i, j] == 0:
                                mask[i, j] = 1
                                delta += 1
                if delta > 0:
                    break
        if delta == 0:
            break
        count += delta
    return mask


# --------------------------------------------------------------------------------------------------------------
# Mixup / CutMix (timm.data.mixup, timm 0.3.2 pinned by the reference's requirements.txt:3; timm is NOT in this image, so this restates
# its published algorithm: PARITY UNPINNED against timm itself, anchored on the reference's call sites run_class_finetuning.py:339-347 and
# engine_for_finetuning.py:87-88). Plain torch CPU ops in timm's order, so the device kernel can be checked bit for bit.
# --------------------------------------------------------------------------------------------------------------
def mixup_images(x: torch.Tensor, lam: float, use_cutmix: bool, box) -> torch.Tensor:
    """Mixup._mix_batch on a copy of x."""
    x = x.clone()
    if lam == 1.0:
        return x
    if use_cutmix:
        yl, yh, xl, xh = box
        x[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
    else:
        x_flipped = x.flip(0).mul_(1.0 - lam)
        x.mul_(lam).add_(x_flipped)
    return x


def mixup_target(target: torch.Tensor, num_classes: int, lam: float = 1.0, smoothing: float = 0.0) -> torch.Tensor:
    off_value = smoothing / num_classes
    on_value = 1.0 - smoothing + off_value

    def one_hot(t):
        t = t.long().view(-1, 1)
        return torch.full((t.size()[0], num_classes), off_value).scatter_(1, t, on_value)
    return one_hot(target) * lam + one_hot(target.flip(0)) * (1.0 - lam)


def mixup_draw(rng, img_shape, mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, correct_lam=True):
    """Mixup._params_per_batch + cutmix_bbox_and_lam / rand_bbox for mode='batch' with numpy's RandomState call order."""
    lam, use_cutmix, box = 1.0, False, (0, 0, 0, 0)
    if rng.rand() < prob:
        if mixup_alpha > 0.0 and cutmix_alpha > 0.0:
            use_cutmix = bool(rng.rand() < switch_prob)
            lam = float(rng.beta(cutmix_alpha, cutmix_alpha) if use_cutmix else rng.beta(mixup_alpha, mixup_alpha))
        elif mixup_alpha > 0.0:
            lam = float(rng.beta(mixup_alpha, mixup_alpha))
        else:
            use_cutmix = True
            lam = float(rng.beta(cutmix_alpha, cutmix_alpha))
    if lam != 1.0 and use_cutmix:
        img_h, img_w = img_shape[-2:]
        ratio = np.sqrt(1 - lam)
        cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
        cy, cx = rng.randint(0, img_h), rng.randint(0, img_w)
        yl, yh = int(np.clip(cy - cut_h // 2, 0, img_h)), int(np.clip(cy + cut_h // 2, 0, img_h))
        xl, xh = int(np.clip(cx - cut_w // 2, 0, img_w)), int(np.clip(cx + cut_w // 2, 0, img_w))
        box = (yl, yh, xl, xh)
        if correct_lam:
            lam = 1.0 - (yh - yl) * (xh - xl) / float(img_h * img_w)
    return lam, use_cutmix, box


class InjectedUniformRng:
    """`random`-module stand-in that replays a given sequence of uniforms in [0,1): uniform(a, b) = a + (b - a) * u exactly as
    random.uniform computes it, randint(a, b) = a + min(b - a, floor(u * (b - a + 1))). Feeding the same sequence to the device
    generator (b200vit_block_masks, uniforms != NULL) must give bit-identical masks."""

    def __init__(self, uniforms):
        self.u = uniforms
        self.i = 0

    def random(self):
        v = float(self.u[self.i])
        self.i += 1
        return v

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def randint(self, a, b):
        n = b - a
        return a + min(n, int(self.random() * (n + 1)))


# --------------------------------------------------------------------------------------------------------------
# MC-sample uncertainty metrics (uncertainty_evaluations.py)
# --------------------------------------------------------------------------------------------------------------
def ece(probs: torch.Tensor, labels: torch.Tensor, n_bins: int = 15, reference_indexing: bool = False) -> float:
    """ECELoss.loss(probs, labels, logits=False) (uncertainty_evaluations.py:110-202): 15 uniform bins (lo,hi],
    sum_b prop_b * |mean conf_b - mean acc_b|.

    reference_indexing=True reproduces what the reference ACTUALLY computes with any torch whose
    Tensor.__array_wrap__ turns numpy bool into uint8 (all releases to date): at :173-176 `in_bin` becomes a uint8
    array, so `accuracies[in_bin]` at :184 is INTEGER fancy indexing (elements 0 and 1 of `accuracies`), i.e.
    bin_acc_b = ((N - n_b) * acc[0] + n_b * acc[1]) / N, while bin_conf_b (a torch mask-index at :185) is the true bin
    mean. The golden vector in tests/golden/metrics.pt pins this variant; the default is the documented ECE.
    """
    conf, pred = probs.max(dim=1)
    acc = pred.eq(labels).numpy().astype(np.float64)
    bounds = np.linspace(0, 1, n_bins + 1)
    total = 0.0
    confn = conf.numpy()
    n = confn.shape[0]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        in_bin = np.greater(confn, lo.item()) * np.less_equal(confn, hi.item())
        prop = np.mean(in_bin)
        if prop.item() > 0:
            if reference_indexing:
                nb = int(in_bin.sum())
                bin_acc = ((n - nb) * acc[0] + nb * acc[1]) / n
            else:
                bin_acc = np.mean(acc[in_bin])
            total += prop * np.abs(np.mean(confn[in_bin]) - bin_acc)
    return float(total)


def nll(logits: torch.Tensor, labels: torch.Tensor) -> float:
    """NLL (uncertainty_evaluations.py:270-272)."""
    return float(-torch.log_softmax(logits.float(), 1).gather(1, labels[:, None]).mean())


def accuracy_topk(logits: torch.Tensor, labels: torch.Tensor, topk=(1, 5)):
    """timm.utils.accuracy (uncertainty_evaluations.py:81)."""
    maxk = max(topk)
    _, pred = logits.topk(maxk, 1, True, True)
    correct = pred.t().eq(labels.reshape(1, -1).expand_as(pred.t()))
    return [float(correct[:k].reshape(-1).float().sum() * 100.0 / labels.shape[0]) for k in topk]


def tace(probs: torch.Tensor, labels: torch.Tensor, threshold: float = 0.01, n_bins: int = 30, reference_indexing: bool = False) -> float:
    """TACELoss.loss(probs, labels, logits=False) (uncertainty_evaluations.py:241-261): probabilities below `threshold` are zeroed, then per
    class the n_bins ADAPTIVE bin edges are sorted_p[i * int(N / n_bins)] (+ 1.0) (compute_bin_boundaries, :119-132) and
    sum_bins prop * |mean conf - mean acc| over bins (lo, hi] (compute_bins, :159-186), averaged over the classes.
    reference_indexing=True: as for ece(), compute_bins indexes the numpy `accuracies` with a uint8 array (:173-184), i.e.
    bin_acc = ((N - n_b) * acc[0] + n_b * acc[1]) / N — what the reference actually prints (golden in tests/golden/metrics.pt)."""
    p = probs.detach().clone().float().numpy()
    p[p < threshold] = 0
    lab = labels.numpy()
    n, k = p.shape
    bin_n = int(n / n_bins)
    total = 0.0
    for i in range(k):
        col = p[:, i]
        srt = np.sort(col)
        bounds = np.append(np.array([srt[j * bin_n] for j in range(n_bins)], dtype=np.float64), 1.0)
        acc = (lab == i).astype("float")
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            in_bin = np.greater(col, np.float32(lo)) * np.less_equal(col, np.float32(hi))
            prop = np.mean(in_bin)
            if prop > 0:
                nb = int(in_bin.sum())
                bin_acc = ((n - nb) * acc[0] + nb * acc[1]) / n if reference_indexing else np.mean(acc[in_bin])
                total += prop * np.abs(np.mean(col[in_bin]) - bin_acc)
    return float(total / k)


def auroc_macro_ovr(probs: torch.Tensor, labels: torch.Tensor) -> float:
    """torchmetrics AUROC(task="multiclass", num_classes=K) as called at uncertainty_evaluations.py:49,85 (defaults: average="macro",
    thresholds=None -> exact curve). torchmetrics is NOT in this image (requirements.txt lists it unpinned): PARITY UNPINNED against the
    package, restated from its published algorithm — softmax if the input is not already in [0,1], one-vs-rest ROC per class, area by the
    trapezoid rule (= Mann-Whitney U with half credit for ties), a class without positives or without negatives contributes 0, plain mean
    over the K classes. tests check it against sklearn.metrics.roc_auc_score(multi_class="ovr") where every class is present."""
    p = probs.double()
    if not bool(((p >= 0) & (p <= 1)).all()):
        p = torch.softmax(p, 1)
    n, k = p.shape
    out = []
    for c in range(k):
        pos = p[labels == c, c]
        neg = p[labels != c, c]
        if pos.numel() == 0 or neg.numel() == 0:
            out.append(0.0)
            continue
        less = (neg[None, :] < pos[:, None]).double().sum()
        tie = (neg[None, :] == pos[:, None]).double().sum()
        out.append(float((less + 0.5 * tie) / (pos.numel() * neg.numel())))
    return float(np.mean(out))


def gaussian_sample(mean: torch.Tensor, cov: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """The reparameterised draw the reference sketches (and leaves commented out) in DistVisionTransformer.forward
    (modeling_finetune_dist.py:314-324: Normal(mean, sqrt(cov)).sample()) with the noise injected: mean + sqrt(max(cov, 0)) * eps."""
    return mean + torch.sqrt(torch.clamp(cov, min=0.0)) * eps


def mc_reduce(logits_snk: torch.Tensor, labels: torch.Tensor) -> Dict[str, object]:
    """evaluate_MC_dropout reduction (uncertainty_evaluations.py:77-85): mean of LOGITS over S, then metrics.
    entropy / variance / mutual information are extensions (no reference function: parity unpinned)."""
    zbar = logits_snk.float().mean(0)
    probs = torch.softmax(zbar, 1)
    a1, a5 = accuracy_topk(zbar, labels, (1, min(5, zbar.shape[1])))
    ps = torch.softmax(logits_snk.float(), -1)
    pbar = ps.mean(0)
    ent = -(pbar * torch.log(pbar.clamp_min(1e-30))).sum(-1)
    ent_s = -(ps * torch.log(ps.clamp_min(1e-30))).sum(-1).mean(0)
    var = ((ps - pbar) ** 2).mean(0).sum(-1)
    return dict(mean_logits=zbar, acc1=a1, acc5=a5, ece=ece(probs, labels), ece_reference=ece(probs, labels, 15, True),
                nll=nll(zbar, labels),
                entropy=ent, variance=var, mutual_info=ent - ent_s, conf=probs.max(1).values, pred=probs.argmax(1))


# --------------------------------------------------------------------------------------------------------------
# one full data2vec optimisation step (engine_for_cyclical.py:45-186) — used by the engine parity test and as the
# CPU baseline ("port") that bench.py times on the host cores
# --------------------------------------------------------------------------------------------------------------
def d2v_step(sd: Dict[str, torch.Tensor], ema: Dict[str, torch.Tensor], opt: Dict[str, Dict[str, torch.Tensor]], arch: Arch, x, mask,
             step: int, noise: Optional[Noise] = None, target_layers=(6, 7, 8, 9, 10, 11), lr=2e-3, wd=0.05, clip=3.0, ema_decay=0.9998,
             l1_beta=2.0, return_grads: bool = False, lam: float = 1e-5, l2_loss: bool = False, target_kwargs: Optional[dict] = None,
             var_w0: float = 0.0, var_margin0: float = 0.5, loss_scale: float = -1, update_ema: bool = True, info: Optional[dict] = None):
    """sd / ema: float state dicts (updated IN PLACE); opt: {'m': {...}, 'v': {...}}. Returns (loss, grad_norm[, grads]).
    target_kwargs: the build_targets switches (target_layer_norm_last, target_batch_norm, target_instance_norm, post_target_instance_norm,
    post_target_layer_norm; default = the README recipe, post layer norm). var_w0 / var_margin0: the column-std hinge (:130-139);
    loss_scale (:162-163); update_ema=False: the `start_lr_decay_at_step` cut-off branch (:182-185). info (dict) receives z0 / std_loss0."""
    tk = dict(post_target_layer_norm=True) if target_kwargs is None else dict(target_kwargs)
    with torch.no_grad():
        t = cyclical_forward(ema, arch, x, None, return_all_tokens=True, layer_results="end")
    if arch.dist:
        t, tc = t
        # the cov targets only ever get the per-layer LayerNorm, the mean and the post LayerNorm: engine_for_cyclical.py:74-86
        ctgt = build_targets(tc, list(target_layers), mask, target_layer_norm_last=tk.get("target_layer_norm_last", True),
                             post_target_layer_norm=tk.get("post_target_layer_norm", False))
    tgt = build_targets(t, list(target_layers), mask, **tk)
    names = [k for k, v in sd.items() if v.is_floating_point()]
    leaf = {k: (sd[k].detach().clone().requires_grad_(True) if k in names else sd[k]) for k in sd}
    out = cyclical_forward(leaf, arch, x, mask, noise=noise)
    if arch.dist:
        out, cout = out
    loss, z0 = d2v_loss(out.float(), tgt, l1_beta, l2_loss)
    std_loss0 = torch.sum(F.relu(var_margin0 - z0)) / z0.size(0) if var_w0 > 0 else 0      # :136-139
    loss = loss + std_loss0 * var_w0
    if info is not None:
        info.update(z0=z0.detach(), std_loss0=float(std_loss0))
    if arch.dist:
        loss = loss + wasserstein_loss(out.float(), cout.float(), tgt, ctgt, lam)               # :152-158
    if loss_scale != -1:
        loss = loss * loss_scale                                                                  # :162-163
    loss.backward()
    grads = {k: (leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(sd[k])) for k in names}
    total, coef = clip_grad_norm(list(grads.values()), clip)
    for k in names:
        if leaf[k].grad is None:
            continue                      # torch.optim.AdamW skips parameters without a gradient (cov_qkv.weight, §A.2-1)
        decay = 0.0 if is_no_decay(k, sd[k].shape) else wd
        adamw_step(sd[k], grads[k] * coef, opt["m"][k], opt["v"][k], step, lr, decay)
    if update_ema:
        # over the WHOLE state dict, like ModelEmaV2._update: the int64 relative_position_index buffer goes through
        # `d * e + (1 - d) * m` in fp32 and is truncated back to int64 by copy_ — for many decays (0.9, and about half of the values on the
        # 0.999 -> 0.9998 anneal) some entries come back as k - 1, so the TEACHER's index table drifts. Pinned by tests/golden/tiny_*_loop.pt.
        ema_update(ema, sd, ema_decay)
    if return_grads:
        return float(loss), float(total), grads
    return float(loss), float(total)


def new_opt_state(sd):
    return {"m": {k: torch.zeros_like(v) for k, v in sd.items() if v.is_floating_point()},
            "v": {k: torch.zeros_like(v) for k, v in sd.items() if v.is_floating_point()}}


# --------------------------------------------------------------------------------------------------------------
# fine-tune criterion + one fine-tune step (engine_for_finetuning_dist.train_class_batch, :286-304; engine_for_finetuning.py:30-43)
# --------------------------------------------------------------------------------------------------------------
def soft_target_cross_entropy(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """timm SoftTargetCrossEntropy / LabelSmoothingCrossEntropy on already-soft targets (run_class_finetuning.py:617-624):
    mean_b sum_k -t_bk log_softmax(z)_bk."""
    return torch.sum(-targets * F.log_softmax(logits.float(), dim=-1), dim=-1).mean()


def smoothed_targets(labels: torch.Tensor, num_classes: int, smoothing: float) -> torch.Tensor:
    """LabelSmoothingCrossEntropy(smoothing) of timm 0.3.2 == soft-target CE with t = (1 - s) onehot + s / K."""
    t = torch.full((labels.shape[0], num_classes), smoothing / num_classes)
    return t.scatter_(1, labels.long().view(-1, 1), 1.0 - smoothing + smoothing / num_classes)


def finetune_loss_and_grads(sd: Dict[str, torch.Tensor], arch: Arch, x, targets, pos=None, neg=None, noise: Optional[Noise] = None,
                            lam_ft=1e-4, lam_pvn=1e-4):
    """train_class_batch: anchor forward in train mode (injected noise); for the dual-stream model the positive / negative forwards run on a
    deep copy in eval() mode (:293-296), i.e. deterministic and without a gradient path into the model. Returns (loss, logits, grads)."""
    names = [k for k, v in sd.items() if v.is_floating_point()]
    leaf = {k: (sd[k].detach().clone().requires_grad_(True) if k in names else sd[k]) for k in sd}
    out = finetune_forward(leaf, arch, x, noise)
    if arch.dist:
        fm, fc, logits = out
    else:
        logits = out
    loss = soft_target_cross_entropy(logits, targets)
    if arch.dist and pos is not None:
        with torch.no_grad():
            pm, pc, _ = finetune_forward(sd, arch, pos, None)
            nm, nc, _ = finetune_forward(sd, arch, neg, None)
        loss = loss + wasserstein_loss_finetune(fm, fc, pm, pc, nm, nc, lam_ft, lam_pvn)
    loss.backward()
    grads = {k: (leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(sd[k])) for k in names}
    return float(loss), logits.detach(), grads
