"""MC-sample uncertainty evaluation (evaluate_MC_dropout, uncertainty_evaluations.py:42-89) on the B200 path.

S stochastic forward passes (model.eval() + enable_dropout: only Dropout modules are re-enabled, DropPath is not), logits stacked
[S, N, K], reduced on device by b200vit_mc_reduce (mean of LOGITS over S, then softmax metrics). With several ranks the S passes are
sharded contiguously across ranks (every rank sees every image) and the logits are all-gathered once."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import ops


def enable_dropout(model) -> None:
    """uncertainty_evaluations.py:35-39."""
    for m in model.modules():
        if m.__class__.__name__.startswith("Dropout"):
            m.train()


def shard_passes(S: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) of the passes each rank runs: contiguous, sizes differ by at most one, every pass exactly once."""
    base, extra = divmod(S, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def gather_passes(local: torch.Tensor, S: int, rank: int, world: int, group=None) -> torch.Tensor:
    """local: [s_local, N, K] of this rank -> [S, N, K] on every rank (all_gather of equal-sized padded shards)."""
    if world == 1:
        return local
    import torch.distributed as dist
    shards = shard_passes(S, world)
    smax = max(e - s for s, e in shards)
    pad = torch.zeros((smax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: e - s] for r, (s, e) in enumerate(shards)], 0)


@torch.no_grad()
def evaluate_mc_dropout(model, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]], forward_passes: int, rank: int = 0, world: int = 1,
                        group=None) -> Dict[str, object]:
    """batches: list of (images [b,3,H,W] on the device, labels [b]). Returns the reference's metrics (+ entropy / variance / MI)."""
    if forward_passes < 2:
        raise ValueError("the reference captures the labels at pass i == 1, so it needs forward_passes >= 2 (uncertainty_evaluations.py:69-70)")
    s0, s1 = shard_passes(forward_passes, world)[rank]
    outs = []
    for _ in range(s0, s1):
        model.eval()
        enable_dropout(model)
        # the dual-stream (--stochastic) model returns (mean_feat, cov_feat, logits): modeling_finetune_dist.py:311-326
        fw = [model(x) for x, _ in batches]
        outs.append(torch.cat([(o[-1] if isinstance(o, (tuple, list)) else o).float() for o in fw], 0))
    dev = batches[0][0].device
    K = model.cfg.num_classes
    N = sum(x.shape[0] for x, _ in batches)
    local = torch.stack(outs) if outs else torch.empty((0, N, K), dtype=torch.float32, device=dev)
    logits = gather_passes(local.contiguous(), forward_passes, rank, world, group)
    labels = torch.cat([y for _, y in batches]).to(dev).to(torch.int32)
    mean_logits, row_stats, hist, summary = ops.mc_reduce(logits.contiguous(), labels)
    s = summary.tolist()
    return dict(acc1=s[0], acc5=s[1], ece=s[2], ece_reference=s[3], nll=s[4], entropy=s[5], variance=s[6], mutual_info=s[7],
                mean_logits=mean_logits, row_stats=row_stats, hist=hist)
