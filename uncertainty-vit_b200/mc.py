"""MC-sample uncertainty evaluation (evaluate_MC_dropout, uncertainty_evaluations.py:42-89) on the B200 path.

S stochastic forward passes (model.eval() + enable_dropout: only Dropout modules are re-enabled, DropPath is not) over N images are S x N
independent (pass, image) rows. They are sharded sample x batch: the IMAGES are split contiguously across the ranks (an even split whatever
S is: 30 passes over 8 ranks would be 4,4,4,4,4,4,3,3) and every rank batches the passes of its images into as few forwards as fit
(`max_rows` images per forward: the dropout key is per (forward, row), so replicas of one image in a batch draw different masks and the
GEMMs see one large M instead of S small ones). Each rank reduces ITS images over all S passes on the device (b200vit_mc_reduce: mean of
LOGITS over S, then softmax metrics); only the per-image statistics [N, 8], the mean logits [N, K] and the 15-bin histogram cross NVLink."""
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


class _GraphedForward:
    """One stochastic forward pass of `model` on a fixed batch shape, replayed from a CUDA graph: ~200 launches whose Python issue time
    (~55 us each) exceeds their run time at eval batch sizes. The per-pass randomness (attention dropout) is keyed from device memory
    (`Noise.seed_dev`), so the captured graph is valid for every pass; it is re-captured when any parameter changes."""

    def __init__(self, model, x):
        from .core import Noise
        self.model = model
        self.sig = self._signature(model)
        calls0 = model._seed_calls                 # warm-up and capture must not advance the model's noise counter
        for _ in range(2):                         # lazy kernel attributes / bf16 weight shadows / allocator warm-up
            model(x)
        self.x = x.clone()
        self.seed = torch.zeros(1, dtype=torch.int64, device=x.device)
        self.graph = torch.cuda.CUDAGraph()
        model._injected = Noise(seed_dev=self.seed)
        torch.cuda.synchronize(x.device)
        with torch.cuda.graph(self.graph):
            self.out = model(self.x)
        model._seed_calls = calls0

    @staticmethod
    def _signature(model):
        return (getattr(model, "_weights_version", 0),) + tuple((p._version, p.data_ptr()) for p in model.parameters())

    def valid_for(self, model, x):
        return tuple(x.shape) == tuple(self.x.shape) and x.dtype == self.x.dtype and self._signature(model) == self.sig

    def __call__(self, x):
        self.model._seed_calls += 1                # the same key sequence as the eager forward (_VitBase._noise)
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * self.model._seed_calls) & 0xFFFFFFFFFFFFFFFF
        self.x.copy_(x, non_blocking=True)
        self.seed.fill_(seed - (1 << 64) if seed >= (1 << 63) else seed)
        self.graph.replay()
        return self.out


def _forward_pass(model, x, use_graph: bool):
    if not use_graph:
        return model(x)
    cache = model.__dict__.setdefault("_mc_graphs", {})
    key = tuple(x.shape)
    g = cache.get(key)
    if g is None or not g.valid_for(model, x):
        g = cache[key] = _GraphedForward(model, x)
    return g(x)


def shard_images(N: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) of the images each rank evaluates: contiguous, sizes differ by at most one."""
    return shard_passes(N, world)


def gather_rows(local: torch.Tensor, N: int, rank: int, world: int, group=None) -> torch.Tensor:
    """local: [n_local, ...] of this rank's images -> [N, ...] on every rank, in image order (all_gather of equal-sized padded shards)."""
    if world == 1:
        return local
    import torch.distributed as dist
    shards = shard_images(N, world)
    nmax = max(e - s for s, e in shards)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: e - s] for r, (s, e) in enumerate(shards)], 0)


@torch.no_grad()
def evaluate_mc_dropout(model, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]], forward_passes: int, rank: int = 0, world: int = 1,
                        group=None, use_graph: bool = True, max_rows: int = 1536) -> Dict[str, object]:
    """batches: list of (images [b,3,H,W] on the device, labels [b]). Returns the reference's metrics (+ entropy / variance / MI)."""
    if forward_passes < 2:
        raise ValueError("the reference captures the labels at pass i == 1, so it needs forward_passes >= 2 (uncertainty_evaluations.py:69-70)")
    S = forward_passes
    dev = batches[0][0].device
    K = model.cfg.num_classes
    images = torch.cat([x for x, _ in batches], 0) if len(batches) > 1 else batches[0][0]
    labels = torch.cat([y for _, y in batches]).to(dev).to(torch.int32)
    N = images.shape[0]
    n0, n1 = shard_images(N, world)[rank]
    nloc = n1 - n0
    model.eval()
    enable_dropout(model)
    outs = []
    if nloc > 0:
        local = images[n0:n1]
        per_fwd = max(1, min(S, max_rows // nloc))       # passes batched into one forward
        s = 0
        while s < S:
            p = min(per_fwd, S - s)
            xb = local.repeat(p, 1, 1, 1) if p > 1 else local
            o = _forward_pass(model, xb.contiguous(), use_graph)
            # the dual-stream (--stochastic) model returns (mean_feat, cov_feat, logits): modeling_finetune_dist.py:311-326
            outs.append((o[-1] if isinstance(o, (tuple, list)) else o).float().reshape(p, nloc, K).clone())
            s += p
        local_logits = torch.cat(outs, 0).contiguous()       # [S, nloc, K]
        mean_l, stats_l, hist, _ = ops.mc_reduce(local_logits, labels[n0:n1].contiguous(), finalize=False)
    else:
        mean_l = torch.zeros((0, K), dtype=torch.float32, device=dev)
        stats_l = torch.zeros((0, 8), dtype=torch.float32, device=dev)
        hist = torch.zeros((15, 3), dtype=torch.float32, device=dev)
    mean_logits = gather_rows(mean_l, N, rank, world, group)
    row_stats = gather_rows(stats_l, N, rank, world, group).contiguous()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(hist, group=group)
    summary = ops.mc_finalize(row_stats, hist, N)
    s_ = summary.tolist()
    res = dict(acc1=s_[0], acc5=s_[1], ece=s_[2], ece_reference=s_[3], nll=s_[4], entropy=s_[5], variance=s_[6], mutual_info=s_[7],
               mean_logits=mean_logits, row_stats=row_stats, hist=hist)
    if N >= 30:      # TACE (30 adaptive bins per class) and AUROC of the printout, on the mean logits (uncertainty_evaluations.py:83,85)
        t = ops.tace_auroc(mean_logits.contiguous(), labels).tolist()
        res.update(tace=t[0], tace_reference=t[1], auroc=t[2])
    return res
