#!/usr/bin/env python
"""BASELINE.json configs[4]: beit_large_patch16_224 --stochastic fine-tune TRAIN step (layer_decay 0.65, drop_path 0.2), batch 64 per GPU, bf16.

    python tools/bench_finetune.py [--steps K] [--warmup W] [--batch 64]            (torchrun for N > 1: DDP-style gradient all-reduce)

One step = anchor forward + backward through the dual-stream classifier (CUDA forward/backward schedules of modeling_dist.py behind ONE
torch.autograd.Function), the positive / negative triplet forwards without gradients (engine_for_finetuning_dist.py:286-304 computes them on a
per-batch deepcopy; the gradients are identical), soft-target cross-entropy (Mixup / CutMix targets, run_class_finetuning.py:339-347) +
WassersteinLossFineTuning (distloss.py:39-70) on the [B, C] features, and AdamW over the reference's layer-decay parameter groups
(optim_factory.py:33-97: lr_scale = 0.65^(25 - layer_id)). --path engine (default) runs the fused flat-arena
FinetuneEngine (engine.py): forward / backward schedules without autograd, gradients in one arena, ONE all-reduce, clip + layer-decay AdamW in
one kernel; --path module runs the nn.Module boundary with torch.optim.AdamW. The [B, 1000] / [B, C] loss terms are a few tiny torch ops."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default="beit_large_patch16_224")
    ap.add_argument("--path", default="engine", choices=["engine", "module"],
                    help="engine: fused flat-arena FinetuneEngine (default); module: nn.Module boundary + torch.optim.AdamW")
    ap.add_argument("--mixup", action="store_true",
                    help="README recipe --mixup 0.8 --cutmix 1.0 --smoothing 0.1: device Mixup / CutMix of a fresh copy of the batch inside every step")
    args = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    import uncertainty_vit_b200  # noqa: F401
    from uncertainty_vit_b200 import engine as E, modeling as M
    torch.manual_seed(rank)
    model = M.create_model(args.model, pretrained=False, stochastic=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.2, attn_drop_rate=0.0,
                           use_mean_pooling=True, init_scale=0.001, use_rel_pos_bias=False, use_shared_rel_pos_bias=True, use_abs_pos_emb=False,
                           init_values=0.1).to(dev)
    if args.path == "engine":
        if world > 1:
            for p in model.parameters():
                dist.broadcast(p.data, 0)
        eng = E.FinetuneEngine(model, lr=5e-4, weight_decay=0.05, layer_decay=0.65, clip_grad=3.0, lambda_finetuning=1e-2, lambda_pvn=1e-4,
                               world_size=world, seed=rank)
    L = model.get_num_layers() + 2
    groups = {}
    for name, p in model.named_parameters():          # optim_factory.get_parameter_groups with LayerDecayValueAssigner(0.65)
        if not p.requires_grad:
            continue
        lid = E.get_num_layer_for_vit(name, L)
        no_decay = p.ndim == 1 or name.endswith(".bias") or name in model.no_weight_decay()
        g = groups.setdefault((lid, no_decay), {"params": [], "weight_decay": 0.0 if no_decay else 0.05, "lr": 5e-4 * 0.65 ** (L - 1 - lid)})
        g["params"].append(p)
    opt = torch.optim.AdamW(list(groups.values()), lr=5e-4, betas=(0.9, 0.999), eps=1e-8, fused=True)
    B = args.batch
    g = torch.Generator().manual_seed(100 + rank)
    x = [torch.randn(B, 3, 224, 224, generator=g).to(dev) for _ in range(3)]            # anchor, positive, negative
    tgt = torch.softmax(torch.randn(B, 1000, generator=g) * 4, -1).to(dev)                # soft (mixup-like) targets

    def wloss_ft(am, ac, pm, pc, nm, nc, lam=1e-2):
        def w2(m1, c1, m2, c2):
            a, b, g_, h = (torch.sigmoid(t.float()) for t in (m1, c1, m2, c2))
            return ((a - g_) ** 2).sum(-1) + ((torch.sqrt(b.clamp_min(1e-24)) - torch.sqrt(h.clamp_min(1e-24))) ** 2).sum(-1)
        pos, neg = w2(am, ac, pm, pc), w2(am, ac, nm, nc)
        l = -torch.log(torch.sigmoid(neg - pos + 1e-24))
        return lam * (l / l.abs().max().clamp_min(1e-30)).sum()

    labels = torch.randint(0, 1000, (B,), generator=g).to(dev)
    mixup_fn = None
    if args.mixup:
        import numpy as np
        from uncertainty_vit_b200.mixup import Mixup
        mixup_fn = Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, label_smoothing=0.1, num_classes=1000,
                         rng=np.random.RandomState(rank))

    def step():
        if args.path == "engine":
            if mixup_fn is not None:        # the loader hands over a fresh batch every step: mix a copy, as Mixup works in place
                xa, t = mixup_fn(x[0].clone(), labels)
                return eng.step(xa, t, x[1], x[2])
            return eng.step(x[0], tgt, x[1], x[2])
        model.train()
        am, ac, logits = model(x[0])
        with torch.no_grad():
            pm, pc, _ = model(x[1])
            nm, nc, _ = model(x[2])
        loss = torch.sum(-tgt * torch.log_softmax(logits.float(), -1), -1).mean() + wloss_ft(am, ac, pm, pc, nm, nc)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            grads = [p.grad for p in model.parameters() if p.grad is not None]
            flat = torch.cat([t.reshape(-1) for t in grads])
            dist.all_reduce(flat)
            flat /= world
            off = 0
            for t in grads:
                t.copy_(flat[off: off + t.numel()].view_as(t)); off += t.numel()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 3.0)
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms.item())
        print(json.dumps({"metric": "ViT-L/16 --stochastic fine-tune train throughput", "value": world * B / (ms * 1e-3), "unit": "img/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": f"{args.model} --stochastic fine-tune train step (anchor fwd+bwd, pos/neg no-grad fwd, soft-target CE + "
                                                 "WassersteinLossFineTuning, AdamW with layer_decay 0.65 groups), drop_path 0.2, batch 64/GPU",
                                     "path": "fused flat-arena FinetuneEngine (clip + layer-decay AdamW in one kernel)" if args.path == "engine"
                                     else "nn.Module boundary + torch.optim.AdamW(fused)", "param_groups": len(groups), "mixup": bool(args.mixup)},
                          "final_loss": float(loss.item()), "params_M": sum(p.numel() for p in model.parameters()) / 1e6}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
