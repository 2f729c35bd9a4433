#!/usr/bin/env python
"""Headline benchmark: ViT-B/16 data2vec pre-training step (BASELINE.json configs[1]) in img/s.

    python bench.py --gpus N --steps K --warmup W            # B200 path (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm (CPU oracle port) on the host cores

One "step" = EMA-teacher forward + target builder / smooth-L1 + student forward/backward + [NCCL grad all-reduce] + clip + AdamW +
EMA on one batch of 128 synthetic 224x224 images per GPU with exactly 120 masked patches per image (README recipe:
drop_path 0.25, attn_drop 0.05, layer-scale 1e-4, target_layers 6-11, post LayerNorm targets, l1_beta 2, EMA 0.9998, clip 3, wd 0.05).
`value`   : device-timed (CUDA events) with the batch already resident in HBM.
`e2e`     : the same step through D2VEngine.stage_host()/launch_staged() (what engine.train_one_epoch calls): every step's pinned host batch is
            copied to the device (one batch ahead, on a copy stream) and its loss read back inside the timed region.
`roofline`: all launches of the tcgen05 GEMM kernel inside instrumented steps: algorithmic FLOPs / CUDA-event time vs measured bf16 peak.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 128
MASKED = 120
CPU_SAMPLE_BATCH = 8


def synth_batch(B, seed, pin=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, 224, 224, generator=g)
    mask = np.zeros((B, 196), dtype=np.uint8)
    for b in range(B):
        mask[b, torch.randperm(196, generator=g)[:MASKED].numpy()] = 1
    return (x.pin_memory() if pin else x), mask.reshape(B, 14, 14)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t_begin = self.t_end = None          # host-time window of the timed regions (value + e2e)

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if self.t_begin is not None and (ts < self.t_begin or (self.t_end is not None and ts > self.t_end + 0.05)):
                continue                       # keep only samples taken while the timed steps were running
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hi = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference algorithm; the only places that execute oracle/)
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, batch=CPU_SAMPLE_BATCH, stochastic=False):
    from oracle import vit_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    arch = O.Arch(kind="cyclical", dist=stochastic, **O.VIT_B)
    sd = O.make_state(arch, 0)
    ema = {k: v.clone() for k, v in sd.items()}
    opt = O.new_opt_state(sd)
    x, mask = synth_batch(batch, 0, pin=False)
    mask_t = torch.from_numpy(mask.astype(np.int64))
    g = torch.Generator().manual_seed(1)
    probs = [float(p) for p in torch.linspace(0, 0.25, arch.depth)]
    noise = O.Noise(drop_path_keep=[(torch.rand(4 if stochastic else 2, batch, generator=g) >= p).float() for p in probs], drop_path_prob=probs,
                    attn_keep=[(torch.rand(batch, arch.num_heads, arch.tokens, arch.tokens, generator=g) >= 0.05).float() for _ in range(arch.depth)],
                    attn_drop=0.05)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.d2v_step(sd, ema, opt, arch, x, mask_t, i + 1, noise)
        times.append(time.perf_counter() - t0)
    t = times[warmup:]
    return batch * len(t) / sum(t), sum(t) / len(t), torch.get_num_threads()


def workload_config(world, stochastic=False):
    return {"workload": ("beit_base_patch16_224 --stochastic (dual-stream mean/cov + Wasserstein loss) " if stochastic else "beit_base_patch16_224 ")
                        + "data2vec cyclical pretrain step (run_cyclical.py recipe), batch 128/GPU, 120 masked patches, "
                        "target_layers 6-11, EMA 0.9998, bf16 GEMMs / fp32 master weights", "global_batch": world * BATCH,
            "parallelism": f"dp{world}", "l2": "working set per step (>9 GB of activations) is far larger than the 126 MB L2; two alternating input batches"}


def run_reference(args):
    """The reference algorithm on the host cores: exactly --steps timed steps after --warmup untimed ones, each step one full data2vec
    optimisation step of the same workload on a BOUNDED sample of CPU_SAMPLE_BATCH images (a 128-image step takes ~11 s on 16 threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, sec, cores = cpu_reference_steps(args.steps, args.warmup, stochastic=args.stochastic)
    sample = (f"{args.steps} timed + {args.warmup} warm-up full data2vec steps, each on a bounded sample of {CPU_SAMPLE_BATCH} images of the workload "
              f"(fp32 oracle port of the reference algorithm: the reference has no packaging and cannot be launched, DESIGN.md section 9), {cores} threads")
    print(json.dumps({"impl": "reference", "metric": "data2vec ViT-B/16 pretrain throughput", "value": v, "unit": "img/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus, args.stochastic),
                      "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
MC_PASSES, MC_BATCH = 30, 192          # BASELINE.json configs[3]: S = 30 passes; eval batch int(1.5 * 128) (run_class_finetuning.py:329-331)


def run_mc_inference(dev, rank, world, M, barrier, max_over_ranks):
    """Second half of the BASELINE metric ("MC-inference img/s"): evaluate_MC_dropout (uncertainty_evaluations.py:42-89) — S stochastic
    passes (eval mode, Dropout re-enabled: attn_drop 0.05) over one eval batch of 192 synthetic images, the S passes sharded across ranks,
    logits all-gathered and reduced on device (mean logits, acc@1/5, 15-bin ECE, NLL, entropy / variance / mutual information).
    Timed with CUDA events, max over ranks; `img_per_s` = images whose S-pass predictive statistics are produced per second (whole job),
    `forward_img_per_s` = S * images / time."""
    import uncertainty_vit_b200 as pkg
    from uncertainty_vit_b200 import mc
    g = torch.Generator().manual_seed(7)
    x = torch.randn(MC_BATCH, 3, 224, 224, generator=g).to(dev)
    y = torch.randint(0, 1000, (MC_BATCH,), generator=g).to(dev)
    out = {}
    for key, kwargs in (("det", {}), ("stochastic", {"stochastic": True})):
        torch.manual_seed(3)
        model = M.create_model("beit_base_patch16_224", pretrained=False, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1, attn_drop_rate=0.05,
                               use_mean_pooling=True, init_scale=0.001, use_rel_pos_bias=False, use_shared_rel_pos_bias=True, use_abs_pos_emb=False,
                               init_values=0.1, **kwargs).to(dev)
        if world > 1:
            import torch.distributed as dist
            for p in model.parameters():
                dist.broadcast(p.data, 0)
        res = mc.evaluate_mc_dropout(model, [(x, y)], MC_PASSES, rank, world)        # warm-up (also allocates)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2
        e0.record()
        for _ in range(reps):
            res = mc.evaluate_mc_dropout(model, [(x, y)], MC_PASSES, rank, world)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        fwd = MC_PASSES * MC_BATCH / (ms * 1e-3)
        gf = 70.25e9 if key == "stochastic" else 35.27e9                       # algorithmic forward FLOPs per image
        pk = peaks()
        out[key] = {"img_per_s": MC_BATCH / (ms * 1e-3), "forward_img_per_s": fwd, "ms_per_eval": ms,
                    "ece": res["ece"], "nll": res["nll"], "entropy": res["entropy"], "tace": res.get("tace"), "auroc": res.get("auroc"),
                    "roofline": {"bound": "tensor", "achieved": fwd * gf / 1e12 / world, "peak": pk["bf16_sustained"], "unit": "TFLOP/s per GPU",
                                 "frac": fwd * gf / 1e12 / world / pk["bf16_sustained"], "flops_per_image": gf}}
        del model
    out["config"] = {"workload": "beit_base_patch16_224 fine-tune model, MC-sample uncertainty eval", "passes": MC_PASSES, "eval_batch": MC_BATCH,
                     "attn_drop": 0.05, "sharding": f"sample x batch: {MC_BATCH} images split over {world} rank(s), the {MC_PASSES} passes of a rank batched into forwards of <= 1536 rows; per-image statistics gathered", "scaling": "strong"}
    return out


def timed_steps(step_fn, steps, warmup, barrier, max_over_ranks):
    for i in range(warmup):
        step_fn(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = step_fn(i)
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / steps, out


def run_stochastic_block(dev, rank, world, E, M, barrier, max_over_ranks, steps):
    """BASELINE.json configs[2]: the --stochastic (dual-stream mean / cov + WassersteinLoss) data2vec step at the same N, same recipe."""
    torch.manual_seed(0 + rank)
    model = M.create_model("beit_base_patch16_224", pretrained=False, drop_path_rate=0.25, drop_rate=0.0, use_shared_rel_pos_bias=True,
                           use_abs_pos_emb=False, init_values=1e-4, attn_drop_rate=0.05, stochastic=True).to(dev)
    eng = E.D2VEngine(model, lr=2e-3, weight_decay=0.05, clip_grad=3.0, ema_decay=0.9998, ema_decay_init=0.999, ema_start_at=0,
                      target_layers=[6, 7, 8, 9, 10, 11], l1_beta=2.0, post_target_layer_norm=True, world_size=world, seed=rank)
    batches = []
    for i in range(2):
        x, m = synth_batch(BATCH, 300 * rank + i, pin=False)
        mu8 = np.ascontiguousarray(m.reshape(BATCH, -1))
        batches.append((x.to(dev), torch.from_numpy(mu8.reshape(-1)).to(dev), torch.from_numpy(eng.rows_from_host_mask(mu8, 197)).to(dev)))
    ms, loss = timed_steps(lambda i: eng.step(*batches[i % 2], lr=1e-3), steps, 3, barrier, max_over_ranks)
    pk = peaks()
    tf = 281.86e9 * BATCH / (ms * 1e-3) / 1e12
    out = {"value": world * BATCH / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms, "steps": steps, "final_loss": float(loss.item()),
           "config": workload_config(world, True),
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s per GPU (whole step, algorithmic 281.86 GFLOP/image)",
                        "frac": tf / pk["bf16_sustained"]}}
    del eng, model, batches
    torch.cuda.empty_cache()
    return out


FT_BATCH = 64


def run_finetune_large_block(dev, rank, world, E, M, barrier, max_over_ranks, steps):
    """BASELINE.json configs[4]: beit_large_patch16_224 --stochastic fine-tune TRAIN step (layer_decay 0.65, drop_path 0.2), batch 64 per GPU:
    anchor forward + backward, eval-mode positive / negative forwards, soft-target CE + WassersteinLossFineTuning (one device kernel),
    gradient all-reduce, fused clip + layer-decay AdamW; forward + loss + backward replayed from a CUDA graph."""
    torch.manual_seed(0 + rank)
    model = M.create_model("beit_large_patch16_224", pretrained=False, stochastic=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.2,
                           attn_drop_rate=0.0, use_mean_pooling=True, init_scale=0.001, use_rel_pos_bias=False, use_shared_rel_pos_bias=True,
                           use_abs_pos_emb=False, init_values=0.1).to(dev)
    eng = E.FinetuneEngine(model, lr=5e-4, weight_decay=0.05, layer_decay=0.65, clip_grad=3.0, lambda_finetuning=1e-2, lambda_pvn=1e-4,
                           world_size=world, seed=rank)
    g = torch.Generator().manual_seed(100 + rank)
    x = [torch.randn(FT_BATCH, 3, 224, 224, generator=g).to(dev) for _ in range(3)]
    tgt = torch.softmax(torch.randn(FT_BATCH, 1000, generator=g) * 4, -1).to(dev)
    ms, loss = timed_steps(lambda i: eng.step(x[0], tgt, x[1], x[2]), steps, 3, barrier, max_over_ranks)
    pk = peaks()
    tf = 1231e9 * FT_BATCH / (ms * 1e-3) / 1e12
    out = {"value": world * FT_BATCH / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms, "steps": steps, "final_loss": float(loss.item()),
           "graphs": len(eng._ft_graphs),
           "config": {"workload": "beit_large_patch16_224 --stochastic fine-tune train step (anchor fwd+bwd, eval-mode pos/neg fwd, soft-target CE + "
                                  "WassersteinLossFineTuning, AdamW with layer_decay 0.65 groups), drop_path 0.2, batch 64/GPU, bf16",
                      "global_batch": world * FT_BATCH, "parallelism": f"dp{world}"},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s per GPU (whole step, algorithmic 1231 GFLOP/image)",
                        "frac": tf / pk["bf16_sustained"]}}
    del eng, model, x
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_fd = None
    if world > 1:
        # NCCL writes its version banner to file descriptor 1: stdout must carry ONE JSON line, so fd 1 is pointed at stderr for the
        # whole run and the JSON line is written through a duplicate of the original stdout.
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev, timeout=__import__("datetime").timedelta(seconds=120))
    import uncertainty_vit_b200 as pkg
    from uncertainty_vit_b200 import engine as E, modeling as M
    ops = pkg.ops

    torch.manual_seed(0 + rank)                                   # run_cyclical.py:315 seed + rank
    extra = {"stochastic": True} if args.stochastic else {}      # --stochastic: the dual-stream (mean / cov) model + Wasserstein loss
    model = M.create_model("beit_base_patch16_224", pretrained=False, drop_path_rate=0.25, drop_rate=0.0, use_shared_rel_pos_bias=True,
                           use_abs_pos_emb=False, init_values=1e-4, attn_drop_rate=0.05, **extra).to(dev)
    eng = E.D2VEngine(model, lr=2e-3, weight_decay=0.05, clip_grad=3.0, ema_decay=0.9998, ema_decay_init=0.999, ema_start_at=0,
                      target_layers=[6, 7, 8, 9, 10, 11], l1_beta=2.0, post_target_layer_norm=True, world_size=world, seed=rank)
    lr_sched = E.cosine_scheduler(2e-3, 1e-5, 800, 10, warmup_epochs=10)       # per-step lr as the runner computes it
    host = [synth_batch(BATCH, 100 * rank + i) for i in range(2)]
    dev_batches = []
    for x, m in host:
        mu8 = np.ascontiguousarray(m.reshape(BATCH, -1))
        dev_batches.append((x.to(dev), torch.from_numpy(mu8.reshape(-1)).to(dev), torch.from_numpy(eng.rows_from_host_mask(mu8, 197)).to(dev)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lr_at = lambda i: float(lr_sched[min(i + 50, len(lr_sched) - 1)])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        eng.step(*dev_batches[i % 2], lr=lr_at(i))
    barrier()
    # ---- value: K steps, inputs resident in HBM, CUDA events
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for i in range(args.steps):
        loss = eng.step(*dev_batches[i % 2], lr=lr_at(i))
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = (ops.LAUNCHES - launches0)
    final_loss = float(loss.item())
    z0_mean = float(eng.z0_dev.mean().item()) if eng.z0_dev is not None else None
    # ---- e2e: pinned host batch -> H2D -> step -> loss read-back, wall clock between device syncs. Every step's inputs are copied from
    # pinned host memory inside the timed region; as in engine.train_one_epoch the copy of batch i+1 is enqueued on the copy stream
    # before step i is launched (a one-batch look-ahead data loader), and the loss of every step is read back (its 4-byte copy is enqueued
    # behind the step and awaited after the NEXT step has been enqueued, D2VEngine.read_loss_async — as train_one_epoch does).
    def e2e_loop(n):
        nxt = eng.stage_host(*host[0])
        pending = None
        for i in range(n):
            loss_dev = eng.launch_staged(nxt, lr=lr_at(i))                       # enqueue step i and the read-back of its loss behind it
            handle = eng.read_loss_async(loss_dev)
            nxt = eng.stage_host(*host[(i + 1) % 2]) if i + 1 < n else None      # stage batch i+1 (host work + H2D) while it runs
            if pending is not None:
                pending.wait()                                                    # the loss of step i-1 reaches the host (every step's does)
            pending = handle
        pending.wait()

    e2e_loop(max(4, args.warmup))        # untimed: the staging buffers of the look-ahead pipeline (three batches alive) come from the allocator's cache afterwards
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    h2d = host[0][0].numel() * 4 + BATCH * 196 + BATCH * MASKED * 4
    # ---- the same e2e loop with the reference's BLOCK-WISE masks (masking_generator.py:29-92) drawn on the device: the masked-row count
    # then differs from batch to batch (a third of the images end 1-10 patches short of 120); one captured graph must serve them all.
    from uncertainty_vit_b200 import masking_generator as MG
    gen = MG.MaskingGenerator(14, MASKED, min_num_patches=16, seed=rank, device=dev)
    counts = []

    def blockwise_loop(n):
        nxt = eng.stage_device_masks(host[0][0], gen)
        pending = None
        for i in range(n):
            loss_dev = eng.launch_staged(nxt, lr=lr_at(i))
            handle = eng.read_loss_async(loss_dev)
            counts.append(nxt[5])
            nxt = eng.stage_device_masks(host[(i + 1) % 2][0], gen) if i + 1 < n else None
            if pending is not None:
                pending.wait()
            pending = handle
        pending.wait()

    blockwise_loop(max(4, args.warmup))
    barrier()
    del counts[:]
    t0 = time.perf_counter()
    blockwise_loop(args.steps)
    barrier()
    bw_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    bw_rows = [int(c.item()) for c in counts]
    blockwise = {"value": world * BATCH / (bw_ms * 1e-3), "unit": "img/s", "ms_per_step": bw_ms, "masked_rows_min": min(bw_rows),
                 "masked_rows_max": max(bw_rows), "graphs": len(eng._graphs), "h2d_bytes_per_step": host[0][0].numel() * 4,
                 "what": "e2e loop with device block-wise masks (variable masked-row count, padded row list, one CUDA graph)"}
    # ---- roofline: CUDA events around every tcgen05 GEMM launch inside 2 instrumented steps
    roof = None
    if rank == 0:
        ops.GEMM_TIMING = []
    # the instrumented steps run the step's launches on ONE stream (the timed steps overlap the EMA-teacher forward with the student's on a
    # second stream; events around a GEMM would then also time the kernels it shares the SMs with)
    serial = {k: os.environ.get(k) for k in ("B200VIT_TEACHER_STREAM",)}
    os.environ.update({k: "0" for k in serial})
    for i in range(2):                       # every rank steps (the step contains the gradient all-reduce); rank 0 records
        eng.step(*dev_batches[i % 2], lr=lr_at(i))
    torch.cuda.synchronize()
    for k, v in serial.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    if rank == 0:
        rec, ops.GEMM_TIMING = ops.GEMM_TIMING, None
        flops = sum(f for _, _, f in rec)
        t_ms = sum(a.elapsed_time(b) for a, b, _ in rec)
        pk = peaks()
        achieved = flops / (t_ms * 1e-3) / 1e12
        traffic = None            # DRAM bytes per GEMM launch (ncu capture of the same step, committed under profiles/)
        tpath = os.path.join(ROOT, "profiles", "r2_gemm_dram.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("bytes_per_launch")
        roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                "traffic": traffic, "traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r2_gemm_dram.json)", "kernel": "gemm_bf16_kernel (tcgen05, all 4 operand-major instantiations)", "launches_timed": len(rec),
                "gemm_ms_per_step": t_ms / 2, "gemm_share_of_step": (t_ms / 2) / ms, "peak_source": pk["source"] + " sustained (kernel timed inside a long step)",
                "frac_of_burst": achieved / pk["bf16_burst"]}
    barrier()
    # ---- the other BASELINE configs at the same N (skipped when this run IS the --stochastic step): driver-visible sub-blocks
    del eng, dev_batches
    torch.cuda.empty_cache()
    sub_steps = max(3, min(args.steps, 10))
    stoch_res = None if (args.stochastic or args.no_sub) else run_stochastic_block(dev, rank, world, E, M, barrier, max_over_ranks, sub_steps)
    ft_res = None if args.no_sub else run_finetune_large_block(dev, rank, world, E, M, barrier, max_over_ranks, sub_steps)
    mc_res = None if args.no_mc else run_mc_inference(dev, rank, world, M, barrier, max_over_ranks)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = 6 if args.stochastic else 14            # ~10 s of host work (0.7 s per 8-image step; the dual-stream step costs twice that)
        v, sec, cores = cpu_reference_steps(n_cpu, 2, stochastic=args.stochastic)
        cpu = {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full data2vec steps of {CPU_SAMPLE_BATCH} images after 2 warm-up ({sec:.2f} s/step, {n_cpu * sec:.1f} s of host work), "
                         f"fp32 oracle port, {cores} threads"}
    if rank == 0:
        emit = (lambda line: os.write(json_fd, (line + "\n").encode())) if json_fd is not None else print
        emit(json.dumps({
            "metric": "data2vec ViT-B/16 pretrain throughput", "value": world * BATCH / (ms * 1e-3), "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, args.stochastic),
            "e2e": {"value": world * BATCH / (e2e_ms * 1e-3), "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "stochastic": stoch_res, "finetune_large": ft_res, "mc_inference": mc_res, "blockwise_masks_e2e": blockwise,
            "final_loss": final_loss, "z0_mean": z0_mean,
            "step_tflops_algorithmic": (281.86e9 if args.stochastic else 140.93e9) * BATCH / (ms * 1e-3) / 1e12}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stochastic", action="store_true",
                    help="BASELINE.json configs[2] variant: the dual-stream --stochastic pre-training step (not the headline metric)")
    ap.add_argument("--no-mc", action="store_true", help="skip the MC-sample uncertainty-inference measurement (BASELINE.json configs[3])")
    ap.add_argument("--no-sub", action="store_true", help="skip the --stochastic step and ViT-L fine-tune sub-blocks (BASELINE.json configs[2] and [4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback); use --impl reference for the host baseline")
        run_b200(args)


if __name__ == "__main__":
    main()
