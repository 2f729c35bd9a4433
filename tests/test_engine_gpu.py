"""Fused data2vec step engine (B200) against the CPU oracle's full step on the same seeded inputs and injected noise."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _setup(cuda, golden_dir, name):
    import uncertainty_vit_b200 as pkg
    from uncertainty_vit_b200 import engine as E, modeling  # noqa: F401
    from tests.test_model_gpu import _build_from_gold
    gold = torch.load(os.path.join(golden_dir, name))
    model, arch, sd = _build_from_gold(pkg, gold, cuda)
    return pkg, E, gold, model, arch, sd


def test_engine_step_matches_oracle_step(cuda, golden_dir):
    from oracle import vit_oracle as O
    pkg, E, gold, model, arch, sd = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    eng = E.D2VEngine(model, lr=1e-3, weight_decay=0.05, clip_grad=3.0, ema_decay=0.99, target_layers=gold["target_layers"], l1_beta=2.0)
    x, mask = gold["x"], gold["mask"]
    n = gold["noise"]
    noise = pkg.core.Noise(seed=1)
    noise.drop_path_scale = torch.stack([k.float() / (1.0 - p) for k, p in zip(n["keep"], n["prob"])]).to(cuda).contiguous()
    noise.attn_keep = [k.to(cuda).contiguous() for k in n["attn_keep"]]
    m = mask.reshape(mask.shape[0], -1).numpy().astype(np.uint8)
    rows = torch.from_numpy(eng.rows_from_host_mask(m, arch.tokens)).to(cuda)
    loss = eng.step(x.to(cuda), torch.from_numpy(m.reshape(-1)).to(cuda), rows, noise=noise)
    # oracle: same step (teacher == student at step 0, as ModelEmaV2's deepcopy)
    sd_o = {k: v.clone() for k, v in sd.items()}
    ema_o = {k: v.clone() for k, v in sd.items()}
    opt = O.new_opt_state(sd_o)
    onoise = O.Noise(drop_path_keep=n["keep"], drop_path_prob=n["prob"], attn_keep=[k.float() for k in n["attn_keep"]], attn_drop=n["attn_drop"])
    loss_o, gnorm_o, grads_o = O.d2v_step(sd_o, ema_o, opt, arch, x, mask, 1, onoise, gold["target_layers"], lr=1e-3, wd=0.05, clip=3.0,
                                          ema_decay=0.99, return_grads=True)
    assert abs(loss.item() - loss_o) / loss_o < 2e-2
    assert abs(loss.item() - gold["loss"]) / gold["loss"] < 2e-2           # and the REAL reference's loss
    assert abs(eng.grad_norm().item() - gnorm_o) / gnorm_o < 3e-2
    for k, g in grads_o.items():
        mine = eng.grads[k].cpu()
        if float(g.norm()) > 1e-7:
            assert rel(mine, g) < 8e-2, (k, rel(mine, g))
    # after clip + AdamW + EMA: update directions agree (Adam's first step is sign-like, so compare by cosine)
    cos_num = cos_a = cos_b = 0.0
    for k in grads_o:
        du = (model.state_dict()[k].cpu() - sd[k]).double().flatten()
        do = (sd_o[k] - sd[k]).double().flatten()
        cos_num += float(du @ do); cos_a += float(du @ du); cos_b += float(do @ do)
    assert cos_num / (cos_a ** 0.5 * cos_b ** 0.5) > 0.97
    ema_sd = eng.ema_state_dict()
    for k in grads_o:
        # EMA of the fused kernel == d*e + (1-d)*p applied to ITS OWN updated weights, exactly
        expect = 0.99 * sd[k].to(cuda) + (1.0 - 0.99) * model.state_dict()[k]
        assert rel(ema_sd[k], expect) < 1e-6, k
    # bf16 shadows track the masters
    assert rel(eng.p16.float(), eng.p32) < 5e-3 and rel(eng.e16.float(), eng.e32) < 5e-3


def test_engine_overfits_fixed_batch(cuda, golden_dir):
    pkg, E, gold, model, arch, sd = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    eng = E.D2VEngine(model, lr=2e-3, weight_decay=0.0, clip_grad=3.0, ema_decay=1.0, target_layers=gold["target_layers"])
    x = gold["x"].pin_memory()
    mask = gold["mask"].numpy()
    losses = [eng.step_host(x, mask) for _ in range(30)]
    assert all(np.isfinite(losses))
    assert losses[-1] < 0.6 * losses[0], losses[::5]


def test_train_one_epoch_loop_and_schedules(cuda, golden_dir):
    pkg, E, gold, model, arch, sd = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    eng = E.D2VEngine(model, ema_decay=0.9998, ema_decay_init=0.999, ema_start_at=4, target_layers=gold["target_layers"])
    lr = E.cosine_scheduler(2e-3, 1e-5, 2, 3, warmup_epochs=1)
    wd = E.cosine_scheduler(0.05, 0.05, 2, 3)
    assert len(lr) == 6 and abs(lr[0]) < 1e-12 and abs(lr[3] - 2e-3) < 1e-9
    loader = [((gold["x"], gold["mask"]), None)] * 3
    stats = E.train_one_epoch(eng, loader, epoch=0, start_steps=0, lr_schedule_values=lr, wd_schedule_values=wd, log=lambda s: None)
    assert np.isfinite(stats["loss"]) and abs(stats["cur_decay"] - (0.999 + 2 * (0.9998 - 0.999) / 4)) < 1e-9


def test_dist_engine_step_matches_oracle_step(cuda, golden_dir):
    """--stochastic data2vec step (dual-stream teacher/student, smooth-L1 + WassersteinLoss) through the fused engine vs the oracle step."""
    import uncertainty_vit_b200 as pkg
    from uncertainty_vit_b200 import engine as E
    from oracle import vit_oracle as O
    from tests.test_model_gpu import _build_dist
    gold = torch.load(os.path.join(golden_dir, "tiny_dist_cyclical.pt"))
    model, arch, sd = _build_dist(pkg, gold, cuda)
    lam = 1e-2          # larger than the README's 1e-5 so that the W-loss gradients matter in the comparison
    eng = E.D2VEngine(model, lr=1e-3, weight_decay=0.05, clip_grad=3.0, ema_decay=0.99, target_layers=gold["target_layers"], lambda_pretraining=lam)
    n = gold["noise"]
    noise = pkg.core.Noise(seed=1)
    noise.drop_path_scale = torch.stack([k.float() / (1.0 - p) for k, p in zip(n["keep"], n["prob"])]).to(cuda).contiguous()
    noise.attn_keep = [k.to(cuda).contiguous() for k in n["attn_keep"]]
    m = gold["mask"].reshape(gold["mask"].shape[0], -1).numpy().astype(np.uint8)
    rows = torch.from_numpy(eng.rows_from_host_mask(m, arch.tokens)).to(cuda)
    w_before = model.state_dict()["blocks.0.attn.cov_qkv.weight"].clone()
    loss = eng.step(gold["x"].to(cuda), torch.from_numpy(m.reshape(-1)).to(cuda), rows, noise=noise)
    sd_o = {k: v.clone() for k, v in sd.items()}
    ema_o = {k: v.clone() for k, v in sd.items()}
    onoise = O.Noise(drop_path_keep=n["keep"], drop_path_prob=n["prob"], attn_keep=[k.float() for k in n["attn_keep"]], attn_drop=n["attn_drop"])
    loss_o, gnorm_o, grads_o = O.d2v_step(sd_o, ema_o, O.new_opt_state(sd_o), arch, gold["x"], gold["mask"], 1, onoise, gold["target_layers"],
                                          lr=1e-3, wd=0.05, clip=3.0, ema_decay=0.99, return_grads=True, lam=lam)
    assert abs(loss.item() - loss_o) / loss_o < 2e-2
    assert abs(eng.grad_norm().item() - gnorm_o) / gnorm_o < 4e-2
    bad = []
    for k, g in grads_o.items():
        if float(g.norm()) > 1e-7:
            e = rel(eng.grads[k].cpu(), g)
            if e > 1e-1:
                bad.append((k, round(e, 3)))
    assert not bad, bad
    # the unused cov_qkv.weight is neither updated nor decayed (torch AdamW skips grad-less parameters)
    assert torch.equal(model.state_dict()["blocks.0.attn.cov_qkv.weight"], w_before)


def test_graphed_step_matches_eager(cuda):
    """The CUDA-graph replay of forward+backward (engine._fwd_bwd_graphed) and the eager launch sequence are the same computation:
    same seeds -> same losses up to the fp32 atomic-order noise of the split-K / column-sum reductions."""
    from functools import partial
    from uncertainty_vit_b200 import engine as E, modeling as M
    losses = {}
    for use_graph in (False, True):
        torch.manual_seed(0)
        model = M.VisionTransformerForCyclicalTraining(img_size=224, patch_size=16, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                                       norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), use_shared_rel_pos_bias=True,
                                                       use_abs_pos_emb=False, init_values=0.1, drop_path_rate=0.1, attn_drop_rate=0.05).to(cuda)
        eng = E.D2VEngine(model, lr=1e-3, target_layers=[0, 1], use_graph=use_graph, seed=5)
        g = torch.Generator().manual_seed(1)
        B = 4
        x = torch.randn(B, 3, 224, 224, generator=g).to(cuda)
        mask = np.zeros((B, 196), dtype=np.uint8)
        for b in range(B):
            mask[b, torch.randperm(196, generator=g)[:120].numpy()] = 1
        rows = torch.from_numpy(eng.rows_from_host_mask(mask, 197)).to(cuda)
        mu8 = torch.from_numpy(mask.reshape(-1)).to(cuda)
        losses[use_graph] = [float(eng.step(x, mu8, rows).item()) for _ in range(6)]
        if use_graph:
            assert len(eng._graphs) == 1, "steps 3.. must have been replayed from the captured graph"
    a, b = np.array(losses[False]), np.array(losses[True])
    assert np.all(np.isfinite(b)) and np.max(np.abs(a - b) / np.abs(a)) < 5e-3, (a, b)


def test_finetune_engine_matches_module_autograd(cuda, golden_dir):
    """FinetuneEngine.step (flat arenas, fused layer-decay AdamW) computes the same loss and parameter gradients as the nn.Module boundary
    + torch autograd (which test_finetune_training_backward_vs_oracle pins to the oracle) on the same inputs and injected noise."""
    import uncertainty_vit_b200 as pkg
    from uncertainty_vit_b200 import engine as E
    from tests.test_model_gpu import _build_from_gold, _build_dist
    for name, builder in (("tiny_det_finetune", _build_from_gold), ("tiny_dist_finetune", lambda p, g, c: _build_dist(p, g, c))):
        gold = dict(torch.load(os.path.join(golden_dir, name + ".pt")), dpr=0.2, attn_drop=0.1)
        model, arch, sd = builder(pkg, gold, cuda)
        B = gold["B"]
        g = torch.Generator().manual_seed(11)
        probs = [float(x) for x in torch.linspace(0, 0.2, arch.depth)]
        draws = 4 if arch.dist else 2
        keeps = [(torch.rand(draws, B, generator=g) >= p).float() for p in probs]
        akeep = [(torch.rand(B, arch.num_heads, arch.tokens, arch.tokens, generator=g) >= 0.1).to(torch.uint8) for _ in range(arch.depth)]
        K = model.cfg.num_classes
        targets = torch.softmax(torch.randn(B, K, generator=g), -1).to(cuda)
        x = gold["x"].to(cuda)
        eng = E.FinetuneEngine(model, lr=1e-3, weight_decay=0.05, layer_decay=0.65, clip_grad=None)
        # module boundary + autograd (parameters alias the engine's arena)
        model.train()
        model.inject_noise(drop_path_keep=keeps, attn_keep=akeep)
        out = model(x)
        logits = out[-1] if isinstance(out, tuple) else out
        loss_m = torch.sum(-targets * torch.log_softmax(logits.float(), -1), -1).mean()
        loss_m.backward()
        ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        before = eng.p32.clone()
        noise = pkg.core.Noise(seed=1)
        noise.drop_path_scale = torch.stack([k.float() / (1.0 - p) for k, p in zip(keeps, probs)]).to(cuda).contiguous()
        noise.attn_keep = [k.to(cuda).contiguous() for k in akeep]
        loss_e = float(eng.step(x, targets, noise=noise).item())
        assert abs(loss_e - float(loss_m)) / abs(float(loss_m)) < 1e-3, (name, loss_e, float(loss_m))
        bad = [(n, round(rel(eng.grads[n], gr), 4)) for n, gr in ref.items() if float(gr.norm()) > 1e-6 and rel(eng.grads[n], gr) > 2e-2]
        assert not bad, (name, bad)
        assert not torch.equal(before, eng.p32), "the fused AdamW must have updated the arena"
        if arch.dist:      # triplet step: positive / negative forwards + WassersteinLossFineTuning on the features
            l2 = float(eng.step(x, targets, x.roll(1, 0).contiguous(), x.flip(0).contiguous()).item())
            assert np.isfinite(l2) and torch.isfinite(eng.g32).all()


def _gold_noise(pkg, gold, cuda):
    n = gold["noise"]
    noise = pkg.core.Noise(seed=1)
    noise.drop_path_scale = torch.stack([k.float() / (1.0 - p) for k, p in zip(n["keep"], n["prob"])]).to(cuda).contiguous()
    noise.attn_keep = [k.to(cuda).contiguous() for k in n["attn_keep"]]
    return noise


@pytest.mark.parametrize("name", ["tiny_det_cyclical.pt", "tiny_dist_cyclical.pt"])
def test_padded_row_list_is_the_same_step(cuda, golden_dir, name):
    """Block-wise masking gives a different masked-row count every batch; the engine keeps one launch shape by padding the row list to a
    capacity and passing the true count in device memory. A step on the padded list (padding = stale but valid row numbers) must equal
    the step on the exact list (itself pinned to the oracle above): same loss, same gradients, same updated weights."""
    import uncertainty_vit_b200 as pkg
    from uncertainty_vit_b200 import engine as E, modeling  # noqa: F401  (pkg.modeling is used by _build_from_gold)
    from tests.test_model_gpu import _build_dist, _build_from_gold
    res = {}
    for padded in (False, True):
        gold = torch.load(os.path.join(golden_dir, name))
        model, arch, sd = (_build_dist if "dist" in name else _build_from_gold)(pkg, gold, cuda)
        eng = E.D2VEngine(model, lr=1e-3, weight_decay=0.05, clip_grad=3.0, ema_decay=0.99, target_layers=gold["target_layers"],
                          lambda_pretraining=1e-2)
        m = gold["mask"].reshape(gold["mask"].shape[0], -1).numpy().astype(np.uint8)
        rows = torch.from_numpy(eng.rows_from_host_mask(m, arch.tokens)).to(cuda)
        R = rows.numel()
        n_valid = None
        if padded:
            pad = torch.randint(0, m.shape[0] * arch.tokens, (13,), dtype=torch.int32, generator=torch.Generator().manual_seed(2)).to(cuda)
            rows = torch.cat([rows, pad])
            n_valid = torch.tensor([R], dtype=torch.int32, device=cuda)
        loss = eng.step(gold["x"].to(cuda), torch.from_numpy(m.reshape(-1)).to(cuda), rows, noise=_gold_noise(pkg, gold, cuda), n_valid=n_valid)
        res[padded] = (float(loss.item()), eng.g32.clone(), eng.p32.clone(), float(eng.grad_norm().item()))
    (l0, g0, p0, n0), (l1, g1, p1, n1) = res[False], res[True]
    assert abs(l0 - l1) / abs(l0) < 1e-5, (l0, l1)
    assert abs(n0 - n1) / n0 < 1e-3
    assert rel(g1, g0) < 2e-3, rel(g1, g0)                 # bf16 dy rounding of the rescaled grad_scale + fp32 atomic order
    assert rel(p1, p0) < 1e-2                               # (Adam's first step is sign-like: only a loose check on the weights)


def test_graph_serves_changing_masked_row_counts(cuda):
    """One captured graph replays batches whose masked-row counts differ (device block-wise masks, no read-back): the loss of each
    replayed step equals the eager step on the exact row list with the same seed."""
    from functools import partial
    from uncertainty_vit_b200 import engine as E, masking_generator as MG, modeling as M
    losses = {}
    for use_graph in (False, True):
        torch.manual_seed(0)
        model = M.VisionTransformerForCyclicalTraining(img_size=224, patch_size=16, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                                       norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), use_shared_rel_pos_bias=True,
                                                       use_abs_pos_emb=False, init_values=0.1, drop_path_rate=0.1, attn_drop_rate=0.05).to(cuda)
        eng = E.D2VEngine(model, lr=1e-3, target_layers=[0, 1], use_graph=use_graph, seed=5)
        gen = MG.MaskingGenerator(14, 120, min_num_patches=16, seed=9, device=cuda)
        g = torch.Generator().manual_seed(1)
        B = 8
        out, counts = [], []
        for i in range(7):
            x = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
            if use_graph:
                staged = eng.stage_device_masks(x, gen)
                staged[3].synchronize()             # (the count lives on the copy stream; the engine itself never reads it back)
                counts.append(int(staged[5].item()))
                out.append(float(eng.launch_staged(staged).item()))
            else:                                   # exact row list on the host from the same device masks
                mask, count, _ = gen.batch(B)
                counts.append(int(count[B].item()))
                out.append(eng.step_host(x, mask.cpu().numpy()))
        losses[use_graph] = out
        if use_graph:
            assert len(eng._graphs) == 1 and eng._eager_steps == 2, "steps 3.. must replay ONE graph whatever the count"
            assert len(set(counts)) > 1, counts      # the counts did change
        else:
            ref_counts = counts
    assert counts == ref_counts
    a, b = np.array(losses[False]), np.array(losses[True])
    assert np.all(np.isfinite(b)) and np.max(np.abs(a - b) / np.abs(a)) < 5e-3, (a, b)


def test_train_one_epoch_with_device_masks(cuda, golden_dir):
    pkg, E, gold, model, arch, sd = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    from uncertainty_vit_b200 import masking_generator as MG
    eng = E.D2VEngine(model, target_layers=gold["target_layers"])
    side = int(round((arch.tokens - 1) ** 0.5))
    gen = MG.MaskingGenerator(side, max(2, (arch.tokens - 1) * 6 // 10), min_num_patches=2, seed=1, device=cuda)
    loader = [(gold["x"], None)] * 5
    stats = E.train_one_epoch(eng, loader, log=lambda s: None, mask_generator=gen)
    assert np.isfinite(stats["loss"]) and gen.images_drawn == 5 * gold["x"].shape[0]


def test_engine_checkpoint_resume_and_torch_adamw_interop(cuda, golden_dir, tmp_path):
    """save_model / auto_load_model over the fused engine (utils.py:462-520): a run resumed from the checkpoint file continues exactly like
    the uninterrupted run, and the stored optimizer entry loads into the torch.optim.AdamW the reference builds (same parameter numbering
    and moments), so the reference runner can resume an engine checkpoint."""
    from uncertainty_vit_b200 import checkpoint as CK
    pkg, E, gold, model, arch, sd = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    kw = dict(lr=1e-3, weight_decay=0.05, clip_grad=3.0, ema_decay=0.99, ema_decay_init=0.9, ema_start_at=10, target_layers=gold["target_layers"],
              use_graph=False, seed=4)
    eng = E.D2VEngine(model, **kw)
    x = gold["x"].pin_memory()
    mask = gold["mask"].numpy()
    for _ in range(3):
        eng.step_host(x, mask)
    path = CK.save_model(str(tmp_path), 0, eng)
    assert path.endswith("checkpoint-0.pth") and CK.latest_checkpoint(str(tmp_path)) == path
    cont = [eng.step_host(x, mask) for _ in range(2)]
    # fresh process: new model (different init), new engine, resume
    _, _, _, model2, _, _ = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    with torch.no_grad():
        for p in model2.parameters():
            p.add_(0.1)
    eng2 = E.D2VEngine(model2, **kw)
    start_epoch = CK.auto_load_model(str(tmp_path), eng2)
    assert start_epoch == 1 and eng2.it == 3 and eng2.opt_step == 3
    resumed = [eng2.step_host(x, mask) for _ in range(2)]
    assert np.allclose(cont, resumed, rtol=1e-4), (cont, resumed)             # fp32 atomic-order noise only
    assert rel(eng2.p32, eng.p32) < 1e-5 and rel(eng2.e32, eng.e32) < 1e-6 and rel(eng2.m32, eng.m32) < 1e-3
    # the reference side: torch AdamW over optim_factory's groups accepts the optimizer entry
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ckpt) >= {"model", "optimizer", "epoch", "model_ema"} and list(ckpt["model_ema"]) == list(ckpt["model"])
    ref_model = {n: torch.nn.Parameter(p.detach().cpu().clone()) for n, p in model2.named_parameters()}
    groups = CK.engine_parameter_groups(eng2)
    opt = torch.optim.AdamW([{"params": [ref_model[n] for n in g["params"]], "weight_decay": g["weight_decay"], "lr_scale": g["lr_scale"]} for g in groups],
                            lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    opt.load_state_dict(ckpt["optimizer"])
    st = opt.state[ref_model["blocks.1.mlp.fc1.weight"]]
    assert float(st["step"]) == 3.0
    for p, s in opt.state.items():
        assert torch.isfinite(s["exp_avg"]).all() and s["exp_avg"].shape == p.shape
    names = [n for g in groups for n in g["params"]]
    k = names.index("blocks.1.mlp.fc1.weight")
    assert torch.equal(ckpt["optimizer"]["state"][k]["exp_avg"], st["exp_avg"])
    # and back: an optimizer state dict produced by torch loads into a fresh engine
    _, _, _, model3, _, _ = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    eng3 = E.D2VEngine(model3, **kw)
    CK.load_optimizer_state_dict(eng3, opt.state_dict())
    assert eng3.opt_step == 3
    ck3 = CK.optimizer_state_dict(eng3)
    assert torch.equal(ck3["state"][k]["exp_avg_sq"], ckpt["optimizer"]["state"][k]["exp_avg_sq"])


def test_finetune_loop_and_evaluate(cuda, golden_dir):
    """finetune_one_epoch (engine_for_finetuning.train_one_epoch / dist_train_one_epoch over the fused engine, with device Mixup) and evaluate
    (engine_for_finetuning.evaluate): the loop learns a fixed batch, reports class_acc only without Mixup, and the evaluation metrics equal
    the oracle's on the model's own logits."""
    import uncertainty_vit_b200 as pkg
    from oracle import vit_oracle as O
    from uncertainty_vit_b200 import engine as E, mixup as MX, modeling  # noqa: F401
    from tests.test_model_gpu import _build_dist, _build_from_gold
    for name, builder in (("tiny_det_finetune", _build_from_gold), ("tiny_dist_finetune", _build_dist)):
        gold = torch.load(os.path.join(golden_dir, name + ".pt"))
        model, arch, sd = builder(pkg, gold, cuda)
        K = model.cfg.num_classes
        g = torch.Generator().manual_seed(3)
        x = torch.randn(4, 3, arch.img_size, arch.img_size, generator=g)
        y = torch.randint(0, K, (4,), generator=g)
        eng = E.FinetuneEngine(model, lr=2e-3, weight_decay=0.0, layer_decay=0.65, clip_grad=3.0)
        model.eval()
        with torch.no_grad():                      # builds the module's bf16 weight shadows on the arena views BEFORE any training
            before = model(x.to(cuda))
        before = (before[-1] if isinstance(before, tuple) else before).float().clone()
        batch = (x, x.roll(1, 0), x.flip(0), y) if arch.dist else (x, y)
        lr = E.cosine_scheduler(2e-3, 1e-4, 1, 25, warmup_epochs=0)
        logs = []
        st = E.finetune_one_epoch(eng, [batch] * 25, lr_schedule_values=lr, num_training_steps_per_epoch=25, log=logs.append)
        first = float(logs[0].split("loss ")[1].split()[0])
        assert np.isfinite(st["loss"]) and st["loss"] < first and st["class_acc"] is not None and 0 <= st["class_acc"] <= 1
        assert st["min_lr"] < st["lr"] <= 2e-3 and np.isfinite(st["grad_norm"]) and len(logs) == 3
        mix = MX.Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, label_smoothing=0.1, num_classes=K, rng=np.random.RandomState(0))
        st2 = E.finetune_one_epoch(eng, [batch] * 4, start_steps=20, lr_schedule_values=lr, mixup_fn=mix, log=lambda s: None)
        assert np.isfinite(st2["loss"]) and st2["class_acc"] is None
        with pytest.raises(NotImplementedError):
            E.finetune_one_epoch(eng, [batch], update_freq=2)
        # the module boundary sees the weights the engine's kernels wrote (its bf16 shadows are keyed on a version the engine bumps): a fresh
        # module loaded from the state dict gives the same logits
        model.eval()
        with torch.no_grad():
            live = model(x.to(cuda))
        gold2 = torch.load(os.path.join(golden_dir, name + ".pt"))
        fresh, _, _ = builder(pkg, gold2, cuda)
        fresh.load_state_dict(model.state_dict())
        fresh.eval()
        with torch.no_grad():
            ref = fresh(x.to(cuda))
        pick = lambda o: (o[-1] if isinstance(o, tuple) else o).float()
        assert rel(pick(live), pick(ref)) < 1e-6 and rel(pick(live), before) > 1e-2
        # evaluation: two batches of different size
        loader = [(x, y), (x[:3], y[:3])]
        ev = E.evaluate(loader, model, cuda, K)
        model.eval()
        with torch.no_grad():
            outs = [model(b.to(cuda)) for b, _ in loader]
        outs = [(o[-1] if isinstance(o, tuple) else o).float().cpu() for o in outs]
        exp = {"acc1": 0.0, "acc5": 0.0, "ECE": 0.0, "NLL": 0.0}
        for o, (_, t) in zip(outs, loader):
            p = torch.softmax(o, 1)
            top5 = o.topk(5, 1).indices
            b = o.shape[0]
            exp["acc1"] += 100.0 * float((o.argmax(1) == t).float().mean()) * b
            exp["acc5"] += 100.0 * float((top5 == t[:, None]).any(1).float().mean()) * b
            exp["ECE"] += float(O.ece(p, t)) * b
            exp["NLL"] += float(-torch.log(p[torch.arange(b), t]).mean()) * b
        for k, v in exp.items():
            assert abs(ev[k] - v / 7) < 2e-3 * max(1.0, abs(v / 7)), (name, k, ev[k], v / 7)
        ce = np.mean([float(torch.nn.functional.cross_entropy(o, t)) for o, (_, t) in zip(outs, loader)])
        assert abs(ev["loss"] - ce) < 1e-3
        auroc = sum(O.auroc_macro_ovr(torch.softmax(o, 1), t) * o.shape[0] for o, (_, t) in zip(outs, loader)) / 7
        assert abs(ev["AUROC"] - auroc) < 1e-4 and np.isnan(ev["TACE"])         # TACE's 30 adaptive bins need >= 30 images per batch
        if arch.dist:      # dist_evaluate: triplet batches, loss = CE + WassersteinLossFineTuning(anchor, positive, negative)
            trip = [(x, x.roll(1, 0), x.flip(0), y)]
            ev3 = E.evaluate(trip, model, cuda, K, dist_criterion=(1e-2, 1e-4))
            with torch.no_grad():
                a, p_, n_ = (model(t.to(cuda)) for t in trip[0][:3])
                expect = torch.nn.functional.cross_entropy(a[2].float(), y.to(cuda)) + O.wasserstein_loss_finetune(a[0].float(), a[1].float(), p_[0].float(), p_[1].float(), n_[0].float(), n_[1].float(), 1e-2, 1e-4)
            assert abs(ev3["loss"] - float(expect)) < 1e-4 * max(1.0, abs(float(expect))) and ev3["loss"] > E.evaluate(trip, model, cuda, K)["loss"]


def test_uint8_staging_is_the_same_step(cuda, golden_dir):
    """A uint8 pixel batch normalised on the device gives the step of the fp32 batch torchvision's ToTensor + Normalize would have produced."""
    losses = {}
    g = torch.Generator().manual_seed(7)
    pix = torch.randint(0, 256, (3, 64, 64, 3), dtype=torch.uint8, generator=g)
    mean, std = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)
    f32 = pix.permute(0, 3, 1, 2).contiguous().float().div(255).sub_(torch.tensor(mean).view(-1, 1, 1)).div_(torch.tensor(std).view(-1, 1, 1))
    for kind in ("f32", "u8"):
        pkg, E, gold, model, arch, sd = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
        eng = E.D2VEngine(model, lr=1e-3, target_layers=gold["target_layers"], use_graph=False, seed=2)
        eng.pixel_norm = (mean, std, True)
        x = (pix if kind == "u8" else f32).pin_memory()
        losses[kind] = [eng.step_host(x, gold["mask"].numpy()) for _ in range(3)]
    assert np.allclose(losses["f32"], losses["u8"], rtol=1e-5), losses
    pkg, E, gold, model, arch, sd = _setup(cuda, golden_dir, "tiny_det_cyclical.pt")
    with pytest.raises(ValueError):
        E.D2VEngine(model, target_layers=gold["target_layers"]).step_host(pix.pin_memory(), gold["mask"].numpy())
