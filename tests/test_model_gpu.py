"""Model-level parity (B200): the nn.Module boundary (reference names / signatures) through the CUDA path against
(1) the golden vectors produced by the REAL reference and (2) the CPU oracle on the same seeded inputs and injected noise.
bf16 compute vs the fp32 reference: tolerance 2e-2 relative (BASELINE.json north_star); integer decisions bit-exact."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(scope="module")
def pkg(cuda):
    import uncertainty_vit_b200 as p
    from uncertainty_vit_b200 import modeling  # noqa: F401
    return p


def _build_from_gold(pkg, gold, cuda, **extra):
    from functools import partial
    from oracle import vit_oracle as O
    a = gold["arch"]
    arch = O.Arch(**a)
    kw = dict(img_size=a["img_size"], patch_size=16, embed_dim=a["embed_dim"], depth=a["depth"], num_heads=a["num_heads"], mlp_ratio=4,
              qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), use_shared_rel_pos_bias=True, use_abs_pos_emb=False,
              init_values=0.1, drop_path_rate=gold.get("dpr", 0.0), attn_drop_rate=gold.get("attn_drop", 0.0), **extra)
    M = pkg.modeling
    if arch.kind == "cyclical":
        model = M.VisionTransformerForCyclicalTraining(**kw)
    else:
        model = M.VisionTransformer(num_classes=a["num_classes"], **kw)
    sd = O.make_state(arch, gold["seed"])
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    return model.to(cuda), arch, sd


def test_state_dict_names_match_reference(pkg, cuda, golden_dir):
    """Checkpoint compatibility: exactly the reference's parameter/buffer names and shapes."""
    from oracle import vit_oracle as O
    for kind in ("cyclical", "finetune"):
        arch = O.Arch(kind=kind, **O.TINY)
        names = {n: s for n, s, _ in O.state_names(arch)}
        names["rel_pos_bias.relative_position_index"] = (arch.tokens, arch.tokens)
        gold = dict(arch=arch.__dict__.copy(), seed=0)
        model, _, _ = _build_from_gold(pkg, gold, cuda)
        own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert own == {k: tuple(v) for k, v in names.items()}
        assert torch.equal(model.rel_pos_bias.relative_position_index.cpu(), O.relative_position_index(arch.grid, arch.grid))
        assert model.get_num_layers() == arch.depth and model.no_weight_decay() == {"pos_embed", "cls_token"}
        assert model.patch_embed.patch_shape == (arch.grid, arch.grid) and model.patch_embed.num_patches == arch.num_patches


def test_tiny_cyclical_against_reference_golden(pkg, cuda, golden_dir):
    gold = torch.load(os.path.join(golden_dir, "tiny_det_cyclical.pt"))
    model, arch, sd = _build_from_gold(pkg, gold, cuda)
    x, mask = gold["x"].to(cuda), gold["mask"].to(cuda)
    # teacher: eval, unmasked, per-layer outputs
    model.eval()
    with torch.no_grad():
        layers = model(x, bool_masked_pos=None, return_all_tokens=True, layer_results="end")
    assert len(layers) == arch.depth and tuple(layers[0].shape) == (gold["B"], arch.num_patches, arch.embed_dim)
    for a, b in zip(layers, gold["teacher_layers"]):
        assert rel(a.cpu(), b) < 2e-2
    # student: train mode, masked, injected drop-path keeps and dropout masks
    model.train()
    n = gold["noise"]
    model.inject_noise(drop_path_keep=n["keep"], attn_keep=n["attn_keep"])
    out = model(x, bool_masked_pos=mask, return_all_tokens=False)
    assert tuple(out.shape) == tuple(gold["outputs"].shape)          # [sum(mask), C], row order of the boolean gather
    assert rel(out.detach().cpu(), gold["outputs"]) < 2e-2
    loss = torch.nn.functional.smooth_l1_loss(out.float(), gold["targets"].to(cuda), beta=2.0)
    assert abs(loss.item() - gold["loss"]) / gold["loss"] < 2e-2
    loss.backward()
    worst = 0.0
    for name, p in model.named_parameters():
        dig = gold["grads"][name]
        assert p.grad is not None, name
        g = p.grad.detach().cpu().flatten()
        e = abs(float(g.double().norm()) - dig["norm"]) / max(dig["norm"], 1e-12)
        worst = max(worst, e)
        assert e < 3e-2, (name, e)
        if float(dig["head"].abs().max()) > 1e-8:
            assert rel(g[:64], dig["head"]) < 6e-2, name
    print("worst grad-norm rel err", worst)


def test_tiny_cyclical_against_oracle_all_modes(pkg, cuda, golden_dir):
    from oracle import vit_oracle as O
    gold = torch.load(os.path.join(golden_dir, "tiny_det_cyclical.pt"))
    model, arch, sd = _build_from_gold(pkg, gold, cuda)
    x, mask = gold["x"], gold["mask"]
    model.eval()
    with torch.no_grad():
        allt = model(x.to(cuda), bool_masked_pos=mask.to(cuda), return_all_tokens=True)
        ref = O.cyclical_forward(sd, arch, x, mask, return_all_tokens=True)
        assert rel(allt.cpu(), ref) < 2e-2
        fc = model(x.to(cuda), None, layer_results="fc")
        reffc = O.cyclical_forward(sd, arch, x, None, layer_results="fc")
        for a, b in zip(fc, reffc):
            assert rel(a.cpu(), b) < 3e-2
        assert model(x.to(cuda), None, layer_results="other") == []
        # empty mask -> zero rows; full mask -> all rows
        empty = model(x.to(cuda), torch.zeros_like(mask).to(cuda))
        assert tuple(empty.shape) == (0, arch.embed_dim)
        full = model(x.to(cuda), torch.ones_like(mask).to(cuda))
        reffull = O.cyclical_forward(sd, arch, x, torch.ones_like(mask))
        assert rel(full.cpu(), reffull) < 2e-2
    # Philox streams: drop-path / dropout change the output between calls, eval does not
    model.train()
    with torch.no_grad():
        a = model(x.to(cuda), mask.to(cuda))
        b = model(x.to(cuda), mask.to(cuda))
    assert not torch.equal(a, b)
    model.eval()
    with torch.no_grad():
        a = model(x.to(cuda), mask.to(cuda))
        b = model(x.to(cuda), mask.to(cuda))
    assert torch.equal(a, b)


def test_tiny_finetune_against_reference_golden(pkg, cuda, golden_dir):
    gold = torch.load(os.path.join(golden_dir, "tiny_det_finetune.pt"))
    model, arch, sd = _build_from_gold(pkg, gold, cuda)
    model.eval()
    with torch.no_grad():
        logits = model(gold["x"].to(cuda))
    assert tuple(logits.shape) == tuple(gold["logits"].shape)
    assert rel(logits.cpu(), gold["logits"]) < 2e-2


def test_vitb_cyclical_b2_against_reference_golden(pkg, cuda, golden_dir):
    """Full-size ViT-B/16 (config 2 shapes at B=2): student output, loss and every parameter-gradient norm vs the reference."""
    from oracle import vit_oracle as O
    gold = torch.load(os.path.join(golden_dir, "vitb_det_cyclical_b2.pt"))
    model, arch, sd = _build_from_gold(pkg, gold, cuda)
    x, mask = gold["x"].to(cuda), gold["mask"].to(cuda)
    model.eval()
    with torch.no_grad():
        layers = model(x, None, layer_results="end")
    tgt = O.build_targets([l.cpu() for l in layers], gold["target_layers"], gold["mask"], post_target_layer_norm=True)
    assert rel(tgt, gold["targets"]) < 2e-2
    model.train()
    n = gold["noise"]
    model.inject_noise(drop_path_keep=n["keep"], attn_keep=n["attn_keep"])
    out = model(x, mask)
    assert rel(out.detach().cpu(), gold["outputs"]) < 2e-2
    loss = torch.nn.functional.smooth_l1_loss(out.float(), gold["targets"].to(cuda), beta=2.0)
    assert abs(loss.item() - gold["loss"]) / gold["loss"] < 2e-2
    loss.backward()
    bad = []
    for name, p in model.named_parameters():
        dig = gold["grads"][name]
        e = abs(float(p.grad.double().norm()) - dig["norm"]) / max(dig["norm"], 1e-12)
        if e > 4e-2:
            bad.append((name, e))
    assert not bad, bad[:10]


def test_registry_boundary(pkg, cuda):
    M = pkg.modeling
    m = M.create_model("beit_base_patch16_224", pretrained=False, drop_path_rate=0.25, drop_rate=0.0, use_shared_rel_pos_bias=True,
                       use_abs_pos_emb=False, init_values=1e-4, attn_drop_rate=0.05, gp_layer=False, gumbel_softmax=False, sinkformer=False,
                       h_sto_trans=False)
    assert isinstance(m, M.VisionTransformerForCyclicalTraining)
    assert sum(p.numel() for p in m.parameters()) == 86_257_152 + 0 or abs(sum(p.numel() for p in m.parameters()) - 86.257e6) < 5e3
    with pytest.raises(NotImplementedError):
        M.create_model("beit_base_patch16_224", use_shared_rel_pos_bias=True, init_values=0.1, sinkformer=True)
    f = M.create_model("beit_base_patch16_224", num_classes=1000, drop_rate=0.0, drop_path_rate=0.1, attn_drop_rate=0.0, drop_block_rate=None,
                       use_mean_pooling=True, init_scale=0.001, use_rel_pos_bias=False, use_shared_rel_pos_bias=True, use_abs_pos_emb=False,
                       init_values=0.1)
    assert isinstance(f, M.VisionTransformer) and f.head.out_features == 1000
    with pytest.raises(pkg._lib.B200VitError):
        m(torch.zeros(1, 3, 224, 224), None)            # CPU tensor: no fallback


def test_mc_dropout_eval_matches_oracle_on_same_masks(pkg, cuda, golden_dir):
    """evaluate_MC_dropout semantics (uncertainty_evaluations.py:42-89): S passes with attention dropout ON in eval mode (enable_dropout),
    DropPath OFF; logits per pass against the oracle run with the SAME Philox masks (materialised by b200vit_dropout_mask); metrics from
    b200vit_mc_reduce against the oracle's reduction."""
    from oracle import vit_oracle as O
    from uncertainty_vit_b200 import mc
    gold = torch.load(os.path.join(golden_dir, "tiny_det_finetune.pt"))
    gold = dict(gold, dpr=0.2, attn_drop=0.1)
    model, arch, sd = _build_from_gold(pkg, gold, cuda)
    x = gold["x"].to(cuda)
    labels = torch.tensor([1, 3, 5])
    S = 4
    model.eval()
    mc.enable_dropout(model)
    assert model.blocks[0].attn.attn_drop.training and not model.blocks[1].drop_path.training
    res = mc.evaluate_mc_dropout(model, [(x, labels)], S)
    # the S passes of the 3 images are batched into ONE forward of S x 3 rows (row s * 3 + n = pass s of image n): re-run it to capture the
    # per-pass logits and the exact Philox masks of every row
    model._seed_calls = 0
    B = gold["B"]
    model.eval(); mc.enable_dropout(model)
    noise_seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * 1) & 0xFFFFFFFFFFFFFFFF
    with torch.no_grad():
        logits_all = model(x.repeat(S, 1, 1, 1).contiguous())
    keeps_all = [pkg.ops.dropout_mask(S * B * arch.num_heads, arch.tokens, 0.1, noise_seed, l, cuda).view(S * B, arch.num_heads, arch.tokens, arch.tokens)
                 for l in range(arch.depth)]
    per_pass = []
    for s in range(S):
        onoise = O.Noise(attn_keep=[k[s * B:(s + 1) * B].float().cpu() for k in keeps_all], attn_drop=0.1)
        ref = O.finetune_forward(sd, arch, gold["x"], noise=onoise)
        assert rel(logits_all[s * B:(s + 1) * B].cpu(), ref) < 2e-2, s
        per_pass.append(ref)
    r = O.mc_reduce(torch.stack(per_pass), labels)
    assert rel(res["mean_logits"].cpu(), r["mean_logits"]) < 2e-2
    assert abs(res["nll"] - r["nll"]) / abs(r["nll"]) < 2e-2 and abs(res["ece"] - r["ece"]) < 2e-2
    assert not torch.equal(per_pass[0], per_pass[1])


def _build_dist(pkg, gold, cuda):
    from functools import partial
    from oracle import vit_oracle as O
    from uncertainty_vit_b200 import modeling_dist as MD
    a = gold["arch"]
    arch = O.Arch(**a)
    kw = dict(img_size=a["img_size"], patch_size=16, embed_dim=a["embed_dim"], depth=a["depth"], num_heads=a["num_heads"], mlp_ratio=4,
              qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), use_shared_rel_pos_bias=True, use_abs_pos_emb=False,
              init_values=0.1, drop_path_rate=gold.get("dpr", 0.0), attn_drop_rate=gold.get("attn_drop", 0.0))
    model = MD.DistVisionTransformerForCyclicalTraining(**kw) if arch.kind == "cyclical" else MD.DistVisionTransformer(num_classes=a["num_classes"], **kw)
    sd = O.make_state(arch, gold["seed"])
    model.load_state_dict(sd, strict=True)            # identical names/shapes incl. the unused cov_qkv.weight
    return model.to(cuda), arch, sd


def test_dist_finetune_forward_against_reference_golden(pkg, cuda, golden_dir):
    """BASELINE.json configs[0]: dual-stream (--stochastic) fine-tune-mode forward, B=8, ViT-B/16, 1000 classes — and the tiny case."""
    for name in ("tiny_dist_finetune", "vitb_dist_finetune_b8"):
        gold = torch.load(os.path.join(golden_dir, name + ".pt"))
        model, arch, sd = _build_dist(pkg, gold, cuda)
        if gold["x"] is None:
            x = torch.randn(gold["B"], 3, arch.img_size, arch.img_size, generator=torch.Generator().manual_seed(gold["seed"] + 1))
        else:
            x = gold["x"]
        model.eval()
        with torch.no_grad():
            mean_feat, cov_feat, logits = model(x.to(cuda))
        assert rel(mean_feat.cpu(), gold["mean_feat"]) < 2e-2, name
        assert rel(cov_feat.cpu(), gold["cov_feat"]) < 2e-2, name
        assert rel(logits.cpu(), gold["logits"]) < 2e-2, name
        assert torch.equal(logits.argmax(1).cpu(), gold["logits"].argmax(1)) or name.startswith("tiny")


def test_dist_cyclical_forward_against_reference_golden(pkg, cuda, golden_dir):
    gold = torch.load(os.path.join(golden_dir, "tiny_dist_cyclical.pt"))
    model, arch, sd = _build_dist(pkg, gold, cuda)
    x, mask = gold["x"].to(cuda), gold["mask"].to(cuda)
    model.eval()
    with torch.no_grad():
        lm, lc = model(x, None, return_all_tokens=True, layer_results="end")
    from oracle import vit_oracle as O
    tgt = O.build_targets([t.cpu() for t in lm], gold["target_layers"], gold["mask"], post_target_layer_norm=True)
    ctgt = O.build_targets([t.cpu() for t in lc], gold["target_layers"], gold["mask"], post_target_layer_norm=True)
    assert rel(tgt, gold["targets"]) < 2e-2 and rel(ctgt, gold["cov_targets"]) < 2e-2
    model.train()
    n = gold["noise"]
    model.inject_noise(drop_path_keep=n["keep"], attn_keep=n["attn_keep"])
    om, oc = model(x, mask)
    assert rel(om.detach().cpu(), gold["outputs"]) < 2e-2 and rel(oc.detach().cpu(), gold["cov_outputs"]) < 2e-2
    # loss of the --stochastic engine (engine_for_cyclical.py:145-158): smooth-L1 + WassersteinLoss, torch autograd on top of the CUDA network
    tg, ctg = gold["targets"].to(cuda), gold["cov_targets"].to(cuda)
    loss = torch.nn.functional.smooth_l1_loss(om.float(), tg, beta=2.0)
    a, b_, g_, h_ = (torch.sigmoid(t) for t in (om.float(), oc.float(), tg, ctg))
    w = ((a - g_) ** 2).sum(-1) + ((torch.sqrt(torch.clamp(b_, min=1e-24)) - torch.sqrt(torch.clamp(h_, min=1e-24))) ** 2).sum(-1)
    w = w / w.abs().max()
    l2 = -torch.log(torch.sigmoid(-w + 1e-24))
    wl = (l2 / l2.abs().max()).sum() * gold["lam"]
    assert abs(float(wl) - gold["wloss"]) / abs(gold["wloss"]) < 3e-2
    (loss + wl).backward()
    bad = []
    for name, p in model.named_parameters():
        dig = gold["grads"][name]
        if dig is None:
            assert p.grad is None and name.endswith("cov_qkv.weight")        # the only parameter without a gradient
            continue
        e = abs(float(p.grad.double().norm()) - dig["norm"]) / max(dig["norm"], 1e-12)
        if e > 5e-2:
            bad.append((name, round(e, 4), dig["norm"]))
    assert not bad, bad


def test_finetune_training_backward_vs_oracle(pkg, cuda, golden_dir):
    """Fine-tune TRAIN step (cross-entropy on the classifier; + WassersteinLossFineTuning on the dual-stream features) against the
    oracle's autograd on the same weights, inputs and injected noise: det and dual-stream."""
    from oracle import vit_oracle as O
    for name, builder in (("tiny_det_finetune", _build_from_gold), ("tiny_dist_finetune", lambda p, g, c: _build_dist(p, g, c))):
        gold = dict(torch.load(os.path.join(golden_dir, name + ".pt")), dpr=0.2, attn_drop=0.1)
        model, arch, sd = builder(pkg, gold, cuda)
        B = gold["B"]
        g = torch.Generator().manual_seed(5)
        probs = [float(x) for x in torch.linspace(0, 0.2, arch.depth)]
        draws = 4 if arch.dist else 2
        keeps = [(torch.rand(draws, B, generator=g) >= p).float() for p in probs]
        akeep = [(torch.rand(B, arch.num_heads, arch.tokens, arch.tokens, generator=g) >= 0.1).to(torch.uint8) for _ in range(arch.depth)]
        labels = torch.tensor([1, 3, 5])
        noise = O.Noise(drop_path_keep=keeps, drop_path_prob=probs, attn_keep=[k.float() for k in akeep], attn_drop=0.1)
        sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
        out = O.finetune_forward(sdg, arch, gold["x"], noise=noise)
        model.train()
        model.inject_noise(drop_path_keep=keeps, attn_keep=akeep)
        mine = model(gold["x"].to(cuda))
        if arch.dist:
            fm, fc, logits = out
            loss_o = torch.nn.functional.cross_entropy(logits, labels) + O.wasserstein_loss(fm, fc, fm.detach().roll(1, 0), fc.detach().roll(1, 0), 1e-2)
            mfm, mfc, mlog = mine
            assert rel(mlog.detach().cpu(), logits.detach()) < 2e-2 and rel(mfc.detach().cpu(), fc.detach()) < 2e-2
            a, b_, g_, h_ = (torch.sigmoid(t) for t in (mfm.float(), mfc.float(), mfm.detach().roll(1, 0).float(), mfc.detach().roll(1, 0).float()))
            w = ((a - g_) ** 2).sum(-1) + ((torch.sqrt(torch.clamp(b_, min=1e-24)) - torch.sqrt(torch.clamp(h_, min=1e-24))) ** 2).sum(-1)
            w = w / w.abs().max()
            l2 = -torch.log(torch.sigmoid(-w + 1e-24))
            loss_m = torch.nn.functional.cross_entropy(mlog.float(), labels.to(cuda)) + (l2 / l2.abs().max()).sum() * 1e-2
        else:
            loss_o = torch.nn.functional.cross_entropy(out, labels)
            assert rel(mine.detach().cpu(), out.detach()) < 2e-2
            loss_m = torch.nn.functional.cross_entropy(mine.float(), labels.to(cuda))
        assert abs(float(loss_m) - float(loss_o)) / abs(float(loss_o)) < 2e-2
        loss_o.backward()
        loss_m.backward()
        bad = []
        for pname, p in model.named_parameters():
            go = sdg[pname].grad
            if go is None:
                assert p.grad is None and pname.endswith("cov_qkv.weight")
                continue
            e = rel(p.grad.cpu(), go)
            if e > 8e-2 and float(go.norm()) > 1e-6:
                bad.append((pname, round(e, 4)))
        assert not bad, (name, bad)
