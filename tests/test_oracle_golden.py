"""CPU: the oracle restatement (oracle/vit_oracle.py) against the golden vectors produced by the REAL reference
(tools/make_golden.py, run where /root/reference exists). fp32, tolerance 2e-5 relative."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import vit_oracle as O


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _noise(gold):
    n = gold["noise"]
    return O.Noise(drop_path_keep=n["keep"], drop_path_prob=n["prob"], attn_keep=[k.float() for k in n["attn_keep"]], attn_drop=n["attn_drop"])


@pytest.mark.parametrize("name", ["tiny_det_cyclical", "tiny_dist_cyclical"])
def test_cyclical_matches_reference(golden_dir, name):
    gold = torch.load(os.path.join(golden_dir, name + ".pt"))
    arch = O.Arch(**gold["arch"])
    sd = O.make_state(arch, gold["seed"])
    chk = float(sum(v.double().sum() for v in sd.values() if v.is_floating_point()))
    assert abs(chk - gold["state_checksum"]) < 1e-6 * max(1.0, abs(chk)), "seeded weights drifted (torch RNG changed?)"
    x, mask = gold["x"], gold["mask"]
    with torch.no_grad():
        t = O.cyclical_forward(sd, arch, x, None, return_all_tokens=True, layer_results="end")
    if arch.dist:
        t, tc = t
    for a, b in zip(t, gold["teacher_layers"]):
        assert rel(a, b) < 2e-5
    tgt = O.build_targets(t, gold["target_layers"], mask, post_target_layer_norm=True)
    assert rel(tgt, gold["targets"]) < 2e-5
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    out = O.cyclical_forward(sdg, arch, x, mask, noise=_noise(gold))
    if arch.dist:
        out, cout = out
        assert rel(cout, gold["cov_outputs"]) < 2e-5
        ctgt = O.build_targets(tc, gold["target_layers"], mask, post_target_layer_norm=True)
        assert rel(ctgt, gold["cov_targets"]) < 2e-5
    assert out.shape[0] == int(mask.sum()) and rel(out, gold["outputs"]) < 2e-5
    loss, _ = O.d2v_loss(out, tgt, 2.0)
    assert abs(float(loss) - gold["loss"]) < 2e-5 * abs(gold["loss"])
    total = loss + (O.wasserstein_loss(out, cout, tgt, ctgt, gold["lam"]) if arch.dist else 0.0)
    assert abs(float(total) - gold["total_loss"]) < 2e-5 * abs(gold["total_loss"])
    total.backward()
    for k, dig in gold["grads"].items():
        g = sdg[k].grad
        if dig is None:          # cov_qkv.weight: allocated but unused by the reference
            assert g is None or float(g.abs().max()) == 0.0
            assert k.endswith("cov_qkv.weight")
            continue
        assert abs(float(g.double().norm()) - dig["norm"]) < 1e-4 * max(dig["norm"], 1e-12), k
        assert rel(g.flatten()[:64], dig["head"]) < 1e-4 or float(dig["head"].abs().max()) < 1e-12, k


@pytest.mark.parametrize("name", ["tiny_det_finetune", "tiny_dist_finetune"])
def test_finetune_matches_reference(golden_dir, name):
    gold = torch.load(os.path.join(golden_dir, name + ".pt"))
    arch = O.Arch(**gold["arch"])
    sd = O.make_state(arch, gold["seed"])
    with torch.no_grad():
        out = O.finetune_forward(sd, arch, gold["x"])
    if arch.dist:
        assert rel(out[0], gold["mean_feat"]) < 2e-5 and rel(out[1], gold["cov_feat"]) < 2e-5 and rel(out[2], gold["logits"]) < 2e-5
    else:
        assert rel(out, gold["logits"]) < 2e-5


def test_metrics_match_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "metrics.pt"))
    r = O.mc_reduce(gold["logits"], gold["labels"])
    assert abs(r["ece_reference"] - gold["ece_reference"]) < 1e-7       # what the reference prints (indexing quirk)
    assert abs(r["ece"] - gold["ece"]) < 1e-7
    assert abs(r["nll"] - gold["nll"]) < 1e-6 and abs(r["acc1"] - gold["acc1"]) < 1e-4 and abs(r["acc5"] - gold["acc5"]) < 1e-4
    t = gold["w_inputs"]
    assert abs(float(O.wasserstein_loss(*t[:4], 1e-5)) - gold["wloss"]) < 1e-9
    assert abs(float(O.wasserstein_loss_finetune(*t, 1e-4, 1e-4)) - gold["wloss_ft"]) < 1e-9
    assert rel(O.wasserstein_distance_matmul(t[0][None], t[1][None], t[2][None], t[3][None]), gold["wdm"]) < 1e-6
    # TACELoss().loss(probs, labels, logits=False) of the reference (the uint8 fancy-index quirk included) + the documented value
    probs = torch.softmax(gold["logits"].float().mean(0), 1)
    assert abs(O.tace(probs, gold["labels"], reference_indexing=True) - gold["tace_reference"]) < 1e-9
    assert abs(O.tace(probs, gold["labels"]) - gold["tace"]) < 1e-9
    p2 = torch.softmax(gold["tace2_logits"], 1)
    assert abs(O.tace(p2, gold["tace2_labels"], reference_indexing=True) - gold["tace2_reference"]) < 1e-9


def test_auroc_restatement_matches_sklearn(golden_dir):
    """torchmetrics is not in the image (parity unpinned against it): the one-vs-rest macro AUROC restatement agrees with scikit-learn's."""
    from sklearn.metrics import roc_auc_score
    gold = torch.load(os.path.join(golden_dir, "metrics.pt"))
    p2 = torch.softmax(gold["tace2_logits"], 1)
    y2 = gold["tace2_labels"]
    assert len(set(y2.tolist())) == p2.shape[1]
    assert abs(O.auroc_macro_ovr(p2, y2) - roc_auc_score(y2.numpy(), p2.double().numpy(), multi_class="ovr", average="macro")) < 1e-9
    q = torch.round(p2 * 20) / 20                                   # heavy ties
    q = q / q.sum(1, keepdim=True)
    assert abs(O.auroc_macro_ovr(q, y2) - roc_auc_score(y2.numpy(), q.double().numpy(), multi_class="ovr", average="macro")) < 1e-9


def _loop_noise(n):
    return O.Noise(drop_path_keep=n["keep"], drop_path_prob=n["prob"], attn_keep=[k.float() for k in n["attn_keep"]], attn_drop=n["attn_drop"])


@pytest.mark.parametrize("name", ["tiny_det_loop", "tiny_dist_loop", "tiny_det_loop_variants", "tiny_det_loop_bn"])
def test_step_oracle_matches_the_references_own_training_loop(golden_dir, name):
    """O.d2v_step against engine_for_cyclical.train_one_epoch of the REAL reference run for 2-3 steps (tools/make_golden.py::case_train_loop):
    per-step loss, mean gradient norm, updated weights, EMA teacher — including the teacher's truncated integer index buffer, the
    instance / batch-norm target variants, the var_w0 hinge, loss_scale and the start_lr_decay_at_step EMA cut-off."""
    gold = torch.load(os.path.join(golden_dir, name + ".pt"))
    arch = O.Arch(**gold["arch"])
    sd = O.make_state(arch, gold["seed"])
    ema = {k: v.clone() for k, v in sd.items()}
    opt = O.new_opt_state(sd)
    kw = gold["loop_kw"]
    tk = {k: kw[k] for k in ("target_layer_norm_last", "target_batch_norm", "target_instance_norm", "post_target_instance_norm",
                             "post_target_layer_norm") if k in kw}
    cur_decay, losses, gns, logged_decay = kw["decay"], [], [], []
    for it in range(gold["steps"]):
        if it < kw["ema_start_at"]:
            cur_decay = kw["decay_init"] + it * (kw["decay"] - kw["decay_init"]) / kw["ema_start_at"]
        sl = kw.get("start_lr_decay_at_step", -1)
        upd = cur_decay != 1 and (sl == -1 or it <= sl)
        x, mask = gold["batches"][it]
        lo, gn = O.d2v_step(sd, ema, opt, arch, x, mask, it + 1, _loop_noise(gold["noises"][it]), gold["target_layers"], lr=gold["lr"][it],
                            wd=gold["wd"][it], clip=gold["clip"], ema_decay=cur_decay, l1_beta=kw["l1_beta"], lam=kw["lambda_pretraining"],
                            l2_loss=kw.get("l2_loss", False), target_kwargs=tk, var_w0=kw.get("var_w0", 0), var_margin0=kw.get("var_margin0", 0.5),
                            loss_scale=kw.get("loss_scale", -1), update_ema=upd)
        if not upd:
            cur_decay = 0
        losses.append(lo)
        gns.append(gn)
        logged_decay.append(cur_decay)               # metric_logger.update(cur_decay=...) after the EMA branch (:206)
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 2e-5 * abs(b), (losses, gold["losses"])
    assert abs(float(np.mean(gns)) - gold["grad_norm"]) <= 2e-5 * gold["grad_norm"]
    assert abs(float(np.mean(logged_decay)) - gold["cur_decay"]) < 1e-9        # the loop returns the meters' global averages
    for k, v in gold["weights"].items():
        assert rel(sd[k], v) < 2e-5, k
    for k, v in gold["ema"].items():
        assert rel(ema[k], v) < 2e-5, k
    assert torch.equal(ema["rel_pos_bias.relative_position_index"], gold["ema_index"].long())
    # decay_init 0.9 does corrupt the teacher's index buffer (the quirk is live in this golden), the student's stays intact
    assert not torch.equal(ema["rel_pos_bias.relative_position_index"], sd["rel_pos_bias.relative_position_index"])


def test_finetune_step_oracle_matches_train_class_batch(golden_dir):
    """O.finetune_loss_and_grads against engine_for_finetuning_dist.train_class_batch of the REAL reference (:286-304): loss, logits and every
    parameter gradient, eval-mode positive / negative forwards."""
    gold = torch.load(os.path.join(golden_dir, "tiny_dist_train_class_batch.pt"))
    arch = O.Arch(**gold["arch"])
    sd = O.make_state(arch, gold["seed"])
    n = gold["noise"]
    noise = O.Noise(drop_path_keep=n["keep"], drop_path_prob=n["prob"])
    loss, logits, grads = O.finetune_loss_and_grads(sd, arch, gold["x"], gold["targets"], gold["pos"], gold["neg"], noise, gold["lam_ft"], gold["lam_pvn"])
    assert abs(loss - gold["loss"]) <= 2e-5 * abs(gold["loss"]) and rel(logits, gold["logits"]) < 2e-5
    for k, g in gold["grads_full"].items():
        if g is None:
            assert float(grads[k].abs().max()) == 0.0, k
        else:
            assert rel(grads[k], g) < 5e-5, k


def test_index_and_block_masks_bit_exact(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "index_masks.pt"))
    assert torch.equal(O.relative_position_index(14, 14), gold["index14"].long())
    rng = random.Random(gold["mask_seed"])
    mine = np.stack([O.blockwise_mask(rng) for _ in range(gold["masks"].shape[0])])
    assert np.array_equal(mine, gold["masks"].numpy())
    assert mine.reshape(mine.shape[0], -1).sum(1).max() <= 120 and set(np.unique(mine)) <= {0, 1}


def test_drop_path_rates_and_param_groups():
    # dpr = linspace(0, rate, depth) (modeling_finetune.py:401); layer ids / no-decay rule (optim_factory.py:33-67)
    dpr = [float(x) for x in torch.linspace(0, 0.25, 12)]
    assert dpr[0] == 0.0 and abs(dpr[-1] - 0.25) < 1e-7 and abs(dpr[1] - 0.0227) < 1e-4
    assert O.get_num_layer_for_vit("cls_token", 14) == 0 and O.get_num_layer_for_vit("patch_embed.proj.weight", 14) == 0
    assert O.get_num_layer_for_vit("blocks.3.attn.qkv.weight", 14) == 4 and O.get_num_layer_for_vit("rel_pos_bias.relative_position_bias_table", 14) == 13
    assert O.get_num_layer_for_vit("cov_cls_token", 14) == 13     # cov_* top-level params fall into the LAST group
    assert O.is_no_decay("blocks.0.gamma_1", (768,)) and O.is_no_decay("cls_token", (1, 1, 768)) and not O.is_no_decay("mask_token", (1, 1, 768))


def test_ema_decay_schedule():
    assert O.ema_decay_at(0, 0.999, 0.9998, 0) == 0.9998        # README recipe: --ema_start_at 0
    assert abs(O.ema_decay_at(5, 0.999, 0.9998, 10) - (0.999 + 5 * (0.9998 - 0.999) / 10)) < 1e-12
