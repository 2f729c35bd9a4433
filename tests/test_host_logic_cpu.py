"""CPU: host-side logic of the engines against the oracle (which is pinned to the reference): label smoothing as soft targets, the
layer-decay group assignment, the lr schedule and the pass sharding."""
import math

import numpy as np
import pytest
import torch


def test_smoothed_targets_are_label_smoothing_cross_entropy():
    """The fine-tune engine turns index labels into (1 - s) one-hot + s / K soft targets and runs the soft-target CE kernel: that is
    timm's LabelSmoothingCrossEntropy(s) (run_class_finetuning.py:619-624), (1 - s) * nll + s * mean(-log p)."""
    from oracle import vit_oracle as O
    g = torch.Generator().manual_seed(1)
    z = torch.randn(5, 17, generator=g)
    hard = torch.tensor([3, 0, 16, 7, 7])
    for s_ in (0.0, 0.1):
        logp = torch.log_softmax(z, -1)
        expect = ((1 - s_) * -logp.gather(1, hard[:, None]).squeeze(1) + s_ * -logp.mean(-1)).mean()
        got = O.soft_target_cross_entropy(z, O.smoothed_targets(hard, 17, s_))
        assert abs(float(got) - float(expect)) < 1e-6
    assert torch.allclose(O.smoothed_targets(hard, 17, 0.1), O.mixup_target(hard, 17, lam=1.0, smoothing=0.1))


def test_layer_ids_match_oracle_for_every_reference_parameter_name():
    from oracle import vit_oracle as O
    from uncertainty_vit_b200 import engine as E
    arch = O.Arch(kind="finetune", dist=True, **O.VIT_B) if hasattr(O, "VIT_B") else None
    names = list(O.make_state(arch, 0).keys()) if arch is not None else []
    names += ["cls_token", "mask_token", "pos_embed", "patch_embed.proj.weight", "rel_pos_bias.relative_position_bias_table", "blocks.0.norm1.weight",
              "blocks.11.mlp.fc2.bias", "fc_norm.weight", "head.weight", "cov_cls_token", "cov_patch_embed.proj.bias"]
    L = 12 + 2
    for n in names:
        assert E.get_num_layer_for_vit(n, L) == O.get_num_layer_for_vit(n, L), n
    # LayerDecayValueAssigner(0.65) of run_class_finetuning.py:569-573: scale = 0.65 ** (L - 1 - layer_id)
    assert math.isclose(0.65 ** (L - 1 - E.get_num_layer_for_vit("blocks.0.attn.qkv.weight", L)), 0.65 ** 12)
    assert E.get_num_layer_for_vit("head.bias", L) == L - 1


def test_cosine_scheduler_shape_and_endpoints():
    from uncertainty_vit_b200 import engine as E
    s = E.cosine_scheduler(2e-3, 1e-5, 10, 7, warmup_epochs=2)
    assert len(s) == 70 and s[0] == 0.0 and abs(s[14] - 2e-3) < 1e-12 and s[-1] > 1e-5 and np.all(np.diff(s[14:]) <= 1e-15)
    s2 = E.cosine_scheduler(1.0, 0.0, 4, 5, warmup_steps=3)
    assert len(s2) == 20 and abs(s2[2] - 1.0) < 1e-12


def test_shard_passes_partitions_every_pass_once():
    from uncertainty_vit_b200 import mc
    for S in (2, 7, 30, 31):
        for world in (1, 2, 3, 8):
            sh = mc.shard_passes(S, world)
            assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [e - s for s, e in sh]
            assert max(sizes) - min(sizes) <= 1


def test_mixup_parameter_draws_follow_timm_call_order():
    """Host side of the device Mixup: (lam, cutmix?, box) per batch == the oracle's restatement of timm 0.3.2 Mixup._params_per_batch /
    cutmix_bbox_and_lam on the same numpy RandomState, for the README recipe (--mixup 0.8 --cutmix 1.0) and the single-mode settings."""
    import numpy as np
    from oracle import vit_oracle as O
    from uncertainty_vit_b200 import mixup as MX
    shape = (8, 3, 224, 224)
    for ma, ca, prob in ((0.8, 1.0, 1.0), (0.8, 0.0, 1.0), (0.0, 1.0, 0.7)):
        mine = MX.Mixup(mixup_alpha=ma, cutmix_alpha=ca, prob=prob, num_classes=10, rng=np.random.RandomState(5))
        r = np.random.RandomState(5)
        seen_cut = seen_mix = seen_off = 0
        for _ in range(300):
            a = mine.draw(shape)
            b = O.mixup_draw(r, shape, mixup_alpha=ma, cutmix_alpha=ca, prob=prob)
            assert a == b
            lam, cut, (yl, yh, xl, xh) = a
            seen_cut += cut and lam != 1.0
            seen_mix += (not cut) and lam != 1.0
            seen_off += lam == 1.0
            if cut and lam != 1.0:
                assert 0 <= yl <= yh <= 224 and 0 <= xl <= xh <= 224 and abs(lam - (1 - (yh - yl) * (xh - xl) / 224 ** 2)) < 1e-12
        assert (seen_cut > 0) == (ca > 0) and (seen_mix > 0) == (ma > 0) and (seen_off > 0) == (prob < 1)
    with pytest.raises(NotImplementedError):
        MX.Mixup(mode="elem")
    t = O.mixup_target(torch.tensor([1, 3, 3, 0]), 5, 0.3, 0.1)
    assert torch.allclose(t.sum(1), torch.ones(4)) and abs(float(t[0, 1]) - (0.3 * 0.92 + 0.7 * 0.02)) < 1e-6


def test_masking_generator_host_interface():
    """The device generator keeps the reference class's host-visible interface (masking_generator.py:29-53): constructor defaults, repr,
    get_shape; plus the image counter that makes a resumed run continue its Philox stream."""
    from uncertainty_vit_b200 import masking_generator as MG
    g = MG.MaskingGenerator((14, 14), num_masking_patches=120, max_num_patches=None, min_num_patches=16, device="cpu")
    assert repr(g) == "Generator(14, 14 -> [16 ~ 120], max = 120, -1.204 ~ 1.204)"
    assert g.get_shape() == (14, 14) and g.num_patches == 196 and g.max_num_patches == 120
    assert g.log_aspect_ratio == (math.log(0.3), math.log(1 / 0.3))
    h = MG.MaskingGenerator(12, 75, min_num_patches=4, max_num_patches=40, min_aspect=0.5, max_aspect=3.0, seed=7, device="cpu")
    assert h.get_shape() == (12, 12) and h.max_num_patches == 40 and h.log_aspect_ratio == (math.log(0.5), math.log(3.0))
    h.images_drawn = 4096
    k = MG.MaskingGenerator(12, 75, device="cpu")
    k.load_state_dict(h.state_dict())
    assert (k.seed, k.images_drawn) == (7, 4096)
    with pytest.raises(Exception):        # no CPU path: drawing needs the CUDA library and CUDA tensors
        g.batch(2)
