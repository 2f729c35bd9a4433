"""Host-side checkpoint / optimizer-group logic (uncertainty-vit_b200/checkpoint.py) against golden outputs of the reference's own
optim_factory.get_parameter_groups and utils.load_state_dict (tests/golden/host_logic.json, written by tools/make_golden.py)."""
import json
import os
from functools import partial

import numpy as np
import pytest
import torch


def _tiny(kind, dist=False, img_size=None, **extra):
    from oracle import vit_oracle as O
    from uncertainty_vit_b200 import modeling as M, modeling_dist as MD
    a = O.Arch(**{**O.TINY, "kind": kind, "dist": dist})
    kw = dict(img_size=img_size or a.img_size, patch_size=16, embed_dim=a.embed_dim, depth=a.depth, num_heads=a.num_heads, mlp_ratio=4, qkv_bias=True,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), use_shared_rel_pos_bias=True, use_abs_pos_emb=False, init_values=0.1, **extra)
    if kind == "cyclical":
        return (MD.DistVisionTransformerForCyclicalTraining if dist else M.VisionTransformerForCyclicalTraining)(**kw), a
    return (MD.DistVisionTransformer if dist else M.VisionTransformer)(num_classes=a.num_classes, **kw), a


@pytest.fixture(scope="module")
def gold(golden_dir):
    return json.load(open(os.path.join(golden_dir, "host_logic.json")))


@pytest.mark.parametrize("label,kind,dist,layer_decay", [("det_cyclical", "cyclical", False, None), ("dist_cyclical", "cyclical", True, None),
                                                         ("det_finetune_ld", "finetune", False, 0.65), ("dist_finetune_ld", "finetune", True, 0.65)])
def test_parameter_groups_match_reference(gold, label, kind, dist, layer_decay):
    """Group order, membership, weight decay and lr scale == optim_factory.get_parameter_groups on the reference module (this order is the
    parameter numbering of the torch.optim.AdamW state dict stored in checkpoints)."""
    from uncertainty_vit_b200 import checkpoint as CK
    model, a = _tiny(kind, dist)
    named = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    if layer_decay is None:
        mine = CK.parameter_groups(named, 0.05, model.no_weight_decay())
    else:
        L = a.depth + 2
        mine = CK.parameter_groups(named, 0.05, model.no_weight_decay(), L, [layer_decay ** (L - 1 - i) for i in range(L)])
    ref = gold["groups"][label]
    assert [g["params"] for g in mine] == [g["params"] for g in ref]
    assert [g["weight_decay"] for g in mine] == [g["weight_decay"] for g in ref]
    assert np.allclose([g["lr_scale"] for g in mine], [g["lr_scale"] for g in ref], rtol=1e-12)


def test_load_state_dict_messages_and_prefix(gold):
    from uncertainty_vit_b200 import checkpoint as CK
    pre, _ = _tiny("cyclical")
    ckpt = {k: v.clone() for k, v in pre.state_dict().items() if "relative_position_index" not in k}
    del ckpt["blocks.1.mlp.fc2.bias"]
    for label, prefix in (("plain", ""), ("prefix", "module.")):
        ft, _ = _tiny("finetune")
        lines = []
        CK.load_state_dict(ft, {prefix + k: v for k, v in ckpt.items()}, prefix=prefix, log=lines.append)
        assert "\n".join(lines) + "\n" == gold["load"][label]
        assert torch.equal(ft.state_dict()["blocks.0.attn.qkv.weight"], pre.state_dict()["blocks.0.attn.qkv.weight"])
        assert torch.equal(ft.state_dict()["rel_pos_bias.relative_position_bias_table"], pre.state_dict()["rel_pos_bias.relative_position_bias_table"])


def test_prepare_finetune_checkpoint():
    """run_class_finetuning.py:400-517 as a function: model_key selection, head of another shape dropped, index buffers dropped,
    same-window tables untouched, pos_embed resized bicubically with the class token kept."""
    from uncertainty_vit_b200 import checkpoint as CK
    pre, _ = _tiny("cyclical")
    ft, _ = _tiny("finetune")
    sd = {k: v.clone() for k, v in pre.state_dict().items()}
    sd["head.weight"] = torch.zeros(7, 128)                       # a head trained for 7 classes: must go
    sd["head.bias"] = torch.zeros(ft.head.bias.shape[0])         # right shape: stays
    msgs = []
    out = CK.prepare_finetune_checkpoint({"model": sd, "epoch": 3}, ft, log=msgs.append)
    assert msgs == ["Removing key head.weight from pretrained checkpoint"]
    assert "head.weight" not in out and "head.bias" in out
    assert not any("relative_position_index" in k for k in out)
    assert out["rel_pos_bias.relative_position_bias_table"] is sd["rel_pos_bias.relative_position_bias_table"]
    missing, unexpected, errors = CK.load_state_dict(ft, out, log=lambda s: None)
    assert missing == ["fc_norm.weight", "fc_norm.bias", "head.weight"] and not errors
    assert unexpected == ["mask_token", "lm_head.weight", "lm_head.bias", "norm.weight", "norm.bias"]
    # no model_key in the file: the checkpoint itself is the state dict; reinit_final_norm drops the norms
    out2 = CK.prepare_finetune_checkpoint(sd, ft, reinit_final_norm=True, log=lambda s: None)
    assert "norm.weight" not in out2 and "norm.bias" not in out2 and "blocks.0.norm1.weight" in out2
    # pos_embed: 4x4 -> 6x6 grid, class token first
    pe = torch.randn(1, 17, 128, generator=torch.Generator().manual_seed(0))
    new = CK.interpolate_pos_embed(pe, 36, 1)
    expect = torch.nn.functional.interpolate(pe[:, 1:].reshape(1, 4, 4, 128).permute(0, 3, 1, 2), size=(6, 6), mode="bicubic", align_corners=False)
    assert new.shape == (1, 37, 128) and torch.equal(new[:, :1], pe[:, :1]) and torch.equal(new[:, 1:], expect.permute(0, 2, 3, 1).flatten(1, 2))
    assert CK.interpolate_pos_embed(pe, 16, 1) is pe


def test_rel_pos_bias_table_resize():
    """Window 14x14 -> 24x24 (the 224 -> 384 fine-tune case): shape, the three cls rows kept, a table that is an affine function of the
    relative offset stays that function (a cubic spline reproduces it exactly, whatever the geometric source spacing)."""
    from uncertainty_vit_b200 import checkpoint as CK
    heads, s, d = 2, 27, 47                      # (2*14-1), (2*24-1)
    x, dx = CK._geometric_positions(s, d)
    assert len(x) == s and len(dx) == d and abs(x[s // 2]) == 0 and np.all(np.diff(x) > 0) and abs(x[-1] - dx[-1]) < 0.05
    gy, gx = np.meshgrid(x, x, indexing="ij")
    body = torch.tensor(np.stack([0.5 * gx - 0.25 * gy + 1.0, -0.1 * gx + 0.3 * gy], -1).reshape(s * s, heads), dtype=torch.float32)
    extra = torch.randn(3, heads)
    out = CK.interpolate_rel_pos_bias_table(torch.cat([body, extra]), d * d + 3, (24, 24))
    assert out.shape == (d * d + 3, heads) and torch.equal(out[-3:], extra)
    ty, tx = np.meshgrid(dx, dx, indexing="ij")
    expect = np.stack([0.5 * tx - 0.25 * ty + 1.0, -0.1 * tx + 0.3 * ty], -1).reshape(d * d, heads)
    assert np.abs(out[:-3].numpy() - expect).max() < 1e-4
    same = torch.cat([body, extra])
    assert CK.interpolate_rel_pos_bias_table(same, s * s + 3, (14, 14)) is same


def test_latest_checkpoint_scan(tmp_path):
    from uncertainty_vit_b200 import checkpoint as CK
    assert CK.latest_checkpoint(str(tmp_path)) is None
    for n in ("checkpoint-3.pth", "checkpoint-12.pth", "checkpoint-best.pth", "checkpoint-7.pth"):
        (tmp_path / n).write_bytes(b"")
    assert CK.latest_checkpoint(str(tmp_path)) == os.path.join(str(tmp_path), "checkpoint-12.pth")
