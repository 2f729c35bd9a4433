#!/usr/bin/env python
"""One data2vec step (config 2) between cudaProfilerStart/Stop, after 2 warm-up steps:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 6 -o gpurun_out/prof python tools/profile_step.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    B = int(os.environ.get("PROFILE_BATCH", bench.BATCH))
    dev = torch.device("cuda:0")
    import uncertainty_vit_b200  # noqa: F401
    from uncertainty_vit_b200 import engine as E, modeling as M
    torch.manual_seed(0)
    extra = {"stochastic": True} if os.environ.get("PROFILE_STOCHASTIC") else {}      # PROFILE_STOCHASTIC=1: the dual-stream step
    model = M.create_model("beit_base_patch16_224", pretrained=False, drop_path_rate=0.25, drop_rate=0.0, use_shared_rel_pos_bias=True,
                           use_abs_pos_emb=False, init_values=1e-4, attn_drop_rate=0.05, **extra).to(dev)
    eng = E.D2VEngine(model, target_layers=[6, 7, 8, 9, 10, 11], use_graph=False)   # same kernels, launched eagerly (ncu sees plain launches)
    x, m = bench.synth_batch(B, 0)
    mu8 = np.ascontiguousarray(m.reshape(B, -1))
    batch = (x.to(dev), torch.from_numpy(mu8.reshape(-1)).to(dev), torch.from_numpy(eng.rows_from_host_mask(mu8, 197)).to(dev))
    for _ in range(2):
        eng.step(*batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.step(*batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("loss", float(eng.loss_dev.item()))


if __name__ == "__main__":
    main()
