#!/usr/bin/env python
"""How long does the HOST need to issue one step (no sync) vs the device time of the step?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import uncertainty_vit_b200  # noqa
from uncertainty_vit_b200 import engine as E, modeling as M
dev = torch.device("cuda:0")
model = M.create_model("beit_base_patch16_224", pretrained=False, drop_path_rate=0.25, drop_rate=0.0, use_shared_rel_pos_bias=True,
                       use_abs_pos_emb=False, init_values=1e-4, attn_drop_rate=0.05).to(dev)
eng = E.D2VEngine(model, target_layers=[6, 7, 8, 9, 10, 11])
x, m = bench.synth_batch(128, 0)
mu8 = np.ascontiguousarray(m.reshape(128, -1))
batch = (x.to(dev), torch.from_numpy(mu8.reshape(-1)).to(dev), torch.from_numpy(eng.rows_from_host_mask(mu8, 197)).to(dev))
for _ in range(3):
    eng.step(*batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    eng.step(*batch)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time per step {1e3 * (t1 - t0) / 10:.2f} ms ; wall per step incl. drain {1e3 * (t2 - t0) / 10:.2f} ms")
