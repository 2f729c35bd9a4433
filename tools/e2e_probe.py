"""Where does the end-to-end step time go? (host-side timing of stage_host / step_staged with the CUDA-graph engine)"""
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
host = [bench.synth_batch(128, i) for i in range(2)]
for i in range(4):
    eng.step_host(*host[i % 2])
torch.cuda.synchronize()
ts, tl, ti = [], [], []
nxt = eng.stage_host(*host[0])
for i in range(10):
    t0 = time.perf_counter()
    cur, nxt = nxt, eng.stage_host(*host[(i + 1) % 2])
    t1 = time.perf_counter()
    images, mask_u8, rows, ev, _k = cur
    c = torch.cuda.current_stream(dev); c.wait_event(ev)
    loss = eng.step(images, mask_u8, rows)
    t2 = time.perf_counter()
    v = float(loss.item())
    t3 = time.perf_counter()
    ts.append(t1 - t0); tl.append(t2 - t1); ti.append(t3 - t2)
print(f"stage_host host {1e3*np.mean(ts):.2f} ms | step launch {1e3*np.mean(tl):.2f} ms | wait for loss {1e3*np.mean(ti):.2f} ms | total {1e3*(np.mean(ts)+np.mean(tl)+np.mean(ti)):.2f}")
# same without staging the next batch (inputs resident)
xs = [(x.to(dev), torch.from_numpy(np.ascontiguousarray(m.reshape(128, -1)).reshape(-1)).to(dev), torch.from_numpy(eng.rows_from_host_mask(np.ascontiguousarray(m.reshape(128, -1)), 197)).to(dev)) for x, m in host]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10):
    float(eng.step(*xs[i % 2]).item())
print(f"resident inputs, loss read every step: {1e2*(time.perf_counter()-t0):.2f} ms/step")
