"""CUDA-event timing of the input-pipeline kernels at the step's shapes (rotating buffers larger than L2):
device Mixup / CutMix (HBM-bound: algorithmic bytes = 8 B per mixed element) and the block-wise mask generator (latency-bound: one
thread per image).

    python tools/aux_bench.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uncertainty_vit_b200 as pkg  # noqa: E402
from uncertainty_vit_b200 import masking_generator as MG  # noqa: E402

ops = pkg.ops
dev = torch.device("cuda:0")


def timeit(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3          # us


def main():
    out = {}
    B = 128
    bufs = [torch.randn(B, 3, 224, 224, device=dev) for _ in range(4)]            # 4 x 77 MB > 126 MB L2
    labels = torch.randint(0, 1000, (B,), device=dev)
    nbytes = bufs[0].numel() * 4
    us = timeit(lambda i: ops.mixup_batch(bufs[i % 4], 0.4, False, (0, 0, 0, 0), labels, 1000, 0.9001, 0.0001))
    out["mixup_b128"] = {"us": us, "algorithmic_GBps": 2 * nbytes / us * 1e-3, "bytes": 2 * nbytes}
    us = timeit(lambda i: ops.mixup_batch(bufs[i % 4], 0.5, True, (32, 192, 40, 180), labels, 1000, 0.9001, 0.0001))
    box = 3 * 160 * 140 * 4 * B
    out["cutmix_b128_box160x140"] = {"us": us, "algorithmic_GBps": 2 * box / us * 1e-3, "bytes": 2 * box}
    pix = [torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(4)]
    us = timeit(lambda i: ops.normalize_u8(pix[i % 4], (0.5, 0.5, 0.5), (0.5, 0.5, 0.5), True, out=bufs[i % 4]))
    out["normalize_u8_b128"] = {"us": us, "algorithmic_GBps": (pix[0].numel() + nbytes) / us * 1e-3, "bytes": pix[0].numel() + nbytes}
    for nb in (128, 4096):
        gen = MG.MaskingGenerator(14, 120, min_num_patches=16, seed=1, device=dev)
        us = timeit(lambda i: gen.batch(nb))
        out[f"block_masks_b{nb}"] = {"us_incl_row_list_and_allocs": us, "images_per_s": nb / us * 1e6}
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        pk = json.load(open(peaks))
        out["hbm_peak"] = {k: v for k, v in pk.items() if "hbm" in k.lower() or "copy" in k.lower()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
