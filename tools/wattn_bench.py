#!/usr/bin/env python
"""CUDA-event timing of the Wasserstein-attention forward (prep + forward kernels) and backward at the --stochastic step's shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uncertainty_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = torch.device("cuda:0")


def timeit(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B, H, N, R = 128, 12, 197, 3
    scale = 64 ** -0.5
    qm = [torch.randn(B, N, 3, H, 64, device=dev).bfloat16() for _ in range(R)]
    qc = [(torch.rand(B, N, 3, H, 64, device=dev) + 0.5).bfloat16() for _ in range(R)]
    om = torch.empty(B, N, H * 64, dtype=torch.bfloat16, device=dev)
    oc = torch.empty_like(om)
    lse = torch.empty(B, H, N, device=dev)
    bits = torch.zeros(B, H, N, 32, dtype=torch.uint8, device=dev)
    bias_f, bias_t = ops.pad_attn_bias(torch.randn(H, N, N, device=dev) * 0.5)
    xw = ops.wattn_workspace(B, H, N, dev)
    for p in (0.0, 0.05):
        us = timeit(lambda i: ops.wattn_fwd(qm[i % R], qc[i % R], bias_f, B, H, N, scale, p, 1, 2, None, om, oc, lse, bits if p > 0 else None, xwork=xw,
                                            keep_ready=p > 0))
        print(f"wattn_fwd (prep + forward, masks pre-drawn) p={p}: {us:8.1f} us", flush=True)
    dom, doc = torch.randn_like(om), torch.randn_like(oc)
    dqm, dqc = torch.empty_like(qm[0]), torch.empty_like(qc[0])
    work = ops.wattn_bwd_workspace(B, H, N, False, dev)
    us = timeit(lambda i: ops.wattn_bwd(qm[0], qc[0], xw, om, oc, dom, doc, lse, bias_t, bits, None, None, B, H, N, scale, 0.05, dqm, dqc, work=work))
    print(f"wattn_bwd (prep + transpose + kv + dX) p=0.05: {us:8.1f} us")


if __name__ == "__main__":
    main()
