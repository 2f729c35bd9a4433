#!/usr/bin/env python
"""CUDA-event timing of the row kernels of the step at its shapes (M = 25 216 rows, C = 768; rotating buffers > L2): algorithmic GB/s per kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uncertainty_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = torch.device("cuda:0")
M, C, T = 25216, 768, 197
R = 4


def timeit(fn, n=40):
    for i in range(4):
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
    bf = torch.bfloat16
    x = [torch.randn(M, C, device=dev) for _ in range(R)]
    dx = [torch.randn(M, C, device=dev) for _ in range(R)]
    dy = [torch.randn(M, C, device=dev).to(bf) for _ in range(R)]
    t = [torch.randn(M, C, device=dev).to(bf) for _ in range(R)]
    dt = [torch.empty(M, C, dtype=bf, device=dev) for _ in range(R)]
    y = [torch.empty(M, C, dtype=bf, device=dev) for _ in range(R)]
    g, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    dg, db, dg2, db2 = (torch.zeros(C, device=dev) for _ in range(4))
    rs = torch.ones(M // T, device=dev)
    rows = []

    def rec(name, us, nbytes):
        rows.append((name, us, nbytes / us * 1e-3))
        print(f"{name:46s} {us:7.1f} us  {nbytes / 1e6:7.1f} MB  {nbytes / us * 1e-3:7.0f} GB/s", flush=True)

    # library reference points for the same tensor sizes: how fast do 2, 3, 4-stream element-wise passes run on this HBM?
    us = timeit(lambda i: dx[i % R].copy_(x[i % R]))
    rec("torch copy_ fp32 (1R + 1W)", us, M * C * 8)
    us = timeit(lambda i: torch.add(x[i % R], dx[i % R], out=x[(i + 1) % R]))
    rec("torch add fp32 (2R + 1W)", us, M * C * 12)
    us = timeit(lambda i: torch.addcmul(x[i % R], dx[i % R], dx[(i + 1) % R], out=x[(i + 2) % R]))
    rec("torch addcmul fp32 (3R + 1W)", us, M * C * 16)
    us = timeit(lambda i: ops.layernorm_fwd(x[i % R], g, b, 1e-6, M, C, y_bf16=y[i % R], mean=mean, rstd=rstd))
    rec("ln_fwd (fp32 in, bf16 out)", us, M * C * 6)
    us = timeit(lambda i: ops.layernorm_bwd(dy[i % R], x[i % R], g, mean, rstd, M, C, dx[i % R], dg, db))
    rec("ln_bwd (dy bf16, x fp32, dx fp32 rmw)", us, M * C * (2 + 4 + 4 + 4))
    us = timeit(lambda i: ops.scale_residual_bwd(dx[i % R], t[i % R], rs, T, g, M, C, dt[i % R], dg2, db2))
    rec("scale_residual_bwd (dx fp32, t bf16, dt bf16)", us, M * C * (4 + 2 + 2))
    us = timeit(lambda i: ops.layernorm_bwd(dy[i % R], x[i % R], g, mean, rstd, M, C, dx[i % R], None, None))
    rec("ln_bwd without dgamma / dbeta", us, M * C * (2 + 4 + 4 + 4))
    us = timeit(lambda i: ops.scale_residual_bwd(dx[i % R], t[i % R], rs, T, g, M, C, dt[i % R], None, None))
    rec("scale_residual_bwd without dgamma / dbias", us, M * C * (4 + 2 + 2))
    os.environ["B200VIT_FUSED_LN"] = "1"
    us = timeit(lambda i: ops.layernorm_bwd_scale_residual(dy[i % R], x[i % R], g, mean, rstd, M, C, dx[i % R], dg, db, t[i % R], rs, T, g, dt[i % R], dg2, db2))
    rec("ln_bwd + scale_residual_bwd fused", us, M * C * (2 + 4 + 4 + 4 + 2 + 2))


if __name__ == "__main__":
    main()
