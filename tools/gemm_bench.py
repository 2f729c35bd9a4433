#!/usr/bin/env python
"""Times the tcgen05 GEMM on the step's shapes (CUDA events, rotating buffers > L2) and prints TFLOP/s per shape/epilogue."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uncertainty_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = torch.device("cuda:0")
M = 25216


def run(name, m, n, k, a_mn=False, b_mn=False, epi=ops.EPI_BF16, reps=20, nbuf=3, no_out2=False, **kw):
    bufs = []
    for _ in range(nbuf):
        a = torch.randn((k, m) if a_mn else (m, k), device=dev).to(torch.bfloat16)
        b = torch.randn((k, n) if b_mn else (n, k), device=dev).to(torch.bfloat16)
        o16 = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        o16b = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        o32 = torch.zeros(m, n, dtype=torch.float32, device=dev)
        res = torch.randn(m, n, device=dev)
        bufs.append((a, b, o16, o16b, o32, res))
    bias = torch.randn(n, device=dev)

    def call(i):
        a, b, o16, o16b, o32, res = bufs[i % nbuf]
        args = dict(a_mn=a_mn, b_mn=b_mn, epilogue=epi, **kw)
        if epi in (ops.EPI_BF16, ops.EPI_ELU1):
            ops.gemm(a, b, m, n, k, bias=bias, out_bf16=o16, **args)
        elif epi == ops.EPI_GELU:
            ops.gemm(a, b, m, n, k, bias=bias, out_bf16=o16, out2_bf16=None if no_out2 else o16b, **args)
        elif epi == ops.EPI_DGELU:
            ops.gemm(a, b, m, n, k, aux=o16b, out_bf16=o16, **args)
        elif epi == ops.EPI_RESIDUAL:
            ops.gemm(a, b, m, n, k, bias=bias, colscale=bias, residual=res, out_f32=o32, out2_bf16=None if no_out2 else o16, **args)
        else:
            ops.gemm(a, b, m, n, k, out_f32=o32, **args)
    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:34s} M={m:6d} N={n:5d} K={k:6d}  {us:8.1f} us  {2.0 * m * n * k / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if "--quick" in sys.argv:
        run("fc1 fwd gelu (teacher: no out2)", M, 3072, 768, epi=ops.EPI_GELU, no_out2=True)
        run("fc1 fwd gelu + gelu' (student)", M, 3072, 768, epi=ops.EPI_GELU)
        run("fc2 dgrad dgelu (acc * aux)", M, 3072, 768, b_mn=True, epi=ops.EPI_DGELU)
        run("fc1 dgrad bf16", M, 768, 3072, b_mn=True)
        run("proj fwd residual", M, 768, 768, epi=ops.EPI_RESIDUAL)
        run("proj fwd residual, no bf16 branch output", M, 768, 768, epi=ops.EPI_RESIDUAL, no_out2=True)
        run("proj shape, fp32 store only (EPI_F32)", M, 768, 768, epi=ops.EPI_F32)
        run("proj shape, bf16 store only (EPI_BF16)", M, 768, 768)
        run("fc2 fwd residual", M, 768, 3072, epi=ops.EPI_RESIDUAL)
        run("qkv fwd bf16", M, 2304, 768)
        run("cov qkv fwd elu+1", M, 2304, 768, epi=ops.EPI_ELU1)
        sys.exit(0)
    for skip in (0, 1):
        tag = " [no epilogue I/O]" if skip else ""
        kw = dict(debug_flags=skip)
        run("qkv fwd bf16" + tag, M, 2304, 768, **kw)
        run("proj fwd residual" + tag, M, 768, 768, epi=ops.EPI_RESIDUAL, **kw)
        run("fc1 fwd gelu" + tag, M, 3072, 768, epi=ops.EPI_GELU, **kw)
        run("fc2 fwd residual" + tag, M, 768, 3072, epi=ops.EPI_RESIDUAL, **kw)
        run("fc2 dgrad dgelu" + tag, M, 3072, 768, b_mn=True, epi=ops.EPI_DGELU, **kw)
        run("fc1 dgrad bf16" + tag, M, 768, 3072, b_mn=True, **kw)
        run("qkv dgrad bf16" + tag, M, 768, 2304, b_mn=True, **kw)
        run("fc1 wgrad atomic" + tag, 3072, 768, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, **kw)
        run("fc2 wgrad atomic" + tag, 768, 3072, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, **kw)
        run("square 8192 bf16" + tag, 8192, 8192, 8192, reps=5, nbuf=1, **kw)
        if not skip:
            for sk in (1, 2, 3, 4, 6, 8, 12):
                run(f"fc1 wgrad split_k={sk}", 3072, 768, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, split_k=sk)
            for sk in (2, 4, 6, 8, 12, 16):
                run(f"proj wgrad split_k={sk}", 768, 768, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, split_k=sk)
            for sk in (1, 2, 3, 4, 6, 8):
                run(f"qkv wgrad split_k={sk}", 2304, 768, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, split_k=sk)
