#!/usr/bin/env python
"""Runs ONE GEMM shape a few times (for ncu): python tools/gemm_one.py <name>  with name in qkv|proj|fc1|fc2|dgelu|wgrad"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uncertainty_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = torch.device("cuda:0")
M = 25216
name = sys.argv[1] if len(sys.argv) > 1 else "fc1"
cfg = {"qkv": (M, 2304, 768, False, False, ops.EPI_BF16), "proj": (M, 768, 768, False, False, ops.EPI_RESIDUAL),
       "fc1": (M, 3072, 768, False, False, ops.EPI_GELU), "fc2": (M, 768, 3072, False, False, ops.EPI_RESIDUAL),
       "dgelu": (M, 3072, 768, False, True, ops.EPI_DGELU), "wgrad": (3072, 768, M, True, True, ops.EPI_F32_ATOMIC)}[name]
m, n, k, a_mn, b_mn, epi = cfg
a = torch.randn((k, m) if a_mn else (m, k), device=dev).to(torch.bfloat16)
b = torch.randn((k, n) if b_mn else (n, k), device=dev).to(torch.bfloat16)
o16 = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
o16b = torch.randn(m, n, device=dev).to(torch.bfloat16)
o32 = torch.zeros(m, n, dtype=torch.float32, device=dev)
res = torch.randn(m, n, device=dev)
bias = torch.randn(n, device=dev)
for _ in range(3):
    if epi == ops.EPI_BF16:
        ops.gemm(a, b, m, n, k, bias=bias, out_bf16=o16, epilogue=epi)
    elif epi == ops.EPI_GELU:
        ops.gemm(a, b, m, n, k, bias=bias, out_bf16=o16, out2_bf16=o16b, epilogue=epi)
    elif epi == ops.EPI_DGELU:
        ops.gemm(a, b, m, n, k, b_mn=True, aux=o16b, out_bf16=o16, epilogue=epi)
    elif epi == ops.EPI_RESIDUAL:
        ops.gemm(a, b, m, n, k, bias=bias, colscale=bias, residual=res, out_f32=o32, out2_bf16=o16, epilogue=epi)
    else:
        ops.gemm(a, b, m, n, k, a_mn=True, b_mn=True, out_f32=o32, epilogue=epi)
torch.cuda.synchronize()
print("ok", name)
