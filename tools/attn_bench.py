"""Parity + CUDA-event timing of the attention kernels at the step's shape (B=128, H=12, N=197, d=64; rotating buffers > L2).

    python tools/attn_bench.py            # forward (no dropout / dropout 0.05) and backward
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uncertainty_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = torch.device("cuda:0")


def ref_attn(qkv, bias, keep, p, scale):
    B, N, _, H, D = qkv.shape
    q, k, v = qkv.float().permute(2, 0, 3, 1, 4)
    s = (q * scale) @ k.transpose(-1, -2) + bias
    pr = s.softmax(-1)
    if keep is not None:
        pr = pr * keep.view(B, H, N, N).float() / (1 - p)
    return (pr @ v).transpose(1, 2).reshape(B, N, H * D), torch.logsumexp(s, -1)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def parity():
    for (B, H, N, p) in [(2, 2, 17, 0.0), (2, 3, 197, 0.0), (3, 2, 197, 0.1), (1, 2, 64, 0.25), (2, 2, 208, 0.05), (2, 2, 129, 0.05), (1, 1, 5, 0.0)]:
        g = torch.Generator().manual_seed(N + B)
        qkv = torch.randn(B, N, 3, H, 64, generator=g).bfloat16().to(dev)
        bias = (torch.randn(H, N, N, generator=g) * 0.5).to(dev)
        scale = 64 ** -0.5
        out = torch.full((B, N, H * 64), float("nan"), dtype=torch.bfloat16, device=dev)
        lse = torch.empty(B, H, N, device=dev)
        bits = torch.zeros(B, H, N, 32, dtype=torch.uint8, device=dev)
        bias_f, _ = ops.pad_attn_bias(bias)
        ops.attn_fwd(qkv, bias_f, B, H, N, scale, p, 1234, 7, None, out, lse, bits if p > 0 else None)
        torch.cuda.synchronize()
        keep = ops.dropout_mask(B * H, N, p, 1234, 7, dev) if p > 0 else None
        ref, lse_ref = ref_attn(qkv, bias, keep, p, scale)
        ok_bits = True
        if keep is not None:
            unpacked = ((bits.view(B * H, N, 32, 1) >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).reshape(B * H, N, 256)[:, :, :N]
            ok_bits = bool(torch.equal(unpacked, keep))
        print(f"fwd B={B} H={H} N={N} p={p}: rel(out)={rel(out.float(), ref):.2e} rel(lse)={rel(lse, lse_ref):.2e} finite={bool(torch.isfinite(out.float()).all())} bits={ok_bits}")
        # no-bias variant
        out2 = torch.empty_like(out)
        ops.attn_fwd(qkv, None, B, H, N, scale, 0.0, 0, 0, None, out2, lse, None)
        ref2, _ = ref_attn(qkv, torch.zeros_like(bias), None, 0.0, scale)
        print(f"    no-bias rel(out)={rel(out2.float(), ref2):.2e}")


def parity_bwd():
    for (B, H, N, p) in [(2, 2, 17, 0.0), (2, 3, 197, 0.0), (3, 2, 197, 0.1), (1, 2, 64, 0.25), (2, 2, 208, 0.05), (20, 12, 197, 0.05)]:
        g = torch.Generator().manual_seed(N + B)
        qkv = torch.randn(B, N, 3, H, 64, generator=g).bfloat16().to(dev)
        bias = (torch.randn(H, N, N, generator=g) * 0.5).to(dev)
        scale = 64 ** -0.5
        out = torch.empty((B, N, H * 64), dtype=torch.bfloat16, device=dev)
        lse = torch.empty(B, H, N, device=dev)
        bits = torch.zeros(B, H, N, 32, dtype=torch.uint8, device=dev)
        bias_f, bias_t = ops.pad_attn_bias(bias)
        ops.attn_fwd(qkv, bias_f, B, H, N, scale, p, 1234, 7, None, out, lse, bits if p > 0 else None)
        keep = ops.dropout_mask(B * H, N, p, 1234, 7, dev) if p > 0 else None
        qr = qkv.float().requires_grad_(True)
        br = bias.clone().requires_grad_(True)
        ref, _ = ref_attn(qr, br, keep, p, scale)
        dout = torch.randn(B, N, H * 64, generator=g).bfloat16().to(dev)
        ref.backward(dout.float())
        idx = torch.randint(0, 50, (N, N), generator=g).to(torch.int32).to(dev)
        dtable = torch.zeros(50, H, device=dev)
        dqkv = torch.full((B, N, 3, H, 64), float("nan"), dtype=torch.bfloat16, device=dev)
        dqb = torch.zeros(H * 64, device=dev)
        dvb = torch.zeros(H * 64, device=dev)
        ops.attn_bwd(qkv, out, dout, lse, bias_t, bits if p > 0 else None, idx, dtable, B, H, N, scale, p, dqkv, dq_bias=dqb, dv_bias=dvb)
        torch.cuda.synchronize()
        ref_tab = torch.zeros(50, H, device=dev)
        ref_tab.index_add_(0, idx.long().flatten(), br.grad.permute(1, 2, 0).reshape(N * N, H))
        gq = qr.grad
        print(f"bwd B={B} H={H} N={N} p={p}: dq={rel(dqkv[:, :, 0].float(), gq[:, :, 0]):.2e} dk={rel(dqkv[:, :, 1].float(), gq[:, :, 1]):.2e} "
              f"dv={rel(dqkv[:, :, 2].float(), gq[:, :, 2]):.2e} dqb={rel(dqb, gq[:, :, 0].sum((0, 1)).flatten()):.2e} "
              f"dvb={rel(dvb, gq[:, :, 2].sum((0, 1)).flatten()):.2e} dtable={rel(dtable, ref_tab):.2e} finite={bool(torch.isfinite(dqkv.float()).all())}")


def timeit(fn, n=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def bench():
    B, H, N = 128, 12, 197
    R = 4                                           # rotating buffers: 4 x (116 + 39) MB > 126 MB L2
    qkvs = [torch.randn(B, N, 3, H, 64, device=dev).bfloat16() for _ in range(R)]
    outs = [torch.empty(B, N, H * 64, dtype=torch.bfloat16, device=dev) for _ in range(R)]
    bias = torch.randn(H, N, N, device=dev) * 0.5
    bias_f, bias_t = ops.pad_attn_bias(bias)
    lse = torch.empty(B, H, N, device=dev)
    bits = torch.zeros(B, H, N, 32, dtype=torch.uint8, device=dev)
    scale = 64 ** -0.5
    for p in (0.0, 0.05):
        us = timeit(lambda i: ops.attn_fwd(qkvs[i % R], bias_f, B, H, N, scale, p, 1, 2, None, outs[i % R], lse, bits if p > 0 else None))
        fl = 4.0 * B * H * N * N * 64
        print(f"attn_fwd p={p}: {us:8.1f} us   {fl / us * 1e-6:7.1f} TFLOP/s algorithmic")
    us = timeit(lambda i: ops.attn_fwd(qkvs[i % R], None, B, H, N, scale, 0.0, 1, 2, None, outs[i % R], lse, None))
    print(f"attn_fwd p=0.0 no bias (no bias ring traffic): {us:8.1f} us")
    table = torch.randn(732, H, device=dev) * 0.5
    from uncertainty_vit_b200 import modeling as _M
    index = _M.relative_position_index(14, 14).to(torch.int32).to(dev).contiguous()      # the reference's index (structured: neighbouring rows -> neighbouring bins)
    bias_i, _ = ops.rel_pos_bias(table, index, N, H, want_bwd=False, want_index_tiles=True)
    bias_d, _ = ops.rel_pos_bias(table, index, N, H, want_bwd=False)
    for name, bb in (("dense fp32 bias ring", bias_d), ("indexed bias (resident index tile + table row)", bias_i)):
        for p in (0.0, 0.05):
            us = timeit(lambda i: ops.attn_fwd(qkvs[i % R], bb, B, H, N, scale, p, 1, 2, None, outs[i % R], lse, bits if p > 0 else None, keep_ready=p > 0))
            print(f"attn_fwd p={p} {name} (masks pre-drawn): {us:8.1f} us")
    if "--fwd-only" in sys.argv:
        return
    if "--bwd" in sys.argv:
        douts = [torch.randn(B, N, H * 64, device=dev).bfloat16() for _ in range(R)]
        dqkv = torch.empty(B, N, 3, H, 64, dtype=torch.bfloat16, device=dev)
        idx = torch.randint(0, 732, (N, N), dtype=torch.int32, device=dev)
        dtable = torch.zeros(732, H, device=dev)
        ds_work = ops.attn_bwd_workspace(B, H, N, dev)
        for p in (0.0, 0.05):
            ops.attn_fwd(qkvs[0], bias_f, B, H, N, scale, p, 1, 2, None, outs[0], lse, bits if p > 0 else None)
            us = timeit(lambda i: ops.attn_bwd(qkvs[0], outs[0], douts[i % R], lse, bias_t, bits if p > 0 else None, idx, dtable, B, H, N, scale, p, dqkv,
                                               ds_work=ds_work))
            print(f"attn_bwd (+relbias_grad) p={p}: {us:8.1f} us")
        us = timeit(lambda i: ops.attn_bwd(qkvs[0], outs[0], douts[i % R], lse, None, None, None, None, B, H, N, scale, 0.0, dqkv, ds_work=ds_work))
        print(f"attn_bwd p=0.0 no bias, no table gradient: {us:8.1f} us")
        us = timeit(lambda i: ops.attn_bwd(qkvs[0], outs[0], douts[i % R], lse, bias_t, None, None, None, B, H, N, scale, 0.0, dqkv, ds_work=ds_work))
        print(f"attn_bwd p=0.0 bias, no table gradient: {us:8.1f} us")


def ncu_run():
    """three launches per variant at the step's shape (for `ncu -k regex:attn`)"""
    B, H, N = 128, 12, 197
    qkv = torch.randn(B, N, 3, H, 64, device=dev).bfloat16()
    out = torch.empty(B, N, H * 64, dtype=torch.bfloat16, device=dev)
    bias_f, bias_t = ops.pad_attn_bias(torch.randn(H, N, N, device=dev) * 0.5)
    lse = torch.empty(B, H, N, device=dev)
    bits = torch.zeros(B, H, N, 32, dtype=torch.uint8, device=dev)
    for p in (0.0, 0.05):
        for _ in range(3):
            ops.attn_fwd(qkv, bias_f, B, H, N, 0.125, p, 1, 2, None, out, lse, bits if p > 0 else None)
    if "--bwd" in sys.argv:
        dout = torch.randn(B, N, H * 64, device=dev).bfloat16()
        dqkv = torch.empty(B, N, 3, H, 64, dtype=torch.bfloat16, device=dev)
        idx = torch.randint(0, 732, (N, N), dtype=torch.int32, device=dev)
        dtable = torch.zeros(732, H, device=dev)
        ds_work = ops.attn_bwd_workspace(B, H, N, dev)
        for _ in range(2):
            ops.attn_bwd(qkv, out, dout, lse, bias_t, bits, idx, dtable, B, H, N, 0.125, 0.05, dqkv, ds_work=ds_work)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if "--ncu" in sys.argv:
        ncu_run()
    elif "--parity-bwd" in sys.argv:
        parity_bwd()
    else:
        parity()
        parity_bwd()
        bench()
