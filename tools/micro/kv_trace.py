import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
# needs a trace build:  B200VIT_EXTRA_NVCC_FLAGS=-DB200VIT_KV_TRACE python uncertainty-vit_b200/build.py --force
import uncertainty_vit_b200 as pkg
ops = pkg.ops; dev = torch.device("cuda:0")
B, H, N = 128, 12, 197
qkv = torch.randn(B, N, 3, H, 64, device=dev).bfloat16(); out = torch.empty(B, N, H * 64, dtype=torch.bfloat16, device=dev)
bias_f, bias_t = ops.pad_attn_bias(torch.randn(H, N, N, device=dev) * 0.5)
lse = torch.empty(B, H, N, device=dev); bits = torch.zeros(B, H, N, 32, dtype=torch.uint8, device=dev)
ops.attn_fwd(qkv, bias_f, B, H, N, 0.125, 0.05, 1, 2, None, out, lse, bits)
dout = torch.randn(B, N, H * 64, device=dev).bfloat16(); dqkv = torch.empty(B, N, 3, H, 64, dtype=torch.bfloat16, device=dev)
idx = torch.randint(0, 732, (N, N), dtype=torch.int32, device=dev); dtable = torch.zeros(732, H, device=dev)
ws = ops.attn_bwd_workspace(B, H, N, dev)
lib = pkg._lib.lib()
buf = (ctypes.c_longlong * (4 * 1536))()
for rep in range(2):
    ops.attn_bwd(qkv, out, dout, lse, bias_t, bits, idx, dtable, B, H, N, 0.125, 0.05, dqkv, ds_work=ws)
    n = lib.b200vit_debug_kv_trace(buf, 1536)
ev = sorted([(buf[4 * i + 3], buf[4 * i], buf[4 * i + 1], buf[4 * i + 2]) for i in range(n) if buf[4 * i + 3] > 0 and buf[4 * i] > 0])
t0 = ev[0][0]
names = {1: "g0 s_full seen", 2: "g0 box done", 3: "g0 acc_full seen", 4: "g0 epilogue done", 5: "g0 item start", 6: "g0   epi tmem loaded", 7: "g0   epi stored", 8: "g0   epi colsum done", 106: "g1   epi tmem loaded", 107: "g1   epi stored", 108: "g1   epi colsum done", 101: "g1 s_full seen", 102: "g1 box done",
         103: "g1 acc_full seen", 104: "g1 epilogue done", 105: "g1 item start", 13: "MMA   fence done", 14: "MMA   acc mmas issued", 20: "MMA   ld_full seen", 21: "MMA   score mmas issued", 10: "MMA p_full seen", 11: "MMA acc issued", 12: "MMA scores issued"}
for t, c, it, bi in ev:
    if it == 1 and c not in (10, 11, 12, 13, 14, 20, 21, 1, 2, 101, 102):
        print(f"{t - t0:8d} cyc  it={it} bi={bi}  {names.get(c, c)}")
