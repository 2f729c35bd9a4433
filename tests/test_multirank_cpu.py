"""CPU, world_size 2, gloo: the host-side logic of the N>1 paths (no CUDA): pass sharding + all-gather order of the MC evaluation,
gradient-arena all-reduce semantics of the data-parallel step (sum / world == DDP's mean), per-rank row indexing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import uncertainty_vit_b200  # noqa: F401
    from uncertainty_vit_b200 import mc
    S, N, K = 7, 5, 4
    full = torch.arange(S * N * K, dtype=torch.float32).reshape(S, N, K)
    s0, s1 = mc.shard_passes(S, world)[rank]
    got = mc.gather_passes(full[s0:s1].clone(), S, rank, world)
    ok_gather = torch.equal(got, full)
    # sample x batch sharding of the MC evaluation: images split across ranks, gathered back in image order
    rows = torch.arange(N * K, dtype=torch.float32).reshape(N, K)
    i0, i1 = mc.shard_images(N, world)[rank]
    ok_gather = ok_gather and torch.equal(mc.gather_rows(rows[i0:i1].clone(), N, rank, world), rows)
    # data-parallel gradient arena: all_reduce(sum) then grad_div = world (what D2VEngine.step does) == mean of per-rank grads
    g = torch.full((1024,), float(rank + 1))
    dist.all_reduce(g)
    ok_grad = torch.allclose(g / world, torch.full((1024,), (1 + world) / 2.0))
    # the engines replace DDP: rank 0's freshly initialised parameter arena is broadcast at construction (run_cyclical.py:316-322 seeds
    # every rank differently, :515-519 DDP makes them equal), and equal arenas + all-reduced gradients stay equal after a step
    from uncertainty_vit_b200.engine import sync_initial_parameters
    torch.manual_seed(100 + rank)
    arena = torch.randn(4096)
    mine_before = arena.clone()
    sync_initial_parameters(arena, world, None)
    grad = torch.randn(4096)                       # per-rank gradient (different seeds)
    dist.all_reduce(grad)
    arena -= 0.1 * grad / world                    # any deterministic optimiser on identical inputs
    gathered = [torch.empty_like(arena) for _ in range(world)]
    dist.all_gather(gathered, arena)
    ok_sync = all(torch.equal(gathered[0], t) for t in gathered) and (rank == 0 or not torch.equal(mine_before, arena))
    q.put((rank, ok_gather, ok_grad and ok_sync))
    dist.destroy_process_group()


def test_world2_gloo_sharding_and_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] for r in res), res


def test_shard_passes_partitions():
    import uncertainty_vit_b200  # noqa: F401
    from uncertainty_vit_b200 import mc
    for S in (2, 7, 30, 31):
        for world in (1, 2, 4, 8):
            sh = mc.shard_passes(S, world)
            assert sh[0][0] == 0 and sh[-1][1] == S and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [e - s for s, e in sh]
            assert max(sizes) - min(sizes) <= 1
    assert mc.shard_passes(30, 8) == [(0, 4), (4, 8), (8, 12), (12, 16), (16, 20), (20, 24), (24, 27), (27, 30)]


def test_rows_from_host_mask_matches_boolean_gather_order():
    import uncertainty_vit_b200  # noqa: F401
    from uncertainty_vit_b200.engine import D2VEngine, cosine_scheduler, get_num_layer_for_vit
    rng = np.random.RandomState(0)
    mask = (rng.rand(3, 196) < 0.6).astype(np.uint8)
    rows = D2VEngine.rows_from_host_mask(mask, 197)
    x = torch.arange(3 * 197).reshape(3, 197)
    expect = x[:, 1:].reshape(-1)[torch.from_numpy(mask.reshape(-1)).bool()]        # modeling_cyclical.py:222-224
    assert np.array_equal(rows, expect.numpy().astype(np.int32))
    assert get_num_layer_for_vit("blocks.11.mlp.fc1.weight", 14) == 12 and get_num_layer_for_vit("lm_head.weight", 14) == 13
    sched = cosine_scheduler(2e-3, 1e-5, 4, 5, warmup_epochs=1)
    assert len(sched) == 20 and sched[0] == 0.0 and abs(sched[5] - 2e-3) < 1e-12 and sched[-1] > 1e-5
