"""N > 1 host logic on CPU: world_size-2 gloo process group (the data path itself needs no collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from enf_pde_b200.dist import field_shard, query_shard, choose_partition, pack, unpack, allreduce_weight_grads, allreduce_grads


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    fields = torch.randn(7, 5, generator=g)                    # 7 fields: uneven split over 2 ranks
    mine = fields[field_shard(7, rank, world)]
    # stand-in for per-rank weight gradients: each leaf = sum over my fields (linear in the fields)
    shapes = [(3, 4), (4,), (2, 2, 2)]
    grads = [mine.sum() * torch.ones(s) * (i + 1) for i, s in enumerate(shapes)]
    red = allreduce_weight_grads(grads)
    want = [fields.sum() * torch.ones(s) * (i + 1) for i, s in enumerate(shapes)]
    ok = all(torch.allclose(a, b, atol=1e-5) for a, b in zip(red, want))
    mean = allreduce_weight_grads(grads, average=True)
    ok = ok and all(torch.allclose(a, b / world, atol=1e-5) for a, b in zip(mean, want))
    out[rank] = ok
    dist.destroy_process_group()


def _query_worker(rank, world, port, out):
    """B = 1 field < 2 ranks: shard the queries; the oracle stands in for the CUDA path (same math, CPU).  Summed over
    ranks, the latent and weight gradients of the per-shard losses must equal those of the full problem (linearity)."""
    from oracle import enf_ref as R
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    cfg = R.EnfConfig(num_in=2, num_hidden=16, num_heads=2, num_out=1, latent_dim=4, invariant_type="rel_pos_periodic")
    params = R.nef_init(cfg, seed=0)
    p, a, sigma = R.init_latents(cfg, 1, 4)
    x = R.make_coords(cfg, (5, 5))[None]
    d_out = torch.randn(1, 25, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    assert choose_partition(1, world) == "queries" and choose_partition(8, world) == "fields"
    sl = query_shard(25, rank, world)
    _, g_loc, dp_loc, da_loc, ds_loc = R.fwd_bwd(cfg, params, x[:, sl], p, a, sigma, d_out[:, sl])
    leaves = list(R.tree_flatten(g_loc["params"]).values())
    wg, (dp, da, ds) = allreduce_grads(leaves, [dp_loc, da_loc, ds_loc])
    _, g_all, dp_all, da_all, ds_all = R.fwd_bwd(cfg, params, x, p, a, sigma, d_out)
    want = list(R.tree_flatten(g_all["params"]).values())
    ok = all(torch.allclose(u, v, atol=1e-10) for u, v in zip(wg, want))
    ok = ok and torch.allclose(dp, dp_all, atol=1e-10) and torch.allclose(da, da_all, atol=1e-10) and torch.allclose(ds, ds_all, atol=1e-10)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_query_sharded_backward_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_query_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_field_shards_partition_the_batch():
    for n in (1, 2, 7, 32, 33):
        for world in (1, 2, 4, 8):
            idx = []
            for r in range(world):
                s = field_shard(n, r, world)
                idx += list(range(n))[s]
            assert idx == list(range(n))
            sizes = [len(range(n)[field_shard(n, r, world)]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_round_trip():
    ts = [torch.randn(3, 4), torch.randn(5), torch.randn(2, 2, 2)]
    back = unpack(pack(ts), ts)
    assert all(torch.equal(a, b) for a, b in zip(ts, back))


def test_weight_grad_allreduce_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
