"""Multi-GPU plumbing for the ENF path: fields shard across ranks, weight gradients all-reduce.

The path is embarrassingly parallel over fields (every tensor's leading axis is the field index and no op
mixes fields: equivariant_cross_attention.py:74-151), and the meta-learning inner loop only touches
per-field latents (pde_trainer.py:199-222), so the data path needs NO collective.  The only exchange is the
outer-loop gradient of the shared weights: one all-reduce of ~0.5 M floats per outer step (NCCL over
NVLink when run with one process per GPU; gloo in the CPU tests).
"""
from typing import Sequence

import torch
import torch.distributed as dist


def field_shard(num_fields: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slice of the global field batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_fields, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def query_shard(num_queries: int, rank: int, world: int) -> slice:
    """Fallback partition when there are fewer fields than ranks (shallow water runs 4 fields, the ball 1):
    every rank takes all fields but a contiguous slice of the C coordinate queries.  The decoded field of a query depends
    on no other query, so the forward needs no exchange; every latent gradient (dp, da, dsigma) and every weight gradient
    is a SUM over queries, so the backward needs one all-reduce of the latent gradients (B*Z*(P+L+1) floats) next to the
    weight-gradient one -- `allreduce_grads` packs both into a single collective."""
    return field_shard(num_queries, rank, world)


def choose_partition(num_fields: int, world: int) -> str:
    """'fields' when every rank gets at least one field (no data-path collective), else 'queries'."""
    return "fields" if num_fields >= world else "queries"


def pack(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors])


def unpack(flat: torch.Tensor, like: Sequence[torch.Tensor]):
    out, off = [], 0
    for t in like:
        out.append(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return out


def allreduce_grads(weight_grads: Sequence[torch.Tensor], latent_grads: Sequence[torch.Tensor] = (), group=None):
    """Query-sharded backward: sum weight gradients AND latent gradients (dp, da, dsigma) over ranks in ONE collective.
    Returns (weight_grads, latent_grads).  With field sharding pass no latent gradients (they are rank-local)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(weight_grads), list(latent_grads)
    both = list(weight_grads) + list(latent_grads)
    flat = pack(both)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out = unpack(flat, both)
    return out[:len(weight_grads)], out[len(weight_grads):]


def allreduce_weight_grads(grads: Sequence[torch.Tensor], group=None, average: bool = False):
    """Sum (or mean) the 46 weight-gradient leaves over ranks with ONE collective on a packed buffer."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(grads)
    flat = pack(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    return unpack(flat, grads)
