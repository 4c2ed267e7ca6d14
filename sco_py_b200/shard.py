"""Multi-GPU sharding of a batch: one process per GPU, problems split contiguously, results
gathered once at the end.  Problems never interact (solver.py:30-59 takes
one Prob), so there is no data-path collective; the only communication is this final gather.

Works on any torch.distributed backend: NCCL on the GPU box (device tensors travel over
NVLink/NVSwitch), gloo on CPU for the host-logic tests.
"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous, balanced split: the first B % world ranks get one extra problem."""
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(B, world):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_results(local, B, group=None):
    """local: dict name -> tensor [b_local, ...] of this rank's shard (shard_range order).
    Returns the same dict with tensors of the full batch [B, ...] on every rank."""
    if not dist.is_available() or not dist.is_initialized():
        return dict(local)
    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    pad = max(sizes)
    out = {}
    for name, t in local.items():
        t = t.contiguous()
        buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[:t.shape[0]] = t
        full = torch.empty((world * pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, buf, group=group)
        full = full.view((world, pad) + tuple(t.shape[1:]))
        out[name] = torch.cat([full[r, :sizes[r]] for r in range(world)], dim=0)
    return out
