"""Multi-GPU plumbing: one process per GPU, rays (or views) sharded across ranks, and ONE flat
all-reduce of the gradients per step (NCCL over NVLink on the GPU box, gloo in the CPU tests).
The path has no other exchange step: sampling, MLPs and compositing are independent per ray
(SURVEY.md section 8e)."""
import torch
import torch.distributed as dist


def shard_rays(n_total, rank, world):
    """Contiguous, equal-size ray shards (the tail shard takes the remainder)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def allreduce_gradients(params, world_size=None, average=True):
    """Flatten every existing .grad into one buffer, all-reduce once, scatter back.
    824 037 floats for the object nets, 1 865 661 for the hand nets: latency-bound, so a single
    collective per step."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world_size = world_size or dist.get_world_size()
    if world_size == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(world_size)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()
