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


def flatten_gradients(params):
    """Every existing .grad packed into one contiguous buffer (the payload of the step's single collective)."""
    grads = [p.grad for p in params if p.grad is not None]
    return torch.cat([g.reshape(-1) for g in grads])


def allreduce_flat(flat):
    """The collective itself; kept separate so a caller can leave it outside a captured CUDA graph."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def unflatten_gradients(params, flat, world_size, average=True):
    off = 0
    for p in params:
        if p.grad is None:
            continue
        n = p.grad.numel()
        src = flat[off:off + n].view_as(p.grad)
        p.grad.copy_(src / world_size if average else src)
        off += n
    return off


def shard_views(n_views, rank, world):
    """Fitting (fitting_single.py:200-291 loops over the views of a frame): contiguous view shards per rank.  The
    reference steps its optimiser once per view; with views sharded the ranks render their views, SUM the pose-parameter
    gradients (allreduce_gradients(pose_params, average=False): 45 F + 9 floats) and take one identical step --
    parity is per render call, not per optimisation trajectory."""
    return shard_rays(n_views, rank, world)
