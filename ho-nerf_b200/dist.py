"""Multi-GPU plumbing: one process per GPU, rays (or views) sharded across ranks, and ONE flat
all-reduce of the gradients per step (NCCL over NVLink on the GPU box, gloo in the CPU tests).
The path has no other exchange step: sampling, MLPs and compositing are independent per ray
(SURVEY.md section 8e).

The recipe that reproduces the single-process gradient exactly (tests/test_dist_cpu.py):

    lo, hi = shard_rays(n_total, rank, world)                  # balanced, never empty when n_total >= world
    div, share = loss_normalisers(true_mask_local, n_total)    # ONE tiny all-reduce: the batch-wide mask_sum + 1e-5
    out  = renderer.render(rays_o[lo:hi], rays_d[lo:hi], ...)
    loss = ops.render_loss(out['color_fine'], out['weight_sum'], rgb, mask, out['gradient_error'],
                           div, 1.0, mask_weight * share, igr_weight * share)[0]
    loss.backward()
    allreduce_gradients(params, average=False)                 # SUM: the shard losses add up to the batch loss

With ``optim.FlatAdam`` the last two lines (all-reduce, then the optimiser step) can be ONE kernel over NVLink peer memory:
``opt.enable_peer_exchange()`` once, then ``opt.step(peer_exchange=True)`` instead of ``allreduce_flat(opt.flat_grad);
opt.step()`` (csrc/peer.cu: two-shot rank-ordered sum + Adam, bit-identical parameters on every rank).

The reference's colour loss is normalised by the BATCH's mask_sum (exp_runner.py:207,221) and its BCE / eikonal terms
are means over the BATCH: a rank-local normaliser followed by an average over ranks is only right for equal shards
with equal mask counts.  Passing the global divisor and weighting the mean terms by the shard's share of the rays makes
the shard losses ADD UP to the batch loss whatever the split (the same rule NeuSRenderer.render_sharded uses for its
per-stream shards inside one GPU).
"""
import torch
import torch.distributed as dist


def shard_rays(n_total, rank, world):
    """Contiguous balanced shards: the first n_total % world ranks take one extra ray; no shard is empty unless
    n_total < world (then the trailing ranks get (n_total, n_total) and must skip the step CONSISTENTLY: they still
    call allreduce_gradients, which sends zeros for every parameter)."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def loss_normalisers(true_mask_local, n_total, eps=1e-5):
    """(div, share): div = the batch-wide mask_sum + eps as a 0-d tensor on the mask's device (ops.render_loss takes it
    as a device scalar: no host sync), share = this rank's fraction of the batch's rays (weight of its mean terms)."""
    s = true_mask_local.sum().reshape(1).float()
    if _active():
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    n_local = true_mask_local.shape[0] if true_mask_local.dim() > 0 else 1
    return (s + eps).reshape(()), float(n_local) / float(max(int(n_total), 1))


def flatten_gradients(params):
    """The gradient of EVERY parameter of the list packed into one contiguous buffer, zeros where a parameter has no
    gradient on this rank: every rank sends the same number of elements whatever it rendered."""
    ref = next((p for p in params), None)
    if ref is None:
        return torch.zeros(0)
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])


def allreduce_flat(flat):
    """The collective itself; kept separate so a caller can leave it outside a captured CUDA graph."""
    if _active():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def unflatten_gradients(params, flat, world_size, average=True):
    """Scatter the reduced buffer back.  A parameter without a local gradient receives one (it may have a gradient on
    another rank) unless its reduced gradient is identically absent everywhere -- which the ranks cannot know without
    another exchange, so every parameter of the list ends up with a .grad."""
    off = 0
    for p in params:
        n = p.numel()
        src = flat[off:off + n].view_as(p)
        val = src / world_size if average else src
        if p.grad is None:
            p.grad = val.clone()
        else:
            p.grad.copy_(val)
        off += n
    return off


def allreduce_gradients(params, world_size=None, average=True):
    """One flat all-reduce over the gradients of `params` (a FIXED list, identical on every rank): 824 037 floats for the
    object nets, 1 865 661 for the hand nets -- latency-bound, hence a single collective per step.  Returns the number of
    elements sent (0 when there is nothing to reduce with)."""
    if not _active():
        return 0
    params = list(params)
    world_size = world_size or dist.get_world_size()
    if world_size == 1 or not params:
        return 0
    flat = flatten_gradients(params)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    unflatten_gradients(params, flat, world_size, average)
    return flat.numel()


def shard_views(n_views, rank, world):
    """Fitting (fitting_single.py:200-291 loops over the views of a frame): contiguous view shards per rank.  The
    reference steps its optimiser once per view; with views sharded the ranks render their views, SUM the pose-parameter
    gradients (allreduce_gradients(pose_params, average=False): 45 F + 9 floats) and take one identical step --
    parity is per render call, not per optimisation trajectory."""
    return shard_rays(n_views, rank, world)


def gather_slabs(local, sizes, dim=0):
    """All-gather of per-rank slabs of different lengths along `dim` (the u lattice of extract_geometry sharded in x
    slabs, full-image renders sharded by view or by ray chunk): every rank returns the concatenation."""
    if not _active():
        return local
    world = dist.get_world_size()
    pad = max(sizes)
    shape = list(local.shape)
    shape[dim] = pad
    buf = local.new_zeros(shape)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf.contiguous())
    return torch.cat([o.narrow(dim, 0, sizes[r]) for r, o in enumerate(outs)], dim=dim)
