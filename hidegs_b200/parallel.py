"""View-sharded data parallelism: Gaussians replicated on every rank, camera views
split `views[rank::world]`, one gradient all-reduce per step.

The backward operator returns every parameter gradient as a view into ONE flat fp32 arena whose
first 59 floats per Gaussian are the trainable parameters (xyz 3 | sh 48 | opacity 1 | scale 3 |
rotation 4).  `flat_view()` recovers that arena from the gradient tensors so the collective runs in
place on a single buffer (NCCL over NVLink on GPUs, gloo in the CPU tests) with no pack kernel; when
the tensors do not share storage (e.g. autograd had to clone one) they are packed first.
"""
import torch
import torch.distributed as dist


def shard_views(views, rank=None, world=None):
    """The views of one step that belong to this rank: views[rank::world]."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return list(views)[rank::world]


def flat_view(tensors):
    """If `tensors` are back-to-back contiguous views of one storage, return the flat 1-D view that
    covers exactly them (no copy); otherwise None."""
    tensors = [t for t in tensors if t is not None and t.numel()]
    if not tensors:
        return None
    first = tensors[0]
    base = first.untyped_storage().data_ptr()
    off = first.storage_offset()
    for t in tensors:
        if (not t.is_contiguous() or t.dtype != first.dtype or t.device != first.device
                or t.untyped_storage().data_ptr() != base or t.storage_offset() != off):
            return None
        off += t.numel()
    n = off - first.storage_offset()
    return torch.as_strided(first, (n,), (1,), first.storage_offset())


def allreduce_gradients(grads, group=None, average=False, async_op=False):
    """Sum (or average) the per-rank gradients of one step over all ranks, in place.

    `grads`: gradient tensors in arena order (means3D, sh, opacity, scales, rotations).  Returns
    (work handle or None, number of bytes sent through the collective)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    flat = flat_view(grads)
    packed = None
    if flat is None:
        packed = torch.cat([g.reshape(-1) for g in grads if g is not None and g.numel()])
        flat = packed
    nbytes = flat.numel() * flat.element_size()
    if world == 1:
        return None, nbytes
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op and packed is None)
    if average:
        flat.div_(world)
    if packed is not None:  # scatter the reduced values back
        off = 0
        for g in grads:
            if g is None or not g.numel():
                continue
            g.copy_(packed[off:off + g.numel()].view_as(g))
            off += g.numel()
    return work, nbytes
