"""Multi-GPU plumbing: one process per GPU, decision groups block-partitioned over ranks, no
collective on the data path.  torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is used only
for the barrier, the max-over-ranks clock and the optional final exchange of per-channel occupancy."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_groups(ngroups, world_size, rank):
    """Contiguous block partition: returns (first_group, count); counts differ by at most one."""
    base, rem = divmod(ngroups, world_size)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world()[0] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def gather_occupancy(decisions, counts=None):
    """Optional epilogue (north_star: 'only an optional final allgather of per-channel occupancy'):
    every rank ends up with the decision code of every group of the whole capture, in group order.
    decisions: int32 tensor of this rank's shard; counts: per-rank shard sizes (needed when uneven)."""
    w, r = world()
    if w == 1:
        return decisions.clone()
    if counts is None:
        counts = [decisions.numel()] * w
    width = max(counts)
    padded = torch.full((width,), -1, dtype=decisions.dtype, device=decisions.device)
    padded[: decisions.numel()] = decisions
    out = [torch.empty_like(padded) for _ in range(w)]
    dist.all_gather(out, padded)
    return torch.cat([o[:c] for o, c in zip(out, counts)])


def occupancy_histogram(decisions):
    """Per-channel occupancy counts [ALL_BUSY, CH1, CH2, CH3] summed over all ranks."""
    h = torch.bincount(decisions.to(torch.int64).clamp(min=0), minlength=4)[:4]
    if world()[0] > 1:
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
    return h


def fuse_across_ranks(masks, nbands, mode, stream=0):
    """Cooperative sensing across GPUs: every rank sensed the same time slots with its own radio;
    the occupancy masks (uint64 per slot, CUDA tensor) are all-gathered over NCCL - 8 bytes per decision, the only
    exchange there is - and fused on the device (OR / majority / AND, crn_fuse_masks_device).  Returns the fused
    mask per slot, identical on every rank."""
    import crn_b200 as crn
    w, _ = world()
    if not masks.is_cuda:
        raise ValueError("fuse_across_ranks needs CUDA tensors: the fusion kernel has no CPU fallback")
    masks = masks.contiguous()
    if w == 1:
        stack = masks.unsqueeze(0)
    else:
        parts = [torch.empty_like(masks) for _ in range(w)]
        dist.all_gather(parts, masks)
        stack = torch.stack(parts)
    fused = torch.empty_like(masks)
    crn.fuse_masks(stack.contiguous(), w, masks.numel(), nbands, mode, fused, masks.device.index or 0, stream)
    return fused
