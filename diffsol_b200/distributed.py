"""Multi-GPU plumbing: the batch shards trivially (instances are independent), so the only collective
on the path is ONE all-gather of the output trajectories (and of the per-instance counters).

Instance i lives on rank i mod G (interleaved, so a parameter-sorted sweep stays balanced); rank r
holds local instance k = global instance r + G k.  Works on any torch.distributed backend: NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def local_count(nbatch, rank, world_size):
    """Number of instances rank owns under the i mod G partition."""
    return (nbatch - rank + world_size - 1) // world_size


def all_gather_batch_major(local, nbatch):
    """local: [..., B_r] tensor (batch-major: instance index fastest) of this rank's shard.
    Returns [..., nbatch] in GLOBAL instance order on every rank, with one all_gather."""
    rank, G = world()
    if G == 1:
        return local
    bmax = local_count(nbatch, 0, G)
    lead = local.shape[:-1]
    padded = local
    if local.shape[-1] != bmax:                      # ragged tail: pad to the common shard size
        padded = local.new_zeros(lead + (bmax,))
        padded[..., : local.shape[-1]] = local
    padded = padded.contiguous()
    # concatenation along dim 0 is the one output layout every backend (NCCL, gloo) accepts
    flat = padded.reshape(-1)
    out = flat.new_empty(G * flat.numel())
    dist.all_gather_into_tensor(out, flat)
    out = out.view((G,) + tuple(padded.shape))
    # out[r, ..., k] is global instance r + G k  ->  [..., k, r] -> flatten
    out = out.movedim(0, -1).reshape(lead + (bmax * G,))
    return out[..., :nbatch]
