"""Batch sharding across the GPUs of one box.  Images are independent and the T-step loop has no cross-image
reduction (channel-LN / GroupNorm / softmax are per image), so the path shards by image with NO data-path
collective: every rank restores its own shard (weights replicated), and the restored shards are gathered with
ONE all_gather per batch (NCCL over NVLink on the GPU box; gloo in the CPU tests).  The reference's
single-process nn.DataParallel (denoising_model.py:41-42) would instead replicate the module and scatter/gather
every denoiser call."""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """Contiguous [begin, end) of the images rank `rank` restores; sizes differ by at most one."""
    base, extra = divmod(n_images, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_restored(local, n_images, group=None):
    """All ranks get the full [n_images, ...] batch in image order.  `local`: this rank's restored shard.
    Ragged shards (n_images not divisible by the world size) are padded to the largest shard for the collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_images, r, world) for r in range(world)]
    largest = max(e - b for b, e in sizes)
    padded = local
    if local.shape[0] < largest:
        pad = torch.zeros((largest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[: e - b] for p, (b, e) in zip(parts, sizes)], dim=0)
