"""Contiguous frame sharding with a one-frame halo (multi-GPU inference).

Frames are independent through the network and the soft-argmax (reference: eval.py:306-345 has no cross-frame op);
only the temporal potential couples t and t+1 (fitdgp.py:1079-1083).  Rank r owns frames
[start_r, stop_r); it needs mu of frame stop_r (the first frame of rank r+1) so that delta_t at its last frame is
exact.  That is nj*2 floats per shard edge; no other data-path exchange exists.
"""
import torch
import torch.distributed as dist


def shard_range(T, rank, world):
    """Contiguous split of T frames; the remainder goes to the low ranks."""
    base, rem = divmod(T, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def exchange_halo(mu_first, group=None):
    """mu_first: (nj, 2) soft-argmax of this rank's FIRST frame.  Returns the next rank's first-frame mu, or None on
    the last rank.  One tiny all_gather (world * nj * 8 bytes); works with NCCL (cuda) and gloo (cpu)."""
    if not dist.is_available() or not dist.is_initialized():
        return None
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return None
    bufs = [torch.empty_like(mu_first) for _ in range(world)]
    dist.all_gather(bufs, mu_first.contiguous(), group=group)
    return bufs[rank + 1] if rank + 1 < world else None


def gather_frames(local, T, group=None):
    """All-gather per-frame outputs (T_r, ...) of contiguous shards back into (T, ...) order on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(T, r, world) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)
