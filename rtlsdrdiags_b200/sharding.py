"""Channel sharding across the GPUs of one box.

Channels are independent, so the data path has no collective: rank g owns the
contiguous channel range [g*N/G, (g+1)*N/G) and runs its own engine. The only
exchange is the optional gather of PCM rows (1/32 of the input volume) to one
rank; torch.distributed carries it (NCCL for CUDA tensors, gloo for host tensors).
"""
import torch


def shard_range(n_total, rank, world):
    """Contiguous channel range of `rank` out of `world`: sizes differ by at most one."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def shard_sizes(n_total, world):
    return [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]


def gather_pcm(local_rows, n_total, dst=0, group=None):
    """Gather [n_local][samples] PCM rows of every rank into [n_total][samples] on `dst`
    (None elsewhere). Rows come back in global channel order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    assert local_rows.shape[0] == sizes[rank], "local rows do not match this rank's shard"
    width = local_rows.shape[1]
    biggest = max(sizes)
    # dist.gather wants equal shapes: pad every shard to the largest. PCM travels as
    # bytes (gloo has no int16).
    padded = local_rows.new_zeros((biggest, width))
    padded[: sizes[rank]] = local_rows
    wire = padded.contiguous().view(torch.uint8)
    bufs = [torch.empty_like(wire) for _ in range(world)] if rank == dst else None
    dist.gather(wire, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r].view(local_rows.dtype)[: sizes[r]] for r in range(world)], dim=0)
