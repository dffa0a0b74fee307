"""One-process-per-GPU sharding of the two blocks that partition.

clFFT: vectors/streams are independent (lib/clFFT_impl.cc:537-541) -> contiguous vector
ranges per rank, no collective.  clXEngine: channels are independent
(lib/clXEngine_impl.cc:742-743) -> rank g owns channels [g*F/G, (g+1)*F/G); its handle
reads only that slab of the [t][station][chan][pol] buffer (clb200_xengine_set_shard ->
cudaMemcpy2D gather) and the per-rank visibility slabs, contiguous in the
[chan][baseline][pol^2] result, are concatenated with one all_gather.
torch.distributed is plumbing only; nothing here computes.
"""


def shard_range(n, rank, world):
    """contiguous near-equal split of n items: (first, count) of `rank`"""
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard_vectors(nvec, rank, world):
    return shard_range(nvec, rank, world)


def shard_channels(num_channels, rank, world):
    return shard_range(num_channels, rank, world)


def gather_visibilities(local_slab, num_channels, items_per_channel, group=None):
    """all_gather the per-rank [chan_count][baseline][pol^2] slabs into the full matrix.

    local_slab: torch tensor (any device the process group supports) holding this rank's
    channels; returns a tensor of num_channels*items_per_channel*(trailing dims) elements
    in channel order on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_channels(num_channels, r, world)[1] for r in range(world)]
    per = items_per_channel * (local_slab.numel() // max(1, counts[rank] * items_per_channel))
    flat = local_slab.reshape(-1)
    assert flat.numel() == counts[rank] * per
    if len(set(counts)) == 1:
        out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
        dist.all_gather_into_tensor(out, flat, group=group) if flat.is_cuda else \
            dist.all_gather(list(out.chunk(world)), flat, group=group)
        return out
    # ragged: pad to the largest slab
    mx = max(counts) * per
    pad = torch.zeros(mx, dtype=flat.dtype, device=flat.device)
    pad[:flat.numel()] = flat
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c * per] for b, c in zip(bufs, counts)])
