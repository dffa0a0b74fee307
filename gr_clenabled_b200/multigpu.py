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


class PeerGather:
    """Fused X-engine all-gather over peer memory (one process per GPU of one box).

    Every rank allocates the FULL visibility matrix c32[num_channels][items_per_channel] with
    clb200_mem_alloc, exports its interprocess handle, opens the peers' (peer access over NVLink is
    enabled on open) and registers all of them with its channel-sharded clXEngine handle
    (clb200_xengine_set_gather): the correlation kernel's epilogue then stores this rank's channel slab
    into every rank's matrix, and no collective follows the kernel.  torch.distributed only carries the
    64-byte handles (once) and the caller's barrier."""

    def __init__(self, xe, device, num_channels, items_per_channel, group=None):
        import ctypes as C

        import torch.distributed as dist

        from . import capi
        self._lib, self.device = capi.load(), device
        self.nbytes = num_channels * items_per_channel * 8
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.own = C.c_void_p()
        capi.check(self._lib.clb200_mem_alloc(device, self.nbytes, C.byref(self.own)))
        h = C.create_string_buffer(64)
        capi.check(self._lib.clb200_ipc_export(device, self.own, h))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        self.ptrs, self._opened = [], []
        for r in range(world):
            if r == rank:
                self.ptrs.append(self.own.value)
                continue
            p = C.c_void_p()
            capi.check(self._lib.clb200_ipc_open(device, C.create_string_buffer(handles[r], 64), C.byref(p)))
            self.ptrs.append(p.value)
            self._opened.append(p)
        xe.set_gather(self.ptrs)

    def result(self):
        """this rank's full matrix as complex64 (call after every rank's stream has drained + a barrier)"""
        import ctypes as C

        import numpy as np

        from . import capi
        out = np.zeros(self.nbytes // 8, np.complex64)
        capi.check(self._lib.clb200_mem_copy_to_host(self.device, self.own, out.ctypes.data_as(C.c_void_p), self.nbytes))
        return out

    def close(self):
        from . import capi
        for p in self._opened:
            self._lib.clb200_ipc_close(self.device, p)
        self._opened = []
        if self.own:
            self._lib.clb200_mem_free(self.device, self.own)
            self.own = None
