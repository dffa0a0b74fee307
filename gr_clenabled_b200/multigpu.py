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


FLAG_STRIDE = 32          # CLB200_XENGINE_FLAG_STRIDE (uint32 words between two ranks' flags: one 128 B line each)


class PeerGather:
    """Fused X-engine all-gather over peer memory (one process per GPU of one box).

    Every rank allocates the FULL visibility matrix c32[num_channels][items_per_channel] and a small flag array
    with clb200_mem_alloc, exports their interprocess handles, opens the peers' (peer access over NVLink is
    enabled on open) and registers all of them with its channel-sharded clXEngine handle
    (clb200_xengine_set_gather / set_gather_sync): the correlation kernel's epilogue then stores this rank's
    channel slab into every rank's matrix (16 B peer stores) and, once its last CTA has stored, releases its flag on
    every rank; xe.gather_wait(stream) makes the stream wait -- on the device -- until every rank's slab of the
    current epoch has landed.  No collective and no host barrier follow the kernel; torch.distributed only carries
    the 64-byte handles, once."""

    def __init__(self, xe, device, num_channels, items_per_channel, group=None):
        import ctypes as C

        import torch.distributed as dist

        from . import capi
        self._lib, self.device = capi.load(), device
        self.nbytes = num_channels * items_per_channel * 8
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.own, self.own_flags = C.c_void_p(), C.c_void_p()
        capi.check(self._lib.clb200_mem_alloc(device, self.nbytes, C.byref(self.own)))
        capi.check(self._lib.clb200_mem_alloc(device, world * FLAG_STRIDE * 4, C.byref(self.own_flags)))
        h, hf = C.create_string_buffer(64), C.create_string_buffer(64)
        capi.check(self._lib.clb200_ipc_export(device, self.own, h))
        capi.check(self._lib.clb200_ipc_export(device, self.own_flags, hf))
        handles = [None] * world
        dist.all_gather_object(handles, (bytes(h.raw), bytes(hf.raw)), group=group)
        self.ptrs, self.flag_ptrs, self._opened = [], [], []
        for r in range(world):
            if r == rank:
                self.ptrs.append(self.own.value)
                self.flag_ptrs.append(self.own_flags.value)
                continue
            p, pf = C.c_void_p(), C.c_void_p()
            capi.check(self._lib.clb200_ipc_open(device, C.create_string_buffer(handles[r][0], 64), C.byref(p)))
            capi.check(self._lib.clb200_ipc_open(device, C.create_string_buffer(handles[r][1], 64), C.byref(pf)))
            self.ptrs.append(p.value)
            self.flag_ptrs.append(pf.value)
            self._opened += [p, pf]
        xe.set_gather(self.ptrs)
        xe.set_gather_sync(rank, self.flag_ptrs)

    def result(self):
        """this rank's full matrix as complex64 (after xe.gather_wait + a stream sync, or a barrier)"""
        import ctypes as C

        import numpy as np

        from . import capi
        out = np.zeros(self.nbytes // 8, np.complex64)
        capi.check(self._lib.clb200_mem_copy_to_host(self.device, self.own, out.ctypes.data_as(C.c_void_p), self.nbytes))
        return out

    def close(self):
        for p in self._opened:
            self._lib.clb200_ipc_close(self.device, p)
        self._opened = []
        if self.own:
            self._lib.clb200_mem_free(self.device, self.own)
            self._lib.clb200_mem_free(self.device, self.own_flags)
            self.own = self.own_flags = None


class MulticastGather:
    """The same fused gather through an NVSwitch MULTICAST object: the matrix and the flag array live in one
    torch symmetric-memory allocation (torch.distributed._symmetric_memory: cuMemCreate + cuMulticastBindMem on every
    rank, the plumbing this module leaves to torch), and the kernel's epilogue issues ONE `multimem.st` per 16 B --
    the switch replicates it into every rank's copy -- instead of one peer store per rank: 1/N of the NVLink egress.
    Raises RuntimeError when the box has no multicast support (then use PeerGather)."""

    def __init__(self, xe, device, num_channels, items_per_channel, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = num_channels * items_per_channel * 8
        mat_words = (self.nbytes + 255) // 256 * 64                       # matrix, padded to 256 B
        self.t = symm_mem.empty(mat_words + world * FLAG_STRIDE, dtype=torch.float32, device=torch.device("cuda", device))
        self.t.zero_()
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:                                                  # noqa: BLE001 -- not needed on newer torch
            pass
        self.hdl = symm_mem.rendezvous(self.t, group.group_name)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("no multicast address for the symmetric allocation (NVLS unavailable)")
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        xe.set_gather(ptrs)
        xe.set_gather_sync(rank, [p + mat_words * 4 for p in ptrs], mc, mc + mat_words * 4)
        self.mat_words = mat_words
        torch.cuda.synchronize()
        dist.barrier(group)

    def result(self):
        import numpy as np
        return self.t[:self.nbytes // 4].cpu().numpy().view(np.complex64)

    def close(self):
        self.hdl = None
        self.t = None
