"""torchrun --nproc-per-node N tools/xe_gather_test.py -- channel-sharded X-engine on N GPUs (BASELINE config 5):
the NCCL all_gather of the visibility slabs vs the gather fused into the kernel's epilogue, over peer memory
(16 B stores into every rank's matrix) and over an NVSwitch multicast object (one multimem.st per 16 B).  All arms
are timed TO CONSUMABLE: the fused arms include the device-side wait for every rank's completion flag
(clb200_xengine_gather_wait).  Checks that every arm gives the same full matrix on every rank and that rank 0's
matches the exact oracle on a slab; CUDA events, max over ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from gr_clenabled_b200 import blocks, capi, multigpu

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sp = torch.cuda.current_stream().cuda_stream
A, T = 32, 1024
F = int(os.environ.get("XE_CHANNELS", "1024"))
nbl = A * (A + 1) // 2
f0, fc = multigpu.shard_channels(F, rank, world)
g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
bufs = [torch.randint(-127, 128, (T * A * fc * 2,), dtype=torch.int8, device="cuda", generator=g) for _ in range(4)]
slab = torch.empty(fc * nbl * 2, dtype=torch.float32, device="cuda")

xe = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])
xe.launch_device(bufs[0].data_ptr(), slab.data_ptr(), False, sp)
want = multigpu.gather_visibilities(slab, F, nbl * 2).cpu().numpy().view(np.complex64)


def timed(step, n=40):
    for i in range(5): step(i)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): step(i)
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def fused_arm(cls):
    xg = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])
    xg.set_shard(F, f0)
    pg = cls(xg, local, F, nbl)
    xg.launch_device_gather(bufs[0].data_ptr(), sp)
    xg.gather_wait(sp)
    torch.cuda.synchronize()                       # NO barrier: the flags are the completion signal
    got = pg.result()
    ok = bool(np.array_equal(got, want))
    dist.barrier()
    def step(i):
        xg.launch_device_gather(bufs[i % 4].data_ptr(), sp)
        xg.gather_wait(sp)
    t = timed(step)
    t_nowait = timed(lambda i: xg.launch_device_gather(bufs[i % 4].data_ptr(), sp))
    # after the timing loop the matrix holds integration (n-1) % 4 of every rank: check it once more
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    pg.close()
    return bool(flag.item()), t, t_nowait, got


def nccl_step(i):
    xe.launch_device(bufs[i % 4].data_ptr(), slab.data_ptr(), False, sp)
    multigpu.gather_visibilities(slab, F, nbl * 2)

t_local = timed(lambda i: xe.launch_device(bufs[i % 4].data_ptr(), slab.data_ptr(), False, sp))
t_nccl = timed(nccl_step)
ok_p, t_peer, t_peer_nw, got = fused_arm(multigpu.PeerGather)
if rank == 0:
    from oracle import oracle as orc
    b0 = bufs[0].cpu().numpy().reshape(T, A, fc, 2)[:, :, :4, :].copy()
    ex = orc.xengine_exact(b0, A, 4, T, 1).astype(np.float64) / (127.0 * 127.0)
    sl = got[:4 * nbl]
    ok_oracle = float(np.max(np.abs(sl.real - ex[:, 0])) + np.max(np.abs(sl.imag - ex[:, 1]))) < 1e-2
try:
    ok_m, t_mc, t_mc_nw, _ = fused_arm(multigpu.MulticastGather)
    mc_err = None
except Exception as e:                       # noqa: BLE001
    ok_m, t_mc, t_mc_nw, mc_err = False, float("nan"), float("nan"), repr(e)[:300]
if rank == 0:
    print("world=%d channels=%d" % (world, F))
    print("kernel only (no gather)                   %7.1f us / integration" % t_local)
    print("kernel + NCCL all_gather                  %7.1f us" % t_nccl)
    print("fused peer-memory gather, to consumable   %7.1f us (%.1f without the flag wait)  same matrix: %s  oracle slab: %s"
          % (t_peer, t_peer_nw, ok_p, ok_oracle))
    print("fused multicast gather, to consumable     %7.1f us (%.1f without the flag wait)  same matrix: %s  %s"
          % (t_mc, t_mc_nw, ok_m, mc_err or ""))
dist.destroy_process_group()
