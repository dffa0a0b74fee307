"""torchrun --nproc-per-node N tools/xe_gather_test.py -- channel-sharded X-engine on N GPUs: the NCCL all_gather of
the visibility slabs vs the fused peer-memory gather (epilogue stores into every rank's matrix over NVLink).
Checks that both give the same full matrix on every rank (and that rank 0's matches the exact oracle on a slab),
then times both with CUDA events, max over ranks."""
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
A, F, T = 32, 1024, 1024
nbl = A * (A + 1) // 2
f0, fc = multigpu.shard_channels(F, rank, world)
g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
bufs = [torch.randint(-127, 128, (T * A * fc * 2,), dtype=torch.int8, device="cuda", generator=g) for _ in range(4)]
slab = torch.empty(fc * nbl * 2, dtype=torch.float32, device="cuda")

xe = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])
xe.launch_device(bufs[0].data_ptr(), slab.data_ptr(), False, sp)
want = multigpu.gather_visibilities(slab, F, nbl * 2).cpu().numpy().view(np.complex64)

xg = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])
xg.set_shard(F, f0)
pg = multigpu.PeerGather(xg, local, F, nbl)
xg.launch_device_gather(bufs[0].data_ptr(), sp)
torch.cuda.synchronize(); dist.barrier()
got = pg.result()
ok = bool(np.array_equal(got, want))
if rank == 0:
    from oracle import oracle as orc
    b0 = bufs[0].cpu().numpy().reshape(T, A, fc, 2)[:, :, :4, :].copy()
    ex = orc.xengine_exact(b0, A, 4, T, 1).astype(np.float64) / (127.0 * 127.0)
    sl = got[:4 * nbl]
    ok = ok and float(np.max(np.abs(sl.real - ex[:, 0])) + np.max(np.abs(sl.imag - ex[:, 1]))) < 1e-2
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)

def timed(step, n=20):
    for i in range(3): step(i)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): step(i)
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def nccl_step(i):
    xe.launch_device(bufs[i % 4].data_ptr(), slab.data_ptr(), False, sp)
    multigpu.gather_visibilities(slab, F, nbl * 2)

t_nccl = timed(nccl_step)
t_fused = timed(lambda i: xg.launch_device_gather(bufs[i % 4].data_ptr(), sp))
t_local = timed(lambda i: xe.launch_device(bufs[i % 4].data_ptr(), slab.data_ptr(), False, sp))
if rank == 0:
    print("world=%d  same matrix on every rank: %s" % (world, bool(flag.item())))
    print("kernel only (no gather)      %7.1f us / integration" % t_local)
    print("kernel + NCCL all_gather     %7.1f us" % t_nccl)
    print("fused peer-memory gather     %7.1f us  (%.0f Msamples/s over %d GPUs)" % (t_fused, A * F * T / t_fused, world))
pg.close()
dist.destroy_process_group()
