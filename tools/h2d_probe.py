"""[torchrun --nproc-per-node N] tools/h2d_probe.py -- the host-link ceiling the e2e numbers are judged against:
bare pinned-memory H2D and D2H copies (no kernel) per rank, one direction at a time and both at once, all ranks
concurrently; reports per-rank and aggregate GB/s (device-timed, max over ranks) plus the CPU affinity / NUMA node of
every GPU.  CLB200_PROBE_BIND=1 binds each rank's thread and its pinned allocation to the GPU's NUMA node first."""
import ctypes, json, os, subprocess, sys
import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def numa_of_gpu(i):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        p = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
        node = int(open(p).read()) if os.path.exists(p) else -1
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bus.lower()[-12:]).read().strip()
        return node, cpus
    except Exception as e:                    # noqa: BLE001
        return -1, "?"


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            out.append(int(part))
    return out


node, cpus = numa_of_gpu(local)
bound = False
if os.environ.get("CLB200_PROBE_BIND") == "1" and cpus != "?":
    try:
        os.sched_setaffinity(0, set(parse_cpulist(cpus)))      # first-touch then places the pinned pages on that node
        bound = True
    except Exception:                          # noqa: BLE001
        pass

n = int(os.environ.get("CLB200_PROBE_MB", "512")) << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(0)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=6):
    def once():
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        once()
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if dist:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return n / (ms / 1e3) / 1e9        # GB/s per direction per rank (slowest rank)


res = {"ranks": world, "mb": n >> 20, "bound_to_gpu_numa_node": bound,
       "h2d_only_GBps_per_rank": run(True, False), "d2h_only_GBps_per_rank": run(False, True),
       "both_GBps_per_rank_per_direction": run(True, True)}
info = [None] * world
if dist:
    dist.all_gather_object(info, (local, node, cpus))
else:
    info = [(local, node, cpus)]
if rank == 0:
    res["aggregate_both_GBps_per_direction"] = res["both_GBps_per_rank_per_direction"] * world
    res["gpu_numa"] = [{"gpu": i, "numa_node": nd, "local_cpus": c} for i, nd, c in info]
    res["host_cpus"] = os.cpu_count()
    try:
        res["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except Exception:                          # noqa: BLE001
        pass
    print(json.dumps(res))
if dist:
    dist.barrier()
    dist.destroy_process_group()
