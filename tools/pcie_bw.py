"""Host-link floor of the end-to-end path: host->device bandwidth of pinned frame clips, one process per
GPU, all ranks copying at the same time.

    python tools/pcie_bw.py                                   one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_bw.py [--affinity] [--wc]

--affinity  pin the rank's host thread (and its allocations, first touch) to the CPUs of the GPU's NUMA node
--wc        source clip in write-combined pinned memory (gofpid_host_alloc)
Prints per rank and aggregated GB/s for 2 MB (one 1080p frame), 6 MB (a group of three) and 64 MB copies, and
the frames/s of 1080p gray frames that bandwidth allows per GPU (the floor bench.py's e2e runs against)."""
import argparse, glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--affinity", action="store_true")
ap.add_argument("--wc", action="store_true")
ap.add_argument("--sizes", default="2,6,64", help="copy sizes in MB")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
numa = None
try:
    bus = torch.cuda.get_device_properties(local).pci_bus_id
    dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
    cand = glob.glob("/sys/bus/pci/devices/%04x:%02x:*.0/numa_node" % (dom, bus))
    numa = int(open(cand[0]).read()) if cand else None
    if args.affinity and numa is not None and numa >= 0:
        cpus = open("/sys/devices/system/node/node%d/cpulist" % numa).read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
except Exception as e:  # topology files are best effort
    print("rank %d: topology lookup failed: %s" % (rank, e))
from pygofpid_b200.hotpath import PinnedFrames
for mb in ([int(v) for v in args.sizes.split(',')]):
    n = mb * 1024 * 1024
    if args.wc:
        keep = PinnedFrames((n,), write_combined=True)
        keep.array[:] = 7
        h = torch.from_numpy(keep.array)
    else:
        h = torch.full((n,), 7, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 400 if mb < 64 else 40
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    gbs = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    tot, mn = t.clone(), t.clone()
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ranks=%d affinity=%d wc=%d numa(gpu0)=%s  H2D %2d MB copies: min rank %.1f GB/s, sum %.1f GB/s  -> floor %.0f "
              "1080p gray frames/s per GPU" % (world, args.affinity, args.wc, numa, mb, float(mn), float(tot),
                                               float(mn) * 1e9 / (1920 * 1080)), flush=True)
    del h, d
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
