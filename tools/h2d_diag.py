"""Host -> device copy bandwidth per rank, alone and with every rank copying at once (512 MB pinned buffers), plus the CPU
affinity / NUMA placement bench.bind_to_gpu_numa ends up with.  Explains the end-to-end leg at 8 GPUs.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_diag.py   [BIND=0 to skip binding]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import torch, torch.distributed as dist
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
import bench
bind = os.environ.get("BIND", "1") == "1"
ncpu = bench.bind_to_gpu_numa(torch, lr) if bind else -1
dist.init_process_group("nccl", device_id=dev)
p = torch.cuda.get_device_properties(lr)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
try:
    node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
except Exception as e:
    node = "?"
aff = sorted(os.sched_getaffinity(0))
def numa_of_cpu(c):
    for n in range(8):
        try:
            txt = open("/sys/devices/system/node/node%d/cpulist" % n).read().strip()
        except Exception:
            continue
        for part in txt.split(","):
            a, _, b = part.partition("-")
            if int(a) <= c <= int(b or a):
                return n
    return -1
nbytes = 512 << 20
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
def bw(n=8):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return n * nbytes / (time.perf_counter() - t0) / 1e9
all_bw = bw()
# one rank at a time
solo = 0.0
for r in range(world):
    dist.barrier()
    if r == rank:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(4): d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(); solo = 4 * nbytes / (time.perf_counter() - t0) / 1e9
    dist.barrier()
print("rank %d gpu %s gpu_numa %s bind_cpus %d affinity %d cpus (first %s -> numa %d) nodes %s | H2D all-ranks %.1f GB/s solo %.1f GB/s" %
      (rank, bus, node, ncpu, len(aff), aff[:2], numa_of_cpu(aff[0]), os.listdir("/sys/devices/system/node")[:6], all_bw, solo), flush=True)
dist.barrier(); dist.destroy_process_group()
