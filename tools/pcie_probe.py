"""PCIe / NUMA probe for the e2e leg of bench.py: pinned-buffer copy bandwidth with and without GPU-local placement."""
import glob
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectron.pytorch_b200 import benchutil

dev = torch.device("cuda:0")
print("affinity cpus:", len(os.sched_getaffinity(0)), "numa nodes:", [os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node[0-9]*")])
loc = benchutil.gpu_local_cpus(0)
print("gpu-local cpus:", None if loc is None else (len(loc), min(loc), max(loc)))
props = torch.cuda.get_device_properties(0)
bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
for f in ("numa_node", "current_link_speed", "current_link_width", "max_link_speed", "max_link_width"):
    try:
        print(f, open("/sys/bus/pci/devices/%s/%s" % (bdf, f)).read().strip())
    except OSError as e:
        print(f, "n/a", e)
N = 64 << 20


def bw(h_src, h_dst, tag):
    d_a = torch.empty(N // 4, device=dev); d_b = torch.randn(N // 4, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = []
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_src, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_dst.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res.append("%s %.1f GB/s" % (mode, 10 * N / dt / 1e9))
    print("%-28s %s" % (tag, "  ".join(res)))


a = torch.randn(N // 4).pin_memory(); b = torch.empty(N // 4).pin_memory()
bw(a, b, "pinned, default placement")
with benchutil.numa_local(0):
    c = torch.randn(N // 4).pin_memory(); d = torch.empty(N // 4).pin_memory()
bw(c, d, "pinned, numa_local()")
for node in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    cl = open(node + "/cpulist").read().strip()
    cpus = set()
    for part in cl.split(","):
        if "-" in part:
            x, y = part.split("-"); cpus.update(range(int(x), int(y) + 1))
        elif part:
            cpus.add(int(part))
    cpus &= os.sched_getaffinity(0)
    if not cpus:
        continue
    prev = os.sched_getaffinity(0)
    os.sched_setaffinity(0, cpus)
    e = torch.randn(N // 4).pin_memory(); f = torch.empty(N // 4).pin_memory()
    os.sched_setaffinity(0, prev)
    bw(e, f, "pinned on %s" % os.path.basename(node))
