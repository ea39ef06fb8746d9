"""Host-to-device bandwidth of a pinned 22.5 MB staging buffer, allocated (first-touched) with the process bound to each
NUMA node's CPUs in turn; prints the node the GPU hangs off.  Development probe for bench.py's e2e leg."""
import os, glob, time, torch
def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part: continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out
p = torch.cuda.get_device_properties(0)
bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
base = "/sys/bus/pci/devices/" + bdf
for f in ("numa_node", "local_cpulist"):
    try: print(f, open(os.path.join(base, f)).read().strip())
    except Exception as e: print(f, "unreadable", e)
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes", [os.path.basename(n) for n in nodes], "cpus allowed", len(os.sched_getaffinity(0)))
torch.zeros(1, device="cuda")
full = os.sched_getaffinity(0)
for n in nodes + [None]:
    if n is not None:
        cpus = set(cpulist(open(n + "/cpulist").read())) & full
        if not cpus: continue
        os.sched_setaffinity(0, cpus)
    else:
        os.sched_setaffinity(0, full)
    h = torch.empty(22_559_232, dtype=torch.uint8).pin_memory(); h.fill_(1)
    d = torch.empty_like(h, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(os.path.basename(n) if n else "unbound", "H2D %.3f ms  %.1f GB/s" % (ms, h.numel() / ms / 1e6))
