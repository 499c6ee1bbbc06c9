"""GPU-box tool: what the host<->device interface can carry (the e2e path of bench.py is bound by it).
Pinned 18.4 MB buffers (one 9000 x 512 fp32 bag), H2D alone, D2H alone, both directions at once; once
with the default CPU affinity and once per NUMA node (first-touch places the pinned pages)."""
import glob, os, subprocess, sys, time
import torch

def cpus_of(node):
    s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

def measure(tag, nbuf=8, n=9000 * 512, reps=20):
    dev = torch.device("cuda", 0)
    hx = [torch.empty(n).pin_memory() for _ in range(nbuf)]
    for h in hx: h.fill_(1.0)                      # touch
    hy = [torch.empty(n).pin_memory() for _ in range(nbuf)]
    for h in hy: h.fill_(0.0)
    dx = [torch.empty(n, device=dev) for _ in range(nbuf)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            for i in range(nbuf):
                if h2d:
                    with torch.cuda.stream(s1): dx[i].copy_(hx[i], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): hy[i].copy_(dx[(i + 3) % nbuf], non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbuf * n * 4 / (time.perf_counter() - t0) / 1e9
    run(True, True)
    print(f"{tag}: H2D {run(True, False):.1f} GB/s | D2H {run(False, True):.1f} GB/s | both {run(True, True):.1f} GB/s each way", flush=True)

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
print("cpus allowed:", len(os.sched_getaffinity(0)), "nodes:", sorted(glob.glob("/sys/devices/system/node/node[0-9]*")))
measure("default affinity")
allowed = os.sched_getaffinity(0)
for nd in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    node = int(nd.rsplit("node", 1)[1])
    c = set(cpus_of(node)) & allowed
    if not c: continue
    os.sched_setaffinity(0, c)
    measure(f"affinity node{node} ({len(c)} cpus)")
os.sched_setaffinity(0, allowed)
