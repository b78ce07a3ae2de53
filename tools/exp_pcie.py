"""Host->device copy rate of pinned buffers vs where they were first touched (GPU-local cores / the rest / unbound)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc.pytorch_b200.parallel import gpu_local_cpus, numa_local
dev = torch.device("cuda:0")
torch.cuda.init()
allowed = os.sched_getaffinity(0)
local = gpu_local_cpus(dev)
print("allowed cpus", len(allowed), "gpu-local", None if local is None else (len(local), min(local), max(local)))
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
db = torch.empty(64 << 20, dtype=torch.float32, device=dev)

def rate(hb, reps=5):
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); db.copy_(hb, non_blocking=True); e1.record(); torch.cuda.synchronize()
        best = max(best, hb.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best

def alloc(cpus):
    saved = os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
    hb = torch.zeros(64 << 20, dtype=torch.float32).pin_memory()
    os.sched_setaffinity(0, saved)
    return hb

print("unbound first touch: %.1f GB/s" % rate(alloc(None)))
if local:
    print("gpu-local cores    : %.1f GB/s" % rate(alloc(local)))
    far = allowed - local
    if far:
        print("far cores          : %.1f GB/s" % rate(alloc(far)))
    os.sched_setaffinity(0, local)
    print("gpu-local, thread also bound: %.1f GB/s" % rate(alloc(local)))
    os.sched_setaffinity(0, allowed)
