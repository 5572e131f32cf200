"""Probe the GPU box: cores, memory, disk, H2D/D2H pinned bandwidth, page-cache read speed."""
import os, time, json, subprocess, torch
out = {}
out["nproc"] = os.cpu_count()
try:
    out["affinity"] = len(os.sched_getaffinity(0))
except Exception as e:
    out["affinity"] = str(e)
out["mem"] = subprocess.run("free -g | head -2", shell=True, capture_output=True, text=True).stdout
out["df"] = subprocess.run("df -h /tmp /root /dev/shm | tail -3", shell=True, capture_output=True, text=True).stdout
out["cpu"] = subprocess.run("lscpu | head -20", shell=True, capture_output=True, text=True).stdout
out["cgroup_cpu"] = subprocess.run("cat /sys/fs/cgroup/cpu.max 2>/dev/null", shell=True, capture_output=True, text=True).stdout
out["cgroup_mem"] = subprocess.run("cat /sys/fs/cgroup/memory.max 2>/dev/null", shell=True, capture_output=True, text=True).stdout
out["smi"] = subprocess.run("nvidia-smi --query-gpu=name,memory.total,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm --format=csv", shell=True, capture_output=True, text=True).stdout
out["tools"] = subprocess.run("which cargo rustc samtools ncu compute-sanitizer; ulimit -l", shell=True, capture_output=True, text=True).stdout
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (a, b) in {"h2d": (d, h), "d2h": (h, d)}.items():
    best = 0
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a.copy_(b, non_blocking=True); e1.record(); torch.cuda.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    out[name + "_GBs"] = best
# host memcpy speed (single thread) pageable -> pinned
p = torch.empty(n, dtype=torch.uint8); p.fill_(1)
t = time.time(); h.copy_(p); out["host_memcpy_GBs_1thread"] = n / (time.time() - t) / 1e9
print(json.dumps(out, indent=1))
open("gpurun_out/probe.json", "w").write(json.dumps(out, indent=1))
