#!/usr/bin/env python
"""Host topology + pinned-memory H2D bandwidth probe (why the e2e arm stopped scaling at 4-8 GPUs in round 1).

Prints one JSON object: CPU affinity, cgroup cpuset, NUMA nodes, each GPU's PCI address / NUMA node / local
cpulist, and the H2D bandwidth of a 1 GiB pinned buffer per (GPU, memory policy) where the policy is the default
first-touch or MPOL_BIND to each NUMA node (set_mempolicy through libc; no numactl in the image)."""
import ctypes
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from realpdebench_b200 import dist as D  # noqa: E402


def read(p):
    try:
        with open(p) as f:
            return f.read().strip()
    except OSError as e:
        return f"<{e.__class__.__name__}>"


def main():
    out = {"affinity": sorted(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(),
           "cpuset_cpus": read("/sys/fs/cgroup/cpuset.cpus.effective"),
           "cpuset_mems": read("/sys/fs/cgroup/cpuset.mems.effective"),
           "numa_nodes": {os.path.basename(n): read(n + "/cpulist") for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}}
    try:
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as e:
        out["topo"] = repr(e)
    gpus = []
    for i in range(torch.cuda.device_count()):
        info = D.gpu_host_locality(i)
        gpus.append(info)
    out["gpus"] = gpus
    nodes = sorted(int(os.path.basename(n)[4:]) for n in glob.glob("/sys/devices/system/node/node[0-9]*"))
    bw = []
    nbytes = 1 << 30
    for i in range(torch.cuda.device_count()):
        torch.cuda.set_device(i)
        dst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}")
        for pol in [None] + nodes:
            ok = D.set_memory_policy(pol)
            try:
                src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                src.fill_(1)
            finally:
                D.set_memory_policy(None)
            best = 0.0
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dst.copy_(src, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
            bw.append({"gpu": i, "policy": "default" if pol is None else f"bind node{pol}", "policy_applied": ok,
                       "h2d_gbs": round(best, 2)})
            del src
        del dst
    out["h2d"] = bw
    print(json.dumps(out))


if __name__ == "__main__":
    main()
