"""Host-side diagnostic for the GPU box: which CPUs are really usable and how the oracle step scales with threads."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R]
import numpy as np
from oracle import pyoracle as po
import bench
print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "usable_cpus", po.usable_cpus())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/cpu.stat"):
    try: print(f, "->", open(f).read().strip().replace("\n", " | ")[:300])
    except Exception as e: print(f, "unreadable", e)
print(os.popen("lscpu | grep -E 'Model name|Socket|Thread|NUMA node\\(s\\)'").read())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for t in [int(a) for a in sys.argv[2:]] or [16, 32, 64, 128]:
    t0 = time.time()
    v, dt = bench.cpu_baseline(B, t)
    print("threads %3d  B=%d  %.2f s/step  %.2f images/s" % (t, B, dt, v), flush=True)
