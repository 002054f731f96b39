"""Scaling of the CPU arm (oracle/parallel.c: persistent threads, independent 8192-pair reference Pippenger sub-MSMs) with
the number of threads on this host, plus what limits the host's CPUs (affinity mask, cgroup quota).  Prints one JSON object."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench

out = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "physical_cores": bench.physical_cores()}
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try:
        out[f] = open(f).read().strip()
    except OSError:
        pass
try:
    out["loadavg"] = open("/proc/loadavg").read().strip()
except OSError:
    pass
rows = []
T = 1
counts = []
while T <= (os.cpu_count() or 1):
    counts.append(T); T *= 2
for T in counts:
    pool = bench.CpuPool(T)
    n = 8192 * T * 2
    sc, pts = pool.msm_inputs(n)
    pool.msm(sc, pts, n)
    dt = min(pool.msm(sc, pts, n)[0] for _ in range(2))
    rows.append({"threads": T, "pairs": n, "M_points_per_s": n / dt / 1e6, "per_thread": n / dt / 1e6 / T})
    pool.close()
out["msm_scaling"] = rows
print(json.dumps(out, indent=1))
