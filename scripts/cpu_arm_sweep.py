#!/usr/bin/env python3
"""Diagnostic for the reference arm: cgroup CPU limits of the box + thread sweep of the in-place oracle arena."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/sys/fs/cgroup/cpuset.cpus.effective", "/sys/fs/cgroup/cpuset/cpuset.cpus"):
  try:
    print(p, open(p).read().strip())
  except OSError:
    pass
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
print(open("/proc/loadavg").read().strip())
def stat():
  for p in ("/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu/cpu.stat"):
    try:
      return open(p).read().replace("\n", " ")
    except OSError:
      pass
  return ""
from bench import CpuArm
for nt in (1, 4, 8, 16, 32, 64, 128):
  if nt > (os.cpu_count() or 1):
    break
  arm = CpuArm("live", 4096 * max(4, nt), nthreads=nt)
  arm.step()
  s0 = stat()
  t = time.perf_counter()
  for _ in range(3):
    arm.step()
  el = time.perf_counter() - t
  print(f"threads {nt:4d}: {3 * arm.B / el:12.0f} steps/s  ({3 * arm.B / el / nt:9.0f} per thread)  B={arm.B}")
  arm.close()
print("cpu.stat before last:", s0)
print("cpu.stat after:", stat())
