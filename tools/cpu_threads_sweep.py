# coding=utf-8
"""How many host threads give the torch-CPU reference restatement its best throughput?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print("cpu_count", os.cpu_count())
for th in (8, 16, 32, 64, 128):
  if th > (os.cpu_count() or 1):
    break
  os.environ["MVB_CPU_THREADS"] = str(th)
  v, dt, t = bench.cpu_reference_run(bench.WORKLOADS["c4"]["cfg"], 1, 1)
  print("threads %d: %.3f traj/s (%.1f s)" % (t, v, dt), flush=True)
