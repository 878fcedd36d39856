#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err; tail -3 gpurun_out/c18_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c18_bench.json").read().strip().splitlines()[-1])
def show(n, d):
  print(n, round(d["value"],1), round(d["ms_per_step"],1), d["clocks"], d["roofline"]["frac"], d["roofline"].get("avg_launch_ms"), d["roofline"].get("cell_share_of_step"), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("h2d_bytes_per_step"), d.get("cpu_baseline") and d["cpu_baseline"]["value"])
show("c4", d)
for k, v in d.get("extra", {}).items(): show(k, v)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/c18_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/c18_launches.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if hdr:
  h=rows[hdr[0]]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
  agg=collections.defaultdict(lambda:[0,0.0])
  for r in rows[hdr[0]+2:]:
    if len(r)>vi:
      name=r[ki].split("(")[0][:60]; v=float(r[vi].replace(",","")); 
      if r[ui]=="us": v/=1e3
      if r[ui]=="ns": v/=1e6
      agg[name][0]+=1; agg[name][1]+=v
  tot=sum(v[1] for v in agg.values())
  for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print("%-62s n=%4d %9.3f ms %5.1f%%"%(k,v[0],v[1],100*v[1]/tot))
PY
