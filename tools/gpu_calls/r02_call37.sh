#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_fwd_kernel -s 40 -c 1 -f -o gpurun_out/r02_cell_final python bench.py --workload c4 --no-extras --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/c37a.log 2>&1
ls -la gpurun_out/r02_cell_final.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_c4_b512_final.csv python bench.py --workload c4 --no-extras --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/c37b.log 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_c4_b512_final.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
  name = re.sub(r"\(.*", "", r[4])[:70]
  v = float(r[-1].replace(",", "")); u = r[-2]
  ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
  agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print("total %.1f ms over %d launches" % (tot, len(rows)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
  print("%-72s %5d %9.2f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
