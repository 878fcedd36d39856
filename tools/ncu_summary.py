# coding=utf-8
"""Print selected metrics of every kernel in an .ncu-rep as a markdown table (run on the CPU box)."""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg.per_second']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
for r in rows[2:]:
  print("**`%s`**\n" % r[hdr.index('Kernel Name')].split('(')[0].replace('void ', ''))
  print("| metric | value |\n|---|---|")
  for w in WANT:
    if w in hdr:
      i = hdr.index(w)
      print("| %s | %s %s |" % (w, r[i], rows[1][i]))
  print()
