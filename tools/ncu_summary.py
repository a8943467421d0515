"""Key metrics of an `ncu --set full` report as JSON: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv"""
import csv
import json
import re
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_uniform.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__cluster_size", "sm__cycles_elapsed.avg", "sm__cycles_active.avg")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = {}
for r in rows[2:]:
    name = re.sub(r"^void |sed::|\(.*$", "", r[hdr.index("Kernel Name")])
    key, n = name, 1
    while key in out:
        n += 1
        key = "%s #%d" % (name, n)
    out[key] = {k: ("%s %s" % (r[hdr.index(k)], units[hdr.index(k)])).strip() for k in KEYS if k in hdr}
print(json.dumps(out, indent=1))
