#!/bin/bash
# usage: tools/ncu_metrics.sh <rep> : per-kernel key metrics of an `ncu --set full` report as text
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rd = csv.reader(sys.stdin)
hdr = next(rd); units = next(rd)
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__cycles_active.avg"]
idx = {h: i for i, h in enumerate(hdr)}
for row in rd:
    out = []
    for w in want:
        cands = [h for h in hdr if h == w or h.startswith(w)]
        if cands:
            i = idx[cands[0]]
            out.append("%s=%s%s" % (w.split(".")[0] if w != "Kernel Name" else "kernel", row[i][:70], (" " + units[i]) if units[i] else ""))
    print(" | ".join(out))
'
