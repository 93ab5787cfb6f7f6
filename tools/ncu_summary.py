"""Summarise an ncu report for profiles/: python tools/ncu_summary.py report.ncu-rep kernel_regex out.json
(reads `ncu -i report --page raw --csv`; keeps the metrics the roofline discussion in DESIGN.md uses)."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]

rep, rx, out = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
res = []
for r in rows[2:]:
    if len(r) != len(hdr) or not rx.search(r[ki]):
        continue
    d = {"kernel": r[ki][:80]}
    for i, h in enumerate(hdr):
        if h in KEEP or h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed") \
                or h.endswith("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"):
            d[h] = [r[i], units[i]]
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
print(f"{len(res)} launches matching {rx.pattern} -> {out}")
