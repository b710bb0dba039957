"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of metrics quoted in DESIGN.md.
Usage: python tools/ncu_summary.py raw.csv out.csv"""
import csv
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_x", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
units = rows[1]
cols = [i for i, h in enumerate(hdr) if h in KEEP or any(h.startswith(k) for k in ("sm__pipe_tensor", "sm__mem_tensor"))]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in cols])
    w.writerow([units[i] for i in cols])
    for r in rows[2:]:
        if len(r) >= len(hdr):
            w.writerow([r[i] for i in cols])
print(f"{len(rows) - 2} launches, {len(cols)} columns -> {sys.argv[2]}")
