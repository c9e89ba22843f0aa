"""Reduces an .ncu-rep (ncu --set full) to the handful of per-launch metrics quoted in profiles/ (not a pytest file).
usage: python tools/summarise_ncu.py report.ncu-rep > summary.csv"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
idx = {k: h.index(k) for k in KEYS if k in h}
w = csv.writer(sys.stdout)
w.writerow(["ID", "Kernel Name"] + list(idx))
w.writerow(["", ""] + [units[i] for i in idx.values()])
ki = h.index("Kernel Name")
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("bdetr::", "")
    w.writerow([r[0], name] + [r[i] for i in idx.values()])
