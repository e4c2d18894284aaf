"""Key metrics of an ncu --set full capture as a small CSV: python tools/ncu_summary.py capture.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
idx = [(w, h.index(w)) for w in WANT if w in h]
w = csv.writer(sys.stdout)
w.writerow([f"{n} [{units[i]}]" if units[i] else n for n, i in idx])
for r in rows[2:]:
    w.writerow([r[i][:110] for _, i in idx])
