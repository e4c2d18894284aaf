"""Summarise an ncu report (raw page) per kernel: duration, tensor-pipe %, DRAM bytes and throughput."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = {"Kernel Name": "name", "gpu__time_duration.sum": "us", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
        "launch__grid_size": "grid", "launch__block_size": "block", "launch__shared_mem_per_block_dynamic": "smem_dyn"}
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
print("| kernel | us | tensor % (active) | DRAM read | DRAM write | DRAM % | regs | grid x block | dyn smem |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    g = lambda k: r[idx[k]] if k in idx else "-"
    u = lambda k: units[idx[k]] if k in idx else ""
    print(f"| {g('Kernel Name')[:70]} | {g('gpu__time_duration.sum')} {u('gpu__time_duration.sum')} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | "
          f"{g('dram__bytes_read.sum')} {u('dram__bytes_read.sum')} | {g('dram__bytes_write.sum')} {u('dram__bytes_write.sum')} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{g('launch__registers_per_thread')} | {g('launch__grid_size')} x {g('launch__block_size')} | {g('launch__shared_mem_per_block_dynamic')} |")
