"""Selected metrics of an `ncu --page raw --csv` export (last launch in the file): python tools/ncu_pick.py file.raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[-1]
d = dict(zip(hdr, vals))
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_red.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor', 'launch__registers_per_thread',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__occupancy_per_block_size']
for k in keys:
    for h in hdr:
        if h == k or h.endswith('.' + k):
            print(f"  {h:95s} {d[h]}")
