"""Cuts an `ncu --page raw --csv` export down to the columns the roofline / README tables use (one row per launch)."""
import csv
import sys

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
        'dram__throughput.avg.pct', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'sm__warps_active.avg.pct',
        'smsp__issue_active.avg.pct', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct',
        'smsp__average_warps_issue_stalled', 'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'sm__throughput.avg.pct')
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ('ID', 'Kernel Name', 'Grid Size', 'Block Size')
        or (any(k in h for k in KEEP) and '.max' not in h and '.min' not in h)]
w = csv.writer(open(sys.argv[2], 'w'))
for r in rows:
    w.writerow([r[i] for i in keep])
