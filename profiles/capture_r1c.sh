# Round-1 (late session) capture 2: N3 sampler tests + timing, ncu --set full of the N1 kernels, then the whole GPU suite.
mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_sampler.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/h_sampler.log 2>&1; tail -4 gpurun_out/h_sampler.log
(timeout 90 python profiles/bench_sampler.py 2>&1 | tail -2) > gpurun_out/h_sampler_bench.json 2>&1; cut -c1-400 gpurun_out/h_sampler_bench.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"loss_norm_kernel|loss_bwd_seg_kernel|loss_pairs_kernel|sgd_kernel" -s 8 -c 5 -o /tmp/h_loss python profiles/bench_loss.py --steps 1 > gpurun_out/h_ncu.log 2>&1; tail -2 gpurun_out/h_ncu.log
ncu -i /tmp/h_loss.ncu-rep --page raw --csv > /tmp/h_loss_raw.csv 2>/dev/null
python - <<'PY'
import csv
try:
    rows=list(csv.reader(open('/tmp/h_loss_raw.csv')))
    hdr=rows[0]
    keep=[i for i,h in enumerate(hdr) if h in ('ID','Kernel Name','Grid Size','Block Size') or any(k in h for k in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct','dram__throughput.avg.pct','lts__throughput.avg.pct','sm__warps_active.avg.pct','smsp__issue_active.avg.pct','launch__registers_per_thread','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct')) and '.max' not in h and '.min' not in h]
    w=csv.writer(open('gpurun_out/h_loss_ncu_sel.csv','w'))
    for r in rows: w.writerow([r[i] for i in keep])
    print("ncu rows", len(rows)-2)
except Exception as e:
    print("ncu export failed", e)
PY
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/h_pytest_all.log 2>&1; tail -3 gpurun_out/h_pytest_all.log
