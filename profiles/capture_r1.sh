# Round-1 capture script: GPU tests, bench (both arms), ncu launch list of one step, ncu --set full of the top kernels.
# Run on a B200 box from the repo root: bash profiles/capture_r1.sh ; outputs land in gpurun_out/ (summaries copied to profiles/).
set -x
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/f_pytest.log 2>&1; cat gpurun_out/f_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -1 gpurun_out/f_bench.json | cut -c1-400
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_ref.json 2>/dev/null; cat gpurun_out/f_ref.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 260 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"split_cols_kernel|split_rels_tasks|edge_fwd_stream_kernel|agg_fwd_stream|agg_bwd_stream|bwd_node|gemm_nn_tc_kernel|gemm_tn_tc_kernel|seg_gather_kernel|seg_gather_tasks" -s 22 -c 22 -o /tmp/f_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu.log 2>&1; tail -2 gpurun_out/f_ncu.log
ncu -i /tmp/f_full.ncu-rep --page raw --csv > /tmp/f_full_raw.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/f_full_raw.csv')))
hdr=rows[0]
keep=[i for i,h in enumerate(hdr) if h in ('ID','Kernel Name','Grid Size','Block Size') or any(k in h for k in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct','dram__throughput.avg.pct','lts__throughput.avg.pct','l1tex__throughput.avg.pct','sm__warps_active.avg.pct','smsp__issue_active.avg.pct','launch__registers_per_thread','smsp__inst_executed.sum','sm__pipe_tensor_cycles_active','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct','smsp__average_warps_issue_stalled','lts__t_sector_hit_rate.pct','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum','sm__throughput.avg.pct')) and '.max' not in h and '.min' not in h]
w=csv.writer(open('gpurun_out/f_full_sel.csv','w'))
for r in rows: w.writerow([r[i] for i in keep])
PY
ls -la gpurun_out/ | tail -8
