timeout 600 ncu --set full --clock-control none --import-source on -k regex:"agg_fwd_stream_kernel|agg_bwd_stream_kernel" -s 2 -c 2 -o /tmp/s_agg python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-strong > gpurun_out/s_ncu.log 2>&1
ncu -i /tmp/s_agg.ncu-rep --page source --csv --kernel-name regex:agg_fwd_stream_kernel > gpurun_out/s_src_fwd.csv 2>/dev/null
ncu -i /tmp/s_agg.ncu-rep --page source --csv --kernel-name regex:agg_bwd_stream_kernel > gpurun_out/s_src_bwd.csv 2>/dev/null
ls -la gpurun_out/s_*; head -c 600 gpurun_out/s_src_fwd.csv
