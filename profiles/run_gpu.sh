#!/bin/bash
# Round-2 GPU command lists (run on a B200 box from the repo root through gpurun; outputs land in gpurun_out/).
#   bash profiles/run_gpu.sh <tag> tests|bench|ref|launches|ncu|all ...
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)    (timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_pytest.log 2>&1; tail -5 gpurun_out/${tag}_pytest.log ;;
bench)    timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.err; tail -1 gpurun_out/${tag}_bench.json | cut -c1-300 ;;
benchq)   timeout 600 python bench.py --steps 10 --warmup 3 --no-strong --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.err; tail -1 gpurun_out/${tag}_bench.json | cut -c1-300 ;;
ref)      timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${tag}_ref.json ;;
smoke)    timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log ;;
launches) timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-strong > /dev/null 2>&1; wc -l gpurun_out/${tag}_launches.csv ;;
ncu)      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"split_|edge_fwd_stream_kernel|agg_fwd_reg|agg_fwd_stream|agg_bwd_stream|agg_bwd_ctx|agg_bwd_pre|agg_dx|agg_table|bwd_node|gemm_nn_tc|gemm_tn_tc|seg_gather_kernel|seg_gather_tasks|residual_norm" -s 40 -c 64 -o /tmp/${tag}_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-strong > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
          ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > /tmp/${tag}_full_raw.csv && python profiles/ncu_select.py /tmp/${tag}_full_raw.csv gpurun_out/${tag}_full_sel.csv ;;
esac
done
ls -la gpurun_out | tail -12
