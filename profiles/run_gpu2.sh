#!/bin/bash
# Multi-GPU command list: bash profiles/run_gpu2.sh <tag> <ngpus> [check] [bench] [benchq]
tag=$1; n=$2; shift; shift
mkdir -p gpurun_out
for what in "$@"; do
case $what in
check)  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/${tag}_check.log 2>&1; grep "DIST PARITY" gpurun_out/${tag}_check.log || tail -30 gpurun_out/${tag}_check.log ;;
bench)  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err; tail -c 1500 gpurun_out/${tag}_bench${n}.err; tail -1 gpurun_out/${tag}_bench${n}.json | cut -c1-400 ;;
benchq) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 10 --warmup 3 --no-strong --no-e2e > gpurun_out/${tag}_benchq${n}.json 2> gpurun_out/${tag}_benchq${n}.err; tail -c 1500 gpurun_out/${tag}_benchq${n}.err; tail -1 gpurun_out/${tag}_benchq${n}.json | cut -c1-400 ;;
esac
done
