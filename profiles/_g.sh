bash profiles/run_gpu2.sh g 2 check
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload c5 --steps 3 --warmup 2 --no-e2e --no-parity > gpurun_out/g_c5_2.json 2> gpurun_out/g_c5_2.err; tail -c 1200 gpurun_out/g_c5_2.err; grep '^{' gpurun_out/g_c5_2.json | tail -1 | cut -c1-500
