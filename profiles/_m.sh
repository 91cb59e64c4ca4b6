(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm_nn" 2>&1 | tail -6) > gpurun_out/m_pytest.log 2>&1; tail -4 gpurun_out/m_pytest.log
SPK_TC_PAIR=0 timeout 200 python profiles/bench_kernels.py --what gemm --reps 10 2>&1 | grep gemm_nn > gpurun_out/m_gemm_pair0.jsonl
SPK_TC_PAIR=1 timeout 200 python profiles/bench_kernels.py --what gemm --reps 10 2>&1 | grep gemm_nn > gpurun_out/m_gemm_pair1.jsonl
paste -d'\n' gpurun_out/m_gemm_pair0.jsonl gpurun_out/m_gemm_pair1.jsonl | cut -c1-160
