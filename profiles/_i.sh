(timeout 900 python -m pytest tests/test_convkb.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/i_pytest.log 2>&1; tail -4 gpurun_out/i_pytest.log
timeout 600 python profiles/bench_nhop.py > gpurun_out/i_nhop.json 2> gpurun_out/i_nhop.err; tail -c 300 gpurun_out/i_nhop.err; cat gpurun_out/i_nhop.json
