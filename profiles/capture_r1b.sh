# Round-1 (late session) capture: the new N1 kernels and the full-size parity test, then a bench line of the same build.
# Run on a B200 box from the repo root: bash profiles/capture_r1b.sh ; outputs land in gpurun_out/.
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_loss.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/g_loss.log 2>&1; tail -5 gpurun_out/g_loss.log
(timeout 60 python profiles/bench_loss.py --steps 5 2>&1 | tail -3) > gpurun_out/g_loss_bench.json 2>&1; cat gpurun_out/g_loss_bench.json | cut -c1-600
(timeout 240 python -m pytest tests/test_scale_properties.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/g_scale.log 2>&1; tail -5 gpurun_out/g_scale.log
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/g_smoke.log 2>&1; cat gpurun_out/g_smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -1 gpurun_out/g_bench.json | cut -c1-500
