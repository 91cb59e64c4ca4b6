(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_scale_properties.py -m gpu -x -q -k "model or layer or golden or locality or hub" 2>&1 | tail -4) > gpurun_out/t_pytest.log 2>&1; tail -3 gpurun_out/t_pytest.log
for v in ring reg; do SPK_AGG_FWD=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
p=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=p['kernels_ms_per_step']
print('$v', round(p['ms_per_step'],3), 'agg_fwd', k.get('agg_fwd'))"; done
