for m in split rows; do SPK_AGG_BWD_MODE=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-strong 2>/dev/null > gpurun_out/u_bench_$m.json; python - <<P
import json
p=json.loads([l for l in open('gpurun_out/u_bench_$m.json').read().splitlines() if l.startswith('{')][-1]); k=p['kernels_ms_per_step']
print('$m', round(p['ms_per_step'],3), {a:b for a,b in k.items() if b>0.25})
P
done
timeout 400 python profiles/trace_gaps.py gpurun_out/u_trace_gaps.json 2>&1 | tail -3
