#!/bin/bash
# scratch: env-variant sweep of the quick bench.  bash profiles/_v.sh <tag> "ENV=.. ENV=.." "ENV=.." ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-strong --no-cpu-baseline --no-e2e > gpurun_out/${tag}_v${i}.json 2> gpurun_out/${tag}_v${i}.err
  echo "== $envs"; python - <<PY
import json
for l in open("gpurun_out/${tag}_v${i}.json"):
    if l.startswith('{'):
        d=json.loads(l); k=d['kernels_ms_per_step']
        print(round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items() if b>0.9})
PY
done
