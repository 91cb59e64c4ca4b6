#!/bin/bash
# scratch: env-variant sweep of the bench incl. the e2e arm.  bash profiles/_e.sh <tag> "ENV=.." ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-strong --no-cpu-baseline > gpurun_out/${tag}_e${i}.json 2> gpurun_out/${tag}_e${i}.err
  echo "== $envs"; python - <<PY
import json
for l in open("gpurun_out/${tag}_e${i}.json"):
    if l.startswith('{'):
        d=json.loads(l)
        print("step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['ms_per_step'],2), "pipelined", round(d['e2e_pipelined']['ms_per_step'],2), "build", round(d['graph_build_ms'],2), d['clocks']['sm_mhz'])
PY
done
