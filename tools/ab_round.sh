#!/bin/bash
# A/B of the library's environment switches on one box: tools/ab_round.sh TAG "VAR=a VAR2=b" "VAR=c" ...
# (each argument after TAG is one environment; bench.py without the CPU and file legs)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
i=0
for envs in "$@"; do
  echo "== $envs"
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-file-to-tree $BENCH_ARGS > $OUT/ab_${TAG}_$i.json 2> $OUT/ab_${TAG}_$i.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open("$OUT/ab_${TAG}_$i.json").read().strip().splitlines()[-1])
print(round(d['value']/1e6), round(d['ms_per_step'],4), round(d['serial_ms_per_step'],4), {k:round(v*1000) for k,v in d['kernel_ms'].items()}, 'e2e', round(d['e2e']['value']/1e6))
PY
  i=$((i+1))
done
