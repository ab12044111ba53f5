#!/bin/bash
# BASELINE.json configs 3-5 at one GPU (per-GPU share of the named read counts) + config 2 at N=2 when two GPUs are visible.
OUT=gpurun_out
TAG=${1:-v15}
python bench.py --config dense --reads 2000000 --cpu-sample 4000 --no-file-to-tree > $OUT/bench_${TAG}_dense.json 2> $OUT/bench_${TAG}_dense.err; tail -c 600 $OUT/bench_${TAG}_dense.err
python bench.py --config multipool --reads 625000 --cpu-sample 4000 --no-file-to-tree > $OUT/bench_${TAG}_multipool.json 2> $OUT/bench_${TAG}_multipool.err
python bench.py --config long --reads 1000000 --cpu-sample 2000 --no-file-to-tree > $OUT/bench_${TAG}_long.json 2> $OUT/bench_${TAG}_long.err
for c in dense multipool long; do python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_$c.json"))
    print("$c", "reads", d["config"]["reads_per_gpu"], "value %.1f M/s" % (d["value"]/1e6), "ms %.3f" % d["ms_per_step"], "e2e %.1f M/s" % (d["e2e"]["value"]/1e6), "gcups %.0f" % d["gcups"], "cells/read %.0f" % d["cells_per_read"], d["kernel_ms"], "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("$c failed", e)
PY
done
