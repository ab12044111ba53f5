#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the stage kernels.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh TAG [skip-tests]
TAG=${1:-dev}
# 9 kernels per resident step on one lane (stage, sliced x 2 primers, finish, start, barcode, 2 x select, scan_compact); skip = split pass (warm-up + step) + attribution warm-up
SKIP=${SKIP:-27}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_$TAG.log
fi
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_primer_sliced|k_primer_finish|k_primer_start|k_primer_long|k_barcode_task|k_select|k_stage_windows|k_rebase_offsets|k_scan" \
  --launch-skip $SKIP -c 11 -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 1 --split 1 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
