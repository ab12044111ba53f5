#!/bin/bash
# ncu --set full capture of the one-lane attribution step only (see tools/gpu_round.sh)
TAG=${1:-dev}; SKIP=${SKIP:-27}; OUT=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_primer_sliced|k_primer_finish|k_primer_start|k_primer_long|k_barcode_task|k_select|k_stage_windows|k_rebase_offsets|k_scan" \
  --launch-skip $SKIP -c 11 -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 1 --split 1 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
