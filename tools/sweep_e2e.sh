#!/bin/bash
# e2e (smx_match_batch, host buffers) over pipeline lanes x chunk size; prints reads/s per combination
for lanes in 3 4 6 8; do for chunk in 49152 65536 98304 131072 196608; do
  SMX_PIPELINE_LANES=$lanes python bench.py --no-cpu-baseline --chunk $chunk 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('lanes $lanes chunk $chunk e2e %.0f M reads/s (%.3f ms, %d chunks) resident %.0f M' % (d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['chunks'], d['value']/1e6))"
done; done
