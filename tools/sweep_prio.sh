#!/bin/bash
# e2e with stream priorities per lane x lanes x chunk
for prio in 0 1; do for lanes in 3 6 8; do for chunk in 98304 131072; do
  SMX_PIPELINE_PRIORITIES=$prio SMX_PIPELINE_LANES=$lanes python bench.py --no-cpu-baseline --chunk $chunk 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('prio $prio lanes $lanes chunk $chunk e2e %.0f M reads/s (%.3f ms, %d chunks) resident %.0f M' % (d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['chunks'], d['value']/1e6))"
done; done; done
SMX_PIPELINE_TRACE=1 SMX_PIPELINE_PRIORITIES=1 SMX_PIPELINE_LANES=6 python bench.py --no-cpu-baseline --steps 2 2>&1 | grep "smx pipeline" | tail -7
