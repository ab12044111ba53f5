#!/usr/bin/env python3
"""e2e probe: smx_match_batch with host buffers (wire-form records) over pipeline lanes x chunk size on the resident
ont037 batch; one traced run per setting when SMX_PROBE_TRACE=1.  usage: python tools/e2e_probe.py [reads]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import helpers as H  # noqa: E402
from specimux_b200.engine import Matcher, PackedBatch  # noqa: E402
from specimux_b200.models import MatchParameters  # noqa: E402
from specimux_b200.orchestration import thresholds_for  # noqa: E402
from specimux_b200.tables import MatchTables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 765_000
ds = bench.dataset("ont037", n, 0)
specimens = H.build_specimens(ds.primers, ds.specimens)
k_idx, k_primers = thresholds_for(specimens)
params = MatchParameters(k_primers, k_idx, ds.search_len, True)
tables = MatchTables(specimens, params)
blob = np.frombuffer(b"ACGT", dtype=np.uint8)[ds.codes].tobytes()
batch = PackedBatch.from_blob(blob, ds.offsets.astype(np.uint64), clip=ds.search_len)
for chain, lanes in ((1, 3), (1, 4), (1, 6), (0, 3)):
    os.environ["SMX_PIPELINE_LANES"] = str(lanes)
    os.environ["SMX_PIPELINE_CHAIN"] = str(chain)
    for chunk in (49152, 65536, 98304, 131072, 196608, 262144):
        with Matcher(tables) as m:
            m.set_pipeline_chunk(chunk)
            for _ in range(3):
                m.match(batch, reuse=True, compact="wire")
            t0 = time.perf_counter()
            for _ in range(10):
                m.match(batch, reuse=True, compact="wire")
            ms = (time.perf_counter() - t0) * 100.0
            print("chain %d lanes %d chunk %6d: %.3f ms  %.0f M reads/s  (%d chunks)" % (chain, lanes, chunk, ms, n / ms / 1e3, m.last_chunk_count()), flush=True)
if os.environ.get("SMX_PROBE_TRACE"):
    os.environ["SMX_PIPELINE_LANES"] = "4"
    os.environ["SMX_PIPELINE_CHAIN"] = "1"
    os.environ["SMX_PIPELINE_TRACE"] = "1"
    with Matcher(tables) as m:
        for _ in range(3):
            m.match(batch, reuse=True, compact="wire")
