"""oracle/cpu_bench.py -- TEST / MEASUREMENT INFRASTRUCTURE.

Times the CPU restatement of the reference path (oracle.pipeline over the C aligner) on host
cores.  Used only by bench.py's `cpu_baseline` leg and `--impl reference` arm.  The real
reference (specimux + edlib) cannot run on the GPU box (edlib/biopython are not installable and
/root/reference is absent there), so this port -- verified op-for-op against the unmodified
reference in oracle/make_goldens.py -- stands in for it (kind = "port").
"""
import multiprocessing as mp
import os
import time

from . import pipeline as orc

_state = {}


def _init(primers, specimens, search_len):
    tables = orc.Tables(primers, specimens)
    params = orc.setup_params(tables, search_len=search_len)
    _state["tables"], _state["params"] = tables, params


def _work(chunk):
    ops, total, matched = orc.process_reads(_state["tables"], _state["params"], chunk)
    return total, matched, len(ops)


def run(primers, specimens, reads, search_len=80, processes=1, batch=250):
    """Process `reads` [(id, bases, quals)] with `processes` workers (1000-read batches in the
    reference, orchestration.py:165; smaller here so short samples still use every core).
    Returns dict(seconds, reads, matched, reads_per_s, cores)."""
    chunks = [reads[i:i + batch] for i in range(0, len(reads), batch)]
    t0 = time.perf_counter()
    if processes <= 1:
        _init(primers, specimens, search_len)
        res = [_work(c) for c in chunks]
    else:
        with mp.get_context("fork").Pool(processes, initializer=_init, initargs=(primers, specimens, search_len)) as pool:
            t0 = time.perf_counter()          # exclude pool start-up / table build, as the reference's timer does
            res = pool.map(_work, chunks)
    dt = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    return dict(seconds=dt, reads=total, matched=sum(r[1] for r in res), records=sum(r[2] for r in res),
                reads_per_s=total / dt if dt > 0 else 0.0, cores=max(1, processes))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# Record-level oracle output for the full-size parity tests (tests/test_gpu_full_configs.py): the
# write-ops of oracle.pipeline.process_reads flattened into one numpy row per op, computed by forked
# workers over index ranges of a shared ASCII blob (no per-read pickling).

import numpy as np

ORACLE_RECORD_DTYPE = np.dtype([("read", "<u4"), ("sample", "<i4"), ("trim_start", "<i4"), ("trim_end", "<i4"),
                                ("p1_loc", "<i4", (2,)), ("p2_loc", "<i4", (2,)), ("b1_loc", "<i4", (2,)),
                                ("b2_loc", "<i4", (2,)), ("pool", "<i2"), ("p1", "<i2"), ("p2", "<i2"),
                                ("dist", "i1", (4,)), ("resolution", "u1"), ("reverse", "u1"), ("trim_empty", "u1")])
_NONE = -(2 ** 31)
_rec_state = {}


def _rec_init(primers, specimens, param_kw, maps):
    tables = orc.Tables(primers, specimens)
    _rec_state["tables"] = tables
    _rec_state["params"] = orc.setup_params(tables, **param_kw)
    _rec_state["maps"] = maps


def _rec_work(span):
    lo, hi = span
    blob, offs = _rec_state["blob"], _rec_state["offs"]
    spec_id, b1_id, b2_id, pool_id, primer_id = _rec_state["maps"]
    reads = [(i, blob[int(offs[i]):int(offs[i + 1])].decode("latin-1"), None) for i in range(lo, hi)]
    ops, _total, matched = orc.process_reads(_rec_state["tables"], _rec_state["params"], reads)
    out = np.zeros(len(ops), dtype=ORACLE_RECORD_DTYPE)
    loc = lambda l: (_NONE, _NONE) if l is None else (l[0], l[1])
    for k, op in enumerate(ops):
        r = out[k]
        r["read"] = op.seq_id
        res = op.resolution_type
        if op.trim_empty:
            sample = -1
        elif res in (orc.FULL_MATCH, orc.DEREPLICATED_FULL, orc.MULTIPLE_SPECIMENS):
            sample = spec_id[op.sample_id]
        elif res == orc.PARTIAL_FORWARD:
            sample = b1_id[op.sample_id[len("barcode_fwd_"):]]
        elif res == orc.PARTIAL_REVERSE:
            sample = b2_id[op.sample_id[len("barcode_rev_"):]]
        else:
            sample = -1
        r["sample"] = sample
        r["trim_start"], r["trim_end"] = op.trim_start, op.trim_end
        r["p1_loc"], r["p2_loc"] = loc(op.p1_location), loc(op.p2_location)
        r["b1_loc"], r["b2_loc"] = loc(op.b1_location), loc(op.b2_location)
        r["pool"] = pool_id.get(op.primer_pool, -1)
        r["p1"] = primer_id.get(op.p1_name, -1)
        r["p2"] = primer_id.get(op.p2_name, -1)
        r["dist"] = [-1 if x == "X" else int(x) for x in op.distance_code.split(",")]
        r["resolution"] = res
        r["reverse"] = 1 if op.is_rc else 0
        r["trim_empty"] = 1 if op.trim_empty else 0
    return lo, out, matched


def run_records(primers, specimens, blob, offsets, maps, param_kw, processes=None, span=500):
    """All write-ops of reads [0, n) as ORACLE_RECORD_DTYPE rows in read order, plus the matched count.
    `blob` / `offsets`: concatenated ASCII reads; `maps` = (specimen id -> row, b1 -> id, b2 -> id, pool -> id,
    primer name -> canonical index) as the product's MatchTables numbers them; `param_kw` goes to
    oracle.pipeline.setup_params."""
    processes = processes or host_cores()
    n = len(offsets) - 1
    _rec_state["blob"], _rec_state["offs"] = blob, offsets          # inherited by the forked workers
    spans = [(lo, min(n, lo + span)) for lo in range(0, n, span)]
    t0 = time.perf_counter()
    if processes <= 1:
        _rec_init(primers, specimens, param_kw, maps)
        res = [_rec_work(s) for s in spans]
    else:
        with mp.get_context("fork").Pool(processes, initializer=_rec_init,
                                         initargs=(primers, specimens, param_kw, maps)) as pool:
            res = pool.map(_rec_work, spans, chunksize=1)
    res.sort(key=lambda x: x[0])
    dt = time.perf_counter() - t0
    return np.concatenate([r[1] for r in res]), sum(r[2] for r in res), dt
