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
