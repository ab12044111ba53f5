#!/usr/bin/env python3
"""Host-side I/O throughput of the native reader / packer / writer (no GPU needed: the records fed to the
writer are synthetic 'unknown' + full-match records).  usage: tools/io_bench.py [n_reads]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def main(n_reads):
    import helpers as H
    from specimux_b200 import _lib, native_io, synth
    from specimux_b200.engine import PackedBatch
    from specimux_b200.models import MatchParameters
    from specimux_b200.tables import MatchTables

    ds = synth.ont037(n_reads=n_reads, with_quals=False)
    tmp = tempfile.mkdtemp(prefix="smx_io_bench_")
    fq = os.path.join(tmp, "reads.fastq")
    t0 = time.time()
    ds.write_fastq(fq)
    size = os.path.getsize(fq)
    print("wrote %s: %d reads, %.1f MB in %.1fs" % (fq, n_reads, size / 1e6, time.time() - t0))
    specimens = H.build_specimens(ds.primers, ds.specimens)
    tables = MatchTables(specimens, MatchParameters({p[1]: 6 for p in ds.primers}, 3, 80, True))

    # reader + packer
    t_read = t_pack = 0.0
    blocks = []
    with native_io.FastxReader(fq, True) as rd:
        batch = None
        while True:
            blk = native_io.ReadBlock()
            t0 = time.perf_counter()
            rd.next_block(65536, blk)
            t_read += time.perf_counter() - t0
            if blk.n_reads == 0:
                break
            t0 = time.perf_counter()
            batch = PackedBatch.from_block(blk, clip=80, reuse=batch)
            t_pack += time.perf_counter() - t0
            blocks.append(blk)
    print("reader: %.3fs  %.0f MB/s  %.2f M reads/s" % (t_read, size / 1e6 / t_read, n_reads / 1e6 / t_read))
    print("packer: %.3fs  %.2f M reads/s" % (t_pack, n_reads / 1e6 / t_pack))

    # writer: every read a dereplicated full match of a random specimen, trimmed, half reversed
    rng = np.random.default_rng(1)
    out = os.path.join(tmp, "out")
    t_write = 0.0
    with native_io.TreeWriter(out, "", True, tables) as wr:
        for blk in blocks:
            n = blk.n_reads
            lens = np.diff(blk.seq_off()).astype(np.int64)
            rec = np.zeros(n, dtype=_lib.RECORD_DTYPE)
            rec["read"] = np.arange(n)
            rec["sample"] = rng.integers(0, len(tables.specimen_ids), size=n)
            rec["trim_start"] = np.minimum(40, lens)
            rec["trim_end"] = np.maximum(lens - 40, rec["trim_start"])
            rec["pool"], rec["p1"], rec["p2"] = 0, 0, 1
            rec["dist"] = rng.integers(0, 4, size=(n, 4))
            rec["resolution"] = 6
            rec["reverse"] = rng.integers(0, 2, size=n)
            t0 = time.perf_counter()
            wr.write(blk, rec)
            t_write += time.perf_counter() - t0
        t0 = time.perf_counter()
    t_write += time.perf_counter() - t0
    n_rec, n_bytes = n_reads, sum(os.path.getsize(os.path.join(d, f)) for d, _s, fs in os.walk(out) for f in fs)
    print("writer: %.3fs  %.0f MB/s written  %.2f M records/s (each record lands in 2 files)" %
          (t_write, n_bytes / 1e6 / t_write, n_rec / 1e6 / t_write))
    import shutil
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200000)
