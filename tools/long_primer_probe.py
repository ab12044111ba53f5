#!/usr/bin/env python3
"""Probe of the warp-cooperative multi-word primer kernel (k_primer_long): a 150-nt forward primer, ITS4 as the reverse
primer, 8 x 96 barcodes, 200 k reads, search window 256.  Prints the CUDA-event time of the stage-1 kernels and checks a
slice against the kernel simulator.  Run it under ncu (-k regex:k_primer_long) for the counters."""
import os
import sys

import numpy as np

os.environ.setdefault("SMX_TEST_SEAM", "1")      # the probe cross-checks 3,000 reads against the kernel simulator

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(n_reads=200_000):
    import helpers as H
    from specimux_b200 import synth
    from specimux_b200.engine import Matcher, PackedBatch
    from specimux_b200.models import MatchParameters
    from specimux_b200.orchestration import thresholds_for
    from specimux_b200.tables import MatchTables
    rng = np.random.default_rng(150)
    long_fwd = "".join(rng.choice(list("ACGT"), size=150))
    fb, rb = synth.make_barcodes(8, 96, 13, 6, 37)
    primers = [("LONG150", long_fwd, "forward", ["P"]), ("ITS4", synth.ITS4, "reverse", ["P"])]
    rows, struct = synth._grid_specimens("LP", "P", fb, rb, "LONG150", "ITS4", long_fwd, synth.ITS4)
    ds = synth._build("longprimer", primers, rows, struct, n_reads, 7, synth._normal_insert(640, 60, 50), 0.07, 256, False)
    specimens = H.build_specimens(ds.primers, ds.specimens)
    k_idx, k_primers = thresholds_for(specimens)
    params = MatchParameters(k_primers, k_idx, ds.search_len, True)
    mt = MatchTables(specimens, params)
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[ds.codes].tobytes()
    offs = ds.offsets.astype(np.uint64)
    batch = PackedBatch.from_blob(blob, offs, clip=ds.search_len)
    with Matcher(mt) as m:
        m.set_resident_split(1)
        m.upload(batch)
        for _ in range(3):
            m.run_resident()
        kt = m.last_kernel_times()
        res = m.download()
        cells, wcols = m.last_work()
        print("k (primer thresholds):", k_primers.values(), "k_idx", k_idx)
        print("matched %d of %d reads; stage-1 kernels (k_primer_search<u64> for ITS4 + k_primer_long<5> for the 150-mer): %.1f us"
              % (res.n_matched, n_reads, 1000 * kt.get("primer_finish_start", kt.get("primer_finish", 0.0))))
        print("kernel_ms", {k: round(v, 4) for k, v in kt.items()})
        hw_cells_long = 2 * 150 * 256 * n_reads
        print("HW cells of the long primer: %.3g -> %.0f GCUPS on that kernel pair" % (hw_cells_long, hw_cells_long / (kt.get("primer_finish_start", kt.get("primer_finish", 0.0)) * 1e-3) / 1e9))
        n_sim = 3000
        sub = PackedBatch.from_blob(blob[:int(offs[n_sim])], offs[:n_sim + 1], clip=ds.search_len)
        gpu = m.match(sub)
    sim = Matcher(mt, binding=H.hostsim_binding()).match(sub)
    assert gpu.records.tobytes() == sim.records.tobytes()
    print("first %d reads identical to the kernel simulator" % n_sim)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200_000)
