"""Record-for-record comparison of the product's smx_record stream with the multi-process oracle on a
whole synthetic config (used by tests/test_gpu_full_configs.py on the GPU and, at a small size through
the kernel simulator, by the CPU tier)."""
import numpy as np

import helpers as H
from specimux_b200 import synth
from specimux_b200.engine import Matcher, PackedBatch
from specimux_b200.models import MatchParameters
from specimux_b200.tables import MatchTables

FIELDS = ("read", "sample", "trim_start", "trim_end", "p1_loc", "p2_loc", "b1_loc", "b2_loc", "pool", "p1", "p2",
          "dist", "resolution", "reverse", "trim_empty")


def sprinkle_non_acgt(ascii_codes, one_in, seed):
    """Replaces ~1 in `one_in` bases by N / IUPAC codes / lower-case letters (reads then take the exact 4-bit
    side stream)."""
    rng = np.random.default_rng(seed)
    out = ascii_codes.copy()
    hits = rng.choice(out.shape[0], size=max(1, out.shape[0] // one_in), replace=False)
    out[hits] = np.frombuffer(b"NNRYKMSWBDHVacgtn", dtype=np.uint8)[rng.integers(0, 17, size=hits.shape[0])]
    return out


_ds_cache = {}


def dataset(cfg, n, seed=None, search_len=None, sprinkle=0):
    kw = {"with_quals": False}
    if seed is not None:
        kw["seed"] = seed
    key = (cfg, n, seed)
    if key not in _ds_cache:
        _ds_cache.clear()                        # one full-size dataset at a time
        _ds_cache[key] = synth.CONFIGS[cfg](n_reads=n, **kw)
    ds = _ds_cache[key]
    ascii_codes = np.frombuffer(b"ACGT", dtype=np.uint8)[ds.codes]
    if sprinkle:
        ascii_codes = sprinkle_non_acgt(ascii_codes, sprinkle, 1000 + n)
    return ds, ascii_codes.tobytes(), ds.offsets.astype(np.uint64)


def compare(cfg, n, flags=None, seed=None, search_len=None, sprinkle=0, binding=None, processes=None, matcher_hook=None,
            index_edit_distance=-1):
    """Runs the oracle (forked workers) and the product on the same reads; asserts identical records.
    Returns a dict of figures for the test log."""
    from oracle import cpu_bench
    from oracle import pipeline as orc
    flags = dict(flags or {})
    ds, blob, offs = dataset(cfg, n, seed, search_len, sprinkle)
    L = search_len or ds.search_len
    specimens = H.build_specimens(ds.primers, ds.specimens)
    otables = orc.Tables(ds.primers, ds.specimens)
    oparams = orc.setup_params(otables, index_edit_distance=index_edit_distance, search_len=L)
    params = MatchParameters(dict(oparams.max_dist_primers), oparams.max_dist_index, L,
                             not flags.get("disable_preorient", False))
    mt = MatchTables(specimens, params, trim=flags.get("trim", "barcodes"), dereplicate=flags.get("dereplicate", "best"),
                     prefilter=not flags.get("disable_prefilter", False))
    maps = ({sid: i for i, sid in enumerate(mt.specimen_ids)}, {b: i for i, b in enumerate(mt.b1)},
            {b: i for i, b in enumerate(mt.b2)}, {p: i for i, p in enumerate(mt.pools)},
            {p: i for i, p in enumerate(mt.primer_names)})
    # Q7: several primer names with one sequence collapse onto the first-registered canonical primer, whose
    # name is what both sides print; the map above is name -> canonical index of that first name.
    param_kw = dict(index_edit_distance=index_edit_distance, search_len=L, preorient=not flags.get("disable_preorient", False),
                    prefilter=not flags.get("disable_prefilter", False), trim=flags.get("trim", "barcodes"),
                    dereplicate=flags.get("dereplicate", "best"))
    want, want_matched, oracle_s = cpu_bench.run_records(ds.primers, ds.specimens, blob, offs, maps, param_kw,
                                                        processes=processes)
    batch = PackedBatch.from_blob(blob, offs, clip=L)
    m = Matcher(mt, binding=binding)
    try:
        if matcher_hook:
            matcher_hook(m)
        got = m.match(batch)
        deferred = m.last_deferred() if binding is None else None
    finally:
        m.close()
    rec = got.records
    assert got.n_matched == want_matched, (cfg, got.n_matched, want_matched)
    assert len(rec) == len(want), (cfg, len(rec), len(want))
    for f in FIELDS:
        a, b = rec[f], want[f]
        if not np.array_equal(a, b):
            bad = np.nonzero((a != b).reshape(len(rec), -1).any(axis=1))[0]
            i = int(bad[0])
            r = int(rec["read"][i])
            raise AssertionError("%s: field %s differs in %d of %d records; first at record %d (read %d, %s):\n got    %r\n oracle %r"
                                 % (cfg, f, len(bad), len(rec), i, r, blob[int(offs[r]):int(offs[r + 1])][:60], rec[i], want[i]))
    counts = np.diff(got.rec_offset)
    assert int(counts.sum()) == len(rec)
    return dict(config=cfg, reads=n, records=len(rec), matched=int(got.n_matched), flagged_reads=batch.n_flagged,
                multi_record_reads=int((counts > 1).sum()), deferred=deferred, oracle_seconds=round(oracle_s, 1))
