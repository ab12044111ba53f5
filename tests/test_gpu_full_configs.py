"""Full-size oracle parity on the GPU box: every record of the CUDA path compared with
the multi-process oracle (oracle.pipeline over the C aligner, forked over all host cores) -- all 765 k reads of
BASELINE config 2, 200 k reads each of configs 3 (both dereplicate modes), 4 (-l 200 and -l 500) and 5 (both
modes), plus batches with non-ACGT bases sprinkled in and a "tie storm" that drives the rare device paths
(hit-list capacity re-runs, k_select_big, record-pool growth) against the oracle instead of the simulator.
Follows the reference's process_sequences (demultiplex.py:108-212) record for record."""
import json
import sys

import pytest

import fullcfg

pytestmark = pytest.mark.gpu


def _report(fig):
    print("full-config parity:", json.dumps(fig), file=sys.stderr)


def test_config2_ont037_all_765k_reads():
    fig = fullcfg.compare("ont037", 765_000)
    _report(fig)
    assert fig["matched"] > 0.9 * fig["reads"] and fig["multi_record_reads"] > 0


@pytest.mark.parametrize("derep", ["best", "none"])
def test_config3_multipool_200k(derep):
    fig = fullcfg.compare("multipool", 200_000, {"dereplicate": derep})
    _report(fig)
    assert fig["matched"] > 0.85 * fig["reads"]


@pytest.mark.parametrize("search_len", [200, 500])
def test_config4_long_amplicon_200k(search_len):
    fig = fullcfg.compare("long", 200_000, search_len=search_len)
    _report(fig)
    assert fig["matched"] > 0.85 * fig["reads"]


@pytest.mark.parametrize("derep", ["best", "none"])
def test_config5_dense_200k(derep):
    fig = fullcfg.compare("dense", 200_000, {"dereplicate": derep})
    _report(fig)
    assert fig["matched"] > 0.75 * fig["reads"]


@pytest.mark.parametrize("cfg,n,flags", [("ont037", 100_000, {}), ("multipool", 60_000, {"trim": "tails"}),
                                         ("dense", 60_000, {"dereplicate": "none", "trim": "primers"})])
def test_non_acgt_sprinkled(cfg, n, flags):
    """1 in 2,000 bases replaced by N / IUPAC / lower case: those reads take the exact 4-bit side stream."""
    fig = fullcfg.compare(cfg, n, flags, sprinkle=2000)
    _report(fig)
    assert fig["flagged_reads"] > 0.03 * n


@pytest.mark.parametrize("derep,tiny", [("best", False), ("none", False), ("best", True)])
def test_tie_storm_rare_device_paths(derep, tiny, monkeypatch):
    """96 x 96 grid of barcodes 1-2 substitutions away from six seeds, searched at k = 3 (synth.tie_storm): every
    flank is within k of many barcodes, so hit sub-lists overflow (4 -> 16 -> 32 re-runs), reads emit many records (record pool growth) and
    overflow the 16 thread-local dereplication groups (k_select_big).  With `tiny` every growable buffer also
    starts far too small (SMX_TEST_TINY_CAPS)."""
    if tiny:
        monkeypatch.setenv("SMX_TEST_TINY_CAPS", "1")
    fig = fullcfg.compare("tiestorm", 40_000, {"dereplicate": derep, "trim": "tails"}, index_edit_distance=3,
                          sprinkle=3000)
    _report(fig)
    if derep == "best":
        assert fig["multi_record_reads"] > 0.05 * fig["reads"]


@pytest.mark.parametrize("cfg,n,flags", [("ont037", 60_000, {"trim": "primers"}), ("multipool", 40_000, {"trim": "tails"})])
def test_switched_off_kernel_forms_still_match_oracle(cfg, n, flags, monkeypatch):
    """The bit-sliced start recovery (k_primer_start_sliced) and the four-entries-to-a-word barcode task (k_barcode_quad)
    are off by default (slower on the box, profiles/r2_k_ab.md); they stay in the library behind their switches and
    have to stay exact."""
    monkeypatch.setenv("SMX_START_SLICED", "1")
    monkeypatch.setenv("SMX_BARCODE_QUAD", "1")
    fig = fullcfg.compare(cfg, n, flags, sprinkle=5000)
    _report(fig)
    assert fig["matched"] > 0.8 * fig["reads"]
