"""CPU tier of the full-config parity machinery (tests/fullcfg.py): the same record-for-record comparison
with the multi-process oracle, at small sizes, through the kernel simulator."""
import pytest

import fullcfg
import helpers as H


@pytest.mark.parametrize("cfg,n,flags,kw", [
    ("ont037", 1500, {}, {}),
    ("multipool", 800, {"dereplicate": "none"}, {"sprinkle": 1500}),
    ("long", 400, {}, {"search_len": 500}),
    ("tiestorm", 500, {"trim": "tails"}, {"index_edit_distance": 3, "sprinkle": 3000}),
])
def test_records_equal_multiprocess_oracle(cfg, n, flags, kw):
    fig = fullcfg.compare(cfg, n, flags, binding=H.hostsim_binding(), processes=2, **kw)
    assert fig["records"] >= n


def test_quad_barcode_form_equals_oracle_on_the_simulator(monkeypatch):
    """SMX_BARCODE_QUAD=1: narrow barcode words take four work entries to a thread (off by default)."""
    monkeypatch.setenv("SMX_BARCODE_QUAD", "1")
    fig = fullcfg.compare("ont037", 1200, {"trim": "primers"}, binding=H.hostsim_binding(), processes=2, sprinkle=4000)
    assert fig["records"] >= 1200
