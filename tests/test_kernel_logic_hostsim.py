"""CPU-side check of the CUDA kernels' per-thread logic (via tests/hostsim, the same
__host__ __device__ code looped over the thread grid) against the committed golden vectors, which
were produced by the oracle and verified against the unmodified reference (oracle/make_goldens.py)."""
import pytest

import helpers as H
from specimux_b200.demultiplex import process_sequences


def _cases():
    for name in H.golden_names():
        g = H.load_golden(name)
        for run in g["runs"]:
            yield name, run


@pytest.mark.parametrize("name,run_name", list(_cases()))
def test_hostsim_matches_golden(name, run_name):
    g = H.load_golden(name)
    run = g["runs"][run_name]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    args = H.make_args(run["flags"])
    params = H.params_from_run(run, specimens)
    ops, total, matched = process_sequences(H.records(g["reads"]), params, specimens, args, H.prefilter_for(args),
                                            None, 0, _binding=H.hostsim_binding())
    assert (total, matched) == (run["total"], run["matched"])
    H.assert_ops_equal([H.op_to_dict(o) for o in ops], run["ops"], "%s/%s" % (name, run_name))
