"""N>1 path on CPU: world_size-2 gloo processes, each running its contiguous shard through the
kernel simulator; rank 0's ordered gather must equal the single-process golden result."""
import os
import socket

import pytest

import helpers as H
from specimux_b200.sharding import shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 40, 765000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, name, run_name, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from specimux_b200.sharding import process_sequences_sharded
        g = H.load_golden(name)
        run = g["runs"][run_name]
        specimens = H.build_specimens(g["primers"], g["specimens"])
        args = H.make_args(run["flags"])
        params = H.params_from_run(run, specimens)
        ops, total, matched = process_sequences_sharded(H.records(g["reads"]), params, specimens, args,
                                                        H.prefilter_for(args), dist, 0, H.hostsim_binding())
        if rank == 0:
            q.put(([H.op_to_dict(o) for o in ops], total, matched))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,run_name", [("fixture", "default"), ("synth_dense", "derep_none")])
def test_two_rank_gloo_matches_golden(name, run_name):
    import torch.multiprocessing as mp
    H.hostsim_binding()                      # build once in the parent
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, run_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    ops, total, matched = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    run = H.load_golden(name)["runs"][run_name]
    assert (total, matched) == (run["total"], run["matched"])
    H.assert_ops_equal(ops, run["ops"], name)
