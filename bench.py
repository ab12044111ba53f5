#!/usr/bin/env python3
"""bench.py -- reads/s demultiplexed on the BASELINE.json headline workload.

One "step" = one pass of the whole hot path (window staging, primer HW search in both
orientations, barcode SHW search, selection/dereplication) over one batch of synthetic reads of
the named shape.  `value` is measured with the packed reads already resident in HBM (CUDA events
around the kernels); `e2e` goes through the public C-ABI call with HOST buffers (H2D + kernels +
D2H inside the timed region).  Reads are independent, so N GPUs = N ranks each processing its own
batch of the same shape (weak scaling), no collective on the data path; torch.distributed is used
only for the barrier and the max-over-ranks of the timings.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the headline metric is quoted on
    "ont037": dict(reads=765_000, desc="ONT037-shape synthetic: 768 specimens (8x96 13-nt barcodes), ITS1F/ITS4, "
                                       "765k reads ~700 bp, 7% ONT-like errors, mixed orientation"),
    "dense": dict(reads=2_000_000, desc="dense 96x96 dual-index grid (9,216 specimens), ~700 bp, 12% errors"),
    "multipool": dict(reads=5_000_000, desc="multi-pool ITS + RPB2 (IUPAC) + shared primers"),
    "long": dict(reads=20_000_000, desc="long amplicon 2-4 kb, widened search windows"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def dataset(name, n_reads, seed_shift):
    """Synthetic dataset (cached under /tmp as .npz so the two arms of one box share it)."""
    from specimux_b200 import synth
    cache = "/tmp/smx_bench_%s_%d_%d.npz" % (name, n_reads, seed_shift)
    base = synth.CONFIGS[name](n_reads=64, with_quals=False)     # tables only (cheap)
    if os.path.exists(cache):
        z = np.load(cache)
        base.codes, base.offsets = z["codes"], z["offsets"]
        return base
    t0 = time.time()
    fn = synth.CONFIGS[name]
    defaults = {"ont037": 37, "multipool": 3, "long": 4, "dense": 5}
    ds = fn(n_reads=n_reads, seed=defaults[name] + 1000 * seed_shift, with_quals=False)
    # keep the barcode tables of the canonical seed (rank-independent tables, rank-dependent reads)
    if seed_shift:
        ds.primers, ds.specimens = base.primers, base.specimens
    log("generated %s x %d reads in %.1fs" % (name, n_reads, time.time() - t0))
    try:
        np.savez(cache + ".tmp.npz", codes=ds.codes, offsets=ds.offsets)
        os.replace(cache + ".tmp.npz", cache)
    except OSError:
        pass
    return ds


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe).  Reads NVML in-process
    (nvidia_ml_py); spawning `nvidia-smi` every 0.2 s from each of N ranks takes the driver's global lock often
    enough to stretch the host-side CUDA calls of the e2e loop (measured at N = 8: 7.9 ms per step instead of ~2),
    so the command-line tool is only the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._th = index, [], threading.Event(), None
        self._nvml = self._handle = None
        self.source = "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml, self._handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flag = lambda bit: "Active" if r & bit else "Not Active"
        return [str(sm), str(mx), "", hex(r), flag(n.nvmlClocksThrottleReasonHwSlowdown),
                flag(n.nvmlClocksThrottleReasonHwThermalSlowdown), flag(n.nvmlClocksThrottleReasonSwThermalSlowdown),
                flag(n.nvmlClocksThrottleReasonSwPowerCap)]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 8:
                        self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.05 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": sorted(reasons), "samples": len(self.samples), "source": self.source}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def _write_fastq(path, ds, n):
    """FASTQ of the first n reads (ids as SynthDataset.read_id, qualities 'I' when the dataset has none)."""
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    blob = lut[ds.codes[:int(ds.offsets[n])]].tobytes()
    offs = ds.offsets
    with open(path + ".tmp", "wb") as fh:
        for lo in range(0, n, 20000):
            hi = min(n, lo + 20000)
            parts = []
            for i in range(lo, hi):
                seq = blob[int(offs[i]):int(offs[i + 1])]
                parts.append(b"@%s_%08d\n%s\n+\n%s\n" % (ds.name.encode(), i, seq, b"I" * len(seq)))
            fh.write(b"".join(parts))
    os.replace(path + ".tmp", path)


def _workload_files(config, ds, n):
    """primers.fasta / specimens.txt / reads.fastq of the first n reads under /tmp (shared by the two arms)."""
    # tmpfs when the box has one with room (both arms, input and output tree): on the GPU boxes /tmp sits on a virtual
    # ext4 disk whose journal / write-back adds up to +-2 s to a 0.5 s run (profiles/r2_o_file_to_tree.md)
    base = "/tmp"
    try:
        import shutil
        if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (16 << 30):
            base = "/dev/shm"
    except OSError:
        pass
    d = "%s/smx_bench_files_%s_%d" % (base, config, n)
    os.makedirs(d, exist_ok=True)
    files = [os.path.join(d, f) for f in ("primers.fasta", "specimens.txt", "reads.fastq")]
    if not all(os.path.exists(f) for f in files):
        ds.write_tables(files[0], files[1])
        _write_fastq(files[2], ds, n)
    return d, files


def _run_cli(pythonpath, files, out_dir, extra, timeout):
    """One run of `python -m specimux.cli <files> -F -O out ...`; returns (reads, elapsed by the CLI's own
    clock -- the 'Elapsed time' line of orchestration.py:226-227 --, wall seconds of the whole process)."""
    import re
    import shutil
    shutil.rmtree(out_dir, ignore_errors=True)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), PYTHONHASHSEED="0", HOME=os.path.dirname(out_dir))
    cmd = [sys.executable, "-m", "specimux.cli"] + files + ["-F", "-O", out_dir] + extra
    t0 = time.perf_counter()
    # cwd away from the repo root: `python -m` puts the working directory first on sys.path, and the root holds
    # this repo's own `specimux` alias package
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout, cwd=os.path.dirname(out_dir))
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("%s failed: %s" % (" ".join(cmd), r.stderr[-2000:]))
    m1 = re.search(r"Processed ([\d,]+) sequences", r.stderr)
    m2 = re.search(r"Elapsed time: ([\d.]+) seconds", r.stderr)
    if not m1 or not m2:
        raise RuntimeError("no 'Processed' / 'Elapsed time' line in the CLI log: %s" % r.stderr[-2000:])
    return int(m1.group(1).replace(",", "")), float(m2.group(1)), wall


REF_COPY = os.path.join(ROOT, "baseline", "_ref")


def bench_config(workload, reads_per_gpu, search_len, k_index):
    """The `config` object of the JSON line, identical in the two arms for one workload."""
    return {"workload": workload, "description": WORKLOADS[workload]["desc"], "reads_per_gpu": int(reads_per_gpu),
            "search_len": int(search_len), "k_index": int(k_index), "dereplicate": "best", "trim": "barcodes",
            "l2": "GPU arm: L2 flushed (512 MB memset) before every timed step, inputs resident in HBM for `value` and in "
                  "pinned host memory for `e2e`; CPU arm: a FASTQ file in the page cache"}


def run_reference(args):
    """Reference arm: the UNMODIFIED reference CLI (baseline/_ref, a git-ignored copy of /root/reference made by
    baseline/make_ref.py) run as `python -m specimux.cli ... -F -t <all host cores>` over the stand-ins for its
    three uninstallable dependencies (edlib -> oracle/edlib_restated.c, Bio, pybloomfilter), on a FASTQ file of a
    bounded sample of the workload.  Falls back to the oracle port only when the copy is absent."""
    rank, _local, world = dist_env()
    if rank != 0:
        return
    from oracle import cpu_bench
    wl = WORKLOADS[args.config]
    cores = cpu_bench.host_cores()
    have_ref = os.path.isdir(os.path.join(REF_COPY, "src", "specimux"))
    n_runs = args.warmup + args.steps
    budget_s = 150.0                       # the whole arm should end within a few minutes
    ds_small = dataset(args.config, 64, 0)
    from oracle import pipeline as orc
    k_index = orc.setup_params(orc.Tables(ds_small.primers, ds_small.specimens), search_len=ds_small.search_len).max_dist_index
    config = bench_config(args.config, args.reads or wl["reads"], ds_small.search_len, k_index)
    if have_ref:
        pythonpath = [os.path.join(ROOT, "oracle", "standins"), os.path.join(REF_COPY, "src"), ROOT]
        flags = ["-t", str(cores), "--disable-prefilter"]
        if args.ref_sample:
            sample = args.ref_sample
        else:
            # calibration run on 2,000 reads per core, then a sample sized for the time budget (whole 1000-read batches)
            probe = max(4000, 2000 * cores)
            ds = dataset(args.config, probe, 0)
            d, files = _workload_files(args.config, ds, probe)
            _n, el, _w = _run_cli(pythonpath, files, os.path.join(d, "out_ref"), flags, 900)
            rate = probe / max(el, 1e-3)
            sample = int(min(wl["reads"], max(probe, rate * budget_s / max(1, n_runs))) // 1000 * 1000)
        ds = dataset(args.config, sample, 0)
        d, files = _workload_files(args.config, ds, sample)
        times, walls = [], []
        for i in range(n_runs):
            n_done, el, wall = _run_cli(pythonpath, files, os.path.join(d, "out_ref"), flags, 1800)
            assert n_done == sample, (n_done, sample)
            if i >= args.warmup:
                times.append(el)
                walls.append(wall)
        kind = "reference"
        what = ("unmodified reference CLI (baseline/_ref: specimux 0.7.0 `python -m specimux.cli -F -t %d --disable-prefilter`, "
                "FASTQ file of the first %d reads -> output tree, both on tmpfs when /dev/shm has room) over stand-ins for edlib (C restatement), Bio and "
                "pybloomfilter; timed by the CLI's own 'Elapsed time' clock; the prefilter is off because its stand-in is a "
                "Python set whose build takes minutes (result-neutral on A/C/G/T reads, 18 %% faster when cached)" % (cores, sample))
    else:
        sample = args.ref_sample if args.ref_sample else max(4000, 2000 * cores)
        ds = dataset(args.config, sample, 0)
        reads = ds.reads(0, sample)
        times = []
        for i in range(n_runs):
            res = cpu_bench.run(ds.primers, ds.specimens, reads, ds.search_len, processes=cores)
            if i >= args.warmup:
                times.append(res["seconds"])
        walls = times
        kind = "port"
        what = ("first %d reads of the workload per step, oracle port of specimux's process_sequences over a C edlib "
                "restatement, %d worker processes (baseline/_ref absent)" % (sample, cores))
    ms = 1000.0 * sum(times) / len(times)
    value = sample / (ms / 1000.0)
    line = {"impl": "reference", "metric": "reads/sec demuxed (whole box)", "value": value, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind, "sample": what,
                             "wall_s_per_step": sum(walls) / len(walls)},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def file_to_tree(args, ds, n_reads, n_gpus):
    """BASELINE.json's whole-box metric: the workload as a FASTQ file on local disk -> finished output tree through
    this package's own CLI (`python -m specimux.cli ... -F -t <n_gpus>`), timed by the CLI's 'Elapsed time' clock
    exactly as the reference arm is; one warm-up run (page cache), then the median of three runs is reported (a run is
    ~0.5 s on a shared 16-vCPU VM: single runs scatter by +-0.2 s; all three are listed)."""
    d, files = _workload_files(args.config, ds, n_reads)
    out = os.path.join(d, "out_b200")
    runs = [_run_cli([ROOT], files, out, ["-t", str(n_gpus)], 900) for _ in range(4)][1:]
    n_done, el, wall = sorted(runs, key=lambda r: r[1])[1]
    size = os.path.getsize(files[2])
    import shutil
    shutil.rmtree(out, ignore_errors=True)
    return {"value": n_done / el, "unit": "reads/s", "n_gpus": n_gpus, "reads": n_done, "elapsed_s": el, "process_wall_s": wall,
            "elapsed_s_runs": [r[1] for r in runs],
            "fastq_bytes": size, "fastq_gbs": size / el / 1e9,
            "medium": "tmpfs (/dev/shm)" if d.startswith("/dev/shm") else "local disk (/tmp), page cache",
            "api": "python -m specimux.cli primers.fasta specimens.txt reads.fastq -F -O <dir> -t %d (native parallel reader / "
                   "packer, smx_match_batch per byte-range chunk, native tree writer)" % n_gpus}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="ont037", choices=list(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (testing only)")
    ap.add_argument("--ref-sample", type=int, default=0, help="reads per step of the reference arm")
    ap.add_argument("--cpu-sample", type=int, default=12000, help="reads of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-to-tree", action="store_true", help="skip the FASTQ file -> output tree leg")
    ap.add_argument("--split", type=int, default=2, help="concurrent sub-batches of the resident run (1 = one lane)")
    ap.add_argument("--chunk", type=int, default=-1, help="reads per pipeline chunk of smx_match_batch (-1 = library default)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: --warmup %d < 3; timing hygiene asks for >= 3" % args.warmup)
    if args.impl == "reference":
        return run_reference(args)

    rank, local, world = dist_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        # Plumbing only (barrier + max-over-ranks of three timings): the data path has no collective
        # (SURVEY.md 8e), so the CPU-side gloo backend is enough and keeps NCCL banners off stdout.
        dist.init_process_group("gloo")

    from specimux_b200 import _lib
    from specimux_b200.engine import Matcher, PackedBatch
    from specimux_b200.models import MatchParameters
    from specimux_b200.tables import MatchTables
    from specimux_b200.orchestration import thresholds_for
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H

    lib = _lib.load()          # raises if the CUDA library is missing -- there is no CPU path
    if lib.smx_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the matching has no CPU fallback")

    numa = _lib.bind_thread_to_gpu_numa_node(local)      # before any pinned allocation (first touch)
    log("rank %d: GPU %d on NUMA node %s" % (rank, local, "%d (%d cpus)" % numa if numa else "unknown"))
    wl = WORKLOADS[args.config]
    n_reads = args.reads or wl["reads"]
    ds = dataset(args.config, n_reads, rank)
    specimens = H.build_specimens(ds.primers, ds.specimens)
    k_idx, k_primers = thresholds_for(specimens, device=local)
    params = MatchParameters(k_primers, k_idx, ds.search_len, True)
    tables = MatchTables(specimens, params, trim="barcodes", dereplicate="best", prefilter=True)
    matcher = Matcher(tables, device=local)
    matcher.set_resident_split(args.split)

    t0 = time.time()
    ascii_blob = np.frombuffer(b"ACGT", dtype=np.uint8)[ds.codes].tobytes()
    batch = PackedBatch.from_blob(ascii_blob, ds.offsets.astype(np.uint64), clip=ds.search_len)
    del ascii_blob
    log("rank %d: packed %d reads (%.1f MB) in %.1fs" % (rank, n_reads, batch.h2d_bytes / 1e6, time.time() - t0))

    import ctypes
    import hashlib
    peaks = (ctypes.c_double * 3)()
    _lib.check(lib.smx_int_alu_peak(local, peaks))
    int_peak = max(peaks)
    copy_peak = (ctypes.c_double * 3)()
    _lib.check(lib.smx_copy_peak(local, 32 << 20, copy_peak))

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing ------------------------------------------------------------
    # (a) the measured configuration: the resident batch runs as concurrent sub-batches (default 2)
    matcher.upload(batch)
    for _ in range(args.warmup):
        matcher.run_resident()
    barrier()
    step_ms = []
    clocks = ClockSampler(local)        # samples across all three timed loops (resident, attribution, e2e)
    clocks.__enter__()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        matcher.flush_l2()                           # evict L2 between timed iterations (not timed)
        matcher.run_resident()                       # synchronises the streams internally
        step_ms.append(matcher.last_timing()[0])
    wall_ms = (time.perf_counter() - t_wall) * 1000.0 / args.steps     # includes the L2 flushes
    barrier()
    launches = matcher.last_launch_count()
    deferred = matcher.last_deferred()
    cells, wcols = matcher.last_work()
    useful = matcher.last_useful_cells()
    res = matcher.download()
    ms = float(np.mean(step_ms))
    # (b) attribution pass: the same batch on ONE lane, kernels back to back, CUDA events around each
    # kernel (with sub-batches the kernels of different lanes overlap and cannot be timed one by one)
    matcher.set_resident_split(1)
    matcher.upload(batch)
    for _ in range(min(args.warmup, 2)):
        matcher.run_resident()
    serial_ms, stage_ms, kernel_ms = [], [], []
    for _ in range(args.steps):
        matcher.flush_l2()
        matcher.run_resident()
        tot, st = matcher.last_timing()
        serial_ms.append(tot)
        stage_ms.append(st)
        kernel_ms.append(matcher.last_kernel_times())
    res_serial = matcher.download()
    if res_serial.records.tobytes() != res.records.tobytes() or not np.array_equal(res_serial.rec_offset, res.rec_offset):
        raise SystemExit("bench.py: split and unsplit resident runs disagree")
    matcher.set_resident_split(args.split)
    st = np.mean(np.array(stage_ms), axis=0)
    ms_serial = float(np.mean(serial_ms))

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    # headline form: fixed-stride 2-bit reads + 16-bit lengths in, 16-byte smx_record16 records out (everything the
    # per-specimen files need); the full 64-byte records (with the four location pairs) are timed beside it
    if args.chunk >= 0:
        matcher.set_pipeline_chunk(args.chunk)

    def time_e2e(**kw):
        for _ in range(min(args.warmup, 2)):
            r_ = matcher.match(batch, reuse=True, **kw)
        barrier()
        t0_ = time.perf_counter()
        for _ in range(args.steps):
            r_ = matcher.match(batch, reuse=True, **kw)
        dt = (time.perf_counter() - t0_) * 1000.0 / args.steps
        return dt, r_.records.nbytes + (r_.rec_offset.nbytes if r_.rec_offset is not None else 0), r_

    e2e_ms, d2h, r_wire = time_e2e(compact="wire")
    matcher_chunks = matcher.last_chunk_count()
    e2e_full_ms, d2h_full, r_full = time_e2e()
    if len(r_wire.records) != len(r_full.records) or not np.array_equal(r_wire.records["sample"], r_full.records["sample"]):
        raise SystemExit("bench.py: wire-form and full records disagree")
    # keep the GPU under the same load until the clock sampler has a few samples (the timed loops above take
    # milliseconds); these extra runs are not part of any figure
    t_hold = time.perf_counter()
    while len(clocks.samples) < 5 and time.perf_counter() - t_hold < 3.0:
        matcher.match(batch, reuse=True, compact="wire")
    clocks.__exit__(None, None, None)
    barrier()

    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_ms, wall_ms, e2e_full_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, wall_ms, e2e_full_ms = [float(x) for x in t.tolist()]

    matcher.close()
    # ---- whole box: FASTQ file on local disk -> output tree through the CLI, all N GPUs (rank 0 runs it) ----
    f2t = None
    if rank == 0 and not args.no_file_to_tree:
        try:
            f2t = file_to_tree(args, ds, n_reads, world)
        except Exception as e:                      # reported, never fatal for the kernel figures
            f2t = {"error": str(e)[-500:]}
    barrier()

    if rank == 0:
        total_reads = n_reads * world
        value = total_reads / (ms / 1000.0)
        alg_ops = 15.0 * (wcols[0] + wcols[1])
        gcups = (cells[0] + cells[1]) / (ms / 1000.0) / 1e9
        # per-kernel durations (CUDA events on the launching stream, mean over the timed steps)
        kt = {k: float(np.mean([d[k] for d in kernel_ms])) for k in kernel_ms[0]}
        # The two DP kernels: stage 1 forward pass (primer HW, bit-sliced across reads) and stage 2 (barcode SHW,
        # bit-sliced across barcodes).  Useful work = bit-sliced DP cells actually evaluated (counted on the device)
        # x 5 three-input logic ops, the cost of one cell of either automaton for its 32 problems; the rate is set
        # against the integer-ALU peak measured in this run.  The dominant kernel by time carries the roofline.
        k_useful = {"primer_sliced": 5.0 * useful[0], "barcode_tasks": 5.0 * useful[1]}
        k_model = {"primer_sliced": 15.0 * wcols[0], "barcode_tasks": 15.0 * wcols[1]}
        dom_name = max(k_useful, key=lambda k: kt[k])
        dp = {}
        for k, ops in k_useful.items():
            a_tops = ops / (kt[k] / 1000.0) / 1e12 if kt[k] > 0 else 0.0
            dp[k] = {"ms": kt[k], "achieved": a_tops, "frac": a_tops / int_peak if int_peak else None,
                     "survey_8d_model_tops": k_model[k] / (kt[k] / 1000.0) / 1e12 if kt[k] > 0 else 0.0}
        achieved = dp[dom_name]["achieved"]
        # ncu figures (executed ALU-pipe utilisation, DRAM bytes) only when the committed capture is of THIS build
        build_id = _lib.source_id()          # identity of the library's sources (a rebuild of the same code keeps it)
        ncu_static = {}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_latest.json")) as fh:
                ncu_static = json.load(fh)
        except OSError:
            pass
        same_build = ncu_static.get("build_id") == build_id
        ncu_dom = ncu_static.get("kernels", {}).get(dom_name, {}) if same_build else {}
        hbm_bytes = batch.h2d_bytes + res.records.nbytes
        # e2e against its own roofline: the slower of the two copy directions at the measured pinned-copy peak, or
        # the kernels (device-resident step time), whichever is larger
        floor_copy = max(batch.h2d_bytes / (copy_peak[0] * 1e9), d2h / (copy_peak[1] * 1e9)) * 1000.0
        floor_ms = max(floor_copy, ms)
        line = {
            "metric": "reads/sec demuxed (whole box)", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            # the same object in both arms (run_reference builds it through the same helper)
            "config": bench_config(args.config, n_reads, ds.search_len, k_idx),
            "l2_note": "%.0f MB packed reads resident, L2 flushed by a 512 MB memset before every timed step"
                       % (batch.h2d_bytes / 1e6),
            "gcups": gcups, "cells_per_read": (cells[0] + cells[1]) / n_reads,
            "stage_ms": {"stage_windows": st[0], "primer_search": st[1], "barcode_search": st[2], "select": st[3]},
            "wall_ms_per_step": wall_ms,
            "resident_sub_batches": args.split,
            "serial_ms_per_step": ms_serial,
            "kernel_ms": kt,
            "kernel_ms_note": "one lane, kernels back to back (attribution pass, same batch, same process); value / "
                              "ms_per_step is the same work as %d concurrent sub-batches" % args.split,
            "roofline": {"bound": "int_alu", "kernel": dom_name, "achieved": achieved, "peak": int_peak,
                         "unit": "Tops/s", "frac": achieved / int_peak if int_peak else None,
                         "whole_step_frac": (5.0 * (useful[0] + useful[1]) / (ms / 1000.0) / 1e12) / int_peak if int_peak else None,
                         "useful_ops": "bit-sliced DP cells evaluated (counted on the device: smx_last_useful_cells) x 5 "
                                       "three-input logic ops per cell of 32 problems",
                         "peak_source": "measured in this run by smx_int_alu_peak (LOP3 %.2f / IADD3 %.2f / mix %.2f Tops/s); "
                                        "MEASURED_PEAKS.json has no integer figure" % (peaks[0], peaks[1], peaks[2]),
                         "dp_kernels": dp,
                         "survey_8d_model": {"ops": "15 int ops x 32-bit word-columns of the one-pattern-per-word model "
                                                    "(SURVEY.md 8d), counted on the device",
                                             "whole_step_tops": alg_ops / (ms / 1000.0) / 1e12,
                                             "note": "kept for continuity; it exceeds the peak because the kernels hold 32 "
                                                     "problems per word and evaluate only the band"},
                         "alu_pipe_pct_ncu": ncu_dom.get("alu_pipe_pct"),
                         "whole_step_alu_pipe_pct_ncu": ncu_static.get("whole_step_alu_pipe_pct") if same_build else None,
                         "traffic": ncu_dom.get("dram_bytes"),
                         "ncu_capture": ncu_static.get("capture") if same_build else
                                        "none for this build (%s); see profiles/ for the captures of committed builds" % build_id,
                         "build_id": build_id,
                         "hbm_sanity_gbs": hbm_bytes / (ms / 1000.0) / 1e9},
            "e2e": {"value": total_reads / (e2e_ms / 1000.0), "unit": "reads/s",
                    "h2d_bytes_per_step": int(batch.h2d_bytes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "chunks": matcher_chunks,
                    "api": "smx_match_batch (pinned host buffers: fixed-stride 2-bit reads clipped to head/tail search_len "
                           "bases + 16-bit lengths in, 16-byte smx_record16 records out; H2D + kernels + D2H, chunks "
                           "pipelined over three streams)",
                    "roofline": {"h2d_peak_gbs": copy_peak[0], "d2h_peak_gbs": copy_peak[1], "both_peak_gbs": copy_peak[2],
                                 "copy_floor_ms": floor_copy, "kernel_floor_ms": ms, "frac": floor_ms / e2e_ms,
                                 "achieved_gbs": (batch.h2d_bytes + d2h) / (e2e_ms / 1000.0) / 1e9,
                                 "note": "floor = max(slower copy direction at the measured pinned-copy peak, device-resident "
                                         "step time); frac = floor / measured"}},
            "e2e_full_records": {"value": total_reads / (e2e_full_ms / 1000.0), "unit": "reads/s",
                                 "ms_per_step": e2e_full_ms, "h2d_bytes_per_step": int(batch.h2d_bytes),
                                 "d2h_bytes_per_step": int(d2h_full),
                                 "note": "same call returning the 64-byte smx_record records (with the four location pairs "
                                         "the reference keeps for --color / traces) and rec_offset; extra figure"},
            "file_to_tree": f2t,
            "gpu_launches": int(launches * args.steps),
            "host_numa_node": numa[0] if numa else None,
            "clocks": clocks.summary(),
            "records_per_step": int(len(res.records)), "matched_reads": int(res.n_matched),
            "reads_on_general_selection_path": int(deferred),
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import cpu_bench
            cores = cpu_bench.host_cores()
            sample = min(max(args.cpu_sample, 1500 * cores), n_reads)
            reads = ds.reads(0, sample)
            cb = cpu_bench.run(ds.primers, ds.specimens, reads, ds.search_len, processes=cores)
            line["cpu_baseline"] = {"value": cb["reads_per_s"], "unit": "reads/s", "cores": cores, "kind": "port",
                                    "sample": "first %d reads of the workload, %.1f s on %d host cores (oracle port of "
                                              "specimux process_sequences over the C edlib restatement; the unmodified "
                                              "reference CLI is the --impl reference arm)" % (sample, cb["seconds"], cores)}
        print(json.dumps(line, default=float), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
