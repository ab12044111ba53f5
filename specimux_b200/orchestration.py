"""Orchestration around the GPU matching path (mirror of the reference's orchestration.py seams).

setup_match_parameters (reference orchestration.py:548-641) keeps its semantics; its O(B^2)
edlib NW loop runs as one batched GPU kernel (smx_pairwise_nw) and no Bloom filter is ever built
(the prefilter's observable behaviour is emulated by the barcode kernel).
"""
import ctypes as C
import itertools
import logging
import math
import os
import timeit
from datetime import datetime
from typing import Dict, List, Tuple

import numpy as np

from . import _lib
from .constants import Primer
from .databases import BloomEmulationPrefilter, PassthroughPrefilter
from .demultiplex import process_sequences
from .io_utils import (OutputManager, cleanup_empty_directories, open_sequence_file, output_write_operation,
                       read_primers_file, read_specimen_file)
from .models import MatchParameters
from .seqio import reverse_complement


def pairwise_nw(seqs: List[str], device: int = 0) -> np.ndarray:
    """All-pairs global edit distance on the GPU (replaces orchestration.py:549-555)."""
    lib = _lib.load()
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    out = np.zeros((n, n), dtype=np.int32)
    _lib.check(lib.smx_pairwise_nw(device, "".join(seqs).encode("ascii"), _lib.ptr(off, _lib.u32p), n,
                                   _lib.ptr(out, _lib.i32p)))
    return out


def _min_pairwise(seqs: List[str], device: int):
    if len(seqs) <= 1:
        return None
    d = pairwise_nw(seqs, device)
    iu = np.triu_indices(len(seqs), k=1)
    return int(d[iu].min())


def _bp_adjusted_length(primer: str) -> float:
    """reference: orchestration.py:564-570."""
    score = 0
    for b in primer:
        if b in "ACGT":
            score += 3
        elif b in "KMRSWY":
            score += 2
        elif b in "BDHV":
            score += 1
    return score / 3.0


def thresholds_for(specimens, index_edit_distance: int = -1, primer_edit_distance: int = -1,
                   device: int = 0, diagnostics=False) -> Tuple[int, Dict[str, int]]:
    """(k_idx, {primer sequence: k}) -- reference orchestration.py:572-614."""
    b1s, b2s = {}, {}
    for p in specimens.get_primers(Primer.FWD):
        b1s.update(p.barcodes)
    for p in specimens.get_primers(Primer.REV):
        b2s.update(p.barcodes)
    combined = list(b1s) + [reverse_complement(b) for b in b2s]
    if index_edit_distance != -1:
        k_idx = index_edit_distance
        if diagnostics:
            for desc, seqs in (("Forward Barcodes", list(b1s)), ("Reverse Barcodes", list(b2s)),
                               ("Forward Barcodes + Reverse Complement of Reverse Barcodes", combined)):
                m = _min_pairwise(seqs, device)
                if m is not None:
                    logging.info(f"Minimum edit distance is {m} for {desc}")
    else:
        min_bc = _min_pairwise(combined, device)
        if diagnostics and min_bc is not None:
            logging.info(f"Minimum edit distance is {min_bc} for Forward Barcodes + Reverse Complement of Reverse Barcodes")
        k_idx = math.ceil(min_bc / 2.0)      # raises TypeError on a single barcode, like the reference
    thr = {}
    for p in specimens.get_primers(Primer.FWD) + specimens.get_primers(Primer.REV):
        thr[p.primer] = primer_edit_distance if primer_edit_distance != -1 else int(_bp_adjusted_length(p.primer) / 3)
    return k_idx, thr


def setup_match_parameters(args, specimens, device: int = 0) -> MatchParameters:
    """reference: orchestration.py:548-641."""
    k_idx, thr = thresholds_for(specimens, args.index_edit_distance, args.primer_edit_distance, device,
                                getattr(args, "diagnostics", None))
    logging.info(f"Using Edit Distance Thresholds {k_idx} for barcode indexes")
    for p, pt in thr.items():
        logging.info(f"Using Edit Distance Threshold {pt} for primer {p}")
    logging.info(f"Using dereplication strategy: {args.dereplicate}")
    if args.disable_preorient:
        logging.info("Sequence pre-orientation disabled, may run slower")
    if not args.disable_prefilter:
        if specimens.b_length() > 13:
            logging.warning("Barcode prefilter not tested for barcodes longer than 13 nt.  You may need to use --disable-prefilter")
        if k_idx > 3:
            logging.warning("Barcode prefilter not tested for edit distance greater than 3.  You may need to use --disable-prefilter")
        logging.info("Using Bloom Filter optimization for barcode matching")
    else:
        logging.info("Barcode prefiltering disabled, may run slower")
    return MatchParameters(thr, k_idx, args.search_len, not args.disable_preorient)


# ---------------------------------------------------------------------------------------------
# Output tree (reference orchestration.py:240-372) and the drivers (reference :153-237, :458-545)

def _write_primers(path_dir: str, fwd_primers, rev_primers):
    with open(os.path.join(path_dir, "primers.fasta"), "w") as f:
        for p in fwd_primers:
            f.write(f">{p.name} position=forward pool={','.join(p.pools)}\n{p.primer}\n")
        for p in rev_primers:
            f.write(f">{p.name} position=reverse pool={','.join(p.pools)}\n{p.primer}\n")
    with open(os.path.join(path_dir, "primers.txt"), "w") as f:
        for p in list(fwd_primers) + list(rev_primers):
            f.write(f">{p.name}\n{p.primer}\n")


def create_output_files(args, specimens):
    """reference: orchestration.py:314-372 -- directory tree + primers.fasta / primers.txt files."""
    if not args.output_to_files:
        return
    out = args.output_dir
    os.makedirs(out, exist_ok=True)
    kinds = ["full", "partial", "unknown"]
    for kind in kinds:
        os.makedirs(os.path.join(out, kind), exist_ok=True)
    for kind in ("partial", "unknown"):
        os.makedirs(os.path.join(out, kind, "unknown", "unknown-unknown"), exist_ok=True)
    registry = specimens._primer_registry
    for pool in registry.get_pools():
        fwd, rev = registry.get_pool_primers(pool, Primer.FWD), registry.get_pool_primers(pool, Primer.REV)
        for kind in kinds:
            os.makedirs(os.path.join(out, kind, pool), exist_ok=True)
        _write_primers(os.path.join(out, "full", pool), fwd, rev)
        for f in fwd:
            for r in rev:
                for kind in kinds:
                    os.makedirs(os.path.join(out, kind, pool, f"{f.name}-{r.name}"), exist_ok=True)
                _write_primers(os.path.join(out, "full", pool, f"{f.name}-{r.name}"), [f], [r])
            for kind in ("partial", "unknown"):
                os.makedirs(os.path.join(out, kind, pool, f"{f.name}-unknown"), exist_ok=True)
        for r in rev:
            for kind in ("partial", "unknown"):
                os.makedirs(os.path.join(out, kind, pool, f"unknown-{r.name}"), exist_ok=True)


def iter_batches(seq_records, batch_size: int, max_seqs: int, all_seqs: bool):
    """reference: orchestration.py:447-456."""
    num = 0
    while all_seqs or num < max_seqs:
        to_read = batch_size if all_seqs else min(batch_size, max_seqs - num)
        batch = list(itertools.islice(seq_records, to_read))
        if not batch:
            break
        yield batch
        num += len(batch)


# reads per GPU call (the reference hands 1000-read batches to CPU workers); SMX_GPU_BATCH_READS overrides
GPU_BATCH_READS = max(128, int(os.environ.get("SMX_GPU_BATCH_READS", "65536")))
TRACE_BATCH_READS = 4096      # with -d the per-search detail arrays come back too (one entry per barcode and read)


def _visible_gpus() -> int:
    return int(_lib.load().smx_device_count())


def _run(args, multi: bool):
    primer_registry = read_primers_file(args.primer_file)
    specimens = read_specimen_file(args.specimen_file, primer_registry)
    specimens.validate()
    parameters = setup_match_parameters(args, specimens)
    n_visible = _visible_gpus()
    if n_visible < 1:
        raise RuntimeError("specimux_b200: no CUDA device visible; the matching has no CPU fallback")
    n_gpus = n_visible if (multi and args.threads <= 0) else max(1, min(n_visible, args.threads if multi else 1))
    prefilter = PassthroughPrefilter() if args.disable_prefilter else BloomEmulationPrefilter()
    if _native_eligible(args):
        create_output_files(args, specimens)
        start = timeit.default_timer()
        n_gpus = _gpus_worth_starting(args, n_gpus)
        logging.info(f"Will run on {n_gpus} GPU(s), {GPU_BATCH_READS} reads per batch, native reader/writer")
        total, matched = _run_native(args, specimens, parameters, n_gpus, prefilter)
        if total > 0:
            logging.info(f"Processed {total:,} sequences, match rate: {matched / total:.1%}")
        logging.info(f"Elapsed time: {timeit.default_timer() - start:.2f} seconds")
        if args.output_to_files:
            cleanup_empty_directories(args.output_dir)
        return
    seq_records = open_sequence_file(args.sequence_file, args)
    create_output_files(args, specimens)
    start = timeit.default_timer()
    if getattr(args, "start_seq", 1) > 1:
        for _ in itertools.islice(seq_records, args.start_seq - 1):
            pass
    all_seqs = args.num_seqs < 0
    logging.info(f"Will run on {n_gpus} GPU(s), {GPU_BATCH_READS} reads per batch")
    stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
    trace_logger = None
    if args.diagnostics and args.output_to_files:
        from .trace import TraceLogger
        trace_logger = TraceLogger(True, args.diagnostics, args.output_dir, "main", stamp)
        if n_gpus > 1:
            # the reference writes one trace file per worker process; here one ordered stream is kept instead
            logging.warning(f"Trace output (-d) runs on one GPU (of {n_gpus}) so that events stay in input order")
            n_gpus = 1

    from concurrent.futures import ThreadPoolExecutor
    from collections import deque
    total = matched = 0
    offset = 0
    output_manager = OutputManager(args.output_dir, args.output_file_prefix, args.isfastq) if args.output_to_files else None
    if output_manager:
        output_manager.__enter__()
    # One single-thread executor PER GPU: a device's Matcher (and the pinned result pool its records are views
    # into) is only ever touched by that device's own thread, which converts a batch's records into
    # WriteOperations before it starts the next match.  Batches are dealt round-robin and their results consumed
    # in submission order, so the per-file record order is the input order (= the reference's `-t 1` order) for
    # any GPU count.
    executors = [ThreadPoolExecutor(max_workers=1) for _ in range(n_gpus)]
    try:
        pending = deque()

        def consume(ops, n, m):
            nonlocal total, matched
            for op in ops:
                output_write_operation(op, output_manager, args, trace_logger)
            total += n
            matched += m

        def drain(limit):
            while len(pending) > limit:
                consume(*pending.popleft().result())

        per_call = TRACE_BATCH_READS if trace_logger is not None else GPU_BATCH_READS
        for i, batch in enumerate(iter_batches(seq_records, per_call, args.num_seqs, all_seqs)):
            if trace_logger is not None:
                # tracing: a batch's search events and its SEQUENCE_OUTPUT events are logged back to back from
                # this thread, in the order a reference worker logs them (multiprocessing_utils.py:89-99)
                consume(*process_sequences(batch, parameters, specimens, args, prefilter, trace_logger, offset, 0))
            else:
                pending.append(executors[i % n_gpus].submit(process_sequences, batch, parameters, specimens, args,
                                                            prefilter, None, offset, i % n_gpus))
                drain(2 * n_gpus)
            offset += len(batch)
        drain(0)
    finally:
        for ex in executors:
            ex.shutdown(wait=True)
        if output_manager:
            output_manager.__exit__(None, None, None)
        if trace_logger:
            trace_logger.close()
    if total > 0:
        logging.info(f"Processed {total:,} sequences, match rate: {matched / total:.1%}")
    logging.info(f"Elapsed time: {timeit.default_timer() - start:.2f} seconds")
    if args.output_to_files:
        cleanup_empty_directories(args.output_dir)


def _native_eligible(args) -> bool:
    """The native reader / writer route carries everything except per-read trace events, which
    need Python-side record objects (trace.py)."""
    return not args.diagnostics and os.environ.get("SMX_NATIVE_IO", "1") != "0"


def _gpus_worth_starting(args, n_gpus: int) -> int:
    """GPUs the file pipeline starts for this input.  One B200 matches ~25x faster than the host side reads, packs and
    writes (SMX_IO_TRACE=1: 0.04 s of GPU work in a 0.47 s run of 765 k reads), while every further device context
    costs ~0.5 s of start-up inside the timed run (measured: -t 2 took 1.6 s where -t 1 took 0.47 s,
    profiles/r2_o_file_to_tree.md).  A further GPU is therefore started per SMX_BYTES_PER_GPU bytes of input
    (default 8 GiB; gzip input counts four-fold), never more than asked for with -t."""
    per_gpu = max(1, int(os.environ.get("SMX_BYTES_PER_GPU", str(8 << 30))))
    try:
        size = os.path.getsize(args.sequence_file)
    except OSError:
        return n_gpus
    if args.sequence_file.endswith((".gz", ".gzip")):
        size *= 4
    want = max(1, -(-size // per_gpu))
    if want < n_gpus:
        logging.info(f"Input of {size >> 20} MiB: starting {want} of {n_gpus} GPU(s) "
                     f"(one more per {per_gpu >> 20} MiB, SMX_BYTES_PER_GPU)")
    return min(n_gpus, want)


def _reader_chunk_bytes() -> int:
    """Bytes of the FASTQ file one parser thread takes at a time = one GPU batch (SMX_READER_CHUNK_BYTES).  Small
    enough that the ring of job slots (blocks + pinned buffers) is re-used many times over a file: a slot's first
    use pays the page faults of its buffers, and those serialise on the process's memory map."""
    return max(1 << 16, int(os.environ.get("SMX_READER_CHUNK_BYTES", str(24 << 20))))


def _reader_threads(n_gpus: int) -> int:
    """Parser + packer threads of the parallel reader (SMX_READER_THREADS overrides): the host cores left beside one
    feeder thread per GPU, the writer's pool and this thread: a quarter of the cores, at most 8 (more parser threads
    only contend for the memory map and the writer's cores: measured on the 16-vCPU box, profiles/r2_*file_to_tree*)."""
    env = os.environ.get("SMX_READER_THREADS")
    if env:
        return max(1, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return max(1, min(8, cores // 4))


def _parallel_read_ok(args, path: str, is_fastq: bool) -> float:
    """The byte-range reader needs a plain (not gzip) FASTQ file in the usual four-line layout, read from its first
    record to its end; the head of the file is checked for the layout.  Returns the mean bytes per record of the
    head (the pipeline sizes its pinned buffers from it), 0.0 when the file does not qualify."""
    if not is_fastq or path.endswith((".gz", ".gzip")) or args.num_seqs >= 0 or getattr(args, "start_seq", 1) > 1:
        return 0.0
    if os.environ.get("SMX_PARALLEL_READER", "1") == "0":
        return 0.0
    try:
        if os.path.getsize(path) < 2 * _reader_chunk_bytes():
            return 0.0
        with open(path, "rb") as fh:
            lines = fh.read(1 << 20).split(b"\n")[:-1]
    except OSError:
        return 0.0
    if len(lines) < 8:
        return 0.0
    n4 = len(lines) - len(lines) % 4
    for i in range(0, n4, 4):
        if not (lines[i].startswith(b"@") and lines[i + 2].startswith(b"+") and len(lines[i + 1]) == len(lines[i + 3])):
            return 0.0
    return (sum(len(x) for x in lines[:n4]) + n4) / (n4 // 4)


def _run_native(args, specimens, parameters, n_gpus: int, prefilter, _binding=None):
    """FASTQ/FASTA file -> output tree with no per-read Python objects: native reader -> 2-bit packer
    -> smx_match_batch on `n_gpus` GPUs -> native writer.  Every stage is a GIL-free C call, run by its own threads:
    K parser threads read and pack (each a byte range of the file at a time, re-synchronised on FASTQ record
    boundaries; one serial reader for gzip / FASTA / -n), one feeder thread per GPU matches, one writer thread formats
    and appends in input order, so per-file record order is the input order (= the reference's `-t 1` order) for any
    thread or GPU count.  Returns (total reads, matched reads)."""
    import queue
    import threading
    from .demultiplex import get_matcher
    from .engine import PackedBatch
    from .io_utils import detect_file_format
    from .native_io import FastxReader, ReadBlock, TreeWriter

    fmt = detect_file_format(args.sequence_file)
    args.isfastq = fmt == "fastq"
    t_setup0 = timeit.default_timer()
    if n_gpus > 1:
        # device contexts come up side by side (each is ~0.5 s of driver work)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=n_gpus) as ex:
            matchers = list(ex.map(lambda dev: get_matcher(parameters, specimens, args, prefilter, dev, _binding),
                                   range(n_gpus)))
    else:
        matchers = [get_matcher(parameters, specimens, args, prefilter, 0, _binding)]
    writer = TreeWriter(args.output_dir if args.output_to_files else None, args.output_file_prefix, args.isfastq,
                        matchers[0].tables)
    record_bytes = _parallel_read_ok(args, args.sequence_file, args.isfastq)
    parallel = record_bytes > 0
    chunk_bytes = _reader_chunk_bytes()
    n_parsers = _reader_threads(n_gpus) if parallel else 1

    class Job:
        def __init__(self):
            self.block, self.batch, self.pool = ReadBlock(), None, {}
            self.result = self.error = None
            self.done = threading.Event()
            self.index = -1

    n_jobs = n_parsers + 2 * n_gpus + 2
    free = queue.Queue()
    for _ in range(n_jobs):
        free.put(Job())
    # Every job slot's pinned buffers (packed reads in, records out) are sized ONCE, before any GPU work: pinned
    # allocation and release synchronise with every device of the process, and a slot that grew its buffers in
    # the middle of a run stalled all feeder threads (with two GPUs: 0.9 s one run, 2.6 s the next).
    reads_hint = int(chunk_bytes / record_bytes * 1.25) + 1024 if parallel else GPU_BATCH_READS
    if _binding is None:
        for _ in range(n_jobs):
            job = free.get()
            job.batch = PackedBatch.reserve(reads_hint, parameters.search_len)
            matchers[0].reserve_results(job.pool, reads_hint, compact="wire")
            free.put(job)
    setup_s = timeit.default_timer() - t_setup0
    gpu_q = [queue.Queue() for _ in range(n_gpus)]
    errors = []
    counts = [0, 0]
    busy = {"parse": 0.0, "pack": 0.0, "match": 0.0, "write": 0.0}     # summed over threads (SMX_IO_TRACE=1 logs them)
    clock = timeit.default_timer
    # jobs reach the writer in input order: `slots[i]` is job i once its parser has claimed it
    order_lock = threading.Condition()
    slots = {}
    state = {"next_index": 0, "n_chunks": None}

    keep_alive = []

    def gpu_worker(dev):
        _lib.bind_thread_to_gpu_numa_node(dev)
        if _binding is None and os.environ.get("SMX_GPU_WARMUP", "1") != "0":
            # While the parser threads work on their first chunks this thread has nothing to do: it runs one batch of
            # `reads_hint` dummy reads so that the device's one-time work -- streams, the lanes' ~100 device buffers at
            # their final size, pinned staging, lazy loading of every kernel on the path -- is done when the first real
            # batch arrives instead of inside it (that first call took 0.03 .. 1.0 s, profiles/r2_o_file_to_tree.md).
            t0 = clock()
            try:
                n_w, l_w = reads_hint, 2 * int(parameters.search_len) + 40
                dummy = PackedBatch.from_blob(b"A" * (n_w * l_w), np.arange(n_w + 1, dtype=np.uint64) * l_w,
                                              clip=parameters.search_len)
                pool = {}
                matchers[dev].match(dummy, reuse=pool, compact="wire")
                keep_alive.append((dummy, pool))       # released after the run: freeing pinned memory synchronises the device
            except BaseException:
                pass                                   # a real batch will surface whatever is wrong
            busy["warmup"] = busy.get("warmup", 0.0) + clock() - t0
        while True:
            job = gpu_q[dev].get()
            if job is None:
                return
            t0 = clock()
            try:
                if job.block.n_reads:
                    job.result = matchers[dev].match(job.batch, reuse=job.pool, compact="wire")   # 16-byte records: all the files need
            except BaseException as e:          # surfaced by the writer thread in input order
                job.error = e
            busy["match"] += clock() - t0
            job.done.set()

    def write_worker():
        # The hand-over to the writer is deferred: writer.write_deferred() returns once job i is planned and with the
        # writer's worker threads, and job i's block stays untouched until the NEXT writer call has returned -- so job
        # i goes back to the free list only after job i + 1 has been handed over (or after the final wait).
        i = 0
        held = None
        try:
            while True:
                with order_lock:
                    while i not in slots and not (state["n_chunks"] is not None and i >= state["n_chunks"]):
                        order_lock.wait()
                    if i not in slots:
                        return
                    job = slots.pop(i)
                job.done.wait()
                wrote = False
                try:
                    if job.error is not None:
                        raise job.error
                    if not errors and job.block.n_reads:
                        t0 = clock()
                        if job.result.records.dtype == _lib.RECORD16_DTYPE and args.output_to_files:
                            writer.write_deferred(job.block, job.result.records)
                            wrote = True
                        else:
                            writer.write(job.block, job.result.records)
                        busy["write"] += clock() - t0
                        counts[0] += job.block.n_reads
                        counts[1] += job.result.n_matched
                except BaseException as e:
                    errors.append(e)
                if held is not None:                   # the call above waited for the held job's formatting
                    release(held)
                    held = None
                if wrote:
                    held = job
                else:
                    release(job)
                i += 1
        finally:
            if held is not None:
                try:
                    writer.wait()
                except BaseException as e:
                    errors.append(e)
                release(held)

    def release(job):
        job.result = job.error = None
        job.done.clear()
        free.put(job)

    def submit(job, index):
        job.index = index
        with order_lock:
            slots[index] = job
            order_lock.notify_all()
        gpu_q[index % n_gpus].put(job)

    def parse_worker(n_chunks):
        # a job slot is taken BEFORE the chunk index, so the lowest outstanding chunk always owns a slot
        while not errors:
            job = free.get()
            with order_lock:
                index = state["next_index"]
                if index < n_chunks:
                    state["next_index"] += 1
            if index >= n_chunks:
                free.put(job)
                return
            try:
                lo = index * chunk_bytes
                t0 = clock()
                with FastxReader(args.sequence_file, True, byte_range=(lo, lo + chunk_bytes)) as rd:
                    rd.next_block(0x7FFFFFFF, job.block)
                t1 = clock()
                if job.block.n_reads:
                    job.batch = PackedBatch.from_block(job.block, clip=parameters.search_len, reuse=job.batch)
                busy["parse"] += t1 - t0
                busy["pack"] += clock() - t1
            except BaseException as e:
                job.error = e
            submit(job, index)
        with order_lock:                        # stopped by an error: no index beyond the claimed ones will come
            state["n_chunks"] = min(state["n_chunks"], state["next_index"])
            order_lock.notify_all()

    threads = [threading.Thread(target=gpu_worker, args=(d,), daemon=True) for d in range(n_gpus)]
    threads.append(threading.Thread(target=write_worker, daemon=True))
    for t in threads:
        t.start()
    try:
        if parallel:
            n_chunks = -(-os.path.getsize(args.sequence_file) // chunk_bytes)
            with order_lock:
                state["n_chunks"] = n_chunks
            logging.info(f"Parallel reader: {n_parsers} parser threads over {n_chunks} byte ranges of "
                         f"{chunk_bytes >> 10} KiB")
            parsers = [threading.Thread(target=parse_worker, args=(n_chunks,), daemon=True) for _ in range(n_parsers)]
            for t in parsers:
                t.start()
            for t in parsers:
                t.join()
        else:
            submitted = 0
            reader = FastxReader(args.sequence_file, args.isfastq)
            try:
                if getattr(args, "start_seq", 1) > 1:
                    reader.skip(args.start_seq - 1)
                remaining = args.num_seqs if args.num_seqs >= 0 else None
                while not errors and (remaining is None or remaining > 0):
                    job = free.get()
                    want = GPU_BATCH_READS if remaining is None else min(GPU_BATCH_READS, remaining)
                    t0 = clock()
                    reader.next_block(want, job.block)
                    t1 = clock()
                    busy["parse"] += t1 - t0
                    if job.block.n_reads == 0:
                        free.put(job)
                        break
                    if remaining is not None:
                        remaining -= job.block.n_reads
                    job.batch = PackedBatch.from_block(job.block, clip=parameters.search_len, reuse=job.batch)
                    busy["pack"] += clock() - t1
                    submit(job, submitted)
                    submitted += 1
            finally:
                reader.close()
                with order_lock:
                    state["n_chunks"] = submitted
                    order_lock.notify_all()
    except BaseException as e:
        errors.append(e)
        with order_lock:
            if state["n_chunks"] is None:
                state["n_chunks"] = 0
            order_lock.notify_all()
    finally:
        for q in gpu_q:
            q.put(None)
        for t in threads:
            t.join()
        t0 = clock()
        try:
            writer.close()
        except BaseException as e:
            errors.append(e)
        busy["write"] += clock() - t0
    if os.environ.get("SMX_IO_TRACE"):
        logging.info("I/O pipeline busy seconds (summed over threads): " + ", ".join("%s %.3f" % kv for kv in busy.items()) +
                     "; %d parser thread(s), %d GPU(s); set-up (device context(s), pinned buffers, writer) %.3f s before the first read" % (n_parsers, n_gpus, setup_s))
    if errors:
        raise errors[0]
    return counts[0], counts[1]


def specimux_mp(args):
    """reference: orchestration.py:153-237 -- file output, all GPUs (one feeder thread per GPU)."""
    _run(args, multi=True)


def specimux(args):
    """reference: orchestration.py:458-545 -- console output, one GPU."""
    _run(args, multi=False)
