"""Batching layer over the C ABI: pack reads, run the GPU path, hand back numpy records.

`Matcher` owns one device context (one per GPU; not re-entrant).  There is no CPU path: a
missing library raises ImportError, a missing GPU raises SmxError(SMX_ERR_NO_DEVICE).
"""
import ctypes as C
import os
import threading
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


class PackedBatch:
    """Reads in the smx_batch encoding (2-bit stream + optional exact 4-bit side stream), held in
    pinned host memory.  `clip` (0 or >= search_len) stores only the first/last `clip` bases of
    longer reads -- the only bases the path looks at -- which cuts packing work and H2D bytes."""

    def __init__(self, bases: Sequence[str], clip: int = 0):
        n = len(bases)
        lens = np.fromiter((len(s) for s in bases), dtype=np.uint64, count=n)
        seq_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(lens, out=seq_off[1:])
        blob = "".join(bases).encode("latin-1", "replace")
        self._init_from_blob(blob, seq_off, clip)

    @classmethod
    def from_blob(cls, blob: bytes, seq_off: np.ndarray, clip: int = 0):
        self = cls.__new__(cls)
        self._init_from_blob(blob, np.ascontiguousarray(seq_off, dtype=np.uint64), clip)
        return self

    @classmethod
    def from_block(cls, block, clip: int = 0, reuse: "PackedBatch" = None):
        """Packs a native_io.ReadBlock straight from the reader's memory (no Python strings).
        `reuse`: a PackedBatch whose pinned buffers are recycled when they are large enough."""
        self = reuse if reuse is not None else cls.__new__(cls)
        self._init_from_blob(block.bases_ptr or b"", block.seq_off(), clip)
        return self

    @classmethod
    def reserve(cls, n_reads: int, clip: int):
        """An empty batch whose pinned buffers already hold `n_reads` clipped reads: a pipeline that re-packs into
        it (from_block(..., reuse=...)) then never allocates or frees pinned memory while the GPU is busy --
        cudaHostAlloc / cudaFreeHost synchronise with every device of the process."""
        self = cls.__new__(cls)
        if clip:
            stride = int(_lib.load().smx_pack_stride(clip))
            self._buffer(0, n_reads * stride + 1, np.uint32)
            for slot, dt in ((2, np.uint32), (3, np.uint64), (4, np.uint16)):
                self._buffer(slot, max(n_reads, 1), dt)
        self.n_reads = 0
        return self

    def _buffer(self, slot, count, dtype):
        """Pinned host buffer `slot`, grown geometrically and kept across repacks."""
        bufs = self.__dict__.setdefault("_bufs", {})
        hb = bufs.get(slot)
        if hb is None or len(hb.array) < count:
            if hb is not None:
                hb.free()
            hb = bufs[slot] = _lib.HostBuffer(max(int(count * 1.25), 1), dtype)
        return hb.array[:count]

    def _init_from_blob(self, blob, seq_off, clip):
        """blob: bytes, or the address of the concatenated bases.  Clipped batches (clip > 0) are packed at a fixed
        stride (no per-read offsets on the wire) with 16-bit lengths when every read is shorter than 65,536."""
        lib = _lib.load()
        n = len(seq_off) - 1
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        w2, w4 = C.c_uint64(0), C.c_uint64(0)
        lib.smx_pack_bound(_lib.ptr(seq_off, _lib.u64p), n, clip, C.byref(w2), C.byref(w4))
        self.n_reads = n
        self.clip = int(clip)
        self.stride = int(lib.smx_pack_stride(clip)) if clip else 0
        self.lengths = self._buffer(2, max(n, 1), np.uint32)
        self.off4 = self._buffer(3, max(n, 1), np.uint64)
        packed4 = np.zeros(int(w4.value), dtype=np.uint32)
        used4, flagged = C.c_uint64(0), C.c_uint32(0)
        src = blob if isinstance(blob, (bytes, bytearray)) else C.cast(C.c_void_p(int(blob)), C.c_char_p)
        if self.stride:
            self.packed2 = self._buffer(0, n * self.stride + 1, np.uint32)
            self.word_off = None
            l16 = self._buffer(4, max(n, 1), np.uint16)
            fit = C.c_int(0)
            _lib.check(lib.smx_pack_reads_fixed(src, _lib.ptr(seq_off, _lib.u64p), n, clip, _lib.ptr(self.packed2, _lib.u32p),
                                                _lib.ptr(self.lengths, _lib.u32p), _lib.ptr(l16, _lib.u16p), C.byref(fit),
                                                _lib.ptr(packed4, _lib.u32p), _lib.ptr(self.off4, _lib.u64p),
                                                C.byref(used4), C.byref(flagged)))
            self.lengths16 = l16 if fit.value else None
        else:
            self.packed2 = self._buffer(0, int(w2.value), np.uint32)
            self.word_off = self._buffer(1, max(n, 1), np.uint64)
            self.lengths16 = None
            _lib.check(lib.smx_pack_reads(src, _lib.ptr(seq_off, _lib.u64p), n, clip, _lib.ptr(self.packed2, _lib.u32p),
                                          _lib.ptr(self.word_off, _lib.u64p), _lib.ptr(self.lengths, _lib.u32p),
                                          _lib.ptr(packed4, _lib.u32p), _lib.ptr(self.off4, _lib.u64p),
                                          C.byref(used4), C.byref(flagged)))
        self.n_flagged = int(flagged.value)
        self.packed4 = packed4[:int(used4.value)].copy() if self.n_flagged else None
        self.h2d_bytes = (self.packed2.nbytes + (self.word_off.nbytes if self.word_off is not None else 0) +
                          (self.lengths16.nbytes if self.lengths16 is not None else self.lengths.nbytes) +
                          (self.packed4.nbytes + self.off4.nbytes if self.n_flagged else 0))

    def c_batch(self) -> "_lib.SmxBatch":
        b = _lib.SmxBatch()
        b.n_reads = self.n_reads
        b.clip_len = self.clip
        b.packed2 = _lib.ptr(self.packed2, _lib.u32p)
        b.packed2_words = len(self.packed2)
        b.stride_words = self.stride
        b.word_off = _lib.ptr(self.word_off, _lib.u64p) if self.word_off is not None else None
        b.lengths = _lib.ptr(self.lengths, _lib.u32p)
        b.lengths16 = _lib.ptr(self.lengths16, _lib.u16p) if self.lengths16 is not None else None
        if self.n_flagged:
            b.packed4 = _lib.ptr(self.packed4, _lib.u32p)
            b.packed4_words = len(self.packed4)
            b.off4 = _lib.ptr(self.off4, _lib.u64p)
        else:
            b.packed4 = None
            b.packed4_words = 0
            b.off4 = None
        return b


class BatchResult:
    def __init__(self, rec_offset, records, n_matched, primer_hits=None, endmask=None, barcode_hits=None, orient_hits=None):
        self.rec_offset = rec_offset
        self.records = records
        self.n_matched = n_matched
        self.primer_hits = primer_hits      # [2*n_primers, n_reads]
        self.endmask = endmask              # [2*n_primers, mask_words, n_reads] uint32
        self.barcode_hits = barcode_hits    # [total_barcode_slots, n_reads]
        self.orient_hits = orient_hits      # [2*n_primers, n_reads] uint8 (explicit orientation test, irregular reads)
        self.barcode_loc_hits = None        # detail="locations": every barcode hit at every primer end location


class Matcher:
    """One device context.  `binding` lets the unit tests drive the same host code through the CPU
    kernel simulator (tests/hostsim); the product never passes it, and it is refused unless the test suite
    has opened the seam (SMX_TEST_SEAM=1, set by tests/conftest.py): there is no CPU execution route
    through the product API."""

    def __init__(self, tables, device: int = 0, binding=None):
        if binding is not None and os.environ.get("SMX_TEST_SEAM") != "1":
            raise RuntimeError("specimux_b200: the simulator binding is a test seam (SMX_TEST_SEAM=1); "
                               "the matching runs on the GPU only")
        self.tables = tables
        self._binding = binding
        self._ctx = C.c_void_p(None)
        self._lock = threading.Lock()
        if binding is None:
            self._lib = _lib.load()
            _lib.check(self._lib.smx_create(device, tables.tables_ref(), tables.params_ref(), C.byref(self._ctx)))
        else:
            self._lib = None

    def close(self):
        if self._lib is not None and self._ctx:
            self._lib.smx_destroy(self._ctx)
            self._ctx = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- results allocation ---------------------------------------------------------------
    def _alloc_results(self, n, detail, cap=None, reuse=False, compact=False):
        t = self.tables
        cap = int(cap if cap is not None else n + n // 8 + 1024)
        res = _lib.SmxResults()
        wire = compact == "wire"
        rec_dtype = _lib.RECORD16_DTYPE if wire else _lib.RECORD32_DTYPE if compact else _lib.RECORD_DTYPE
        rec_key = "rec16" if wire else "rec32" if compact else "rec"
        if reuse is not None and reuse is not False:
            # pinned pool reused across calls: the previous call's result arrays are overwritten.
            # reuse=True: one pool per Matcher; reuse=<dict>: a pool owned by the caller (lets several
            # results stay alive, e.g. while a writer thread still formats an earlier batch)
            pool = reuse if isinstance(reuse, dict) else self.__dict__.setdefault("_result_pool", {})
            # grown geometrically: every growth is a cudaHostAlloc + (later) cudaFreeHost, and both synchronise with
            # the devices -- batches of slightly different sizes must not reallocate each time
            if pool.get("n", -1) < n + 1:
                grown = n + 1 if "n" not in pool else n + n // 4 + 1024
                pool["n"], pool["off"] = grown, _lib.HostBuffer(grown, np.uint32)
            if pool.get("cap_" + rec_key, -1) < cap:
                grown = cap if ("cap_" + rec_key) not in pool else cap + cap // 4 + 1024
                pool["cap_" + rec_key], pool[rec_key] = grown, _lib.HostBuffer(grown, rec_dtype)
            rec_offset = pool["off"].array[:n + 1]
            records = pool[rec_key].array[:cap]
        else:
            rec_offset = np.empty(n + 1, dtype=np.uint32)
            records = np.empty(cap, dtype=rec_dtype)
        if wire:
            rec_offset = None                       # the last-of-read flags carry the grouping
            res.rec_offset = None
            res.records16 = records.ctypes.data
        else:
            res.rec_offset = _lib.ptr(rec_offset, _lib.u32p)
            if compact:
                res.records32 = records.ctypes.data
            else:
                res.records = records.ctypes.data
        res.records_cap = cap
        ph = em = bh = oh = None
        if detail:
            ph = np.zeros((2 * t.n_primers, n), dtype=_lib.PRIMER_HIT_DTYPE)
            em = np.zeros((2 * t.n_primers, t.mask_words, n), dtype=np.uint32)
            bh = np.zeros((max(t.total_barcode_slots, 1), n), dtype=_lib.BARCODE_HIT_DTYPE)
            res.primer_hits = ph.ctypes.data
            res.endmask_bits = em.ctypes.data
            oh = np.zeros((2 * t.n_primers, n), dtype=np.uint8)
            res.barcode_hits = bh.ctypes.data
            res.orient_hits = oh.ctypes.data
            lh = None
            if detail == "locations":
                lh = np.zeros(max(int(self.__dict__.get("_loc_cap", 0)), 64 * n + 1024), dtype=_lib.BARCODE_LOC_HIT_DTYPE)
                res.barcode_loc_hits = lh.ctypes.data
                res.barcode_loc_cap = len(lh)
            return res, rec_offset, records, ph, em, (bh, oh, lh)
        return res, rec_offset, records, ph, em, (bh, None, None)

    def _finish(self, res, rec_offset, records, ph, em, bh):
        bh, oh, lh = bh
        out = BatchResult(rec_offset, records[:int(res.n_records)], int(res.n_matched), ph, em, bh, oh)
        if lh is not None:
            out.barcode_loc_hits = lh[:min(int(res.n_barcode_loc_hits), len(lh))]
        return out

    def reserve_results(self, pool: dict, n: int, compact=False):
        """Sizes a caller-owned pinned result pool for batches of up to `n` reads ahead of the first match."""
        self._alloc_results(int(n), False, None, pool, compact)

    # -- whole path with host buffers (H2D + kernels + D2H) ---------------------------------
    def match(self, batch: PackedBatch, detail: bool = False, reuse=False, compact: bool = False) -> BatchResult:
        """`reuse=True` returns views into a pinned pool that the next reuse=True call overwrites;
        `reuse=<dict>` uses (and grows) a caller-owned pool instead.  `compact=True` returns 32-byte
        smx_record32 records (no location pairs); `compact="wire"` the 16-byte smx_record16 records in read
        order (what the output-tree writer needs; no rec_offset).  A context is not
        re-entrant: calls from several threads are serialised here."""
        with self._lock:
            return self._match_locked(batch, detail, reuse, compact)

    def _match_locked(self, batch, detail, reuse, compact=False):
        cb = batch.c_batch()
        cap = None
        while True:
            res, rec_offset, records, ph, em, bh = self._alloc_results(batch.n_reads, detail, cap, reuse, compact)
            if self._binding is not None:
                rc = self._binding.hostsim_match_batch(self.tables.tables_ref(), self.tables.params_ref(),
                                                       C.byref(cb), C.byref(res))
                if rc == _lib.SMX_ERR_CAPACITY:
                    cap = int(res.n_records)
                    continue
                if rc != 0:
                    raise _lib.SmxError(rc, self._binding.hostsim_last_error().decode())
            else:
                rc = self._lib.smx_match_batch(self._ctx, C.byref(cb), C.byref(res))
                if rc == _lib.SMX_ERR_CAPACITY:
                    cap = int(res.n_records)
                    continue
                _lib.check(rc)
            if detail == "locations" and int(res.n_barcode_loc_hits) > int(res.barcode_loc_cap):
                self._loc_cap = int(res.n_barcode_loc_hits) + 1024
                continue
            return self._finish(res, rec_offset, records, ph, em, bh)

    # -- split form -------------------------------------------------------------------------
    def upload(self, batch: PackedBatch):
        self._resident_n = batch.n_reads
        cb = batch.c_batch()
        _lib.check(self._lib.smx_upload_batch(self._ctx, C.byref(cb)))

    def run_resident(self):
        _lib.check(self._lib.smx_run_resident(self._ctx))

    def download(self, detail: bool = False, reuse: bool = False) -> BatchResult:
        cap = None
        while True:
            res, rec_offset, records, ph, em, bh = self._alloc_results(self._resident_n, detail, cap, reuse)
            rc = self._lib.smx_download_results(self._ctx, C.byref(res))
            if rc == _lib.SMX_ERR_CAPACITY:
                cap = int(res.n_records)
                continue
            _lib.check(rc)
            return self._finish(res, rec_offset, records, ph, em, bh)

    def set_pipeline_chunk(self, reads_per_chunk: int):
        """Chunk size of the pipelined smx_match_batch (0 = always one shot)."""
        _lib.check(self._lib.smx_set_pipeline_chunk(self._ctx, int(reads_per_chunk)))

    def set_resident_split(self, n_sub_batches: int):
        """Sub-batches the resident form runs concurrently (1 = one lane, per-kernel times available)."""
        _lib.check(self._lib.smx_set_resident_split(self._ctx, int(n_sub_batches)))

    def last_chunk_count(self) -> int:
        return int(self._lib.smx_last_chunk_count(self._ctx))

    # slot 3 (once the separate start-recovery kernel) is the gap between stage 1 and stage 2: ~0
    KERNEL_NAMES = ("stage_windows", "primer_sliced", "primer_finish_start", "stage_gap", "barcode_tasks",
                    "select_fast", "select_general", "scan_compact", "rebase_offsets")

    def last_kernel_times(self):
        """{kernel: ms} of the last run_resident (CUDA events on the launching stream)."""
        ms = (C.c_float * len(self.KERNEL_NAMES))()
        n = self._lib.smx_last_kernel_times(self._ctx, ms, len(self.KERNEL_NAMES))
        return {k: float(ms[i]) for i, k in enumerate(self.KERNEL_NAMES[:n])}

    def last_deferred(self) -> int:
        return int(self._lib.smx_last_deferred(self._ctx))

    def flush_l2(self):
        _lib.check(self._lib.smx_flush_l2(self._ctx))

    def last_timing(self):
        total = C.c_float(0)
        stages = (C.c_float * 4)()
        _lib.check(self._lib.smx_last_timing(self._ctx, C.byref(total), stages))
        return float(total.value), [float(x) for x in stages]

    def last_launch_count(self) -> int:
        return int(self._lib.smx_last_launch_count(self._ctx))

    def last_useful_cells(self):
        """(stage 1, stage 2) bit-sliced DP cells evaluated by the last run_resident (5 LOP3 per cell)."""
        cells = (C.c_uint64 * 2)()
        _lib.check(self._lib.smx_last_useful_cells(self._ctx, cells))
        return int(cells[0]), int(cells[1])

    def last_work(self):
        cells = (C.c_uint64 * 2)()
        wcols = (C.c_uint64 * 2)()
        _lib.check(self._lib.smx_last_work(self._ctx, cells, wcols))
        return [int(x) for x in cells], [int(x) for x in wcols]
