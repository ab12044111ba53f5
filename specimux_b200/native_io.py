"""ctypes binding of libspecimux_io.so (include/specimux_io.h): the native FASTQ/FASTA(.gz) reader
and the native per-specimen output-tree writer either side of the GPU matching path
(SURVEY.md 8f rank 1).  Text never becomes Python objects on this route: the reader's block feeds
smx_pack_reads by pointer, the writer formats smx_records straight from the block.

Byte-compatibility targets: Bio.SeqIO.parse as the reference uses it (io_utils.py:429-450) and
OutputManager / output_write_operation (io_utils.py:179-268, 452-471); both are also implemented
in Python (seqio.py, io_utils.py) and the two are compared file-for-file in tests/test_native_io.py.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .constants import SampleId

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspecimux_io.so")

u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
strp = C.POINTER(C.c_char_p)

EXPORTS = ["smx_io_abi_version", "smx_io_last_error", "smx_reader_open", "smx_reader_open_range", "smx_reader_close", "smx_block_create",
           "smx_block_destroy", "smx_block_get", "smx_reader_next", "smx_reader_skip", "smx_writer_open",
           "smx_writer_write", "smx_writer_write32", "smx_writer_write16", "smx_writer_write16_deferred", "smx_writer_wait",
           "smx_writer_close", "smx_writer_stats"]


class SmxBlockView(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("bases", C.c_void_p), ("seq_off", u64p), ("quals", C.c_void_p),
                ("titles", C.c_void_p), ("title_off", u64p), ("id_start", u32p), ("id_len", u32p)]


class SmxNames(C.Structure):
    _fields_ = [("n_specimens", C.c_uint32), ("specimen_id", strp), ("specimen_file", strp),
                ("n_b1", C.c_uint32), ("b1_id", strp), ("b1_file", strp),
                ("n_b2", C.c_uint32), ("b2_id", strp), ("b2_file", strp),
                ("n_pools", C.c_uint32), ("pool", strp),
                ("n_primers", C.c_uint32), ("primer_name", strp)]


class SmxIoError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_io = None


def load():
    global _io
    if _io is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `make -C specimux_b200/csrc` (or __graft_entry__.build())" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.smx_io_last_error.restype = C.c_char_p
        lib.smx_reader_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.smx_reader_open_range.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]
        lib.smx_reader_close.argtypes = [C.c_void_p]
        lib.smx_reader_close.restype = None
        lib.smx_block_create.restype = C.c_void_p
        lib.smx_block_destroy.argtypes = [C.c_void_p]
        lib.smx_block_destroy.restype = None
        lib.smx_block_get.argtypes = [C.c_void_p, C.POINTER(SmxBlockView)]
        lib.smx_block_get.restype = None
        lib.smx_reader_next.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.smx_reader_skip.argtypes = [C.c_void_p, C.c_uint64, u64p]
        lib.smx_writer_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(SmxNames), C.POINTER(C.c_void_p)]
        lib.smx_writer_write.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.smx_writer_write32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.smx_writer_write16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.smx_writer_write16_deferred.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.smx_writer_wait.argtypes = [C.c_void_p]
        lib.smx_writer_close.argtypes = [C.c_void_p]
        lib.smx_writer_stats.argtypes = [C.c_void_p, u64p, u64p]
        lib.smx_writer_stats.restype = None
        if lib.smx_io_abi_version() != 2:
            raise ImportError("libspecimux_io.so ABI version mismatch")
        _io = lib
    return _io


def _check(rc):
    if rc != 0:
        msg = load().smx_io_last_error().decode("utf-8", "replace")
        # the reference surfaces parser problems as ValueError (Bio.SeqIO) and I/O problems as OSError
        if rc == 3:
            raise ValueError(msg)
        if rc in (2, 4):
            raise OSError(msg)
        raise SmxIoError(rc, msg)


class ReadBlock:
    """One batch of parsed reads held by the native library."""

    def __init__(self):
        self._lib = load()
        self._h = C.c_void_p(self._lib.smx_block_create())
        self._view = SmxBlockView()

    def close(self):
        if self._h:
            self._lib.smx_block_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh(self):
        self._lib.smx_block_get(self._h, C.byref(self._view))
        return self

    @property
    def handle(self):
        return self._h

    @property
    def n_reads(self) -> int:
        return int(self._view.n_reads)

    @property
    def has_quality(self) -> bool:
        return bool(self._view.quals)

    @property
    def bases_ptr(self) -> int:
        return int(self._view.bases or 0)

    def seq_off(self) -> np.ndarray:
        """n_reads + 1 offsets (a view into the block; valid until the block is refilled)."""
        n = self.n_reads
        return np.ctypeslib.as_array(self._view.seq_off, shape=(n + 1,))

    def n_bases(self) -> int:
        return int(self._view.seq_off[self.n_reads]) if self.n_reads else 0

    # -- Python-object accessors (tests, trace path); the fast route never calls these -----------
    def _text(self, base, lo, hi):
        return C.string_at(base + lo, hi - lo).decode("utf-8", "surrogateescape") if hi > lo else ""

    def read(self, r):
        """(id, description, bases, quality or None) of read r."""
        v = self._view
        s0, s1 = int(v.seq_off[r]), int(v.seq_off[r + 1])
        t0, t1 = int(v.title_off[r]), int(v.title_off[r + 1])
        title = self._text(v.titles, t0, t1)
        i0 = t0 + int(v.id_start[r])
        rid = self._text(v.titles, i0, i0 + int(v.id_len[r]))
        return rid, title, self._text(v.bases, s0, s1), (self._text(v.quals, s0, s1) if v.quals else None)

    def records(self):
        from .seqio import SeqRecord
        out = []
        for r in range(self.n_reads):
            rid, title, seq, qual = self.read(r)
            out.append(SeqRecord(seq, rid, title, qual))
        return out


class FastxReader:
    """Native FASTQ / FASTA reader (plain or gzip)."""

    def __init__(self, path: str, is_fastq: bool, byte_range=None):
        """byte_range=(start, end): only the records that start inside that range of a plain four-line FASTQ file
        (both ends moved forward to a record boundary; consecutive ranges see every record exactly once)."""
        self._lib = load()
        self._h = C.c_void_p(None)
        if byte_range is None:
            _check(self._lib.smx_reader_open(os.fsencode(path), 1 if is_fastq else 0, C.byref(self._h)))
        else:
            _check(self._lib.smx_reader_open_range(os.fsencode(path), 1 if is_fastq else 0, int(byte_range[0]),
                                                   int(byte_range[1]), C.byref(self._h)))
        self.is_fastq = is_fastq

    def close(self):
        if self._h:
            self._lib.smx_reader_close(self._h)
            self._h = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def next_block(self, max_reads: int, block: ReadBlock = None) -> ReadBlock:
        """Fills `block` (or a new one) with up to max_reads records; n_reads == 0 at end of file."""
        block = block or ReadBlock()
        _check(self._lib.smx_reader_next(self._h, int(max_reads), block.handle))
        return block.refresh()

    def skip(self, n: int) -> int:
        done = C.c_uint64(0)
        _check(self._lib.smx_reader_skip(self._h, int(n), C.byref(done)))
        return int(done.value)


def safe_file_id(sample_id: str) -> str:
    """OutputManager._make_filename's safe_id (reference io_utils.py:215-216), Python's character classes."""
    return "".join(c if c.isalnum() or c in "._-$#" else "_" for c in sample_id)


def _str_array(strings):
    arr = (C.c_char_p * max(len(strings), 1))()
    for i, s in enumerate(strings):
        arr[i] = s.encode("utf-8", "surrogateescape")
    return arr


class TreeWriter:
    """Native output writer: smx_records + the block they index -> the per-specimen FASTQ/FASTA tree
    (output_dir given) or the console form on stdout (output_dir None)."""

    def __init__(self, output_dir, prefix: str, is_fastq: bool, tables):
        self._lib = load()
        spec_ids = list(tables.specimen_ids)
        b1_ids = [SampleId.PREFIX_FWD_MATCH + b for b in tables.b1]
        b2_ids = [SampleId.PREFIX_REV_MATCH + b for b in tables.b2]
        k = self._keep = {
            "spec": _str_array(spec_ids), "spec_f": _str_array([safe_file_id(s) for s in spec_ids]),
            "b1": _str_array(b1_ids), "b1_f": _str_array([safe_file_id(s) for s in b1_ids]),
            "b2": _str_array(b2_ids), "b2_f": _str_array([safe_file_id(s) for s in b2_ids]),
            "pool": _str_array(list(tables.pools)), "primer": _str_array(list(tables.primer_names)),
        }
        nm = SmxNames()
        nm.n_specimens, nm.specimen_id, nm.specimen_file = len(spec_ids), k["spec"], k["spec_f"]
        nm.n_b1, nm.b1_id, nm.b1_file = len(b1_ids), k["b1"], k["b1_f"]
        nm.n_b2, nm.b2_id, nm.b2_file = len(b2_ids), k["b2"], k["b2_f"]
        nm.n_pools, nm.pool = len(tables.pools), k["pool"]
        nm.n_primers, nm.primer_name = len(tables.primer_names), k["primer"]
        self._h = C.c_void_p(None)
        _check(self._lib.smx_writer_open(None if output_dir is None else os.fsencode(output_dir),
                                         (prefix or "").encode("utf-8"), 1 if is_fastq else 0, C.byref(nm),
                                         C.byref(self._h)))

    def write(self, block: ReadBlock, records: np.ndarray):
        """records: smx_record (RECORD_DTYPE), the compact smx_record32 or the 16-byte wire form smx_record16."""
        rec = np.ascontiguousarray(records)
        if rec.dtype == _lib.RECORD16_DTYPE:
            _check(self._lib.smx_writer_write16(self._h, block.handle, rec.ctypes.data, len(rec)))
        elif rec.dtype == _lib.RECORD32_DTYPE:
            _check(self._lib.smx_writer_write32(self._h, block.handle, rec.ctypes.data, len(rec)))
        else:
            assert rec.dtype == _lib.RECORD_DTYPE
            _check(self._lib.smx_writer_write(self._h, block.handle, rec.ctypes.data, len(rec)))

    def write_deferred(self, block: ReadBlock, records: np.ndarray):
        """smx_record16 records only: returns once the records are planned and with the worker threads; `block` must stay
        untouched until the next write / write_deferred / wait / close call on this writer has returned."""
        rec = np.ascontiguousarray(records)
        assert rec.dtype == _lib.RECORD16_DTYPE
        _check(self._lib.smx_writer_write16_deferred(self._h, block.handle, rec.ctypes.data, len(rec)))

    def wait(self):
        _check(self._lib.smx_writer_wait(self._h))

    def stats(self):
        n, b = C.c_uint64(0), C.c_uint64(0)
        self._lib.smx_writer_stats(self._h, C.byref(n), C.byref(b))
        return int(n.value), int(b.value)

    def close(self):
        if self._h:
            h, self._h = self._h, C.c_void_p(None)
            _check(self._lib.smx_writer_close(h))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        if exc_type is None:
            self.close()
        else:
            try:
                self.close()
            except Exception:
                pass
        return False
