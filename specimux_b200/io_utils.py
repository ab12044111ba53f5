"""File formats either side of the matching path (mirror of the reference's io_utils.py interface).

primers.fasta / specimens.txt parsing (reference io_utils.py:270-377), sequence file opening
(:380-450) and the per-specimen output tree (:179-268, :452-471).  Host-side I/O only.
"""
import csv
import gzip
import logging
import os
import sys
from collections import OrderedDict
from typing import Optional

from . import seqio
from .constants import Primer, ResolutionType, SampleId
from .databases import PrimerDatabase, Specimens
from .models import PrimerInfo, WriteOperation


def read_primers_file(filename: str) -> PrimerDatabase:
    """reference: io_utils.py:270-322."""
    registry = PrimerDatabase()
    for file_index, record in enumerate(seqio.parse(filename, "fasta")):
        pool_names, position = [], None
        for field in record.description.split():
            if field.startswith("pool="):
                pool_names = [p.strip() for p in field[5:].replace(";", ",").split(",")]
            elif field.startswith("position="):
                position = field[9:]
        if not pool_names:
            raise ValueError(f"Missing pool specification for primer {record.id}")
        if not position:
            raise ValueError(f"Missing position specification for primer {record.id}")
        if position == "forward":
            direction = Primer.FWD
        elif position == "reverse":
            direction = Primer.REV
        else:
            raise ValueError(f"Invalid primer position '{position}' for {record.id}")
        registry.add_primer(PrimerInfo(record.id, record.seq, direction, pool_names, file_index=file_index), pool_names)
    registry.validate_pools()
    stats = registry.get_pool_stats()
    logging.info(f"Loaded {stats['total_primers']} primers in {stats['total_pools']} pools")
    for pool, ps in stats["pools"].items():
        logging.info(f"Pool {pool}: {ps['forward_primers']} forward, {ps['reverse_primers']} reverse primers")
    return registry


def read_specimen_file(filename: str, primer_registry: PrimerDatabase) -> Specimens:
    """reference: io_utils.py:324-377."""
    specimens = Specimens(primer_registry)
    expected = {"SampleID", "PrimerPool", "FwIndex", "FwPrimer", "RvIndex", "RvPrimer"}
    with open(filename, "r", newline="") as fh:
        reader = csv.DictReader(fh, delimiter="\t")
        missing = expected - set(reader.fieldnames or [])
        if missing:
            raise ValueError(f"Missing required columns in specimen file: {missing}")
        empty = []
        for row_num, row in enumerate(reader, start=1):
            try:
                b1, b2 = row["FwIndex"].upper(), row["RvIndex"].upper()
                if not b1.strip() or not b2.strip():
                    empty.append(f"Row {row_num} ({row['SampleID']}): "
                                 f"{'FwIndex is empty' if not b1.strip() else 'RvIndex is empty'}")
                    continue
                specimens.add_specimen(specimen_id=row["SampleID"], pool=row["PrimerPool"], b1=b1,
                                       p1=row["FwPrimer"], b2=b2, p2=row["RvPrimer"])
            except (KeyError, ValueError) as e:
                raise ValueError(f"Error processing row {row_num}: {e}")
        if empty:
            raise ValueError(f"Empty barcodes found in {len(empty)} specimen(s). "
                             f"Single-indexed demultiplexing is not supported.\n" + "\n".join(empty[:10])
                             + (f"\n... and {len(empty) - 10} more" if len(empty) > 10 else ""))
    if not specimens._specimens:
        raise ValueError("No valid data found in the specimen file")
    return specimens


def detect_file_format(filename: str) -> str:
    """reference: io_utils.py:380-427."""
    base = os.path.basename(filename)
    root, ext = os.path.splitext(base)
    while ext.lower() in (".gz", ".gzip", ".bz2", ".zip"):
        base = root
        root, ext = os.path.splitext(base)
    low = base.lower()
    if low.endswith((".fastq", ".fq")):
        return "fastq"
    if low.endswith((".fasta", ".fa", ".fna")):
        return "fasta"
    try:
        opener = gzip.open if filename.endswith((".gz", ".gzip")) else open
        with opener(filename, "rt") as fh:
            first = fh.read(1)
        if first == "@":
            return "fastq"
    except Exception:
        pass
    return "fasta"


def open_sequence_file(filename, args):
    """reference: io_utils.py:429-450 (sets args.isfastq)."""
    fmt = detect_file_format(filename)
    args.isfastq = fmt == "fastq"
    return seqio.parse(filename, fmt)


class OutputManager:
    """Per-specimen output tree (reference: io_utils.py:179-268).  Path rules, header format and the
    pool-level duplicate of full matches are identical; writes are buffered per file and appended
    in arrival order by a single writer, which gives the `-t 1` record order of the reference."""

    def __init__(self, output_dir: str, prefix: str, is_fastq: bool, max_open_files: int = 200,
                 buffer_size: int = 500):
        self.output_dir = output_dir
        self.prefix = prefix
        self.is_fastq = is_fastq
        self.buffer_size = buffer_size
        self._buffers = OrderedDict()
        self._made_dirs = set()

    def __enter__(self):
        os.makedirs(self.output_dir, exist_ok=True)
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.flush_all()
        return False

    def _make_filename(self, sample_id, pool, p1, p2, resolution_type: ResolutionType) -> str:
        ext = ".fastq" if self.is_fastq else ".fasta"
        sample_id = sample_id or SampleId.UNKNOWN
        pool, p1, p2 = pool or "unknown", p1 or "unknown", p2 or "unknown"
        safe_id = "".join(c if c.isalnum() or c in "._-$#" else "_" for c in sample_id)
        top = "unknown" if resolution_type.is_unknown() else "partial" if resolution_type.is_partial_match() else "full"
        return os.path.join(self.output_dir, top, pool, f"{p1}-{p2}", f"{self.prefix}{safe_id}{ext}")

    def _append(self, filename: str, data: str):
        buf = self._buffers.setdefault(filename, [])
        buf.append(data)
        if len(buf) >= self.buffer_size:
            self._flush(filename)

    def _flush(self, filename: str):
        buf = self._buffers.get(filename)
        if not buf:
            return
        d = os.path.dirname(filename)
        if d not in self._made_dirs:
            os.makedirs(d, exist_ok=True)
            self._made_dirs.add(d)
        with open(filename, "a") as fh:
            fh.write("".join(buf))
        buf.clear()

    def flush_all(self):
        for filename in list(self._buffers):
            self._flush(filename)

    def write_sequence(self, write_op: WriteOperation, trace_logger=None):
        filename = self._make_filename(write_op.sample_id, write_op.primer_pool, write_op.p1_name,
                                       write_op.p2_name, write_op.resolution_type)
        if trace_logger:
            trace_logger.log_sequence_output(write_op.trace_sequence_id, write_op.sample_id, write_op.primer_pool,
                                             f"{write_op.p1_name}-{write_op.p2_name}",
                                             os.path.relpath(filename, self.output_dir))
        header = (f"{write_op.seq_id} {write_op.distance_code} pool={write_op.primer_pool} "
                  f"primers={write_op.p1_name}+{write_op.p2_name} {write_op.sample_id}")
        if self.is_fastq:
            content = f"@{header}\n{write_op.sequence}\n+\n{write_op.quality_sequence}\n"
        else:
            content = f">{header}\n{write_op.sequence}\n"
        self._append(filename, content)
        if write_op.resolution_type.is_full_match():
            ext = ".fastq" if self.is_fastq else ".fasta"
            safe_id = "".join(c if c.isalnum() or c in "._-$#" else "_" for c in write_op.sample_id)
            self._append(os.path.join(self.output_dir, "full", write_op.primer_pool, f"{self.prefix}{safe_id}{ext}"),
                         content)


def output_write_operation(write_op: WriteOperation, output_manager: Optional[OutputManager], args,
                           trace_logger=None) -> None:
    """reference: io_utils.py:452-471."""
    if not args.output_to_files:
        symbol = "@" if args.isfastq else ">"
        sys.stdout.write(f"{symbol}{write_op.seq_id} {write_op.distance_code} {write_op.sample_id}\n")
        sys.stdout.write(f"{write_op.sequence}\n")
        if args.isfastq:
            sys.stdout.write("+\n" + write_op.quality_sequence + "\n")
    else:
        output_manager.write_sequence(write_op, trace_logger)


def cleanup_empty_directories(output_dir: str):
    """reference: io_utils.py:521-568 -- prune directories holding nothing but primers files."""
    ignorable = {"primers.fasta", "primers.txt"}
    for root, dirs, files in os.walk(output_dir, topdown=False):
        if root == output_dir or os.path.basename(root) in ("trace",):
            continue
        real_files = [f for f in files if f not in ignorable]
        live_dirs = [d for d in dirs if os.path.exists(os.path.join(root, d))]
        if not real_files and not live_dirs:
            for f in files:
                os.remove(os.path.join(root, f))
            try:
                os.rmdir(root)
            except OSError:
                pass
