"""specimine on the GPU: recover partial-barcode reads of a specimen by aligning the specimen's full-match reads
into them (reference: src/specimux/specimine.py).

The reference loops `edlib.align(full_seq, partial_seq, mode="HW", task="path", k=int(len(full) * (1 - min_identity)))`
over every (partial, full) pair and reads only `editDistance` from the result (specimine.py:226-248).  Here the whole
|full| x |partial| distance matrix comes from one call of the C ABI (smx_hw_distances: warp-cooperative multi-word
Myers, hand-written CUDA); file discovery, the identity rule (first full read with the strictly highest identity
>= --min-identity wins) and the output naming follow the reference.  No CPU path: without a GPU the call raises.
"""
import argparse
import glob
import logging
import os
import re
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from . import seqio

# at most this many int32 distances per GPU call (the matrix is cut along the partial reads)
_MATRIX_BUDGET = 32 << 20


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser(description="Mine additional candidate sequences from partial matches.")
    ap.add_argument("--index", required=True, help="Path to specimen index file (same as used with specimux)")
    ap.add_argument("--fastq", required=True, help="Path to full match FASTQ file for a specimen")
    ap.add_argument("--partial-forward", action="store_true", default=False,
                    help="Include forward partial matches (default: False)")
    ap.add_argument("--no-partial-reverse", action="store_true", default=False,
                    help="Exclude reverse partial matches (included by default)")
    ap.add_argument("--min-identity", type=float, default=0.85, help="Minimum alignment identity for a match (default: 0.85)")
    ap.add_argument("--debug", action="store_true", help="Enable debug logging")
    return ap.parse_args(argv)


def extract_specimen_id(fastq_path: str) -> str:
    """specimine.py:54-62: the file name minus `.fastq` (and minus the legacy `sample_` prefix)."""
    name = os.path.basename(fastq_path)
    m = re.match(r"(?:sample_)?(.+)\.fastq", name)
    if not m:
        raise ValueError(f"Could not extract specimen ID from filename: {name}")
    return m.group(1)


def find_barcodes(specimen_id: str, index_file: str) -> Tuple[Optional[str], Optional[str]]:
    """specimine.py:65-83: (FwIndex, RvIndex) of the specimen's row, upper-cased."""
    with open(index_file, "r") as fh:
        header = next(fh).strip().split("\t")
        col = lambda name, default: header.index(name) if name in header else default
        i_id, i_fw, i_rv = col("SampleID", 0), col("FwIndex", 2), col("RvIndex", 4)
        for line in fh:
            fields = line.strip().split("\t")
            if len(fields) > max(i_id, i_fw, i_rv) and fields[i_id] == specimen_id:
                return fields[i_fw].upper(), fields[i_rv].upper()
    logging.error(f"Could not find specimen {specimen_id} in index file")
    return None, None


def detect_input_level(fastq_path: str) -> Tuple[str, str, Optional[str]]:
    """specimine.py:86-117: (output root, pool, primer pair or None) from .../full/<pool>[/<pair>]/<specimen>.fastq."""
    parts = os.path.abspath(fastq_path).split(os.sep)
    if "full" not in parts:
        raise ValueError(f"Could not find 'full' directory in path: {fastq_path}")
    at = parts.index("full")
    below = parts[at + 1:-1]
    if len(below) not in (1, 2):
        raise ValueError(f"Unexpected path structure: {fastq_path}")
    return os.sep.join(parts[:at]), below[0], (below[1] if len(below) == 2 else None)


def _partials_in(directory: str, fwd_barcode: str, rev_barcode: str, found: Dict[str, List[str]]) -> None:
    """specimine.py:120-150: barcode_fwd_<b1>.fastq / barcode_rev_<b2>.fastq (or the legacy sample_ names)."""
    if not os.path.isdir(directory):
        return
    for kind, tag, barcode in (("forward", "fwd", fwd_barcode), ("reverse", "rev", rev_barcode)):
        if not barcode:
            continue
        for name in (f"barcode_{tag}_{barcode}.fastq", f"sample_barcode_{tag}_{barcode}.fastq"):
            path = os.path.join(directory, name)
            if os.path.exists(path):
                found[kind].append(path)
                break


def derive_partial_match_filenames(fastq_path: str, fwd_barcode: str, rev_barcode: str) -> Dict[str, List[str]]:
    """specimine.py:153-194: the partial files next to a pool-level or primer-pair-level full-match file."""
    found: Dict[str, List[str]] = {"forward": [], "reverse": []}
    root, pool, pair = detect_input_level(fastq_path)
    if pair is not None:
        _partials_in(os.path.join(root, "partial", pool, pair), fwd_barcode, rev_barcode, found)
    else:
        pool_dir = os.path.join(root, "partial", pool)
        if os.path.isdir(pool_dir):
            for d in glob.glob(os.path.join(pool_dir, "*")):
                if os.path.isdir(d):
                    _partials_in(d, fwd_barcode, rev_barcode, found)
    if fwd_barcode and not found["forward"]:
        logging.warning(f"No forward partial match files found for barcode: {fwd_barcode}")
    if rev_barcode and not found["reverse"]:
        logging.warning(f"No reverse partial match files found for barcode: {rev_barcode}")
    return {k: v for k, v in found.items() if v}


def _blob(seqs: List[str]):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    return "".join(seqs).encode("latin-1", "replace"), off


def hw_distances(patterns: List[str], max_dist: List[int], texts: List[str], device: int = 0) -> np.ndarray:
    """[len(patterns), len(texts)] int32: HW edit distance of pattern i in text j, -1 when above max_dist[i].
    One smx_hw_distances call per slab of texts (GPU only)."""
    lib = _lib.load()
    out = np.empty((len(patterns), len(texts)), dtype=np.int32)
    if not patterns or not texts:
        return out
    pblob, poff = _blob(patterns)
    pk = np.asarray(max_dist, dtype=np.int32)
    per = max(1, _MATRIX_BUDGET // max(1, len(patterns)))
    for lo in range(0, len(texts), per):
        hi = min(len(texts), lo + per)
        tblob, toff = _blob(texts[lo:hi])
        slab = np.empty((len(patterns), hi - lo), dtype=np.int32)
        _lib.check(lib.smx_hw_distances(device, pblob, _lib.ptr(poff, _lib.u64p), _lib.ptr(pk, _lib.i32p), len(patterns),
                                        tblob, _lib.ptr(toff, _lib.u64p), hi - lo, _lib.ptr(slab, _lib.i32p)))
        out[:, lo:hi] = slab
    return out


def mine_sequences(full_match_file: str, partial_match_files: Dict[str, List[str]], min_identity: float, device: int = 0):
    """specimine.py:205-268.  Returns the mined partial records (id / description rewritten) in the reference's order."""
    full = list(seqio.parse(full_match_file, "fastq"))
    if not full:
        logging.error(f"No sequences found in full match file: {full_match_file}")
        return []
    logging.info(f"Loaded {len(full)} sequences from full match file")
    full_seqs = [str(r.seq) for r in full]
    full_len = np.array([len(s) for s in full_seqs], dtype=np.float64)
    max_dist = [int(len(s) * (1 - min_identity)) for s in full_seqs]          # specimine.py:238
    mined = []
    for kind, files in partial_match_files.items():
        logging.info(f"Processing {kind} partial matches from {len(files)} file(s)")
        partial = []
        for path in files:
            logging.debug(f"  Loading: {path}")
            partial.extend(seqio.parse(path, "fastq"))
        logging.info(f"Found {len(partial)} sequences across all {kind} partial match files")
        dist = hw_distances(full_seqs, max_dist, [str(r.seq) for r in partial], device)
        # identity = 1 - d / len(full) where the alignment exists (specimine.py:197-202, :243-244); the first full read
        # with the strictly highest identity >= min_identity (and > 0, the initial best) wins (:246-248)
        with np.errstate(divide="ignore", invalid="ignore"):
            ident = 1 - (dist / full_len[:, None])
        ok = (dist != -1) & (ident >= min_identity) & (ident > 0)
        ident = np.where(ok, ident, -np.inf)
        n_hit = 0
        for j, rec in enumerate(partial):
            col = ident[:, j] if len(partial) else ident[:, :0]
            i = int(np.argmax(col))
            if col[i] == -np.inf:
                continue
            best = float(col[i])
            n_hit += 1
            rec.id = f"{rec.id}_mined_{kind}_{best:.2f}"
            rec.description = f"{rec.description} mined_{kind} identity={best:.2f}"
            mined.append(rec)
        logging.info(f"Matched {n_hit}/{len(partial)} sequences from {kind} partial matches")
    return mined


def fastq_title(rec) -> str:
    """Title line Biopython's FASTQ writer gives a record (Bio.SeqIO.QualityIO.FastqPhredWriter.write_record): the
    description when it already starts with the id, else "<id> <description>"; newlines become spaces."""
    clean = lambda s: s.replace("\n", " ").replace("\r", " ")
    ident = clean(rec.id) if rec.id else ""
    desc = clean(rec.description or "")
    if desc and desc.split(None, 1)[0] == ident:
        return desc
    return f"{ident} {desc}" if desc else ident


def write_fastq(records, path: str) -> int:
    with open(path, "w") as fh:
        for rec in records:
            qtext = rec.qual if getattr(rec, "qual", None) is not None else "".join(
                chr(q + 33) for q in rec.letter_annotations["phred_quality"])
            fh.write("@%s\n%s\n+\n%s\n" % (fastq_title(rec), str(rec.seq), qtext))
    return len(records)


def main(argv=None):
    args = parse_arguments(argv)
    logging.basicConfig(level=logging.DEBUG if args.debug else logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    specimen = extract_specimen_id(args.fastq)
    logging.info(f"Processing specimen: {specimen}")
    fwd, rev = find_barcodes(specimen, args.index)
    if not (fwd and rev):
        sys.exit(1)
    logging.info(f"Found barcodes - Forward: {fwd}, Reverse: {rev}")
    files = derive_partial_match_filenames(args.fastq, fwd, rev)
    for kind, paths in files.items():
        logging.info(f"Found {len(paths)} {kind} partial match file(s)")
        for p in paths:
            logging.debug(f"  - {p}")
    if not args.partial_forward:
        files.pop("forward", None)
    if args.no_partial_reverse:
        files.pop("reverse", None)
    if not files:
        logging.error("No partial match files found or selected")
        sys.exit(1)
    mined = mine_sequences(args.fastq, files, args.min_identity)
    logging.info(f"Found {len(mined)} mined sequences")
    out = f"{args.fastq}.mined"
    write_fastq(mined, out)
    logging.info(f"Wrote {len(mined)} sequences to {out}")


if __name__ == "__main__":
    main()
