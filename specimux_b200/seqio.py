"""Minimal FASTA/FASTQ reader and sequence record (Biopython is not a dependency here).

Mirrors what the reference takes from Bio (io_utils.py:429-450, demultiplex.py:142-144): record
`id` = first whitespace-delimited token of the title, `description` = whole title, Sanger
phred+33 qualities, and reverse_complement with the ambiguous-DNA complement table
(case preserved, U->A, unknown characters unchanged).
"""
import gzip
from typing import Iterator, List, Optional

_SRC = "ACGTMRWSYKVHDBXN"
_DST = "TGCAKYWSRMBDHVXN"
_COMPLEMENT = str.maketrans(_SRC + _SRC.lower() + "Uu", _DST + _DST.lower() + "Aa")


def reverse_complement(seq: str) -> str:
    return str(seq).translate(_COMPLEMENT)[::-1]


class SeqRecord:
    """Sequence record with the attribute names process_sequences consumes (id, description, seq,
    letter_annotations['phred_quality']).  `seq` is a plain str; `qual` keeps the raw quality
    string so no per-base Python list is ever built on the hot path."""
    __slots__ = ("id", "description", "seq", "qual")

    def __init__(self, seq: str, id: str = "<unknown id>", description: str = "<unknown description>",
                 qual: Optional[str] = None):
        self.seq = seq
        self.id = id
        self.description = description
        self.qual = qual

    name = property(lambda self: self.id)

    def __len__(self):
        return len(self.seq)

    @property
    def letter_annotations(self):
        if self.qual is None:
            return {}
        return {"phred_quality": [ord(c) - 33 for c in self.qual]}

    def reverse_complement(self):
        return SeqRecord(reverse_complement(self.seq), self.id, self.description,
                         None if self.qual is None else self.qual[::-1])


def _open_text(path):
    if str(path).endswith((".gz", ".gzip")):
        return gzip.open(path, "rt")
    return open(path, "rt")


def parse_fasta(handle) -> Iterator[SeqRecord]:
    title, chunks = None, []
    for line in handle:
        if line.startswith(">"):
            if title is not None:
                yield SeqRecord("".join(chunks), (title.split(None, 1) or [""])[0], title)
            title, chunks = line[1:].rstrip(), []
        elif title is not None:
            chunks.append(line.strip())
    if title is not None:
        yield SeqRecord("".join(chunks), (title.split(None, 1) or [""])[0], title)


def parse_fastq(handle) -> Iterator[SeqRecord]:
    readline = handle.readline
    while True:
        line = readline()
        if not line:
            return
        if not line.strip():
            continue
        if line[0] != "@":
            raise ValueError("Records in Fastq files should start with '@' character")
        title = line[1:].rstrip()
        seq = readline().strip()
        plus = readline()
        while plus and not plus.startswith("+"):       # multi-line sequence
            seq += plus.strip()
            plus = readline()
        qual = readline().strip()
        while len(qual) < len(seq):
            more = readline()
            if not more:
                break
            qual += more.strip()
        if len(qual) != len(seq):
            raise ValueError("Lengths of sequence and quality values differs for %s" % title)
        yield SeqRecord(seq, (title.split(None, 1) or [""])[0], title, qual)


def parse(path_or_handle, fmt: str) -> Iterator[SeqRecord]:
    own = isinstance(path_or_handle, (str, bytes)) or hasattr(path_or_handle, "__fspath__")
    handle = _open_text(path_or_handle) if own else path_or_handle
    try:
        if fmt == "fasta":
            yield from parse_fasta(handle)
        elif fmt == "fastq":
            yield from parse_fastq(handle)
        else:
            raise ValueError("unsupported sequence format %r" % fmt)
    finally:
        if own:
            handle.close()


def get_bases_and_quality(rec) -> "tuple[str, Optional[str]]":
    """(bases, ASCII quality or None) of a record of this module or of a Biopython-like SeqRecord."""
    bases = rec.seq if isinstance(rec.seq, str) else str(rec.seq)
    qual = getattr(rec, "qual", None)
    if qual is None:
        ann = getattr(rec, "letter_annotations", None)
        if ann and "phred_quality" in ann:
            qual = "".join(chr(q + 33) for q in ann["phred_quality"])
    return bases, qual


def records_from_tuples(reads: List[tuple]) -> List[SeqRecord]:
    return [SeqRecord(s, rid, rid, q) for rid, s, q in reads]
