"""SeqIO.parse subset: FASTA and 4-line/multi-line FASTQ, Biopython id/description rules."""
from .Seq import Seq
from .SeqRecord import SeqRecord


def _open(handle):
    if isinstance(handle, (str, bytes)) or hasattr(handle, "__fspath__"):
        return open(handle, "rt"), True
    return handle, False


def _title_fields(title):
    parts = title.split(None, 1)
    ident = parts[0] if parts else ""
    return ident, title


def _parse_fasta(fh):
    title, chunks = None, []
    for line in fh:
        if line.startswith(">"):
            if title is not None:
                ident, desc = _title_fields(title)
                yield SeqRecord(Seq("".join(chunks)), id=ident, name=ident, description=desc)
            title, chunks = line[1:].rstrip(), []
        elif title is not None:
            chunks.append(line.strip())
    if title is not None:
        ident, desc = _title_fields(title)
        yield SeqRecord(Seq("".join(chunks)), id=ident, name=ident, description=desc)


def _parse_fastq(fh):
    while True:
        line = fh.readline()
        if not line:
            return
        if not line.strip():
            continue
        if not line.startswith("@"):
            raise ValueError("Records in Fastq files should start with '@' character")
        title = line[1:].rstrip()
        seq_chunks = []
        line = fh.readline()
        while line and not line.startswith("+"):
            seq_chunks.append(line.strip())
            line = fh.readline()
        seq = "".join(seq_chunks)
        qual = ""
        while len(qual) < len(seq):
            line = fh.readline()
            if not line:
                break
            qual += line.strip()
        if len(qual) != len(seq):
            raise ValueError("Lengths of sequence and quality values differs for %s" % title)
        ident, desc = _title_fields(title)
        rec = SeqRecord(Seq(seq), id=ident, name=ident, description=desc)
        rec.letter_annotations["phred_quality"] = [ord(c) - 33 for c in qual]
        yield rec


def parse(handle, fmt):
    fh, owned = _open(handle)
    try:
        if fmt == "fasta":
            yield from _parse_fasta(fh)
        elif fmt == "fastq":
            yield from _parse_fastq(fh)
        else:
            raise ValueError("unsupported format %r" % fmt)
    finally:
        if owned:
            fh.close()


def write(records, handle, fmt):
    fh, owned = (open(handle, "wt"), True) if isinstance(handle, str) else (handle, False)
    n = 0
    for r in records:
        fh.write(r.format(fmt))
        n += 1
    if owned:
        fh.close()
    return n
