"""SeqRecord subset (id, description, seq, letter_annotations, reverse_complement, slicing)."""
from .Seq import Seq


class SeqRecord:
    def __init__(self, seq, id="<unknown id>", name="<unknown name>", description="<unknown description>",
                 letter_annotations=None):
        self.seq = seq if isinstance(seq, Seq) else Seq(seq)
        self.id = id
        self.name = name
        self.description = description
        self.letter_annotations = letter_annotations if letter_annotations is not None else {}

    def __len__(self):
        return len(self.seq)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            ann = {k: v[idx] for k, v in self.letter_annotations.items()}
            return SeqRecord(self.seq[idx], self.id, self.name, self.description, ann)
        return self.seq[idx]

    def reverse_complement(self, id=False, name=False, description=False):
        ann = {k: v[::-1] for k, v in self.letter_annotations.items()}
        return SeqRecord(self.seq.reverse_complement(),
                         self.id if id is True else (id or "<unknown id>"),
                         self.name if name is True else (name or "<unknown name>"),
                         self.description if description is True else (description or "<unknown description>"),
                         ann)

    def format(self, fmt):
        if fmt == "fasta":
            return ">%s\n%s\n" % (self.description, str(self.seq))
        if fmt == "fastq":
            # Bio.SeqIO.QualityIO.FastqPhredWriter.write_record: the description when it already starts with the id,
            # else "<id> <description>"
            clean = lambda s: s.replace("\n", " ").replace("\r", " ")
            ident = clean(self.id) if self.id else ""
            desc = clean(self.description or "")
            if desc and desc.split(None, 1)[0] == ident:
                title = desc
            else:
                title = "%s %s" % (ident, desc) if desc else ident
            q = "".join(chr(x + 33) for x in self.letter_annotations["phred_quality"])
            return "@%s\n%s\n+\n%s\n" % (title, str(self.seq), q)
        raise ValueError(fmt)
