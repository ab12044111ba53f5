"""Seq / reverse_complement subset with Biopython's ambiguous-DNA complement table."""

_SRC = "ACGTMRWSYKVHDBXN"
_DST = "TGCAKYWSRMBDHVXN"
# Biopython also maps U/u -> A/a in its DNA complement table.
_TABLE = bytes.maketrans((_SRC + _SRC.lower() + "Uu").encode(), (_DST + _DST.lower() + "Aa").encode())


def complement(sequence):
    if isinstance(sequence, Seq):
        return Seq(str(sequence).translate(_STR_TABLE))
    return sequence.translate(_STR_TABLE)


_STR_TABLE = {k: v for k, v in zip(_SRC + _SRC.lower() + "Uu", _DST + _DST.lower() + "Aa")}
_STR_TABLE = str.maketrans(_STR_TABLE)


def reverse_complement(sequence):
    if isinstance(sequence, Seq):
        return Seq(str(sequence).translate(_STR_TABLE)[::-1])
    return str(sequence).translate(_STR_TABLE)[::-1]


class Seq:
    __slots__ = ("_data",)

    def __init__(self, data=""):
        self._data = str(data)

    def __str__(self):
        return self._data

    def __repr__(self):
        return "Seq(%r)" % self._data

    def __len__(self):
        return len(self._data)

    def __iter__(self):
        return iter(self._data)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return Seq(self._data[idx])
        return self._data[idx]

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash(self._data)

    def __add__(self, other):
        return Seq(self._data + str(other))

    def upper(self):
        return Seq(self._data.upper())

    def lower(self):
        return Seq(self._data.lower())

    def reverse_complement(self):
        return reverse_complement(self)

    def complement(self):
        return complement(self)
