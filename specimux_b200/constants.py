"""Constants and enumerations (mirror of the reference's src/specimux/constants.py values)."""
from enum import Enum

# code -> base set; the reference lists the same relation as (code, base) pairs (constants.py:13-20)
IUPAC_BASES = {"Y": "CT", "R": "AG", "N": "ACGT", "W": "AT", "M": "AC", "S": "CG", "K": "GT",
               "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG"}
IUPAC_EQUIV = [(code, base) for code, bases in IUPAC_BASES.items() for base in bases]
IUPAC_CODES = set(IUPAC_BASES) | set("ACGT")


class AlignMode:
    GLOBAL = "NW"
    INFIX = "HW"
    PREFIX = "SHW"


class SampleId:
    UNKNOWN = "unknown"
    PREFIX_FWD_MATCH = "barcode_fwd_"
    PREFIX_REV_MATCH = "barcode_rev_"


class TrimMode:
    PRIMERS = "primers"
    BARCODES = "barcodes"
    TAILS = "tails"
    NONE = "none"


class MultipleMatchStrategy:
    NONE = "none"
    BEST = "best"


class ResolutionType(Enum):
    FULL_MATCH = 1
    PARTIAL_FORWARD = 2
    PARTIAL_REVERSE = 3
    MULTIPLE_SPECIMENS = 4
    UNKNOWN = 5
    DEREPLICATED_FULL = 6

    def to_string(self) -> str:
        return {1: "full_match", 2: "partial_forward", 3: "partial_reverse", 4: "multiple_specimens",
                6: "dereplicated_full"}.get(self.value, "unknown")

    def is_full_match(self) -> bool:
        return self in (ResolutionType.FULL_MATCH, ResolutionType.DEREPLICATED_FULL)

    def is_partial_match(self) -> bool:
        return self in (ResolutionType.PARTIAL_FORWARD, ResolutionType.PARTIAL_REVERSE)

    def is_unknown(self) -> bool:
        return self is ResolutionType.UNKNOWN


class Barcode(Enum):
    B1 = 1
    B2 = 2

    def to_string(self) -> str:
        return "forward" if self is Barcode.B1 else "reverse"


class Primer(Enum):
    FWD = 3
    REV = 4

    def to_string(self) -> str:
        return "forward" if self is Primer.FWD else "reverse"


class Orientation(Enum):
    FORWARD = 1
    REVERSE = 2
    UNKNOWN = 3

    def to_string(self) -> str:
        return self.name.lower()
