"""Primer / specimen registries (mirror of the reference's src/specimux/databases.py interface).

Same method names and lookup semantics (including the by-sequence canonical primer registry,
SURVEY.md Q7); in this package they feed the device-resident match table (tables.py) instead of
being queried per read.
"""
import logging
from typing import Dict, List, Optional, Protocol, Set

from .constants import Primer
from .models import PrimerInfo


class PrimerDatabase:
    """reference: databases.py:17-120."""

    def __init__(self):
        self._primers: Dict[str, PrimerInfo] = {}
        self._pools: Dict[str, Set[str]] = {}
        self._pool_primers: Dict[str, Dict[Primer, List[PrimerInfo]]] = {}

    def add_primer(self, primer: PrimerInfo, pools: List[str]) -> None:
        if primer.name in self._primers:
            raise ValueError(f"Duplicate primer name: {primer.name}")
        self._primers[primer.name] = primer
        for pool in pools:
            if pool not in self._pools:
                self._pools[pool] = set()
                self._pool_primers[pool] = {Primer.FWD: [], Primer.REV: []}
            self._pools[pool].add(primer.name)
            self._pool_primers[pool][primer.direction].append(primer)

    def get_primer(self, name: str) -> Optional[PrimerInfo]:
        return self._primers.get(name)

    def get_primers_in_pool(self, pool: str) -> List[PrimerInfo]:
        if pool not in self._pools:
            return []
        return self._pool_primers[pool][Primer.FWD] + self._pool_primers[pool][Primer.REV]

    def get_pools(self) -> List[str]:
        return list(self._pools)

    def get_pool_primers(self, pool: str, direction: Optional[Primer] = None) -> List[PrimerInfo]:
        if pool not in self._pools:
            return []
        if direction:
            return self._pool_primers[pool][direction]
        return self.get_primers_in_pool(pool)

    def primer_in_pool(self, primer_name: str, pool: str) -> bool:
        return pool in self._pools and primer_name in self._pools[pool]

    def validate_pools(self) -> None:
        for pool, by_dir in self._pool_primers.items():
            if not by_dir[Primer.FWD]:
                raise ValueError(f"Pool {pool} has no forward primers")
            if not by_dir[Primer.REV]:
                raise ValueError(f"Pool {pool} has no reverse primers")

    def get_pool_stats(self) -> Dict:
        return {"total_primers": len(self._primers), "total_pools": len(self._pools),
                "pools": {pool: {"forward_primers": len(d[Primer.FWD]), "reverse_primers": len(d[Primer.REV]),
                                 "total_primers": len(d[Primer.FWD]) + len(d[Primer.REV])}
                          for pool, d in self._pool_primers.items()}}


class Specimens:
    """reference: databases.py:123-308."""

    def __init__(self, primer_registry: PrimerDatabase):
        self._specimens = []            # (id, pool, b1, [PrimerInfo], b2, [PrimerInfo]) in file order
        self._barcode_length = 0
        self._primers: Dict[str, PrimerInfo] = {}    # canonical primers keyed by SEQUENCE
        self._specimen_ids = set()
        self._primer_pairings: Dict[str, List[PrimerInfo]] = {}
        self._primer_registry = primer_registry
        self._active_pools = set()

    def add_specimen(self, specimen_id: str, pool: str, b1: str, p1: str, b2: str, p2: str):
        if specimen_id in self._specimen_ids:
            raise ValueError(f"Duplicate specimen id in index file: {specimen_id}")
        self._specimen_ids.add(specimen_id)
        self._active_pools.add(pool)
        self._barcode_length = max(self._barcode_length, len(b1), len(b2))
        p1_list = self._resolve_primer_name(p1, pool, Primer.FWD)
        p2_list = self._resolve_primer_name(p2, pool, Primer.REV)
        for plist, barcode in ((p1_list, b1), (p2_list, b2)):
            for info in plist:
                canon = self._primers.setdefault(info.primer, info)
                canon.barcodes[barcode] = None
                canon.specimens.add(specimen_id)
        self._specimens.append((specimen_id, pool, b1, p1_list, b2, p2_list))

    def prune_unused_pools(self):
        unused = set(self._primer_registry.get_pools()) - self._active_pools
        if not unused:
            return
        logging.info(f"Removing unused pools: {unused}")
        for primer in self._primers.values():
            primer.pools = [p for p in primer.pools if p in self._active_pools]
        for pool in unused:
            self._primer_registry._pools.pop(pool, None)
            self._primer_registry._pool_primers.pop(pool, None)
        stats = self._primer_registry.get_pool_stats()
        logging.info(f"After pruning: {stats['total_primers']} primers in {stats['total_pools']} pools")
        for pool, ps in stats["pools"].items():
            logging.info(f"Pool {pool}: {ps['forward_primers']} forward, {ps['reverse_primers']} reverse primers")

    def _resolve_primer_name(self, primer_name: str, pool: str, direction: Primer) -> List[PrimerInfo]:
        if primer_name in ("-", "*"):
            primers = [p for p in self._primer_registry.get_primers_in_pool(pool) if p.direction == direction]
            if not primers:
                raise ValueError(f"No {direction.name} primers found in pool {pool}")
            return primers
        primer = self._primer_registry.get_primer(primer_name)
        if not primer:
            raise ValueError(f"Primer not found: {primer_name}")
        if primer.direction != direction:
            raise ValueError(f"Primer {primer_name} is not a {direction.name} primer")
        if not self._primer_registry.primer_in_pool(primer_name, pool):
            raise ValueError(f"Primer {primer_name} is not in pool {pool}")
        return [primer]

    def specimens_for_barcodes_and_primers(self, b1_list, b2_list, p1_matched, p2_matched) -> List[str]:
        return [sid for sid, _pool, b1, p1s, b2, p2s in self._specimens
                if any(p1_matched is x for x in p1s) and any(p2_matched is x for x in p2s)
                and b1.upper() in b1_list and b2.upper() in b2_list]

    def specimen_for_exact_match(self, b1: str, b2: str, p1: PrimerInfo, p2: PrimerInfo) -> Optional[str]:
        for sid, _pool, sb1, p1s, sb2, p2s in self._specimens:
            if (any(p1 is x for x in p1s) and any(p2 is x for x in p2s)
                    and sb1.upper() == b1.upper() and sb2.upper() == b2.upper()):
                return sid
        return None

    def get_primers(self, direction: Primer) -> List[PrimerInfo]:
        return [p for p in self._primers.values() if p.direction == direction]

    def get_paired_primers(self, primer: str) -> List[PrimerInfo]:
        if primer not in self._primer_pairings:
            me = self._primers[primer]
            self._primer_pairings[primer] = [pi for pi in self._primers.values()
                                             if pi.direction != me.direction and pi.specimens & me.specimens]
        return self._primer_pairings[primer]

    def get_specimen_pool(self, specimen_id: str) -> Optional[str]:
        for row in self._specimens:
            if row[0] == specimen_id:
                return row[1]
        return None

    def b_length(self):
        return self._barcode_length

    def validate(self):
        self._validate_barcodes_globally_unique()
        self._validate_barcode_lengths()
        self.prune_unused_pools()

    def _barcode_sets(self):
        b1s, b2s = {}, {}
        for primer in self._primers.values():
            (b1s if primer.direction == Primer.FWD else b2s).update(primer.barcodes)
        return b1s, b2s

    def _validate_barcodes_globally_unique(self):
        b1s, b2s = self._barcode_sets()
        dups = set(b1s) & set(b2s)
        if dups:
            logging.warning(f"Duplicate Barcodes ({len(dups)}) in Fwd and Rev: {dups}")

    def _validate_barcode_lengths(self):
        b1s, b2s = self._barcode_sets()
        if len({len(b) for b in b1s}) > 1:
            logging.warning("Forward barcodes have inconsistent lengths")
        if len({len(b) for b in b2s}) > 1:
            logging.warning("Reverse barcodes have inconsistent lengths")


class BarcodePrefilter(Protocol):
    """reference: databases.py:311-316.  The GPU path searches every barcode exhaustively; a
    prefilter object only selects whether the Bloom filter's *observable* behaviour is emulated."""

    def match(self, barcode: str, sequence: str) -> bool:
        ...


class PassthroughPrefilter:
    """reference: databases.py:319-324 -- "no filtering"."""

    def match(self, barcode: str, sequence: str) -> bool:
        return True


class BloomEmulationPrefilter:
    """Marker for the default (--disable-prefilter not given): the device emulates the reference
    BloomPrefilter's result (bloom_filter.py:176-186) exactly, without any filter being built."""

    def match(self, barcode: str, sequence: str) -> bool:      # never called on the hot path
        raise NotImplementedError("the Bloom prefilter is emulated on the GPU")
