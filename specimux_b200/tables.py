"""Specimens + MatchParameters + flags -> the smx_tables / smx_params the device consumes.

This is where the reference's per-read Python lookups become a device-resident match table:
canonical primers in iteration order (databases.py:247-249), per-primer barcode lists in pinned
order (models.py:28), primer pairs in candidate order (demultiplex.py:699-700) with
their pool (demultiplex.py:640-665), specimen rows with primer identity masks
(databases.py:219-245).
"""
import ctypes as C
from typing import List, Optional

import numpy as np

from . import _lib
from .constants import Primer
from .seqio import reverse_complement


def pool_from_primers(p1, p2) -> Optional[str]:
    """reference: demultiplex.py:640-665 get_pool_from_primers."""
    if p1 and p2:
        common = set(p1.pools) & set(p2.pools)
        return sorted(common)[0] if common else None
    one = p1 or p2
    if one:
        return sorted(one.pools)[0] if one.pools else None
    return None


def _concat(strings: List[str]):
    off = np.zeros(len(strings) + 1, dtype=np.uint32)
    np.cumsum([len(s) for s in strings], out=off[1:])
    return "".join(strings).encode("ascii"), off


class MatchTables:
    """Host-side image of the device tables plus the id <-> name maps used to decode records."""

    def __init__(self, specimens, parameters, trim="barcodes", dereplicate="best", prefilter=True,
                 min_length=-1, max_length=-1):
        self.specimens = specimens
        self.parameters = parameters
        primers = list(specimens._primers.values())
        self.primers = primers
        index = {id(p): i for i, p in enumerate(primers)}
        self.primer_names = [p.name for p in primers]

        # barcode namespaces, first appearance over primers in iteration order
        self.b1: List[str] = []
        self.b2: List[str] = []
        b1_id, b2_id = {}, {}
        pb_off = [0]
        pb_barcode: List[int] = []
        for p in primers:
            space, ids = (self.b1, b1_id) if p.direction == Primer.FWD else (self.b2, b2_id)
            for bc in p.barcodes:
                if bc not in ids:
                    ids[bc] = len(space)
                    space.append(bc)
                pb_barcode.append(ids[bc])
            pb_off.append(len(pb_barcode))

        # pools: registry pools, then any specimen-declared pool not in the registry
        self.pools: List[str] = []
        pool_id = {}

        def pid(name):
            if name is None:
                return -1
            if name not in pool_id:
                pool_id[name] = len(self.pools)
                self.pools.append(name)
            return pool_id[name]

        pairs = []
        for fwd in specimens.get_primers(Primer.FWD):
            for rev in specimens.get_paired_primers(fwd.primer):
                pairs.append((index[id(fwd)], index[id(rev)], pid(pool_from_primers(fwd, rev))))
        self.pairs = pairs

        rows = specimens._specimens
        self.specimen_ids = [r[0] for r in rows]
        self.specimen_pools = [r[1] for r in rows]

        mask_words = (len(primers) + 63) // 64

        def mask(plist):
            words = [0] * mask_words
            for p in plist:
                i = index.get(id(p))
                if i is not None:
                    words[i >> 6] |= 1 << (i & 63)
            return words

        # keep every array alive for the lifetime of the object (ctypes only borrows pointers)
        k = self._keep = {}
        k["primer_seq"], k["primer_off"] = _concat([p.primer for p in primers])
        k["primer_rc"], _ = _concat([p.primer_rc for p in primers])
        k["primer_dir"] = np.array([0 if p.direction == Primer.FWD else 1 for p in primers], dtype=np.uint8)
        k["primer_k"] = np.array([parameters.max_dist_primers[p.primer] for p in primers], dtype=np.int32)
        k["primer_fidx"] = np.array([p.file_index for p in primers], dtype=np.int32)
        k["pb_off"] = np.array(pb_off, dtype=np.uint32)
        k["pb_barcode"] = np.array(pb_barcode if pb_barcode else [0], dtype=np.uint32)
        k["b1_rc"], k["b1_off"] = _concat([reverse_complement(b) for b in self.b1])
        k["b2_rc"], k["b2_off"] = _concat([reverse_complement(b) for b in self.b2])
        k["pair_fwd"] = np.array([p[0] for p in pairs] or [0], dtype=np.uint32)
        k["pair_rev"] = np.array([p[1] for p in pairs] or [0], dtype=np.uint32)
        k["pair_pool"] = np.array([p[2] for p in pairs] or [0], dtype=np.int32)
        k["spec_b1"] = np.array([b1_id[r[2].upper()] if r[2].upper() in b1_id else b1_id[r[2]] for r in rows], dtype=np.uint32)
        k["spec_b2"] = np.array([b2_id[r[4].upper()] if r[4].upper() in b2_id else b2_id[r[4]] for r in rows], dtype=np.uint32)
        k["spec_p1"] = np.array([mask(r[3]) for r in rows] or [[0] * mask_words], dtype=np.uint64).reshape(-1)
        k["spec_p2"] = np.array([mask(r[5]) for r in rows] or [[0] * mask_words], dtype=np.uint64).reshape(-1)
        k["spec_pool"] = np.array([pid(r[1]) for r in rows], dtype=np.int32)

        t = _lib.SmxTables()
        t.n_primers = len(primers)
        t.primer_seq, t.primer_rc = k["primer_seq"], k["primer_rc"]
        t.primer_off = _lib.ptr(k["primer_off"], _lib.u32p)
        t.primer_dir = _lib.ptr(k["primer_dir"], _lib.u8p)
        t.primer_k = _lib.ptr(k["primer_k"], _lib.i32p)
        t.primer_file_index = _lib.ptr(k["primer_fidx"], _lib.i32p)
        t.pb_off = _lib.ptr(k["pb_off"], _lib.u32p)
        t.pb_barcode = _lib.ptr(k["pb_barcode"], _lib.u32p)
        t.n_b1, t.n_b2 = len(self.b1), len(self.b2)
        t.b1_rc, t.b2_rc = k["b1_rc"], k["b2_rc"]
        t.b1_off = _lib.ptr(k["b1_off"], _lib.u32p)
        t.b2_off = _lib.ptr(k["b2_off"], _lib.u32p)
        t.n_pairs = len(pairs)
        t.pair_fwd = _lib.ptr(k["pair_fwd"], _lib.u32p)
        t.pair_rev = _lib.ptr(k["pair_rev"], _lib.u32p)
        t.pair_pool = _lib.ptr(k["pair_pool"], _lib.i32p)
        t.n_specimens = len(rows)
        t.spec_b1 = _lib.ptr(k["spec_b1"], _lib.u32p)
        t.spec_b2 = _lib.ptr(k["spec_b2"], _lib.u32p)
        t.spec_p1_mask = _lib.ptr(k["spec_p1"], _lib.u64p)
        t.spec_p2_mask = _lib.ptr(k["spec_p2"], _lib.u64p)
        t.spec_pool = _lib.ptr(k["spec_pool"], _lib.i32p)
        self.c_tables = t

        pr = _lib.SmxParams()
        pr.search_len = parameters.search_len
        pr.max_dist_index = parameters.max_dist_index
        pr.barcode_length = specimens.b_length()
        pr.preorient = 1 if parameters.preorient else 0
        pr.prefilter = 1 if prefilter else 0
        pr.trim = _lib.TRIM_CODES[trim]
        pr.dereplicate_best = 1 if dereplicate == "best" else 0
        pr.min_length = min_length
        pr.max_length = max_length
        self.c_params = pr
        self.n_primers = len(primers)
        self.pb_off = pb_off
        self.total_barcode_slots = 2 * len(pb_barcode)
        self.mask_words = (parameters.search_len + 31) // 32

    def tables_ref(self):
        return C.byref(self.c_tables)

    def params_ref(self):
        return C.byref(self.c_params)
