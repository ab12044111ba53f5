"""ctypes binding for oracle/edlib_restated.c (TEST INFRASTRUCTURE ONLY).

`align()` mirrors the dict edlib.align returns for the arguments specimux uses
(reference: src/specimux/alignment.py:42, src/specimux/orchestration.py:552).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

# reference: src/specimux/constants.py:13-20 (IUPAC_EQUIV) -- restated as code -> bases
IUPAC_SETS = {"Y": "CT", "R": "AG", "N": "ACGT", "W": "AT", "M": "AC", "S": "CG",
              "K": "GT", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG"}
IUPAC_PAIRS = [(code, base) for code, bases in IUPAC_SETS.items() for base in bases]

_MODES = {"NW": 0, "SHW": 1, "HW": 2}
_NONE = -(2 ** 31)
_lib = None
_eq_cache = {}


def build():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "edlib_restated.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_align.restype = ctypes.c_int
    return _lib


_eq_by_id = {}


def bloom_yes(barcode, flank, k):
    """C version of oracle.pipeline._bloom_yes_py (same semantics, see edlib_restated.c)."""
    b = barcode.encode("latin-1")
    f = flank[:max(0, len(barcode) - k)].encode("latin-1", "replace")
    return bool(lib().orc_bloom_yes(b, len(b), f, len(f), int(k)))


def equality_table(pairs):
    hit = _eq_by_id.get(id(pairs))
    if hit is not None and hit[0] is pairs:
        return hit[1]
    tab = _equality_table(pairs)
    if pairs is not None:
        _eq_by_id[id(pairs)] = (pairs, tab)
    return tab


def _equality_table(pairs):
    key = tuple((str(a), str(b)) for a, b in pairs) if pairs else ()
    tab = _eq_cache.get(key)
    if tab is None:
        flat = bytes(ord(c) for ab in key for c in ab)
        tab = ctypes.create_string_buffer(65536)
        lib().orc_build_equality(flat, len(key), tab)
        _eq_cache[key] = tab
    return tab


def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None, cap=4096):
    q = query.encode("latin-1") if isinstance(query, str) else bytes(query)
    t = target.encode("latin-1") if isinstance(target, str) else bytes(target)
    dist = ctypes.c_int(0)
    nloc = ctypes.c_int(0)
    starts = (ctypes.c_int * cap)()
    ends = (ctypes.c_int * cap)()
    rc = lib().orc_align(q, len(q), t, len(t), _MODES[mode], int(k),
                         equality_table(additionalEqualities),
                         ctypes.byref(dist), ctypes.byref(nloc), starts, ends, cap)
    if rc != 0:
        raise MemoryError("orc_align failed")
    if nloc.value > cap:
        return align(query, target, mode, task, k, additionalEqualities, cap=nloc.value)
    locs = []
    if dist.value >= 0:
        for i in range(nloc.value):
            s = None if (starts[i] == _NONE or task == "distance") else starts[i]
            locs.append((s, ends[i]))
    return {"editDistance": dist.value, "alphabetLength": 0, "locations": locs, "cigar": None}


def align_py(query, target, mode="NW", k=-1, pairs=IUPAC_PAIRS):
    """Independent pure-Python DP (cross-check of the C restatement; small inputs only)."""
    eqs = set(pairs) | {(b, a) for a, b in pairs}
    E = lambda a, b: a == b or (a, b) in eqs
    m, n = len(query), len(target)
    if m == 0 or n == 0:
        if mode == "NW":
            return {"editDistance": max(m, n), "locations": [(None, n - 1)]}
        return {"editDistance": m, "locations": [(None, -1)]}

    def last_row(q, t, free):
        prev = [0 if free else j for j in range(len(t) + 1)]
        for i in range(1, len(q) + 1):
            cur = [i] + [0] * len(t)
            for j in range(1, len(t) + 1):
                cur[j] = min(prev[j - 1] + (0 if E(q[i - 1], t[j - 1]) else 1), prev[j] + 1, cur[j - 1] + 1)
            prev = cur
        return prev

    row = last_row(query, target, mode == "HW")
    if mode == "NW":
        d = row[n]
        if 0 <= k < d:
            return {"editDistance": -1, "locations": []}
        return {"editDistance": d, "locations": [(0, n - 1)]}
    j0 = 0 if m % 64 else 1
    best = min(row[j0:])
    if 0 <= k < best:
        return {"editDistance": -1, "locations": []}
    locs = []
    for j in range(j0, n + 1):
        if row[j] != best:
            continue
        e = j - 1
        s = 0
        if mode == "HW" and e >= 0:
            rrow = last_row(query[::-1], target[:e + 1][::-1], False)
            rb = min(rrow[1:])
            rpos = max(x - 1 for x in range(1, e + 2) if rrow[x] == rb)
            s = e - rpos
        locs.append((s, e))
    return {"editDistance": best, "locations": locs}
