"""Randomised tables and reads (mixed barcode lengths, IUPAC-degenerate primers, shared primers,
wildcards, every k from 0 to 5, N / IUPAC / lower-case symbols, reads shorter than the search
window): CUDA per-thread routines (via the CPU kernel simulator) against the live oracle."""
import random

import pytest

import helpers as H
from oracle import pipeline as orc
from specimux_b200.demultiplex import process_sequences
from specimux_b200.models import MatchParameters
from specimux_b200.seqio import reverse_complement

IUPAC = {"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT",
         "V": "ACG", "N": "ACGT"}


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, rate):
    out = []
    for c in s:
        u = rng.random()
        if u < rate * 0.4:
            out.append(rng.choice("ACGT"))
        elif u < rate * 0.7:
            out.append(rng.choice("ACGT") + c)
        elif u < rate:
            continue
        else:
            out.append(c)
    return "".join(out)


def instantiate(rng, s):
    return "".join(rng.choice(IUPAC[c]) if c in IUPAC else c for c in s)


def make_case(seed, long_primers=False):
    """long_primers: some primers of 65..300 nt (the warp-cooperative multi-word search) and search
    windows wide enough to hold them."""
    rng = random.Random(seed)
    n_fwd, n_rev = rng.randint(1, 3), rng.randint(1, 3)
    pools = ["P%d" % i for i in range(rng.randint(1, 2))]
    primers = []
    for i in range(n_fwd + n_rev):
        seq = rand_seq(rng, rng.randint(16, 26))
        if long_primers and (i == 0 or rng.random() < 0.5):
            seq = rand_seq(rng, rng.choice([65, 66, 96, 97, 127, 128, 129, 150, 200, 256, 257, 300]))
        if rng.random() < 0.5:
            seq = "".join(rng.choice("RYNWM") if rng.random() < 0.12 else c for c in seq)
        ppools = sorted(rng.sample(pools, rng.randint(1, len(pools))))
        primers.append(("F%d" % i if i < n_fwd else "R%d" % i, seq, "forward" if i < n_fwd else "reverse", ppools))
    for pool in pools:      # every pool needs both directions
        for direction in ("forward", "reverse"):
            if not any(pool in p[3] and p[2] == direction for p in primers):
                idx = next(i for i, p in enumerate(primers) if p[2] == direction)
                primers[idx] = (primers[idx][0], primers[idx][1], direction, sorted(set(primers[idx][3]) | {pool}))
    lens = [rng.choice([9, 11, 13, 13, 13, 16]) for _ in range(2)]
    mixed = rng.random() < 0.4
    nb1, nb2 = rng.randint(2, 40), rng.randint(2, 70)

    def barcodes(n, ln):
        out = set()
        while len(out) < n:
            out.add(rand_seq(rng, rng.choice([ln, ln - 2, ln + 1]) if mixed else ln))
        return sorted(out, key=lambda _: rng.random())
    b1s, b2s = barcodes(nb1, lens[0]), barcodes(nb2, lens[1])
    specimens = []
    seen = set()
    for i in range(rng.randint(3, 120)):
        pool = rng.choice(pools)
        fw = [p for p in primers if p[2] == "forward" and pool in p[3]]
        rv = [p for p in primers if p[2] == "reverse" and pool in p[3]]
        p1 = "*" if rng.random() < 0.15 else rng.choice(fw)[0]
        p2 = "-" if rng.random() < 0.1 else rng.choice(rv)[0]
        b1, b2 = rng.choice(b1s), rng.choice(b2s)
        if (pool, b1, b2, p1, p2) in seen:
            continue
        seen.add((pool, b1, b2, p1, p2))
        specimens.append(("S%03d" % i, pool, b1, p1, b2, p2))
    k_idx = rng.choice([0, 1, 2, 3, 3, 4, 5])
    k_idx = min(k_idx, min(len(b) for b in b1s + b2s) - 1)
    L = rng.choice([40, 80, 80, 120, 200])
    if long_primers:
        L = rng.choice([120, 200, 333, 512])
    reads = []
    by_name = {p[0]: p for p in primers}
    for r in range(rng.randint(40, 90)):
        sp = rng.choice(specimens)
        fw = [p for p in primers if p[2] == "forward" and sp[1] in p[3]]
        rv = [p for p in primers if p[2] == "reverse" and sp[1] in p[3]]
        p1 = by_name[sp[3]] if sp[3] in by_name else rng.choice(fw)
        p2 = by_name[sp[5]] if sp[5] in by_name else rng.choice(rv)
        kind = rng.random()
        insert = rand_seq(rng, rng.randint(0, 300))
        left = rand_seq(rng, rng.randint(0, 25)) + sp[2] + instantiate(rng, p1[1])
        right = reverse_complement(instantiate(rng, p2[1])) + reverse_complement(sp[4]) + rand_seq(rng, rng.randint(0, 25))
        if kind < 0.1:
            s = insert
        elif kind < 0.2:
            s = left + insert
        elif kind < 0.3:
            s = insert + right
        else:
            s = left + insert + right
        s = mutate(rng, s, rng.choice([0.0, 0.05, 0.12]))
        if rng.random() < 0.5:
            s = reverse_complement(s)
        u = rng.random()
        if u < 0.1:
            s = "".join("N" if rng.random() < 0.05 else c for c in s)
        elif u < 0.15:
            s = "".join(rng.choice("RYKMSWBDHV") if rng.random() < 0.05 else c for c in s)
        elif u < 0.2:
            s = s[:30].lower() + s[30:]
        elif u < 0.3:
            s = s[:rng.randint(0, L + 3)]
        elif u < 0.35:
            s = s[len(s) - min(len(s), rng.randint(0, L + 3)):]
        reads.append(("r%04d" % r, s, "".join(chr(33 + rng.randint(2, 40)) for _ in s)))
    flags = dict(search_len=L, trim=rng.choice(["barcodes", "primers", "tails", "none"]),
                 dereplicate=rng.choice(["best", "none"]), disable_preorient=rng.random() < 0.3,
                 disable_prefilter=rng.random() < 0.5)
    return primers, specimens, reads, k_idx, flags


def make_iupac_barcode_case(seed):
    """Barcodes that hold IUPAC codes, searched with the Bloom prefilter ON (the default flags): the reference builds
    the filter's variants from the barcode STRING (bloom_filter.py:70-101), so an N in a barcode only passes the
    filter where the read has a literal N or an edit is spent on it, while edlib itself matches it for free.
    Uniform barcode length (the reference takes the filter's key length from one barcode)."""
    rng = random.Random(7000 + seed)
    primers = [("F0", rand_seq(rng, 20), "forward", ["P0"]), ("R0", rand_seq(rng, 22), "reverse", ["P0"])]
    if rng.random() < 0.5:
        primers.append(("F1", rand_seq(rng, 18), "forward", ["P0"]))
    blen = rng.choice([11, 13, 13])

    def barcode():
        s = list(rand_seq(rng, blen))
        if rng.random() < 0.6:
            for _ in range(rng.randint(1, 2)):
                i = rng.randrange(blen)
                s[i] = rng.choice([c for c, v in IUPAC.items() if s[i] in v or s[i] == c])
        return "".join(s)
    b1s = sorted({barcode() for _ in range(rng.randint(3, 20))})
    b2s = sorted({barcode() for _ in range(rng.randint(3, 40))})
    specimens, seen = [], set()
    for i in range(rng.randint(5, 60)):
        p1 = rng.choice([p[0] for p in primers if p[2] == "forward"])
        b1, b2 = rng.choice(b1s), rng.choice(b2s)
        if (b1, b2, p1) in seen:
            continue
        seen.add((b1, b2, p1))
        specimens.append(("S%03d" % i, "P0", b1, p1, b2, "R0"))
    by_name = {p[0]: p for p in primers}
    reads = []
    for r in range(70):
        sp = rng.choice(specimens)

        def concrete(bc):           # what a read carries where the barcode has a code: a base of the set, or a literal N
            return "".join((("N" if rng.random() < 0.3 else rng.choice(IUPAC[c])) if c in IUPAC else c) for c in bc)
        s = (rand_seq(rng, rng.randint(0, 20)) + concrete(sp[2]) + by_name[sp[3]][1] + rand_seq(rng, rng.randint(30, 200)) +
             reverse_complement(by_name[sp[5]][1]) + reverse_complement(concrete(sp[4])) + rand_seq(rng, rng.randint(0, 20)))
        s = mutate(rng, s, rng.choice([0.0, 0.03, 0.08]))
        if rng.random() < 0.5:
            s = reverse_complement(s)
        reads.append(("q%04d" % r, s, "I" * len(s)))
    flags = dict(search_len=80, trim=rng.choice(["barcodes", "tails"]), dereplicate=rng.choice(["best", "none"]),
                 disable_preorient=False, disable_prefilter=False)
    return primers, specimens, reads, rng.choice([1, 2, 3]), flags


@pytest.mark.parametrize("seed", range(12))
def test_iupac_barcodes_with_prefilter(seed):
    run_case(*make_iupac_barcode_case(seed), tag="iupac barcode seed %d" % seed)


def make_large_k_case(seed):
    """Long barcodes searched with thresholds beyond 8 (the reference derives k = ceil(min pairwise distance / 2),
    orchestration.py:583-600): the general band walk of stage 2 with up to 25 diagonals."""
    rng = random.Random(9000 + seed)
    primers = [("F0", rand_seq(rng, 21), "forward", ["P0"]), ("R0", rand_seq(rng, 19), "reverse", ["P0"])]
    blen = rng.choice([24, 28, 30])
    b1s = sorted({rand_seq(rng, blen) for _ in range(rng.randint(2, 12))})
    b2s = sorted({rand_seq(rng, blen) for _ in range(rng.randint(2, 40))})
    specimens = [("S%03d" % i, "P0", rng.choice(b1s), "F0", rng.choice(b2s), "R0") for i in range(25)]
    specimens = list({(s[2], s[4]): s for s in specimens}.values())
    reads = []
    for r in range(50):
        sp = rng.choice(specimens)
        s = (rand_seq(rng, rng.randint(0, 20)) + sp[2] + primers[0][1] + rand_seq(rng, rng.randint(30, 200)) +
             reverse_complement(primers[1][1]) + reverse_complement(sp[4]) + rand_seq(rng, rng.randint(0, 20)))
        s = mutate(rng, s, rng.choice([0.05, 0.15, 0.3]))
        if rng.random() < 0.5:
            s = reverse_complement(s)
        if rng.random() < 0.1:
            s = s[:40] + "N" + s[41:]
        reads.append(("k%04d" % r, s, "I" * len(s)))
    flags = dict(search_len=rng.choice([80, 120]), trim=rng.choice(["barcodes", "tails"]), dereplicate=rng.choice(["best", "none"]),
                 disable_preorient=False, disable_prefilter=rng.random() < 0.5)
    return primers, specimens, reads, rng.choice([9, 10, 11, 12]), flags


@pytest.mark.parametrize("seed", range(8))
def test_barcode_threshold_beyond_eight(seed):
    run_case(*make_large_k_case(seed), tag="large k seed %d" % seed)


def make_many_primers_case(seed, n_fwd=100, n_rev=100, n_pairs=300):
    """200 canonical primers in 300 primer pairs (the reference is unbounded, databases.py:247-264): multi-word
    primer-identity masks, the 254-primer selection instantiation, one sliced launch per primer; one barcode holds N."""
    rng = random.Random(11000 + seed)
    pools = ["PA", "PB", "PC"]
    seqs = set()
    while len(seqs) < n_fwd + n_rev:
        seqs.add(rand_seq(rng, rng.randint(18, 24)))
    seqs = sorted(seqs, key=lambda _: rng.random())
    primers = []
    for i, sq in enumerate(seqs):
        pool = pools[i % 3]
        primers.append(("F%03d" % i if i < n_fwd else "R%03d" % i, sq, "forward" if i < n_fwd else "reverse", [pool]))
    b1s = [rand_seq(rng, 13) for _ in range(12)]
    b2s = [rand_seq(rng, 13) for _ in range(24)]
    b1s[3] = b1s[3][:5] + "N" + b1s[3][6:]
    specimens, seen = [], set()
    while len(seen) < n_pairs:
        pool = rng.choice(pools)
        f = rng.choice([p for p in primers if p[2] == "forward" and p[3] == [pool]])
        r = rng.choice([p for p in primers if p[2] == "reverse" and p[3] == [pool]])
        if (f[0], r[0]) in seen:
            continue
        seen.add((f[0], r[0]))
        specimens.append(("S%04d" % len(specimens), pool, rng.choice(b1s), f[0], rng.choice(b2s), r[0]))
    by_name = {p[0]: p for p in primers}
    reads = []
    for r in range(60):
        sp = rng.choice(specimens)
        b1 = sp[2].replace("N", rng.choice("ACGTN"))
        s = (rand_seq(rng, rng.randint(0, 20)) + b1 + by_name[sp[3]][1] + rand_seq(rng, rng.randint(30, 150)) +
             reverse_complement(by_name[sp[5]][1]) + reverse_complement(sp[4]) + rand_seq(rng, rng.randint(0, 20)))
        s = mutate(rng, s, rng.choice([0.0, 0.04, 0.08]))
        if rng.random() < 0.5:
            s = reverse_complement(s)
        reads.append(("m%04d" % r, s, "I" * len(s)))
    flags = dict(search_len=80, trim="barcodes", dereplicate=rng.choice(["best", "none"]), disable_preorient=rng.random() < 0.5,
                 disable_prefilter=False)
    return primers, specimens, reads, 3, flags


@pytest.mark.parametrize("seed", range(3))
def test_two_hundred_primers_three_hundred_pairs(seed):
    run_case(*make_many_primers_case(seed), tag="many primers seed %d" % seed)


def test_long_primer_carry_lookahead():
    """The kernel's one-addition carry-lookahead over the words of every segment of a warp
    (long_carry_in<SW>) against a rippled carry chain, 2M random generate/propagate masks per width."""
    lib = H.hostsim_binding()
    assert lib.hostsim_check_long_carry(2_000_000) == 0


@pytest.mark.parametrize("seed", range(1000, 1016))
def test_random_case_long_primers(seed):
    run_case(*make_case(seed, long_primers=True), tag="long seed %d" % seed)


@pytest.mark.parametrize("seed", range(60))
def test_random_case(seed):
    run_case(*make_case(seed), tag="seed %d" % seed)


def run_case(primers, specimens, reads, k_idx, flags, tag, binding=None):
    """binding="cuda": through the CUDA library (GPU tests); default: the CPU kernel simulator."""
    try:
        tables = orc.Tables(primers, specimens)
    except ValueError:
        pytest.skip("generator produced an invalid table")
    oparams = orc.setup_params(tables, index_edit_distance=k_idx, search_len=flags["search_len"],
                               preorient=not flags["disable_preorient"], prefilter=not flags["disable_prefilter"],
                               trim=flags["trim"], dereplicate=flags["dereplicate"])
    expected, total, matched = orc.process_reads(tables, oparams, reads)
    sp = H.build_specimens(primers, specimens)
    params = MatchParameters(dict(oparams.max_dist_primers), k_idx, flags["search_len"], not flags["disable_preorient"])
    args = H.make_args(flags)
    ops, n, m = process_sequences(H.records(reads), params, sp, args, H.prefilter_for(args), None, 0,
                                  _binding=None if binding == "cuda" else H.hostsim_binding())
    assert (n, m) == (total, matched)
    H.assert_ops_equal([H.op_to_dict(o) for o in ops], [H.op_to_dict(o) for o in expected], tag)
