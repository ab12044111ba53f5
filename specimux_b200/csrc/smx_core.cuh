// smx_core.cuh -- host/device core of the B200 read-matching path.
//
// Everything here is pure integer logic written as __host__ __device__ inline functions so the CUDA
// kernels (smx_kernels.cuh) and the CPU-side kernel simulator used by the unit tests
// (tests/hostsim) execute the very same code.  Reference semantics are cited per function
// (paths relative to /root/reference/src/specimux/).
#pragma once
#include <stdint.h>

#include "../../include/specimux_b200.h"

#if defined(__CUDACC__)
#define SMX_HD __host__ __device__ __forceinline__
#else
#define SMX_HD inline
#endif

namespace smx {

static_assert(sizeof(smx_record) == 64, "smx_record layout");
static_assert(sizeof(smx_record32) == 32, "smx_record32 layout");
static_assert(sizeof(smx_record16) == 16, "smx_record16 layout");
static_assert(sizeof(smx_primer_hit) == 12, "smx_primer_hit layout");
static_assert(sizeof(smx_barcode_hit) == 16, "smx_barcode_hit layout");

typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

constexpr int kSymOther = 15;       // read symbol that matches nothing (lower case, unknown bytes)
constexpr int kNone = INT32_MIN;    // Python None in coordinates
constexpr int kMaxPairs = SMX_MAX_PAIRS;   // candidate slots per read = 2 * pairs
constexpr int kSmallGroups = 16;    // dereplication groups tracked per read in the first pass
constexpr int kBigGroups = 4096;
constexpr int kMaxTaskWords = 4;    // bwords one stage-2 thread evaluates side by side
constexpr int kMaxBarcodeK = 12;     // largest barcode threshold k_idx (the reference derives ceil(min distance / 2))
constexpr int kQuadSyms = 5, kQuadRow = 625;   // four-entry narrow-word tables: symbols per entry, entries per row (5^4)
constexpr int kMaxTaskK = 4;        // largest k_idx with multi-word stage-2 tasks (beyond: one bword per task)
// control block of a batch: kCtrWords 64-bit counters followed by the 2 * SMX_MAX_PRIMERS u32 per-slot entry counts
constexpr int kCtrDeferred = 8;     // (u32) reads left to the general selection kernel
constexpr int kCtrBig = 9;          // (u32) reads left to the second selection pass (k_select_big)
constexpr int kCtrUseful1 = 10;     // stage 1: bit-sliced DP cells (pattern rows x columns per 32 reads) actually evaluated
constexpr int kCtrUseful2 = 11;     // stage 2: bit-sliced DP cells (band cells x bwords per work entry) actually evaluated
constexpr int kCtrWords = 12;

// ---------------------------------------------------------------------------------------------
// Symbols.  4-bit read codes: A0 C1 G2 T3 R4 Y5 S6 W7 K8 M9 B10 D11 H12 V13 N14 other15.

SMX_HD int sym_complement(int c) {
    // Biopython ambiguous-DNA complement restricted to the 16 codes (Bio.Seq).
    // A<->T C<->G R<->Y S W K<->M B<->V D<->H N other
    const u64 table = 0xFE'A'B'C'D'8'9'7'6'4'5'0'1'2'3ull;   // nibble i = complement of code i
    return (int)((table >> (4 * c)) & 15);
}

// ---------------------------------------------------------------------------------------------
// smx_record -> the compact wire forms.

SMX_HD void pack_record32(const smx_record &r, smx_record32 &o) {
    o.read = r.read; o.sample = r.sample; o.trim_start = r.trim_start; o.trim_end = r.trim_end;
    o.pool = r.pool; o.p1 = r.p1; o.p2 = r.p2;
    o.dist[0] = r.dist[0]; o.dist[1] = r.dist[1]; o.dist[2] = r.dist[2]; o.dist[3] = r.dist[3];
    o.resolution = r.resolution;
    o.flags = (uint8_t)((r.reverse ? 1 : 0) | (r.trim_empty ? 2 : 0));
    o.candidate = r.candidate; o.pad[0] = o.pad[1] = o.pad[2] = 0;
}

// n: true length of the record's read; last: last record of its read.  Returns false when an extent does not
// fit 16 bits (see smx_record16).
SMX_HD bool pack_record16(const smx_record &r, int n, bool last, smx_record16 &o) {
    const int tail = n - r.trim_end;
    o.sample = r.sample;
    o.trim_start = (uint16_t)r.trim_start; o.trim_tail = (uint16_t)tail;
    o.pool = r.pool;
    o.p1 = (uint8_t)(r.p1 < 0 ? 0xFF : r.p1); o.p2 = (uint8_t)(r.p2 < 0 ? 0xFF : r.p2);
    o.dist_p1 = (uint8_t)(r.dist[0] < 0 ? 0xFF : r.dist[0]); o.dist_p2 = (uint8_t)(r.dist[3] < 0 ? 0xFF : r.dist[3]);
    o.dist_b = (uint8_t)((r.dist[1] < 0 ? 0xF : r.dist[1]) | ((r.dist[2] < 0 ? 0xF : r.dist[2]) << 4));
    o.flags = (uint8_t)(r.resolution | (r.reverse ? 8 : 0) | (r.trim_empty ? 16 : 0) | (last ? 32 : 0));
    return r.trim_start >= 0 && r.trim_start <= 0xFFFF && tail >= 0 && tail <= 0xFFFF && r.p1 < 0xFF && r.p2 < 0xFF &&
           r.dist[1] < 0xF && r.dist[2] < 0xF;
}

// ---------------------------------------------------------------------------------------------
// Window geometry.  Exact Python-slice semantics of align_seq (alignment.py:37-49) for the primer
// search window of match_one_end (demultiplex.py:757-758), including reads shorter than the
// search length.

struct Geo {
    int n;       // read length
    int woff;    // X index of staged symbol 0; the staged buffer is X[woff : n]
    int wl;      // staged symbols = min(n, L)
    int start;   // staged index where the primer search window begins
    int wlen;    // primer search window length
    int delta;   // reported coordinate = X index + delta
    bool regular;
};

SMX_HD Geo make_geo(int n, int L) {
    Geo g;
    g.n = n;
    int raw = n - L;                       // search_start (demultiplex.py:757)
    g.woff = raw > 0 ? raw : 0;
    g.wl = n - g.woff;
    int off, shift;
    if (raw >= 0) { off = raw; shift = raw; }
    else if (raw == -1) { off = 0; shift = 0; }                    // alignment.py:37
    else { off = n + raw; if (off < 0) off = 0; shift = raw; }      // negative Python slice start
    g.start = off - g.woff;
    g.wlen = n - off;
    g.delta = shift - off;
    g.regular = raw >= -1;
    return g;
}

// Barcode flank of match_one_end (demultiplex.py:787-800) for a primer end reported at e_rep.
struct Flank {
    int bs;        // barcode_search_start as passed to align_seq
    int a_align;   // X index where the aligned flank starts (alignment.py:37-40)
    int a_pref;    // X index where the prefilter's flank starts (demultiplex.py:789)
    int bshift;    // value added to SHW locations (alignment.py:49)
};

SMX_HD Flank make_flank(int e_rep, int n) {
    Flank f;
    f.bs = e_rep + 1;
    if (f.bs == -1) {
        f.a_align = 0; f.bshift = 0;
        f.a_pref = n > 0 ? n - 1 : 0;
    } else if (f.bs < 0) {
        int a = n + f.bs; if (a < 0) a = 0;
        f.a_align = f.a_pref = a; f.bshift = f.bs;
    } else {
        int a = f.bs < n ? f.bs : n;
        f.a_align = f.a_pref = a; f.bshift = f.bs;
    }
    return f;
}

// ---------------------------------------------------------------------------------------------
// Myers / Hyyro bit-vector column step.  The pattern is TOP-aligned in the word (row m is the
// sign bit) so the score delta of the last row is a plain shift.  Bits below the pattern are
// neutral: HW initialises them Pv=1 (they keep emitting Ph=Mh=0, i.e. D[0][j]=0), SHW/NW shift a 1
// in at bit 0 every column (D[0][j]=j).  Restates edlib's calculateBlock (the arithmetic behind
// alignment.py:42) for a single block.

template <typename W> struct WordBits;
template <> struct WordBits<u32> { static constexpr int bits = 32; };
template <> struct WordBits<u64> { static constexpr int bits = 64; };

template <typename W, bool kShiftInOne>
SMX_HD int myers_step(W Eq, W &Pv, W &Mv) {
    constexpr int B = WordBits<W>::bits;
    W Xv = Eq | Mv;
    W Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    W Ph = Mv | ~(Xh | Pv);
    W Mh = Pv & Xh;
    int d = (int)(Ph >> (B - 1)) - (int)(Mh >> (B - 1));
    Ph = (Ph << 1) | (W)(kShiftInOne ? 1 : 0);
    Mh = Mh << 1;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return d;
}

template <typename W> SMX_HD W pattern_mask(int m) {
    constexpr int B = WordBits<W>::bits;
    return m >= B ? ~(W)0 : (~(W)0) << (B - m);
}

// Peq tables are stored as u64 with the pattern top-aligned at bit 63; a 32-bit kernel uses the
// high half.
template <typename W> SMX_HD W peq_word(u64 v);
template <> SMX_HD u32 peq_word<u32>(u64 v) { return (u32)(v >> 32); }
template <> SMX_HD u64 peq_word<u64>(u64 v) { return v; }

// ---------------------------------------------------------------------------------------------
// Device-resident tables (plain pointers; filled by smx_api.cu).

struct Tables {
    int n_primers, n_pairs, n_specimens, n_keys;
    int L, wpw, mw;                 // search_len, staged words per window (8 syms/word), mask words
    int nw2;                        // 2-bit window words per strand (16 symbols each) = blocks of the sliced primer search
    int sliced;                     // stage 1 runs bit-sliced across reads (all primers <= 32 nt)
    int k_idx, blen_max;
    int preorient, prefilter, trim, derep_best, min_length, max_length;
    int total_bslots;               // sum over (strand, primer) of barcode-list lengths
    int use64;                      // any primer longer than 32
    int hit_cap;                    // entries per (read, strand, bword) hit sub-list

    unsigned short p_len[SMX_MAX_PRIMERS];
    short p_k[SMX_MAX_PRIMERS];
    unsigned char p_sw[SMX_MAX_PRIMERS];   // long primers (> 64 nt): lanes per problem of the warp-cooperative search, else 0
    int p_long[SMX_MAX_PRIMERS];           // long primers: word offset of their tables in peq_long
    unsigned char p_dir[SMX_MAX_PRIMERS];
    int p_fidx[SMX_MAX_PRIMERS];
    u32 pb_off[SMX_MAX_PRIMERS + 1];
    u32 bslot_base[2 * SMX_MAX_PRIMERS];

    const u64 *peq_rc;       // [primer][16]   primer_rc, top-aligned
    const u64 *peq_rcrev;    // [primer][16]   reversed primer_rc (start-recovery pass)
    const u64 *peq_fw;       // [primer][16]   forward-sense primer (explicit orientation test)
    const u32 *peq_long;     // long primers: [rc | rcrev | fw][16 symbols][p_sw words], pattern top-aligned in 32*p_sw bits
    const unsigned char *b_len;    // [list entry]
    const unsigned char *b_codes;  // barcode_rc symbols (4-bit codes, one per byte) of every list entry, concatenated
    const u32 *b_code_off;         // [list entry] offset into b_codes
    const u32 *bw_iupac;           // [bword] bit lanes whose barcode holds an IUPAC code (exact Bloom emulation, see
                                   // bloom_literal_yes)
    const u32 *pb_barcode;   // [list entry] global barcode id

    // Bit-sliced barcode match table: barcodes of one primer and one length are grouped 32 to a
    // "bword"; beq[(bw_row[g] + i) * 16 + sym] has bit q set iff row i (0-based) of the q-th
    // barcode_rc of bword g equals read symbol `sym` (edlib additionalEqualities semantics).
    int n_bwords;
    u32 bw_off[SMX_MAX_PRIMERS + 1];   // primer p owns bwords [bw_off[p], bw_off[p+1])
    const unsigned char *bw_len;       // [bword] barcode length m
    const unsigned char *bw_primer;    // [bword] owning primer
    const u32 *bw_row;                 // [bword] first row in beq
    const u32 *bw_valid;               // [bword] mask of populated bit lanes
    const unsigned short *bw_list;     // [bword][32] position j in the primer's barcode list
    const u32 *beq;
    // Stage-2 tasks: up to four bwords of one primer and one barcode length.  One thread evaluates a work
    // entry against all words of its task (flank set-up shared, one vector table load per DP cell).
    int n_btasks;
    u32 bt_off[SMX_MAX_PRIMERS + 1];   // primer p owns tasks [bt_off[p], bt_off[p+1])
    const unsigned short *bt_g0;       // [task] first bword
    const unsigned char *bt_nw;        // [task] number of bwords (1..4)
    const u32 *bt_row;                 // [task] word offset of the task's table in bt_eq
    const u32 *bt_eq;                  // per task [row][16 symbols][S], S = 1, 2 or 4 (bt_nw rounded up to a power of two)
    const i32 *bt_quad_row;            // [task] word offset of the task's four-entry table in bt_quad ([row][625]), -1 = none
    const u32 *bt_quad;

    const u32 *pair_fwd, *pair_rev;
    const i32 *pair_pool;

    const i32 *spec_dense;   // optional [b1 * n_b2 + b2] -> key index or -1 (nullptr when too large)
    int n_b2;
    const u64 *spec_key;     // sorted unique (b1 << 32 | b2)
    const u32 *spec_key_off; // n_keys+1 -> rows of that key in file order
    const u32 *spec_row;     // row indices
    const u64 *spec_p1_mask, *spec_p2_mask;   // by row, pmask_words words each (bit c of word c / 64 = canonical primer c)
    int pmask_words;
    const i32 *spec_pool;    // by row
};

// Per-(slot, read) digest of the barcode hit lists, produced by the summary kernel so that the
// selection stage does not walk the lists in the common case.
struct SlotSum {
    u64 first_mask;            // lowest equal-best end column of the first equal-best barcode, as a one-bit mask
    i32 first_ss;              // its barcode_search_start
    unsigned short first_best; // its list position
    unsigned short nhits;      // barcodes within k_idx (saturating)
    unsigned short nbest;      // barcodes at the best distance (saturating)
    signed char bd;            // best distance, -1 = none
    unsigned char pad;
};
static_assert(sizeof(SlotSum) == 24, "SlotSum layout");

// Stage-2 tasks of one (words per task, barcode length): bt_class_tasks[off .. off + count), one kernel launch.
// quad: narrow single-word tasks (<= 8 barcodes) evaluated four work entries to a word (barcode_quad_thread).
struct BtClass { int nw, m, quad; u32 off, count; };

// Digest of the barcode hits of one (work entry, stage-2 task), written by the barcode kernel so that the
// selection stage reads 16 bytes per entry instead of walking the hit sub-lists in the common case.
struct BarcodeDigest {
    i32 search_start;          // barcode_search_start of this primer end location
    unsigned short jmin, jmax; // smallest / largest list position among the hits at distance `bd`
    unsigned short count;      // hits at distance `bd` (saturating)
    signed char bd;            // smallest distance, -1 = no hit
    unsigned char first_col;   // lowest equal-best SHW end column of the hit at jmin
    u32 nhits;                 // hits within k_idx
};
static_assert(sizeof(BarcodeDigest) == 16, "BarcodeDigest layout");

// Per-batch device buffers.
struct Batch {
    u32 n_reads, n_pad;
    const u32 *packed2;
    const u64 *word_off;
    const u32 *lengths;
    const u32 *packed4;
    const u64 *off4;         // nullptr when no read is flagged
    u32 clip;                // 0, or reads longer than 2*clip are stored as head clip + tail clip bases
    u32 stride;              // 0, or every read owns `stride` words of packed2 (word_off unused)
    u64 word_base;           // word_off values are absolute stream offsets; packed2[0] is stream word word_base
    u32 read_base;           // index of this (sub-)batch's read 0 in the caller's batch (smx_record.read)
    u32 *win;                // staged 4-bit windows [(strand*wpw + w) * n_pad + read]
    u32 *win2;               // staged 2-bit windows [(strand*nw2 + w2) * n_pad + read], 16 symbols/word (unflagged reads)
    u32 *tmix;               // sliced primer search output [(slot*nw2 + blk) * n_pad + read]: bit c = equal-best so far,
                             // bit 16+c = improvement, for column 16*blk + c
    // level-1 results
    smx_primer_hit *phit;    // [slot * n_pad + read], slot = strand*n_primers + primer
    u32 *endmask;            // [(slot*mw + w) * n_pad + read]
    u32 *impmask;            // same shape: columns where the running best improved (stage-1 scratch)
    unsigned char *orient_hit;   // [slot * n_pad + read]  explicit orientation test (irregular reads)
    // One work entry per (matched slot, equal-best primer end location); a read's entries are
    // consecutive and in ascending end order.  e_cap entries are allocated per slot.
    u32 e_cap;
    u32 *slot_count;         // [slot] entries appended (may exceed e_cap -> the library re-runs larger)
    u32 *ent_base;           // [slot * n_pad + read] first entry of the read in this slot
    u32 *ent_read;           // [slot * e_cap + e]
    unsigned short *ent_pos; // [slot * e_cap + e] staged position of the primer end
    // Barcode hits of entry e for bword g, gslot = strand * n_bwords + g:
    unsigned char *bh_count; // [gslot * e_cap + e]; may exceed hit_cap (-> re-run with a larger cap)
    smx_barcode_hit *bh_list;    // [(gslot * hit_cap + h) * e_cap + e], ascending barcode position
    BarcodeDigest *bdig;         // [(strand * n_btasks + task) * e_cap + e]
    u32 *defer_list;         // reads the fast selection kernel left to the general one (count: counters[kCtrDeferred])
    u32 *big_list;           // reads the general selection left to k_select_big (count: counters[kCtrBig])
    // single-pass selection: first record of every read + pool for the (rare) further records
    smx_record *rec_stage;   // [read]
    smx_record *rec_pool;    // extra records, contiguous per read
    u32 *rec_extra;          // [read] index of the read's 2nd record in rec_pool
    u32 pool_cap;
    // level-2 results
    u32 *rec_count;          // per read
    u32 *rec_offset;         // exclusive scan (n_reads + 1)
    smx_record *records;
    unsigned char *read_flags;   // bit0: read had a full match; bit1: internal cap overflow
    unsigned long long *counters;    // [0] HW cells [1] SHW cells [2] HW wordcols [3] SHW wordcols
                                     // [4] matched reads [5] lo: select overflow, hi: hit-list overflow [6] records
};

SMX_HD void counter_add(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

// First word of read r inside this (sub-)batch's packed2 buffer.
SMX_HD u64 read_word0(const Batch &b, u32 r) {
    return b.stride ? (u64)r * b.stride : b.word_off[r] - b.word_base;
}

SMX_HD bool read_is_flagged(const Batch &b, u32 r) {
    return b.off4 != nullptr && b.off4[r] != ~0ull;
}

// Stored position of base x of an n-long strand under clipping (only x < clip or x >= n - clip
// are ever requested when clip >= search_len).
SMX_HD int stored_pos(const Batch &b, int x, int n) {
    if (b.clip == 0 || n <= 2 * (int)b.clip) return x;
    return x < (int)b.clip ? x : x - (n - 2 * (int)b.clip);
}
SMX_HD int stored_len(const Batch &b, int n) {
    return (b.clip == 0 || n <= 2 * (int)b.clip) ? n : 2 * (int)b.clip;
}

// Symbol x of strand `strand` of read r, straight from the packed streams.
SMX_HD int sym_at(const Batch &b, u32 r, int strand, int x, int n) {
    if (read_is_flagged(b, r)) {
        u64 base = b.off4[r] + (strand ? (u64)((stored_len(b, n) + 7) >> 3) : 0);
        int xs = stored_pos(b, x, n);
        u32 w = b.packed4[base + (u64)(xs >> 3)];
        return (int)((w >> (4 * (xs & 7))) & 15);
    }
    int i = stored_pos(b, strand ? n - 1 - x : x, n);
    u32 w = b.packed2[read_word0(b, r) + (u64)(i >> 4)];
    int c = (int)((w >> (2 * (i & 15))) & 3);
    return strand ? 3 - c : c;
}

// ---------------------------------------------------------------------------------------------
// Primer HW search over the staged window (demultiplex.py:765 -> alignment.py:21-50 -> edlib HW).
// `load(p)` returns the 4-bit symbol at staged position p.

template <typename W> struct HwResult {
    int best;       // min over columns of D[m][j]
    int first;      // staged position of the first column attaining `best`
};

// Start recovery (edlib: reverse SHW pass with k = best, LAST equal-best end wins).
// Returns the number of columns walked back from the end to the start: start = end - ret.
template <typename W, typename Load>
SMX_HD int hw_start_back(const u64 *peq_rev, int m, int best, int e_pos, int start_pos, Load load) {
    W Pv = pattern_mask<W>(m), Mv = 0;
    int score = m;
    int cols = e_pos - start_pos + 1;
    int lim = m + best;
    if (cols > lim) cols = lim;
    int last = m - 1;           // falls back to an m-long hit (never used: a best column exists)
    for (int j = 0; j < cols; ++j) {
        int c = load(e_pos - j);
        W Eq = peq_word<W>(peq_rev[c]);
        score += myers_step<W, true>(Eq, Pv, Mv);
        if (score == best) last = j;
    }
    return last;
}

// ---------------------------------------------------------------------------------------------
// Specimen lookup (databases.py:219-245).

SMX_HD int spec_find_key(const Tables &t, u32 b1, u32 b2) {
    if (t.spec_dense) return t.spec_dense[(u64)b1 * t.n_b2 + b2];
    u64 key = ((u64)b1 << 32) | b2;
    int lo = 0, hi = t.n_keys - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        u64 v = t.spec_key[mid];
        if (v == key) return mid;
        if (v < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

// Row `row` lists canonical primers p1 / p2 among its resolved primers (identity test, databases.py:224-225,241).
SMX_HD bool spec_row_has(const Tables &t, u32 row, int p1, int p2) {
    const u64 w1 = t.spec_p1_mask[(u64)row * t.pmask_words + (p1 >> 6)], w2 = t.spec_p2_mask[(u64)row * t.pmask_words + (p2 >> 6)];
    return ((w1 >> (p1 & 63)) & 1) && ((w2 >> (p2 & 63)) & 1);
}

// First specimen row (file order) with exactly (b1, b2, p1, p2); -1 if none.  specimen_for_exact_match.
SMX_HD int spec_exact(const Tables &t, u32 b1, u32 b2, int p1, int p2) {
    int k = spec_find_key(t, b1, b2);
    if (k < 0) return -1;
    for (u32 i = t.spec_key_off[k]; i < t.spec_key_off[k + 1]; ++i) {
        u32 row = t.spec_row[i];
        if (spec_row_has(t, row, p1, p2)) return (int)row;
    }
    return -1;
}

// Count of rows matching (b1, b2, p1, p2) and the smallest such row.  specimens_for_barcodes_and_primers.
SMX_HD void spec_all(const Tables &t, u32 b1, u32 b2, int p1, int p2, int &count, int &min_row) {
    int k = spec_find_key(t, b1, b2);
    if (k < 0) return;
    for (u32 i = t.spec_key_off[k]; i < t.spec_key_off[k + 1]; ++i) {
        u32 row = t.spec_row[i];
        if (spec_row_has(t, row, p1, p2)) {
            ++count;
            if (min_row < 0 || (int)row < min_row) min_row = (int)row;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Selection / dereplication / specimen resolution / trimming for one read
// (demultiplex.py:126-210, 216-598; models.py:72-328).  One thread runs this per read.

struct EndInfo {            // one (strand, primer) slot as seen by a candidate (kept small: it lives in local memory)
    u64 first_mask;         // lowest equal-best SHW end column of the first equal-best barcode (one-bit mask)
    int ps, pe;             // first primer location, reported X coordinates
    int first_ss;           // barcode_search_start of the first equal-best barcode
    int first_best;         // its list position (pinned order), -1 = none
    short pd;               // primer distance
    unsigned short nhits;   // barcodes within k_idx
    unsigned short nbest;   // number of barcodes at bd
    signed char bd;         // best barcode distance
    unsigned char matched;
    unsigned char strand, primer;
};
static_assert(sizeof(EndInfo) == 40, "EndInfo layout");

// Gathered equal-best barcode list of one end (see gather_best).
constexpr int kBestCache = 12;
struct BestList {
    short n;                               // -1: not gathered yet, -2: more than kBestCache (walk the lists instead)
    unsigned short j[kBestCache];
};

struct SelectCtx {
    const Tables *t;
    const Batch *b;
    u32 read;
    int n;
    BestList *best = nullptr;              // optional [2 * n_primers] cache, entries start at n = -1 (general selection)
};

SMX_HD u32 slot_index(const Tables &t, int strand, int primer) { return (u32)(strand * t.n_primers + primer); }

// Barcode hits of one (strand, primer) slot live in per-(end location, bword) sub-lists written by
// stage 2, each ascending in barcode position.  next_hit() walks their union in ascending
// position; for a barcode found at several primer end locations the strictly smallest distance
// wins and ties keep the earliest location (demultiplex.py:786-812).
SMX_HD bool next_hit(const SelectCtx &c, int strand, int primer, int after_j, smx_barcode_hit &out) {
    const Tables &t = *c.t;
    const Batch &b = *c.b;
    const u32 slot = (u32)(strand * t.n_primers + primer);
    const u64 hidx = (u64)slot * b.n_pad + c.read;
    const u32 e0 = b.ent_base[hidx];
    u32 nloc = b.phit[hidx].n_locations;
    if (e0 >= b.e_cap) return false;
    if (e0 + nloc > b.e_cap) nloc = b.e_cap - e0;
    bool found = false;
    for (u32 l = 0; l < nloc; ++l) {
        const u64 e = (u64)e0 + l;
        for (u32 g = t.bw_off[primer]; g < t.bw_off[primer + 1]; ++g) {
            const u64 gslot = (u64)strand * t.n_bwords + g;
            int cnt = b.bh_count[gslot * b.e_cap + e];
            if (cnt > t.hit_cap) cnt = t.hit_cap;
            for (int x = 0; x < cnt; ++x) {
                const smx_barcode_hit &h = b.bh_list[(gslot * t.hit_cap + x) * b.e_cap + e];
                int j = (int)h.barcode;
                if (j <= after_j) continue;
                if (!found || j < (int)out.barcode || (j == (int)out.barcode && h.distance < out.distance)) {
                    out = h; found = true;
                }
                break;      // sub-list is ascending: the first j > after_j is this list's candidate
            }
        }
    }
    return found;
}

// Digest of the barcode hits of one matched slot, merged from the per-(location, task) digests of stage 2.
// A barcode found at several primer end locations keeps its strictly smallest distance, ties the
// earliest location (demultiplex.py:786-812); the best distance over the merged barcodes is simply
// the minimum over all hits, the equal-best barcodes are those with some hit at that distance, the
// first of them (pinned order) is the smallest list position, and its location is the earliest one
// carrying that hit.  nhits counts hits (only zero / non-zero is ever used); nbest is exact for a
// single location and saturates at 2 otherwise (only 0 / 1 / more is ever used).
SMX_HD void summarize_slot(const SelectCtx &c, int strand, int primer, SlotSum &o, u32 e0, u32 nloc,
                           const BarcodeDigest *d0 = nullptr) {
    // e0 / nloc: the slot's first work entry and its number of equal-best primer ends (loaded by the caller, so that a
    // caller with several slots can have all of those loads in flight at once); d0: the digest of (entry e0, the
    // primer's first task) if the caller fetched it ahead as well.
    o.first_mask = 0; o.first_ss = 0; o.first_best = 0; o.nhits = 0; o.nbest = 0; o.bd = -1; o.pad = 0;
    const Tables &t = *c.t;
    const Batch &b = *c.b;
    if (e0 >= b.e_cap) return;
    if (e0 + nloc > b.e_cap) nloc = b.e_cap - e0;
    int bd = 1 << 20, count = 0, jmin = 1 << 20, jmax = -1;
    u32 nhits = 0;
    for (u32 l = 0; l < nloc; ++l) {
        const u64 e = (u64)e0 + l;
        for (u32 tk = t.bt_off[primer]; tk < t.bt_off[primer + 1]; ++tk) {
            const BarcodeDigest d = (d0 && l == 0 && tk == t.bt_off[primer])
                                        ? *d0 : b.bdig[((u64)strand * t.n_btasks + tk) * b.e_cap + e];
            if (!d.nhits) continue;
            nhits += d.nhits;
            if (d.bd < bd) { bd = d.bd; count = 0; jmin = 1 << 20; jmax = -1; }
            if (d.bd == bd) {
                count += d.count;
                if ((int)d.jmin < jmin) {
                    jmin = d.jmin; o.first_best = d.jmin; o.first_mask = 1ull << d.first_col; o.first_ss = d.search_start;
                }
                if ((int)d.jmax > jmax) jmax = d.jmax;
            }
        }
    }
    if (!nhits) return;
    const int nbest = nloc == 1 ? count : (jmin == jmax ? 1 : 2);
    o.bd = (signed char)bd;
    o.nhits = (unsigned short)(nhits > 65535 ? 65535 : nhits);
    o.nbest = (unsigned short)(nbest > 65535 ? 65535 : nbest);
}

// EndInfo of one slot from its (already loaded) primer hit, first work entry and, optionally, first digest.
SMX_HD void fill_end(const SelectCtx &c, int strand, int primer, const smx_primer_hit &ph, u32 e0, const BarcodeDigest *d0,
                     EndInfo &e) {
    e.strand = (unsigned char)strand; e.primer = (unsigned char)primer;
    e.matched = ph.distance >= 0;
    e.pd = ph.distance; e.ps = ph.first_start; e.pe = ph.first_end;
    e.nhits = 0; e.bd = -1; e.nbest = 0; e.first_best = -1; e.first_ss = 0; e.first_mask = 0;
    if (!e.matched) return;
    SlotSum ss;
    summarize_slot(c, strand, primer, ss, e0, ph.n_locations, d0);
    e.nhits = ss.nhits; e.bd = ss.bd; e.nbest = ss.nbest;
    e.first_best = ss.nhits ? (int)ss.first_best : -1; e.first_ss = ss.first_ss; e.first_mask = ss.first_mask;
}

SMX_HD void load_end(const SelectCtx &c, int strand, int primer, EndInfo &e) {
    const Tables &t = *c.t;
    const u64 idx = (u64)slot_index(t, strand, primer) * c.b->n_pad + c.read;
    const smx_primer_hit ph = c.b->phit[idx];
    fill_end(c, strand, primer, ph, ph.distance >= 0 ? c.b->ent_base[idx] : 0u, nullptr, e);
}

// All 2 * NP slots of a read when the number of primers is a compile-time constant: the primer hits and first work
// entries of every slot are loaded first (2 * NP independent pairs of loads in flight), then the first digest of every
// matched slot, and only then is anything consumed.  The one-slot-at-a-time form above is a chain of three dependent
// loads per slot, one slot after the other: 12 round trips to L2 / HBM per read for two primers, and k_select_fast
// spent most of its time waiting on them (profiles/r2_j: long-scoreboard 14 stalls per issue).
template <int NP>
SMX_HD void load_ends_ahead(const SelectCtx &c, EndInfo *ends) {
    constexpr int NS = 2 * NP;
    const Tables &t = *c.t;
    const Batch &b = *c.b;
    smx_primer_hit ph[NS];
    u32 e0[NS];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NS; ++i) {
        ph[i] = b.phit[(u64)i * b.n_pad + c.read];
        e0[i] = b.ent_base[(u64)i * b.n_pad + c.read];       // only meaningful where the slot matched
    }
    BarcodeDigest d0[NS];
    bool have[NS];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NS; ++i) {
        const int strand = i / NP, primer = i % NP;
        have[i] = ph[i].distance >= 0 && e0[i] < b.e_cap && t.bt_off[primer] < t.bt_off[primer + 1];
        if (have[i]) d0[i] = b.bdig[((u64)strand * t.n_btasks + t.bt_off[primer]) * b.e_cap + e0[i]];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NS; ++i) fill_end(c, i / NP, i % NP, ph[i], e0[i], have[i] ? &d0[i] : nullptr, ends[i]);
}

// Equal-best barcodes of an end in pinned order (models.py:116-126 best_b1 / best_b2): the list
// position after `after_j`, or -1.
SMX_HD int next_best(const SelectCtx &c, const EndInfo &e, int after_j) {
    smx_barcode_hit h;
    while (next_hit(c, e.strand, e.primer, after_j, h)) {
        after_j = (int)h.barcode;
        if (h.distance == e.bd) return after_j;
    }
    return -1;
}

// The equal-best barcodes of an end, gathered ONCE: every hit at the end's best distance e.bd, over all of its
// sub-lists, sorted by list position without duplicates -- the sequence next_best() walks (a barcode's merged distance
// is the minimum over its hits and e.bd the minimum over all of them, so the barcode is equal-best exactly when one of
// its hits is at e.bd).  next_best() re-walks every sub-list per call and the general selection calls it in nested
// loops (dereplication keys, in_none, resolve): for a deferred read that was a chain of hundreds of dependent loads.
SMX_HD void gather_best(const SelectCtx &c, const EndInfo &e, BestList &bl) {
    const Tables &t = *c.t;
    const Batch &b = *c.b;
    bl.n = 0;
    if (!e.matched || e.nhits == 0) return;
    const u32 slot = (u32)(e.strand * t.n_primers + e.primer);
    const u64 hidx = (u64)slot * b.n_pad + c.read;
    const u32 e0 = b.ent_base[hidx];
    u32 nloc = b.phit[hidx].n_locations;
    if (e0 >= b.e_cap) return;
    if (e0 + nloc > b.e_cap) nloc = b.e_cap - e0;
    for (u32 l = 0; l < nloc; ++l) {
        const u64 en = (u64)e0 + l;
        for (u32 g = t.bw_off[e.primer]; g < t.bw_off[e.primer + 1]; ++g) {
            const u64 gslot = (u64)e.strand * t.n_bwords + g;
            int cnt = b.bh_count[gslot * b.e_cap + en];
            if (cnt > t.hit_cap) cnt = t.hit_cap;
            for (int x = 0; x < cnt; ++x) {
                const smx_barcode_hit &h = b.bh_list[(gslot * t.hit_cap + x) * b.e_cap + en];
                if (h.distance != e.bd) continue;
                const unsigned short j = h.barcode;
                int pos = 0;
                while (pos < bl.n && bl.j[pos] < j) ++pos;
                if (pos < bl.n && bl.j[pos] == j) continue;
                if (bl.n == kBestCache) { bl.n = -2; return; }
                for (int q = bl.n; q > pos; --q) bl.j[q] = bl.j[q - 1];
                bl.j[pos] = j;
                ++bl.n;
            }
        }
    }
}

// k-th equal-best barcode of end e (list position), -1 past the last: from the gathered list when the context carries
// a cache (indexed like the EndInfo cache), else by walking on from prev_j.
SMX_HD int best_at(const SelectCtx &c, const EndInfo &e, int k, int prev_j) {
    if (c.best) {
        BestList &bl = c.best[(int)e.strand * c.t->n_primers + (int)e.primer];
        if (bl.n == -1) gather_best(c, e, bl);
        if (bl.n >= 0) return k < bl.n ? (int)bl.j[k] : -1;
    }
    return next_best(c, e, prev_j);
}

struct Cand {               // CandidateMatch (models.py:72-95) by reference to its two ends
    int pair, rc;           // rc = 1: candidate built on the reverse-complemented read
    int e1, e2;             // indices into the per-read EndInfo cache (always valid)
};

SMX_HD int cand_score(const EndInfo &a, const EndInfo &b) {   // demultiplex.py:226-236
    bool p1 = a.matched, p2 = b.matched, b1 = a.nhits > 0, b2 = b.nhits > 0;
    if (p1 && p2 && b1 && b2) return 5;
    if (p1 && p2 && (b1 || b2)) return 4;
    if ((p1 || p2) && (b1 || b2)) return 3;
    if (p1 && p2) return 2;
    if (p1 || p2) return 1;
    return 0;
}

// Band cells the stage-2 automaton evaluates for an m-row pattern (columns j = i-K .. i+K inside 1 .. 16, or the
// whole band for the general form).
SMX_HD int band_cells(int m, int K, bool small) {
    int n = 0;
    for (int i = 1; i <= m; ++i) {
        int lo = i - K < 1 ? 1 : i - K, hi = i + K;
        if (small && hi > 16) hi = 16;
        n += hi - lo + 1;
    }
    return n;
}

// Emits records for one read.  `emit` = nullptr counts only.  Returns the number of records.
struct Emitter {
    smx_record *out;        // may be nullptr (count pass)
    u32 cap;                // records `out` can hold; further records are only counted
    u32 count;
    bool full;
};

struct TrimState {          // cumulative trim_locations() shift per candidate
    int *cand;
    int *shift;
    int n, cap;
};

SMX_HD int trim_shift_get(const TrimState &s, int cand) {
    for (int i = 0; i < s.n; ++i) if (s.cand[i] == cand) return s.shift[i];
    return 0;
}
SMX_HD void trim_shift_add(TrimState &s, int cand, int v, bool &overflow) {
    for (int i = 0; i < s.n; ++i) if (s.cand[i] == cand) { s.shift[i] += v; return; }
    if (s.n < s.cap) { s.cand[s.n] = cand; s.shift[s.n] = v; ++s.n; } else overflow = true;
}

// intertail_extent's folds over barcode locations with the reference's -1 sentinel
// (models.py:300-319).  Hits are visited in stable distance order, locations ascending.
SMX_HD int lowest_bit64(u64 m) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}

SMX_HD void tails_fold(const SelectCtx &c, const EndInfo &e, bool is_b1, int shift, int &acc) {
    if (!e.matched || e.nhits == 0) return;
    const Tables &t = *c.t;
    for (int d = 0; d <= t.k_idx; ++d) {
        smx_barcode_hit h;
        int after = -1;
        while (next_hit(c, e.strand, e.primer, after, h)) {
            after = (int)h.barcode;
            if (h.distance != d) continue;
            int bshift = h.search_start == -1 ? 0 : h.search_start;
            u64 m = h.end_mask;
            while (m) {
                int col = lowest_bit64(m);
                m &= m - 1;
                int e_x = bshift + col;                        // SHW location (bshift, e_x) in X coordinates
                if (is_b1) {
                    int l0 = c.n - e_x - 1 - shift;             // reversed(): (len-e-1, len-s-1)
                    acc = (acc == -1) ? l0 : (l0 < acc ? l0 : acc);
                } else {
                    int l1 = e_x + 1 - shift;
                    acc = (acc == -1) ? l1 : (l1 > acc ? l1 : acc);
                }
            }
        }
    }
}

template <bool kWithTails = true>
SMX_HD void emit_record(const SelectCtx &c, Emitter &em, TrimState &ts, bool &overflow,
                        int cand_idx, const Cand &cd, const EndInfo &e1, const EndInfo &e2,
                        int sample, int resolution, int pool) {
    const Tables &t = *c.t;
    int n = c.n;
    int shift = trim_shift_get(ts, cand_idx);
    bool m1 = e1.matched, m2 = e2.matched;
    // candidate coordinates (p1/b1 were searched on the opposite strand and reversed, models.py:52-63)
    int p1s = kNone, p1e = kNone, p2s = kNone, p2e = kNone;
    if (m1) { p1s = n - e1.pe - 1 - shift; p1e = n - e1.ps - 1 - shift; }
    if (m2) { p2s = e2.ps - shift; p2e = e2.pe - shift; }
    int b1s = kNone, b1e = kNone, b2s = kNone, b2e = kNone;
    if (m1 && e1.nhits) {
        int bshift = e1.first_ss == -1 ? 0 : e1.first_ss;
        int col = lowest_bit64(e1.first_mask);
        b1s = n - (bshift + col) - 1 - shift; b1e = n - bshift - 1 - shift;
    }
    if (m2 && e2.nhits) {
        int bshift = e2.first_ss == -1 ? 0 : e2.first_ss;
        int col = lowest_bit64(e2.first_mask);
        b2s = bshift - shift; b2e = bshift + col - shift;
    }
    int s = 0, e = n;
    bool empty = false;
    if (t.trim != SMX_TRIM_NONE) {
        int ps = m1 ? p1e + 1 : 0;                 // interprimer_extent, models.py:278-287
        int pe = m2 ? p2s : n;
        if (t.trim == SMX_TRIM_PRIMERS) { s = ps; e = pe; }
        else if (t.trim == SMX_TRIM_BARCODES) {     // models.py:289-298
            s = m1 ? p1s : 0;
            e = m2 ? p2e + 1 : n;
        } else {                                    // models.py:300-319
            int fs = -1, fe = -1;
            if (kWithTails) {
                tails_fold(c, e1, true, shift, fs);
                tails_fold(c, e2, false, shift, fe);
            }
            if (fs == -1) { fs = ps - t.blen_max; if (fs < 0) fs = 0; }
            if (fe == -1) { fe = pe + t.blen_max; if (fe > n) fe = n; }
            s = fs; e = fe;
        }
        if (s >= e) empty = true;                   // demultiplex.py:47
    }
    if (em.out && em.count < em.cap) {
        smx_record &r = em.out[em.count];
        r.read = c.read + c.b->read_base;
        r.reverse = (unsigned char)cd.rc;
        r.candidate = (unsigned char)cand_idx;
        r.dist[0] = (signed char)(m1 ? e1.pd : -1);
        r.dist[1] = (signed char)((m1 && e1.nhits) ? e1.bd : -1);
        r.dist[2] = (signed char)((m2 && e2.nhits) ? e2.bd : -1);
        r.dist[3] = (signed char)(m2 ? e2.pd : -1);
        r.pad[0] = r.pad[1] = 0;
        if (empty) {
            r.sample = -1; r.resolution = SMX_RES_UNKNOWN; r.pool = -1; r.p1 = -1; r.p2 = -1;
            r.trim_start = 0; r.trim_end = n; r.trim_empty = 1;
            r.p1_loc[0] = p1s; r.p1_loc[1] = p1e; r.p2_loc[0] = p2s; r.p2_loc[1] = p2e;
            r.b1_loc[0] = b1s; r.b1_loc[1] = b1e; r.b2_loc[0] = b2s; r.b2_loc[1] = b2e;
        } else {
            int ds = (t.trim != SMX_TRIM_NONE) ? s : 0;   // trim_locations(s), demultiplex.py:76
            r.sample = sample; r.resolution = (unsigned char)resolution; r.pool = (int16_t)pool;
            r.p1 = (int16_t)(m1 ? e1.primer : -1); r.p2 = (int16_t)(m2 ? e2.primer : -1);
            r.trim_start = s; r.trim_end = e; r.trim_empty = 0;
            r.p1_loc[0] = m1 ? p1s - ds : kNone; r.p1_loc[1] = m1 ? p1e - ds : kNone;
            r.p2_loc[0] = m2 ? p2s - ds : kNone; r.p2_loc[1] = m2 ? p2e - ds : kNone;
            r.b1_loc[0] = b1s == kNone ? kNone : b1s - ds; r.b1_loc[1] = b1e == kNone ? kNone : b1e - ds;
            r.b2_loc[0] = b2s == kNone ? kNone : b2s - ds; r.b2_loc[1] = b2e == kNone ? kNone : b2e - ds;
        }
    }
    if (!empty && t.trim != SMX_TRIM_NONE) trim_shift_add(ts, cand_idx, s, overflow);
    ++em.count;
}

// resolve_specimen (demultiplex.py:541-598).
SMX_HD void resolve(const SelectCtx &c, const EndInfo &e1, const EndInfo &e2, int pair_pool,
                    int &sample, int &resolution, int &pool) {
    const Tables &t = *c.t;
    sample = -1; pool = pair_pool; resolution = SMX_RES_UNKNOWN;
    bool b1 = e1.matched && e1.nhits > 0, b2 = e2.matched && e2.nhits > 0;
    if (e1.matched && e2.matched && b1 && b2) {
        int count = 0, min_row = -1;
        for (int q1 = 0, j1 = best_at(c, e1, 0, -1); j1 >= 0; ++q1, j1 = best_at(c, e1, q1, j1))
            for (int q2 = 0, j2 = best_at(c, e2, 0, -1); j2 >= 0; ++q2, j2 = best_at(c, e2, q2, j2))
                spec_all(t, t.pb_barcode[t.pb_off[e1.primer] + j1], t.pb_barcode[t.pb_off[e2.primer] + j2],
                         e1.primer, e2.primer, count, min_row);
        if (count > 1) { sample = min_row; resolution = SMX_RES_MULTIPLE_SPECIMENS; pool = t.spec_pool[min_row]; }
        else if (count == 1) { sample = min_row; resolution = SMX_RES_FULL_MATCH; pool = t.spec_pool[min_row]; }
        return;
    }
    if (b1 && !b2 && e1.nbest == 1) {
        resolution = SMX_RES_PARTIAL_FORWARD;
        sample = (int)t.pb_barcode[t.pb_off[e1.primer] + e1.first_best];
    } else if (b2 && !b1 && e2.nbest == 1) {
        resolution = SMX_RES_PARTIAL_REVERSE;
        sample = (int)t.pb_barcode[t.pb_off[e2.primer] + e2.first_best];
    }
}

// resolve_specimen when each end has at most one equal-best barcode (then best_b1 / best_b2 are
// the single barcodes of the digests and no hit list needs walking).
SMX_HD void resolve_single(const SelectCtx &c, const EndInfo &e1, const EndInfo &e2, int pair_pool,
                           int &sample, int &resolution, int &pool) {
    const Tables &t = *c.t;
    sample = -1; pool = pair_pool; resolution = SMX_RES_UNKNOWN;
    bool b1 = e1.matched && e1.nhits > 0, b2 = e2.matched && e2.nhits > 0;
    if (e1.matched && e2.matched && b1 && b2) {
        int count = 0, min_row = -1;
        spec_all(t, t.pb_barcode[t.pb_off[e1.primer] + e1.first_best], t.pb_barcode[t.pb_off[e2.primer] + e2.first_best],
                 e1.primer, e2.primer, count, min_row);
        if (count > 1) { sample = min_row; resolution = SMX_RES_MULTIPLE_SPECIMENS; pool = t.spec_pool[min_row]; }
        else if (count == 1) { sample = min_row; resolution = SMX_RES_FULL_MATCH; pool = t.spec_pool[min_row]; }
        return;
    }
    if (b1 && !b2 && e1.nbest == 1) {
        resolution = SMX_RES_PARTIAL_FORWARD;
        sample = (int)t.pb_barcode[t.pb_off[e1.primer] + e1.first_best];
    } else if (b2 && !b1 && e2.nbest == 1) {
        resolution = SMX_RES_PARTIAL_REVERSE;
        sample = (int)t.pb_barcode[t.pb_off[e2.primer] + e2.first_best];
    }
}

struct Group {      // one dereplication group (demultiplex.py:322-382 / :416-467)
    int key;        // specimen row, or for partial groups (direction << 30 | global barcode id)
    int cand;       // best candidate so far
    int k0, k1, k2, k3;   // sort key of the best
};

SMX_HD bool key_less(int a0, int a1, int a2, int a3, const Group &g) {
    if (a0 != g.k0) return a0 < g.k0;
    if (a1 != g.k1) return a1 < g.k1;
    if (a2 != g.k2) return a2 < g.k2;
    return a3 < g.k3;
}

// Working storage of select_read.  The first pass keeps it in thread-local arrays of kSmallGroups
// entries; reads that overflow are re-run by a second GPU pass on kBigGroups-entry global scratch.
struct SelectStore {
    Group *groups; Cand *gcand;     // specimen groups
    Group *pg; Cand *pcand;         // partial (direction, barcode) groups
    int *ts_cand, *ts_shift;        // trim shifts
    int cap;
};

constexpr unsigned char kFlagDeferred = 8;   // fast-only pass: the read needs the general routine

// Whole per-read selection.  ends: cache of 2*n_primers EndInfo (index strand*n_primers+primer).
// kFastOnly = true compiles only the common single-candidate path (no grouping, no list walks,
// no TAILS fold): reads that need more come back with flags = kFlagDeferred and no record.
// NP > 0: the number of primers as a compile-time constant (the caller checked t.n_primers == NP): the slots' loads
// are then issued ahead (load_ends_ahead); NP = 0: any number, one slot at a time.
template <bool kFastOnly, int NP = 0>
SMX_HD u32 select_read_impl(const SelectCtx &c, EndInfo *ends, const SelectStore &st, smx_record *out, u32 out_cap,
                            unsigned char &flags) {
    const Tables &t = *c.t;
    const Batch &b = *c.b;
    Emitter em; em.out = out; em.cap = out_cap; em.count = 0; em.full = false;
    TrimState ts; ts.n = 0; ts.cap = st.cap; ts.cand = st.ts_cand; ts.shift = st.ts_shift;
    bool overflow = false;
    int n = c.n;
    flags = 0;
    const int kMaxGroups = st.cap;
    if ((t.min_length != -1 && n < t.min_length) || (t.max_length != -1 && n > t.max_length)) return 0;
    if (kFastOnly && t.trim == SMX_TRIM_TAILS) { flags = kFlagDeferred; return 0; }     // the tails fold walks the hit lists

    Geo g = make_geo(n, t.L);
    bool irregular = !g.regular || read_is_flagged(b, c.read);
    if constexpr (NP > 0) {
        load_ends_ahead<NP>(c, ends);
    } else {
        for (int s = 0; s < 2; ++s)
            for (int p = 0; p < t.n_primers; ++p) load_end(c, s, p, ends[s * t.n_primers + p]);
    }

    // determine_orientation (demultiplex.py:602-638).  For regular reads the head-window test of a
    // forward-sense primer equals the tail-window match of its reverse complement on the other
    // strand, so it is read off the slots; irregular reads ran it explicitly.
    int orient = 3;
    if (t.preorient) {
        int fwd = 0, rev = 0;
        for (int p = 0; p < t.n_primers; ++p) {
            int hit_s, hit_rs;   // forward-sense primer found in head of s / head of rs
            if (irregular) {
                hit_s = b.orient_hit[(u64)slot_index(t, 0, p) * b.n_pad + c.read];
                hit_rs = b.orient_hit[(u64)slot_index(t, 1, p) * b.n_pad + c.read];
            } else {
                hit_s = ends[1 * t.n_primers + p].matched;
                hit_rs = ends[0 * t.n_primers + p].matched;
            }
            if (t.p_dir[p] == 0) { fwd += hit_s; rev += hit_rs; }
            else { fwd += hit_rs; rev += hit_s; }
        }
        orient = (fwd > 0 && rev == 0) ? 1 : (rev > 0 && fwd == 0) ? 2 : 3;
    }

    // find_candidate_matches (demultiplex.py:699-741) + select_best_matches (:216-259)
    int top = 0;
    for (int pr = 0; pr < t.n_pairs; ++pr)
        for (int rc = 0; rc < 2; ++rc) {
            if (rc == 0 ? (orient == 2) : (orient == 1)) continue;
            const EndInfo &a = ends[(rc ? 0 : 1) * t.n_primers + (int)t.pair_fwd[pr]];
            const EndInfo &z = ends[(rc ? 1 : 0) * t.n_primers + (int)t.pair_rev[pr]];
            int sc = cand_score(a, z);
            if (sc > top) top = sc;
        }
    if (top == 0) {
        // no candidate at all: minimal match object (demultiplex.py:202-210)
        EndInfo none; none.matched = 0; none.nhits = 0; none.nbest = 0; none.first_best = -1; none.pd = -1; none.bd = -1;
        none.first_ss = 0; none.first_mask = 0;
        none.ps = none.pe = 0; none.strand = 0; none.primer = 0;
        Cand cd; cd.pair = -1; cd.rc = 0; cd.e1 = cd.e2 = 0;
        emit_record<!kFastOnly>(c, em, ts, overflow, 0, cd, none, none, -1, SMX_RES_UNKNOWN, -1);
        if (overflow) flags |= 2;
        return em.count;
    }

    // Walk the equal-best candidates in reference order.  Candidate index = position among *kept*
    // candidates (match_counter, demultiplex.py:702,720).
    auto for_each_top = [&](auto &&fn) {
        int kept = 0;
        for (int pr = 0; pr < t.n_pairs; ++pr)
            for (int rc = 0; rc < 2; ++rc) {
                if (rc == 0 ? (orient == 2) : (orient == 1)) continue;
                int i1 = (rc ? 0 : 1) * t.n_primers + (int)t.pair_fwd[pr];
                int i2 = (rc ? 1 : 0) * t.n_primers + (int)t.pair_rev[pr];
                int sc = cand_score(ends[i1], ends[i2]);
                if (sc == 0) continue;
                int idx = kept++;
                if (sc != top) continue;
                Cand cd; cd.pair = pr; cd.rc = rc; cd.e1 = i1; cd.e2 = i2;
                fn(idx, cd);
            }
    };

    // Fast path (the overwhelmingly common case): a single equal-best candidate whose ends carry at
    // most one equal-best barcode each.  Every branch of dereplicate_matches / resolve_specimen
    // then reduces to one record, with no grouping.
    {
        int n_top = 0, top_idx = 0;
        Cand top_cd; top_cd.pair = 0; top_cd.rc = 0; top_cd.e1 = top_cd.e2 = 0;
        for_each_top([&](int idx, const Cand &cd) { if (n_top++ == 0) { top_idx = idx; top_cd = cd; } });
        const EndInfo &a = ends[top_cd.e1];
        const EndInfo &z = ends[top_cd.e2];
        if (n_top == 1 && a.nbest <= 1 && z.nbest <= 1) {
            int sample = -1, res = SMX_RES_UNKNOWN, pool = t.pair_pool[top_cd.pair];
            bool done = false;
            if (t.derep_best && a.matched && z.matched && a.nhits > 0 && z.nhits > 0) {
                int row = spec_exact(t, t.pb_barcode[t.pb_off[a.primer] + a.first_best],
                                     t.pb_barcode[t.pb_off[z.primer] + z.first_best], a.primer, z.primer);
                if (row >= 0) { sample = row; res = SMX_RES_DEREPLICATED_FULL; pool = t.spec_pool[row]; em.full = true; done = true; }
            }
            if (!done) {
                resolve_single(c, a, z, t.pair_pool[top_cd.pair], sample, res, pool);
                if (res == SMX_RES_FULL_MATCH) em.full = true;
            }
            emit_record<!kFastOnly>(c, em, ts, overflow, top_idx, top_cd, a, z, sample, res, pool);
            if (em.full) flags |= 1;
            if (overflow) flags |= 2;
            return em.count;
        }
    }
    if (kFastOnly) { flags = kFlagDeferred; return 0; }

    if (!t.derep_best) {
        for_each_top([&](int idx, const Cand &cd) {
            int sample, res, pool;
            resolve(c, ends[cd.e1], ends[cd.e2], t.pair_pool[cd.pair], sample, res, pool);
            emit_record(c, em, ts, overflow, idx, cd, ends[cd.e1], ends[cd.e2], sample, res, pool);
            if (res == SMX_RES_FULL_MATCH) em.full = true;
        });
        if (em.full) flags |= 1;
        if (overflow) flags |= 2;
        return em.count;
    }

    // dereplicate_matches (demultiplex.py:262-393).  Groups in dict-insertion order; the None group
    // is one entry (key -1) whose members are re-walked afterwards.
    Group *groups = st.groups;
    Cand *gcand = st.gcand;
    int ng = 0;
    int none_pos = -1;
    for_each_top([&](int idx, const Cand &cd) {
        const EndInfo &a = ends[cd.e1];
        const EndInfo &z = ends[cd.e2];
        bool full = a.matched && z.matched && a.nhits > 0 && z.nhits > 0;
        bool found = false;
        if (full) {
            for (int q1 = 0, j1 = best_at(c, a, 0, -1); j1 >= 0; ++q1, j1 = best_at(c, a, q1, j1))
                for (int q2 = 0, j2 = best_at(c, z, 0, -1); j2 >= 0; ++q2, j2 = best_at(c, z, q2, j2)) {
                    int row = spec_exact(t, t.pb_barcode[t.pb_off[a.primer] + j1],
                                         t.pb_barcode[t.pb_off[z.primer] + j2], a.primer, z.primer);
                    if (row < 0) continue;
                    found = true;
                    int k0 = a.bd + z.bd, k1 = a.pd + z.pd, k2 = t.p_fidx[a.primer] + t.p_fidx[z.primer];
                    int gi = -1;
                    for (int q = 0; q < ng; ++q) if (groups[q].key == row) { gi = q; break; }
                    if (gi < 0) {
                        if (ng >= kMaxGroups) { overflow = true; continue; }
                        gi = ng++;
                        groups[gi].key = row; groups[gi].cand = idx; gcand[gi] = cd;
                        groups[gi].k0 = k0; groups[gi].k1 = k1; groups[gi].k2 = k2; groups[gi].k3 = 0;
                    } else if (key_less(k0, k1, k2, 0, groups[gi])) {
                        groups[gi].cand = idx; gcand[gi] = cd;
                        groups[gi].k0 = k0; groups[gi].k1 = k1; groups[gi].k2 = k2;
                    }
                }
        }
        if (!found && none_pos < 0) {
            if (ng >= kMaxGroups) { overflow = true; return; }
            none_pos = ng;
            groups[ng].key = -1; groups[ng].cand = -1; ++ng;
        }
    });

    for (int gi = 0; gi < ng; ++gi) {
        if (groups[gi].key >= 0) {
            const Cand &cd = gcand[gi];
            int row = groups[gi].key;
            emit_record(c, em, ts, overflow, groups[gi].cand, cd, ends[cd.e1], ends[cd.e2], row,
                        SMX_RES_DEREPLICATED_FULL, t.spec_pool[row]);
            em.full = true;
            continue;
        }
        // --- the None group: partials, unknowns, others (demultiplex.py:332-365) ---
        auto in_none = [&](const Cand &cd) {
            const EndInfo &a = ends[cd.e1];
            const EndInfo &z = ends[cd.e2];
            bool full = a.matched && z.matched && a.nhits > 0 && z.nhits > 0;
            if (!full) return true;
            for (int q1 = 0, j1 = best_at(c, a, 0, -1); j1 >= 0; ++q1, j1 = best_at(c, a, q1, j1))
                for (int q2 = 0, j2 = best_at(c, z, 0, -1); j2 >= 0; ++q2, j2 = best_at(c, z, q2, j2))
                    if (spec_exact(t, t.pb_barcode[t.pb_off[a.primer] + j1], t.pb_barcode[t.pb_off[z.primer] + j2],
                                   a.primer, z.primer) >= 0)
                        return false;
            return true;
        };
        // dereplicate_partial_matches (:396-477)
        Group *pg = st.pg;
        Cand *pcand = st.pcand;
        int npg = 0;
        for_each_top([&](int idx, const Cand &cd) {
            if (!in_none(cd)) return;
            const EndInfo &a = ends[cd.e1];
            const EndInfo &z = ends[cd.e2];
            bool hb1 = a.matched && a.nhits > 0, hb2 = z.matched && z.nhits > 0;
            if (hb1 == hb2) return;
            const EndInfo &be = hb1 ? a : z;
            int pcnt = (a.matched ? 1 : 0) + (z.matched ? 1 : 0);
            int pdist = (a.matched ? a.pd : 0) + (z.matched ? z.pd : 0);
            int fidx = (a.matched ? t.p_fidx[a.primer] : 0) + (z.matched ? t.p_fidx[z.primer] : 0);
            for (int qb = 0, jb = best_at(c, be, 0, -1); jb >= 0; ++qb, jb = best_at(c, be, qb, jb)) {
                int key = (hb1 ? 0 : (1 << 30)) | (int)t.pb_barcode[t.pb_off[be.primer] + jb];
                int q = -1;
                for (int x = 0; x < npg; ++x) if (pg[x].key == key) { q = x; break; }
                if (q < 0) {
                    if (npg >= kMaxGroups) { overflow = true; continue; }
                    q = npg++;
                    pg[q].key = key; pg[q].cand = idx; pcand[q] = cd;
                    pg[q].k0 = be.bd; pg[q].k1 = -pcnt; pg[q].k2 = pdist; pg[q].k3 = fidx;
                } else if (key_less(be.bd, -pcnt, pdist, fidx, pg[q])) {
                    pg[q].cand = idx; pcand[q] = cd;
                    pg[q].k0 = be.bd; pg[q].k1 = -pcnt; pg[q].k2 = pdist; pg[q].k3 = fidx;
                }
            }
        });
        for (int q = 0; q < npg; ++q) {
            const Cand &cd = pcand[q];
            int sample, res, pool;
            resolve(c, ends[cd.e1], ends[cd.e2], t.pair_pool[cd.pair], sample, res, pool);
            emit_record(c, em, ts, overflow, pg[q].cand, cd, ends[cd.e1], ends[cd.e2], sample, res, pool);
            if (res == SMX_RES_FULL_MATCH) em.full = true;
        }
        // dereplicate_unknown_matches (:480-538)
        Group ug; ug.cand = -1; ug.key = 0; ug.k0 = ug.k1 = ug.k2 = ug.k3 = 0;
        Cand ucand; ucand.pair = 0; ucand.rc = 0; ucand.e1 = ucand.e2 = 0;
        for_each_top([&](int idx, const Cand &cd) {
            if (!in_none(cd)) return;
            const EndInfo &a = ends[cd.e1];
            const EndInfo &z = ends[cd.e2];
            bool hb1 = a.matched && a.nhits > 0, hb2 = z.matched && z.nhits > 0;
            if (hb1 || hb2) return;
            int pcnt = (a.matched ? 1 : 0) + (z.matched ? 1 : 0);
            int pdist = (a.matched ? a.pd : 0) + (z.matched ? z.pd : 0);
            int fidx = (a.matched ? t.p_fidx[a.primer] : 999) + (z.matched ? t.p_fidx[z.primer] : 999);
            if (ug.cand < 0 || key_less(-pcnt, pdist, fidx, 0, ug)) {
                ug.cand = idx; ucand = cd; ug.k0 = -pcnt; ug.k1 = pdist; ug.k2 = fidx; ug.k3 = 0;
            }
        });
        if (ug.cand >= 0) {
            int sample, res, pool;
            resolve(c, ends[ucand.e1], ends[ucand.e2], t.pair_pool[ucand.pair], sample, res, pool);
            emit_record(c, em, ts, overflow, ug.cand, ucand, ends[ucand.e1], ends[ucand.e2], sample, res, pool);
        }
        // others: both barcodes but no specimen (:361-363)
        for_each_top([&](int idx, const Cand &cd) {
            if (!in_none(cd)) return;
            const EndInfo &a = ends[cd.e1];
            const EndInfo &z = ends[cd.e2];
            bool hb1 = a.matched && a.nhits > 0, hb2 = z.matched && z.nhits > 0;
            if (!(hb1 && hb2)) return;
            int sample, res, pool;
            resolve(c, a, z, t.pair_pool[cd.pair], sample, res, pool);
            emit_record(c, em, ts, overflow, idx, cd, a, z, sample, res, pool);
            if (res == SMX_RES_FULL_MATCH) em.full = true;
        });
    }
    if (em.full) flags |= 1;
    if (overflow) flags |= 2;
    return em.count;
}

template <int NP = 0>
SMX_HD u32 select_read(const SelectCtx &c, EndInfo *ends, const SelectStore &st, smx_record *out, u32 out_cap,
                       unsigned char &flags) {
    return select_read_impl<false, NP>(c, ends, st, out, out_cap, flags);
}

// Bytes of global scratch one read needs in the second pass.
constexpr size_t kBigScratchBytes = (size_t)kBigGroups * (2 * sizeof(Group) + 2 * sizeof(Cand) + 2 * sizeof(int)) +
                                    2 * SMX_MAX_PRIMERS * sizeof(EndInfo);

SMX_HD SelectStore big_store(unsigned char *base, EndInfo *&ends) {
    SelectStore st;
    st.groups = (Group *)base; base += kBigGroups * sizeof(Group);
    st.pg = (Group *)base; base += kBigGroups * sizeof(Group);
    st.gcand = (Cand *)base; base += kBigGroups * sizeof(Cand);
    st.pcand = (Cand *)base; base += kBigGroups * sizeof(Cand);
    st.ts_cand = (int *)base; base += kBigGroups * sizeof(int);
    st.ts_shift = (int *)base; base += kBigGroups * sizeof(int);
    ends = (EndInfo *)base;
    st.cap = kBigGroups;
    return st;
}

}  // namespace smx
