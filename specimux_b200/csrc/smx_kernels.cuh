// smx_kernels.cuh -- per-thread search routines (host/device) and their __global__ wrappers.
//
// Stage 0  stage_windows   : 2-bit / 4-bit packed reads -> 4-bit search windows of both strands
// Stage 1  primer_search   : one thread per (read window, primer): Myers HW + start recovery
// Stage 2  barcode_search  : one thread per (primer hit, barcode): Myers SHW in the flank
// Stage 3  select_reads    : per-read selection / dereplication / trimming -> smx_record
#pragma once
#include "smx_core.cuh"

namespace smx {

// ---------------------------------------------------------------------------------------------
// Stage 0.  Staged buffer of strand X = X[woff : n] (the last min(n, L) symbols), 8 symbols/word.
// Strand 1 is the reverse complement (demultiplex.py:142; Bio.Seq complement table).

// 8 consecutive 2-bit codes -> 8 nibbles
SMX_HD u32 spread2to4(u32 v) {
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    return v;
}

SMX_HD u32 bit_reverse32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}

// 16 consecutive 2-bit codes -> the same 16 bases read backwards and complemented
SMX_HD u32 revcomp16(u32 v) {
    v = bit_reverse32(v);
    v = ((v & 0x55555555u) << 1) | ((v >> 1) & 0x55555555u);
    return ~v;
}

// Reads whose primer search runs bit-sliced (regular full-length A/C/G/T window).  Only the OTHER reads get the 4-bit
// staged windows (`win`): sliced reads are served from the 2-bit words everywhere (stage 1, start recovery, the
// barcode flank), which cuts the staging kernel's stores from 120 to 40 bytes per read.
SMX_HD bool sliced_eligible(const Tables &t, const Batch &b, u32 read) {
    return t.sliced && !read_is_flagged(b, read) && (int)b.lengths[read] >= t.L;
}

// One thread stages 16 symbols of one strand: the 2-bit word win2[w2] (input of the sliced primer
// search) and the two 4-bit words win[2*w2], win[2*w2+1] (barcode stage, start recovery, classic
// primer search).  Positions past the staged length hold kSymOther in `win`.
// `src2` / `src2_origin`: where the read's 2-bit words are read from -- src2[read_word0(read) - src2_origin + i];
// the packed buffer itself (b.packed2, origin 0) or a block's shared-memory copy of its reads' words (origin =
// first word of the block's first read).
// What every 16-symbol word of one (read, strand) shares: length, window geometry, where the read's 2-bit words start.
struct StageCtx {
    int n;
    Geo g;
    bool flagged, want4;            // want4: the read also gets 4-bit windows (it is off the sliced path)
    const u32 *src;                 // the read's first 2-bit word inside src2
};

SMX_HD StageCtx stage_ctx(const Tables &t, const Batch &b, u32 read, const u32 *src2, u64 src2_origin) {
    StageCtx c;
    c.n = (int)b.lengths[read];
    c.g = make_geo(c.n, t.L);
    c.flagged = read_is_flagged(b, read);
    c.want4 = c.flagged || !(t.sliced && c.n >= t.L);
    c.src = src2 + (read_word0(b, read) - src2_origin);
    return c;
}

// w2p: the (strand, w2) word of this read in win2 (the caller steps it by n_pad per word: no 64-bit multiply per word).
SMX_HD void stage_window_word(const Tables &t, const Batch &b, const StageCtx &c, u32 read, int strand, int w2, u32 *w2p) {
    const int n = c.n;
    const Geo &g = c.g;
    int valid = g.wl - 16 * w2;
    valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
    if (!c.flagged) {
        // fast path: the staged symbols are one contiguous run of the 2-bit stream (the tail of the
        // read for strand 0, its head read backwards and complemented for strand 1)
        u32 v = 0;
        if (valid) {
            const int x0 = g.woff + 16 * w2;                  // strand coordinate of symbol 0
            const int first = strand ? (n - 1 - x0) - 15 : stored_pos(b, x0, n);   // stored index of the lowest base needed
            const int lo = first < 0 ? 0 : first;
            const u32 *src = c.src + (lo >> 4);
            const u64 pair = (u64)src[0] | ((u64)src[1] << 32);
            v = (u32)(pair >> (2 * (lo & 15)));
            if (first < 0) v <<= 2 * (-first);                // bases before the read start: masked below
            if (strand) v = revcomp16(v);
        }
        *w2p = v;
        if (!c.want4) return;                                 // sliced read: no 4-bit window (staged_sym reads win2)
        u32 *wlo = b.win + ((u64)strand * t.wpw + 2 * w2) * b.n_pad + read;
        const bool has_hi = 2 * w2 + 1 < t.wpw;
        const int vlo = valid > 8 ? 8 : valid, vhi = valid > 8 ? valid - 8 : 0;
        u32 out = spread2to4(v);
        if (vlo < 8) out |= ~0u << (4 * vlo);
        wlo[0] = out;
        if (has_hi) {
            out = spread2to4(v >> 16);
            if (vhi < 8) out |= ~0u << (4 * vhi);
            wlo[b.n_pad] = out;
        }
        return;
    }
    u32 *wlo = b.win + ((u64)strand * t.wpw + 2 * w2) * b.n_pad + read;
    const bool has_hi = 2 * w2 + 1 < t.wpw;
    *w2p = 0;                                                     // flagged reads take the classic search
    for (int h = 0; h < (has_hi ? 2 : 1); ++h) {
        u32 out = 0;
        for (int i = 0; i < 8; ++i) {
            int p = w2 * 16 + h * 8 + i;
            int c4 = kSymOther;
            if (p < g.wl) c4 = sym_at(b, read, strand, g.woff + p, n);
            out |= (u32)c4 << (4 * i);
        }
        wlo[(u64)h * b.n_pad] = out;
    }
}

// All window words of one (read, strand): the per-read part (length, geometry, flags, stream position) is worked out
// once -- computing it per 16-symbol word was three quarters of the staging kernel's instructions
// (profiles/r2_m ncu source view: 147 instructions per word, ~25 of them the word's own).
SMX_HD void stage_windows_thread(const Tables &t, const Batch &b, u32 read, int strand, const u32 *src2, u64 src2_origin) {
    const StageCtx c = stage_ctx(t, b, read, src2, src2_origin);
    u32 *w2p = b.win2 + (u64)strand * t.nw2 * b.n_pad + read;
    for (int w2 = 0; w2 < t.nw2; ++w2, w2p += b.n_pad) stage_window_word(t, b, c, read, strand, w2, w2p);
}

SMX_HD int staged_sym(const Tables &t, const Batch &b, u32 read, int strand, int p) {
    if (sliced_eligible(t, b, read)) {
        const u32 w2 = b.win2[((u64)strand * t.nw2 + (p >> 4)) * b.n_pad + read];
        return (int)((w2 >> (2 * (p & 15))) & 3);
    }
    u32 w = b.win[((u64)strand * t.wpw + (p >> 3)) * b.n_pad + read];
    return (int)((w >> (4 * (p & 7))) & 15);
}

SMX_HD int popcount32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

SMX_HD int count_leading_zeros32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}

SMX_HD int lowest_bit32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// ---------------------------------------------------------------------------------------------
// Stage 1.  match_one_end's primer search (demultiplex.py:757-766) for one (read, strand, primer).

SMX_HD u32 funnel_in_sign(u32 hist, u32 x) {      // (hist << 1) | (x >> 31): one SHF on the GPU
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, hist, 1);
#else
    return (hist << 1) | (x >> 31);
#endif
}

// Classic forward pass (one thread per (read window, primer), single-word Myers/Hyyro HW).
// Leaves "score == running best" / "score improved the running best" column histories in
// endmask / impmask and returns the best distance (m + 1 when the window is empty).
//
// Bookkeeping is kept off the critical ALU path: per column only the running minimum is updated
// (1 op) and two sign bits are shifted into history words (1 op each).  After the loop the last
// improvement marks the first equal-best end and every earlier equality bit is stale
// (primer_tail).
template <typename W>
SMX_HD int primer_forward_classic(const Tables &t, const Batch &b, u32 read, int strand, int primer, const u64 *peq) {
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    const int m = t.p_len[primer];
    const u32 slot = slot_index(t, strand, primer);
    const u32 *win = b.win + (u64)strand * t.wpw * b.n_pad + read;
    u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + read;
    u32 *imask = b.impmask + (u64)slot * t.mw * b.n_pad + read;

    W Pv = ~(W)0, Mv = 0;
    int score = m, best = m + 1;
    const int p_begin = g.start, p_end = g.wl;
    for (int mwi = 0; mwi < t.mw; ++mwi) {
        u32 eqh = 0, imh = 0;
        int done = 0;                                       // columns shifted into the histories
        for (int q = 0; q < 4; ++q) {
            int w = mwi * 4 + q;
            if (w >= t.wpw || w * 8 >= p_end) break;
            u32 word = win[(u64)w * b.n_pad];
            if (g.regular && w * 8 + 8 <= p_end) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < 8; ++i) {
                    W Eq = peq_word<W>(peq[word & 15]);
                    word >>= 4;
                    score += myers_step<W, false>(Eq, Pv, Mv);
                    imh = funnel_in_sign(imh, (u32)(score - best));          // score < best
                    best = score < best ? score : best;
                    eqh = funnel_in_sign(eqh, (u32)(score - best - 1));       // score == best
                }
                done += 8;
            } else {
                for (int i = 0; i < 8; ++i) {
                    int p = w * 8 + i;
                    int c = (int)(word & 15);
                    word >>= 4;
                    u32 im = 0, eq = 0;
                    if (p >= p_begin && p < p_end) {
                        W Eq = peq_word<W>(peq[c]);
                        score += myers_step<W, false>(Eq, Pv, Mv);
                        im = score < best ? 0x80000000u : 0u;
                        best = score < best ? score : best;
                        eq = score == best ? 0x80000000u : 0u;
                    }
                    imh = funnel_in_sign(imh, im);
                    eqh = funnel_in_sign(eqh, eq);
                }
                done += 8;
            }
        }
        // histories hold column c of this word at bit (done-1-c): reverse into bit c
        eqh = done ? bit_reverse32(eqh) >> (32 - done) : 0u;
        imh = done ? bit_reverse32(imh) >> (32 - done) : 0u;
        emask[(u64)mwi * b.n_pad] = eqh;
        imask[(u64)mwi * b.n_pad] = imh;
    }
    return best;
}

// From the column histories to the primer hit: the last improvement is the first equal-best end,
// equality bits before it are stale.  Returns the number of equal-best end locations (0 = no
// match within k); the start of the first location is recovered later by primer_start_thread.
SMX_HD int primer_tail(const Tables &t, const Batch &b, u32 read, int strand, int primer, int best) {
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    const int k = t.p_k[primer];
    const u32 slot = slot_index(t, strand, primer);
    const u64 hit_idx = (u64)slot * b.n_pad + read;
    u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + read;
    const u32 *imask = b.impmask + (u64)slot * t.mw * b.n_pad + read;
    smx_primer_hit h;
    h.distance = -1; h.n_locations = 0; h.first_start = 0; h.first_end = 0;
    int nloc = 0;
    if (best <= k) {
        int first = 0;
        for (int mwi = t.mw - 1; mwi >= 0; --mwi) {
            u32 v = imask[(u64)mwi * b.n_pad];
            if (v) { first = mwi * 32 + 31 - count_leading_zeros32(v); break; }
        }
        for (int mwi = 0; mwi < t.mw; ++mwi) {
            u32 v = emask[(u64)mwi * b.n_pad];
            int lo = mwi * 32;
            if (lo + 32 <= first) v = 0;
            else if (lo < first) v &= ~0u << (first - lo);
            emask[(u64)mwi * b.n_pad] = v;
            nloc += popcount32(v);
        }
        h.distance = (int16_t)best;
        h.n_locations = (uint16_t)nloc;
        h.first_end = g.woff + first + g.delta;
        h.first_start = h.first_end;                       // filled in by primer_start_thread
    }
    b.phit[hit_idx] = h;
    return nloc;
}

// determine_orientation (demultiplex.py:602-638), explicit form, only where the tail-window
// equivalence does not hold: reads shorter than search_len-1 (Q1) or with non-ACGT symbols.
template <typename W>
SMX_HD void primer_orient_explicit(const Tables &t, const Batch &b, u32 read, int strand, int primer, const u64 *peq_fw) {
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    const int m = t.p_len[primer], k = t.p_k[primer];
    unsigned char ohit = 0;
    if (t.preorient && (!g.regular || read_is_flagged(b, read))) {
        int cols = n < t.L ? n : t.L;
        W pv = ~(W)0, mv = 0;
        int sc = m, bst = m + 1;
        for (int x = 0; x < cols; ++x) {
            W Eq = peq_word<W>(peq_fw[sym_at(b, read, strand, x, n)]);
            sc += myers_step<W, false>(Eq, pv, mv);
            if (sc < bst) bst = sc;
        }
        ohit = bst <= k;
    }
    b.orient_hit[(u64)slot_index(t, strand, primer) * b.n_pad + read] = ohit;
}

// The whole classic stage-1 search of one (read, strand, primer).
template <typename W>
SMX_HD int primer_search_thread(const Tables &t, const Batch &b, u32 read, int strand, int primer,
                                const u64 *peq, const u64 *peq_fw) {
    int best = primer_forward_classic<W>(t, b, read, strand, primer, peq);
    int nloc = primer_tail(t, b, read, strand, primer, best);
    primer_orient_explicit<W>(t, b, read, strand, primer, peq_fw);
    return nloc;
}

// ---------------------------------------------------------------------------------------------
// Stage 1, sliced form: the HW search evaluated BIT-SLICED ACROSS READS.  One thread owns a group
// of 32 reads (bit r of every word = read 32*group + r) and one (strand, primer); a Myers/Hyyro
// cell update (6 three-input logic ops) advances cell (i, j) of all 32 windows at once, so a
// column of an m-row pattern costs 6m + ~20 integer ops per 32 reads instead of ~29 per read.
// Only reads whose window is the regular full-length A/C/G/T case take part (n >= search_len,
// not on the 4-bit side stream); the others run the classic per-thread search.
//
// Per block of 16 columns: the group's 32 two-bit window words are transposed in registers into
// 32 bit-planes (2 per column), the 16 columns are evaluated, each column's "equal to the running
// best" / "improved the running best" planes replace the consumed input planes, and one more
// transpose turns them into a per-read word of interleaved (equal, improved) bits.  The running
// best itself is never materialised: best = m - (number of improvements).

// Bit-sliced full adder carry / majority on planes.
SMX_HD u32 maj3(u32 a, u32 b, u32 c) { return (a & b) | (a & c) | (b & c); }

// In-register 32x32 bit-matrix transpose: afterwards bit r of a[c] is the former bit c of a[r].
SMX_HD void transpose32(u32 (&a)[32]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int s = 0; s < 5; ++s) {
        const int j = 16 >> s;
        const u32 m = s == 0 ? 0x0000FFFFu : s == 1 ? 0x00FF00FFu : s == 2 ? 0x0F0F0F0Fu : s == 3 ? 0x33333333u : 0x55555555u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 32; ++k) {
            if (k & j) continue;
            u32 x = ((a[k] >> j) ^ a[k + j]) & m;
            a[k + j] ^= x;
            a[k] ^= x << j;
        }
    }
}

// Byte offsets (into the thread's code-plane scratch) of the Eq plane of every pattern row.
struct RowOffsets { unsigned short off[32]; };

constexpr int kSlicedCodes = 15;        // scratch planes per thread: one per IUPAC pattern code

template <int M> struct SlicedBits { static constexpr int NB = M < 2 ? 1 : M < 4 ? 2 : M < 8 ? 3 : M < 16 ? 4 : M < 32 ? 5 : 6; };

// sp: this thread's kSlicedCodes code planes, sa: its 32-word plane buffer followed by 16 more
// planes for the "improved" results (all with element stride STRIDE: shared memory columns on the
// GPU).  The column loop and the two transposes are
// kept ROLLED: the generation that unrolled 16 columns was instruction-fetch bound (ncu:
// stall_no_instruction the top stall at 40 % issue utilisation).
template <int M, int STRIDE>
SMX_HD void primer_sliced_thread(const Tables &t, const Batch &b, u32 group, int strand, int primer,
                                 const RowOffsets &ro, bool degenerate, u32 *sp, u32 *sa) {
    constexpr int NB = SlicedBits<M>::NB;
    const u32 slot = slot_index(t, strand, primer);
    u32 VP[M], VM[M];                                   // vertical deltas of the previous column
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < M; ++i) { VP[i] = ~0u; VM[i] = 0u; }        // D[i][0] = i
    u32 d[NB];                                          // score - running best, bit-sliced
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < NB; ++q) d[q] = 0u;
    u32 zero = ~0u;                                     // score == running best
    for (int blk = 0; blk < t.nw2; ++blk) {
        u32 a[32];
        const u32 *src = b.win2 + ((u64)strand * t.nw2 + blk) * b.n_pad + (u64)group * 32;
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int r = 0; r < 32; r += 4) {
            uint4 v = *reinterpret_cast<const uint4 *>(src + r);
            a[r] = v.x; a[r + 1] = v.y; a[r + 2] = v.z; a[r + 3] = v.w;
        }
#else
        for (int r = 0; r < 32; ++r) a[r] = src[r];
#endif
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int phase = 0; phase < 2; ++phase) {
            // phase 0: reads x code bits -> code-bit planes x reads; phase 1: column results -> per-read words
            transpose32(a);
            if (phase) break;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int r = 0; r < 32; ++r) sa[r * STRIDE] = a[r];     // sa[2c] / sa[2c+1]: low / high code bit of column c
            int ncols = t.L - blk * 16;
            ncols = ncols > 16 ? 16 : ncols;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int c = 0; c < ncols; ++c) {
                const u32 b0 = sa[(2 * c) * STRIDE], b1 = sa[(2 * c + 1) * STRIDE];
                sp[0 * STRIDE] = ~(b0 | b1);            // A
                sp[1 * STRIDE] = b0 & ~b1;              // C
                sp[2 * STRIDE] = b1 & ~b0;              // G
                sp[3 * STRIDE] = b0 & b1;               // T
                if (degenerate) {                       // uniform; IUPAC base sets (constants.py:13-20)
                    sp[4 * STRIDE] = ~b0;               // R = A|G
                    sp[5 * STRIDE] = b0;                // Y = C|T
                    sp[6 * STRIDE] = b0 ^ b1;           // S = C|G
                    sp[7 * STRIDE] = ~(b0 ^ b1);        // W = A|T
                    sp[8 * STRIDE] = b1;                // K = G|T
                    sp[9 * STRIDE] = ~b1;               // M = A|C
                    sp[10 * STRIDE] = b0 | b1;          // B = not A
                    sp[11 * STRIDE] = ~(b0 & ~b1);      // D = not C
                    sp[12 * STRIDE] = ~(b1 & ~b0);      // H = not G
                    sp[13 * STRIDE] = ~(b0 & b1);       // V = not T
                    sp[14 * STRIDE] = ~0u;              // N
                }
                u32 Ph = 0u, Mh = 0u;                   // row 0: D[0][j] = 0 (HW)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < M; ++i) {
                    const u32 Eq = *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(sp) + ro.off[i]);
                    // Pv & Mv = 0 and Ph & Mh = 0 (a delta is +1, 0 or -1), so the two "match or negative delta"
                    // terms Eq | Mv and Eq | Mh can share ONE three-input term X = Eq | Mv | Mh: wherever X adds a
                    // bit to one of them, the factor it is combined with is zero.  5 LOP3 per cell instead of 6.
                    const u32 Pv = VP[i], Mv = VM[i];
                    const u32 X = Eq | Mv | Mh;
                    VP[i] = Mh | ~(X | Ph);
                    VM[i] = Ph & X;
                    const u32 nPh = Mv | ~(X | Pv);
                    Mh = Pv & X;
                    Ph = nPh;
                }
                // score += Ph - Mh against the running best: the difference saturates at 0 from below
                const u32 imp = Mh & zero, dec = Mh & ~zero;
                u32 x = Ph | dec;
                u32 carry = d[0] & x;
                d[0] ^= x;
                u32 any = d[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int q = 1; q < NB; ++q) {          // + (dec ? all-ones : 0) + carry
                    const u32 sum = d[q] ^ dec ^ carry;
                    carry = (d[q] & dec) | (d[q] & carry) | (dec & carry);
                    d[q] = sum;
                    any |= sum;
                }
                zero = ~any;
                // results: "equal" plane of column c over the (already consumed) input plane c,
                // "improved" plane into the 16 extra planes -- so that after the second transpose a
                // read's word is (improved bits << 16) | equal bits with no bit shuffling left to do
                sa[c * STRIDE] = zero;
                sa[(32 + c) * STRIDE] = imp;
            }
            for (int c = ncols; c < 16; ++c) { sa[c * STRIDE] = 0u; sa[(32 + c) * STRIDE] = 0u; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int r = 0; r < 16; ++r) { a[r] = sa[r * STRIDE]; a[16 + r] = sa[(32 + r) * STRIDE]; }
        }
        // a[r]: read r's word of the block's 16 columns: bit c = equal-best so far, bit 16 + c = improvement
        u32 *dst = b.tmix + ((u64)slot * t.nw2 + blk) * b.n_pad + (u64)group * 32;
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int r = 0; r < 32; r += 4) *reinterpret_cast<uint4 *>(dst + r) = make_uint4(a[r], a[r + 1], a[r + 2], a[r + 3]);
#else
        for (int r = 0; r < 32; ++r) dst[r] = a[r];
#endif
    }
}

// Stage 1 per (read, strand, primer) after the sliced pass: eligible reads decode their column
// histories from tmix, every other read runs the classic search.  Returns the number of
// equal-best end locations.
// `ev` (optional, kFinishMaskWords words): receives the slot's equal-best end mask when it is short enough
// to stay in registers, `have_ev` says so; the caller then writes the work entries from registers
// instead of reading the mask back from memory (write_entries).
constexpr int kFinishMaskWords = 4, kFinishMixWords = 8;

template <typename W>
SMX_HD int primer_finish_thread(const Tables &t, const Batch &b, u32 read, int strand, int primer,
                                const u64 *peq, const u64 *peq_fw, u32 *ev = nullptr, bool *have_ev = nullptr) {
    if (have_ev) *have_ev = false;
    if (!sliced_eligible(t, b, read)) return primer_search_thread<W>(t, b, read, strand, primer, peq, peq_fw);
    const u32 slot = slot_index(t, strand, primer);
    const u64 hit_idx = (u64)slot * b.n_pad + read;
    u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + read;
    const u32 *mix = b.tmix + (u64)slot * t.nw2 * b.n_pad + read;
    const bool in_regs = t.nw2 <= kFinishMixWords && t.mw <= kFinishMaskWords;
    // pass 1: number of improvements (-> best) and the column of the last one (-> first equal-best end);
    // the words stay in registers for pass 2 when there are few enough of them
    u32 mw_[kFinishMixWords];
    int improvements = 0, first = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int blk = 0; blk < kFinishMixWords; ++blk) {
        mw_[blk] = (in_regs && blk < t.nw2) ? mix[(u64)blk * b.n_pad] : 0u;
        const u32 im = mw_[blk] >> 16;
        improvements += popcount32(im);
        if (im) first = blk * 16 + 31 - count_leading_zeros32(im);
    }
    if (!in_regs)
        for (int blk = 0; blk < t.nw2; ++blk) {
            const u32 im = mix[(u64)blk * b.n_pad] >> 16;
            improvements += popcount32(im);
            if (im) first = blk * 16 + 31 - count_leading_zeros32(im);
        }
    const int best = (int)t.p_len[primer] - improvements;
    b.orient_hit[hit_idx] = 0;                          // eligible reads never need the explicit test
    smx_primer_hit h;
    h.distance = -1; h.n_locations = 0; h.first_start = 0; h.first_end = 0;
    int nloc = 0;
    const bool hit = best <= t.p_k[primer];
    if (in_regs) {
        // pass 2 from registers: equality bits from the last improvement on are the equal-best ends
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int mwi = 0; mwi < kFinishMaskWords; ++mwi) {
            if (mwi >= t.mw) break;
            u32 v = 0;
            if (hit) {
                v = (mw_[2 * mwi] & 0xFFFFu) | (mw_[2 * mwi + 1] << 16);
                const int lo = mwi * 32;
                if (lo + 32 <= first) v = 0;
                else if (lo < first) v &= ~0u << (first - lo);
            }
            emask[(u64)mwi * b.n_pad] = v;
            nloc += popcount32(v);
            if (ev) ev[mwi] = v;
        }
        if (have_ev && ev) *have_ev = true;
    } else if (hit) {
        for (int mwi = 0; mwi < t.mw; ++mwi) {
            u32 v = mix[(u64)(2 * mwi) * b.n_pad] & 0xFFFFu;
            if (2 * mwi + 1 < t.nw2) v |= mix[(u64)(2 * mwi + 1) * b.n_pad] << 16;
            const int lo = mwi * 32;
            if (lo + 32 <= first) v = 0;
            else if (lo < first) v &= ~0u << (first - lo);
            emask[(u64)mwi * b.n_pad] = v;
            nloc += popcount32(v);
        }
    } else {
        for (int mwi = 0; mwi < t.mw; ++mwi) emask[(u64)mwi * b.n_pad] = 0;
    }
    if (hit) {
        const Geo g = make_geo((int)b.lengths[read], t.L);
        h.distance = (int16_t)best;
        h.n_locations = (uint16_t)nloc;
        h.first_end = g.woff + first + g.delta;
        h.first_start = h.first_end;                    // filled in by primer_start_thread
    }
    b.phit[hit_idx] = h;
    return nloc;
}

// Start recovery over the staged 2-bit window (reads without non-ACGT symbols): same recurrence as
// hw_start_back, but the up to m + best symbols before the end are cut out of the window words ONCE into a
// 64-bit register (32 symbols, the end on top) and shifted out two bits per column -- the per-column window
// load of the first form was 30 % of the start-recovery kernel's instructions (profiles/r2_a_*).
// win2: the strand's 2-bit window words of this read, `stride` words apart.
template <typename W>
SMX_HD int hw_start_back_win2(const u64 *peq_rev, int m, int best, int e_pos, int start_pos, const u32 *win2, u64 stride) {
    W Pv = pattern_mask<W>(m), Mv = 0;
    int score = m;
    int cols = e_pos - start_pos + 1;
    const int lim = m + best;
    if (cols > lim) cols = lim;
    int last = m - 1;
    for (int j0 = 0; j0 < cols; j0 += 32) {
        // symbols e-31 .. e of the window with e = e_pos - j0, symbol e in the top two bits
        const int e = e_pos - j0, hi = e >> 4, sh = 30 - 2 * (e & 15);
        const u32 w0 = win2[(u64)hi * stride];
        const u32 w1 = hi >= 1 ? win2[(u64)(hi - 1) * stride] : 0u;
        const u32 w2 = hi >= 2 ? win2[(u64)(hi - 2) * stride] : 0u;
        u64 X = (((u64)w0 << 32) | w1) << sh;
        if (sh) X |= (u64)(w2 >> (32 - sh));
        const int jn = cols - j0 < 32 ? cols - j0 : 32;
        for (int j = 0; j < jn; ++j) {
            const int c = (int)(X >> 62);
            X <<= 2;
            const W Eq = peq_word<W>(peq_rev[c]);
            score += myers_step<W, true>(Eq, Pv, Mv);
            if (score == best) last = j0 + j;
        }
    }
    return last;
}

// Start recovery for the first equal-best end (staged position `first`) of a matched slot (edlib: reverse SHW
// pass with k = best, the LAST equal-best reverse end wins = longest alignment).
template <typename W>
SMX_HD void primer_start_slot(const Tables &t, const Batch &b, u32 read, int strand, int primer, int first, const u64 *peq_rev) {
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    smx_primer_hit &h = b.phit[(u64)slot_index(t, strand, primer) * b.n_pad + read];
    int back;
    if (!read_is_flagged(b, read)) {
        back = hw_start_back_win2<W>(peq_rev, t.p_len[primer], h.distance, first, g.start,
                                     b.win2 + (u64)strand * t.nw2 * b.n_pad + read, b.n_pad);
    } else {
        auto load = [&](int p) { return staged_sym(t, b, read, strand, p); };
        back = hw_start_back<W>(peq_rev, t.p_len[primer], h.distance, first, g.start, load);
    }
    h.first_start = h.first_end - back;
}

// The same over the compact work-entry lists: one thread per work entry; only a read's first entry does the work.
template <typename W>
SMX_HD void primer_start_thread(const Tables &t, const Batch &b, u32 slot, u32 entry, const u64 *peq_rev) {
    const u32 read = b.ent_read[(u64)slot * b.e_cap + entry];
    const u64 hit_idx = (u64)slot * b.n_pad + read;
    if (b.ent_base[hit_idx] != entry) return;              // not the first location of its read
    const int strand = (int)slot / t.n_primers, primer = (int)slot % t.n_primers;
    primer_start_slot<W>(t, b, read, strand, primer, (int)b.ent_pos[(u64)slot * b.e_cap + entry], peq_rev);
}

// ---------------------------------------------------------------------------------------------
// Start recovery, sliced form: the reverse SHW pass of 32 work entries at once (bit r of every word = entry
// e0 + r of one slot), the same bit-sliced cell as the forward search.  Per entry the up to 32 window symbols that
// end at its first equal-best end are cut out in REVERSE order (symbol first - j in column j), the 32 entries'
// words are transposed into code-bit planes, the m x (m + k) cells are evaluated (5 LOP3 per cell for 32 entries),
// the score of the last row is tracked on six bit-planes and compared with the entries' best distances, and one
// more transpose turns the per-column "score == best" planes into a word per entry whose top set bit is the
// number of columns walked back (edlib: the LAST equal-best reverse end = the longest alignment).  ~250 thread
// instructions per entry against ~1,070 for the single-word form (profiles/r2_c_*), which stays in charge of
// reads on the 4-bit side stream and of primers with m + k > 32.

SMX_HD bool start_sliced_ok(const Tables &t, int primer) {
    return t.sliced && !t.p_sw[primer] && (int)t.p_len[primer] + (int)t.p_k[primer] <= 32;
}

// 16 consecutive 2-bit symbols starting at staged position p0 (may be negative: symbols before the window read 0)
SMX_HD u32 win2_extract16(const u32 *win2, u64 stride, int nw2, int p0) {
    const int idx = p0 >> 4, sh = 2 * (p0 & 15);              // arithmetic shift: floor for negative p0
    const u32 lo = (idx >= 0 && idx < nw2) ? win2[(u64)idx * stride] : 0u;
    const u32 hi = (idx + 1 >= 0 && idx + 1 < nw2) ? win2[(u64)(idx + 1) * stride] : 0u;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}

// the 16 two-bit groups of a word in reverse order
SMX_HD u32 reverse_pairs16(u32 v) {
    v = bit_reverse32(v);
    return ((v & 0x55555555u) << 1) | ((v >> 1) & 0x55555555u);
}

constexpr int kStartPlanes = 128;       // per-thread plane buffer of the sliced start recovery (words): XH, XL, ACT, BEST

// Gather of ONE work entry for the sliced start recovery: the reversed window words (symbol first - j in column j,
// 16 columns per word), the active-column mask and the best distance (63 = the entry takes no part: not the first
// location of its read, beyond the list, or a read on the 4-bit side stream).
SMX_HD void start_gather_entry(const Tables &t, const Batch &b, u32 slot, u32 e, u32 n_entries, int m,
                               u32 &xh, u32 &xl, u32 &act, u32 &best) {
    xh = 0; xl = 0; act = 0; best = 63;
    if (e >= n_entries) return;
    const u32 read = b.ent_read[(u64)slot * b.e_cap + e];
    const u64 hit_idx = (u64)slot * b.n_pad + read;
    if (b.ent_base[hit_idx] != e || read_is_flagged(b, read)) return;
    const int strand = (int)slot / t.n_primers;
    const smx_primer_hit h = b.phit[hit_idx];
    const Geo g = make_geo((int)b.lengths[read], t.L);
    const int first = h.first_end - g.woff - g.delta;
    int cols = first - g.start + 1;
    const int lim = m + (int)h.distance;
    if (cols > lim) cols = lim;
    const u32 *w2 = b.win2 + (u64)strand * t.nw2 * b.n_pad + read;
    xh = reverse_pairs16(win2_extract16(w2, b.n_pad, t.nw2, first - 15));
    xl = reverse_pairs16(win2_extract16(w2, b.n_pad, t.nw2, first - 31));
    act = cols >= 32 ? ~0u : ((1u << cols) - 1u);
    best = (u32)h.distance;
}

// `last` = number of columns walked back from the first equal-best end (edlib: the LAST reverse end at the best score)
SMX_HD void start_store_entry(const Tables &t, const Batch &b, u32 slot, u32 e, int last) {
    const u32 read = b.ent_read[(u64)slot * b.e_cap + e];
    smx_primer_hit &h = b.phit[(u64)slot * b.n_pad + read];
    h.first_start = h.first_end - last;
}

// The bit-sliced reverse SHW pass over the 32 gathered entries of one thread.  sp: kSlicedCodes code planes; sa:
// kStartPlanes words (element stride STRIDE) holding the gathered XH[32], XL[32], ACT[32], BEST[32] words; on return
// XH[r] = columns walked back for entry r, or 0xFFFFFFFF when the entry takes no part.  ro: byte offsets of the Eq
// plane of every row of the REVERSED primer_rc.
template <int M, int STRIDE>
SMX_HD void primer_start_sliced_thread(int ncols, const RowOffsets &ro, bool degenerate, u32 *sp, u32 *sa) {
    u32 *XH = sa, *XL = sa + 32 * STRIDE, *ACT = sa + 64 * STRIDE, *BP = sa + 96 * STRIDE;
    u32 a[32];
    u32 part = 0;                       // entries that take part (best != 63)
    // ---- entries x bits -> bit planes x entries
    for (int which = 0; which < 4; ++which) {
        u32 *buf = which == 0 ? XH : which == 1 ? XL : which == 2 ? ACT : BP;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 32; ++r) a[r] = buf[r * STRIDE];
        if (which == 3) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int r = 0; r < 32; ++r) part |= (a[r] != 63u ? 1u : 0u) << r;
        }
        transpose32(a);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 32; ++r) buf[r * STRIDE] = a[r];
    }
    // ---- the reverse SHW pass (D[0][j] = j, D[i][0] = i), score of row m on six planes
    u32 VP[M], VM[M];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < M; ++i) { VP[i] = ~0u; VM[i] = 0u; }
    u32 sc[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int l = 0; l < 6; ++l) sc[l] = (M >> l) & 1 ? ~0u : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int c = 0; c < ncols; ++c) {
        const u32 *X = c < 16 ? XH : XL;
        const u32 b0 = X[(2 * (c & 15)) * STRIDE], b1 = X[(2 * (c & 15) + 1) * STRIDE];
        sp[0 * STRIDE] = ~(b0 | b1);            // A
        sp[1 * STRIDE] = b0 & ~b1;              // C
        sp[2 * STRIDE] = b1 & ~b0;              // G
        sp[3 * STRIDE] = b0 & b1;               // T
        if (degenerate) {                       // IUPAC base sets (constants.py:13-20)
            sp[4 * STRIDE] = ~b0; sp[5 * STRIDE] = b0; sp[6 * STRIDE] = b0 ^ b1; sp[7 * STRIDE] = ~(b0 ^ b1);
            sp[8 * STRIDE] = b1; sp[9 * STRIDE] = ~b1; sp[10 * STRIDE] = b0 | b1; sp[11 * STRIDE] = ~(b0 & ~b1);
            sp[12 * STRIDE] = ~(b1 & ~b0); sp[13 * STRIDE] = ~(b0 & b1); sp[14 * STRIDE] = ~0u;
        }
        u32 Ph = ~0u, Mh = 0u;                  // row 0: D[0][j] = j (SHW)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < M; ++i) {
            const u32 Eq = *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(sp) + ro.off[i]);
            const u32 Pv = VP[i], Mv = VM[i];
            const u32 X3 = Eq | Mv | Mh;
            VP[i] = Mh | ~(X3 | Ph);
            VM[i] = Ph & X3;
            const u32 nPh = Mv | ~(X3 | Pv);
            Mh = Pv & X3;
            Ph = nPh;
        }
        // score += Ph - Mh in one ripple (+1 as the carry-in, -1 as an all-ones addend); equal to the entry's best?
        u32 carry = Ph, diff = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < 6; ++l) {
            const u32 v = sc[l];
            sc[l] = v ^ Mh ^ carry;
            carry = maj3(v, Mh, carry);
            diff |= sc[l] ^ BP[l * STRIDE];
        }
        // the consumed input planes of column c make room for its result (slot c <= 2c)
        const u32 eq = ~diff & ACT[c * STRIDE];
        if (c < 16) XH[c * STRIDE] = eq; else XL[(c - 16) * STRIDE] = eq;
    }
    // ---- per-column planes -> a word per entry: the top set bit is the last column whose score equals the best
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 32; ++c) a[c] = c < ncols ? (c < 16 ? XH[c * STRIDE] : XL[(c - 16) * STRIDE]) : 0u;
    transpose32(a);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 32; ++r)
        XH[r * STRIDE] = ((part >> r) & 1u) ? (a[r] ? (u32)(31 - count_leading_zeros32(a[r])) : (u32)(M - 1)) : 0xFFFFFFFFu;
}

// ---------------------------------------------------------------------------------------------
// Stage 2.  match_one_end's barcode loop (demultiplex.py:781-815): for one matched (read, strand,
// primer) slot and one bword (<= 32 barcodes of one length), the SHW distance of every barcode at
// every equal-best primer end; strictly-smallest distance over the ends wins (first end on ties).
//
// The DP is evaluated BIT-SLICED ACROSS BARCODES: bit q of every machine word belongs to barcode q,
// and one Myers/Hyyro cell update (6 three-input logic ops on the +-1 delta planes) advances cell
// (i, j) of all 32 barcodes at once.  Only the Ukkonen band |i - j| <= k is evaluated (anything
// outside costs > k); cells beyond the band edge are assumed one worse than their in-band
// neighbour, which keeps every value <= k exact.  Row m is read out through a small bit-sliced
// counter anchored on the band's lower diagonal (D[k][0] = k).
// Plain-logic restatement of edlib SHW (alignment.py:42 with mode='SHW') for bounded k.

template <int K> struct BitSliced {
    static constexpr int NT = 2 * K + 1;                       // band width
    static constexpr int NB = K <= 1 ? 3 : (K <= 4 ? 4 : (K <= 10 ? 5 : 6));   // counter bits: values up to 3K+1 matter

    struct Out { u32 rp[NT > 1 ? NT - 1 : 1], rm[NT > 1 ? NT - 1 : 1], cnt[NB], over; };

    // One cell of the automaton for 32 barcodes.
    static SMX_HD void cell(u32 Eq, u32 Pv, u32 Mv, u32 Ph, u32 Mh, u32 &oPv, u32 &oMv, u32 &oPh, u32 &oMh, u32 &inc) {
        // Pv & Mv = 0 and Ph & Mh = 0, so Eq | Mv and Eq | Mh share one term X = Eq | Mv | Mh (see the sliced
        // primer search): 5 LOP3 per cell, and the diagonal increment is simply ~X.
        const u32 X = Eq | Mv | Mh;
        oPv = Mh | ~(X | Ph);
        oMv = Ph & X;
        oPh = Mv | ~(X | Pv);
        oMh = Pv & X;
        inc = ~X;                                              // D[i][j] - D[i-1][j-1]
    }

    // beq_rows: table rows of this bword ([i][16]); fsym[j - 1]: 4-bit symbol of flank column j (1-based),
    // kSymOther beyond the flank (m + K entries).
    static SMX_HD void run(const u32 *beq_rows, int m, const unsigned char *fsym, Out &o) {
        u32 HP[NT], HM[NT];                                    // horizontal deltas of the previous row
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 0; t < NT; ++t) { HP[t] = ~0u; HM[t] = 0; }    // row 0: D[0][j] = j
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < NB; ++b) o.cnt[b] = (K >> b) & 1 ? ~0u : 0u;    // anchor D[K][0] = K
        o.over = 0;
        for (int i = 1; i <= m; ++i) {
            const u32 *row = beq_rows + (i - 1) * 16;
            u32 Pv = ~0u, Mv = 0;                              // column 0 / below the band: +1
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < NT; ++t) {
                int j = i - K + t;
                if (j < 1) continue;                           // uniform over the warp
                u32 Ph = t < NT - 1 ? HP[t + 1] : ~0u;          // above the band: +1
                u32 Mh = t < NT - 1 ? HM[t + 1] : 0u;
                u32 Eq = row[j <= m + K ? fsym[j - 1] : (unsigned char)kSymOther];
                u32 nPv, nMv, nPh, nMh, inc;
                cell(Eq, Pv, Mv, Ph, Mh, nPv, nMv, nPh, nMh, inc);
                if (t == 0) {                                   // lower diagonal: anchor += inc
                    u32 x = inc;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int b = 0; b < NB; ++b) { u32 c = o.cnt[b] & x; o.cnt[b] ^= x; x = c; }
                    o.over |= x;
                }
                HP[t] = nPh; HM[t] = nMh;                       // becomes index t-1+1 of the next row
                Pv = nPv; Mv = nMv;
            }
            // shift the window by one column for the next row: entry t of the next row is column
            // (i+1)-K+t = this row's t+1
            // (done implicitly: next row reads HP[t+1], this row wrote HP[t] for its own column t)
        }
        // After the last row HP[t]/HM[t] hold Delta-h of row m at column m-K+t.
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 1; t < NT; ++t) { o.rp[t - 1] = HP[t]; o.rm[t - 1] = HM[t]; }
    }

};
// value <= K on NB bit-planes (K compile-time)
template <int K, int NB> SMX_HD u32 planes_le(const u32 *v) {
    u32 gt = 0, eq = ~0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = NB - 1; b >= 0; --b) {
        if ((K >> b) & 1) eq &= v[b];
        else { gt |= eq & v[b]; eq &= ~v[b]; }
    }
    return ~gt;
}

constexpr int kMaxWordHits = 32;

// Eq words of one (row, symbol) of a task table: S consecutive words, one vector load on the GPU.
template <int S> SMX_HD void load_eq(const u32 *p, u32 (&eq)[S]) {
#if defined(__CUDA_ARCH__)
    if (S == 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        eq[0] = v.x; eq[1 % S] = v.y; eq[2 % S] = v.z; eq[3 % S] = v.w;
    } else if (S == 2) {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        eq[0] = v.x; eq[1 % S] = v.y;
    } else {
        eq[0] = *p;
    }
#else
    for (int q = 0; q < S; ++q) eq[q] = p[q];
#endif
}

// Result of the small form for one bword: D[m][m-K] as five bit-planes (values <= 16 there), and the horizontal
// deltas of row m to its right.
template <int K> struct SmallOut {
    static constexpr int NT = 2 * K + 1;
    u32 v0[5];
    u32 rp[NT > 1 ? NT - 1 : 1], rm[NT > 1 ? NT - 1 : 1];
};

// Carry-save count of one-bit planes: value = sum s[l] << l + sum p[l] << l.  Two planes of one weight meet in a
// full adder (2 LOP3) whose carry moves up a level, so n planes cost ~2n LOP3 instead of the 8n of rippling
// every plane through a 4-bit counter (the per-row diagonal increments were 5 % of the barcode kernel's
// instructions, profiles/r2_c_*).  `r` = planes added so far (a compile-time value in the unrolled row loop).
struct CsaCount {
    u32 s[5], p[4];
    SMX_HD void init() {
        for (int l = 0; l < 5; ++l) s[l] = 0;
        for (int l = 0; l < 4; ++l) p[l] = 0;
    }
    SMX_HD void add(u32 x, int r) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < 4; ++l) {
            if (((r >> l) & 1) == 0) { p[l] = x; return; }
            const u32 c = maj3(s[l], p[l], x);
            s[l] = s[l] ^ p[l] ^ x;
            p[l] = 0;
            x = c;
        }
        s[4] ^= x;                                  // at most 16 planes: no carry out of the top level
    }
    // s + p + constant k -> five planes
    SMX_HD void finish(int k, u32 (&v)[5]) const {
        u32 c = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < 5; ++l) {               // s + p
            const u32 pl = l < 4 ? p[l] : 0u;
            v[l] = s[l] ^ pl ^ c;
            c = maj3(s[l], pl, c);
        }
        c = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < 5; ++l) {               // + k (constant planes fold away)
            const u32 kl = (k >> l) & 1 ? ~0u : 0u;
            const u32 a = v[l];
            v[l] = a ^ kl ^ c;
            c = maj3(a, kl, c);
        }
    }
};

// Small form of the automaton (m + K <= 16) for the NWQ bwords of one task at once.  The whole flank sits in
// one 64-bit register F (nibble j-1 = symbol of column j, kSymOther beyond the flank).  One table address is
// formed per COLUMN (row 0 of that column's symbol); the rows are fully unrolled, so every cell's Eq words are
// one vector load at a compile-time offset from its column's address, shared by the task's words, and the
// NWQ independent automata interleave in the instruction stream.  Rows beyond m leave the unrolled sequence
// through one early exit, so the horizontal-delta registers are renamed from row to row without moves.
// col[j]: address of row 0 of flank column j + 1 in the task table; ROWSTRIDE: words from one row to the next.
template <int K, int NWQ, int MF, int ROWSTRIDE>
SMX_HD void bitsliced_small_rows_cols(const u32 *const (&col)[16], int m, SmallOut<K> (&o)[NWQ]) {
    typedef BitSliced<K> BS;
    constexpr int NT = BS::NT, S = NWQ == 1 ? 1 : NWQ == 2 ? 2 : 4;
    constexpr int kRows = 16 - K > 0 ? 16 - K : 0;
    u32 HP[NWQ][NT], HM[NWQ][NT];
    CsaCount cnt[NWQ];                                      // D[i][i-K] - K along the band's lower diagonal
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < NWQ; ++q) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 0; t < NT; ++t) { HP[q][t] = ~0u; HM[q][t] = 0; }
        cnt[q].init();
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i <= (MF ? MF : kRows); ++i) {
        if (!MF && i > m) break;                            // uniform over the block; MF: the length is a template argument
        u32 Pv[NWQ], Mv[NWQ];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < NWQ; ++q) { Pv[q] = ~0u; Mv[q] = 0; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 0; t < NT; ++t) {
            const int j = i - K + t;                        // compile-time
            if (j < 1 || j > 16) continue;
            u32 eq[S];
            load_eq<S>(col[j - 1] + (i - 1) * ROWSTRIDE, eq);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int q = 0; q < NWQ; ++q) {
                const u32 Ph = t < NT - 1 ? HP[q][t + 1] : ~0u;
                const u32 Mh = t < NT - 1 ? HM[q][t + 1] : 0u;
                u32 nPv, nMv, nPh, nMh, inc;
                BS::cell(eq[q], Pv[q], Mv[q], Ph, Mh, nPv, nMv, nPh, nMh, inc);
                if (t == 0) cnt[q].add(inc, i - K - 1);     // lower diagonal: D[i][i-K] = D[i-1][i-1-K] + inc
                HP[q][t] = nPh; HM[q][t] = nMh;
                Pv[q] = nPv; Mv[q] = nMv;
            }
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < NWQ; ++q) {
        cnt[q].finish(K, o[q].v0);                          // anchor D[K][0] = K
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 1; t < NT; ++t) { o[q].rp[t - 1] = HP[q][t]; o[q].rm[t - 1] = HM[q][t]; }
    }
}

template <int K, int NWQ, int MF = 0>
SMX_HD void bitsliced_small_rows(const u32 *tab, int m, u64 F, SmallOut<K> (&o)[NWQ]) {
    constexpr int S = NWQ == 1 ? 1 : NWQ == 2 ? 2 : 4;
    const u32 *col[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 16; ++j) col[j] = tab + (u32)((F >> (4 * j)) & 15u) * S;
    bitsliced_small_rows_cols<K, NWQ, MF, 16 * S>(col, m, o);
}

// BloomPrefilter.match for a barcode that holds IUPAC codes, with an exact set (bloom_filter.py:70-101, 176-186): the
// filter's keys are barcode + variant[:m-k] over all variants within k edits of the barcode STRING whose substituted
// / inserted characters come from "ACGT"; unedited positions keep the barcode's own character.  So the key
// barcode + flank[:m-k] is present iff the flank prefix is within k such edits of some prefix of the barcode, a
// flank symbol matching a barcode position only when it is the same character (a read's N equals the barcode's N,
// its A does not) and a flank symbol that is not A/C/G/T never being created by an edit.  For A/C/G/T-only barcodes
// this reduces to "the prefix is A/C/G/T" (no false negatives) and is tested on the whole word instead.
// bc: the barcode_rc symbols (4-bit codes); sym(x): symbol x of the flank the prefilter looks at.
template <typename Sym>
SMX_HD bool bloom_literal_yes(const unsigned char *bc, int m, int K, Sym sym) {
    const int need = m - K;
    if (need <= 0) return true;
    int prev[SMX_MAX_PATTERN + 1], cur[SMX_MAX_PATTERN + 1];
    const int kInf = 1 << 20;
    for (int j = 0; j <= m; ++j) prev[j] = j;
    for (int i = 0; i < need; ++i) {
        const int ch = sym(i);
        const bool creatable = ch <= 3;
        cur[0] = creatable ? prev[0] + 1 : kInf;
        for (int j = 1; j <= m; ++j) {
            int best = (ch == (int)bc[j - 1] && ch != kSymOther) ? prev[j - 1] : (creatable ? prev[j - 1] + 1 : kInf);
            if (creatable && prev[j] + 1 < best) best = prev[j] + 1;
            if (cur[j - 1] + 1 < best) best = cur[j - 1] + 1;
            cur[j] = best < kInf ? best : kInf;
        }
        for (int j = 0; j <= m; ++j) prev[j] = cur[j];
    }
    int best = prev[0];
    for (int j = 1; j <= m; ++j) if (prev[j] < best) best = prev[j];
    return best <= K;
}

// The scalar part of the read-out: exact values of ONE flagged barcode (bit q of a bword): smallest D[m][j] over
// the in-range end columns and the mask of the columns attaining it; v0: NV bit-planes of D[m][m-K].
struct DigestAcc { int bd, count, jmin, jmax, first_col; u32 nhits; };

template <int K, int NV>
SMX_HD int barcode_lane_value(const u32 *v0, const u32 *rp, const u32 *rm, int q, int m, int cols, u64 &mask) {
    constexpr int NT = 2 * K + 1;
    int val = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int bb = 0; bb < NV; ++bb) val |= (int)((v0[bb] >> q) & 1) << bb;
    int best = 1 << 20;
    mask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int tt = 0; tt < NT; ++tt) {
        if (tt > 0) val += (int)((rp[tt - 1] >> q) & 1) - (int)((rm[tt - 1] >> q) & 1);
        const int col = m - K + tt;
        if (col <= cols) {
            if (val < best) { best = val; mask = 0; }
            if (val == best) mask |= 1ull << (col - 1);
        }
    }
    return best;
}

// Hit record + digest update for one barcode hit (list position j) of (gslot, entry); nh = hits of this bword so far.
SMX_HD void barcode_add_hit(const Tables &t, const Batch &b, u64 gslot, u64 entry, int nh, int j, int best, u64 mask,
                            int search_start, DigestAcc &acc) {
    if (nh < t.hit_cap) {
        smx_barcode_hit h;
        h.barcode = (uint16_t)j; h.distance = (int16_t)best; h.end_mask = mask; h.search_start = search_start;
        b.bh_list[(gslot * t.hit_cap + nh) * b.e_cap + entry] = h;
    }
    ++acc.nhits;
    if (best < acc.bd) { acc.bd = best; acc.count = 0; acc.jmin = 1 << 20; acc.jmax = -1; }
    if (best == acc.bd) {
        ++acc.count;
        if (j < acc.jmin) { acc.jmin = j; acc.first_col = lowest_bit64(mask); }
        if (j > acc.jmax) acc.jmax = j;
    }
}

template <int K, int NV, typename LaneOk>
SMX_HD void barcode_emit_hits(const Tables &t, const Batch &b, u32 flag, const u32 *v0, const u32 *rp, const u32 *rm,
                              u32 g, u64 gslot, u64 entry, int m, int cols, int search_start, DigestAcc &acc, LaneOk lane_ok) {
    int nh = 0;
    while (flag) {
        int q = lowest_bit32(flag);
        flag &= flag - 1;
        u64 mask;
        const int best = barcode_lane_value<K, NV>(v0, rp, rm, q, m, cols, mask);
        if (best > K || !lane_ok(g, q)) continue;
        barcode_add_hit(t, b, gslot, entry, nh, (int)t.bw_list[(u64)g * 32 + q], best, mask, search_start, acc);
        ++nh;
    }
    b.bh_count[gslot * b.e_cap + entry] = (unsigned char)nh;
    if (nh > t.hit_cap) counter_add(&b.counters[7], 1);      // the library re-runs with a larger cap
}

// Flag planes of the small form: "some in-range end column has D[m][j] <= K" as a threshold test on the biased value
// u = D + (15 - K): u fits five planes (D[m][m-K] <= m <= 16 - K, at most 2K columns further right) and D <= K
// exactly when bit 4 of u is clear.  Walking one column to the right adds the +-1 horizontal delta in ONE ripple:
// the +1 enters as the carry-in, the -1 as an all-ones addend (2 LOP3 per plane).
template <int K>
SMX_HD u32 barcode_flags_small(const SmallOut<K> &o, int m, int cols) {
    constexpr int NT = 2 * K + 1;
    u32 u[5];
    {
        u32 c = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < 5; ++l) {
            const u32 kl = ((15 - K) >> l) & 1 ? ~0u : 0u;
            u[l] = o.v0[l] ^ kl ^ c;
            c = maj3(o.v0[l], kl, c);
        }
    }
    u32 flag = (m - K <= cols) ? ~u[4] : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int tt = 1; tt < NT; ++tt) {
        u32 c = o.rp[tt - 1];
        const u32 neg = o.rm[tt - 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < 5; ++l) {
            const u32 a = u[l];
            u[l] = a ^ neg ^ c;
            c = maj3(a, neg, c);
        }
        if (m - K + tt <= cols) flag |= ~u[4];
    }
    return flag;
}

// Read-out of the general form (BitSliced<K>::run: long barcodes): counter planes with a saturation flag.
template <int K, typename LaneOk>
SMX_HD void barcode_readout(const Tables &t, const Batch &b, const typename BitSliced<K>::Out &o, u32 g, u64 gslot, u64 entry,
                            int m, int cols, int search_start, DigestAcc &acc, LaneOk lane_ok) {
    typedef BitSliced<K> BS;
    u32 v[BS::NB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < BS::NB; ++q) v[q] = o.cnt[q];
    u32 ov = o.over;
    u32 flag = (m - K <= cols) ? (planes_le<K, BS::NB>(v) & ~ov) : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int tt = 1; tt < BS::NT; ++tt) {
        u32 x = o.rp[tt - 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < BS::NB; ++q) { u32 c = v[q] & x; v[q] ^= x; x = c; }
        ov |= x;
        x = o.rm[tt - 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < BS::NB; ++q) { u32 br = ~v[q] & x; v[q] ^= x; x = br; }
        if (m - K + tt <= cols) flag |= planes_le<K, BS::NB>(v) & ~ov;
    }
    flag &= t.bw_valid[g];
    barcode_emit_hits<K, BS::NB>(t, b, flag, o.cnt, o.rp, o.rm, g, gslot, entry, m, cols, search_start, acc, lane_ok);
}

// One work entry (matched slot of a read, one equal-best primer end at staged position p) against the NWQ bwords
// of stage-2 task `task` (all of one barcode length m).  `tab` points at the task's [m][16][S] table (shared
// memory in the CUDA launch).  Returns lanes x columns of the SURVEY.md 8d work formula (x m = cells) in the low
// 20 bits and, from bit 20 up, the number of bwords whose automaton actually ran (0 or NWQ).
template <int K, int NWQ, int MF = 0>
SMX_HD u32 barcode_task_thread(const Tables &t, const Batch &b, u32 read, int p, u64 entry, int strand, int primer,
                               u32 task, const u32 *tab) {
    typedef BitSliced<K> BS;
    const u32 g0 = t.bt_g0[task];
    const u64 gslot0 = (u64)strand * t.n_bwords + g0;
    BarcodeDigest dg;
    dg.search_start = 0; dg.jmin = 0; dg.jmax = 0; dg.count = 0; dg.bd = -1; dg.first_col = 0; dg.nhits = 0;
    BarcodeDigest &dg_out = b.bdig[((u64)strand * t.n_btasks + task) * b.e_cap + entry];
    const int n = (int)b.lengths[read];
    const Geo geo = make_geo(n, t.L);
    const int m = MF ? MF : (int)t.bw_len[g0];
    int nbits = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < NWQ; ++q) nbits += popcount32(t.bw_valid[g0 + q]);
    const bool small = m + K <= 16;            // whole flank fits one 64-bit register

    const Flank f = make_flank(geo.woff + p + geo.delta, n);
    dg.search_start = f.bs;
    const int fl = n - f.a_align;
    const int cols = fl < m + K ? fl : m + K;
    const u32 work = cols > 0 ? (u32)(nbits * cols) : 0u;
    const int base = f.a_align - geo.woff;          // staged index of flank column 1
    u64 F = ~0ull;
    if (small && cols > 0) {
        // 16 symbols starting at staged position `base`, symbols >= cols forced to "other"
        if (sliced_eligible(t, b, read)) {
            // from the 2-bit window words: 16 symbols = 32 bits, spread into nibbles
            const u32 *w2p = b.win2 + (u64)strand * t.nw2 * b.n_pad + read;
            const int w0 = base >> 4, sh = 2 * (base & 15);
            const u32 a0 = w0 < t.nw2 ? w2p[(u64)w0 * b.n_pad] : 0u;
            const u32 a1 = w0 + 1 < t.nw2 ? w2p[(u64)(w0 + 1) * b.n_pad] : 0u;
            const u32 v = sh ? (a0 >> sh) | (a1 << (32 - sh)) : a0;
            F = (u64)spread2to4(v) | ((u64)spread2to4(v >> 16) << 32);
        } else {
            const u32 *win = b.win + (u64)strand * t.wpw * b.n_pad + read;
            int w0 = base >> 3, sh = 4 * (base & 7);
            u32 a0 = w0 < t.wpw ? win[(u64)w0 * b.n_pad] : ~0u;
            u32 a1 = w0 + 1 < t.wpw ? win[(u64)(w0 + 1) * b.n_pad] : ~0u;
            u32 a2 = w0 + 2 < t.wpw ? win[(u64)(w0 + 2) * b.n_pad] : ~0u;
            u64 lo = ((u64)a1 << 32) | a0, hi = a2;
            F = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        }
        if (cols < 16) F |= ~0ull << (4 * cols);
    }
    bool skip = cols < m - K || cols <= 0;           // D[m][j] >= m - j > K for every column
    bool flank_exotic = false;                       // the prefilter's flank prefix holds a symbol other than A/C/G/T
    u32 iupac_any = 0;                               // some barcode of the task holds an IUPAC code
    if (t.prefilter && !skip) {
        // BloomPrefilter.match (bloom_filter.py:176-186) with an exact set: the key barcode_rc + flank[:m-k] can only
        // be present if those m-k symbols exist; for an A/C/G/T-only barcode they must also all be A/C/G/T, and for
        // such flanks the filter has no false negatives.  Barcodes with IUPAC codes get the exact
        // per-barcode test (bloom_literal_yes) on the few lanes that survive the search.
        const int need = m - K;
        if (n - f.a_pref < need) skip = true;
        else if (small && f.a_pref == f.a_align) {
            u64 chk = need >= 16 ? ~0ull : ((1ull << (4 * need)) - 1);
            flank_exotic = (F & chk & 0xCCCCCCCCCCCCCCCCull) != 0;
        } else {
            for (int x = 0; x < need; ++x)
                if (staged_sym(t, b, read, strand, f.a_pref - geo.woff + x) > 3) { flank_exotic = true; break; }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < NWQ; ++q) iupac_any |= t.bw_iupac[g0 + q];
        if (flank_exotic && !iupac_any) skip = true;
    }
    // a flagged barcode (bit `bit` of bword g) passes the prefilter emulation
    auto lane_ok = [&](u32 g, int bit) -> bool {
        if (!t.prefilter) return true;
        if (!((t.bw_iupac[g] >> bit) & 1u)) return !flank_exotic;
        const u32 e = t.pb_off[primer] + t.bw_list[(u64)g * 32 + bit];
        return bloom_literal_yes(t.b_codes + t.b_code_off[e], m, K,
                                 [&](int x) { return staged_sym(t, b, read, strand, f.a_pref - geo.woff + x); });
    };
    if (skip) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < NWQ; ++q) b.bh_count[(gslot0 + q) * b.e_cap + entry] = 0;
        dg_out = dg;
        return work;
    }
    DigestAcc acc;
    acc.bd = 1 << 20; acc.count = 0; acc.jmin = 1 << 20; acc.jmax = -1; acc.first_col = 0; acc.nhits = 0;
    if (small) {
        SmallOut<K> o[NWQ];
        bitsliced_small_rows<K, NWQ, MF>(tab, m, F, o);
        u32 flag[NWQ];
        int nh[NWQ];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < NWQ; ++q) { flag[q] = barcode_flags_small<K>(o[q], m, cols) & t.bw_valid[g0 + q]; nh[q] = 0; }
        // ONE scalar loop over the flagged barcodes of all the task's words (ascending word, then bit = ascending list
        // position): a work entry usually has a single hit, in whichever word its barcode lives, so a warp runs about
        // one iteration instead of one per word (the per-word loops were 20 % of the kernel, profiles/r2_d_*).
        for (;;) {
            int w = -1;
            u32 fw = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int q = NWQ - 1; q >= 0; --q) if (flag[q]) { w = q; fw = flag[q]; }
            if (w < 0) break;
            const int bit = lowest_bit32(fw);
            u32 v0[5], rp[BS::NT > 1 ? BS::NT - 1 : 1], rm[BS::NT > 1 ? BS::NT - 1 : 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int q = 0; q < NWQ; ++q) {                 // the chosen word's planes (register selects)
                if (q == w || q == 0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int l = 0; l < 5; ++l) v0[l] = o[q].v0[l];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int l = 0; l < BS::NT - 1; ++l) { rp[l] = o[q].rp[l]; rm[l] = o[q].rm[l]; }
                }
                if (q == w) flag[q] &= flag[q] - 1;
            }
            u64 mask;
            const int best = barcode_lane_value<K, 5>(v0, rp, rm, bit, m, cols, mask);
            if (best > K || ((flank_exotic || iupac_any) && !lane_ok(g0 + w, bit))) continue;
            int nhw = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int q = 0; q < NWQ; ++q) if (q == w) nhw = nh[q]++;
            barcode_add_hit(t, b, gslot0 + w, entry, nhw, (int)t.bw_list[(u64)(g0 + w) * 32 + bit], best, mask, f.bs, acc);
        }
        bool over_cap = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < NWQ; ++q) {
            b.bh_count[(gslot0 + q) * b.e_cap + entry] = (unsigned char)nh[q];
            over_cap |= nh[q] > t.hit_cap;
        }
        if (over_cap) counter_add(&b.counters[7], 1);        // the library re-runs with a larger cap
    } else {
        // long barcodes (m + K > 16): the general band walk (tasks of such lengths hold one word)
        typename BS::Out o;
        unsigned char fsym[SMX_MAX_PATTERN];                    // m + K <= SMX_MAX_PATTERN flank symbols
        for (int j = 1; j <= m + K; ++j)
            fsym[j - 1] = (unsigned char)(j <= cols ? staged_sym(t, b, read, strand, base + j - 1) : kSymOther);
        BS::run(tab, m, fsym, o);
        barcode_readout<K>(t, b, o, g0, gslot0, entry, m, cols, f.bs, acc, lane_ok);
    }
    if (acc.nhits) {
        dg.nhits = acc.nhits; dg.bd = (signed char)acc.bd; dg.count = (unsigned short)(acc.count > 65535 ? 65535 : acc.count);
        dg.jmin = (unsigned short)acc.jmin; dg.jmax = (unsigned short)acc.jmax; dg.first_col = (unsigned char)acc.first_col;
    }
    dg_out = dg;
    return work | ((u32)NWQ << 20);
}

// Narrow bwords (at most 8 barcodes: config 2's forward primer): FOUR work entries share one machine word, entry q in
// byte lane q.  The Eq word of a cell is one load from a table indexed by the four entries' symbols of that column
// (five symbols each: A/C/G/T or "beyond the flank" -> 625 combinations per row), so the automaton costs the same
// 5 LOP3 + 1 LDS per cell as a full word while serving four entries: the one-entry-per-thread form spent a third of
// stage 2 (66 of 199 us, profiles/r2_h_ncu_full.md) on a word with a quarter of its lanes in use.
// Entries the quad form cannot take (reads on the 4-bit side stream, the Python-slice corner where the prefilter's
// flank differs from the aligned one) run the one-entry routine from the same thread.

template <int K, int MF>
SMX_HD u32 barcode_quad_thread(const Tables &t, const Batch &b, u32 slot, u32 e0, u32 n_entries, int strand, int primer,
                               u32 task, const u32 *tab4, const u32 *tab1) {
    typedef BitSliced<K> BS;
    const u32 g0 = t.bt_g0[task];
    const u64 gslot0 = (u64)strand * t.n_bwords + g0;
    const int m = MF ? MF : (int)t.bw_len[g0];
    const u32 valid8 = t.bw_valid[g0] & 0xFFu, iupac8 = t.bw_iupac[g0] & 0xFFu;
    const int nbits = popcount32(valid8);
    u32 work = 0, ran = 0;
    u32 F2[4];                      // 16 flank symbols, 2 bits each
    int cols[4], bs[4], apref[4], woff[4];
    u32 rd[4];
    bool act[4];
    u32 lanes = 0;                  // byte lanes of the entries whose automaton runs
    for (int q = 0; q < 4; ++q) {
        act[q] = false; F2[q] = 0; cols[q] = 0; bs[q] = 0; apref[q] = 0; woff[q] = 0; rd[q] = 0;
        const u32 e = e0 + (u32)q;
        if (e >= n_entries) continue;
        const u32 read = b.ent_read[(u64)slot * b.e_cap + e];
        const int p = b.ent_pos[(u64)slot * b.e_cap + e];
        const int n = (int)b.lengths[read];
        const Geo geo = make_geo(n, t.L);
        const Flank f = make_flank(geo.woff + p + geo.delta, n);
        if (!sliced_eligible(t, b, read) || f.a_pref != f.a_align) {
            const u32 w1 = barcode_task_thread<K, 1, MF>(t, b, read, p, e, strand, primer, task, tab1);
            work += w1 & 0xFFFFFu; ran += w1 >> 20;
            continue;
        }
        const int fl = n - f.a_align;
        const int c = fl < m + K ? fl : m + K;
        if (c > 0) work += (u32)(nbits * c);
        BarcodeDigest dg;
        dg.search_start = f.bs; dg.jmin = 0; dg.jmax = 0; dg.count = 0; dg.bd = -1; dg.first_col = 0; dg.nhits = 0;
        b.bdig[((u64)strand * t.n_btasks + task) * b.e_cap + e] = dg;          // overwritten below when hits turn up
        b.bh_count[gslot0 * b.e_cap + e] = 0;
        if (c < m - K || c <= 0 || (t.prefilter && n - f.a_pref < m - K)) continue;     // no barcode can match
        const int base = f.a_align - geo.woff;
        const u32 *w2p = b.win2 + (u64)strand * t.nw2 * b.n_pad + read;
        const int w0 = base >> 4, sh = 2 * (base & 15);
        const u32 a0 = w0 < t.nw2 ? w2p[(u64)w0 * b.n_pad] : 0u;
        const u32 a1 = w0 + 1 < t.nw2 ? w2p[(u64)(w0 + 1) * b.n_pad] : 0u;
        F2[q] = sh ? (a0 >> sh) | (a1 << (32 - sh)) : a0;
        cols[q] = c; bs[q] = f.bs; apref[q] = f.a_pref; woff[q] = geo.woff; rd[q] = read;
        act[q] = true;
        lanes |= valid8 << (8 * q);
        ++ran;
    }
    if (!lanes) return work | (ran << 20);
    // table address of every column: the four entries' symbols (4 = beyond the entry's flank, or no entry)
    const u32 *col[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 16; ++j) {
        u32 idx = 0, mul = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q) {
            const u32 sym = (act[q] && j < cols[q]) ? (F2[q] >> (2 * j)) & 3u : 4u;
            idx += sym * mul;
            mul *= kQuadSyms;
        }
        col[j] = tab4 + idx;
    }
    SmallOut<K> o[1];
    bitsliced_small_rows_cols<K, 1, MF, kQuadRow>(col, m, o);
    // flags over all end columns (an entry's own column limit is applied by the exact scalar read-out)
    u32 flag = barcode_flags_small<K>(o[0], m, 16) & lanes;
    DigestAcc acc[4];
    int nh[4];
    for (int q = 0; q < 4; ++q) { acc[q].bd = 1 << 20; acc[q].count = 0; acc[q].jmin = 1 << 20; acc[q].jmax = -1; acc[q].first_col = 0; acc[q].nhits = 0; nh[q] = 0; }
    while (flag) {
        const int bit = lowest_bit32(flag);
        flag &= flag - 1;
        const int q = bit >> 3, lane = bit & 7;
        u64 mask;
        const int best = barcode_lane_value<K, 5>(o[0].v0, o[0].rp, o[0].rm, bit, m, cols[q], mask);
        if (best > K) continue;
        if (t.prefilter && ((iupac8 >> lane) & 1u)) {           // exact per-barcode Bloom emulation (see bloom_literal_yes)
            const u32 le = t.pb_off[primer] + t.bw_list[(u64)g0 * 32 + lane];
            const u32 read = rd[q];
            const int off = apref[q] - woff[q];
            if (!bloom_literal_yes(t.b_codes + t.b_code_off[le], m, K, [&](int x) { return staged_sym(t, b, read, strand, off + x); }))
                continue;
        }
        barcode_add_hit(t, b, gslot0, (u64)e0 + q, nh[q], (int)t.bw_list[(u64)g0 * 32 + lane], best, mask, bs[q], acc[q]);
        ++nh[q];
    }
    bool over_cap = false;
    for (int q = 0; q < 4; ++q) {
        if (!act[q] || !nh[q]) continue;
        const u64 e = (u64)e0 + q;
        b.bh_count[gslot0 * b.e_cap + e] = (unsigned char)nh[q];
        over_cap |= nh[q] > t.hit_cap;
        BarcodeDigest dg;
        dg.search_start = bs[q]; dg.nhits = acc[q].nhits; dg.bd = (signed char)acc[q].bd;
        dg.count = (unsigned short)(acc[q].count > 65535 ? 65535 : acc[q].count);
        dg.jmin = (unsigned short)acc[q].jmin; dg.jmax = (unsigned short)acc[q].jmax; dg.first_col = (unsigned char)acc[q].first_col;
        b.bdig[((u64)strand * t.n_btasks + task) * b.e_cap + e] = dg;
    }
    if (over_cap) counter_add(&b.counters[7], 1);
    return work | (ran << 20);
}

// Work-entry bookkeeping of stage 1: entries [base, base + nloc) of `slot` for one matched read.
// `ev`: the end mask in registers (primer_finish_thread), or nullptr to read it from memory.
SMX_HD void write_entries(const Tables &t, const Batch &b, u32 slot, u32 read, u32 base, const u32 *ev = nullptr) {
    b.ent_base[(u64)slot * b.n_pad + read] = base;
    const u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + read;
    u32 e = base;
    for (int mwi = 0; mwi < t.mw; ++mwi) {
        u32 word = ev ? ev[mwi] : emask[(u64)mwi * b.n_pad];
        while (word) {
            int p = mwi * 32 + lowest_bit32(word);
            word &= word - 1;
            if (e < b.e_cap) {
                b.ent_read[(u64)slot * b.e_cap + e] = read;
                b.ent_pos[(u64)slot * b.e_cap + e] = (unsigned short)p;
            }
            ++e;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Long primers: carry-lookahead across the SW words of every segment of a warp at once.  Bit l of
// G / P says word (lane) l generates a carry / would pass an incoming carry on; the result has bit
// l set iff a carry enters word l.  One integer addition ripples all segments: the generate bits
// are injected one position up, the propagate bits let the machine carry run through, and word 0
// of every segment (never a carry target) is masked out so nothing crosses a segment boundary.
// Segments need not be a power of two wide: a warp holds 32 / SW whole segments (lanes beyond the last whole segment
// idle) -- a 150-nt primer is five words, six reads to a warp instead of the four of an eight-lane segment.
template <int SW> struct LongSeg {
    static_assert(SW >= 2 && SW <= 32, "segment width");
    static constexpr int kReads = 32 / SW;                 // whole segments (reads) per warp
    static constexpr u32 start_mask() {                    // word 0 of every segment, and every idle lane
        u32 m = 0;
        for (int l = 0; l < 32; ++l)
            if (l >= kReads * SW || l % SW == 0) m |= 1u << l;
        return m;
    }
    static constexpr u32 kStart = start_mask();
};

template <int SW> SMX_HD u32 long_carry_in(u32 G, u32 P) {
    constexpr u32 inner = ~LongSeg<SW>::kStart;
    const u32 V = (G << 1) & inner, Pm = P & inner;
    return ((Pm + V) ^ Pm) & inner;
}

}  // namespace smx
