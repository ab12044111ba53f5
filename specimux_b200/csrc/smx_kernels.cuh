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

// One thread stages 16 symbols of one strand: the 2-bit word win2[w2] (input of the sliced primer
// search) and the two 4-bit words win[2*w2], win[2*w2+1] (barcode stage, start recovery, classic
// primer search).  Positions past the staged length hold kSymOther in `win`.
// `src2` / `src2_origin`: where the read's 2-bit words are read from -- src2[word_off[read] - src2_origin + i];
// the packed stream itself (b.packed2, b.word_base) or a block's shared-memory copy of its reads' words.
SMX_HD void stage_window_pair(const Tables &t, const Batch &b, u32 read, int strand, int w2, const u32 *src2, u64 src2_origin) {
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    u32 *wlo = b.win + ((u64)strand * t.wpw + 2 * w2) * b.n_pad + read;
    const bool has_hi = 2 * w2 + 1 < t.wpw;
    int valid = g.wl - 16 * w2;
    valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
    if (!read_is_flagged(b, read)) {
        // fast path: the staged symbols are one contiguous run of the 2-bit stream (the tail of the
        // read for strand 0, its head read backwards and complemented for strand 1)
        u32 v = 0;
        if (valid) {
            const int x0 = g.woff + 16 * w2;                  // strand coordinate of symbol 0
            const int first = strand ? (n - 1 - x0) - 15 : stored_pos(b, x0, n);   // stored index of the lowest base needed
            const int lo = first < 0 ? 0 : first;
            const u32 *src = src2 + (b.word_off[read] - src2_origin) + (u64)(lo >> 4);
            const u64 pair = (u64)src[0] | ((u64)src[1] << 32);
            v = (u32)(pair >> (2 * (lo & 15)));
            if (first < 0) v <<= 2 * (-first);                // bases before the read start: masked below
            if (strand) v = revcomp16(v);
        }
        b.win2[((u64)strand * t.nw2 + w2) * b.n_pad + read] = v;
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
    b.win2[((u64)strand * t.nw2 + w2) * b.n_pad + read] = 0;      // flagged reads take the classic search
    for (int h = 0; h < (has_hi ? 2 : 1); ++h) {
        u32 out = 0;
        for (int i = 0; i < 8; ++i) {
            int p = w2 * 16 + h * 8 + i;
            int c = kSymOther;
            if (p < g.wl) c = sym_at(b, read, strand, g.woff + p, n);
            out |= (u32)c << (4 * i);
        }
        wlo[(u64)h * b.n_pad] = out;
    }
}

SMX_HD int staged_sym(const Tables &t, const Batch &b, u32 read, int strand, int p) {
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
                    const u32 Pv = VP[i], Mv = VM[i];
                    const u32 Xv = Eq | Mv, Xh = Eq | Mh;
                    VP[i] = Mh | ~(Xv | Ph);
                    VM[i] = Ph & Xv;
                    const u32 nPh = Mv | ~(Xh | Pv);
                    Mh = Pv & Xh;
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

SMX_HD bool sliced_eligible(const Tables &t, const Batch &b, u32 read) {
    return t.sliced && !read_is_flagged(b, read) && (int)b.lengths[read] >= t.L;
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

// Start recovery for the first equal-best end of a matched slot (edlib: reverse SHW pass with
// k = best, the LAST equal-best reverse end wins = longest alignment).  One thread per work entry;
// only a read's first entry does the work.
template <typename W>
SMX_HD void primer_start_thread(const Tables &t, const Batch &b, u32 slot, u32 entry, const u64 *peq_rev) {
    const u32 read = b.ent_read[(u64)slot * b.e_cap + entry];
    const u64 hit_idx = (u64)slot * b.n_pad + read;
    if (b.ent_base[hit_idx] != entry) return;              // not the first location of its read
    const int strand = (int)slot / t.n_primers, primer = (int)slot % t.n_primers;
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    smx_primer_hit &h = b.phit[hit_idx];
    const int first = (int)b.ent_pos[(u64)slot * b.e_cap + entry];
    auto load = [&](int p) { return staged_sym(t, b, read, strand, p); };
    int back = hw_start_back<W>(peq_rev, t.p_len[primer], h.distance, first, g.start, load);
    h.first_start = h.first_end - back;
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
    static constexpr int NB = K <= 1 ? 3 : (K <= 4 ? 4 : 5);   // counter bits: values up to 3K+1 matter

    struct Out { u32 rp[NT > 1 ? NT - 1 : 1], rm[NT > 1 ? NT - 1 : 1], cnt[NB], over; };

    // One cell of the automaton for 32 barcodes.
    static SMX_HD void cell(u32 Eq, u32 Pv, u32 Mv, u32 Ph, u32 Mh, u32 &oPv, u32 &oMv, u32 &oPh, u32 &oMh, u32 &inc) {
        u32 Xv = Eq | Mv, Xh = Eq | Mh;
        oPv = Mh | ~(Xv | Ph);
        oMv = Ph & Xv;
        oPh = Mv | ~(Xh | Pv);
        oMh = Pv & Xh;
        inc = ~(Xv | Mh);                                      // D[i][j] - D[i-1][j-1]
    }

    // beq_rows: table rows of this bword ([i][16]); rowwin(i): nibble t = 4-bit symbol of flank
    // column i-K+t (1-based columns), kSymOther beyond the flank.
    template <typename RowWin>
    static SMX_HD void run(const u32 *beq_rows, int m, RowWin rowwin, Out &o) {
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
            const u64 W = rowwin(i);
            u32 Pv = ~0u, Mv = 0;                              // column 0 / below the band: +1
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < NT; ++t) {
                int j = i - K + t;
                if (j < 1) continue;                           // uniform over the warp
                u32 Ph = t < NT - 1 ? HP[t + 1] : ~0u;          // above the band: +1
                u32 Mh = t < NT - 1 ? HM[t + 1] : 0u;
                u32 Eq = row[(u32)(W >> (4 * t)) & 15u];
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

    // Small form (m + K <= 16): the whole flank sits in one 64-bit register F (nibble j-1 = symbol of
    // column j, kSymOther beyond the flank).  One table address is formed per COLUMN (beq row 0 of
    // that column's symbol); the rows are fully unrolled, so every cell's Eq is a single load at a
    // compile-time offset from its column's address -- no per-cell shift / mask / scale arithmetic.
    static constexpr int kSmallRows = 16 - K > 0 ? 16 - K : 0;
    static SMX_HD void run_small(const u32 *beq_rows, int m, u64 F, Out &o) {
        const u32 *col[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 16; ++j) col[j] = beq_rows + (u32)((F >> (4 * j)) & 15u);
        u32 HP[NT], HM[NT];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 0; t < NT; ++t) { HP[t] = ~0u; HM[t] = 0; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < NB; ++b) o.cnt[b] = (K >> b) & 1 ? ~0u : 0u;
        o.over = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 1; i <= kSmallRows; ++i) {
            if (i <= m) {                                       // uniform over the block
                u32 Pv = ~0u, Mv = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int t = 0; t < NT; ++t) {
                    const int j = i - K + t;                    // compile-time
                    if (j < 1 || j > 16) continue;
                    u32 Ph = t < NT - 1 ? HP[t + 1] : ~0u;
                    u32 Mh = t < NT - 1 ? HM[t + 1] : 0u;
                    u32 Eq = col[j - 1][(i - 1) * 16];
                    u32 nPv, nMv, nPh, nMh, inc;
                    cell(Eq, Pv, Mv, Ph, Mh, nPv, nMv, nPh, nMh, inc);
                    if (t == 0) {
                        u32 x = inc;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                        for (int b = 0; b < NB; ++b) { u32 c = o.cnt[b] & x; o.cnt[b] ^= x; x = c; }
                        o.over |= x;
                    }
                    HP[t] = nPh; HM[t] = nMh;
                    Pv = nPv; Mv = nMv;
                }
            }
        }
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

// One work entry (matched slot of a read, one equal-best primer end at staged position p) and one
// bword.  `beq_rows` points at the bword's [m][16] table (shared memory in the CUDA launch).
// Accumulates the SURVEY.md 8d work formula into cells / wcols.
template <int K>
SMX_HD void barcode_bitsliced_thread(const Tables &t, const Batch &b, u32 read, int p, u64 entry, int strand,
                                     int primer, u32 g, const u32 *beq_rows,
                                     unsigned long long &cells, unsigned long long &wcols) {
    typedef BitSliced<K> BS;
    const u64 gslot = (u64)strand * t.n_bwords + g;
    unsigned char &out_count = b.bh_count[gslot * b.e_cap + entry];
    out_count = 0;
    const int n = (int)b.lengths[read];
    const Geo geo = make_geo(n, t.L);
    const u32 *win = b.win + (u64)strand * t.wpw * b.n_pad + read;
    const int m = t.bw_len[g];
    const u32 valid = t.bw_valid[g];
    const int nbits = popcount32(valid);
    const bool small = m + K <= 16;            // whole flank fits one 64-bit register

    const Flank f = make_flank(geo.woff + p + geo.delta, n);
    const int fl = n - f.a_align;
    const int cols = fl < m + K ? fl : m + K;
    if (cols > 0) {
        cells += (unsigned long long)nbits * m * cols;
        wcols += (unsigned long long)nbits * ((m + 31) >> 5) * cols;
    }
    const int base = f.a_align - geo.woff;          // staged index of flank column 1
    u64 F = ~0ull;
    if (small && cols > 0) {
        // 16 symbols starting at staged position `base`, symbols >= cols forced to "other"
        int w0 = base >> 3, sh = 4 * (base & 7);
        u32 a0 = w0 < t.wpw ? win[(u64)w0 * b.n_pad] : ~0u;
        u32 a1 = w0 + 1 < t.wpw ? win[(u64)(w0 + 1) * b.n_pad] : ~0u;
        u32 a2 = w0 + 2 < t.wpw ? win[(u64)(w0 + 2) * b.n_pad] : ~0u;
        u64 lo = ((u64)a1 << 32) | a0, hi = a2;
        F = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        if (cols < 16) F |= ~0ull << (4 * cols);
    }
    if (t.prefilter) {
        // BloomPrefilter.match (bloom_filter.py:176-186) with an exact set: the key
        // barcode_rc + flank[:m-k] can only be present if those m-k symbols exist and are
        // all A/C/G/T; for such flanks the filter has no false negatives.
        const int need = m - K;
        if (n - f.a_pref < need) return;
        bool acgt = true;
        if (small && f.a_pref == f.a_align) {
            u64 chk = need >= 16 ? ~0ull : ((1ull << (4 * need)) - 1);
            acgt = (F & chk & 0xCCCCCCCCCCCCCCCCull) == 0;
        } else {
            for (int x = 0; x < need; ++x)
                if (staged_sym(t, b, read, strand, f.a_pref - geo.woff + x) > 3) { acgt = false; break; }
        }
        if (!acgt) return;
    }
    if (cols < m - K || cols <= 0) return;           // D[m][j] >= m - j > K for every column
    typename BS::Out o;
    if (small) {
        BS::run_small(beq_rows, m, F, o);
    } else {
        auto rowwin = [&](int i) -> u64 {
            u64 W = 0;
            for (int tt = 0; tt < BS::NT; ++tt) {
                int j = i - K + tt;
                u64 c = (j >= 1 && j <= cols) ? (u64)staged_sym(t, b, read, strand, base + j - 1) : 15ull;
                W |= c << (4 * tt);
            }
            return W;
        };
        BS::run(beq_rows, m, rowwin, o);
    }
    // bit-sliced test "some in-range column has D[m][j] <= K"
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
    flag &= valid;
    // exact scalar read-out of the few flagged barcodes (ascending bit = ascending list position)
    int nh = 0;
    while (flag) {
        int q = lowest_bit32(flag);
        flag &= flag - 1;
        int val = 0;
        for (int bb = 0; bb < BS::NB; ++bb) val |= (int)((o.cnt[bb] >> q) & 1) << bb;
        int best = 1 << 20;
        u64 mask = 0;
        for (int tt = 0; tt < BS::NT; ++tt) {
            if (tt > 0) val += (int)((o.rp[tt - 1] >> q) & 1) - (int)((o.rm[tt - 1] >> q) & 1);
            int col = m - K + tt;
            if (col > cols) break;
            if (val < best) { best = val; mask = 0; }
            if (val == best) mask |= 1ull << (col - 1);
        }
        if (best > K) continue;
        if (nh < t.hit_cap) {
            smx_barcode_hit h;
            h.barcode = t.bw_list[(u64)g * 32 + q]; h.distance = (int16_t)best; h.end_mask = mask; h.search_start = f.bs;
            b.bh_list[(gslot * t.hit_cap + nh) * b.e_cap + entry] = h;
        }
        ++nh;
    }
    out_count = (unsigned char)nh;
    if (nh > t.hit_cap) counter_add(&b.counters[7], 1);      // the library re-runs with a larger cap
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
template <int SW> struct LongSeg {
    static_assert(SW == 4 || SW == 8 || SW == 16 || SW == 32, "segment width");
    static constexpr u32 kStart = SW == 32 ? 0x00000001u : SW == 16 ? 0x00010001u : SW == 8 ? 0x01010101u : 0x11111111u;
};

template <int SW> SMX_HD u32 long_carry_in(u32 G, u32 P) {
    constexpr u32 inner = ~LongSeg<SW>::kStart;
    const u32 V = (G << 1) & inner, Pm = P & inner;
    return ((Pm + V) ^ Pm) & inner;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// __global__ wrappers.  Tables travel as __grid_constant__ parameters, Peq masks are staged once per block
// in shared memory.

// Tables and Batch travel as __grid_constant__ kernel parameters (constant bank, per launch), so
// several contexts / pipeline lanes can be in flight on one device without sharing a symbol.
#define SMX_KARGS const __grid_constant__ Tables c_tables, const __grid_constant__ Batch b

// One block stages the windows of 128 consecutive reads.  Their 2-bit words are one contiguous range
// of the packed stream (reads are packed back to back), so the block first copies that range into
// shared memory with fully coalesced loads and every thread then cuts its read's 2 * nw2 windows
// out of the copy; the stores are coalesced across the block's reads as before.  The first form --
// one thread per (read, window word), each loading its two words straight from the stream -- made a
// warp touch 32 different reads' lines per load and was the top kernel of the long-amplicon config
// (457 us of 1,279; profiles/r1_v20_bench_long.json).  The tile is dynamic shared memory sized by the
// host for the batch's clip length (128 reads x the most words a clipped read can have); blocks whose
// reads span more than that (unclipped long reads) read the stream directly.
constexpr int kStageBlock = 128;                // reads per block; blockDim = (kStageBlock, 2 strands)

// tile_words: capacity of the dynamic shared-memory tile in words (0 = always read the stream directly)
__global__ void __launch_bounds__(2 * kStageBlock) k_stage_windows(SMX_KARGS, u32 tile_words) {
    extern __shared__ u32 s_src[];
    const Tables &t = c_tables;
    const u32 r0 = blockIdx.x * kStageBlock;
    const u32 r1 = r0 + kStageBlock < b.n_reads ? r0 + kStageBlock : b.n_reads;
    const u64 w0 = b.word_off[r0];
    const u64 w1 = b.word_off[r1 - 1] + (u64)((stored_len(b, (int)b.lengths[r1 - 1]) + 15) >> 4);
    const bool tiled = w1 - w0 + 2 <= tile_words;                             // window extraction reads one word past the read
    if (tiled) {
        const u32 span = (u32)(w1 - w0) + 2;
        const u32 *g = b.packed2 + (w0 - b.word_base);
        for (u32 i = threadIdx.y * kStageBlock + threadIdx.x; i < span; i += 2 * kStageBlock) s_src[i] = g[i];
    }
    __syncthreads();
    const u32 read = r0 + threadIdx.x;
    if (read >= b.n_reads) return;
    const u32 *src2 = tiled ? s_src : b.packed2;
    const u64 origin = tiled ? w0 : b.word_base;
    const int strand = (int)threadIdx.y;
    for (int w2 = 0; w2 < t.nw2; ++w2) stage_window_pair(t, b, read, strand, w2, src2, origin);
}

// Sliced primer search: one thread per (group of 32 reads, strand) for one primer of length M.
constexpr int kSlicedBlock = 64;        // small blocks: equal-length tasks, let the block scheduler balance the SMs
template <int M>
__global__ void __launch_bounds__(kSlicedBlock) k_primer_sliced(SMX_KARGS, int primer, const __grid_constant__ RowOffsets ro, int degenerate) {
    __shared__ u32 s_planes[(kSlicedCodes + 48) * kSlicedBlock];
    const u32 group = blockIdx.x * kSlicedBlock + threadIdx.x;
    if (group >= b.n_pad / 32) return;
    primer_sliced_thread<M, kSlicedBlock>(c_tables, b, group, (int)blockIdx.y, primer, ro, degenerate != 0,
                                          s_planes + threadIdx.x, s_planes + kSlicedCodes * kSlicedBlock + threadIdx.x);
}

constexpr int kFinishBlock = 256;

// Work counters.  Every thread of a block contributes a small 32-bit count `v` (DP columns, or
// barcode lanes x columns); the block adds v_total * mul0 and v_total * mul1 to two 64-bit device
// counters.  One REDUX per warp, one shared atomic per warp, two global atomics per block (the
// first form -- a shuffle tree per counter and two barriers each -- was 9 % of the barcode
// kernel's instructions and 14 % of its stall samples: profiles/r1_v14_ncu_full.md).
__device__ __forceinline__ void block_work_add(u32 v, unsigned long long mul0, unsigned long long *dst0,
                                               unsigned long long mul1, unsigned long long *dst1, u32 *s_acc /*1, zeroed*/) {
    const u32 wsum = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(s_acc, wsum);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long tot = *s_acc;
        if (tot) { atomicAdd(dst0, tot * mul0); atomicAdd(dst1, tot * mul1); }
    }
}

template <typename W>
__global__ void __launch_bounds__(kFinishBlock) k_primer_search(SMX_KARGS) {
    // grid: x over reads, y = strand * n_primers + primer
    __shared__ u64 s_peq[3][16];
    __shared__ u32 s_wtot[kFinishBlock / 32 + 1];
    __shared__ u32 s_acc;
    const int primer = blockIdx.y % c_tables.n_primers, strand = blockIdx.y / c_tables.n_primers;
    if (c_tables.p_sw[primer]) return;                  // long primer: k_primer_long owns this slot
    if (threadIdx.x == 64) s_acc = 0;
    if (threadIdx.x < 48) {
        const u64 *src = threadIdx.x < 16 ? c_tables.peq_rc : threadIdx.x < 32 ? c_tables.peq_rcrev : c_tables.peq_fw;
        s_peq[threadIdx.x >> 4][threadIdx.x & 15] = src[primer * 16 + (threadIdx.x & 15)];
    }
    __syncthreads();
    u32 read = blockIdx.x * blockDim.x + threadIdx.x;
    u32 cells = 0;
    int nloc = 0;
    u32 ev[kFinishMaskWords];
    bool have_ev = false;
    if (read < b.n_reads) {
        nloc = primer_finish_thread<W>(c_tables, b, read, strand, primer, s_peq[0], s_peq[2], ev, &have_ev);
        int n = (int)b.lengths[read];
        cells = (u32)(n < c_tables.L ? n : c_tables.L);                      // HW columns of this search
    }
    // work entries, one per equal-best end location: block-aggregated allocation (one atomic per
    // block), a read's entries stay consecutive (the order of reads inside the list is irrelevant)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int incl = nloc;
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wtot[warp] = (u32)incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 run = 0;
            for (int w = 0; w < kFinishBlock / 32; ++w) { u32 v = s_wtot[w]; s_wtot[w] = run; run += v; }
            u32 base = run ? atomicAdd(&b.slot_count[blockIdx.y], run) : 0u;
            s_wtot[kFinishBlock / 32] = base;
        }
        __syncthreads();
        if (nloc) write_entries(c_tables, b, blockIdx.y, read, s_wtot[kFinishBlock / 32] + s_wtot[warp] + (u32)(incl - nloc),
                                have_ev ? ev : nullptr);
    }
    const int m = c_tables.p_len[primer];
    block_work_add(cells, (unsigned long long)m, &b.counters[0], (unsigned long long)((m + 31) >> 5), &b.counters[2], &s_acc);
}

// ---------------------------------------------------------------------------------------------
// Long primers (65 .. 1024 nt): warp-cooperative multi-word Myers/Hyyro.  The pattern occupies the
// top m bits of a 32*SW-bit vector spread over SW consecutive lanes (lane `sub` holds word `sub`;
// row m is the sign bit of the top lane), so a warp works on 32/SW reads at once.  Per column the
// only cross-lane traffic is the carry of (Eq & Pv) + Pv -- resolved for all segments at once from
// two ballots (generate / propagate masks, carry-lookahead by one integer addition) -- and the
// one-bit shifts of Ph / Mh (__shfl_up).  Same recurrences as myers_step<>, i.e. edlib's
// calculateBlock over several blocks (alignment.py:42).  One kernel does the forward HW pass, the
// hit bookkeeping, the reverse pass that recovers the start of the first location, the work
// entries and, for irregular reads, the explicit orientation test.

// One DP column for every segment of the warp.  Returns the score delta of the last row (only
// meaningful in a segment's top lane).  Must be called by all 32 lanes.
template <int SW, bool kShiftInOne>
__device__ __forceinline__ int long_step(u32 Eq, u32 &Pv, u32 &Mv, int sub, int lane) {
    const u32 a = Eq & Pv;
    u32 sum = a + Pv;
    const u32 G = __ballot_sync(0xffffffffu, sum < a);                 // word generates a carry
    const u32 P = __ballot_sync(0xffffffffu, sum == 0xFFFFFFFFu);      // word propagates an incoming carry
    sum += (long_carry_in<SW>(G, P) >> lane) & 1u;
    const u32 Xh = (sum ^ Pv) | Eq;
    const u32 Xv = Eq | Mv;
    u32 Ph = Mv | ~(Xh | Pv);
    u32 Mh = Pv & Xh;
    const int d = (int)(Ph >> 31) - (int)(Mh >> 31);
    u32 Ph_lo = __shfl_up_sync(0xffffffffu, Ph, 1), Mh_lo = __shfl_up_sync(0xffffffffu, Mh, 1);
    if (sub == 0) { Ph_lo = kShiftInOne ? 0x80000000u : 0u; Mh_lo = 0u; }
    Ph = __funnelshift_l(Ph_lo, Ph, 1);
    Mh = __funnelshift_l(Mh_lo, Mh, 1);
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return d;
}

template <int SW>
__global__ void __launch_bounds__(128) k_primer_long(SMX_KARGS, int primer) {
    // grid: x over (read, word) pairs, y = strand
    __shared__ u32 s_peq[3][16 * SW];
    const Tables &t = c_tables;
    {
        const u32 *src = t.peq_long + t.p_long[primer];
        for (int i = threadIdx.x; i < 3 * 16 * SW; i += blockDim.x) s_peq[i / (16 * SW)][i % (16 * SW)] = src[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane % SW;
    const bool top = sub == SW - 1;
    const int strand = (int)blockIdx.y;
    const u32 read = (u32)(((u64)blockIdx.x * blockDim.x + threadIdx.x) / SW);
    const bool valid = read < b.n_reads;
    const int m = t.p_len[primer], k = t.p_k[primer];
    const u32 slot = slot_index(t, strand, primer);
    const int n = valid ? (int)b.lengths[read] : 0;
    const Geo g = make_geo(n, t.L);
    const u64 hit_idx = (u64)slot * b.n_pad + (valid ? read : 0);
    u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + (valid ? read : 0);
    u32 *imask = b.impmask + (u64)slot * t.mw * b.n_pad + (valid ? read : 0);

    // ---- forward HW pass over the staged window [g.start, g.wl)
    const int p_begin = g.start, cols = valid ? g.wl - g.start : 0;
    const u32 *wwin = b.win + (u64)strand * t.wpw * b.n_pad + (valid ? read : 0);     // this read's 4-bit window words
    u32 wcur = 0;
    int wcur_idx = -1;
    if (valid && top) for (int w = 0; w < t.mw; ++w) { emask[(u64)w * b.n_pad] = 0; imask[(u64)w * b.n_pad] = 0; }
    u32 Pv = ~0u, Mv = 0u;
    int score = m, best = m + 1;
    {
        u32 eqw = 0, imw = 0;
        int cur = p_begin >> 5;
        const int maxcols = __reduce_max_sync(0xffffffffu, cols);
        for (int j = 0; j < maxcols; ++j) {
            const bool active = j < cols;
            const int p = p_begin + j;
            int c = kSymOther;
            if (active) {                                   // one window word per 8 columns, not one load per column
                if ((p >> 3) != wcur_idx) { wcur_idx = p >> 3; wcur = wwin[(u64)wcur_idx * b.n_pad]; }
                c = (int)((wcur >> (4 * (p & 7))) & 15u);
            }
            const u32 sPv = Pv, sMv = Mv;
            const int d = long_step<SW, false>(s_peq[0][c * SW + sub], Pv, Mv, sub, lane);
            if (!active) { Pv = sPv; Mv = sMv; }
            else if (top) {
                if ((p >> 5) != cur) { emask[(u64)cur * b.n_pad] = eqw; imask[(u64)cur * b.n_pad] = imw; eqw = imw = 0; cur = p >> 5; }
                score += d;
                if (score < best) { best = score; imw |= 1u << (p & 31); }
                if (score == best) eqw |= 1u << (p & 31);
            }
        }
        if (valid && top && cols > 0) { emask[(u64)cur * b.n_pad] = eqw; imask[(u64)cur * b.n_pad] = imw; }
    }
    // ---- hit bookkeeping (top lane), then the segment learns (nloc, first, best)
    int nloc = 0, first = 0;
    if (valid && top) {
        nloc = primer_tail(t, b, read, strand, primer, best);
        if (nloc) first = b.phit[hit_idx].first_end - g.woff - g.delta;
    }
    const int src_lane = lane - sub + SW - 1;
    nloc = __shfl_sync(0xffffffffu, nloc, src_lane);
    first = __shfl_sync(0xffffffffu, first, src_lane);
    best = __shfl_sync(0xffffffffu, best, src_lane);
    // ---- reverse SHW pass from the first equal-best end: the LAST column with score == best is the
    //      longest alignment (edlib start recovery)
    {
        int rcols = 0;
        if (nloc) { rcols = first - p_begin + 1; if (rcols > m + best) rcols = m + best; }
        const int maxr = __reduce_max_sync(0xffffffffu, rcols);
        // pattern mask: bits at positions >= 32*SW - m of the 32*SW-bit vector
        const int lo = 32 * SW - m - 32 * sub;                // first pattern bit inside this word
        Pv = lo <= 0 ? ~0u : (lo >= 32 ? 0u : ~0u << lo);
        Mv = 0u;
        int rs = m, last = m - 1;
        for (int j = 0; j < maxr; ++j) {
            const bool active = j < rcols;
            int c = kSymOther;
            if (active) {
                const int p = first - j;
                if ((p >> 3) != wcur_idx) { wcur_idx = p >> 3; wcur = wwin[(u64)wcur_idx * b.n_pad]; }
                c = (int)((wcur >> (4 * (p & 7))) & 15u);
            }
            const u32 sPv = Pv, sMv = Mv;
            const int d = long_step<SW, true>(s_peq[1][c * SW + sub], Pv, Mv, sub, lane);
            if (!active) { Pv = sPv; Mv = sMv; }
            else if (top) { rs += d; if (rs == best) last = j; }
        }
        if (valid && top && nloc) b.phit[hit_idx].first_start = b.phit[hit_idx].first_end - last;
    }
    // ---- work entries and counters (top lane; this kernel is the rare path, plain atomics)
    if (valid && top) {
        if (nloc) write_entries(t, b, slot, read, atomicAdd(&b.slot_count[slot], (u32)nloc));
        const unsigned long long hw_cols = (unsigned long long)(n < t.L ? n : t.L);
        atomicAdd(&b.counters[0], hw_cols * (unsigned long long)m);
        atomicAdd(&b.counters[2], hw_cols * (unsigned long long)((m + 31) >> 5));
    }
    // ---- determine_orientation, explicit form (demultiplex.py:602-638), irregular reads only
    {
        const bool need = valid && t.preorient && (!g.regular || read_is_flagged(b, read));
        const int ocols = need ? (n < t.L ? n : t.L) : 0;
        const int maxo = __reduce_max_sync(0xffffffffu, ocols);
        Pv = ~0u; Mv = 0u;
        int sc = m, bst = m + 1;
        for (int x = 0; x < maxo; ++x) {
            const bool active = x < ocols;
            const int c = active ? sym_at(b, read, strand, x, n) : kSymOther;
            const u32 sPv = Pv, sMv = Mv;
            const int d = long_step<SW, false>(s_peq[2][c * SW + sub], Pv, Mv, sub, lane);
            if (!active) { Pv = sPv; Mv = sMv; }
            else if (top) { sc += d; if (sc < bst) bst = sc; }
        }
        if (valid && top) b.orient_hit[hit_idx] = (unsigned char)(need && bst <= k);
    }
}

// Start recovery over the compact work-entry lists (full warps instead of the ~50 % matched lanes).
template <typename W>
__global__ void __launch_bounds__(128) k_primer_start(SMX_KARGS) {
    __shared__ u64 s_rev[16];
    const u32 slot = blockIdx.y;
    if (c_tables.p_sw[slot % c_tables.n_primers]) return;   // long primer: start recovered by k_primer_long
    u32 cnt = b.slot_count[slot];
    if (cnt > b.e_cap) cnt = b.e_cap;
    if (blockIdx.x * blockDim.x >= cnt) return;
    if (threadIdx.x < 16) s_rev[threadIdx.x] = c_tables.peq_rcrev[(slot % c_tables.n_primers) * 16 + threadIdx.x];
    __syncthreads();
    const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < cnt) primer_start_thread<W>(c_tables, b, slot, idx, s_rev);
}

// One thread per (matched slot entry, bword); the bword's bit-sliced table sits in shared memory.
template <int K>
__global__ void __launch_bounds__(128) k_barcode_bitsliced(SMX_KARGS) {
    __shared__ u32 s_beq[SMX_MAX_PATTERN * 16];
    const Tables &t = c_tables;
    const u32 g = blockIdx.y % t.n_bwords;
    const int strand = blockIdx.y / t.n_bwords;
    const int primer = t.bw_primer[g];
    const u32 slot = slot_index(t, strand, primer);
    u32 cnt = b.slot_count[slot];
    if (cnt > b.e_cap) cnt = b.e_cap;
    if (blockIdx.x * blockDim.x >= cnt) return;
    const int m = t.bw_len[g];
    __shared__ u32 s_acc;
    if (threadIdx.x == 0) s_acc = 0;
    for (int i = threadIdx.x; i < m * 16; i += blockDim.x) s_beq[i] = t.beq[(u64)t.bw_row[g] * 16 + i];
    __syncthreads();
    const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0, wcols = 0;
    if (idx < cnt) {
        const u32 read = b.ent_read[(u64)slot * b.e_cap + idx];
        const int p = b.ent_pos[(u64)slot * b.e_cap + idx];
        barcode_bitsliced_thread<K>(t, b, read, p, idx, strand, primer, g, s_beq, cells, wcols);
    }
    // m is uniform over the block: cells = m * S, word-columns = ceil(m/32) * S with S = sum of lanes x columns
    const unsigned long long wmul = (unsigned long long)((m + 31) >> 5);
    (void)cells;
    block_work_add((u32)(wcols / wmul), (unsigned long long)m, &b.counters[1], wmul, &b.counters[3], &s_acc);
}

constexpr int kInlineRecords = 4;

// Stage 3a, fast selection: one thread per read, only the common single-candidate path of the
// selection routine (slot digests computed on the fly, no grouping, small register footprint so
// the dependent global loads are hidden by occupancy).  Reads that need the general routine
// (several equal-best candidates, tied barcodes, TAILS trimming) are appended to defer_list.
template <int MAXP>
__global__ void __launch_bounds__(128) k_select_fast(SMX_KARGS) {
    u32 read = blockIdx.x * blockDim.x + threadIdx.x;
    bool defer = false;
    if (read < b.n_reads) {
        EndInfo ends[2 * MAXP];
        int ts_cand[1], ts_shift[1];
        smx_record rec;
        SelectStore st;
        st.groups = nullptr; st.gcand = nullptr; st.pg = nullptr; st.pcand = nullptr;
        st.ts_cand = ts_cand; st.ts_shift = ts_shift; st.cap = 1;
        SelectCtx c; c.t = &c_tables; c.b = &b; c.read = read; c.n = (int)b.lengths[read];
        unsigned char flags;
        u32 cnt = select_read_impl<true>(c, ends, st, &rec, 1, flags);
        defer = (flags & kFlagDeferred) != 0;
        if (!defer) {
            b.rec_count[read] = cnt;
            b.read_flags[read] = flags & 1;
            if (cnt) {
                uint4 *dst = reinterpret_cast<uint4 *>(b.rec_stage + read);
                const uint4 *src = reinterpret_cast<const uint4 *>(&rec);
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, defer);
    if (m) {
        const int lane = threadIdx.x & 31;
        u32 base = 0;
        if (lane == 0) base = atomicAdd((unsigned int *)&b.counters[kCtrDeferred], (u32)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (defer) b.defer_list[base + __popc(m & ((1u << lane) - 1))] = read;
    }
}

// Stage 3b, general selection over the deferred reads: working storage in thread-local arrays.
// Records are produced once into a small local buffer; the first goes to rec_stage[read], further
// ones (rare) to a contiguous block of rec_pool.  Reads whose groups overflow kSmallGroups or that
// emit more than kInlineRecords records are flagged (bit1) for k_select_big.
template <int MAXP>
__global__ void __launch_bounds__(128) k_select(SMX_KARGS) {
    // The routine is long and branchy: when few reads are deferred they are spread one per 8 lanes
    // so that a warp serialises 4 divergent reads instead of 32.
    const u32 n_def = (u32)b.counters[kCtrDeferred];
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool spread = (u64)n_def * 8 <= (u64)gridDim.x * blockDim.x;
    if (spread && (tid & 7)) return;
    const u32 i = spread ? tid >> 3 : tid;
    if (i >= n_def) return;
    const u32 read = b.defer_list[i];
    EndInfo ends[2 * MAXP];
    Group groups[kSmallGroups], pg[kSmallGroups];
    Cand gcand[kSmallGroups], pcand[kSmallGroups];
    int ts_cand[kSmallGroups], ts_shift[kSmallGroups];
    smx_record local[kInlineRecords];
    SelectStore st;
    st.groups = groups; st.gcand = gcand; st.pg = pg; st.pcand = pcand;
    st.ts_cand = ts_cand; st.ts_shift = ts_shift; st.cap = kSmallGroups;
    SelectCtx c; c.t = &c_tables; c.b = &b; c.read = read; c.n = (int)b.lengths[read];
    unsigned char flags;
    u32 cnt = select_read(c, ends, st, local, kInlineRecords, flags);
    if (cnt > kInlineRecords) flags |= 2;
    b.rec_count[read] = cnt;
    b.read_flags[read] = flags;
    if (flags & 2) {                                    // second pass: listed on the device, no host round trip
        b.big_list[atomicAdd((unsigned int *)&b.counters[kCtrBig], 1u)] = read;
        return;
    }
    if (cnt >= 1) b.rec_stage[read] = local[0];
    if (cnt >= 2) {
        u32 base = atomicAdd((unsigned int *)&b.counters[6] + 1, cnt - 1);
        b.rec_extra[read] = base;
        if (base + cnt - 1 <= b.pool_cap)
            for (u32 i2 = 1; i2 < cnt; ++i2) b.rec_pool[base + i2 - 1] = local[i2];
    }
}

// Second pass over the (rare) reads flagged by the first: same routine, kBigGroups-entry storage
// in global scratch.  One thread per flagged read, list built on the device by k_select.  The
// routine runs twice in the thread (count, then write) and the records take the same route as
// everybody else's -- first record staged per read, the rest in a contiguous pool block -- so that
// the scan + compaction that follows needs no special case.  Reads that overflow even this pass
// keep bit 1 and get bit 2 (the library reports them).
__global__ void __launch_bounds__(32) k_select_big(SMX_KARGS, const u32 *list, u32 n_list, unsigned char *scratch) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    u32 read = list[i];
    EndInfo *ends;
    SelectStore st = big_store(scratch + (size_t)i * kBigScratchBytes, ends);
    SelectCtx c; c.t = &c_tables; c.b = &b; c.read = read; c.n = (int)b.lengths[read];
    unsigned char flags;
    const u32 cnt = select_read(c, ends, st, nullptr, 0xFFFFFFFFu, flags);
    b.rec_count[read] = cnt;
    if (flags & 2) {
        b.read_flags[read] = (unsigned char)((flags & 1) | 2 | 4);
        return;
    }
    b.read_flags[read] = (unsigned char)(flags & 1);
    if (!cnt) return;
    const u32 base = atomicAdd((unsigned int *)&b.counters[6] + 1, cnt);
    if ((u64)base + cnt > b.pool_cap) return;           // pool overflow: the library grows it and re-runs selection
    select_read(c, ends, st, b.rec_pool + base, cnt, flags);
    b.rec_stage[read] = b.rec_pool[base];
    b.rec_extra[read] = base + 1;
}

// Stage 3c.  rec_count -> rec_offset (exclusive scan, n + 1 entries), the per-read flag counters, and
// the read-ordered compaction of the staged records, in ONE pass: a single-pass scan with decoupled
// look-back over 1024-read tiles (tile ids are taken from a ticket counter, so a tile only ever waits
// for tiles that already started), then each tile moves its own records (four threads per 64-byte
// record, 16-byte quarters, coalesced both ways).  Replaces a two-launch scan plus a compaction
// launch that needed a host round trip in between (profiles/r1_v14_ncu_full.md: 67 us for 3 MB of
// counts and 49 MB of records).
//
// Tile status word: epoch (30 bits) | state (2 bits: 1 = tile aggregate, 2 = inclusive prefix) |
// value (32 bits).  The epoch changes with every launch, so the status array is never cleared.
constexpr int kScanTile = 1024, kScanThreads = 256;

__global__ void __launch_bounds__(kScanThreads) k_scan_compact(SMX_KARGS, u32 rec_cap, unsigned long long *tile_status,
                                                               u32 *ticket, u32 ticket_base, u32 epoch) {
    __shared__ u32 s_off[kScanTile + 1];
    __shared__ unsigned char s_big[kScanTile];
    __shared__ u32 s_warp[kScanThreads / 32];
    __shared__ u32 s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;
    __syncthreads();
    const u32 tile = s_tile, n = b.n_reads;
    const u32 r0 = tile * kScanTile + threadIdx.x * 4;
    // four consecutive reads per thread
    u32 c[4] = {0, 0, 0, 0};
    unsigned f[4] = {0, 0, 0, 0};
    if (r0 + 3 < n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(b.rec_count + r0);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
        const uchar4 g = *reinterpret_cast<const uchar4 *>(b.read_flags + r0);
        f[0] = g.x; f[1] = g.y; f[2] = g.z; f[3] = g.w;
    } else {
        for (int i = 0; i < 4; ++i) if (r0 + i < n) { c[i] = b.rec_count[r0 + i]; f[i] = b.read_flags[r0 + i]; }
    }
    const u32 fm = (f[0] & 1) + (f[1] & 1) + (f[2] & 1) + (f[3] & 1);
    const u32 fo = ((f[0] >> 1) & 1) + ((f[1] >> 1) & 1) + ((f[2] >> 1) & 1) + ((f[3] >> 1) & 1);
    const u32 fh = ((f[0] >> 2) & 1) + ((f[1] >> 2) & 1) + ((f[2] >> 2) & 1) + ((f[3] >> 2) & 1);
    const u32 wm = __reduce_add_sync(0xffffffffu, fm), wo = __reduce_add_sync(0xffffffffu, fo), wh = __reduce_add_sync(0xffffffffu, fh);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        if (wm) atomicAdd(&b.counters[4], (unsigned long long)wm);            // reads with a full match
        if (wo) atomicAdd((unsigned int *)&b.counters[5], wo);                 // reads needing the big pass
        if (wh) atomicAdd((unsigned int *)&b.counters[5] + 1, wh);             // reads beyond even that
    }
    const u32 tsum = c[0] + c[1] + c[2] + c[3];
    u32 incl = tsum;
    for (int o = 1; o < 32; o <<= 1) {
        const u32 x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < kScanThreads / 32 ? s_warp[lane] : 0u, wi = w;
        for (int o = 1; o < kScanThreads / 32; o <<= 1) {
            const u32 x = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += x;
        }
        if (lane < kScanThreads / 32) s_warp[lane] = wi - w;                   // exclusive prefix of the warp totals
        const u32 total = __shfl_sync(0xffffffffu, wi, kScanThreads / 32 - 1);
        if (lane == 0) {
            const unsigned long long tag = (unsigned long long)epoch << 34;
            volatile unsigned long long *st = tile_status;
            u32 excl = 0;
            if (tile == 0) {
                st[0] = tag | (2ull << 32) | total;
            } else {
                st[tile] = tag | (1ull << 32) | total;
                __threadfence();
                for (u32 j = tile; j-- > 0;) {
                    unsigned long long v;
                    do { v = st[j]; } while ((v >> 34) != epoch || ((v >> 32) & 3ull) == 0);
                    excl += (u32)v;
                    if (((v >> 32) & 3ull) == 2ull) break;
                }
                st[tile] = tag | (2ull << 32) | (u32)(excl + total);
            }
            s_prefix = excl;
            if ((u64)(tile + 1) * kScanTile >= n) {                            // last tile: grand total
                b.rec_offset[n] = excl + total;
                *(u32 *)&b.counters[6] = excl + total;
            }
        }
    }
    __syncthreads();
    u32 off = s_prefix + s_warp[warp] + incl - tsum;
    for (int i = 0; i < 4; ++i) {
        s_off[threadIdx.x * 4 + i] = off;
        s_big[threadIdx.x * 4 + i] = (unsigned char)(f[i] & 2);
        if (r0 + i < n) b.rec_offset[r0 + i] = off;
        off += c[i];
    }
    if (threadIdx.x == kScanThreads - 1) s_off[kScanTile] = off;
    __syncthreads();
    // compaction of the tile's records: four quarters in flight per thread (loads first, then stores)
    const u32 tile_r0 = tile * kScanTile;
#pragma unroll 1
    for (u32 base = 0; base < kScanTile * 4; base += 4 * kScanThreads) {
        uint4 v[4];
        u32 o[4], cnt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 idx = base + k * kScanThreads + threadIdx.x, lr = idx >> 2, read = tile_r0 + lr;
            o[k] = s_off[lr];
            cnt[k] = read < n ? s_off[lr + 1] - o[k] : 0u;
            if (cnt[k] && (s_big[lr] || (u64)o[k] + cnt[k] > rec_cap)) cnt[k] = 0;      // k_select_big writes flagged reads
            if (cnt[k]) v[k] = reinterpret_cast<const uint4 *>(b.rec_stage + read)[idx & 3];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!cnt[k]) continue;
            const u32 idx = base + k * kScanThreads + threadIdx.x, q = idx & 3, read = tile_r0 + (idx >> 2);
            uint4 *dst = reinterpret_cast<uint4 *>(b.records + o[k]);
            dst[q] = v[k];
            if (cnt[k] > 1) {
                const uint4 *src = reinterpret_cast<const uint4 *>(b.rec_pool + b.rec_extra[read]);
                for (u32 i = q; i < 4 * (cnt[k] - 1); i += 4) dst[4 + i] = src[i];
            }
        }
    }
}

// smx_record -> smx_record32 (drops the four location pairs) ahead of the copy-out.
__global__ void __launch_bounds__(256) k_pack_records32(const smx_record *in, u32 n, smx_record32 *out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const smx_record r = in[i];
    smx_record32 o;
    o.read = r.read; o.sample = r.sample; o.trim_start = r.trim_start; o.trim_end = r.trim_end;
    o.pool = r.pool; o.p1 = r.p1; o.p2 = r.p2;
    o.dist[0] = r.dist[0]; o.dist[1] = r.dist[1]; o.dist[2] = r.dist[2]; o.dist[3] = r.dist[3];
    o.resolution = r.resolution;
    o.flags = (uint8_t)((r.reverse ? 1 : 0) | (r.trim_empty ? 2 : 0));
    o.candidate = r.candidate; o.pad[0] = o.pad[1] = o.pad[2] = 0;
    out[i] = o;
}

// rec_offset of a sub-batch in the caller's whole batch (its records start at rec_base).
__global__ void __launch_bounds__(256) k_rebase_offsets(const u32 *in, u32 n, u32 rec_base, u32 *out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + rec_base;
}

// Batched global (NW) distances for setup_match_parameters (orchestration.py:549-555):
// one thread per ordered pair (i, j), Myers with D[0][j] = j, score read at the last column.
__global__ void k_pairwise_nw(const char *seqs, const u32 *off, u32 n, i32 *out) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)n * n) return;
    u32 i = (u32)(idx / n), j = (u32)(idx % n);
    const char *a = seqs + off[i], *bseq = seqs + off[j];
    int m = (int)(off[i + 1] - off[i]), len = (int)(off[j + 1] - off[j]);
    if (m == 0 || len == 0) { out[idx] = m > len ? m : len; return; }
    // plain equality (edlib default alphabet, no additionalEqualities at orchestration.py:552)
    u64 Pv = pattern_mask<u64>(m), Mv = 0;
    int score = m;
    for (int x = 0; x < len; ++x) {
        char ch = bseq[x];
        u64 Eq = 0;
        for (int r = 0; r < m; ++r) if (a[r] == ch) Eq |= 1ull << (64 - m + r);
        score += myers_step<u64, true>(Eq, Pv, Mv);
    }
    out[idx] = score;
}
// Integer-ALU peak microbenchmark: 8 independent chains per thread, fully unrolled.
// MODE 0: LOP3 only, 1: IADD3 only, 2: alternating.
template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(u32 *out, int iters, u32 seed) {
    u32 a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed * (threadIdx.x + 1) + j * 0x9E3779B9u + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                u32 x = a[j], y = a[(j + 1) & 7], z = a[(j + 3) & 7];
                bool logic = MODE == 0 || (MODE == 2 && ((j + u) & 1));
                a[j] = logic ? ((x & y) ^ z) : (x + y + z);
            }
        }
    }
    u32 r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= a[j];
    if (r == 0x12345678u) out[0] = r;       // practically never; keeps the chains alive
}
#endif  // __CUDACC__

}  // namespace smx
