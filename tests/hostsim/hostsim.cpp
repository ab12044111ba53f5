// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY (never loaded by the specimux_b200 package).
//
// CPU simulator of the CUDA kernels: runs the very same __host__ __device__ per-thread routines
// (specimux_b200/csrc/smx_core.cuh, smx_kernels.cuh) in plain loops over the thread grid, so the
// kernel logic can be checked against the oracle in the GPU-less build container.  It is not a
// fallback: the product library has no CPU path and fails with SMX_ERR_NO_DEVICE without a GPU.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../specimux_b200/csrc/smx_host_tables.hpp"
#include "../../specimux_b200/csrc/smx_kernels.cuh"

using namespace smx;

static char g_err[512] = "";

// ---------------------------------------------------------------------------------------------
// Long primers (k_primer_long).  The CUDA kernel spreads the 32*SW-bit vectors over SW lanes and
// resolves carries with ballots; here the same recurrences run over an array of SW words with the
// carry rippled sequentially.  Same Peq layout (peq_long), same bookkeeping routines (primer_tail,
// write_entries).  long_carry_in<> -- the kernel's carry-lookahead -- is checked separately by
// hostsim_check_long_carry().
struct LongVec {
    int sw;
    std::vector<u32> Pv, Mv;
    explicit LongVec(int sw_) : sw(sw_), Pv(sw_), Mv(sw_) {}
    int step(const u32 *Eq /*sw words*/, bool shift_in_one) {
        std::vector<u32> Ph(sw), Mh(sw), Xv(sw);
        u32 carry = 0;
        for (int w = 0; w < sw; ++w) {
            const u32 a = Eq[w] & Pv[w];
            const unsigned long long full = (unsigned long long)a + Pv[w] + carry;
            carry = (u32)(full >> 32);
            const u32 Xh = ((u32)full ^ Pv[w]) | Eq[w];
            Xv[w] = Eq[w] | Mv[w];
            Ph[w] = Mv[w] | ~(Xh | Pv[w]);
            Mh[w] = Pv[w] & Xh;
        }
        const int d = (int)(Ph[sw - 1] >> 31) - (int)(Mh[sw - 1] >> 31);
        for (int w = sw - 1; w >= 0; --w) {
            const u32 plo = w ? Ph[w - 1] >> 31 : (shift_in_one ? 1u : 0u), mlo = w ? Mh[w - 1] >> 31 : 0u;
            Ph[w] = (Ph[w] << 1) | plo;
            Mh[w] = (Mh[w] << 1) | mlo;
        }
        for (int w = 0; w < sw; ++w) { Pv[w] = Mh[w] | ~(Xv[w] | Ph[w]); Mv[w] = Ph[w] & Xv[w]; }
        return d;
    }
};

// Returns the number of equal-best end locations of the slot (entries are written by the caller).
static int primer_long_host(const Tables &t, const Batch &b, u32 read, int strand, int primer) {
    const int sw = t.p_sw[primer], m = t.p_len[primer], k = t.p_k[primer];
    const u32 *peq = t.peq_long + t.p_long[primer];
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    const u32 slot = slot_index(t, strand, primer);
    const u64 hit_idx = (u64)slot * b.n_pad + read;
    u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + read;
    u32 *imask = b.impmask + (u64)slot * t.mw * b.n_pad + read;
    for (int w = 0; w < t.mw; ++w) { emask[(u64)w * b.n_pad] = 0; imask[(u64)w * b.n_pad] = 0; }
    int best = m + 1;
    {
        LongVec v(sw);
        std::fill(v.Pv.begin(), v.Pv.end(), ~0u);
        int score = m;
        for (int p = g.start; p < g.wl; ++p) {
            const int c = staged_sym(t, b, read, strand, p);
            score += v.step(peq + (size_t)c * sw, false);
            if (score < best) { best = score; imask[(u64)(p >> 5) * b.n_pad] |= 1u << (p & 31); }
            if (score == best) emask[(u64)(p >> 5) * b.n_pad] |= 1u << (p & 31);
        }
    }
    const int nloc = primer_tail(t, b, read, strand, primer, best);
    if (nloc) {
        const int first = b.phit[hit_idx].first_end - g.woff - g.delta;
        int rcols = first - g.start + 1;
        if (rcols > m + best) rcols = m + best;
        LongVec v(sw);
        for (int w = 0; w < sw; ++w) {
            const int lo = 32 * sw - m - 32 * w;
            v.Pv[w] = lo <= 0 ? ~0u : (lo >= 32 ? 0u : ~0u << lo);
        }
        int rs = m, last = m - 1;
        for (int j = 0; j < rcols; ++j) {
            const int c = staged_sym(t, b, read, strand, first - j);
            rs += v.step(peq + (size_t)16 * sw + (size_t)c * sw, true);
            if (rs == best) last = j;
        }
        b.phit[hit_idx].first_start = b.phit[hit_idx].first_end - last;
    }
    const unsigned long long hw_cols = (unsigned long long)(n < t.L ? n : t.L);
    b.counters[0] += hw_cols * (unsigned long long)m;
    b.counters[2] += hw_cols * (unsigned long long)((m + 31) >> 5);
    unsigned char ohit = 0;
    if (t.preorient && (!g.regular || read_is_flagged(b, read))) {
        LongVec v(sw);
        std::fill(v.Pv.begin(), v.Pv.end(), ~0u);
        int sc = m, bst = m + 1;
        const int cols = n < t.L ? n : t.L;
        for (int x = 0; x < cols; ++x) {
            sc += v.step(peq + (size_t)32 * sw + (size_t)sym_at(b, read, strand, x, n) * sw, false);
            if (sc < bst) bst = sc;
        }
        ohit = bst <= k;
    }
    b.orient_hit[hit_idx] = ohit;
    return nloc;
}

// Exhaustive / randomised check of the kernel's carry-lookahead against a rippled carry chain.
template <int SW> static int check_long_carry(unsigned long long trials) {
    unsigned long long x = 0x9E3779B97F4A7C15ull;
    for (unsigned long long it = 0; it < trials; ++it) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const u32 G = (u32)x & ~(u32)(x >> 32);       // generate and propagate are mutually exclusive
        const u32 P = (u32)(x >> 32) & ~G;
        u32 want = 0;
        for (int seg = 0; seg < 32 / SW; ++seg) {
            u32 c = 0;
            for (int w = 0; w < SW; ++w) {
                const int l = seg * SW + w;
                if (c) want |= 1u << l;
                c = ((G >> l) & 1u) | (((P >> l) & 1u) & c);
            }
        }
        if (long_carry_in<SW>(G, P) != want) return 1;
    }
    return 0;
}

extern "C" int hostsim_check_long_carry(unsigned long long trials) {
    return check_long_carry<3>(trials) | check_long_carry<4>(trials) | check_long_carry<5>(trials) | check_long_carry<6>(trials) |
           check_long_carry<8>(trials) | check_long_carry<10>(trials) | check_long_carry<16>(trials) | check_long_carry<32>(trials);
}

extern "C" const char *hostsim_last_error(void) { return g_err; }

extern "C" int hostsim_match_batch(const smx_tables *tb, const smx_params *pr, const smx_batch *in, smx_results *out) {
    HostTables ht;
    if (!ht.build(tb, pr)) { snprintf(g_err, sizeof(g_err), "%s", ht.error.c_str()); return SMX_ERR_ARG; }
    const Tables &t = ht.t;
    const u32 n = in->n_reads, n_pad = (n + 31u) & ~31u;      // whole groups of 32 reads for the sliced search
    const int nP = t.n_primers;
    std::vector<u32> win((size_t)2 * t.wpw * n_pad), win2((size_t)2 * t.nw2 * n_pad, 0), tmix((size_t)2 * nP * t.nw2 * n_pad + 1, 0), endmask((size_t)2 * nP * t.mw * n_pad), impmask((size_t)2 * nP * t.mw * n_pad), rec_count(n), rec_offset(n + 1);
    std::vector<smx_primer_hit> phit((size_t)2 * nP * n_pad);
    std::vector<unsigned char> orient_hit((size_t)2 * nP * n_pad), flags(n);
    std::vector<u32> slot_count((size_t)2 * nP + 1, 0), ent_base((size_t)2 * nP * n_pad + 1), ent_read;
    std::vector<unsigned short> ent_pos;
    std::vector<unsigned char> bh_count;
    std::vector<smx_barcode_hit> bh_list;
    std::vector<BarcodeDigest> bdig;
    u32 e_cap = 16;        // deliberately tiny: exercises the capacity re-run
    unsigned long long counters[kCtrWords] = {0};
    Batch b;
    memset(&b, 0, sizeof(b));
    b.n_reads = n; b.n_pad = n_pad; b.clip = in->clip_len; b.stride = in->stride_words;
    std::vector<u32> lengths32;
    if (in->lengths16) {                // 16-bit lengths of the wire form: what k_expand_lengths does
        lengths32.assign(in->lengths16, in->lengths16 + n);
    }
    b.packed2 = in->packed2; b.word_off = in->word_off; b.lengths = in->lengths16 ? lengths32.data() : in->lengths;
    bool flagged = in->packed4 && in->off4 && in->packed4_words;
    b.packed4 = flagged ? in->packed4 : nullptr; b.off4 = flagged ? in->off4 : nullptr;
    b.win = win.data(); b.win2 = win2.data(); b.tmix = tmix.data(); b.phit = phit.data(); b.endmask = endmask.data(); b.impmask = impmask.data(); b.orient_hit = orient_hit.data();
    b.slot_count = slot_count.data(); b.ent_base = ent_base.data();
    b.rec_count = rec_count.data(); b.rec_offset = rec_offset.data();
    b.read_flags = flags.data(); b.counters = counters;

    for (u32 r = 0; r < n; ++r)
        for (int s = 0; s < 2; ++s)
            stage_windows_thread(t, b, r, s, b.packed2, 0);
    if (t.sliced)       // forward pass bit-sliced across reads: one "thread" per (group of 32 reads, strand, primer)
        for (int p = 0; p < nP; ++p) {
            RowOffsets ro;
            bool degenerate = false;
            for (int i = 0; i < 32; ++i) {
                int code = i < t.p_len[p] ? ht.prow_code[(size_t)p * 32 + i] : 0;
                degenerate |= code > 3;
                ro.off[i] = (unsigned short)(code * sizeof(u32));
            }
            u32 scratch[kSlicedCodes], planes[48];
            for (int s = 0; s < 2; ++s)
                for (u32 g = 0; g < n_pad / 32; ++g)
                    switch (t.p_len[p]) {
#define SMX_M(MM) case MM: primer_sliced_thread<MM, 1>(t, b, g, s, p, ro, degenerate, scratch, planes); break;
                        SMX_M(1) SMX_M(2) SMX_M(3) SMX_M(4) SMX_M(5) SMX_M(6) SMX_M(7) SMX_M(8) SMX_M(9) SMX_M(10) SMX_M(11)
                        SMX_M(12) SMX_M(13) SMX_M(14) SMX_M(15) SMX_M(16) SMX_M(17) SMX_M(18) SMX_M(19) SMX_M(20) SMX_M(21)
                        SMX_M(22) SMX_M(23) SMX_M(24) SMX_M(25) SMX_M(26) SMX_M(27) SMX_M(28) SMX_M(29) SMX_M(30) SMX_M(31)
                        SMX_M(32)
#undef SMX_M
                        default: snprintf(g_err, sizeof(g_err), "sliced primer length"); return SMX_ERR_INTERNAL;
                    }
        }
    Tables &tm = ht.t;
    for (;;) {      // stage 1 + 2, re-run on capacity overflow exactly as the CUDA library does
        ent_read.assign((size_t)2 * nP * e_cap + 1, 0); ent_pos.assign((size_t)2 * nP * e_cap + 1, 0);
        bh_count.assign((size_t)2 * t.n_bwords * e_cap + 1, 0);
        bh_list.assign((size_t)2 * t.n_bwords * t.hit_cap * e_cap + 1, smx_barcode_hit());
        b.e_cap = e_cap; b.ent_read = ent_read.data(); b.ent_pos = ent_pos.data();
        b.bh_count = bh_count.data(); b.bh_list = bh_list.data();
        std::fill(slot_count.begin(), slot_count.end(), 0u);
        counters[1] = counters[3] = counters[7] = 0;
        for (int s = 0; s < 2; ++s)
            for (int p = 0; p < nP; ++p)
                for (u32 r = 0; r < n; ++r) {
                    int nloc;
                    if (t.p_sw[p]) nloc = primer_long_host(t, b, r, s, p);
                    else if (t.use64) nloc = primer_finish_thread<u64>(t, b, r, s, p, t.peq_rc + p * 16, t.peq_fw + p * 16);
                    else nloc = primer_finish_thread<u32>(t, b, r, s, p, t.peq_rc + p * 16, t.peq_fw + p * 16);
                    if (nloc) {
                        u32 slot = (u32)(s * nP + p);
                        write_entries(t, b, slot, r, slot_count[slot]);
                        slot_count[slot] += (u32)nloc;
                    }
                }
        u32 max_entries = 0;
        for (u32 v : slot_count) max_entries = std::max(max_entries, v);
        if (max_entries > e_cap) { e_cap = max_entries; continue; }
        for (u32 slot = 0; slot < (u32)(2 * nP); ++slot) {
            const int p = (int)(slot % nP);
            if (t.p_sw[p]) continue;                 // long primer: start already recovered
            const u64 *rev = t.peq_rcrev + (size_t)p * 16;
            const bool sliced_start = start_sliced_ok(t, p);
            // what k_primer_finish keeps of the single-word form: every slot of a primer without the sliced start,
            // otherwise only the reads on the 4-bit side stream
            for (u32 e = 0; e < slot_count[slot]; ++e) {
                if (sliced_start && !read_is_flagged(b, ent_read[(size_t)slot * e_cap + e])) continue;
                if (t.use64) primer_start_thread<u64>(t, b, slot, e, rev); else primer_start_thread<u32>(t, b, slot, e, rev);
            }
            if (!sliced_start) continue;
            RowOffsets ro;
            bool degenerate = false;
            const int m = t.p_len[p];
            for (int i = 0; i < 32; ++i) {
                int code = i < m ? ht.prow_code[(size_t)p * 32 + (m - 1 - i)] : 0;
                degenerate |= code > 3;
                ro.off[i] = (unsigned short)(code * sizeof(u32));
            }
            u32 scratch[kSlicedCodes], planes[kStartPlanes];
            for (u32 e0 = 0; e0 < slot_count[slot]; e0 += 32) {
                for (u32 r = 0; r < 32; ++r)
                    start_gather_entry(t, b, slot, e0 + r, slot_count[slot], m, planes[r], planes[32 + r], planes[64 + r], planes[96 + r]);
                const int ncols = m + (int)t.p_k[p];
                switch (m) {
#define SMX_M(MM) case MM: primer_start_sliced_thread<MM, 1>(ncols, ro, degenerate, scratch, planes); break;
                    SMX_M(1) SMX_M(2) SMX_M(3) SMX_M(4) SMX_M(5) SMX_M(6) SMX_M(7) SMX_M(8) SMX_M(9) SMX_M(10) SMX_M(11)
                    SMX_M(12) SMX_M(13) SMX_M(14) SMX_M(15) SMX_M(16) SMX_M(17) SMX_M(18) SMX_M(19) SMX_M(20) SMX_M(21)
                    SMX_M(22) SMX_M(23) SMX_M(24) SMX_M(25) SMX_M(26) SMX_M(27) SMX_M(28) SMX_M(29) SMX_M(30) SMX_M(31)
                    SMX_M(32)
#undef SMX_M
                    default: snprintf(g_err, sizeof(g_err), "sliced start: primer length"); return SMX_ERR_INTERNAL;
                }
                for (u32 r = 0; r < 32; ++r)
                    if (planes[r] != 0xFFFFFFFFu) start_store_entry(t, b, slot, e0 + r, (int)planes[r]);
            }
        }
        bdig.assign((size_t)2 * t.n_btasks * e_cap + 1, BarcodeDigest());
        b.bdig = bdig.data();
        for (int s = 0; s < 2; ++s)
            for (int tk = 0; tk < t.n_btasks; ++tk) {
                const u32 g0 = t.bt_g0[tk];
                const u32 *tab = t.bt_eq + t.bt_row[tk];
                const int p = t.bw_primer[g0], m = t.bw_len[g0], nw = t.bt_nw[tk];
                u32 slot = (u32)(s * nP + p);
                if (t.bt_quad_row[tk] >= 0 && t.k_idx <= kMaxTaskK) {          // narrow word: four entries to a "thread"
                    const u32 *tab4 = t.bt_quad + t.bt_quad_row[tk];
                    for (u32 e0 = 0; e0 < slot_count[slot]; e0 += 4) {
                        u32 work = 0;
                        switch (t.k_idx) {
#define SMX_Q(KK) case KK: work = (m == 13 && KK >= 1 && KK <= 3) \
        ? barcode_quad_thread<KK, (KK >= 1 && KK <= 3) ? 13 : 0>(t, b, slot, e0, slot_count[slot], s, p, (u32)tk, tab4, tab) \
        : barcode_quad_thread<KK, 0>(t, b, slot, e0, slot_count[slot], s, p, (u32)tk, tab4, tab); break;
                            SMX_Q(0) SMX_Q(1) SMX_Q(2) SMX_Q(3) SMX_Q(4)
#undef SMX_Q
                            default: break;
                        }
                        counters[1] += (unsigned long long)(work & 0xFFFFFu) * m;
                        counters[3] += (unsigned long long)(work & 0xFFFFFu) * ((m + 31) >> 5);
                    }
                    continue;
                }
                for (u32 e = 0; e < slot_count[slot]; ++e) {
                    u32 r = ent_read[(size_t)slot * e_cap + e];
                    int pos = ent_pos[(size_t)slot * e_cap + e];
                    u32 work = 0;
                    switch (t.k_idx * 8 + nw) {
// 13-nt barcodes take the fixed-length instantiation, as the CUDA launcher does (smx_k_stage2.cu)
#define SMX_K2(KK, NN) case KK * 8 + NN: work = (m == 13 && KK >= 1 && KK <= 3) \
        ? barcode_task_thread<KK, NN, (KK >= 1 && KK <= 3) ? 13 : 0>(t, b, r, pos, e, s, p, (u32)tk, tab) \
        : barcode_task_thread<KK, NN>(t, b, r, pos, e, s, p, (u32)tk, tab); break;
#define SMX_K2M(KK) SMX_K2(KK, 1) SMX_K2(KK, 2) SMX_K2(KK, 3) SMX_K2(KK, 4)
                        SMX_K2M(0) SMX_K2M(1) SMX_K2M(2) SMX_K2M(3) SMX_K2M(4)
                        SMX_K2(5, 1) SMX_K2(6, 1) SMX_K2(7, 1) SMX_K2(8, 1) SMX_K2(9, 1) SMX_K2(10, 1) SMX_K2(11, 1) SMX_K2(12, 1)
#undef SMX_K2M
#undef SMX_K2
                        default: snprintf(g_err, sizeof(g_err), "unsupported k_idx / task width"); return SMX_ERR_ARG;
                    }
                    counters[1] += (unsigned long long)(work & 0xFFFFFu) * m;
                    counters[3] += (unsigned long long)(work & 0xFFFFFu) * ((m + 31) >> 5);
                }
            }
        if (counters[7] == 0) break;
        if (tm.hit_cap >= kMaxWordHits) { snprintf(g_err, sizeof(g_err), "hit list overflow"); return SMX_ERR_INTERNAL; }
        tm.hit_cap = std::min(kMaxWordHits, tm.hit_cap * 4);
    }
    std::vector<unsigned char> big(kBigScratchBytes);
    auto run_select = [&](u32 r, smx_record *dst, unsigned char &f) -> u32 {
        SelectCtx c; c.t = &t; c.b = &b; c.read = r; c.n = (int)b.lengths[r];
        EndInfo ends[2 * SMX_MAX_PRIMERS];
        Group groups[kSmallGroups], pg[kSmallGroups];
        Cand gcand[kSmallGroups], pcand[kSmallGroups];
        int ts_cand[kSmallGroups], ts_shift[kSmallGroups];
        SelectStore st;
        st.groups = groups; st.gcand = gcand; st.pg = pg; st.pcand = pcand;
        st.ts_cand = ts_cand; st.ts_shift = ts_shift; st.cap = kSmallGroups;
        // the kernels' dispatch: a compile-time primer count (slots loaded ahead) for 1..4 primers; the general pass
        // carries the equal-best list cache, the second pass on the big scratch walks the lists
        const int np = t.n_primers;
        {   // the fast-only pass first, exactly as k_select_fast does; deferred reads take the general routine
            SelectStore fst = st;
            fst.cap = 1;
            smx_record one;
            u32 fc = np == 1 ? select_read_impl<true, 1>(c, ends, fst, dst ? &one : nullptr, 1, f)
                   : np == 2 ? select_read_impl<true, 2>(c, ends, fst, dst ? &one : nullptr, 1, f)
                   : np == 3 ? select_read_impl<true, 3>(c, ends, fst, dst ? &one : nullptr, 1, f)
                   : np == 4 ? select_read_impl<true, 4>(c, ends, fst, dst ? &one : nullptr, 1, f)
                             : select_read_impl<true>(c, ends, fst, dst ? &one : nullptr, 1, f);
            if (!(f & kFlagDeferred)) {
                if (dst && fc) dst[0] = one;
                f &= 1;
                return fc;
            }
        }
        std::vector<BestList> best(2 * (size_t)np);
        for (auto &bl : best) bl.n = -1;
        c.best = best.data();
        u32 cnt = np == 2 ? select_read<2>(c, ends, st, dst, 0xFFFFFFFFu, f) : select_read(c, ends, st, dst, 0xFFFFFFFFu, f);
        c.best = nullptr;
        if (f & 2) {            // second pass on the big scratch, as k_select_big does
            EndInfo *bends;
            SelectStore bst = big_store(big.data(), bends);
            cnt = select_read(c, bends, bst, dst, 0xFFFFFFFFu, f);
            if (f & 2) f |= 4;
        }
        return cnt;
    };
    u64 total = 0, matched = 0;
    for (u32 r = 0; r < n; ++r) {
        unsigned char f;
        rec_count[r] = run_select(r, nullptr, f);
        flags[r] = f;
        rec_offset[r] = (u32)total;
        total += rec_count[r];
        matched += f & 1;
        if (f & 4) { snprintf(g_err, sizeof(g_err), "read %u exceeds %d dereplication groups", r, kBigGroups); return SMX_ERR_INTERNAL; }
    }
    rec_offset[n] = (u32)total;
    out->n_records = total; out->n_matched = matched;
    if (total > out->records_cap) { snprintf(g_err, sizeof(g_err), "capacity"); return SMX_ERR_CAPACITY; }
    std::vector<smx_record> records(total + 1);
    b.records = records.data();
    for (u32 r = 0; r < n; ++r) {
        unsigned char f;
        run_select(r, records.data() + rec_offset[r], f);
    }
    if (out->rec_offset) memcpy(out->rec_offset, rec_offset.data(), (size_t)(n + 1) * sizeof(u32));
    if (out->records && total) memcpy(out->records, records.data(), total * sizeof(smx_record));
    if (!out->records && out->records32)
        for (u64 i = 0; i < total; ++i) pack_record32(records[i], out->records32[i]);      // what k_pack_records32 does
    if (!out->records && !out->records32 && out->records16)
        for (u64 i = 0; i < total; ++i) {           // what k_pack_records16 does
            const bool last = i + 1 >= total || records[i + 1].read != records[i].read;
            if (!pack_record16(records[i], (int)b.lengths[records[i].read], last, out->records16[i])) {
                snprintf(g_err, sizeof(g_err), "record does not fit smx_record16");
                return SMX_ERR_INTERNAL;
            }
        }
    // level-1 detail is kept padded ([row][n_pad]) and returned dense ([row][n])
    if (out->primer_hits)
        for (size_t row = 0; row < (size_t)2 * nP; ++row)
            memcpy(out->primer_hits + row * n, phit.data() + row * n_pad, (size_t)n * sizeof(smx_primer_hit));
    if (out->orient_hits)
        for (size_t row = 0; row < (size_t)2 * nP; ++row)
            memcpy(out->orient_hits + row * n, orient_hit.data() + row * n_pad, (size_t)n);
    if (out->endmask_bits)
        for (size_t row = 0; row < (size_t)2 * nP * t.mw; ++row)
            memcpy(out->endmask_bits + row * n * sizeof(u32), endmask.data() + row * n_pad, (size_t)n * sizeof(u32));
    if (out->barcode_hits && t.total_bslots) {
        smx_barcode_hit none;
        none.end_mask = 0; none.search_start = 0; none.distance = -1; none.barcode = 0;
        for (size_t i = 0; i < (size_t)t.total_bslots * n; ++i) out->barcode_hits[i] = none;
        uint64_t n_loc_hits = 0;
        for (int sd = 0; sd < 2; ++sd)
            for (int g = 0; g < t.n_bwords; ++g) {
                int p = t.bw_primer[g];
                u32 slot = (u32)(sd * nP + p);
                u64 gslot = (u64)sd * t.n_bwords + g;
                for (u32 r = 0; r < n; ++r) {
                    const smx_primer_hit &h0 = phit[(size_t)slot * n_pad + r];
                    if (h0.distance < 0) continue;
                    for (u32 l = 0; l < h0.n_locations; ++l) {
                        u64 e = (u64)ent_base[(size_t)slot * n_pad + r] + l;
                        int k = std::min<int>(bh_count[gslot * e_cap + e], t.hit_cap);
                        for (int x = 0; x < k; ++x) {
                            const smx_barcode_hit &h = bh_list[(gslot * t.hit_cap + x) * e_cap + e];
                            smx_barcode_hit &dst = out->barcode_hits[((size_t)t.bslot_base[slot] + h.barcode) * n + r];
                            if (dst.distance < 0 || h.distance < dst.distance) dst = h;
                            if (out->barcode_loc_hits) {
                                if (n_loc_hits < out->barcode_loc_cap) {
                                    smx_barcode_loc_hit &lh = out->barcode_loc_hits[n_loc_hits];
                                    lh.read = r; lh.slot = (uint16_t)slot; lh.location = (uint16_t)l; lh.hit = h;
                                }
                                ++n_loc_hits;
                            }
                        }
                    }
                }
            }
        out->n_barcode_loc_hits = n_loc_hits;
    }
    return SMX_OK;
}
