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

extern "C" const char *hostsim_last_error(void) { return g_err; }

extern "C" int hostsim_match_batch(const smx_tables *tb, const smx_params *pr, const smx_batch *in, smx_results *out) {
    HostTables ht;
    if (!ht.build(tb, pr)) { snprintf(g_err, sizeof(g_err), "%s", ht.error.c_str()); return SMX_ERR_ARG; }
    const Tables &t = ht.t;
    const u32 n = in->n_reads, n_pad = n;
    const int nP = t.n_primers;
    std::vector<u32> win((size_t)2 * t.wpw * n_pad), endmask((size_t)2 * nP * t.mw * n_pad), rec_count(n), rec_offset(n + 1);
    std::vector<smx_primer_hit> phit((size_t)2 * nP * n_pad);
    std::vector<unsigned char> orient_hit((size_t)2 * nP * n_pad), flags(n);
    std::vector<smx_barcode_hit> bhit((size_t)t.total_bslots * n_pad);
    unsigned long long counters[8] = {0};
    Batch b;
    memset(&b, 0, sizeof(b));
    b.n_reads = n; b.n_pad = n_pad;
    b.packed2 = in->packed2; b.word_off = in->word_off; b.lengths = in->lengths;
    bool flagged = in->packed4 && in->off4 && in->packed4_words;
    b.packed4 = flagged ? in->packed4 : nullptr; b.off4 = flagged ? in->off4 : nullptr;
    b.win = win.data(); b.phit = phit.data(); b.endmask = endmask.data(); b.orient_hit = orient_hit.data();
    b.bhit = bhit.data(); b.rec_count = rec_count.data(); b.rec_offset = rec_offset.data();
    b.read_flags = flags.data(); b.counters = counters;

    for (u32 r = 0; r < n; ++r)
        for (int s = 0; s < 2; ++s)
            for (int w = 0; w < t.wpw; ++w) stage_window_word(t, b, r, s, w);
    for (int s = 0; s < 2; ++s)
        for (int p = 0; p < nP; ++p)
            for (u32 r = 0; r < n; ++r) {
                if (t.use64) primer_search_thread<u64>(t, b, r, s, p, t.peq_rc + p * 16, t.peq_rcrev + p * 16, t.peq_fw + p * 16);
                else primer_search_thread<u32>(t, b, r, s, p, t.peq_rc + p * 16, t.peq_rcrev + p * 16, t.peq_fw + p * 16);
            }
    for (int s = 0; s < 2; ++s)
        for (int p = 0; p < nP; ++p) {
            int nb = (int)(t.pb_off[p + 1] - t.pb_off[p]);
            for (u32 r = 0; r < n; ++r)
                for (int j = 0; j < nb; ++j) {
                    const u64 *peq = t.bpeq + ((size_t)t.pb_off[p] + j) * 16;
                    int m = t.b_len[t.pb_off[p] + j];
                    if (t.buse64) barcode_search_thread<u64>(t, b, r, s, p, j, peq, m, counters[1], counters[3]);
                    else barcode_search_thread<u32>(t, b, r, s, p, j, peq, m, counters[1], counters[3]);
                }
        }
    std::vector<EndInfo> ends(2 * SMX_MAX_PRIMERS);
    u64 total = 0, matched = 0;
    for (u32 r = 0; r < n; ++r) {
        SelectCtx c; c.t = &t; c.b = &b; c.read = r; c.n = (int)b.lengths[r];
        unsigned char f;
        rec_count[r] = select_read(c, ends.data(), nullptr, f);
        flags[r] = f;
        rec_offset[r] = (u32)total;
        total += rec_count[r];
        matched += f & 1;
        if (f & 2) { snprintf(g_err, sizeof(g_err), "read %u exceeded an internal tie/group capacity", r); return SMX_ERR_INTERNAL; }
    }
    rec_offset[n] = (u32)total;
    out->n_records = total; out->n_matched = matched;
    if (total > out->records_cap) { snprintf(g_err, sizeof(g_err), "capacity"); return SMX_ERR_CAPACITY; }
    std::vector<smx_record> records(total + 1);
    b.records = records.data();
    for (u32 r = 0; r < n; ++r) {
        SelectCtx c; c.t = &t; c.b = &b; c.read = r; c.n = (int)b.lengths[r];
        unsigned char f;
        select_read(c, ends.data(), records.data() + rec_offset[r], f);
    }
    if (out->rec_offset) memcpy(out->rec_offset, rec_offset.data(), (size_t)(n + 1) * sizeof(u32));
    if (out->records && total) memcpy(out->records, records.data(), total * sizeof(smx_record));
    if (out->primer_hits) memcpy(out->primer_hits, phit.data(), phit.size() * sizeof(smx_primer_hit));
    if (out->endmask_bits) memcpy(out->endmask_bits, endmask.data(), endmask.size() * sizeof(u32));
    if (out->barcode_hits && !bhit.empty()) memcpy(out->barcode_hits, bhit.data(), bhit.size() * sizeof(smx_barcode_hit));
    return SMX_OK;
}
