// smx_api.cu -- C ABI of libspecimux_b200.so (see include/specimux_b200.h).
// Context / table construction, batch buffers, kernel launches.  No CPU matching path exists:
// without a usable GPU every compute entry point fails with SMX_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "smx_host_tables.hpp"
#include "smx_kernels.cuh"

using namespace smx;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver          \
                            ? SMX_ERR_NO_DEVICE : SMX_ERR_CUDA,                               \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T> cudaError_t upload(DevBuf<T> &d, const std::vector<T> &h) {
    cudaError_t e = d.ensure(h.size() ? h.size() : 1);
    if (e != cudaSuccess) return e;
    if (h.empty()) return cudaSuccess;
    return cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

struct smx_ctx {
    int device = 0;
    Tables t;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // table storage
    DevBuf<u64> peq_rc, peq_rcrev, peq_fw, spec_key, spec_p1, spec_p2;
    DevBuf<unsigned char> b_len, bw_len, bw_primer;
    DevBuf<u32> pb_barcode, pair_fwd, pair_rev, spec_key_off, spec_row, bw_row, bw_valid, beq;
    DevBuf<unsigned short> bw_list;
    DevBuf<i32> pair_pool, spec_pool, spec_dense;
    DevBuf<SlotSum> ssum;
    int max_nb = 0;
    // batch storage
    Batch b;
    DevBuf<u32> packed2, lengths, packed4, win, endmask, impmask, rec_count, rec_offset, block_sums;
    DevBuf<u64> word_off, off4;
    DevBuf<smx_primer_hit> phit;
    DevBuf<unsigned char> orient_hit, read_flags, bh_count;
    DevBuf<smx_barcode_hit> bh_list;
    DevBuf<u32> slot_count, ent_base, ent_read, rec_extra, big_list;
    DevBuf<unsigned short> ent_pos;
    DevBuf<smx_record> rec_stage, rec_pool;
    u32 e_cap = 0, pool_cap = 0;
    DevBuf<unsigned char> big_scratch;
    DevBuf<smx_record> records;
    DevBuf<unsigned long long> counters;   // 4 work counters + matched + (u32) overflow + (u32) total
    bool have_batch = false, have_results = false;
    u64 n_records = 0, n_matched = 0;
    unsigned long long work[4] = {0, 0, 0, 0};
    float total_ms = 0, stage_ms[4] = {0, 0, 0, 0};
    int launches = 0;
};

static cudaError_t ensure_entry_buffers(smx_ctx *c) {
    const Tables &t = c->t;
    cudaError_t e;
    if ((e = c->ent_read.ensure((size_t)2 * t.n_primers * c->e_cap)) != cudaSuccess) return e;
    if ((e = c->ent_pos.ensure((size_t)2 * t.n_primers * c->e_cap)) != cudaSuccess) return e;
    if ((e = c->bh_count.ensure((size_t)2 * t.n_bwords * c->e_cap + 1)) != cudaSuccess) return e;
    if ((e = c->bh_list.ensure((size_t)2 * t.n_bwords * t.hit_cap * c->e_cap + 1)) != cudaSuccess) return e;
    return c->rec_pool.ensure(c->pool_cap);
}

static void bind_entry_buffers(smx_ctx *c) {
    Batch &b = c->b;
    b.e_cap = c->e_cap; b.pool_cap = c->pool_cap;
    b.ent_read = c->ent_read.p; b.ent_pos = c->ent_pos.p; b.bh_count = c->bh_count.p; b.bh_list = c->bh_list.p;
    b.rec_pool = c->rec_pool.p;
}

// rec_count -> rec_offset (exclusive scan), flag counters, then one D2H of the 8 counters.
static cudaError_t scan_and_count(smx_ctx *c, int &launches, unsigned long long host_counters[8]) {
    Batch &b = c->b;
    const u32 n = b.n_reads;
    cudaStream_t st = c->stream;
    unsigned sblocks = (n + kScanBlock - 1) / kScanBlock;
    k_scan_block_sums<<<sblocks, kScanBlock, 0, st>>>(b.rec_count, n, c->block_sums.p);
    k_scan_spine<<<1, kScanBlock, 0, st>>>(c->block_sums.p, sblocks, (u32 *)(c->counters.p + 6));
    k_scan_apply<<<sblocks, kScanBlock, 0, st>>>(b.rec_count, n, c->block_sums.p, b.rec_offset);
    k_count_flags<<<(n + 255) / 256, 256, 0, st>>>(b.read_flags, n, c->counters.p + 4, (unsigned *)(c->counters.p + 5));
    launches += 4;
    cudaError_t e = cudaMemcpyAsync(host_counters, c->counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

extern "C" {

int smx_abi_version(void) { return SMX_ABI_VERSION; }

const char *smx_last_error(void) { return g_err; }

int smx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void smx_destroy(smx_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->peq_rc.release(); c->peq_rcrev.release(); c->peq_fw.release(); c->bw_len.release(); c->bw_primer.release(); c->bw_row.release();
    c->bw_valid.release(); c->beq.release(); c->bw_list.release(); c->bh_count.release(); c->bh_list.release();
    c->slot_count.release(); c->ent_base.release(); c->ent_read.release(); c->ent_pos.release();
    c->rec_extra.release(); c->rec_stage.release(); c->rec_pool.release(); c->big_list.release(); c->big_scratch.release();
    c->spec_key.release(); c->spec_p1.release(); c->spec_p2.release(); c->b_len.release();
    c->pb_barcode.release(); c->pair_fwd.release(); c->pair_rev.release(); c->spec_key_off.release();
    c->spec_row.release(); c->pair_pool.release(); c->spec_pool.release(); c->spec_dense.release(); c->ssum.release();
    c->packed2.release(); c->lengths.release(); c->packed4.release(); c->win.release(); c->endmask.release(); c->impmask.release();
    c->rec_count.release(); c->rec_offset.release(); c->block_sums.release(); c->word_off.release();
    c->off4.release(); c->phit.release(); c->orient_hit.release(); c->read_flags.release();
    c->records.release(); c->counters.release();
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int smx_create(int device, const smx_tables *tb, const smx_params *pr, smx_ctx **out) {
    if (!tb || !pr || !out) return fail(SMX_ERR_ARG, "smx_create: null argument");
    *out = nullptr;
    HostTables ht;
    if (!ht.build(tb, pr)) return fail(SMX_ERR_ARG, "smx_create: %s", ht.error.c_str());
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SMX_ERR_NO_DEVICE, "smx_create: no CUDA device (%s); this library has no CPU path",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    }
    if (device < 0 || device >= ndev) return fail(SMX_ERR_ARG, "smx_create: device %d of %d", device, ndev);

    smx_ctx *c = new smx_ctx();
    c->device = device;
    c->max_nb = ht.max_nb;
#define CUC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            int rc_ = fail(SMX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));             \
            smx_destroy(c);                                                                           \
            return rc_;                                                                               \
        }                                                                                             \
    } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CUC(cudaEventCreate(&e));
    CUC(upload(c->peq_rc, ht.peq_rc)); CUC(upload(c->peq_rcrev, ht.peq_rcrev)); CUC(upload(c->peq_fw, ht.peq_fw));
    CUC(upload(c->b_len, ht.b_len)); CUC(upload(c->pb_barcode, ht.pb_barcode));
    CUC(upload(c->bw_len, ht.bw_len)); CUC(upload(c->bw_primer, ht.bw_primer)); CUC(upload(c->bw_row, ht.bw_row));
    CUC(upload(c->bw_valid, ht.bw_valid)); CUC(upload(c->bw_list, ht.bw_list)); CUC(upload(c->beq, ht.beq));
    CUC(upload(c->pair_fwd, ht.pair_fwd)); CUC(upload(c->pair_rev, ht.pair_rev)); CUC(upload(c->pair_pool, ht.pair_pool));
    CUC(upload(c->spec_key, ht.spec_key)); CUC(upload(c->spec_key_off, ht.spec_key_off));
    CUC(upload(c->spec_row, ht.spec_row)); CUC(upload(c->spec_p1, ht.spec_p1)); CUC(upload(c->spec_p2, ht.spec_p2));
    CUC(upload(c->spec_pool, ht.spec_pool));
    CUC(upload(c->spec_dense, ht.spec_dense));
    CUC(c->counters.ensure(8));
    ht.set_bword_pointers(c->bw_len.p, c->bw_primer.p, c->bw_row.p, c->bw_valid.p, c->bw_list.p, c->beq.p);
    ht.set_pointers(c->peq_rc.p, c->peq_rcrev.p, c->peq_fw.p, c->b_len.p, c->pb_barcode.p,
                    c->pair_fwd.p, c->pair_rev.p, c->pair_pool.p, c->spec_key.p, c->spec_key_off.p,
                    c->spec_row.p, c->spec_p1.p, c->spec_p2.p, c->spec_pool.p);
    c->t = ht.t;
    c->t.spec_dense = ht.spec_dense.empty() ? nullptr : c->spec_dense.p;
    memset(&c->b, 0, sizeof(c->b));
    if (ht.t.k_idx > 8) {
        smx_destroy(c);
        return fail(SMX_ERR_ARG, "smx_create: barcode distance threshold %d exceeds the supported 8", ht.t.k_idx);
    }
#undef CUC
    *out = c;
    return SMX_OK;
}

uint64_t smx_result_bound(const smx_ctx *c, uint32_t n_reads) {
    // Every read yields >= 1 record; multi-record reads (tied specimens / partial barcodes) are rare.
    // smx_download_results reports SMX_ERR_CAPACITY with the exact need if this is ever exceeded.
    (void)c;
    return (uint64_t)n_reads * 2 + 1024;
}

int smx_upload_batch(smx_ctx *c, const smx_batch *in) {
    if (!c || !in) return fail(SMX_ERR_ARG, "smx_upload_batch: null argument");
    if (in->n_reads == 0) return fail(SMX_ERR_ARG, "smx_upload_batch: empty batch");
    if (in->clip_len && (int)in->clip_len < c->t.L)
        return fail(SMX_ERR_ARG, "smx_upload_batch: clip_len %u is below search_len %d", in->clip_len, c->t.L);
    if ((in->packed4 == nullptr) != (in->off4 == nullptr) && in->packed4_words)
        return fail(SMX_ERR_ARG, "smx_upload_batch: packed4 and off4 must be given together");
    CU(cudaSetDevice(c->device));
    const Tables &t = c->t;
    const u32 n = in->n_reads, n_pad = (n + 127u) & ~127u;
    const int nP = t.n_primers;
    CU(c->packed2.ensure(in->packed2_words + 1)); CU(c->word_off.ensure(n)); CU(c->lengths.ensure(n));
    bool flagged = in->packed4 && in->off4 && in->packed4_words;
    if (flagged) { CU(c->packed4.ensure(in->packed4_words)); CU(c->off4.ensure(n)); }
    CU(c->win.ensure((size_t)2 * t.wpw * n_pad));
    CU(c->phit.ensure((size_t)2 * nP * n_pad));
    CU(c->endmask.ensure((size_t)2 * nP * t.mw * n_pad));
    CU(c->impmask.ensure((size_t)2 * nP * t.mw * n_pad));
    CU(c->orient_hit.ensure((size_t)2 * nP * n_pad));
    CU(c->slot_count.ensure((size_t)2 * nP)); CU(c->ent_base.ensure((size_t)2 * nP * n_pad));
    CU(c->ssum.ensure((size_t)2 * nP * n_pad));
    if (c->e_cap < n_pad) c->e_cap = n_pad;
    if (c->pool_cap < n_pad / 8 + 1024) c->pool_cap = n_pad / 8 + 1024;
    CU(ensure_entry_buffers(c));
    CU(c->rec_stage.ensure(n_pad)); CU(c->rec_extra.ensure(n_pad)); CU(c->rec_pool.ensure(c->pool_cap));
    CU(c->rec_count.ensure(n)); CU(c->rec_offset.ensure((size_t)n + 1)); CU(c->read_flags.ensure(n));
    CU(c->block_sums.ensure((n + kScanBlock - 1) / kScanBlock + 1));
    CU(cudaMemcpyAsync(c->packed2.p, in->packed2, in->packed2_words * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->word_off.p, in->word_off, (size_t)n * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->lengths.p, in->lengths, (size_t)n * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
    if (flagged) {
        CU(cudaMemcpyAsync(c->packed4.p, in->packed4, in->packed4_words * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->off4.p, in->off4, (size_t)n * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    }
    Batch &b = c->b;
    b.n_reads = n; b.n_pad = n_pad; b.clip = in->clip_len;
    b.packed2 = c->packed2.p; b.word_off = c->word_off.p; b.lengths = c->lengths.p;
    b.packed4 = flagged ? c->packed4.p : nullptr; b.off4 = flagged ? c->off4.p : nullptr;
    b.win = c->win.p; b.phit = c->phit.p; b.endmask = c->endmask.p; b.impmask = c->impmask.p; b.orient_hit = c->orient_hit.p;
    b.slot_count = c->slot_count.p; b.ent_base = c->ent_base.p; b.ssum = c->ssum.p;
    b.rec_stage = c->rec_stage.p; b.rec_extra = c->rec_extra.p;
    bind_entry_buffers(c);
    b.rec_count = c->rec_count.p; b.rec_offset = c->rec_offset.p;
    b.records = nullptr; b.read_flags = c->read_flags.p; b.counters = c->counters.p;
    CU(cudaStreamSynchronize(c->stream));
    c->have_batch = true; c->have_results = false;
    return SMX_OK;
}

int smx_run_resident(smx_ctx *c) {
    if (!c) return fail(SMX_ERR_ARG, "smx_run_resident: null context");
    if (!c->have_batch) return fail(SMX_ERR_ARG, "smx_run_resident: no batch uploaded");
    CU(cudaSetDevice(c->device));
    const Tables &t = c->t;
    Batch &b = c->b;
    const u32 n = b.n_reads;
    const int nP = t.n_primers;
    cudaStream_t st = c->stream;
    int launches = 0;
    CU(cudaMemcpyToSymbolAsync(c_tables, &c->t, sizeof(Tables), 0, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), st));
    CU(cudaEventRecord(c->ev[0], st));
    {   // stage 0
        dim3 grid((n + 127) / 128, 2 * t.wpw);
        k_stage_windows<<<grid, 128, 0, st>>>(b);
        ++launches;
    }
    const unsigned blocks = (n + 127) / 128;
    unsigned long long host_counters[8];
    std::vector<u32> slot_counts((size_t)2 * nP);
    int from = 1;                      // stage to (re)start from
    for (;;) {
        if (from <= 1) {   // stage 1
            CU(cudaMemsetAsync(c->slot_count.p, 0, (size_t)2 * nP * sizeof(u32), st));
            CU(cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long), st));
            CU(cudaMemsetAsync(c->counters.p + 2, 0, sizeof(unsigned long long), st));
            CU(cudaEventRecord(c->ev[1], st));
            dim3 grid(blocks, 2 * nP);
            if (t.use64) k_primer_search<u64><<<grid, 128, 0, st>>>(b);
            else k_primer_search<u32><<<grid, 128, 0, st>>>(b);
            dim3 sgrid((b.e_cap + 127) / 128, 2 * nP);
            if (t.use64) k_primer_start<u64><<<sgrid, 128, 0, st>>>(b);
            else k_primer_start<u32><<<sgrid, 128, 0, st>>>(b);
            launches += 2;
        }
        if (from <= 2) {   // stage 2
            CU(cudaMemsetAsync(c->counters.p + 1, 0, sizeof(unsigned long long), st));
            CU(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(unsigned long long), st));
            CU(cudaMemsetAsync(c->counters.p + 7, 0, sizeof(unsigned long long), st));
            CU(cudaEventRecord(c->ev[2], st));
            if (t.n_bwords) {
                dim3 grid((b.e_cap + 127) / 128, 2 * t.n_bwords);
                switch (t.k_idx) {
#define SMX_K2(KK) case KK: k_barcode_bitsliced<KK><<<grid, 128, 0, st>>>(b); break;
                    SMX_K2(0) SMX_K2(1) SMX_K2(2) SMX_K2(3) SMX_K2(4) SMX_K2(5) SMX_K2(6) SMX_K2(7) SMX_K2(8)
#undef SMX_K2
                    default: return fail(SMX_ERR_INTERNAL, "unsupported k_idx");
                }
                ++launches;
            }
        }
        // stage 3: single-pass selection, scan
        CU(cudaMemsetAsync(c->counters.p + 4, 0, 3 * sizeof(unsigned long long), st));
        CU(cudaEventRecord(c->ev[3], st));
        {
            dim3 sgrid((n + 255) / 256, 2 * nP);
            k_slot_summary<<<sgrid, 256, 0, st>>>(b);
        }
        if (nP <= 8) k_select<8><<<blocks, 128, 0, st>>>(b); else k_select<SMX_MAX_PRIMERS><<<blocks, 128, 0, st>>>(b);
        launches += 2;
        CU(cudaMemcpyAsync(slot_counts.data(), c->slot_count.p, slot_counts.size() * sizeof(u32), cudaMemcpyDeviceToHost, st));
        CU(scan_and_count(c, launches, host_counters));
        // capacity checks: every overflow is resolved by a second GPU pass with larger buffers
        u32 max_entries = 0;
        for (u32 v : slot_counts) max_entries = std::max(max_entries, v);
        if (max_entries > c->e_cap) {
            c->e_cap = (max_entries + 127u) & ~127u;
            CU(ensure_entry_buffers(c)); bind_entry_buffers(c);
            from = 1;
            continue;
        }
        if (host_counters[7]) {
            if (c->t.hit_cap >= kMaxWordHits) return fail(SMX_ERR_INTERNAL, "hit list overflow at maximum capacity");
            c->t.hit_cap = std::min(kMaxWordHits, c->t.hit_cap * 4);
            CU(ensure_entry_buffers(c)); bind_entry_buffers(c);
            CU(cudaMemcpyToSymbolAsync(c_tables, &c->t, sizeof(Tables), 0, cudaMemcpyHostToDevice, st));
            from = 2;
            continue;
        }
        u32 pool_used = (u32)(host_counters[6] >> 32);
        if (pool_used > c->pool_cap) {
            c->pool_cap = pool_used + 1024;
            CU(ensure_entry_buffers(c)); bind_entry_buffers(c);
            from = 3;
            continue;
        }
        break;
    }
    std::vector<u32> big_list;
    if ((u32)host_counters[5]) {
        // some reads overflowed the thread-local group storage (or emit many records): second GPU
        // pass for those reads only, on kBigGroups-entry global scratch
        std::vector<unsigned char> flags(n);
        CU(cudaMemcpy(flags.data(), b.read_flags, n, cudaMemcpyDeviceToHost));
        for (u32 r = 0; r < n; ++r) if (flags[r] & 2) big_list.push_back(r);
        CU(c->big_list.ensure(big_list.size()));
        CU(cudaMemcpy(c->big_list.p, big_list.data(), big_list.size() * sizeof(u32), cudaMemcpyHostToDevice));
        const size_t chunk = 512;
        CU(c->big_scratch.ensure(std::min(chunk, big_list.size()) * kBigScratchBytes));
        for (size_t off = 0; off < big_list.size(); off += chunk) {
            u32 cnt = (u32)std::min(chunk, big_list.size() - off);
            k_select_big<<<(cnt + 31) / 32, 32, 0, st>>>(b, c->big_list.p + off, cnt, c->big_scratch.p, 0);
            ++launches;
        }
        CU(cudaMemsetAsync(c->counters.p + 4, 0, 2 * sizeof(unsigned long long), st));
        CU(cudaMemsetAsync(c->counters.p + 6, 0, sizeof(u32), st));
        CU(scan_and_count(c, launches, host_counters));
        if ((u32)(host_counters[5] >> 32))
            return fail(SMX_ERR_INTERNAL, "selection: %u read(s) exceed %d dereplication groups",
                        (unsigned)(host_counters[5] >> 32), kBigGroups);
    }
    {
        u64 total = (u32)host_counters[6];
        CU(c->records.ensure(total + 1));
        b.records = c->records.p;
        k_compact_records<<<(n + 255) / 256, 256, 0, st>>>(b);
        ++launches;
        const size_t chunk = 512;
        for (size_t off = 0; off < big_list.size(); off += chunk) {
            u32 cnt = (u32)std::min(chunk, big_list.size() - off);
            k_select_big<<<(cnt + 31) / 32, 32, 0, st>>>(b, c->big_list.p + off, cnt, c->big_scratch.p, 1);
            ++launches;
        }
        c->n_records = total;
        c->n_matched = host_counters[4];
        for (int i = 0; i < 4; ++i) c->work[i] = host_counters[i];
    }
    CU(cudaEventRecord(c->ev[4], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]));
    CU(cudaEventElapsedTime(&c->total_ms, c->ev[0], c->ev[4]));
    c->launches = launches;
    c->have_results = true;
    return SMX_OK;
}

int smx_download_results(smx_ctx *c, smx_results *out) {
    if (!c || !out) return fail(SMX_ERR_ARG, "smx_download_results: null argument");
    if (!c->have_results) return fail(SMX_ERR_ARG, "smx_download_results: nothing to download");
    CU(cudaSetDevice(c->device));
    const Batch &b = c->b;
    const Tables &t = c->t;
    const u32 n = b.n_reads;
    out->n_records = c->n_records;
    out->n_matched = c->n_matched;
    if (c->n_records > out->records_cap)
        return fail(SMX_ERR_CAPACITY, "smx_download_results: %llu records, capacity %llu",
                    (unsigned long long)c->n_records, (unsigned long long)out->records_cap);
    cudaStream_t st = c->stream;
    if (out->rec_offset)
        CU(cudaMemcpyAsync(out->rec_offset, b.rec_offset, ((size_t)n + 1) * sizeof(u32), cudaMemcpyDeviceToHost, st));
    if (out->records && c->n_records)
        CU(cudaMemcpyAsync(out->records, b.records, c->n_records * sizeof(smx_record), cudaMemcpyDeviceToHost, st));
    // level-1 detail is stored padded ([slot][n_pad]) on the device and returned dense ([slot][n])
    if (out->primer_hits)
        CU(cudaMemcpy2DAsync(out->primer_hits, (size_t)n * sizeof(smx_primer_hit), b.phit,
                             (size_t)b.n_pad * sizeof(smx_primer_hit), (size_t)n * sizeof(smx_primer_hit),
                             (size_t)2 * t.n_primers, cudaMemcpyDeviceToHost, st));
    if (out->endmask_bits)
        CU(cudaMemcpy2DAsync(out->endmask_bits, (size_t)n * sizeof(u32), b.endmask, (size_t)b.n_pad * sizeof(u32),
                             (size_t)n * sizeof(u32), (size_t)2 * t.n_primers * t.mw, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (out->barcode_hits && t.total_bslots) {
        // expand the compact per-(entry, bword) hit lists into the dense [barcode slot][read] detail
        // layout, merging a barcode's hits over the primer end locations (strictly smaller wins)
        std::vector<smx_primer_hit> ph((size_t)2 * t.n_primers * b.n_pad);
        std::vector<u32> ebase((size_t)2 * t.n_primers * b.n_pad);
        std::vector<unsigned char> cnt((size_t)2 * t.n_bwords * b.e_cap);
        std::vector<smx_barcode_hit> lst((size_t)2 * t.n_bwords * t.hit_cap * b.e_cap);
        std::vector<unsigned char> bwp(t.n_bwords);
        CU(cudaMemcpy(ph.data(), b.phit, ph.size() * sizeof(smx_primer_hit), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(ebase.data(), b.ent_base, ebase.size() * sizeof(u32), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(cnt.data(), b.bh_count, cnt.size(), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(lst.data(), b.bh_list, lst.size() * sizeof(smx_barcode_hit), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(bwp.data(), t.bw_primer, bwp.size(), cudaMemcpyDeviceToHost));
        smx_barcode_hit none;
        none.end_mask = 0; none.search_start = 0; none.distance = -1; none.barcode = 0;
        for (size_t i = 0; i < (size_t)t.total_bslots * n; ++i) out->barcode_hits[i] = none;
        for (int sd = 0; sd < 2; ++sd)
            for (int g = 0; g < t.n_bwords; ++g) {
                int p = bwp[g];
                u32 slot = (u32)(sd * t.n_primers + p);
                u64 gslot = (u64)sd * t.n_bwords + g;
                for (u32 r = 0; r < n; ++r) {
                    const smx_primer_hit &h0 = ph[(size_t)slot * b.n_pad + r];
                    if (h0.distance < 0) continue;
                    for (u32 l = 0; l < h0.n_locations; ++l) {
                        u64 e = (u64)ebase[(size_t)slot * b.n_pad + r] + l;
                        int k = std::min<int>(cnt[gslot * b.e_cap + e], t.hit_cap);
                        for (int x = 0; x < k; ++x) {
                            const smx_barcode_hit &h = lst[(gslot * t.hit_cap + x) * b.e_cap + e];
                            smx_barcode_hit &dst = out->barcode_hits[((size_t)t.bslot_base[slot] + h.barcode) * n + r];
                            if (dst.distance < 0 || h.distance < dst.distance) dst = h;
                        }
                    }
                }
            }
    }
    return SMX_OK;
}

int smx_match_batch(smx_ctx *c, const smx_batch *in, smx_results *out) {
    int rc = smx_upload_batch(c, in);
    if (rc) return rc;
    rc = smx_run_resident(c);
    if (rc) return rc;
    return smx_download_results(c, out);
}

int smx_last_timing(const smx_ctx *c, float *total_ms, float stage_ms[4]) {
    if (!c) return fail(SMX_ERR_ARG, "smx_last_timing: null context");
    if (total_ms) *total_ms = c->total_ms;
    if (stage_ms) for (int i = 0; i < 4; ++i) stage_ms[i] = c->stage_ms[i];
    return SMX_OK;
}

int smx_last_launch_count(const smx_ctx *c) { return c ? c->launches : 0; }

int smx_last_work(const smx_ctx *c, uint64_t cells[2], uint64_t wordcols[2]) {
    if (!c || !c->have_results) return fail(SMX_ERR_ARG, "smx_last_work: no results");
    cells[0] = c->work[0]; cells[1] = c->work[1];
    wordcols[0] = c->work[2]; wordcols[1] = c->work[3];
    return SMX_OK;
}

int smx_pairwise_nw(int device, const char *seqs, const uint32_t *seq_off, uint32_t n, int32_t *out) {
    if (!seqs || !seq_off || !out) return fail(SMX_ERR_ARG, "smx_pairwise_nw: null argument");
    if (n == 0) return SMX_OK;
    for (u32 i = 0; i < n; ++i)
        if (seq_off[i + 1] - seq_off[i] > 64) return fail(SMX_ERR_ARG, "smx_pairwise_nw: sequence %u longer than 64", i);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SMX_ERR_NO_DEVICE, "smx_pairwise_nw: no CUDA device; this library has no CPU path");
    }
    CU(cudaSetDevice(device));
    char *d_seq = nullptr; u32 *d_off = nullptr; i32 *d_out = nullptr;
    size_t bytes = seq_off[n];
    CU(cudaMalloc((void **)&d_seq, bytes ? bytes : 1));
    CU(cudaMalloc((void **)&d_off, (size_t)(n + 1) * sizeof(u32)));
    CU(cudaMalloc((void **)&d_out, (size_t)n * n * sizeof(i32)));
    CU(cudaMemcpy(d_seq, seqs, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_off, seq_off, (size_t)(n + 1) * sizeof(u32), cudaMemcpyHostToDevice));
    u64 total = (u64)n * n;
    k_pairwise_nw<<<(unsigned)((total + 127) / 128), 128>>>(d_seq, d_off, n, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d_out, total * sizeof(i32), cudaMemcpyDeviceToHost));
    cudaFree(d_seq); cudaFree(d_off); cudaFree(d_out);
    return SMX_OK;
}

int smx_int_alu_peak(int device, double out_tops[3]) {
    if (!out_tops) return fail(SMX_ERR_ARG, "smx_int_alu_peak: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SMX_ERR_NO_DEVICE, "smx_int_alu_peak: no CUDA device");
    }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    u32 *d_out = nullptr;
    CU(cudaMalloc((void **)&d_out, 4));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    const double ops = (double)blocks * threads * iters * 64.0;
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CU(cudaEventRecord(e0));
            if (mode == 0) k_int_peak<0><<<blocks, threads>>>(d_out, iters, 17u + rep);
            else if (mode == 1) k_int_peak<1><<<blocks, threads>>>(d_out, iters, 17u + rep);
            else k_int_peak<2><<<blocks, threads>>>(d_out, iters, 17u + rep);
            CU(cudaEventRecord(e1));
            CU(cudaEventSynchronize(e1));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        out_tops[mode] = ops / (best * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    return SMX_OK;
}

int smx_flush_l2(smx_ctx *c) {
    if (!c) return fail(SMX_ERR_ARG, "smx_flush_l2: null context");
    CU(cudaSetDevice(c->device));
    const size_t bytes = 512ull << 20;
    CU(c->big_scratch.ensure(bytes));
    CU(cudaMemsetAsync(c->big_scratch.p, 0x5a, bytes, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SMX_OK;
}

void *smx_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        fail(SMX_ERR_CUDA, "smx_host_alloc: cudaHostAlloc(%llu) failed", (unsigned long long)bytes);
        return nullptr;
    }
    return p;
}

void smx_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------
// Host-side packer (batching layer; no matching happens here).

void smx_pack_bound(const uint64_t *seq_off, uint32_t n_reads, uint32_t clip_len,
                    uint64_t *packed2_words, uint64_t *packed4_words_max) {
    uint64_t w2 = 0, w4 = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        uint64_t len = seq_off[r + 1] - seq_off[r];
        if (clip_len && len > 2ull * clip_len) len = 2ull * clip_len;
        w2 += (len + 15) / 16;
        w4 += 2 * ((len + 7) / 8);
    }
    if (packed2_words) *packed2_words = w2 + 1;
    if (packed4_words_max) *packed4_words_max = w4 + 1;
}

int smx_pack_reads(const char *bases, const uint64_t *seq_off, uint32_t n_reads, uint32_t clip_len,
                   uint32_t *packed2, uint64_t *word_off, uint32_t *lengths,
                   uint32_t *packed4, uint64_t *off4, uint64_t *packed4_words, uint32_t *n_flagged) {
    if (!bases || !seq_off || !packed2 || !word_off || !lengths)
        return fail(SMX_ERR_ARG, "smx_pack_reads: null argument");
    uint64_t w2 = 0, w4 = 0;
    uint32_t flagged = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        const unsigned char *s = (const unsigned char *)bases + seq_off[r];
        const uint64_t len = seq_off[r + 1] - seq_off[r];
        if (len > 0x7FFFFFFFull) return fail(SMX_ERR_ARG, "smx_pack_reads: read %u too long", r);
        const bool clipped = clip_len && len > 2ull * clip_len;
        const uint64_t slen = clipped ? 2ull * clip_len : len;     // stored bases
        const uint64_t skip = len - slen;                          // bases dropped from the middle
        // stored index i -> source index
        auto src = [&](uint64_t i) { return (clipped && i >= clip_len) ? i + skip : i; };
        word_off[r] = w2;
        lengths[r] = (uint32_t)len;
        bool exotic = false;
        const uint64_t nw = (slen + 15) / 16;
        for (uint64_t w = 0; w < nw; ++w) {
            uint32_t v = 0;
            uint64_t lim = std::min<uint64_t>(16, slen - w * 16);
            for (uint64_t i = 0; i < lim; ++i) {
                int c = read_code(s[src(w * 16 + i)]);
                if (c > 3) { exotic = true; c = 0; }
                v |= (uint32_t)c << (2 * i);
            }
            packed2[w2 + w] = v;
        }
        w2 += nw;
        if (off4) off4[r] = ~0ull;
        if (exotic) {
            if (!packed4 || !off4) return fail(SMX_ERR_ARG, "smx_pack_reads: read %u needs the packed4 stream", r);
            off4[r] = w4;
            const uint64_t n4 = (slen + 7) / 8;
            for (uint64_t w = 0; w < 2 * n4; ++w) packed4[w4 + w] = 0;
            for (uint64_t i = 0; i < slen; ++i) {
                unsigned char ch = s[src(i)];
                packed4[w4 + i / 8] |= (uint32_t)read_code(ch) << (4 * (i % 8));
                // the reverse-complement strand is clipped the same way: its stored index of the
                // base at source position p is the stored index of (len-1-p) mirrored
                uint64_t x = slen - 1 - i;
                packed4[w4 + n4 + x / 8] |= (uint32_t)read_code_rc(ch) << (4 * (x % 8));
            }
            w4 += 2 * n4;
            ++flagged;
        }
    }
    if (packed4_words) *packed4_words = w4;
    if (n_flagged) *n_flagged = flagged;
    return SMX_OK;
}

}  // extern "C"
