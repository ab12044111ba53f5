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

SMX_HD void stage_window_word(const Tables &t, const Batch &b, u32 read, int strand, int w) {
    int n = (int)b.lengths[read];
    Geo g = make_geo(n, t.L);
    u32 out = 0;
    for (int i = 0; i < 8; ++i) {
        int p = w * 8 + i;
        int c = kSymOther;
        if (p < g.wl) c = sym_at(b, read, strand, g.woff + p, n);
        out |= (u32)c << (4 * i);
    }
    b.win[((u64)strand * t.wpw + w) * b.n_pad + read] = out;
}

SMX_HD int staged_sym(const Tables &t, const Batch &b, u32 read, int strand, int p) {
    u32 w = b.win[((u64)strand * t.wpw + (p >> 3)) * b.n_pad + read];
    return (int)((w >> (4 * (p & 7))) & 15);
}

// ---------------------------------------------------------------------------------------------
// Stage 1.  match_one_end's primer search (demultiplex.py:757-766) for one (read, strand, primer).

template <typename W>
SMX_HD void primer_search_thread(const Tables &t, const Batch &b, u32 read, int strand, int primer,
                                 const u64 *peq, const u64 *peq_rev, const u64 *peq_fw) {
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    const int m = t.p_len[primer], k = t.p_k[primer];
    const u32 slot = slot_index(t, strand, primer);
    const u64 hit_idx = (u64)slot * b.n_pad + read;
    const u32 *win = b.win + (u64)strand * t.wpw * b.n_pad + read;
    u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + read;

    W Pv = ~(W)0, Mv = 0;
    int score = m, best = m + 1, first = 0;
    const int p_begin = g.start, p_end = g.wl;
    for (int mwi = 0; mwi < t.mw; ++mwi) {
        u32 mask = 0;
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
                    int p = w * 8 + i;
                    if (score < best) { best = score; first = p; }
                    if (score == best) mask |= 1u << (q * 8 + i);
                }
            } else {
                for (int i = 0; i < 8; ++i) {
                    int p = w * 8 + i;
                    int c = (int)(word & 15);
                    word >>= 4;
                    if (p < p_begin || p >= p_end) continue;
                    W Eq = peq_word<W>(peq[c]);
                    score += myers_step<W, false>(Eq, Pv, Mv);
                    if (score < best) { best = score; first = p; }
                    if (score == best) mask |= 1u << (q * 8 + i);
                }
            }
        }
        emask[(u64)mwi * b.n_pad] = mask;
    }

    smx_primer_hit h;
    h.distance = -1; h.n_locations = 0; h.first_start = 0; h.first_end = 0;
    if (best <= k) {
        // bits set before `first` belong to an older (larger) best: clear them, count the rest
        int nloc = 0;
        for (int mwi = 0; mwi < t.mw; ++mwi) {
            u32 v = emask[(u64)mwi * b.n_pad];
            int lo = mwi * 32;
            if (lo + 32 <= first) v = 0;
            else if (lo < first) v &= ~0u << (first - lo);
            emask[(u64)mwi * b.n_pad] = v;
#if defined(__CUDA_ARCH__)
            nloc += __popc(v);
#else
            nloc += __builtin_popcount(v);
#endif
        }
        auto load = [&](int p) { return staged_sym(t, b, read, strand, p); };
        int back = hw_start_back<W>(peq_rev, m, best, first, g.start, load);
        h.distance = (int16_t)best;
        h.n_locations = (uint16_t)nloc;
        h.first_end = g.woff + first + g.delta;
        h.first_start = g.woff + first - back + g.delta;
    }
    b.phit[hit_idx] = h;

    // determine_orientation (demultiplex.py:602-638), explicit form, only where the tail-window
    // equivalence does not hold: reads shorter than search_len-1 (Q1) or with non-ACGT symbols.
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
    b.orient_hit[hit_idx] = ohit;
}

// ---------------------------------------------------------------------------------------------
// Stage 2.  match_one_end's barcode loop (demultiplex.py:781-815) for one (read, strand, primer,
// barcode j): SHW search at every equal-best primer end, strictly-smallest distance wins.
// Returns the SHW cells / word-columns this thread accounts for (SURVEY.md 8d formula).

template <typename W>
SMX_HD void barcode_search_thread(const Tables &t, const Batch &b, u32 read, int strand, int primer, int j,
                                  const u64 *peq, int m, unsigned long long &cells, unsigned long long &wcols) {
    const u32 slot = slot_index(t, strand, primer);
    const smx_primer_hit ph = b.phit[(u64)slot * b.n_pad + read];
    if (ph.distance < 0) return;
    const int n = (int)b.lengths[read];
    const Geo g = make_geo(n, t.L);
    const u32 *emask = b.endmask + (u64)slot * t.mw * b.n_pad + read;
    const int k = t.k_idx;

    smx_barcode_hit out;
    out.distance = -1; out.end_mask = 0; out.search_start = 0; out.pad = 0;
    for (int mwi = 0; mwi < t.mw; ++mwi) {
        u32 word = emask[(u64)mwi * b.n_pad];
        while (word) {
#if defined(__CUDA_ARCH__)
            int bit = __ffs((int)word) - 1;
#else
            int bit = __builtin_ctz(word);
#endif
            word &= word - 1;
            int p = mwi * 32 + bit;
            int e_rep = g.woff + p + g.delta;
            Flank f = make_flank(e_rep, n);
            int fl = n - f.a_align;
            int cols = fl < m + k ? fl : m + k;
            cells += (unsigned long long)m * (unsigned long long)cols;
            wcols += (unsigned long long)((m + 31) >> 5) * (unsigned long long)cols;
            if (t.prefilter) {
                // BloomPrefilter.match (bloom_filter.py:176-186) with an exact set: the key
                // barcode_rc + flank[:m-k] can only be present if those m-k symbols exist and are
                // all A/C/G/T; for such flanks the filter has no false negatives.
                int need = m - k;
                if (n - f.a_pref < need) continue;
                bool acgt = true;
                for (int x = 0; x < need; ++x)
                    if (staged_sym(t, b, read, strand, f.a_pref - g.woff + x) > 3) { acgt = false; break; }
                if (!acgt) continue;
            }
            if (cols <= 0) continue;                 // empty target: edlib reports distance m > k
            int base = f.a_align - g.woff;
            auto load = [&](int x) { return staged_sym(t, b, read, strand, base + x); };
            int best; u64 mask;
            shw_search<W>(peq, m, cols, load, best, mask);
            if (best <= k && (out.distance < 0 || best < out.distance)) {
                out.distance = (int16_t)best; out.end_mask = mask; out.search_start = f.bs;
            }
        }
    }
    u64 bslot = (u64)t.bslot_base[slot] + (u64)j;
    b.bhit[bslot * b.n_pad + read] = out;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// __global__ wrappers.  Tables live in __constant__ memory, Peq masks are staged once per block
// in shared memory.

__constant__ Tables c_tables;

__global__ void k_stage_windows(Batch b) {
    // grid: x over reads, y over (strand, word)
    u32 read = blockIdx.x * blockDim.x + threadIdx.x;
    if (read >= b.n_reads) return;
    int strand = blockIdx.y / c_tables.wpw, w = blockIdx.y % c_tables.wpw;
    stage_window_word(c_tables, b, read, strand, w);
}

template <typename W>
__global__ void __launch_bounds__(128) k_primer_search(Batch b) {
    // grid: x over reads, y = strand * n_primers + primer
    __shared__ u64 s_peq[3][16];
    const int primer = blockIdx.y % c_tables.n_primers, strand = blockIdx.y / c_tables.n_primers;
    if (threadIdx.x < 48) {
        const u64 *src = threadIdx.x < 16 ? c_tables.peq_rc : threadIdx.x < 32 ? c_tables.peq_rcrev : c_tables.peq_fw;
        s_peq[threadIdx.x >> 4][threadIdx.x & 15] = src[primer * 16 + (threadIdx.x & 15)];
    }
    __syncthreads();
    u32 read = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0;
    if (read < b.n_reads) {
        primer_search_thread<W>(c_tables, b, read, strand, primer, s_peq[0], s_peq[1], s_peq[2]);
        int n = (int)b.lengths[read];
        cells = (unsigned long long)(n < c_tables.L ? n : c_tables.L);       // HW columns of this search
    }
    for (int o = 16; o; o >>= 1) cells += __shfl_down_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) {
        int m = c_tables.p_len[primer];
        atomicAdd(&b.counters[0], cells * (unsigned long long)m);
        atomicAdd(&b.counters[2], cells * (unsigned long long)((m + 31) >> 5));
    }
}

// One warp per (read, group of 32 barcodes); lanes = barcodes.  Shared Peq layout [barcode][16].
template <typename W>
__global__ void __launch_bounds__(256) k_barcode_search(Batch b, int primer, int strand) {
    extern __shared__ u64 s_bpeq[];            // nb * 16
    const Tables &t = c_tables;
    const int nb = (int)(t.pb_off[primer + 1] - t.pb_off[primer]);
    for (int i = threadIdx.x; i < nb * 16; i += blockDim.x) s_bpeq[i] = t.bpeq[(u64)t.pb_off[primer] * 16 + i];
    __syncthreads();
    const int groups = (nb + 31) >> 5;
    const int lane = threadIdx.x & 31;
    u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u32 read = (u32)(warp / groups);
    int j = (int)(warp % groups) * 32 + lane;
    unsigned long long cells = 0, wcols = 0;
    if (read < b.n_reads && j < nb)
        barcode_search_thread<W>(t, b, read, strand, primer, j, s_bpeq + j * 16, t.b_len[t.pb_off[primer] + j],
                                 cells, wcols);
    for (int o = 16; o; o >>= 1) {
        cells += __shfl_down_sync(0xffffffffu, cells, o);
        wcols += __shfl_down_sync(0xffffffffu, wcols, o);
    }
    if (lane == 0 && cells) {
        atomicAdd(&b.counters[1], cells);
        atomicAdd(&b.counters[3], wcols);
    }
}

template <int MAXP>
__global__ void __launch_bounds__(128) k_select(Batch b, int write_pass) {
    u32 read = blockIdx.x * blockDim.x + threadIdx.x;
    if (read >= b.n_reads) return;
    EndInfo ends[2 * MAXP];
    SelectCtx c; c.t = &c_tables; c.b = &b; c.read = read; c.n = (int)b.lengths[read];
    unsigned char flags;
    smx_record *out = write_pass ? b.records + b.rec_offset[read] : nullptr;
    u32 cnt = select_read(c, ends, out, flags);
    if (!write_pass) { b.rec_count[read] = cnt; b.read_flags[read] = flags; }
}

// Exclusive scan of rec_count -> rec_offset (n+1 entries), three small kernels.
constexpr int kScanBlock = 1024;

__global__ void k_scan_block_sums(const u32 *in, u32 n, u32 *block_sums) {
    __shared__ u32 s[32];
    u32 i = blockIdx.x * kScanBlock + threadIdx.x;
    u32 v = i < n ? in[i] : 0;
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = s[threadIdx.x];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

__global__ void k_scan_spine(u32 *block_sums, u32 nblocks, u32 *total) {
    // single block; serial over chunks of 1024 with a running carry
    __shared__ u32 s[kScanBlock];
    __shared__ u32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < nblocks; base += kScanBlock) {
        u32 i = base + threadIdx.x;
        u32 v = i < nblocks ? block_sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kScanBlock; o <<= 1) {
            u32 x = threadIdx.x >= (u32)o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += x;
            __syncthreads();
        }
        if (i < nblocks) block_sums[i] = carry + s[threadIdx.x] - v;     // exclusive
        __syncthreads();
        if (threadIdx.x == 0) carry += s[kScanBlock - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_scan_apply(const u32 *in, u32 n, const u32 *block_sums, u32 *out) {
    __shared__ u32 s[kScanBlock];
    u32 i = blockIdx.x * kScanBlock + threadIdx.x;
    u32 v = i < n ? in[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
        u32 x = threadIdx.x >= (u32)o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += x;
        __syncthreads();
    }
    if (i < n) out[i] = block_sums[blockIdx.x] + s[threadIdx.x] - v;
    if (i == n - 1) out[n] = block_sums[blockIdx.x] + s[threadIdx.x];
}

__global__ void k_count_flags(const unsigned char *flags, u32 n, unsigned long long *matched, unsigned int *overflow) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned f = i < n ? flags[i] : 0;
    unsigned m = __ballot_sync(0xffffffffu, f & 1);
    unsigned o = __ballot_sync(0xffffffffu, f & 2);
    if ((threadIdx.x & 31) == 0) {
        if (m) atomicAdd(matched, (unsigned long long)__popc(m));
        if (o) atomicAdd(overflow, (unsigned)__popc(o));
    }
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
