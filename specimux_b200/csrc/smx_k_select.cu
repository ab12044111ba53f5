// smx_k_select.cu -- stage 3 (selection, scan + compaction, record packing) and the small utility kernels.
#include <cstdlib>
#include <cuda_runtime.h>

#include "smx_device.cuh"
#include "smx_launch.hpp"

namespace smx {

constexpr int kInlineRecords = 4;

// Stage 3a, fast selection: one thread per read, only the common single-candidate path of the
// selection routine (slot digests computed on the fly, no grouping, small register footprint so
// the dependent global loads are hidden by occupancy).  Reads that need the general routine
// (several equal-best candidates, tied barcodes, TAILS trimming) are appended to defer_list.
// 16 resident blocks asked for = at most 32 registers: the kernel lives on warps in flight (measured on config 2:
// 52 registers 89 us, 38 registers 79 us, 32 registers 76 us; profiles/r2_u_ab.md).
template <int MAXP>
__global__ void __launch_bounds__(128, 16) k_select_fast(SMX_KARGS) {
    // (loading all slots ahead -- select_read_impl<true, NP> -- was measured SLOWER here: 93 vs 80 us, 62 vs 38
    // registers; the kernel lives on occupancy, profiles/r2_l_ab.md)
    constexpr int NP = 0;
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
        u32 cnt = select_read_impl<true, NP>(c, ends, st, &rec, 1, flags);
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
    // so that a warp serialises 4 divergent reads instead of 32.  The grid is a fixed few blocks per SM walking the
    // list in strides (a grid over all reads spent a fifth of the kernel's instructions on threads that only looked
    // at the count and left, profiles/r2_j).
    constexpr int NP = MAXP <= 4 ? MAXP : 0;
    const u32 n_def = (u32)b.counters[kCtrDeferred];
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
    // lanes per read: a whole warp when the list is short enough (divergent reads of one warp run one after the other,
    // and a read is ~3 k dependent instructions: four to a warp was a 12 k-instruction chain, 50 us), else 8, else 1
    const u32 shift = (u64)n_def * 32 <= (u64)total ? 5u : (u64)n_def * 8 <= (u64)total ? 3u : 0u;
    if (tid & ((1u << shift) - 1u)) return;
    const u32 first = tid >> shift, step = total >> shift;
    for (u32 i = first; i < n_def; i += step) {
        const u32 read = b.defer_list[i];
        EndInfo ends[2 * MAXP];
        Group groups[kSmallGroups], pg[kSmallGroups];
        Cand gcand[kSmallGroups], pcand[kSmallGroups];
        int ts_cand[kSmallGroups], ts_shift[kSmallGroups];
        BestList best[2 * MAXP];
        smx_record local[kInlineRecords];
        SelectStore st;
        st.groups = groups; st.gcand = gcand; st.pg = pg; st.pcand = pcand;
        st.ts_cand = ts_cand; st.ts_shift = ts_shift; st.cap = kSmallGroups;
        SelectCtx c; c.t = &c_tables; c.b = &b; c.read = read; c.n = (int)b.lengths[read];
        for (int q = 0; q < 2 * c_tables.n_primers; ++q) best[q].n = -1;
        c.best = best;
        unsigned char flags;
        u32 cnt = select_read<NP>(c, ends, st, local, kInlineRecords, flags);
        if (cnt > kInlineRecords) flags |= 2;
        b.rec_count[read] = cnt;
        b.read_flags[read] = flags;
        if (flags & 2) {                                    // second pass: listed on the device, no host round trip
            b.big_list[atomicAdd((unsigned int *)&b.counters[kCtrBig], 1u)] = read;
            continue;
        }
        if (cnt >= 1) b.rec_stage[read] = local[0];
        if (cnt >= 2) {
            u32 base = atomicAdd((unsigned int *)&b.counters[6] + 1, cnt - 1);
            b.rec_extra[read] = base;
            if (base + cnt - 1 <= b.pool_cap)
                for (u32 i2 = 1; i2 < cnt; ++i2) b.rec_pool[base + i2 - 1] = local[i2];
        }
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
    if ((u64)base + cnt > b.pool_cap) {                 // pool overflow: the library grows it and re-runs selection
        b.rec_extra[read] = 0xFFFFFFFFu;                // keeps k_scan_compact away from a stale pool index
        return;
    }
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
constexpr int kScanThreads = 256;

__global__ void __launch_bounds__(kScanThreads, 6) k_scan_compact(SMX_KARGS, u32 rec_cap, unsigned long long *tile_status,
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
        // Decoupled look-back, 32 predecessors per step: lane l looks at tile (j - 1 - l), waits until that tile has
        // published at least its aggregate, and the warp sums the aggregates up to the nearest tile that already holds
        // an inclusive prefix (tiles before tile 0 count as "prefix 0").  One lane walking back tile by tile was a chain
        // of dependent loads as long as the distance to the nearest finished tile -- with all 748 tiles of config 2
        // resident at once, 39 % of the kernel's stall samples sat behind it (profiles/r2_m ncu source view).
        const unsigned long long tag = (unsigned long long)epoch << 34;
        volatile unsigned long long *st = tile_status;
        u32 excl = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = tag | (2ull << 32) | total;
        } else {
            if (lane == 0) { st[tile] = tag | (1ull << 32) | total; __threadfence(); }
            __syncwarp();
            for (u32 j = tile;;) {
                const bool real = j >= 1u + (u32)lane;               // tile j - 1 - lane exists
                unsigned long long v = (2ull << 32);                 // before the first tile: inclusive prefix 0
                if (real) {
                    const u32 idx = j - 1u - (u32)lane;
                    do { v = st[idx]; } while ((v >> 34) != epoch || ((v >> 32) & 3ull) == 0);
                }
                const bool is_prefix = ((v >> 32) & 3ull) == 2ull;
                const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                const int nearest = pm ? __ffs((int)pm) - 1 : 31;    // lanes 0..nearest contribute
                excl += __reduce_add_sync(0xffffffffu, lane <= nearest ? (u32)v : 0u);
                if (pm) break;
                j -= 32u;                                            // no prefix among 32 real tiles: j > 32 here
            }
            if (lane == 0) st[tile] = tag | (2ull << 32) | (u32)(excl + total);
        }
        if (lane == 0) {
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
                // further records sit in the pool -- unless the pool overflowed (the host then grows it and
                // re-runs selection): never read past it
                const u32 extra = b.rec_extra[read];
                if ((u64)extra + cnt[k] - 1 <= b.pool_cap) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(b.rec_pool + extra);
                    for (u32 i = q; i < 4 * (cnt[k] - 1); i += 4) dst[4 + i] = src[i];
                }
            }
        }
    }
}

// smx_record -> smx_record32 (drops the four location pairs) ahead of the copy-out.
__global__ void __launch_bounds__(256) k_pack_records32(const smx_record *in, u32 n, smx_record32 *out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    smx_record32 o;
    pack_record32(in[i], o);
    out[i] = o;
}

// smx_record -> smx_record16 (the 16-byte wire form).  read_base: index of the (sub-)batch's first read in the
// caller's batch (smx_record.read is batch-wide); bad: counts records whose extents do not fit.
__global__ void __launch_bounds__(256) k_pack_records16(const smx_record *in, u32 n, const u32 *lengths, u32 read_base,
                                                        smx_record16 *out, u32 *bad) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const smx_record r = in[i];
    const bool last = i + 1 >= n || in[i + 1].read != r.read;
    smx_record16 o;
    if (!pack_record16(r, (int)lengths[r.read - read_base], last, o)) atomicAdd(bad, 1u);
    out[i] = o;
}

// 16-bit lengths of the wire form -> the 32-bit array every kernel reads.
__global__ void __launch_bounds__(256) k_expand_lengths(const unsigned short *in, u32 n, u32 *out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
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

// SMs of the current device (cached per device ordinal).
static int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

cudaError_t launch_select_fast(const Tables &t, const Batch &b, cudaStream_t st) {
    const unsigned blocks = (b.n_reads + 127) / 128;
    if (t.n_primers <= 8) k_select_fast<8><<<blocks, 128, 0, st>>>(t, b);
    else if (t.n_primers <= 64) k_select_fast<64><<<blocks, 128, 0, st>>>(t, b);
    else k_select_fast<SMX_MAX_PRIMERS><<<blocks, 128, 0, st>>>(t, b);
    return cudaGetLastError();
}

cudaError_t launch_select(const Tables &t, const Batch &b, cudaStream_t st) {
    // a fixed grid (a few blocks per SM) strides over the deferred list, whose length only the device knows
    unsigned blocks = (b.n_reads + 127) / 128;
    const unsigned cap = 6u * (unsigned)sm_count();
    if (blocks > cap) blocks = cap;
    if (t.n_primers == 2) k_select<2><<<blocks, 128, 0, st>>>(t, b);
    else if (t.n_primers <= 8) k_select<8><<<blocks, 128, 0, st>>>(t, b);
    else if (t.n_primers <= 64) k_select<64><<<blocks, 128, 0, st>>>(t, b);
    else k_select<SMX_MAX_PRIMERS><<<blocks, 128, 0, st>>>(t, b);
    return cudaGetLastError();
}

cudaError_t launch_select_big(const Tables &t, const Batch &b, const u32 *list, u32 n_list, unsigned char *scratch, cudaStream_t st) {
    k_select_big<<<(n_list + 31) / 32, 32, 0, st>>>(t, b, list, n_list, scratch);
    return cudaGetLastError();
}

cudaError_t launch_scan_compact(const Tables &t, const Batch &b, u32 rec_cap, unsigned long long *tile_status, u32 *ticket,
                                u32 ticket_base, u32 epoch, cudaStream_t st) {
    const unsigned tiles = (b.n_reads + kScanTile - 1) / kScanTile;
    k_scan_compact<<<tiles, kScanThreads, 0, st>>>(t, b, rec_cap, tile_status, ticket, ticket_base, epoch);
    return cudaGetLastError();
}

cudaError_t launch_pack_records32(const smx_record *in, u32 n, smx_record32 *out, cudaStream_t st) {
    k_pack_records32<<<(n + 255) / 256, 256, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t launch_pack_records16(const smx_record *in, u32 n, const u32 *lengths, u32 read_base, smx_record16 *out, u32 *bad,
                                  cudaStream_t st) {
    k_pack_records16<<<(n + 255) / 256, 256, 0, st>>>(in, n, lengths, read_base, out, bad);
    return cudaGetLastError();
}

cudaError_t launch_expand_lengths(const unsigned short *in, u32 n, u32 *out, cudaStream_t st) {
    k_expand_lengths<<<(n + 255) / 256, 256, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t launch_rebase_offsets(const u32 *in, u32 n, u32 rec_base, u32 *out, cudaStream_t st) {
    k_rebase_offsets<<<(n + 255) / 256, 256, 0, st>>>(in, n, rec_base, out);
    return cudaGetLastError();
}

cudaError_t launch_pairwise_nw(const char *seqs, const u32 *off, u32 n, i32 *out, cudaStream_t st) {
    const u64 total = (u64)n * n;
    k_pairwise_nw<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(seqs, off, n, out);
    return cudaGetLastError();
}

cudaError_t launch_int_peak(int mode, int blocks, int threads, u32 *out, int iters, u32 seed, cudaStream_t st) {
    if (mode == 0) k_int_peak<0><<<blocks, threads, 0, st>>>(out, iters, seed);
    else if (mode == 1) k_int_peak<1><<<blocks, threads, 0, st>>>(out, iters, seed);
    else k_int_peak<2><<<blocks, threads, 0, st>>>(out, iters, seed);
    return cudaGetLastError();
}

}  // namespace smx
