// smx_k_stage1.cu -- stage 0 / 1 kernels other than the sliced forward pass: window staging, primer finish /
// classic search, long primers, start recovery.  Per-thread logic lives in smx_kernels.cuh.
#include <cstdlib>
#include <cuda_runtime.h>

#include "smx_device.cuh"
#include "smx_launch.hpp"

namespace smx {

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
// 8 resident blocks = 32 registers (was 40 / 6 blocks): 47 -> 42 us on config 2 (profiles/r2_u_ab.md)
#ifndef SMX_STAGE_MINB
#define SMX_STAGE_MINB 8
#endif
__global__ void __launch_bounds__(2 * kStageBlock, SMX_STAGE_MINB) k_stage_windows(SMX_KARGS, u32 tile_words) {
    extern __shared__ u32 s_src[];
    const Tables &t = c_tables;
    const u32 r0 = blockIdx.x * kStageBlock;
    const u32 r1 = r0 + kStageBlock < b.n_reads ? r0 + kStageBlock : b.n_reads;
    const u64 w0 = read_word0(b, r0);
    const u64 w1 = read_word0(b, r1 - 1) + (u64)((stored_len(b, (int)b.lengths[r1 - 1]) + 15) >> 4);
    const bool tiled = w1 - w0 + 2 <= tile_words;                             // window extraction reads one word past the read
    if (tiled) {
        const u32 span = (u32)(w1 - w0) + 2;
        const u32 *g = b.packed2 + w0;
        for (u32 i = threadIdx.y * kStageBlock + threadIdx.x; i < span; i += 2 * kStageBlock) s_src[i] = g[i];
    }
    __syncthreads();
    const u32 read = r0 + threadIdx.x;
    if (read >= b.n_reads) return;
    const u32 *src2 = tiled ? s_src : b.packed2;
    const u64 origin = tiled ? w0 : 0;
    const int strand = (int)threadIdx.y;
    stage_windows_thread(t, b, read, strand, src2, origin);
}

#ifndef SMX_FINISH_BLOCK
#define SMX_FINISH_BLOCK 256
#endif
constexpr int kFinishBlock = SMX_FINISH_BLOCK;

// Stage 1 finish: one thread per (read, primer) closes BOTH strands' searches -- eligible reads decode the column
// histories of the sliced pass, the others run the classic single-word search -- allocates the work entries
// (block-aggregated: one atomic per block and slot, a read's entries stay consecutive) and recovers the start of
// the first location.  A primer is found on at most one strand of nearly every read, so handing a thread both
// strands keeps the warp full through the start-recovery pass (a thread per (read, strand, primer) leaves half
// the lanes idle there, which is why the first form ran it as a separate kernel over the compact entry lists:
// 84 us, 30 % of its instructions per-column window loads -- profiles/r2_a_ncu_head.md).
// Six resident blocks asked for = at most 40 registers (measured on config 2: 56 registers 136 us, 48 registers 127 us,
// 40 registers 124 us, 32 registers with spills 126 us; profiles/r2_u_ab.md).
template <typename W>
__global__ void __launch_bounds__(kFinishBlock, 1536 / kFinishBlock) k_primer_finish(SMX_KARGS, int with_start) {
    // grid: x over reads, y = primer
    __shared__ u64 s_peq[3][16];
    __shared__ u32 s_wtot[2][kFinishBlock / 32 + 1];
    __shared__ u32 s_acc;
    const Tables &t = c_tables;
    const int primer = blockIdx.y;
    if (t.p_sw[primer]) return;                         // long primer: k_primer_long owns these slots
    if (threadIdx.x == 64) s_acc = 0;
    if (threadIdx.x < 48) {
        const u64 *src = threadIdx.x < 16 ? t.peq_rc : threadIdx.x < 32 ? t.peq_rcrev : t.peq_fw;
        s_peq[threadIdx.x >> 4][threadIdx.x & 15] = src[primer * 16 + (threadIdx.x & 15)];
    }
    __syncthreads();
    const u32 read = blockIdx.x * blockDim.x + threadIdx.x;
    u32 cells = 0;
    int nloc[2] = {0, 0};
    u32 ev[2][kFinishMaskWords];
    bool have_ev[2] = {false, false};
    if (read < b.n_reads) {
#pragma unroll
        for (int s = 0; s < 2; ++s) nloc[s] = primer_finish_thread<W>(t, b, read, s, primer, s_peq[0], s_peq[2], ev[s], &have_ev[s]);
        const int n = (int)b.lengths[read];
        cells = 2u * (u32)(n < t.L ? n : t.L);                               // HW columns of the two searches
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        incl[s] = nloc[s];
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl[s], o);
            if (lane >= o) incl[s] += v;
        }
        if (lane == 31) s_wtot[s][warp] = (u32)incl[s];
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int s = threadIdx.x;
        u32 run = 0;
        for (int w = 0; w < kFinishBlock / 32; ++w) { const u32 v = s_wtot[s][w]; s_wtot[s][w] = run; run += v; }
        s_wtot[s][kFinishBlock / 32] = run ? atomicAdd(&b.slot_count[s * t.n_primers + primer], run) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s)
        if (nloc[s])
            write_entries(t, b, (u32)(s * t.n_primers + primer), read,
                          s_wtot[s][kFinishBlock / 32] + s_wtot[s][warp] + (u32)(incl[s] - nloc[s]), have_ev[s] ? ev[s] : nullptr);
    // start of the first location, single-word form: every matched slot when with_start == 2, else only what the
    // sliced start kernel leaves out (reads on the 4-bit side stream; primers with m + k > 32)
    if (with_start && (nloc[0] | nloc[1]) &&
        (with_start == 2 || !start_sliced_ok(t, primer) || read_is_flagged(b, read))) {
        const Geo g = make_geo((int)b.lengths[read], t.L);
        const int s1 = nloc[0] ? 0 : 1;                 // the strand that matched (per lane: the warp stays converged)
        {
            const smx_primer_hit &h = b.phit[(u64)slot_index(t, s1, primer) * b.n_pad + read];
            primer_start_slot<W>(t, b, read, s1, primer, h.first_end - g.woff - g.delta, s_peq[1]);
        }
        if (nloc[0] && nloc[1]) {                        // both strands matched (rare)
            const smx_primer_hit &h = b.phit[(u64)slot_index(t, 1, primer) * b.n_pad + read];
            primer_start_slot<W>(t, b, read, 1, primer, h.first_end - g.woff - g.delta, s_peq[1]);
        }
    }
    const int m = t.p_len[primer];
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
    constexpr int kReads = LongSeg<SW>::kReads;          // reads (whole segments) per warp; lanes beyond them idle
    const int lane = threadIdx.x & 31, sub = lane % SW;
    const bool top = sub == SW - 1;
    const int strand = (int)blockIdx.y;
    const u32 warp_global = (u32)(((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const u32 read = warp_global * (u32)kReads + (u32)(lane / SW);
    const bool valid = lane < kReads * SW && read < b.n_reads;
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
    if (valid && top) for (int w = 0; w < t.mw; ++w) { emask[(u64)w * b.n_pad] = 0; imask[(u64)w * b.n_pad] = 0; }
    u32 Pv = ~0u, Mv = 0u;
    int score = m, best = m + 1;
    {
        // The warp walks the window positions p in step (the same p in every lane; a read whose window starts later or
        // ends earlier -- short reads -- sits those columns out), eight columns = one 4-bit window word at a time, so
        // the word load, the history-word stores and the column's bit are uniform across the warp and every lane runs
        // the same straight-line code per column: the running best / equal / improved bits are plain selects that
        // only mean something in a segment's top lane.  The first form branched per column for the word load, for the
        // top lane's bookkeeping and for idle columns: ~100 executed instructions per lane and column in the worst
        // path (profiles/r1_v30_long_primer_ncu.md: 72 on average).
        const int p_end = p_begin + cols;
        const int p_lo = __reduce_min_sync(0xffffffffu, cols > 0 ? p_begin : 0x7fffffff);
        const int p_hi = __reduce_max_sync(0xffffffffu, cols > 0 ? p_end : 0);
        u32 eqw = 0, imw = 0;
        for (int p0 = p_lo & ~7; p0 < p_hi; p0 += 8) {
            const bool touches = p0 + 8 > p_begin && p0 < p_end;
            const u32 w = touches ? wwin[(u64)(p0 >> 3) * b.n_pad] : 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int p = p0 + q;
                const bool active = p >= p_begin && p < p_end;
                const int c = active ? (int)((w >> (4 * q)) & 15u) : kSymOther;
                const u32 sPv = Pv, sMv = Mv;
                const int d = long_step<SW, false>(s_peq[0][c * SW + sub], Pv, Mv, sub, lane);
                Pv = active ? Pv : sPv;
                Mv = active ? Mv : sMv;
                const int ns = score + d;
                const bool improved = active && ns < best;
                score = active ? ns : score;
                best = improved ? ns : best;
                const u32 bit = 1u << (p & 31);
                imw |= improved ? bit : 0u;
                eqw |= (active && score == best) ? bit : 0u;
            }
            if ((p0 & 31) == 24 || p0 + 8 >= p_hi) {        // uniform: the 32-column history words are complete
                if (top && valid && p0 + 8 > p_begin && (p0 & ~31) < p_end) {
                    emask[(u64)(p0 >> 5) * b.n_pad] = eqw; imask[(u64)(p0 >> 5) * b.n_pad] = imw;
                }
                eqw = imw = 0;
            }
        }
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
            const int p = first - j;
            if (active && (j == 0 || (p & 7) == 7)) wcur = wwin[(u64)(p >> 3) * b.n_pad];      // walking down: a new word at its top symbol
            const int c = active ? (int)((wcur >> (4 * (p & 7))) & 15u) : kSymOther;
            const u32 sPv = Pv, sMv = Mv;
            const int d = long_step<SW, true>(s_peq[1][c * SW + sub], Pv, Mv, sub, lane);
            Pv = active ? Pv : sPv;
            Mv = active ? Mv : sMv;
            rs = active ? rs + d : rs;
            last = (active && rs == best) ? j : last;
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

cudaError_t launch_barcode_tasks(const Tables &t, const Batch &b, const unsigned short *class_tasks,
                                 const BtClass *classes, int n_classes, cudaStream_t st, int *launches) {
    for (int i = 0; i < n_classes; ++i) {
        const BtClass &c = classes[i];
        const unsigned short *list = class_tasks + c.off;
        cudaError_t e;
        switch (t.k_idx) {
#define SMX_K2(KK) case KK: e = launch_barcode_class_k##KK(t, b, list, (int)c.count, c.nw, c.m, c.quad, st); break;
            SMX_K2(0) SMX_K2(1) SMX_K2(2) SMX_K2(3) SMX_K2(4) SMX_K2(5) SMX_K2(6) SMX_K2(7) SMX_K2(8)
            SMX_K2(9) SMX_K2(10) SMX_K2(11) SMX_K2(12)
#undef SMX_K2
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
    }
    return cudaSuccess;
}

cudaError_t launch_stage_windows(const Tables &t, const Batch &b, cudaStream_t st) {
    // shared-memory tile: 128 reads x the words of a clipped read (+1 word of slack per read, +2 per tile)
    u32 tile_words = 0;
    if (b.clip) tile_words = (u32)kStageBlock * ((2 * b.clip + 15) / 16 + 1) + 2;
    if (tile_words * sizeof(u32) > 48u * 1024u) tile_words = 0;           // beyond the default dynamic limit: direct loads
    k_stage_windows<<<(b.n_reads + kStageBlock - 1) / kStageBlock, dim3(kStageBlock, 2), tile_words * sizeof(u32), st>>>(t, b, tile_words);
    return cudaGetLastError();
}

cudaError_t launch_primer_finish(const Tables &t, const Batch &b, int with_start, cudaStream_t st) {
    dim3 grid((b.n_reads + kFinishBlock - 1) / kFinishBlock, t.n_primers);
    if (t.use64) k_primer_finish<u64><<<grid, kFinishBlock, 0, st>>>(t, b, with_start);
    else k_primer_finish<u32><<<grid, kFinishBlock, 0, st>>>(t, b, with_start);
    return cudaGetLastError();
}

cudaError_t launch_primer_long(const Tables &t, const Batch &b, int primer, cudaStream_t st) {
    const int sw = t.p_sw[primer];
    const unsigned reads_per_block = 4u * (unsigned)(32 / sw);         // 128 threads = 4 warps of 32 / sw reads each
    dim3 lgrid((b.n_reads + reads_per_block - 1) / reads_per_block, 2);
    switch (sw) {
        case 3: k_primer_long<3><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        case 4: k_primer_long<4><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        case 5: k_primer_long<5><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        case 6: k_primer_long<6><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        case 8: k_primer_long<8><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        case 10: k_primer_long<10><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        case 16: k_primer_long<16><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        case 32: k_primer_long<32><<<lgrid, 128, 0, st>>>(t, b, primer); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace smx
