// smx_k_stage2.cu -- stage 2: barcode SHW search, bit-sliced across barcodes (smx_kernels.cuh: barcode_task_thread).
// Compiled once per barcode threshold: -DSMX_STAGE2_K=<k_idx> (Makefile), so the instantiations of the nine
// thresholds build side by side.
#include <cuda_runtime.h>

#include "smx_device.cuh"
#include "smx_launch.hpp"

#ifndef SMX_STAGE2_K
#error "compile with -DSMX_STAGE2_K=<0..12>"
#endif

namespace smx {

// One thread per work entry of a matched slot; the thread evaluates the entry against the NWQ bwords of one
// stage-2 task.  blockIdx.y = strand * n_list + (index into task_list); the task's interleaved table sits in
// shared memory.  MF > 0: the barcode length is a template argument (the unrolled row sequence then has no early
// exit: registers are renamed from row to row without moves and the carry-save counter's bookkeeping folds
// away -- 2,944 instead of 3,960 instructions for K = 3, three words, 13 rows); MF = 0: any length.
// Resident blocks per SM the register allocation aims for (knobs of the build for A/B runs: -DSMX_BT_MINB1 one-word
// tasks, -DSMX_BT_MINB3 multi-word tasks).  Measured on config 2 (profiles/r2_u_ab.md): one-word task 91 registers
// (5 blocks) -> 64 registers (8 blocks, a few spills): stage 2 200 -> 194 us; three-word task: 128 registers (4 blocks)
// kept, 96 registers with 140 B of spills was no faster.
#ifndef SMX_BT_MINB1
#define SMX_BT_MINB1 8
#endif
#ifndef SMX_BT_MINB3
#define SMX_BT_MINB3 4
#endif
template <int K, int NWQ, int MF>
__global__ void __launch_bounds__(128, NWQ == 1 ? SMX_BT_MINB1 : SMX_BT_MINB3) k_barcode_task(SMX_KARGS, const unsigned short *task_list, int n_list) {
    constexpr int S = NWQ == 1 ? 1 : NWQ == 2 ? 2 : 4;
    constexpr int kRows = NWQ == 1 ? SMX_MAX_PATTERN : 16;      // multi-word tasks only exist for m + K <= 16
    __shared__ __align__(16) u32 s_tab[kRows * 16 * S];
    __shared__ u32 s_acc;
    const Tables &t = c_tables;
    const u32 task = task_list[blockIdx.y % n_list];
    const int strand = blockIdx.y / n_list;
    const u32 g0 = t.bt_g0[task];
    const int primer = t.bw_primer[g0];
    const u32 slot = slot_index(t, strand, primer);
    u32 cnt = b.slot_count[slot];
    if (cnt > b.e_cap) cnt = b.e_cap;
    if (blockIdx.x * blockDim.x >= cnt) return;
    const int m = MF ? MF : (int)t.bw_len[g0];
    if (threadIdx.x == 0) s_acc = 0;
    {
        const u32 *src = t.bt_eq + t.bt_row[task];
        for (int i = threadIdx.x; i < m * 16 * S; i += blockDim.x) s_tab[i] = src[i];
    }
    __syncthreads();
    const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    u32 work = 0;
    if (idx < cnt) {
        const u32 read = b.ent_read[(u64)slot * b.e_cap + idx];
        const int p = b.ent_pos[(u64)slot * b.e_cap + idx];
        work = barcode_task_thread<K, NWQ, MF>(t, b, read, p, idx, strand, primer, task, s_tab);
    }
    // m is uniform over the block: cells = m * W, word-columns = ceil(m/32) * W with W = sum of lanes x columns (low 20
    // bits of `work`); the bits above count the bwords whose automaton ran (x band cells = evaluated bit-sliced cells)
    const u32 wsum = __reduce_add_sync(0xffffffffu, work);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&s_acc, wsum);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long tot = s_acc & 0xFFFFFu, ran = s_acc >> 20;
        if (tot) {
            atomicAdd(&b.counters[1], tot * (unsigned long long)m);
            atomicAdd(&b.counters[3], tot * (unsigned long long)((m + 31) >> 5));
        }
        if (ran) atomicAdd(&b.counters[kCtrUseful2], ran * (unsigned long long)band_cells(m, K, m + K <= 16));
    }
}

// Narrow single-word tasks, four work entries to a thread (barcode_quad_thread): blockIdx.x walks the entries in
// groups of 4 * blockDim.x; the 625-entry-per-row table and the task's ordinary table (for the entries that fall back
// to the one-entry routine) sit in shared memory.
template <int K, int MF>
__global__ void __launch_bounds__(128) k_barcode_quad(SMX_KARGS, const unsigned short *task_list, int n_list) {
    __shared__ u32 s_tab4[16 * kQuadRow];
    __shared__ u32 s_tab1[16 * 16];
    __shared__ u32 s_acc;
    const Tables &t = c_tables;
    const u32 task = task_list[blockIdx.y % n_list];
    const int strand = blockIdx.y / n_list;
    const u32 g0 = t.bt_g0[task];
    const int primer = t.bw_primer[g0];
    const u32 slot = slot_index(t, strand, primer);
    u32 cnt = b.slot_count[slot];
    if (cnt > b.e_cap) cnt = b.e_cap;
    if (blockIdx.x * blockDim.x * 4u >= cnt) return;
    const int m = MF ? MF : (int)t.bw_len[g0];
    if (threadIdx.x == 0) s_acc = 0;
    {
        const u32 *src4 = t.bt_quad + t.bt_quad_row[task], *src1 = t.bt_eq + t.bt_row[task];
        for (int i = threadIdx.x; i < m * kQuadRow; i += blockDim.x) s_tab4[i] = src4[i];
        for (int i = threadIdx.x; i < m * 16; i += blockDim.x) s_tab1[i] = src1[i];
    }
    __syncthreads();
    const u32 e0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    u32 work = 0;
    if (e0 < cnt) work = barcode_quad_thread<K, MF>(t, b, slot, e0, cnt, strand, primer, task, s_tab4, s_tab1);
    const u32 wsum = __reduce_add_sync(0xffffffffu, work);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&s_acc, wsum);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long tot = s_acc & 0xFFFFFu, ran = s_acc >> 20;
        if (tot) {
            atomicAdd(&b.counters[1], tot * (unsigned long long)m);
            atomicAdd(&b.counters[3], tot * (unsigned long long)((m + 31) >> 5));
        }
        if (ran) atomicAdd(&b.counters[kCtrUseful2], ran * (unsigned long long)band_cells(m, K, true));
    }
}

// barcode lengths with their own instantiation (the lengths barcode sets are built in; every other length runs the
// any-length form)
template <int K, int M> struct FixedLen { static constexpr int value = (K >= 1 && K <= kMaxTaskK && M > K && M + K <= 16) ? M : 0; };

template <int K, int NWQ>
static cudaError_t launch_class(const Tables &t, const Batch &b, const unsigned short *list, int n_list, int m, cudaStream_t st) {
    dim3 grid((b.e_cap + 127) / 128, 2 * n_list);
    switch (m) {
#define SMX_MF(MM)                                                                                         \
        case MM:                                                                                           \
            if (FixedLen<K, MM>::value) {                                                                  \
                k_barcode_task<K, NWQ, FixedLen<K, MM>::value><<<grid, 128, 0, st>>>(t, b, list, n_list);  \
                return cudaGetLastError();                                                                 \
            }                                                                                              \
            break;
        SMX_MF(8) SMX_MF(10) SMX_MF(12) SMX_MF(13)
#undef SMX_MF
        default: break;
    }
    k_barcode_task<K, NWQ, 0><<<grid, 128, 0, st>>>(t, b, list, n_list);
    return cudaGetLastError();
}

#define SMX_CAT_(a, b) a##b
#define SMX_CAT(a, b) SMX_CAT_(a, b)

template <int K>
static cudaError_t launch_quad(const Tables &t, const Batch &b, const unsigned short *list, int n_list, int m, cudaStream_t st) {
    dim3 grid((b.e_cap / 4 + 128) / 128, 2 * n_list);
    switch (m) {
#define SMX_MF(MM)                                                                                         \
        case MM:                                                                                           \
            if (FixedLen<K, MM>::value) {                                                                  \
                k_barcode_quad<K, FixedLen<K, MM>::value><<<grid, 128, 0, st>>>(t, b, list, n_list);       \
                return cudaGetLastError();                                                                 \
            }                                                                                              \
            break;
        SMX_MF(8) SMX_MF(10) SMX_MF(12) SMX_MF(13)
#undef SMX_MF
        default: break;
    }
    k_barcode_quad<K, 0><<<grid, 128, 0, st>>>(t, b, list, n_list);
    return cudaGetLastError();
}

cudaError_t SMX_CAT(launch_barcode_class_k, SMX_STAGE2_K)(const Tables &t, const Batch &b, const unsigned short *list,
                                                          int n_list, int nw, int m, int quad, cudaStream_t st) {
    constexpr int K = SMX_STAGE2_K;
    if (quad) return launch_quad<K>(t, b, list, n_list, m, st);
    if (nw == 1) return launch_class<K, 1>(t, b, list, n_list, m, st);
    if (K > kMaxTaskK) return cudaErrorInvalidValue;            // the host never builds multi-word tasks there
    constexpr int KM = K > kMaxTaskK ? 0 : K;
    if (nw == 2) return launch_class<KM, 2>(t, b, list, n_list, m, st);
    if (nw == 3) return launch_class<KM, 3>(t, b, list, n_list, m, st);
    return launch_class<KM, 4>(t, b, list, n_list, m, st);
}

}  // namespace smx
