// smx_k_stage2.cu -- stage 2: barcode SHW search, bit-sliced across barcodes (smx_kernels.cuh: barcode_task_thread).
#include <cuda_runtime.h>

#include "smx_device.cuh"
#include "smx_launch.hpp"

namespace smx {

// One thread per work entry of a matched slot; the thread evaluates the entry against the NWQ bwords of one
// stage-2 task.  blockIdx.y = strand * n_list + (index into task_list); the task's interleaved table sits in
// shared memory.
template <int K, int NWQ>
__global__ void __launch_bounds__(128) k_barcode_task(SMX_KARGS, const unsigned short *task_list, int n_list) {
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
    const int m = t.bw_len[g0];
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
        work = barcode_task_thread<K, NWQ>(t, b, read, p, idx, strand, primer, task, s_tab);
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

template <int K>
static cudaError_t launch_k(const Tables &t, const Batch &b, const unsigned short *class_tasks,
                            const u32 class_off[kMaxTaskWords + 1], cudaStream_t st, int *launches) {
    for (int w = 1; w <= kMaxTaskWords; ++w) {
        const int n_list = (int)(class_off[w] - class_off[w - 1]);
        if (!n_list) continue;
        const unsigned short *list = class_tasks + class_off[w - 1];
        dim3 grid((b.e_cap + 127) / 128, 2 * n_list);
        if (w == 1) k_barcode_task<K, 1><<<grid, 128, 0, st>>>(t, b, list, n_list);
        else if (K > kMaxTaskK) return cudaErrorInvalidValue;          // the host never builds such tasks
        else if (w == 2) k_barcode_task<(K > kMaxTaskK ? 0 : K), 2><<<grid, 128, 0, st>>>(t, b, list, n_list);
        else if (w == 3) k_barcode_task<(K > kMaxTaskK ? 0 : K), 3><<<grid, 128, 0, st>>>(t, b, list, n_list);
        else k_barcode_task<(K > kMaxTaskK ? 0 : K), 4><<<grid, 128, 0, st>>>(t, b, list, n_list);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_barcode_tasks(const Tables &t, const Batch &b, const unsigned short *class_tasks,
                                 const u32 class_off[kMaxTaskWords + 1], cudaStream_t st, int *launches) {
    switch (t.k_idx) {
#define SMX_K2(KK) case KK: return launch_k<KK>(t, b, class_tasks, class_off, st, launches);
        SMX_K2(0) SMX_K2(1) SMX_K2(2) SMX_K2(3) SMX_K2(4) SMX_K2(5) SMX_K2(6) SMX_K2(7) SMX_K2(8)
#undef SMX_K2
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace smx
