// smx_k_sliced.cu -- stage 1 forward pass, bit-sliced across reads (one instantiation per primer length).
#include <cuda_runtime.h>

#include "smx_device.cuh"
#include "smx_launch.hpp"

namespace smx {

// Sliced primer search: one thread per (group of 32 reads, strand) for one primer of length M.
#ifndef SMX_SLICED_BLOCK
#define SMX_SLICED_BLOCK 64
#endif
constexpr int kSlicedBlock = SMX_SLICED_BLOCK;        // small blocks: equal-length tasks, let the block scheduler balance the SMs
template <int M>
__global__ void __launch_bounds__(kSlicedBlock) k_primer_sliced(SMX_KARGS, int primer, const __grid_constant__ RowOffsets ro, int degenerate) {
    __shared__ u32 s_planes[(kSlicedCodes + 48) * kSlicedBlock];
    const u32 group = blockIdx.x * kSlicedBlock + threadIdx.x;
    if (threadIdx.x == 0) {             // evaluated bit-sliced cells: M rows x L columns per thread of the block
        const u32 groups = b.n_pad / 32, first = blockIdx.x * kSlicedBlock;
        const u32 act = groups > first ? (groups - first < (u32)kSlicedBlock ? groups - first : (u32)kSlicedBlock) : 0u;
        if (act) atomicAdd(&b.counters[kCtrUseful1], (unsigned long long)act * M * c_tables.L);
    }
    if (group >= b.n_pad / 32) return;
    primer_sliced_thread<M, kSlicedBlock>(c_tables, b, group, (int)blockIdx.y, primer, ro, degenerate != 0,
                                          s_planes + threadIdx.x, s_planes + kSlicedCodes * kSlicedBlock + threadIdx.x);
}


// Start recovery, bit-sliced across work entries: one thread owns 32 consecutive entries of a slot, a block of 64
// threads 2,048.  The entries' window words are gathered by the WHOLE block, one entry per thread and step (the
// first form let every thread gather its own 32 entries one after the other: 258 us at 10 % ALU, a chain of
// dependent loads per entry -- profiles/r2_i_ncu_full.md), the reverse passes run bit-sliced, and the starts go
// back out the same cooperative way.
template <int M>
__global__ void __launch_bounds__(kSlicedBlock) k_primer_start_sliced(SMX_KARGS, int primer, const __grid_constant__ RowOffsets ro, int degenerate) {
    __shared__ u32 s_planes[(kSlicedCodes + kStartPlanes) * kSlicedBlock];
    const Tables &t = c_tables;
    const u32 slot = blockIdx.y * t.n_primers + primer;
    u32 cnt = b.slot_count[slot];
    if (cnt > b.e_cap) cnt = b.e_cap;
    const u32 block_e0 = blockIdx.x * kSlicedBlock * 32u;
    if (block_e0 >= cnt) return;
    u32 *sa = s_planes + kSlicedCodes * kSlicedBlock;              // [plane word][owner thread]
#pragma unroll 4
    for (int it = 0; it < 32; ++it) {
        const u32 idx = (u32)it * kSlicedBlock + threadIdx.x;      // entry block_e0 + idx: owner idx / 32, its word idx % 32
        u32 xh, xl, act, best;
        start_gather_entry(t, b, slot, block_e0 + idx, cnt, M, xh, xl, act, best);
        u32 *dst = sa + (idx & 31u) * kSlicedBlock + (idx >> 5);
        dst[0] = xh; dst[32 * kSlicedBlock] = xl; dst[64 * kSlicedBlock] = act; dst[96 * kSlicedBlock] = best;
    }
    __syncthreads();
    if (block_e0 + threadIdx.x * 32u < cnt)
        primer_start_sliced_thread<M, kSlicedBlock>((int)t.p_len[primer] + (int)t.p_k[primer], ro, degenerate != 0,
                                                    s_planes + threadIdx.x, sa + threadIdx.x);
    __syncthreads();
#pragma unroll 4
    for (int it = 0; it < 32; ++it) {
        const u32 idx = (u32)it * kSlicedBlock + threadIdx.x;
        if (block_e0 + (idx & ~31u) >= cnt) continue;              // its owner did not run
        const u32 last = sa[(idx & 31u) * kSlicedBlock + (idx >> 5)];
        if (last != 0xFFFFFFFFu) start_store_entry(t, b, slot, block_e0 + idx, (int)last);
    }
}

cudaError_t launch_primer_start_sliced(const Tables &t, const Batch &b, int primer, const unsigned char *prow_code, cudaStream_t st) {
    const int m = t.p_len[primer];
    dim3 sgrid((b.e_cap / 32 + kSlicedBlock) / kSlicedBlock, 2);
    RowOffsets ro;
    bool degenerate = false;
    for (int i = 0; i < 32; ++i) {                      // rows of the REVERSED primer_rc
        int code = i < m ? prow_code[m - 1 - i] : 0;
        degenerate |= code > 3;
        ro.off[i] = (unsigned short)(code * kSlicedBlock * sizeof(u32));
    }
    switch (m) {
#define SMX_M(MM) case MM: k_primer_start_sliced<MM><<<sgrid, kSlicedBlock, 0, st>>>(t, b, primer, ro, degenerate ? 1 : 0); break;
        SMX_M(1) SMX_M(2) SMX_M(3) SMX_M(4) SMX_M(5) SMX_M(6) SMX_M(7) SMX_M(8) SMX_M(9) SMX_M(10) SMX_M(11)
        SMX_M(12) SMX_M(13) SMX_M(14) SMX_M(15) SMX_M(16) SMX_M(17) SMX_M(18) SMX_M(19) SMX_M(20) SMX_M(21)
        SMX_M(22) SMX_M(23) SMX_M(24) SMX_M(25) SMX_M(26) SMX_M(27) SMX_M(28) SMX_M(29) SMX_M(30) SMX_M(31)
        SMX_M(32)
#undef SMX_M
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_primer_sliced(const Tables &t, const Batch &b, int primer, const unsigned char *prow_code, cudaStream_t st) {
    dim3 sgrid((b.n_pad / 32 + kSlicedBlock - 1) / kSlicedBlock, 2);
    RowOffsets ro;
    bool degenerate = false;
    for (int i = 0; i < 32; ++i) {
        int code = i < t.p_len[primer] ? prow_code[i] : 0;
        degenerate |= code > 3;
        ro.off[i] = (unsigned short)(code * kSlicedBlock * sizeof(u32));
    }
    switch (t.p_len[primer]) {
#define SMX_M(MM) case MM: k_primer_sliced<MM><<<sgrid, kSlicedBlock, 0, st>>>(t, b, primer, ro, degenerate ? 1 : 0); break;
        SMX_M(1) SMX_M(2) SMX_M(3) SMX_M(4) SMX_M(5) SMX_M(6) SMX_M(7) SMX_M(8) SMX_M(9) SMX_M(10) SMX_M(11)
        SMX_M(12) SMX_M(13) SMX_M(14) SMX_M(15) SMX_M(16) SMX_M(17) SMX_M(18) SMX_M(19) SMX_M(20) SMX_M(21)
        SMX_M(22) SMX_M(23) SMX_M(24) SMX_M(25) SMX_M(26) SMX_M(27) SMX_M(28) SMX_M(29) SMX_M(30) SMX_M(31)
        SMX_M(32)
#undef SMX_M
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace smx
