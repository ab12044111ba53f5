// smx_device.cuh -- pieces shared by the kernel translation units (CUDA only).
#pragma once
#include "smx_kernels.cuh"

namespace smx {

// Tables and Batch travel as __grid_constant__ kernel parameters (constant bank, per launch), so
// several contexts / pipeline lanes can be in flight on one device without sharing a symbol.
#define SMX_KARGS const __grid_constant__ Tables c_tables, const __grid_constant__ Batch b

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

}  // namespace smx
