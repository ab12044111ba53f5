// smx_launch.hpp -- host-callable launchers of the kernels (one translation unit per stage, see Makefile).
// Every launcher enqueues on `st` and returns the launch error (cudaGetLastError), nothing synchronises.
#pragma once
#include <cuda_runtime.h>

#include "smx_core.cuh"

namespace smx {

constexpr int kScanTile = 1024;         // reads per tile of k_scan_compact (sizes the tile status array)

// stage 0 / 1
cudaError_t launch_stage_windows(const Tables &t, const Batch &b, cudaStream_t st);
// prow_code: the primer's 32 pattern-row IUPAC codes (HostTables::prow_code + 32 * primer)
cudaError_t launch_primer_sliced(const Tables &t, const Batch &b, int primer, const unsigned char *prow_code, cudaStream_t st);
// finish of both strands per (read, primer) + work entries + start of the first location in the single-word form:
// with_start 0 = none, 1 = only what launch_primer_start_sliced leaves out, 2 = every matched slot
cudaError_t launch_primer_finish(const Tables &t, const Batch &b, int with_start, cudaStream_t st);
// start of the first location, bit-sliced across 32 work entries (primers with start_sliced_ok); prow_code as above
cudaError_t launch_primer_start_sliced(const Tables &t, const Batch &b, int primer, const unsigned char *prow_code, cudaStream_t st);
cudaError_t launch_primer_long(const Tables &t, const Batch &b, int primer, cudaStream_t st);
// stage 2: one launch per (task width, barcode length) class; class_tasks is the DEVICE copy of
// HostTables::bt_class_tasks.  *launches receives the number of kernels enqueued.
cudaError_t launch_barcode_tasks(const Tables &t, const Batch &b, const unsigned short *class_tasks,
                                 const BtClass *classes, int n_classes, cudaStream_t st, int *launches);
// one translation unit per barcode threshold (smx_k_stage2.cu, -DSMX_STAGE2_K=k)
#define SMX_DECL_K2(KK) cudaError_t launch_barcode_class_k##KK(const Tables &t, const Batch &b, const unsigned short *list, \
                                                                int n_list, int nw, int m, int quad, cudaStream_t st);
SMX_DECL_K2(0) SMX_DECL_K2(1) SMX_DECL_K2(2) SMX_DECL_K2(3) SMX_DECL_K2(4) SMX_DECL_K2(5) SMX_DECL_K2(6) SMX_DECL_K2(7) SMX_DECL_K2(8)
SMX_DECL_K2(9) SMX_DECL_K2(10) SMX_DECL_K2(11) SMX_DECL_K2(12)
#undef SMX_DECL_K2
// stage 3
cudaError_t launch_select_fast(const Tables &t, const Batch &b, cudaStream_t st);
cudaError_t launch_select(const Tables &t, const Batch &b, cudaStream_t st);
cudaError_t launch_select_big(const Tables &t, const Batch &b, const u32 *list, u32 n_list, unsigned char *scratch, cudaStream_t st);
cudaError_t launch_scan_compact(const Tables &t, const Batch &b, u32 rec_cap, unsigned long long *tile_status, u32 *ticket,
                                u32 ticket_base, u32 epoch, cudaStream_t st);
cudaError_t launch_pack_records32(const smx_record *in, u32 n, smx_record32 *out, cudaStream_t st);
cudaError_t launch_pack_records16(const smx_record *in, u32 n, const u32 *lengths, u32 read_base, smx_record16 *out, u32 *bad,
                                  cudaStream_t st);
cudaError_t launch_expand_lengths(const unsigned short *in, u32 n, u32 *out, cudaStream_t st);
cudaError_t launch_rebase_offsets(const u32 *in, u32 n, u32 rec_base, u32 *out, cudaStream_t st);
// specimine: HW distances of the patterns listed in pat_list (all of one (lanes, words per lane) class) in every text
cudaError_t launch_hw_distance(int sw, int w, const unsigned char *pat, const u64 *pat_off, const i32 *pat_k, const u32 *pat_list,
                               u32 n_list, const unsigned char *txt, const u64 *txt_off, u32 n_txt, i32 *out, u32 *bad,
                               cudaStream_t st);
// utilities
cudaError_t launch_pairwise_nw(const char *seqs, const u32 *off, u32 n, i32 *out, cudaStream_t st);
cudaError_t launch_int_peak(int mode, int blocks, int threads, u32 *out, int iters, u32 seed, cudaStream_t st);

}  // namespace smx
