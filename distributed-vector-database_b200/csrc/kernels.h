// kernels.h -- launcher interface between the C-ABI host layer (vdb_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vdbk {

// ---- K1 scan (scan_topk.cu) ---------------------------------------------------------------
struct ScanParams {
    const void* rows;        // [n_rows][ld] T, row stride row_bytes
    uint32_t row_bytes;      // ld * sizeof(T), multiple of 512
    uint32_t ld;             // padded row length in elements
    uint32_t n_rows;
    const uint32_t* labels;  // [n_rows], or null: label == label_base + row (the usual case: ids handed out in insertion
    uint32_t label_base;     // order).  A row that passes the threshold then costs the consumers no global load.
    const uint32_t* tomb;    // bitmap, 1 = deleted, or null
    const float* q;          // [nq][ld] prepared queries (fp32, zero padded) -- or:
    const float* q_raw;      // [nq][dim] raw queries, prepared inside the kernel (saves a launch)
    int dim, normalize;
    int nq;                  // 1..8
    int k;
    int metric;              // 0 = squared L2 (direct form), 1 = 1 - dot
    int stages, ring;        // filled by the launcher (copy-ring stages, candidate-ring slots)
    int dbg;                 // experiments (VDB_SCAN_DBG): 1 = no FMA loop, 2 = no L2 policy hint, 4 = no select
    uint64_t* out_keys;      // [nq][grid][k]
    // In-kernel merge (merge_group > 0): the last CTA of every group of `merge_group` CTAs reduces the group's lists,
    // the last group to finish reduces the group results and writes the final answer -- no merge launch after the scan.
    int merge_group;         // CTAs per group (>= grid: one level); 0 = leave the per-CTA lists to the merge kernel
    unsigned int* merge_ctr; // [1 + groups] arrival counters, zero before the launch; the kernel leaves them zero
    uint64_t* group_keys;    // [nq][groups][k] group results (two levels only)
    int64_t* out_ids;        // [nq][k] final results (-1 / +inf padded), out_counts optional
    float* out_dist;
    int* out_counts;
    // Candidate mode (in-kernel merge only; the scan ran over an APPROXIMATE plane of the rows, e.g. the fp16 shadow):
    // instead of final results the k best keys (distance, label) go to cand_keys[q][0..k), their number to cand_cnt[q]
    // and cand_tau[q] = the k-th distance (+inf with fewer than k live rows): every row that is not a candidate lies
    // at or above it.  The exact re-rank (K4w) takes it from there.
    uint64_t* cand_keys; size_t cand_stride; int* cand_cnt; float* cand_tau;
};
struct ScanPlan { int grid, ctas_per_sm, stages; size_t smem; int merge_group, merge_groups; };   // grid == 0: does not fit
constexpr int SCAN_MERGE_KEYS = 4096;    // keys one in-kernel merge step holds in shared memory
ScanPlan scan_plan(int nq, uint32_t ld, uint32_t row_bytes, int k, uint32_t n_rows, int num_sms);
cudaError_t launch_scan_topk(ScanParams p, bool f16, const ScanPlan& pl, cudaStream_t st);
int scan_max_k(int nq_t, int ld, uint32_t row_bytes);

// ---- K5 merge (merge_topk.cu) -------------------------------------------------------------
struct MergeParams {
    // input mode A: packed keys, per query a contiguous block of n_in keys
    const uint64_t* in_keys;   // [nq][n_in]
    // input mode B: (dist,id) lists laid out [G][nq][k_in]; id < 0 is padding
    const float* in_dist;
    const int64_t* in_ids;
    int G, k_in;               // mode C: in_keys != null and G > 0: packed keys [G][key_stride_g] with query rows of k_in
    size_t key_stride_g;
    size_t nq;
    int n_in;                  // keys per query (mode A) or G*k_in (mode B)
    int k_out;
    uint64_t* out_keys;        // optional [nq][k_out]
    int64_t* out_ids;          // optional [nq][k_out]  (-1 padded)
    float* out_dist;           // optional [nq][k_out]  (+inf padded)
    int* out_counts;           // optional [nq]
};
cudaError_t launch_merge_topk(const MergeParams& p, cudaStream_t st);

// ---- K5x exchange + merge over NVLink peer memory (merge_topk.cu) ---------------------------
constexpr int XCHG_MAX_WORLD = 16;
enum { XCHG_ERR_TIMEOUT = 1, XCHG_ERR_SHAPE = 2, XCHG_ERR_STEP = 3 };
struct XchgParams {
    int rank, world, parity;
    uint32_t step;                 // 1, 2, 3, ... (flags start at 0)
    uint32_t shape;                // hash of (nq, k): must be the same on every rank of a step
    unsigned long long timeout_ns; // bound on the wait for the peers
    uint32_t* err;                 // host-visible (mapped pinned) [4]: code, step, peer, -
    size_t nq, slice, owned;       // global batch; slice = ceil(nq / world); queries this rank owns (<= slice)
    int k;
    const int64_t* ids;            // this rank's lists [nq][k] (id < 0 = padding)
    const float* dist;
    uint64_t* peer_buf[XCHG_MAX_WORLD];    // receive buffer of every rank: [2][world][stride_src] keys
    uint64_t* peer_flag[XCHG_MAX_WORLD];   // flag words of every rank: [2][world] of (shape << 32 | step)
    const uint64_t* local_flag;
    size_t stride_parity, stride_src;
    unsigned int* done_counter;    // local, zeroed by the launcher
};
cudaError_t launch_exchange_merge(const XchgParams& x, const MergeParams& mp, int num_sms, cudaStream_t st);
// arrival signal / wait of the copy-engine query all-gather (vdb_xchg_gather_queries / vdb_xchg_wait_queries)
struct XchgSignal { int world; uint32_t value; uint32_t* flag[XCHG_MAX_WORLD]; };   // this rank's word in every rank's array
struct XchgWait { int world; uint32_t value; const uint32_t* flag; unsigned long long timeout_ns; uint32_t* err; };
cudaError_t launch_xchg_signal(const XchgSignal& s, cudaStream_t st);
cudaError_t launch_xchg_wait(const XchgWait& w, cudaStream_t st);

// ---- K0 / K3 / utilities (insert.cu) ------------------------------------------------------
// queries [nq][dim] fp32 -> prepared [nq][ld] fp32 (normalised when cosine, zero padded),
// qn2[nq] = sum of squares of the prepared query.  Optional: fp16 copy [nq][ld16] (operand of the kind::f16
// contraction) and two flag arrays to clear (zero_per_query[nq], *zero_one).
cudaError_t launch_prepare_queries(const float* q, size_t nq, int dim, int ld, bool normalize, float* out,
                                   float* qn2, cudaStream_t st, void* out16 = nullptr, int ld16 = 0,
                                   int* zero_per_query = nullptr, int* zero_one = nullptr);
// src [n][dim] fp32 -> shard rows [row0 .. row0+n) (T = f32/f16, stride ld), sqnorm, max norm.
cudaError_t launch_insert_rows(const float* src, size_t n, int dim, int ld, bool normalize, bool f16, void* rows,
                               float* sqnorm, size_t row0, unsigned int* max_sqnorm_bits, cudaStream_t st);
// fp32 rows [row0, row0+n) (stride ld) -> fp16 shadow plane (stride ld16 <= ld)
cudaError_t launch_shadow_rows(const float* rows, int ld, void* shadow, int ld16, size_t row0, size_t n, cudaStream_t st);
cudaError_t launch_synth_rows(uint64_t seed, uint64_t row_start, size_t n, int dim, float* out, cudaStream_t st);
cudaError_t launch_set_bits(uint32_t* bitmap, const uint32_t* rows, size_t n, bool set, cudaStream_t st);
cudaError_t launch_gather_rows(const void* rows, int ld, int dim, bool f16, const uint32_t* idx, size_t n, float* out,
                               cudaStream_t st);
cudaError_t launch_iota_u32(uint32_t* out, size_t n, uint32_t start, cudaStream_t st);

// ---- K2 / K2s / K4w batched tensor-core path (gemm_topk.cu) ---------------------------
struct GemmTopkPlan;  // opaque, owns tensor maps

uint64_t launch_count();
void count_launch();

}  // namespace vdbk
