// gemm_topk.h -- host interface of the batched tensor-core search (K2 + K4), gemm_topk.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace vdbk {

struct GemmPlan {          // per shard: cached tensor maps for the row slab
    void* impl = nullptr;
};
struct GemmWorkspace {     // per in-flight call: candidate lists, thresholds, fallback flags
    void* impl = nullptr;
};

struct GemmSearchArgs {
    const void* rows; int ld; int dim; bool f16; uint32_t n_rows;
    // optional fp16 shadow plane of fp32 rows [n_rows][ld16]: when set, K2 contracts it (kind::f16) against
    // fp16-rounded queries instead of the fp32 rows (kind::tf32); K4w always re-ranks from `rows`
    const void* shadow = nullptr; int ld16 = 0;
    const float* sqnorm; const uint32_t* labels; const uint32_t* tomb;
    bool prepped = false;  // the caller filled gemm_topk_prep_targets() while preparing the queries
    const float* q;        // prepared queries [nq][ld] fp32
    const float* qn2;      // [nq]
    size_t nq; int k; int metric;   // 0 = L2, 1 = 1 - dot
    const unsigned int* d_max_sqnorm_bits;
    int num_sms;
    int64_t* out_ids; float* out_dist; int* out_counts;
    // optional: called around every launch of the dominant (tensor-core) kernel so the caller can time it
    void (*prof_begin)(void* ctx, cudaStream_t st) = nullptr;
    void (*prof_end)(void* ctx, cudaStream_t st) = nullptr;
    void* prof_ctx = nullptr;
};

// Buffers of the workspace that the caller's query-preparation kernel fills / clears for the next
// gemm_topk_search (saves a conversion kernel and two memsets per batch): fp16 query plane [nq][gld] when the
// contraction runs in kind::f16, the per-query overflow flags and the flagged-query counter.
struct GemmPrepTargets { void* q16 = nullptr; int gld = 0; int* overflow = nullptr; int* n_flagged = nullptr; };
cudaError_t gemm_topk_prep_targets(GemmWorkspace& ws, const struct GemmSearchArgs& a, GemmPrepTargets* out);

// the launch sequence of one search (see gemm_topk.cu): probe, then levels with the rank of the threshold published
// after each (0 = last level, no select)
struct LevelPlan {
    int kp, kq, cap, growth, query_blocks, n_tiles, bits, n_pos, probe_tiles, probe_rank;
    bool small_batch;
    struct Level { int p0, p1, rank_after; };
    std::vector<Level> levels;
};
LevelPlan gemm_topk_level_plan(size_t nq, size_t n_rows, int k);

// Exact re-rank of candidates that another kernel collected (the scan over the fp16 shadow plane, vdb_api.cu): per
// query up to kp keys (approximate DISTANCE, row), their count and tau = a bound from below on the approximate
// distance of every row that is not a candidate.  candidate_buffers hands out the workspace's arrays to fill;
// rerank_candidates enqueues K4w (window re-rank + certificate) and K4x (exact re-search of queries whose
// certificate failed).  a.q / a.qn2 are the prepared queries, a.rows the exact rows; the per-query overflow flags
// and the flagged-query counter must have been cleared (gemm_topk_prep_targets + the prepare kernel).
struct CandidateBuffers { uint64_t* keys = nullptr; size_t stride = 0; int* cnt = nullptr; float* tau = nullptr; };
int gemm_topk_candidate_kp(int k);       // candidates to collect for a top-k (0: k too large for this route)
cudaError_t gemm_topk_candidate_buffers(GemmWorkspace& ws, size_t nq, int kp, CandidateBuffers* out);
cudaError_t gemm_topk_rerank_candidates(GemmWorkspace& ws, const GemmSearchArgs& a, int kp, cudaStream_t st);

bool gemm_topk_supported(int dim, int ld, bool f16, int k, size_t n_rows);
cudaError_t gemm_topk_search(GemmPlan& plan, GemmWorkspace& ws, const GemmSearchArgs& a, cudaStream_t st, std::string& err);
// queries this workspace has re-searched exactly on the device so far (certificate failed / buffer overflowed);
// synchronises the device -- statistics only
long gemm_workspace_fallbacks(const GemmWorkspace& ws);
void gemm_plan_free(GemmPlan& plan);
void gemm_workspace_free(GemmWorkspace& ws);

}  // namespace vdbk
