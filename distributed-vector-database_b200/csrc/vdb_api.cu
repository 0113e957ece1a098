// vdb_api.cu -- host side of the C ABI declared in include/vdb.h.
//
// Owns the GPU-resident shard (rows, norms, labels, tombstone bitmap), the label->row map,
// a pool of per-call workspaces (stream + scratch) and dispatches a search to the scan kernel
// (K1) or the tensor-core kernel (K2) followed by the merge (K5).  No CPU compute path exists:
// if CUDA is unavailable every entry point fails with VDB_ECUDA.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vdb.h"
#include "common.cuh"
#include "gemm_topk.h"
#include "growbuf.h"
#include "kernels.h"

namespace vdbk {
static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace vdbk

using namespace vdbk;

static thread_local std::string t_err;
static int fail(int code, const std::string& msg) {
    t_err = msg;
    return code;
}
#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            cudaGetLastError();                                                                       \
            return fail(VDB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
        }                                                                                             \
    } while (0)

namespace {

constexpr int K_MAX = 1024;
constexpr size_t STAGE_ROWS = 32768;   // insert staging chunk
constexpr int MAX_WORKSPACES = 8;

struct Workspace {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool used = false;                // a device-pointer call left work in flight on `last_stream` (see `done`)
    cudaStream_t last_stream = nullptr;
    // device scratch (grow only)
    float* d_q_in = nullptr;  size_t q_in_cap = 0;     // raw queries (host path)
    float* d_q = nullptr;     size_t q_cap = 0;        // prepared queries [nq][ld]
    float* d_qn2 = nullptr;   size_t qn2_cap = 0;
    uint64_t* d_keys = nullptr; size_t keys_cap = 0;   // per-CTA candidate lists
    unsigned int* d_scan_ctr = nullptr;                // arrival counters of the scan kernel's in-kernel merge (kept zero)
    uint64_t* d_group_keys = nullptr;                  // its group results: 2 queries x SCAN_MERGE_KEYS keys
    uint8_t* d_out = nullptr; size_t out_cap = 0;      // results of a host-buffer call: ids [nq*k] | dist [nq*k] | counts [nq],
                                                       // one block, so that small results come back in ONE copy
    // pinned host staging
    float* h_q = nullptr; size_t h_q_cap = 0;
    uint8_t* h_out = nullptr; size_t h_out_cap = 0;
    GemmWorkspace gemm;
};

template <typename T>
cudaError_t grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(need, (size_t)4096);
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
}
template <typename T>
cudaError_t grow_host(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(need, (size_t)4096);
    cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
}

}  // namespace

// Concurrency (SURVEY 8b "Threading"; reference: one RLock around everything, src/datanode/handler.py:23):
//   * searches take `mu` SHARED and snapshot `count` (acquire): they read rows [0, count) only
//   * writers (add / mark_deleted / get_rows / save / resize) serialise on `wmu` and ALSO take `mu` shared: an
//     append fills rows at and above `count` -- which no search looks at -- and publishes the new count (release)
//     once the insert kernels have completed; tombstone bits flip while searches run (a search sees the delete or
//     it does not).  A search never waits for a writer and the other way round.
//   * the slabs grow in place (growbuf.h): resize maps more memory behind the same base pointers.  Only when a
//     slab's address reservation is exhausted (8x the initial capacity) is `mu` taken EXCLUSIVE to move to a
//     larger reservation -- the one operation that makes searches wait.
struct vdb {
    int dim = 0, ld = 0, metric = 0, dtype = 0, device = 0, num_sms = 148;
    std::atomic<size_t> capacity{0};
    std::atomic<size_t> count{0};
    std::atomic<size_t> live{0};
    size_t va_rows = 0;               // rows the address reservations hold
    GrowBuf b_rows, b_shadow, b_sqnorm, b_labels, b_tomb;
    size_t tomb_words_zeroed = 0;
    size_t image_saved = 0;           // rows of this shard known to be in its on-disk image (vdb_save_image)
    void* rows = nullptr;             // == b_rows.ptr() etc.: stable while `mu` is held shared
    void* shadow = nullptr;           // fp16 copy of fp32 rows [capacity][ld16]: operand plane of the tensor path
    int ld16 = 0;
    float* sqnorm = nullptr;
    uint32_t* labels = nullptr;
    uint32_t* tomb = nullptr;
    unsigned int* d_max_sqnorm = nullptr;
    float* d_stage = nullptr;
    uint32_t* d_idx = nullptr; size_t idx_cap = 0;
    cudaStream_t wstream = nullptr;   // writer stream
    mutable std::shared_mutex mu;
    std::mutex wmu;                   // writers, one at a time (lock order: wmu, then mu)
    // label -> row
    std::atomic<bool> affine{true};   // labels are label_base + row: kernels need no label array for the rows they see
    int64_t label_base = 0;
    std::unordered_map<int64_t, uint32_t> map;
    std::vector<uint64_t> h_dead;     // host mirror of the tombstone bitmap (writers only)
    std::atomic<bool> any_dead{false};
    // workspaces
    std::mutex ws_mu;
    std::condition_variable ws_cv;
    std::vector<std::unique_ptr<Workspace>> ws_all;
    std::vector<Workspace*> ws_free;
    // options / stats
    std::atomic<long> opt_path{0};          // 0 auto, 1 force scan, 2 force tensor
    // nq <= this takes the scan kernel in auto mode.  Measured on 1M x 512 (tools/small_batch.py): one scan pass
    // of 1/2/4/8 fp32 queries costs 297/300/341/570 us (beyond 4 queries the pass is LDS-bound, not HBM-bound),
    // the tensor path ~320 us for any small batch -> 4 for fp32 rows; fp16 rows do twice the work per byte -> 2
    std::atomic<long> opt_scan_batch{4};
    std::atomic<long> opt_small_batch_tensor_rows{500000};   // 2..scan_batch queries take the tensor path from this many rows on
    std::atomic<long> opt_shadow_scan_nq{1};          // ... for batches of up to this many queries (at most 2)
    std::atomic<long> opt_shadow_scan_rows{360000};  // single queries scan the fp16 shadow plane from this many rows on (measured on
                                                     // 512-d shards: 250k rows 104 vs 85 us for the fp32 scan, 500k 121 vs 154, 1M 181 vs 294:
                                                     // four launches against one)
    std::atomic<long> opt_shadow{1};        // 1 = the tensor path contracts the fp16 shadow plane (fp32 shards)
    std::atomic<long> stat_fallback{0}, stat_tensor_batches{0}, stat_scan_passes{0}, stat_shadow_scans{0};
    GemmPlan gemm_plan;
    // opt-in timing of the dominant kernel (scan or tensor) with CUDA events on the launching stream
    std::atomic<long> opt_profile{0};
    std::mutex prof_mu;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;

    size_t elem() const { return dtype == VDB_F16 ? 2 : 4; }
    size_t row_bytes() const { return (size_t)ld * elem(); }
    bool dead(size_t r) const { return (h_dead[r >> 6] >> (r & 63)) & 1; }
    void set_dead(size_t r, bool v) {
        if (v) h_dead[r >> 6] |= (1ull << (r & 63));
        else h_dead[r >> 6] &= ~(1ull << (r & 63));
    }
    // returns row or -1
    long long row_of(int64_t label) const {
        if (affine) {
            const int64_t r = label - label_base;
            return (r >= 0 && (size_t)r < count.load()) ? r : -1;
        }
        auto it = map.find(label);
        return it == map.end() ? -1 : (long long)it->second;
    }
};

namespace {

static size_t tomb_words_for(size_t cap) { return (cap + 31) / 32 + 4; }

// Back every slab for `cap` rows (maps the missing tail; the prefix and the base pointers stay).  Caller holds wmu.
int map_capacity(vdb* db, size_t cap) {
    cap = std::max(cap, (size_t)1);
    std::string err;
    if (!db->b_rows.ensure(cap * db->row_bytes(), err) ||
        (db->ld16 && !db->b_shadow.ensure(cap * (size_t)db->ld16 * 2, err)) ||
        !db->b_sqnorm.ensure(cap * sizeof(float), err) || !db->b_labels.ensure(cap * sizeof(uint32_t), err) ||
        !db->b_tomb.ensure(tomb_words_for(cap) * sizeof(uint32_t), err))
        return fail(err.find("out of memory") != std::string::npos ? VDB_ENOMEM : VDB_ECUDA, err);
    const size_t words = tomb_words_for(cap);
    if (words > db->tomb_words_zeroed) {      // freshly mapped memory is not zero
        CU_TRY(cudaMemsetAsync(db->tomb + db->tomb_words_zeroed, 0, (words - db->tomb_words_zeroed) * sizeof(uint32_t), db->wstream));
        CU_TRY(cudaStreamSynchronize(db->wstream));
        db->tomb_words_zeroed = words;
    }
    if (db->h_dead.size() < (cap + 63) / 64) db->h_dead.resize((cap + 63) / 64, 0);
    return VDB_OK;
}

// Address reservations for `va_rows` rows per slab.  First call: reserves; later (reservation exhausted): moves the
// mapped chunks to larger ranges -- the base pointers change, the caller holds `mu` exclusive and the device is idle.
int reserve_rows(vdb* db, size_t va_rows) {
    std::string err;
    const bool first = db->rows == nullptr;
    auto one = [&](GrowBuf& b, size_t bytes) { return first ? b.reserve(db->device, bytes, err) : b.rebase(bytes, err); };
    if (!one(db->b_rows, va_rows * db->row_bytes()) || (db->ld16 && !one(db->b_shadow, va_rows * (size_t)db->ld16 * 2)) ||
        !one(db->b_sqnorm, va_rows * sizeof(float)) || !one(db->b_labels, va_rows * sizeof(uint32_t)) ||
        !one(db->b_tomb, tomb_words_for(va_rows) * sizeof(uint32_t)))
        return fail(VDB_ECUDA, err);
    db->va_rows = va_rows;
    db->rows = db->b_rows.ptr();
    db->shadow = db->ld16 ? db->b_shadow.ptr() : nullptr;
    db->sqnorm = static_cast<float*>(db->b_sqnorm.ptr());
    db->labels = static_cast<uint32_t*>(db->b_labels.ptr());
    db->tomb = static_cast<uint32_t*>(db->b_tomb.ptr());
    return VDB_OK;
}

int alloc_shard(vdb* db) {
    CU_TRY(cudaSetDevice(db->device));
    CU_TRY(cudaFree(nullptr));                                   // make sure the primary context exists and is current
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, db->device));
    db->num_sms = prop.multiProcessorCount;
    if (prop.major != 10)
        return fail(VDB_ECUDA, std::string("this library is built for sm_100a (B200); device is ") + prop.name);
    CU_TRY(cudaStreamCreateWithFlags(&db->wstream, cudaStreamNonBlocking));
    const size_t cap = std::max(db->capacity.load(), (size_t)1);
    // room to grow 8x in place (address space only; at least 1M rows so that small shards never re-reserve)
    int rc = reserve_rows(db, std::max(cap * 8, (size_t)1 << 20));
    if (rc) return rc;
    rc = map_capacity(db, cap);
    if (rc) return rc;
    CU_TRY(cudaMalloc((void**)&db->d_max_sqnorm, sizeof(unsigned int)));
    CU_TRY(cudaMemset(db->d_max_sqnorm, 0, sizeof(unsigned int)));
    CU_TRY(cudaMalloc((void**)&db->d_stage, STAGE_ROWS * (size_t)db->dim * sizeof(float)));
    return VDB_OK;
}

void free_workspace(Workspace* w) {
    if (w->d_q_in) cudaFree(w->d_q_in);
    if (w->d_q) cudaFree(w->d_q);
    if (w->d_qn2) cudaFree(w->d_qn2);
    if (w->d_keys) cudaFree(w->d_keys);
    if (w->d_scan_ctr) cudaFree(w->d_scan_ctr);
    if (w->d_group_keys) cudaFree(w->d_group_keys);
    if (w->d_out) cudaFree(w->d_out);
    if (w->h_q) cudaFreeHost(w->h_q);
    if (w->h_out) cudaFreeHost(w->h_out);
    gemm_workspace_free(w->gemm);
    if (w->done) cudaEventDestroy(w->done);
    if (w->stream) cudaStreamDestroy(w->stream);
}

// `st`: the stream a device-pointer call will enqueue on (null for host-buffer calls, which run on the workspace's
// own stream and hold it until they return).  Order of preference: the workspace this stream used last (stream
// order already protects its scratch, and a caller looping on one stream keeps one warm workspace); one whose last
// device-side user has finished; a new one; any (the caller then waits for its event on the GPU).  A caller that
// enqueues on several streams gets its searches side by side instead of chained through one scratch buffer.
Workspace* acquire_ws(vdb* db, cudaStream_t st = nullptr, bool dev_call = false) {
    std::unique_lock<std::mutex> lk(db->ws_mu);
    for (;;) {
        if (dev_call)
            for (size_t i = db->ws_free.size(); i-- > 0;)
                if (db->ws_free[i]->used && db->ws_free[i]->last_stream == st) {
                    Workspace* w = db->ws_free[i];
                    db->ws_free.erase(db->ws_free.begin() + (long)i);
                    return w;
                }
        for (size_t i = db->ws_free.size(); i-- > 0;) {
            Workspace* w = db->ws_free[i];
            if (!w->used || cudaEventQuery(w->done) == cudaSuccess) {
                db->ws_free.erase(db->ws_free.begin() + (long)i);
                return w;
            }
        }
        cudaGetLastError();                                   // cudaErrorNotReady from the queries above is not an error
        if ((int)db->ws_all.size() < MAX_WORKSPACES) {
            auto w = std::make_unique<Workspace>();
            if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&w->done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
            db->ws_all.push_back(std::move(w));
            return db->ws_all.back().get();
        }
        if (!db->ws_free.empty()) {                           // all busy and the pool is full: queue behind one
            Workspace* w = db->ws_free.back();
            db->ws_free.pop_back();
            return w;
        }
        db->ws_cv.wait(lk);
    }
}
void release_ws(vdb* db, Workspace* w) {
    {
        std::lock_guard<std::mutex> lk(db->ws_mu);
        db->ws_free.push_back(w);
    }
    db->ws_cv.notify_one();
}
struct WsGuard {
    vdb* db; Workspace* w;
    ~WsGuard() { if (w) release_ws(db, w); }
};

struct ProfScope {   // records start/stop events around the dominant kernel when profiling is on
    vdb* db; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(vdb* d, cudaStream_t s) : db(d), st(s) { begin(); }
    ProfScope(vdb* d, cudaStream_t s, bool deferred) : db(d), st(s) { (void)deferred; }
    void begin() {
        if (!db->opt_profile.load()) return;
        std::lock_guard<std::mutex> lk(db->prof_mu);
        if (!db->prof_pool.empty()) { a = db->prof_pool.back().first; b = db->prof_pool.back().second; db->prof_pool.pop_back(); }
        else { cudaEventCreate(&a); cudaEventCreate(&b); }
        cudaEventRecord(a, st);
    }
    void end() {
        if (!a) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(db->prof_mu);
        db->prof_events.emplace_back(a, b);
        a = b = nullptr;
    }
    ~ProfScope() { end(); }
};
static void prof_begin_cb(void* ctx, cudaStream_t) { static_cast<ProfScope*>(ctx)->begin(); }
static void prof_end_cb(void* ctx, cudaStream_t) { static_cast<ProfScope*>(ctx)->end(); }

// Queries one scan pass takes: up to 8, fewer when the rows are so long that 8 query vectors + k-lists no longer fit
// beside the copy ring in shared memory (dim > 1152 fp32 / 1792 fp16).  0: k does not fit even for one query.
int scan_group(const vdb* db, int k, size_t nq) {
    int g = nq > 4 ? 8 : nq > 2 ? 4 : nq > 1 ? 2 : 1;      // one pass when the batch fits a kernel variant
    while (g > 1 && scan_max_k(g, db->ld, (uint32_t)db->row_bytes()) < k) g >>= 1;
    return scan_max_k(g, db->ld, (uint32_t)db->row_bytes()) >= k ? g : 0;
}

// Exact scan (K1 + K5) of nq PREPARED queries [nq][ld]; groups of up to 8 queries per pass.
int scan_prepared(vdb* db, Workspace* ws, const float* d_qp, size_t nq, int k, int64_t* d_ids, float* d_dist,
                  int* d_cnt, cudaStream_t st, size_t n, bool raw = false) {
    const bool f16 = db->dtype == VDB_F16;
    ScanParams sp{};
    sp.rows = db->rows;
    sp.row_bytes = (uint32_t)db->row_bytes();
    sp.ld = db->ld;
    sp.n_rows = (uint32_t)n;
    // a search's snapshot [0, n) is affine if the shard is when the search starts: rows that break the pattern are
    // appended behind it
    const bool affine = db->affine.load();
    sp.labels = affine ? nullptr : db->labels;
    sp.label_base = affine ? (uint32_t)db->label_base : 0u;
    sp.tomb = db->any_dead.load() ? db->tomb : nullptr;
    sp.k = k;
    sp.metric = db->metric == VDB_L2 ? 0 : 1;
    // every pass of this call uses the launch shape of the widest pass
    const int group = scan_group(db, k, nq);
    if (group == 0) return fail(VDB_EINVAL, "k too large for the scan kernel at this dimension");
    const ScanPlan pl = scan_plan(group, (uint32_t)db->ld, sp.row_bytes, k, (uint32_t)n, db->num_sms);
    if (pl.grid == 0) return fail(VDB_EINVAL, "k too large for the scan kernel at this dimension");
    const int grid = pl.grid;
    CU_TRY(grow(ws->d_keys, ws->keys_cap, nq * (size_t)grid * k));
    const bool fused_merge = pl.merge_group > 0 && nq <= 2;     // one pass whose last CTAs also merge the lists
    if (fused_merge) {
        if (!ws->d_scan_ctr) {
            CU_TRY(cudaMalloc((void**)&ws->d_scan_ctr, (SCAN_MERGE_KEYS + 1) * sizeof(unsigned int)));
            CU_TRY(cudaMemsetAsync(ws->d_scan_ctr, 0, (SCAN_MERGE_KEYS + 1) * sizeof(unsigned int), st));
            CU_TRY(cudaMalloc((void**)&ws->d_group_keys, 2 * (size_t)SCAN_MERGE_KEYS * sizeof(uint64_t)));
        }
        sp.merge_group = pl.merge_group;
        sp.merge_ctr = ws->d_scan_ctr;
        sp.group_keys = ws->d_group_keys;
        sp.out_ids = d_ids; sp.out_dist = d_dist; sp.out_counts = d_cnt;
    }
    for (size_t g = 0; g < nq; g += (size_t)group) {
        sp.nq = (int)std::min<size_t>((size_t)group, nq - g);
        if (raw) {   // queries as the caller gave them: normalised / padded inside the kernel
            sp.q_raw = d_qp + g * (size_t)db->dim;
            sp.dim = db->dim;
            sp.normalize = db->metric == VDB_COSINE ? 1 : 0;
        } else {
            sp.q = d_qp + g * (size_t)db->ld;
        }
        sp.out_keys = ws->d_keys + g * (size_t)grid * k;
        {
            ProfScope prof(db, st);
            CU_TRY(launch_scan_topk(sp, f16, pl, st));
        }
        db->stat_scan_passes.fetch_add(1);
    }
    if (fused_merge) return VDB_OK;
    MergeParams mp{};
    mp.nq = nq;
    mp.k_out = k;
    mp.out_ids = d_ids;
    mp.out_dist = d_dist;
    mp.out_counts = d_cnt;
    mp.in_keys = ws->d_keys;
    mp.n_in = grid * k;
    CU_TRY(launch_merge_topk(mp, st));
    return VDB_OK;
}

// One or two queries against an fp32 shard that has an fp16 shadow plane: K1 streams the SHADOW (half the bytes of
// the rows: the search is HBM-bound) and keeps the k' best approximate candidates per query (in-kernel merge), K4w
// recomputes the few that can still matter from the fp32 rows and certifies that no other row can enter the top-k
// (same rigorous error bound as the batched path), K4x re-searches exactly the rare query whose certificate fails.
// Results are bit-identical to the fp32 scan.  VDB_ENOTSUP_INTERNAL: not applicable here, take the fp32 scan.
constexpr int VDB_ENOTSUP_INTERNAL = -1000;
int shadow_scan(vdb* db, Workspace* ws, const float* d_q_raw, size_t nq, int k, int64_t* d_ids, float* d_dist,
                int* d_cnt, cudaStream_t st, size_t n) {
    if (db->dtype == VDB_F16 || !db->shadow || !db->opt_shadow.load() || nq > (size_t)db->opt_shadow_scan_nq.load()) return VDB_ENOTSUP_INTERNAL;
    if ((long)n < db->opt_shadow_scan_rows.load()) return VDB_ENOTSUP_INTERNAL;      // small shards: launches dominate
    const uint32_t row_bytes16 = (uint32_t)db->ld16 * 2;
    const int kp = gemm_topk_candidate_kp(k);
    if (kp == 0 || row_bytes16 % 512 != 0 || n <= (size_t)kp) return VDB_ENOTSUP_INTERNAL;
    const int group = nq > 1 ? 2 : 1;
    if (scan_max_k(group, db->ld16, row_bytes16) < kp) return VDB_ENOTSUP_INTERNAL;
    const ScanPlan pl = scan_plan(group, (uint32_t)db->ld16, row_bytes16, kp, (uint32_t)n, db->num_sms);
    if (pl.grid == 0 || pl.merge_group == 0) return VDB_ENOTSUP_INTERNAL;

    CU_TRY(grow(ws->d_q, ws->q_cap, nq * (size_t)db->ld));
    CU_TRY(grow(ws->d_qn2, ws->qn2_cap, nq));
    GemmSearchArgs a{};
    a.rows = db->rows; a.ld = db->ld; a.dim = db->dim; a.f16 = false; a.n_rows = (uint32_t)n;
    a.sqnorm = db->sqnorm; a.labels = db->labels; a.tomb = db->any_dead.load() ? db->tomb : nullptr;
    a.q = ws->d_q; a.qn2 = ws->d_qn2; a.nq = nq; a.k = k;
    a.metric = db->metric == VDB_L2 ? 0 : 1;
    a.d_max_sqnorm_bits = db->d_max_sqnorm;
    a.num_sms = db->num_sms;
    a.out_ids = d_ids; a.out_dist = d_dist; a.out_counts = d_cnt;
    GemmPrepTargets pt;
    CU_TRY(gemm_topk_prep_targets(ws->gemm, a, &pt));                 // no shadow in `a`: no fp16 query plane wanted
    CandidateBuffers cb;
    CU_TRY(gemm_topk_candidate_buffers(ws->gemm, nq, kp, &cb));
    CU_TRY(launch_prepare_queries(d_q_raw, nq, db->dim, db->ld, db->metric == VDB_COSINE, ws->d_q, ws->d_qn2, st, nullptr, 0,
                                  pt.overflow, pt.n_flagged));

    ScanParams sp{};
    sp.rows = db->shadow;
    sp.row_bytes = row_bytes16;
    sp.ld = (uint32_t)db->ld16;
    sp.n_rows = (uint32_t)n;
    sp.labels = nullptr; sp.label_base = 0;                            // candidates carry ROW numbers: K4w maps them to labels
    sp.tomb = a.tomb;
    sp.k = kp;
    sp.metric = a.metric;
    sp.nq = (int)nq;
    sp.q_raw = d_q_raw; sp.dim = db->dim; sp.normalize = db->metric == VDB_COSINE ? 1 : 0;
    CU_TRY(grow(ws->d_keys, ws->keys_cap, nq * (size_t)pl.grid * kp));
    if (!ws->d_scan_ctr) {
        CU_TRY(cudaMalloc((void**)&ws->d_scan_ctr, (SCAN_MERGE_KEYS + 1) * sizeof(unsigned int)));
        CU_TRY(cudaMemsetAsync(ws->d_scan_ctr, 0, (SCAN_MERGE_KEYS + 1) * sizeof(unsigned int), st));
        CU_TRY(cudaMalloc((void**)&ws->d_group_keys, 2 * (size_t)SCAN_MERGE_KEYS * sizeof(uint64_t)));
    }
    sp.out_keys = ws->d_keys;
    sp.merge_group = pl.merge_group; sp.merge_ctr = ws->d_scan_ctr; sp.group_keys = ws->d_group_keys;
    sp.cand_keys = cb.keys; sp.cand_stride = cb.stride; sp.cand_cnt = cb.cnt; sp.cand_tau = cb.tau;
    {
        ProfScope prof(db, st);
        CU_TRY(launch_scan_topk(sp, /*f16=*/true, pl, st));
    }
    db->stat_scan_passes.fetch_add(1);
    db->stat_shadow_scans.fetch_add(1);
    CU_TRY(gemm_topk_rerank_candidates(ws->gemm, a, kp, st));
    return VDB_OK;
}

// Enqueue a search of nq device-resident raw queries on `st`.  Outputs are device pointers.  Nothing here waits for
// the device: a query whose tensor-path certificate fails is re-searched exactly by a kernel of the same enqueue.
int search_core(vdb* db, Workspace* ws, const float* d_q_raw, size_t nq, int k, int64_t* d_ids, float* d_dist,
                int* d_cnt, cudaStream_t st, size_t n) {
    const bool f16 = db->dtype == VDB_F16;
    MergeParams mp{};
    mp.nq = nq;
    mp.k_out = k;
    mp.out_ids = d_ids;
    mp.out_dist = d_dist;
    mp.out_counts = d_cnt;
    if (n == 0) {  // empty shard: all padding (reference: success + empty lists, handler.py:353-354)
        CU_TRY(grow(ws->d_keys, ws->keys_cap, 16));
        mp.in_keys = ws->d_keys;
        mp.n_in = 0;
        CU_TRY(launch_merge_topk(mp, st));
        return VDB_OK;
    }
    const long path = db->opt_path.load();
    bool tensor = false;
    if (path == 2) tensor = true;
    else if (path == 0) {
        tensor = (long)nq > db->opt_scan_batch.load();
        // two to four queries against a large fp32 shard with a shadow plane: the tensor path streams half the bytes
        // (1M x 512: 235 us against 300 - 340 us for the fp32 scan; its six launches lose below ~0.5M rows)
        if (!tensor && nq >= 2 && !f16 && db->shadow && db->opt_shadow.load() && (long)n >= db->opt_small_batch_tensor_rows.load())
            tensor = true;
    }
    if (tensor && !gemm_topk_supported(db->dim, db->ld, f16, k, n)) {
        if (path == 2) return fail(VDB_EINVAL, "tensor path forced but unsupported for this shape/k");
        tensor = false;
    }
    if (!tensor && path == 0) {
        const int rc = shadow_scan(db, ws, d_q_raw, nq, k, d_ids, d_dist, d_cnt, st, n);
        if (rc != VDB_ENOTSUP_INTERNAL) return rc;
    }
    if (!tensor) return scan_prepared(db, ws, d_q_raw, nq, k, d_ids, d_dist, d_cnt, st, n, /*raw=*/true);

    CU_TRY(grow(ws->d_q, ws->q_cap, nq * (size_t)db->ld));
    CU_TRY(grow(ws->d_qn2, ws->qn2_cap, nq));
    if (tensor) {
        GemmSearchArgs a{};
        a.rows = db->rows; a.ld = db->ld; a.dim = db->dim; a.f16 = f16; a.n_rows = (uint32_t)n;
        if (db->shadow && db->opt_shadow.load()) { a.shadow = db->shadow; a.ld16 = db->ld16; }
        a.sqnorm = db->sqnorm; a.labels = db->labels; a.tomb = db->any_dead.load() ? db->tomb : nullptr;
        a.q = ws->d_q; a.qn2 = ws->d_qn2; a.nq = nq; a.k = k;
        a.metric = db->metric == VDB_L2 ? 0 : 1;
        a.d_max_sqnorm_bits = db->d_max_sqnorm;
        a.num_sms = db->num_sms;
        a.out_ids = d_ids; a.out_dist = d_dist; a.out_counts = d_cnt;
        // one kernel prepares the queries (pad / normalise / norms), writes their fp16 operand copy and clears
        // the per-batch flags of the tensor path
        GemmPrepTargets pt;
        CU_TRY(gemm_topk_prep_targets(ws->gemm, a, &pt));
        CU_TRY(launch_prepare_queries(d_q_raw, nq, db->dim, db->ld, db->metric == VDB_COSINE, ws->d_q, ws->d_qn2, st,
                                      pt.q16, pt.gld, pt.overflow, pt.n_flagged));
        a.prepped = true;
        std::string err;
        ProfScope prof(db, st, true);       // one event pair per launch of the tensor-core kernel
        a.prof_begin = prof_begin_cb; a.prof_end = prof_end_cb; a.prof_ctx = &prof;
        cudaError_t e = gemm_topk_search(db->gemm_plan, ws->gemm, a, st, err);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(VDB_ECUDA, "gemm_topk_search: " + (err.empty() ? std::string(cudaGetErrorString(e)) : err));
        }
        db->stat_tensor_batches.fetch_add(1);
        return VDB_OK;
    }

    return fail(VDB_ECUDA, "internal: unreachable search path");
}

}  // namespace
namespace vdbk {
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("VDB_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
}  // namespace vdbk
namespace {

bool is_pinned_host(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int check_k(const vdb* db, int k, size_t nq) {
    if (k < 1 || k > K_MAX) return fail(VDB_EINVAL, "k must be in [1, 1024]");
    if (scan_group(db, k, nq) == 0) return fail(VDB_EINVAL, "k too large for this dim");
    return VDB_OK;
}

}  // namespace

extern "C" {

const char* vdb_last_error(void) { return t_err.c_str(); }
const char* vdb_version(void) { return "vdb_b200 0.1 (sm_100a)"; }
uint64_t vdb_launch_count(void) { return launch_count(); }

int vdb_debug_level_plan(size_t nq, size_t n_rows, int k, int* out, int out_len) {
    if (nq < 1 || n_rows < 1 || k < 1 || k > 128 || n_rows > 0xFFFFFFFFull) return fail(VDB_EINVAL, "bad argument");
    const LevelPlan lp = gemm_topk_level_plan(nq, n_rows, k);
    std::vector<int> v = {lp.kp, lp.kq, lp.cap, lp.growth, lp.query_blocks, lp.n_tiles, lp.n_pos, lp.probe_tiles,
                          lp.probe_rank, (int)lp.levels.size()};
    for (const LevelPlan::Level& lv : lp.levels) { v.push_back(lv.p0); v.push_back(lv.p1); v.push_back(lv.rank_after); }
    for (int i = 0; i < (int)v.size() && i < out_len; ++i)
        if (out) out[i] = v[i];
    return (int)v.size();
}

int vdb_create(int dim, int metric, int store_dtype, size_t capacity, int device, vdb_t** out) {
    if (!out) return fail(VDB_EINVAL, "out is null");
    *out = nullptr;
    if (dim < 1 || dim > 8192) return fail(VDB_EINVAL, "dim out of range");
    if (metric < VDB_L2 || metric > VDB_COSINE) return fail(VDB_EINVAL, "metric must be 0 (l2), 1 (ip) or 2 (cosine)");
    if (store_dtype != VDB_F32 && store_dtype != VDB_F16) return fail(VDB_EINVAL, "store_dtype must be 0 (f32) or 1 (f16)");
    if (capacity >= 0xFFFFFFF0ull) return fail(VDB_EINVAL, "capacity must be below 2^32 rows per shard");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(VDB_ECUDA, "no CUDA device: this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(VDB_EINVAL, "device index out of range");
    auto db = std::make_unique<vdb>();
    db->dim = dim;
    db->metric = metric;
    db->dtype = store_dtype;
    db->device = device;
    db->capacity = capacity;
    db->opt_scan_batch.store(store_dtype == VDB_F16 ? 2 : 4);
    const int unit = store_dtype == VDB_F16 ? 256 : 128;   // one 512-byte warp load
    db->ld = (dim + unit - 1) / unit * unit;
    if (scan_max_k(1, db->ld, (uint32_t)db->row_bytes()) < 1)
        return fail(VDB_EINVAL, "dim too large: a shard row must fit the scan kernel's shared-memory ring (dim <= 1408 for "
                                "fp32 rows, <= 2816 for fp16 rows)");
    // fp32 shards keep an fp16 shadow plane for the batched tensor path (+50 % HBM); VDB_SHADOW=0 opts out
    const char* sh = getenv("VDB_SHADOW");
    if (store_dtype == VDB_F32 && !(sh && sh[0] == '0')) db->ld16 = (dim + 63) / 64 * 64;
    int rc = alloc_shard(db.get());
    if (rc != VDB_OK) {
        vdb_destroy(db.release());
        return rc;
    }
    *out = db.release();
    return VDB_OK;
}

void vdb_destroy(vdb_t* db) {
    if (!db) return;
    cudaSetDevice(db->device);
    cudaDeviceSynchronize();
    for (auto& w : db->ws_all) free_workspace(w.get());
    for (auto& ev : db->prof_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto& ev : db->prof_pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    gemm_plan_free(db->gemm_plan);
    // the slabs (GrowBuf members) unmap and release themselves
    if (db->d_max_sqnorm) cudaFree(db->d_max_sqnorm);
    if (db->d_stage) cudaFree(db->d_stage);
    if (db->d_idx) cudaFree(db->d_idx);
    if (db->wstream) cudaStreamDestroy(db->wstream);
    delete db;
}

size_t vdb_count(const vdb_t* db) { return db ? db->count.load() : 0; }
size_t vdb_live_count(const vdb_t* db) { return db ? db->live.load() : 0; }
size_t vdb_capacity(const vdb_t* db) { return db ? db->capacity.load() : 0; }
int vdb_dim(const vdb_t* db) { return db ? db->dim : 0; }

// caller holds wmu.  Tombstones rows on host mirror + device (searches in flight see the old or the new bit).
static int tombstone_rows(vdb* db, const std::vector<uint32_t>& rows, bool set) {
    if (rows.empty()) return VDB_OK;
    CU_TRY(grow(db->d_idx, db->idx_cap, rows.size()));
    CU_TRY(cudaMemcpyAsync(db->d_idx, rows.data(), rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, db->wstream));
    CU_TRY(launch_set_bits(db->tomb, db->d_idx, rows.size(), set, db->wstream));
    CU_TRY(cudaStreamSynchronize(db->wstream));
    for (uint32_t r : rows) db->set_dead(r, set);
    if (set) db->any_dead.store(true);
    return VDB_OK;
}

// caller holds wmu.  Registers labels for rows [row0, row0+n) (not yet visible to searches) and uploads them.
static int register_labels(vdb* db, const int64_t* labels, size_t n, size_t row0) {
    for (size_t i = 0; i < n; ++i)
        if (labels[i] < 0 || labels[i] > (int64_t)LABEL_MAX) return fail(VDB_EINVAL, "labels must lie in [0, 2^32-2]");
    // existing live label => tombstone the old row (hnswlib would update in place)
    std::vector<uint32_t> kill;
    for (size_t i = 0; i < n; ++i) {
        long long r = db->row_of(labels[i]);
        if (r >= 0 && !db->dead((size_t)r)) kill.push_back((uint32_t)r);
    }
    // duplicates inside this batch: the last one wins
    std::unordered_map<int64_t, size_t> last;
    bool dup = false;
    if (n > 1) {
        last.reserve(n * 2);
        for (size_t i = 0; i < n; ++i) {
            auto it = last.find(labels[i]);
            if (it != last.end()) { kill.push_back((uint32_t)(row0 + it->second)); it->second = i; dup = true; }
            else last.emplace(labels[i], i);
        }
    }
    // does the batch keep label == base + row ?
    bool keeps_affine = db->affine && !dup;
    if (keeps_affine) {
        if (row0 == 0) db->label_base = labels[0];
        for (size_t i = 0; i < n && keeps_affine; ++i) keeps_affine = labels[i] == db->label_base + (int64_t)(row0 + i);
    }
    if (db->affine && !keeps_affine) {   // materialise the map for what is already there
        db->map.reserve((row0 + n) * 2);
        for (size_t r = 0; r < row0; ++r) db->map[db->label_base + (int64_t)r] = (uint32_t)r;
        db->affine = false;
    }
    if (!db->affine)
        for (size_t i = 0; i < n; ++i) db->map[labels[i]] = (uint32_t)(row0 + i);
    std::vector<uint32_t> l32(n);
    for (size_t i = 0; i < n; ++i) l32[i] = (uint32_t)labels[i];
    CU_TRY(cudaMemcpyAsync(db->labels + row0, l32.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, db->wstream));
    CU_TRY(cudaStreamSynchronize(db->wstream));
    if (!kill.empty()) {
        std::sort(kill.begin(), kill.end());
        kill.erase(std::unique(kill.begin(), kill.end()), kill.end());
        // rows killed inside this batch are not yet counted live
        int rc = tombstone_rows(db, kill, true);
        if (rc) return rc;
        db->live.fetch_sub(kill.size());   // in-batch duplicates are compensated by the caller's += n
    }
    return VDB_OK;
}

static int add_impl(vdb* db, const float* rows, bool rows_on_device, const int64_t* labels, size_t n, cudaStream_t user) {
    if (!db) return fail(VDB_EINVAL, "db is null");
    if (n == 0) return VDB_OK;
    if (!rows || !labels) return fail(VDB_EINVAL, "rows/labels is null");
    std::lock_guard<std::mutex> wlk(db->wmu);                 // one writer at a time ...
    std::shared_lock<std::shared_mutex> lk(db->mu);           // ... next to any number of searches
    CU_TRY(cudaSetDevice(db->device));
    const size_t row0 = db->count.load();
    if (row0 + n > db->capacity.load())
        return fail(VDB_EFULL, "The number of elements exceeds the specified limit");   // hnswlib's message
    if (rows_on_device && user) CU_TRY(cudaStreamSynchronize(user));
    int rc = register_labels(db, labels, n, row0);
    if (rc) return rc;
    for (size_t off = 0; off < n; off += STAGE_ROWS) {
        const size_t m = std::min(STAGE_ROWS, n - off);
        const float* src = rows + off * (size_t)db->dim;
        if (!rows_on_device) {
            CU_TRY(cudaMemcpyAsync(db->d_stage, src, m * (size_t)db->dim * sizeof(float), cudaMemcpyHostToDevice, db->wstream));
            src = db->d_stage;
        }
        CU_TRY(launch_insert_rows(src, m, db->dim, db->ld, db->metric == VDB_COSINE, db->dtype == VDB_F16, db->rows,
                                  db->sqnorm, row0 + off, db->d_max_sqnorm, db->wstream));
        if (db->shadow)
            CU_TRY(launch_shadow_rows((const float*)db->rows, db->ld, db->shadow, db->ld16, row0 + off, m, db->wstream));
        if (!rows_on_device) CU_TRY(cudaStreamSynchronize(db->wstream));   // staging buffer is reused
    }
    CU_TRY(cudaStreamSynchronize(db->wstream));
    db->live.fetch_add(n);
    db->count.store(row0 + n, std::memory_order_release);     // the rows are complete: searches may see them
    return VDB_OK;
}

int vdb_add(vdb_t* db, const float* rows, const int64_t* labels, size_t n) { return add_impl(db, rows, false, labels, n, nullptr); }
int vdb_add_dev(vdb_t* db, const float* d_rows, const int64_t* h_labels, size_t n, void* stream) {
    return add_impl(db, d_rows, true, h_labels, n, (cudaStream_t)stream);
}

int vdb_synth_dev(uint64_t seed, uint64_t row_start, size_t n, int dim, float* d_out, void* stream) {
    if (!d_out || dim < 1) return fail(VDB_EINVAL, "bad argument");
    CU_TRY(launch_synth_rows(seed, row_start, n, dim, d_out, (cudaStream_t)stream));
    return VDB_OK;
}

int vdb_add_synthetic(vdb_t* db, uint64_t seed, uint64_t row_start, size_t n, int64_t label_start) {
    if (!db) return fail(VDB_EINVAL, "db is null");
    if (n == 0) return VDB_OK;
    std::lock_guard<std::mutex> wlk(db->wmu);
    std::shared_lock<std::shared_mutex> lk(db->mu);
    CU_TRY(cudaSetDevice(db->device));
    const size_t row0 = db->count.load();
    if (row0 + n > db->capacity.load()) return fail(VDB_EFULL, "The number of elements exceeds the specified limit");
    if (label_start < 0 || label_start + (int64_t)n - 1 > (int64_t)LABEL_MAX) return fail(VDB_EINVAL, "labels must lie in [0, 2^32-2]");
    // labels are label_start + i: keep the affine map when possible, else fall back to the generic path
    const bool keeps_affine = db->affine && (row0 == 0 || label_start == db->label_base + (int64_t)row0);
    if (!keeps_affine) {
        std::vector<int64_t> l(n);
        for (size_t i = 0; i < n; ++i) l[i] = label_start + (int64_t)i;
        int rc = register_labels(db, l.data(), n, row0);
        if (rc) return rc;
    } else {
        if (row0 == 0) db->label_base = label_start;
        CU_TRY(launch_iota_u32(db->labels + row0, n, (uint32_t)label_start, db->wstream));
    }
    for (size_t off = 0; off < n; off += STAGE_ROWS) {
        const size_t m = std::min(STAGE_ROWS, n - off);
        CU_TRY(launch_synth_rows(seed, row_start + off, m, db->dim, db->d_stage, db->wstream));
        CU_TRY(launch_insert_rows(db->d_stage, m, db->dim, db->ld, db->metric == VDB_COSINE, db->dtype == VDB_F16,
                                  db->rows, db->sqnorm, row0 + off, db->d_max_sqnorm, db->wstream));
        if (db->shadow)
            CU_TRY(launch_shadow_rows((const float*)db->rows, db->ld, db->shadow, db->ld16, row0 + off, m, db->wstream));
    }
    CU_TRY(cudaStreamSynchronize(db->wstream));
    db->live.fetch_add(n);
    db->count.store(row0 + n, std::memory_order_release);
    return VDB_OK;
}

static int mark_impl(vdb* db, const int64_t* labels, size_t n, bool set) {
    if (!db) return fail(VDB_EINVAL, "db is null");
    if (n == 0) return VDB_OK;
    if (!labels) return fail(VDB_EINVAL, "labels is null");
    std::lock_guard<std::mutex> wlk(db->wmu);
    std::shared_lock<std::shared_mutex> lk(db->mu);
    CU_TRY(cudaSetDevice(db->device));
    std::vector<uint32_t> rows;
    rows.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        long long r = db->row_of(labels[i]);
        if (r < 0) return fail(VDB_ENOTFOUND, "Label not found");   // hnswlib's message
        if (db->dead((size_t)r) != set) rows.push_back((uint32_t)r);
    }
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    int rc = tombstone_rows(db, rows, set);
    if (rc) return rc;
    if (set) db->live.fetch_sub(rows.size());
    else db->live.fetch_add(rows.size());
    return VDB_OK;
}
int vdb_mark_deleted(vdb_t* db, const int64_t* labels, size_t n) { return mark_impl(db, labels, n, true); }
int vdb_unmark_deleted(vdb_t* db, const int64_t* labels, size_t n) { return mark_impl(db, labels, n, false); }

// A host-buffer search in two halves: submit enqueues upload + search + download on a workspace's own stream and
// returns; collect waits for it.  vdb_search is submit + collect; a caller that keeps two tickets open has the upload
// of batch i+1 (and, on the SMs, the short launches of its search) behind the kernels of batch i.
struct vdb_ticket {
    vdb* db = nullptr;
    Workspace* ws = nullptr;
    size_t nq = 0, nout = 0;
    bool out_pinned = false;
    int64_t* out_labels = nullptr; float* out_dist = nullptr; int* out_counts = nullptr;
};

int vdb_search_submit(vdb_t* db, const float* queries, size_t nq, int k, int64_t* out_labels, float* out_dist, int* out_counts,
                      vdb_ticket_t** ticket) {
    if (!ticket) return fail(VDB_EINVAL, "ticket is null");
    *ticket = nullptr;
    if (!db) return fail(VDB_EINVAL, "db is null");
    if (nq == 0) { *ticket = new vdb_ticket(); return VDB_OK; }
    if (!queries || !out_labels || !out_dist) return fail(VDB_EINVAL, "null buffer");
    int rc = check_k(db, k, nq);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> lk(db->mu);
    CU_TRY(cudaSetDevice(db->device));
    const size_t n = db->count.load(std::memory_order_acquire);     // snapshot: rows appended later are not looked at
    Workspace* ws = acquire_ws(db);
    if (!ws) return fail(VDB_ECUDA, "cannot create a search workspace (stream)");
    WsGuard guard{db, ws};
    cudaStream_t st = ws->stream;
    // a vdb_search_dev call may have returned this workspace to the pool while its kernels were still running on
    // the CALLER's stream: order this call's use of the scratch after them on the GPU
    if (ws->used) {
        CU_TRY(cudaStreamWaitEvent(st, ws->done, 0));
        ws->used = false;          // the ticket holds the workspace until collect has synchronised `st`
    }
    const size_t qelems = nq * (size_t)db->dim, nout = nq * (size_t)k;
    CU_TRY(grow(ws->d_q_in, ws->q_in_cap, qelems));
    CU_TRY(grow_host(ws->h_q, ws->h_q_cap, qelems));
    const size_t out_bytes = nout * (sizeof(int64_t) + sizeof(float)) + nq * sizeof(int);
    CU_TRY(grow(ws->d_out, ws->out_cap, out_bytes));
    CU_TRY(grow_host(ws->h_out, ws->h_out_cap, out_bytes));
    int64_t* d_ids = reinterpret_cast<int64_t*>(ws->d_out);
    float* d_dist = reinterpret_cast<float*>(ws->d_out + nout * sizeof(int64_t));
    int* d_cnt = reinterpret_cast<int*>(ws->d_out + nout * (sizeof(int64_t) + sizeof(float)));
    // caller buffers that are already page-locked (vdb_host_alloc, cudaHostAlloc/Register, torch pin_memory) are
    // DMA'd directly; pageable ones go through the workspace's pinned staging buffers
    const bool q_pinned = qelems * sizeof(float) > 16384 && is_pinned_host(queries);   // a few rows: staged, no attribute query
    // results: page-locked caller buffers take the DMA directly when they are large; small results (a single query: 124
    // bytes in three arrays) come back as ONE copy of the contiguous device block into the pinned staging buffer and are
    // copied out by the CPU in collect -- two enqueues and two copy-engine round trips less per request
    const bool out_pinned = out_bytes > 16384 && is_pinned_host(out_labels) && is_pinned_host(out_dist) &&
                            (!out_counts || is_pinned_host(out_counts));
    if (!q_pinned) memcpy(ws->h_q, queries, qelems * sizeof(float));
    CU_TRY(cudaMemcpyAsync(ws->d_q_in, q_pinned ? queries : ws->h_q, qelems * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = search_core(db, ws, ws->d_q_in, nq, k, d_ids, d_dist, d_cnt, st, n);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    int64_t* h_ids = reinterpret_cast<int64_t*>(ws->h_out);
    float* h_dist = reinterpret_cast<float*>(ws->h_out + nout * sizeof(int64_t));
    int* h_cnt = reinterpret_cast<int*>(ws->h_out + nout * (sizeof(int64_t) + sizeof(float)));
    if (out_pinned) { h_ids = out_labels; h_dist = out_dist; h_cnt = out_counts; }
    cudaError_t ce;
    if (!out_pinned) {
        ce = cudaMemcpyAsync(ws->h_out, ws->d_out, out_bytes, cudaMemcpyDeviceToHost, st);
    } else {
        ce = cudaMemcpyAsync(h_ids, d_ids, nout * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_dist, d_dist, nout * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess && h_cnt) ce = cudaMemcpyAsync(h_cnt, d_cnt, nq * sizeof(int), cudaMemcpyDeviceToHost, st);
    }
    if (ce != cudaSuccess) { cudaStreamSynchronize(st); return fail(VDB_ECUDA, cudaGetErrorString(ce)); }
    auto* t = new vdb_ticket();
    t->db = db; t->ws = ws; t->nq = nq; t->nout = nout; t->out_pinned = out_pinned;
    t->out_labels = out_labels; t->out_dist = out_dist; t->out_counts = out_counts;
    guard.w = nullptr;             // the ticket owns the workspace now
    *ticket = t;
    return VDB_OK;
}

int vdb_search_collect(vdb_ticket_t* t) {
    if (!t) return fail(VDB_EINVAL, "ticket is null");
    std::unique_ptr<vdb_ticket> own(t);
    if (!t->ws) return VDB_OK;     // empty batch
    WsGuard guard{t->db, t->ws};
    cudaSetDevice(t->db->device);
    CU_TRY(cudaStreamSynchronize(t->ws->stream));
    if (!t->out_pinned) {
        const uint8_t* h = t->ws->h_out;
        memcpy(t->out_labels, h, t->nout * sizeof(int64_t));
        memcpy(t->out_dist, h + t->nout * sizeof(int64_t), t->nout * sizeof(float));
        if (t->out_counts) memcpy(t->out_counts, h + t->nout * (sizeof(int64_t) + sizeof(float)), t->nq * sizeof(int));
    }
    return VDB_OK;
}

int vdb_search(vdb_t* db, const float* queries, size_t nq, int k, int64_t* out_labels, float* out_dist, int* out_counts) {
    if (!db) return fail(VDB_EINVAL, "db is null");
    if (nq == 0) return VDB_OK;
    vdb_ticket_t* t = nullptr;
    int rc = vdb_search_submit(db, queries, nq, k, out_labels, out_dist, out_counts, &t);
    return rc ? rc : vdb_search_collect(t);
}

void* vdb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        fail(VDB_ENOMEM, "cudaMallocHost failed");
        return nullptr;
    }
    return p;
}
void vdb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int vdb_search_dev(vdb_t* db, const float* d_queries, size_t nq, int k, int64_t* d_labels, float* d_dist, int* d_counts,
                   void* stream) {
    if (!db) return fail(VDB_EINVAL, "db is null");
    if (nq == 0) return VDB_OK;
    if (!d_queries || !d_labels || !d_dist) return fail(VDB_EINVAL, "null buffer");
    int rc = check_k(db, k, nq);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> lk(db->mu);
    CU_TRY(cudaSetDevice(db->device));
    const size_t n = db->count.load(std::memory_order_acquire);
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = acquire_ws(db, st, true);
    if (!ws) return fail(VDB_ECUDA, "cannot create a search workspace (stream)");
    WsGuard guard{db, ws};
    // the scratch may still be in use by a call on ANOTHER stream: order after it on the GPU (the same stream orders itself)
    if (ws->used && ws->last_stream != st) CU_TRY(cudaStreamWaitEvent(st, ws->done, 0));
    rc = search_core(db, ws, d_queries, nq, k, d_labels, d_dist, d_counts, st, n);
    CU_TRY(cudaEventRecord(ws->done, st));
    ws->used = true;
    ws->last_stream = st;
    return rc;
}

int vdb_resize(vdb_t* db, size_t new_capacity) {
    if (!db) return fail(VDB_EINVAL, "db is null");
    std::lock_guard<std::mutex> wlk(db->wmu);
    if (new_capacity < db->count.load()) return fail(VDB_EINVAL, "new capacity below current count");
    if (new_capacity >= 0xFFFFFFF0ull) return fail(VDB_EINVAL, "capacity must be below 2^32 rows per shard");
    if (new_capacity > db->va_rows) {
        // the address reservations are exhausted: move the mapped chunks to larger ranges (no copy).  The base
        // pointers change, so this one step keeps searches out and waits for the device.
        std::unique_lock<std::shared_mutex> lk(db->mu);
        CU_TRY(cudaSetDevice(db->device));
        CU_TRY(cudaDeviceSynchronize());
        int rc = reserve_rows(db, std::max(new_capacity, db->va_rows * 8));
        if (rc) return rc;
    }
    // growing maps more physical memory behind the same pointers: searches keep running
    std::shared_lock<std::shared_mutex> lk(db->mu);
    CU_TRY(cudaSetDevice(db->device));
    int rc = map_capacity(db, new_capacity);
    if (rc) return rc;
    db->capacity.store(new_capacity);      // shrinking (>= count) lowers the limit; the memory stays mapped
    return VDB_OK;
}

int vdb_get_rows(vdb_t* db, const int64_t* labels, size_t n, float* out) {
    if (!db) return fail(VDB_EINVAL, "db is null");
    if (n == 0) return VDB_OK;
    if (!labels || !out) return fail(VDB_EINVAL, "null buffer");
    std::lock_guard<std::mutex> wlk(db->wmu);          // uses the writer stream + d_idx + the label map
    std::shared_lock<std::shared_mutex> lk(db->mu);
    CU_TRY(cudaSetDevice(db->device));
    std::vector<uint32_t> rows(n);
    for (size_t i = 0; i < n; ++i) {
        long long r = db->row_of(labels[i]);
        if (r < 0 || db->dead((size_t)r)) return fail(VDB_ENOTFOUND, "Label not found");
        rows[i] = (uint32_t)r;
    }
    float* d_out = nullptr;
    CU_TRY(grow(db->d_idx, db->idx_cap, n));
    CU_TRY(cudaMalloc((void**)&d_out, n * (size_t)db->dim * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(db->d_idx, rows.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, db->wstream);
    if (e == cudaSuccess) e = launch_gather_rows(db->rows, db->ld, db->dim, db->dtype == VDB_F16, db->d_idx, n, d_out, db->wstream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * (size_t)db->dim * sizeof(float), cudaMemcpyDeviceToHost, db->wstream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(db->wstream);
    cudaFree(d_out);
    CU_TRY(e);
    return VDB_OK;
}

// ---- snapshot: header | rows | sqnorm | labels(u32) | tombstone words(u64 host mirror) ----
struct SnapHeader {
    char magic[8];
    uint32_t version, dim, ld, metric, dtype, affine;
    uint64_t count, live;
    int64_t label_base;
    uint32_t max_sqnorm_bits, pad;
};

int vdb_save(vdb_t* db, const char* path) {
    if (!db || !path) return fail(VDB_EINVAL, "null argument");
    std::lock_guard<std::mutex> wlk(db->wmu);          // no writer changes the shard while it is read out;
    std::shared_lock<std::shared_mutex> lk(db->mu);    // searches carry on
    CU_TRY(cudaSetDevice(db->device));
    CU_TRY(cudaStreamSynchronize(db->wstream));
    const size_t n = db->count.load();
    std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(VDB_EIO, std::string("cannot open ") + tmp);
    SnapHeader h{};
    memcpy(h.magic, "VDBB200\0", 8);
    h.version = 1; h.dim = db->dim; h.ld = db->ld; h.metric = db->metric; h.dtype = db->dtype; h.affine = db->affine;
    h.count = n; h.live = db->live.load(); h.label_base = db->label_base;
    cudaMemcpy(&h.max_sqnorm_bits, db->d_max_sqnorm, 4, cudaMemcpyDeviceToHost);
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / db->row_bytes());
    std::vector<uint8_t> buf(chunk_rows * db->row_bytes());
    for (size_t r = 0; ok && r < n; r += chunk_rows) {
        const size_t m = std::min(chunk_rows, n - r);
        if (cudaMemcpy(buf.data(), (uint8_t*)db->rows + r * db->row_bytes(), m * db->row_bytes(), cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
        else ok = fwrite(buf.data(), db->row_bytes(), m, f) == m;
    }
    std::vector<uint32_t> tmp32(n);
    if (ok && n) {
        ok = cudaMemcpy(tmp32.data(), db->sqnorm, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess && fwrite(tmp32.data(), 4, n, f) == n;
        ok = ok && cudaMemcpy(tmp32.data(), db->labels, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess && fwrite(tmp32.data(), 4, n, f) == n;
        const size_t w = (n + 63) / 64;
        ok = ok && fwrite(db->h_dead.data(), 8, w, f) == w;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); cudaGetLastError(); return fail(VDB_EIO, std::string("write failed: ") + path); }
    if (rename(tmp.c_str(), path) != 0) return fail(VDB_EIO, std::string("rename failed: ") + path);
    return VDB_OK;
}

int vdb_load(const char* path, size_t capacity, int device, vdb_t** out) {
    if (!path || !out) return fail(VDB_EINVAL, "null argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(VDB_EIO, std::string("cannot open ") + path);
    SnapHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "VDBB200\0", 8) != 0 || h.version != 1) {
        fclose(f);
        return fail(VDB_EIO, "not a vdb_b200 snapshot");
    }
    const size_t n = h.count;
    if (capacity < n) capacity = n;
    vdb_t* db = nullptr;
    int rc = vdb_create((int)h.dim, (int)h.metric, (int)h.dtype, capacity, device, &db);
    if (rc) { fclose(f); return rc; }
    if ((uint32_t)db->ld != h.ld) { fclose(f); vdb_destroy(db); return fail(VDB_EIO, "snapshot row stride mismatch"); }
    bool ok = true;
    const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / db->row_bytes());
    std::vector<uint8_t> buf(chunk_rows * db->row_bytes());
    for (size_t r = 0; ok && r < n; r += chunk_rows) {
        const size_t m = std::min(chunk_rows, n - r);
        ok = fread(buf.data(), db->row_bytes(), m, f) == m &&
             cudaMemcpy((uint8_t*)db->rows + r * db->row_bytes(), buf.data(), m * db->row_bytes(), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    std::vector<uint32_t> tmp32(n);
    if (ok && n) {
        ok = fread(tmp32.data(), 4, n, f) == n && cudaMemcpy(db->sqnorm, tmp32.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && fread(tmp32.data(), 4, n, f) == n && cudaMemcpy(db->labels, tmp32.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok && !h.affine) {
            db->affine = false;
            db->map.reserve(n * 2);
            for (size_t r = 0; r < n; ++r) db->map[(int64_t)tmp32[r]] = (uint32_t)r;   // later rows win
        }
        const size_t w = (n + 63) / 64;
        ok = ok && fread(db->h_dead.data(), 8, w, f) == w;
        if (ok) {
            std::vector<uint32_t> words((n + 31) / 32 + 1, 0);
            bool any = false;
            for (size_t r = 0; r < n; ++r)
                if (db->dead(r)) { words[r >> 5] |= 1u << (r & 31); any = true; }
            db->any_dead.store(any);
            ok = cudaMemcpy(db->tomb, words.data(), ((n + 31) / 32) * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        }
    }
    fclose(f);
    if (!ok) { cudaGetLastError(); vdb_destroy(db); return fail(VDB_EIO, "snapshot truncated or copy failed"); }
    cudaMemcpy(db->d_max_sqnorm, &h.max_sqnorm_bits, 4, cudaMemcpyHostToDevice);
    if (db->shadow && n) {   // derived data: rebuilt from the fp32 rows, not stored in the snapshot
        if (launch_shadow_rows((const float*)db->rows, db->ld, db->shadow, db->ld16, 0, n, db->wstream) != cudaSuccess ||
            cudaStreamSynchronize(db->wstream) != cudaSuccess) {
            cudaGetLastError();
            vdb_destroy(db);
            return fail(VDB_ECUDA, "rebuilding the fp16 shadow plane failed");
        }
    }
    db->label_base = h.label_base;
    db->live.store(h.live);
    db->count.store(n);
    *out = db;
    return VDB_OK;
}

// ---- append-only shard image -----------------------------------------------------------------------------------
// rows / norms / labels of an appended row never change, so the on-disk form of a shard can be append-only too:
//   <image_dir>/rows.bin  sqnorm.bin  labels.bin     grow by the rows added since the last save
//   <meta_path>                                       header + tombstone words of THIS moment (small; tmp + rename)
// A checkpoint costs O(new rows) instead of rewriting the shard (2 GB per million rows with vdb_save); loading reads
// the prefix [0, count) that the meta names.  Rows the image holds beyond `image_saved` (a save that crashed before
// its meta was renamed, or a restart from an older meta) are cut off before the next append.
namespace {
struct Fd {
    int fd = -1;
    ~Fd() { if (fd >= 0) ::close(fd); }
};
bool write_all(int fd, const void* p, size_t n, off_t off) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        ssize_t w = ::pwrite(fd, c, n, off);
        if (w <= 0) return false;
        c += w; n -= (size_t)w; off += w;
    }
    return true;
}
bool read_all(int fd, void* p, size_t n, off_t off) {
    char* c = static_cast<char*>(p);
    while (n) {
        ssize_t r = ::pread(fd, c, n, off);
        if (r <= 0) return false;
        c += r; n -= (size_t)r; off += r;
    }
    return true;
}
}  // namespace

int vdb_save_image(vdb_t* db, const char* image_dir, const char* meta_path) {
    if (!db || !image_dir || !meta_path) return fail(VDB_EINVAL, "null argument");
    std::lock_guard<std::mutex> wlk(db->wmu);
    std::shared_lock<std::shared_mutex> lk(db->mu);
    CU_TRY(cudaSetDevice(db->device));
    CU_TRY(cudaStreamSynchronize(db->wstream));
    const size_t n = db->count.load();
    ::mkdir(image_dir, 0755);
    const std::string dir(image_dir);
    struct Part { const char* name; const uint8_t* dev; size_t bytes_per_row; };
    const Part parts[3] = {{"rows.bin", (const uint8_t*)db->rows, db->row_bytes()},
                           {"sqnorm.bin", (const uint8_t*)db->sqnorm, sizeof(float)},
                           {"labels.bin", (const uint8_t*)db->labels, sizeof(uint32_t)}};
    size_t have = db->image_saved;
    for (const Part& pt : parts) {                 // what the files really hold (a foreign or shorter image: start over)
        struct stat sb;
        const std::string f = dir + "/" + pt.name;
        if (::stat(f.c_str(), &sb) != 0) { have = 0; break; }
        have = std::min(have, (size_t)sb.st_size / pt.bytes_per_row);
    }
    have = std::min(have, n);
    std::vector<uint8_t> buf;
    for (const Part& pt : parts) {
        Fd f;
        f.fd = ::open((dir + "/" + pt.name).c_str(), O_RDWR | O_CREAT, 0644);
        if (f.fd < 0) return fail(VDB_EIO, std::string("cannot open ") + dir + "/" + pt.name);
        if (::ftruncate(f.fd, (off_t)(have * pt.bytes_per_row)) != 0) return fail(VDB_EIO, "ftruncate failed");
        const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / pt.bytes_per_row);
        buf.resize(std::min(chunk_rows, n - have + 1) * pt.bytes_per_row);
        for (size_t r = have; r < n; r += chunk_rows) {
            const size_t m = std::min(chunk_rows, n - r);
            CU_TRY(cudaMemcpy(buf.data(), pt.dev + r * pt.bytes_per_row, m * pt.bytes_per_row, cudaMemcpyDeviceToHost));
            if (!write_all(f.fd, buf.data(), m * pt.bytes_per_row, (off_t)(r * pt.bytes_per_row)))
                return fail(VDB_EIO, std::string("write failed: ") + pt.name);
        }
        if (::fsync(f.fd) != 0) return fail(VDB_EIO, "fsync failed");
    }
    // the meta names how many rows count, with this moment's tombstones: written last, atomically
    SnapHeader h{};
    memcpy(h.magic, "VDBIMG2\0", 8);
    h.version = 2; h.dim = db->dim; h.ld = db->ld; h.metric = db->metric; h.dtype = db->dtype; h.affine = db->affine.load();
    h.count = n; h.live = db->live.load(); h.label_base = db->label_base;
    CU_TRY(cudaMemcpy(&h.max_sqnorm_bits, db->d_max_sqnorm, 4, cudaMemcpyDeviceToHost));
    const std::string tmp = std::string(meta_path) + ".tmp";
    {
        Fd f;
        f.fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (f.fd < 0) return fail(VDB_EIO, std::string("cannot open ") + tmp);
        const size_t w = (n + 63) / 64;
        if (!write_all(f.fd, &h, sizeof(h), 0) || (w && !write_all(f.fd, db->h_dead.data(), w * 8, sizeof(h))) || ::fsync(f.fd) != 0)
            return fail(VDB_EIO, std::string("write failed: ") + tmp);
    }
    if (::rename(tmp.c_str(), meta_path) != 0) return fail(VDB_EIO, std::string("rename failed: ") + meta_path);
    db->image_saved = n;
    return VDB_OK;
}

int vdb_load_image(const char* image_dir, const char* meta_path, size_t capacity, int device, vdb_t** out) {
    if (!image_dir || !meta_path || !out) return fail(VDB_EINVAL, "null argument");
    *out = nullptr;
    Fd mf;
    mf.fd = ::open(meta_path, O_RDONLY);
    if (mf.fd < 0) return fail(VDB_EIO, std::string("cannot open ") + meta_path);
    SnapHeader h{};
    if (!read_all(mf.fd, &h, sizeof(h), 0) || memcmp(h.magic, "VDBIMG2\0", 8) != 0 || h.version != 2)
        return fail(VDB_EIO, "not a vdb_b200 image meta file");
    const size_t n = h.count;
    if (capacity < n) capacity = n;
    vdb_t* db = nullptr;
    int rc = vdb_create((int)h.dim, (int)h.metric, (int)h.dtype, capacity, device, &db);
    if (rc) return rc;
    auto bail = [&](const std::string& msg) { cudaGetLastError(); vdb_destroy(db); return fail(VDB_EIO, msg); };
    if ((uint32_t)db->ld != h.ld) return bail("image row stride mismatch");
    const std::string dir(image_dir);
    struct Part { const char* name; uint8_t* dev; size_t bytes_per_row; };
    const Part parts[3] = {{"rows.bin", (uint8_t*)db->rows, db->row_bytes()},
                           {"sqnorm.bin", (uint8_t*)db->sqnorm, sizeof(float)},
                           {"labels.bin", (uint8_t*)db->labels, sizeof(uint32_t)}};
    std::vector<uint8_t> buf;
    std::vector<uint32_t> labels32;
    for (const Part& pt : parts) {
        Fd f;
        f.fd = ::open((dir + "/" + pt.name).c_str(), O_RDONLY);
        if (f.fd < 0) return bail(std::string("cannot open ") + dir + "/" + pt.name);
        const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / pt.bytes_per_row);
        buf.resize(std::min(chunk_rows, n + 1) * pt.bytes_per_row);
        for (size_t r = 0; r < n; r += chunk_rows) {
            const size_t m = std::min(chunk_rows, n - r);
            if (!read_all(f.fd, buf.data(), m * pt.bytes_per_row, (off_t)(r * pt.bytes_per_row)))
                return bail(std::string("image shorter than its meta says: ") + pt.name);
            if (cudaMemcpy(pt.dev + r * pt.bytes_per_row, buf.data(), m * pt.bytes_per_row, cudaMemcpyHostToDevice) != cudaSuccess)
                return bail("copy to the device failed");
            if (!h.affine && pt.dev == (uint8_t*)db->labels) {
                const uint32_t* l = reinterpret_cast<const uint32_t*>(buf.data());
                labels32.insert(labels32.end(), l, l + m);
            }
        }
    }
    if (!h.affine) {
        db->affine.store(false);
        db->map.reserve(n * 2);
        for (size_t r = 0; r < n; ++r) db->map[(int64_t)labels32[r]] = (uint32_t)r;   // later rows win
    }
    const size_t w = (n + 63) / 64;
    if (w && !read_all(mf.fd, db->h_dead.data(), w * 8, sizeof(h))) return bail("meta file truncated");
    {
        std::vector<uint32_t> words((n + 31) / 32 + 1, 0);
        bool any = false;
        for (size_t r = 0; r < n; ++r)
            if (db->dead(r)) { words[r >> 5] |= 1u << (r & 31); any = true; }
        db->any_dead.store(any);
        if (n && cudaMemcpy(db->tomb, words.data(), ((n + 31) / 32) * 4, cudaMemcpyHostToDevice) != cudaSuccess) return bail("copy failed");
    }
    cudaMemcpy(db->d_max_sqnorm, &h.max_sqnorm_bits, 4, cudaMemcpyHostToDevice);
    if (db->shadow && n) {   // derived data: rebuilt from the fp32 rows
        if (launch_shadow_rows((const float*)db->rows, db->ld, db->shadow, db->ld16, 0, n, db->wstream) != cudaSuccess ||
            cudaStreamSynchronize(db->wstream) != cudaSuccess)
            return bail("rebuilding the fp16 shadow plane failed");
    }
    db->label_base = h.label_base;
    db->live.store(h.live);
    db->count.store(n);
    db->image_saved = n;
    *out = db;
    return VDB_OK;
}

int vdb_merge_topk(const float* dist, const int64_t* ids, int G, size_t nq, int k_in, int k_out, float* o_dist,
                   int64_t* o_ids, int on_device, int device, void* stream) {
    if (nq == 0) return VDB_OK;
    if (!dist || !ids || !o_dist || !o_ids) return fail(VDB_EINVAL, "null buffer");
    if (G < 1 || k_in < 1 || k_out < 1 || k_out > K_MAX) return fail(VDB_EINVAL, "bad G/k");
    if ((size_t)G * (size_t)k_in > (size_t)1 << 30) return fail(VDB_EINVAL, "too many candidates");
    MergeParams mp{};
    mp.G = G; mp.k_in = k_in; mp.nq = nq; mp.n_in = G * k_in; mp.k_out = k_out;
    if (on_device) {
        mp.in_dist = dist; mp.in_ids = ids; mp.out_dist = o_dist; mp.out_ids = o_ids;
        CU_TRY(launch_merge_topk(mp, (cudaStream_t)stream));
        return VDB_OK;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(VDB_ECUDA, "no CUDA device: this library has no CPU fallback");
    }
    CU_TRY(cudaSetDevice(device));
    const size_t nin = (size_t)G * nq * k_in, nout = nq * (size_t)k_out;
    float *d_dist = nullptr, *d_od = nullptr;
    int64_t *d_ids = nullptr, *d_oi = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_dist, nin * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_ids, nin * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_od, nout * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_oi, nout * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_dist, dist, nin * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_ids, ids, nin * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        mp.in_dist = d_dist; mp.in_ids = d_ids; mp.out_dist = d_od; mp.out_ids = d_oi;
        e = launch_merge_topk(mp, nullptr);
    }
    if (e == cudaSuccess) e = cudaMemcpy(o_dist, d_od, nout * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(o_ids, d_oi, nout * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_dist); cudaFree(d_ids); cudaFree(d_od); cudaFree(d_oi);
    CU_TRY(e);
    return VDB_OK;
}


// ---- K5x: exchange + merge over NVLink peer memory -------------------------------------------
struct vdb_xchg {
    int device = 0, rank = 0, world = 1, num_sms = 148;
    size_t max_slice = 0; int max_k = 0;
    size_t stride_src = 0, stride_parity = 0, flag_bytes = 0, total_bytes = 0;
    uint8_t* base = nullptr;                       // local allocation: [flags][keys]
    uint8_t* peer_base[XCHG_MAX_WORLD] = {};       // peer mappings ([rank] == base)
    bool connected = false;
    unsigned int* done_counter = nullptr;
    // query all-gather over the copy engines (vdb_xchg_gather_queries): [2 slots][world] arrival words, then 2 query slots
    size_t q_flag_off = 0, q_buf_off = 0, q_slot_bytes = 0;
    uint32_t* h_err = nullptr;                     // pinned, mapped: [code, step, peer, -] written by the kernel
    uint32_t* d_err = nullptr;                     // device alias of h_err
    unsigned long long timeout_ns = 30ull * 1000000000ull;
    uint32_t step = 0;
    std::mutex mu;
};

static int xchg_error(vdb_xchg* x) {
    const uint32_t code = *reinterpret_cast<volatile uint32_t*>(x->h_err);
    if (!code) return VDB_OK;
    const uint32_t step = x->h_err[1], peer = x->h_err[2];
    const char* what = code == XCHG_ERR_TIMEOUT ? "timed out waiting for rank " :
                       code == XCHG_ERR_SHAPE ? "batch size / k differs from rank " : "step counter out of sync with rank ";
    return fail(VDB_ECUDA, std::string("vdb_xchg: step ") + std::to_string(step) + " " + what + std::to_string(peer) +
                               "; the exchange is unusable (every rank must call vdb_xchg_merge_dev once per step with the same nq and k)");
}

int vdb_xchg_create(int device, int rank, int world, size_t max_slice, int max_k, vdb_xchg_t** out, unsigned char* handle64) {
    return vdb_xchg_create_q(device, rank, world, max_slice, max_k, 0, out, handle64);
}

int vdb_xchg_create_q(int device, int rank, int world, size_t max_slice, int max_k, size_t query_slot_bytes, vdb_xchg_t** out,
                      unsigned char* handle64) {
    if (!out || !handle64) return fail(VDB_EINVAL, "null argument");
    *out = nullptr;
    if (world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world) return fail(VDB_EINVAL, "bad rank/world (at most 16 ranks)");
    if (max_slice < 1 || max_k < 1 || max_k > K_MAX) return fail(VDB_EINVAL, "bad max_slice/max_k");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU_TRY(cudaSetDevice(device));
    auto x = std::make_unique<vdb_xchg>();
    x->device = device; x->rank = rank; x->world = world; x->max_slice = max_slice; x->max_k = max_k;
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    x->num_sms = prop.multiProcessorCount;
    x->flag_bytes = 256;                                             // [2 parities][16 ranks] x 8 bytes (shape << 32 | step)
    static_assert(2 * XCHG_MAX_WORLD * sizeof(uint64_t) <= 256, "flag area");
    if (const char* t = getenv("VDB_XCHG_TIMEOUT_MS")) {
        const long ms = atol(t);
        if (ms > 0) x->timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    x->stride_src = max_slice * (size_t)max_k;                       // keys per source rank
    x->stride_parity = x->stride_src * world;
    x->total_bytes = x->flag_bytes + 2 * x->stride_parity * sizeof(uint64_t);
    if (query_slot_bytes) {                                          // [arrival words 256 B][slot 0][slot 1], 256-byte aligned
        x->q_flag_off = (x->total_bytes + 255) / 256 * 256;
        x->q_slot_bytes = (query_slot_bytes + 255) / 256 * 256;
        x->q_buf_off = x->q_flag_off + 256;
        x->total_bytes = x->q_buf_off + 2 * x->q_slot_bytes;
    }
    CU_TRY(cudaMalloc((void**)&x->base, x->total_bytes));
    CU_TRY(cudaMemset(x->base, 0, x->flag_bytes));
    if (query_slot_bytes) CU_TRY(cudaMemset(x->base + x->q_flag_off, 0, 256));
    CU_TRY(cudaMalloc((void**)&x->done_counter, sizeof(unsigned int)));
    CU_TRY(cudaHostAlloc((void**)&x->h_err, 4 * sizeof(uint32_t), cudaHostAllocMapped));
    memset(x->h_err, 0, 4 * sizeof(uint32_t));
    CU_TRY(cudaHostGetDevicePointer((void**)&x->d_err, x->h_err, 0));
    CU_TRY(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, x->base));
    memcpy(handle64, &h, 64);
    x->peer_base[rank] = x->base;
    *out = x.release();
    return VDB_OK;
}

int vdb_xchg_connect(vdb_xchg_t* x, const unsigned char* handles) {
    if (!x || !handles) return fail(VDB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(x->mu);
    if (x->connected) return VDB_OK;
    CU_TRY(cudaSetDevice(x->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer_base[r] = (uint8_t*)p;
    }
    x->connected = true;
    return VDB_OK;
}

int vdb_xchg_merge_dev(vdb_xchg_t* x, const float* d_dist, const int64_t* d_ids, size_t nq, int k, float* o_dist,
                       int64_t* o_ids, void* stream) {
    if (!x || !d_dist || !d_ids || !o_dist || !o_ids) return fail(VDB_EINVAL, "null argument");
    if (!x->connected && x->world > 1) return fail(VDB_EINVAL, "vdb_xchg_connect has not been called");
    if (nq == 0) return fail(VDB_EINVAL, "empty batch");
    const size_t slice = (nq + x->world - 1) / x->world;     // rank r owns queries [r*slice, min(nq, (r+1)*slice))
    if (slice > x->max_slice || k < 1 || k > x->max_k) return fail(VDB_EINVAL, "batch slice or k beyond what the exchange was created for");
    std::lock_guard<std::mutex> lk(x->mu);
    if (int rc = xchg_error(x)) return rc;          // an earlier step failed: the ranks are no longer in lockstep
    CU_TRY(cudaSetDevice(x->device));
    XchgParams xp{};
    xp.rank = x->rank; xp.world = x->world;
    xp.step = ++x->step;
    xp.shape = (uint32_t)((nq * 2654435761ull) ^ ((uint64_t)k << 20) ^ (nq >> 12));
    xp.timeout_ns = x->timeout_ns;
    xp.err = x->d_err;
    xp.parity = (int)(xp.step & 1);
    xp.nq = nq; xp.slice = slice; xp.k = k;
    { const size_t lo = std::min(nq, (size_t)x->rank * slice), hi = std::min(nq, lo + slice); xp.owned = hi - lo; }
    xp.ids = d_ids; xp.dist = d_dist;
    // strides follow this call's k and slice: every rank computes the same values from the same (nq, k)
    xp.stride_src = slice * (size_t)k;
    xp.stride_parity = x->stride_parity;
    for (int r = 0; r < x->world; ++r) {
        xp.peer_flag[r] = reinterpret_cast<uint64_t*>(x->peer_base[r]);
        xp.peer_buf[r] = reinterpret_cast<uint64_t*>(x->peer_base[r] + x->flag_bytes);
    }
    xp.local_flag = reinterpret_cast<const uint64_t*>(x->base);
    xp.done_counter = x->done_counter;
    MergeParams mp{};
    mp.in_keys = reinterpret_cast<const uint64_t*>(x->base + x->flag_bytes) + (size_t)xp.parity * x->stride_parity;
    mp.G = x->world; mp.k_in = k; mp.key_stride_g = xp.stride_src;
    mp.nq = xp.owned; mp.n_in = x->world * k; mp.k_out = k;
    mp.out_ids = o_ids; mp.out_dist = o_dist;
    CU_TRY(launch_exchange_merge(xp, mp, x->num_sms, (cudaStream_t)stream));
    return VDB_OK;
}

// ---- query all-gather over the copy engines ---------------------------------------------------------------------
void* vdb_xchg_query_slot(vdb_xchg_t* x, int slot) {
    if (!x || !x->q_slot_bytes || slot < 0 || slot > 1) return nullptr;
    return x->base + x->q_buf_off + (size_t)slot * x->q_slot_bytes;
}

int vdb_xchg_gather_queries(vdb_xchg_t* x, const void* slice, size_t slice_bytes, int slot, uint32_t batch_no, void* stream) {
    if (!x || !slice) return fail(VDB_EINVAL, "null argument");
    if (!x->q_slot_bytes) return fail(VDB_EINVAL, "the exchange was created without query slots (vdb_xchg_create_q)");
    if (!x->connected && x->world > 1) return fail(VDB_EINVAL, "vdb_xchg_connect has not been called");
    if (slot < 0 || slot > 1 || slice_bytes * (size_t)x->world > x->q_slot_bytes || batch_no == 0)
        return fail(VDB_EINVAL, "bad slot / slice size / batch number");
    if (int rc = xchg_error(x)) return rc;
    CU_TRY(cudaSetDevice(x->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t off = x->q_buf_off + (size_t)slot * x->q_slot_bytes + (size_t)x->rank * slice_bytes;
    // own slot first (host or device source), then the same bytes into every peer's slot: DMA over NVLink, no SM
    CU_TRY(cudaMemcpyAsync(x->base + off, slice, slice_bytes, cudaMemcpyDefault, st));
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank) CU_TRY(cudaMemcpyAsync(x->peer_base[r] + off, x->base + off, slice_bytes, cudaMemcpyDeviceToDevice, st));
    XchgSignal sg{};
    sg.world = x->world;
    for (int r = 0; r < x->world; ++r)
        sg.flag[r] = reinterpret_cast<uint32_t*>(x->peer_base[r] + x->q_flag_off) + slot * XCHG_MAX_WORLD + x->rank;
    sg.value = batch_no;
    CU_TRY(launch_xchg_signal(sg, st));
    return VDB_OK;
}

int vdb_xchg_wait_queries(vdb_xchg_t* x, int slot, uint32_t batch_no, void* stream) {
    if (!x) return fail(VDB_EINVAL, "null argument");
    if (!x->q_slot_bytes || slot < 0 || slot > 1) return fail(VDB_EINVAL, "bad slot, or no query slots");
    if (int rc = xchg_error(x)) return rc;
    CU_TRY(cudaSetDevice(x->device));
    XchgWait w{};
    w.world = x->world;
    w.flag = reinterpret_cast<const uint32_t*>(x->base + x->q_flag_off) + slot * XCHG_MAX_WORLD;
    w.value = batch_no;
    w.timeout_ns = x->timeout_ns;
    w.err = x->d_err;
    CU_TRY(launch_xchg_wait(w, (cudaStream_t)stream));
    return VDB_OK;
}

void vdb_xchg_destroy(vdb_xchg_t* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
    if (x->base) cudaFree(x->base);
    if (x->done_counter) cudaFree(x->done_counter);
    if (x->h_err) cudaFreeHost(x->h_err);
    delete x;
}

int vdb_xchg_status(vdb_xchg_t* x) {
    if (!x) return fail(VDB_EINVAL, "null argument");
    return xchg_error(x);
}

int vdb_set_option(vdb_t* db, const char* name, long value) {
    if (!db || !name) return fail(VDB_EINVAL, "null argument");
    if (!strcmp(name, "path")) { db->opt_path.store(value); return VDB_OK; }
    if (!strcmp(name, "scan_batch")) { db->opt_scan_batch.store(value); return VDB_OK; }
    if (!strcmp(name, "shadow")) { db->opt_shadow.store(value); return VDB_OK; }
    if (!strcmp(name, "small_batch_tensor_rows")) { db->opt_small_batch_tensor_rows.store(value); return VDB_OK; }
    if (!strcmp(name, "shadow_scan_rows")) { db->opt_shadow_scan_rows.store(value); return VDB_OK; }
    if (!strcmp(name, "shadow_scan_nq")) { db->opt_shadow_scan_nq.store(value < 0 ? 0 : value > 2 ? 2 : value); return VDB_OK; }
    if (!strcmp(name, "profile")) { db->opt_profile.store(value); return VDB_OK; }
    return fail(VDB_EINVAL, std::string("unknown option ") + name);
}
long vdb_get_stat(vdb_t* db, const char* name) {
    if (!db || !name) return -1;
    if (!strcmp(name, "fallback_queries")) {      // counted on the device, per workspace (synchronises)
        long total = db->stat_fallback.load();
        cudaSetDevice(db->device);
        std::lock_guard<std::mutex> lk(db->ws_mu);
        for (auto& w : db->ws_all) total += gemm_workspace_fallbacks(w->gemm);
        return total;
    }
    if (!strcmp(name, "tensor_batches")) return db->stat_tensor_batches.load();
    if (!strcmp(name, "scan_passes")) return db->stat_scan_passes.load();
    if (!strcmp(name, "shadow_scans")) return db->stat_shadow_scans.load();
    if (!strcmp(name, "num_sms")) return db->num_sms;
    if (!strcmp(name, "profile_ns") || !strcmp(name, "profile_count")) {
        // sum of (stop - start) over the dominant-kernel launches recorded since the last read
        std::lock_guard<std::mutex> lk(db->prof_mu);
        if (!strcmp(name, "profile_count")) return (long)db->prof_events.size();
        double total_ms = 0.0;
        for (auto& ev : db->prof_events) {
            cudaEventSynchronize(ev.second);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) total_ms += ms;
            db->prof_pool.push_back(ev);
        }
        db->prof_events.clear();
        return (long)(total_ms * 1e6);
    }
    if (!strcmp(name, "ld")) return db->ld;
    if (!strcmp(name, "ld16")) return db->shadow ? db->ld16 : 0;
    if (!strcmp(name, "shadow")) return db->shadow ? 1 : 0;
    return -1;
}

}  // extern "C"
