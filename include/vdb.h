/* vdb.h -- C ABI of the B200 exact-kNN shard ("libvdb_b200.so").
 *
 * This is the drop-in boundary for the datanode search hot path of
 * f1ybaozii/Distributed-Vector-Database.  The reference has no FFI of its own: the seam is
 * the set of hnswlib.Index method calls made by VectorNodeHandler plus the coordinator's
 * merge.  Each entry point below names the reference call it replaces
 * (paths relative to the reference checkout).  Python binds these with ctypes
 * (distributed-vector-database_b200/_ffi.py); INTEGRATION.md shows the stub a maintainer
 * of the reference would add.
 *
 * Conventions: every function returning int gives 0 on success, a negative VDB_E* code on
 * failure and leaves a message for vdb_last_error() (thread-local).  Unless a name ends in
 * _dev, pointers are HOST pointers owned by the caller; the library owns all device memory
 * behind the opaque handle.  Calls on one handle are thread-safe: searches run concurrently
 * (one internal stream + workspace per in-flight call), writers are exclusive.
 *
 * Labels (hnswlib "ids") must lie in [0, 2^32-2]; results order by (distance, label).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef VDB_B200_H
#define VDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vdb vdb_t;

enum vdb_metric { VDB_L2 = 0, VDB_IP = 1, VDB_COSINE = 2 };  /* hnswlib space 'l2' | 'ip' | 'cosine' */
enum vdb_dtype { VDB_F32 = 0, VDB_F16 = 1 };                 /* storage type of the shard rows      */
enum vdb_error {
    VDB_OK = 0,
    VDB_EINVAL = -1,   /* bad argument                                  */
    VDB_ECUDA = -2,    /* CUDA runtime/driver error (message has it)   */
    VDB_EFULL = -3,    /* add beyond capacity (hnswlib: RuntimeError)  */
    VDB_ENOMEM = -4,
    VDB_EIO = -5,
    VDB_ENOTFOUND = -6 /* label not present                            */
};

/* hnswlib.Index(space, dim) + init_index(max_elements, ...)   src/datanode/handler.py:46,86
 * ef_construction / M have no meaning for an exact index and are not taken.
 * dim <= 1408 for fp32 rows, <= 2816 for fp16 rows (16 rows of a shard must fit one stage of the scan kernel's
 * shared-memory ring; CLIP-style embeddings are 512 - 1024 wide); larger dims fail with VDB_EINVAL. */
int vdb_create(int dim, int metric, int store_dtype, size_t capacity, int device, vdb_t **out);
void vdb_destroy(vdb_t *db);

/* hnswlib.Index.add_items(float32[n,dim], int64[n])           handler.py:112,268-271,279-282
 * cosine rows are L2-normalised on insert as hnswlib does; ||d||^2 is stored beside each row.
 * A label that is already live is tombstoned and re-appended (same observable result as
 * hnswlib's in-place update).  VDB_EFULL when count + n > capacity. */
int vdb_add(vdb_t *db, const float *rows, const int64_t *labels, size_t n);
int vdb_add_dev(vdb_t *db, const float *d_rows, const int64_t *h_labels, size_t n, void *stream);
/* Bench/test utility: append rows [row_start, row_start+n) of the synthetic unit-norm set
 * `seed` (definition: oracle/cpu_ref.py synth_rows), generated on the device, labels
 * label_start + i. */
int vdb_add_synthetic(vdb_t *db, uint64_t seed, uint64_t row_start, size_t n, int64_t label_start);
/* Same generator into a caller-owned device buffer [n, dim] fp32 (queries for benches). */
int vdb_synth_dev(uint64_t seed, uint64_t row_start, size_t n, int dim, float *d_out, void *stream);

/* Tombstones: `if hnsw_id in self.deleted_ids: continue`      handler.py:256,332,378
 * (hnswlib.Index.mark_deleted / unmark_deleted).  Masked inside the scan. */
int vdb_mark_deleted(vdb_t *db, const int64_t *labels, size_t n);
int vdb_unmark_deleted(vdb_t *db, const int64_t *labels, size_t n);

/* hnswlib.Index.knn_query(float32[nq,dim], k) -> (labels[nq,k], distances[nq,k]) ascending
 *                                                              handler.py:364
 * Exact over all live rows.  Rows with fewer than k live neighbours are padded with label -1
 * and distance +inf; out_counts[q] (optional) is the number of real results.
 * nq <= 4 (2 for fp16 rows) takes the HBM-streaming scan kernel, larger nq the tcgen05 kernel + exact
 * fp32 re-rank.  On an fp32 shard with an fp16 shadow plane, a single query over >= 360k rows scans the shadow
 * (half the bytes) and the candidates that can still matter are recomputed from the fp32 rows, and 2..4 queries
 * over >= 500k rows take the tcgen05 path: the results are bit-identical to the fp32 scan's either way. */
int vdb_search(vdb_t *db, const float *queries, size_t nq, int k, int64_t *out_labels,
               float *out_dist, int *out_counts);
/* The same call in two halves, for a server that keeps more than one batch in flight (the reference's Thrift
 * server runs one handler call per client thread: datanode/server.py; here one thread can overlap them):
 * submit enqueues upload + search + download on a stream of its own and returns a ticket; collect waits until the
 * results are in the caller's buffers and frees the ticket (always call it once per successful submit).  The
 * buffers must stay valid and untouched in between.  vdb_search == submit + collect. */
typedef struct vdb_ticket vdb_ticket_t;
int vdb_search_submit(vdb_t *db, const float *queries, size_t nq, int k, int64_t *out_labels, float *out_dist,
                      int *out_counts, vdb_ticket_t **ticket);
int vdb_search_collect(vdb_ticket_t *ticket);
/* Page-locked host buffers for queries / results: vdb_search moves them by DMA without the staging copy it
 * needs for pageable memory (any page-locked buffer is recognised, these two are a convenience). */
void *vdb_host_alloc(size_t bytes);
void vdb_host_free(void *p);
/* Same with DEVICE pointers, enqueued on `stream` (a cudaStream_t; NULL = default stream),
 * no host synchronisation. */
int vdb_search_dev(vdb_t *db, const float *d_queries, size_t nq, int k, int64_t *d_labels,
                   float *d_dist, int *d_counts, void *stream);

/* get_current_count / get_max_elements                         handler.py:82,196,237-238,350 */
size_t vdb_count(const vdb_t *db);       /* rows appended, tombstoned ones included (hnswlib semantics) */
size_t vdb_live_count(const vdb_t *db);
size_t vdb_capacity(const vdb_t *db);
int vdb_dim(const vdb_t *db);
/* hnswlib.Index.resize_index(new_max) -- the reference rebuilds instead (handler.py:91-120) */
int vdb_resize(vdb_t *db, size_t new_capacity);
/* hnswlib.Index.get_items(labels): rows AS STORED (normalised / fp16-rounded), fp32 out */
int vdb_get_rows(vdb_t *db, const int64_t *labels, size_t n, float *out);

/* save_index(path) / load_index(path, max_elements)            handler.py:65,80,115,164,195,302
 * A flat shard snapshot (header, rows, norms, labels, tombstones) replaces hnswlib's
 * private index.bin. */
int vdb_save(vdb_t *db, const char *path);
int vdb_load(const char *path, size_t capacity, int device, vdb_t **out);

/* The same for checkpoints that come often (every 2000 puts in the reference, handler.py:316-317): rows, norms and
 * labels of an appended row never change, so the shard keeps an APPEND-ONLY image on disk -- <image_dir>/rows.bin,
 * sqnorm.bin, labels.bin grow by the rows added since the last save -- and a checkpoint is a small meta file (header
 * + the tombstone bitmap of that moment, written to meta_path by tmp + rename).  O(new rows) instead of rewriting the
 * shard.  vdb_load_image reads the prefix the meta names; older metas stay loadable (their prefix never changes). */
int vdb_save_image(vdb_t *db, const char *image_dir, const char *meta_path);
int vdb_load_image(const char *image_dir, const char *meta_path, size_t capacity, int device, vdb_t **out);

/* CoordinatorHandler.search merge: sorted(range(n), key=score)[:top_k]
 *                                                              src/coordinator/handler.py:212-216
 * dist/ids [G, nq, k_in] (id < 0 = padding) -> ascending (distance, id) top k_out per query.
 * on_device != 0: all four pointers are device pointers and the work is enqueued on
 * `stream`; otherwise host pointers, synchronous. */
int vdb_merge_topk(const float *dist, const int64_t *ids, int G, size_t nq, int k_in, int k_out,
                   float *o_dist, int64_t *o_ids, int on_device, int device, void *stream);

/* The same gather + merge (coordinator/handler.py:191-216) for G ranks of ONE box, fused into one kernel over
 * NVLink peer memory instead of a collective library call: every rank has searched its shard for the whole
 * batch; rank r owns the answers of queries [r*nq/G, (r+1)*nq/G).  The kernel stores each query's list into
 * the owner's receive buffer (CUDA-IPC peer pointers), signals, waits for the G-1 peers and merges its slice.
 *   create : allocates this rank's receive buffer, returns its 64-byte IPC handle
 *   connect: takes all G handles (rank order; exchange them with any out-of-band channel, e.g. an all-gather)
 *   merge  : d_dist/d_ids [nq,k] = this rank's lists (id < 0 = padding) -> o_dist/o_ids [slice,k] with
 *            slice = ceil(nq/G) (rank r owns queries [r*slice, min(nq,(r+1)*slice)); rows beyond what it owns
 *            are left untouched), enqueued on `stream`.  Collective: every rank must call it once per step with the same nq and k; one step in
 *            flight per rank.  The kernel is launched cooperatively (its whole grid is placed at once), so other work
 *            on the same GPU -- searches on other streams -- can delay but not deadlock it.  The one thing that may
 *            NOT run beside it is another kernel that waits for the same peers (a collective of a communication
 *            library on another stream): two ranks can then each hold the resources the other's waiting kernel needs.
 *            Enqueue collectives on the stream that carries this call. */
typedef struct vdb_xchg vdb_xchg_t;
int vdb_xchg_create(int device, int rank, int world, size_t max_slice, int max_k, vdb_xchg_t **out,
                    unsigned char *handle64);
int vdb_xchg_connect(vdb_xchg_t *x, const unsigned char *handles /* world * 64 bytes */);
/* The broadcast of the queries (coordinator/handler.py:186-197: the same SearchRequest to every node) for ranks that
 * each hold a slice of the batch, over the COPY ENGINES: create_q reserves two query slots of query_slot_bytes in the
 * exported buffer; gather_queries copies this rank's slice (host or device pointer) into its place of slot `slot`
 * here and -- by DMA over NVLink -- on every peer, then publishes batch_no (1, 2, 3, ...; slot = batch_no & 1 by
 * convention) in every rank's arrival array, all enqueued on `stream` (a copy stream); wait_queries enqueues, on the
 * stream that will search, a one-warp kernel that waits until every rank's slice of batch_no has landed in the local
 * slot (bounded like the merge; errors through vdb_xchg_status).  No collective-library kernel is involved, so it
 * may overlap the previous batch's search and exchange.  The caller must not gather batch i into a slot before its
 * own merge of batch i-2 has completed (then every peer has finished searching that slot). */
int vdb_xchg_create_q(int device, int rank, int world, size_t max_slice, int max_k, size_t query_slot_bytes,
                      vdb_xchg_t **out, unsigned char *handle64);
void *vdb_xchg_query_slot(vdb_xchg_t *x, int slot);   /* device pointer of the local slot: [world][slice] rows */
int vdb_xchg_gather_queries(vdb_xchg_t *x, const void *slice, size_t slice_bytes, int slot, uint32_t batch_no, void *stream);
int vdb_xchg_wait_queries(vdb_xchg_t *x, int slot, uint32_t batch_no, void *stream);
int vdb_xchg_merge_dev(vdb_xchg_t *x, const float *d_dist, const int64_t *d_ids, size_t nq, int k,
                       float *o_dist, int64_t *o_ids, void *stream);
/* 0, or VDB_ECUDA once a step has failed: a peer did not arrive within the timeout (30 s; VDB_XCHG_TIMEOUT_MS),
 * or arrived with another (nq, k) / step.  The failing step's outputs are padding (-1 / +inf); the exchange stays
 * failed (vdb_xchg_merge_dev returns the same error).  Read it after synchronising the stream of the step. */
int vdb_xchg_status(vdb_xchg_t *x);
void vdb_xchg_destroy(vdb_xchg_t *x);

/* Introspection for bench.py / tests */
uint64_t vdb_launch_count(void);              /* kernels this library has launched so far   */
/* options: "path" (0 auto, 1 scan, 2 tensor), "scan_batch" (largest batch the scan takes in auto mode), "shadow"
 * (0: tensor path on the stored rows), "shadow_scan_nq" (0..2: batches up to this size may scan the shadow plane;
 * default 1), "shadow_scan_rows" / "small_batch_tensor_rows" (shard sizes from which those routes apply), "profile" */
int vdb_set_option(vdb_t *db, const char *name, long value);
long vdb_get_stat(vdb_t *db, const char *name); /* "fallback_queries", "tensor_batches", ... */
const char *vdb_last_error(void);
const char *vdb_version(void);
/* The launch sequence the batched tensor-core search would use for nq queries over n_rows rows at top-k (pure
 * host arithmetic, no device needed).  out[0..9] = k', tight rank, buffer keys per query, growth, 256-query blocks,
 * 256-row tiles, positions (tiles rounded up to a power of two), probe positions, probe threshold rank, levels L;
 * then L triples (first position, end position, threshold rank published after the level; 0 = last level).
 * Returns the number of ints the plan needs (written up to out_len), or a negative error code. */
int vdb_debug_level_plan(size_t nq, size_t n_rows, int k, int *out, int out_len);

#ifdef __cplusplus
}
#endif
#endif /* VDB_B200_H */
