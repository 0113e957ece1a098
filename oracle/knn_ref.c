/* CPU oracle, C restatement -- TEST INFRASTRUCTURE ONLY (see oracle/cpu_ref.py header).
 *
 * PARITY UNPINNED: the reference has no tests/golden vectors for this arithmetic and its
 * native dependency (hnswlib, unpinned, not vendored) cannot be installed here.  This file
 * restates the published hnswlib v0.8.0 algorithm for the three spaces the path uses and is
 * itself checked bit-for-bit against oracle/cpu_ref.py (numpy) in tests/test_oracle.py.
 *
 *   distance kernels   hnswlib space_l2.h L2SqrSIMD16ExtAVX / space_ip.h
 *                      InnerProductSIMD16ExtAVX (8 fp32 lanes, mul then add, lanes reduced
 *                      left to right; residual elements summed sequentially and added last)
 *   normalisation      hnswlib python_bindings normalize_vector (sequential fp32 sum)
 *   top-k              hnswlib bruteforce.h searchKnn: every live row, result ordered by the
 *                      std::pair (distance, label) => ties to the smaller label
 *   call sites         reference src/datanode/handler.py:46 (space), :268 (add), :364 (query)
 *
 * Build: gcc -O3 -mavx2 -ffp-contract=off -fopenmp -shared -fPIC  (see oracle/Makefile).
 * -ffp-contract=off keeps "mul then add" as two roundings, as the intrinsics are written.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORACLE_L2 = 0, ORACLE_IP = 1 };

static inline float l2_lanes8(const float *a, const float *b, int d) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int d16 = d & ~15;
    for (int j = 0; j < d16; j += 8)
        for (int l = 0; l < 8; ++l) {
            float t = a[j + l] - b[j + l];
            s[l] = s[l] + t * t;
        }
    float r = 0.0f;
    if (d16) {
        r = s[0];
        for (int l = 1; l < 8; ++l) r = r + s[l];
    }
    if (d16 < d) {
        float tail = 0.0f;
        for (int j = d16; j < d; ++j) {
            float t = a[j] - b[j];
            tail = tail + t * t;
        }
        r = r + tail;
    }
    return r;
}

static inline float ip_lanes8(const float *a, const float *b, int d) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int d16 = d & ~15;
    for (int j = 0; j < d16; j += 8)
        for (int l = 0; l < 8; ++l) s[l] = s[l] + a[j + l] * b[j + l];
    float r = 0.0f;
    if (d16) {
        r = s[0];
        for (int l = 1; l < 8; ++l) r = r + s[l];
    }
    if (d16 < d) {
        float tail = 0.0f;
        for (int j = d16; j < d; ++j) tail = tail + a[j] * b[j];
        r = r + tail;
    }
    return r;
}

static inline float dist_one(const float *q, const float *row, int d, int metric) {
    return metric == ORACLE_L2 ? l2_lanes8(q, row, d) : 1.0f - ip_lanes8(q, row, d);
}

void oracle_distances(const float *q, const float *rows, size_t n, int dim, size_t ld, int metric,
                      float *out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) out[i] = dist_one(q, rows + (size_t)i * ld, dim, metric);
}

void oracle_normalize(const float *x, size_t n, int dim, float *out) {
    for (size_t i = 0; i < n; ++i) {
        const float *r = x + i * (size_t)dim;
        float norm = 0.0f;
        for (int j = 0; j < dim; ++j) norm = norm + r[j] * r[j];
        norm = 1.0f / (sqrtf(norm) + 1e-30f);
        for (int j = 0; j < dim; ++j) out[i * (size_t)dim + j] = r[j] * norm;
    }
}

/* ---- synthetic rows: same definition as cpu_ref.synth_rows and csrc/synth.cuh ---- */
static inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_synth_rows(uint64_t seed, uint64_t row_start, size_t n, int dim, float *out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        uint64_t base = splitmix64(seed * 0x9E3779B97F4A7C15ull + (row_start + (uint64_t)i));
        int64_t ss = 0;
        float *o = out + (size_t)i * dim;
        for (int c = 0; c < dim; ++c) {
            uint64_t h = splitmix64(base + (uint64_t)c);
            int64_t v = (int64_t)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 131070;
            ss += v * v;
            o[c] = (float)v; /* exact: |v| < 2^18 */
        }
        double inv = sqrt((double)ss);
        for (int c = 0; c < dim; ++c) o[c] = (float)((double)o[c] / inv);
    }
}

/* ---- exact top-k: max-heap of (dist,label) pairs, lexicographic ---- */
typedef struct { float d; int64_t l; } pair_t;

static inline int pair_less(pair_t a, pair_t b) { return a.d < b.d || (a.d == b.d && a.l < b.l); }

static void heap_sift_down(pair_t *h, int n, int i) {
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && pair_less(h[c], h[c + 1])) c++;
        if (!pair_less(h[i], h[c])) break;
        pair_t t = h[i]; h[i] = h[c]; h[c] = t;
        i = c;
    }
}
static void heap_sift_up(pair_t *h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!pair_less(h[p], h[i])) break;
        pair_t t = h[i]; h[i] = h[p]; h[p] = t;
        i = p;
    }
}
static inline void heap_offer(pair_t *h, int *n, int k, pair_t x) {
    if (*n < k) { h[*n] = x; heap_sift_up(h, (*n)++); }
    else if (pair_less(x, h[0])) { h[0] = x; heap_sift_down(h, k, 0); }
}
static int pair_cmp(const void *a, const void *b) {
    pair_t x = *(const pair_t *)a, y = *(const pair_t *)b;
    return pair_less(x, y) ? -1 : (pair_less(y, x) ? 1 : 0);
}

/* queries [nq, dim] (already normalised by the caller for cosine); rows [n, ld] as stored;
 * dead[n] optional (non-zero = tombstoned).  Output padded with label -1 / +inf. */
int oracle_knn(const float *queries, size_t nq, const float *rows, const int64_t *labels, size_t n,
               int dim, size_t ld, int metric, const uint8_t *dead, int k, int64_t *out_labels,
               float *out_dist, int *out_counts, int nthreads) {
    if (k <= 0) return -1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    /* one heap per (thread, query); threads split the rows so a single query also scales */
    pair_t *heaps = (pair_t *)malloc(sizeof(pair_t) * (size_t)nthreads * nq * (size_t)k);
    int *hn = (int *)calloc((size_t)nthreads * nq, sizeof(int));
    if (!heaps || !hn) { free(heaps); free(hn); return -2; }
    const size_t RB = 256; /* row block kept in cache while all queries of a group visit it */
    size_t nblocks = (n + RB - 1) / RB;
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        pair_t *myh = heaps + (size_t)t * nq * (size_t)k;
        int *myn = hn + (size_t)t * nq;
#pragma omp for schedule(static)
        for (long long b = 0; b < (long long)nblocks; ++b) {
            size_t r0 = (size_t)b * RB, r1 = r0 + RB < n ? r0 + RB : n;
            for (size_t qi = 0; qi < nq; ++qi) {
                const float *q = queries + qi * (size_t)dim;
                pair_t *h = myh + qi * (size_t)k;
                for (size_t r = r0; r < r1; ++r) {
                    if (dead && dead[r]) continue;
                    pair_t x;
                    x.d = dist_one(q, rows + r * ld, dim, metric);
                    x.l = labels ? labels[r] : (int64_t)r;
                    heap_offer(h, &myn[qi], k, x);
                }
            }
        }
    }
    pair_t *tmp = (pair_t *)malloc(sizeof(pair_t) * (size_t)nthreads * (size_t)k);
    for (size_t qi = 0; qi < nq; ++qi) {
        int m = 0;
        for (int t = 0; t < nthreads; ++t) {
            int c = hn[(size_t)t * nq + qi];
            memcpy(tmp + m, heaps + ((size_t)t * nq + qi) * (size_t)k, sizeof(pair_t) * (size_t)c);
            m += c;
        }
        qsort(tmp, (size_t)m, sizeof(pair_t), pair_cmp);
        int c = m < k ? m : k;
        for (int i = 0; i < k; ++i) {
            out_labels[qi * (size_t)k + i] = i < c ? tmp[i].l : -1;
            out_dist[qi * (size_t)k + i] = i < c ? tmp[i].d : INFINITY;
        }
        if (out_counts) out_counts[qi] = c;
    }
    free(tmp); free(heaps); free(hn);
    return 0;
}

/* OpenMP thread count for every function of this file (torchrun exports OMP_NUM_THREADS=1 to its workers:
 * a CPU baseline timed there must ask for the host's cores explicitly). */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* In place: fp32 -> fp16 (round to nearest even) -> fp32, the value an fp16-stored shard holds
 * (same as numpy's astype(float16).astype(float32)). */
void oracle_round_f16(float *x, size_t n) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) x[i] = (float)(_Float16)x[i];
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
