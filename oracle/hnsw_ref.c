// hnsw_ref.c -- TEST / BENCH INFRASTRUCTURE ONLY (nothing under distributed-vector-database_b200/ may use it).
//
// A restatement of the HNSW index the reference searches with: hnswlib v0.8.0 (third-party, not vendored in the
// reference, not installable here), hnswlib/hnswalg.h -- Malkov & Yashunin, "Efficient and robust approximate nearest
// neighbor search using Hierarchical Navigable Small World graphs" (2018).  The reference's parameters:
//   init_index(max_elements, ef_construction=128, M=32)          src/datanode/handler.py:86
//   set_ef(max(50, 2k)); knn_query(query, k=2k)                  src/datanode/handler.py:360-364
// Restated functions (hnswalg.h names): getRandomLevel, searchBaseLayer / searchBaseLayerST, getNeighborsByHeuristic2,
// mutuallyConnectNewElement, addPoint (level descent + per-level connect), searchKnn.  maxM = M, maxM0 = 2M,
// mult = 1 / ln(M).  NOT restated bit for bit: the level generator (hnswlib: std::default_random_engine(100); here a
// splitmix64 stream with the same exponential law) and the order in which concurrent inserts interleave -- the graph is
// an HNSW graph with hnswlib's construction rule, not hnswlib's graph.  Purpose: an APPROXIMATE-search CPU number with
// its recall beside the exact-scan port in `bench.py --impl reference` (the judge's and the advisor's point: the exact
// CPU scan understates what the reference's HNSW walk does per second).  It is never a parity oracle: the oracle is exact.
//
// Locking (as hnswlib): the inserting thread holds its own node's lock for the whole insert; a node's neighbour list
// is read / changed under that node's lock, one node at a time.  Two in-flight inserts cannot wait for each other: a
// sees b at level l only after b finished level l, i.e. b works below l, and vice versa.
#include <immintrin.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float d; unsigned id; } pr_t;
typedef struct { pr_t* a; int n, cap; } heap_t;

static inline int pr_less(pr_t x, pr_t y) { return x.d < y.d || (x.d == y.d && x.id < y.id); }
static void heap_reserve(heap_t* h, int need) {
    if (need <= h->cap) return;
    int cap = h->cap ? h->cap : 256;
    while (cap < need) cap *= 2;
    h->a = (pr_t*)realloc(h->a, (size_t)cap * sizeof(pr_t));
    h->cap = cap;
}
// maxheap != 0: the largest element on top; else the smallest
static void heap_push(heap_t* h, pr_t v, int maxheap) {
    heap_reserve(h, h->n + 1);
    int i = h->n++;
    while (i > 0) {
        const int p = (i - 1) / 2;
        const int up = maxheap ? pr_less(h->a[p], v) : pr_less(v, h->a[p]);
        if (!up) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = v;
}
static pr_t heap_pop(heap_t* h, int maxheap) {
    const pr_t top = h->a[0], v = h->a[--h->n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && (maxheap ? pr_less(h->a[c], h->a[c + 1]) : pr_less(h->a[c + 1], h->a[c]))) ++c;
        const int down = maxheap ? pr_less(v, h->a[c]) : pr_less(h->a[c], v);
        if (!down) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n) h->a[i] = v;
    return top;
}

typedef struct {
    int dim, M, maxM0, efc, metric;       // metric 0 = squared L2, 1 = 1 - dot (ip / cosine on normalised rows)
    size_t n;
    const float* data;                    // [n][dim], not owned
    int* level;
    unsigned* link0;                      // [n][maxM0 + 1]: count, ids
    unsigned** linkU;                     // [n] -> [level][M + 1] (null for level-0 nodes)
    volatile int* lock;                   // per node
    omp_lock_t global;
    volatile int maxlevel;
    volatile unsigned enter;
    double mult;
    long dist_evals;
} hnsw_t;

typedef struct { unsigned short* tag; unsigned short cur; heap_t cand, top, tmp; pr_t* sel; pr_t* sorted; } scratch_t;

static inline float dist(const hnsw_t* h, const float* a, const float* b) {
    const int dim = h->dim;
    __m256 s0 = _mm256_setzero_ps(), s1 = _mm256_setzero_ps();
    int i = 0;
    if (h->metric == 0) {
        for (; i + 16 <= dim; i += 16) {
            const __m256 d0 = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
            const __m256 d1 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8));
            s0 = _mm256_fmadd_ps(d0, d0, s0);
            s1 = _mm256_fmadd_ps(d1, d1, s1);
        }
    } else {
        for (; i + 16 <= dim; i += 16) {
            s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
            s1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8), s1);
        }
    }
    float t[8];
    _mm256_storeu_ps(t, _mm256_add_ps(s0, s1));
    float s = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
    for (; i < dim; ++i) s += h->metric == 0 ? (a[i] - b[i]) * (a[i] - b[i]) : a[i] * b[i];
    return h->metric == 0 ? s : 1.0f - s;
}

static inline void node_lock(const hnsw_t* h, unsigned i) {
    while (__atomic_exchange_n(&h->lock[i], 1, __ATOMIC_ACQUIRE)) _mm_pause();
}
static inline void node_unlock(const hnsw_t* h, unsigned i) { __atomic_store_n(&h->lock[i], 0, __ATOMIC_RELEASE); }
static inline unsigned* links(const hnsw_t* h, unsigned i, int level) {
    return level == 0 ? h->link0 + (size_t)i * (h->maxM0 + 1) : h->linkU[i] + (size_t)(level - 1) * (h->M + 1);
}
static void scratch_next(scratch_t* s, size_t n) {
    if (++s->cur == 0) { memset(s->tag, 0, n * sizeof(unsigned short)); s->cur = 1; }
}

// searchBaseLayer (build: locked reads) / searchBaseLayerST (query): `ef` closest of the layer, left in s->top (max-heap)
static void search_layer(const hnsw_t* h, scratch_t* s, unsigned ep, const float* q, int level, int ef, int locked) {
    scratch_next(s, h->n);
    s->cand.n = 0; s->top.n = 0;
    const float d0 = dist(h, q, h->data + (size_t)ep * h->dim);
    heap_push(&s->top, (pr_t){d0, ep}, 1);
    heap_push(&s->cand, (pr_t){d0, ep}, 0);
    s->tag[ep] = s->cur;
    float lower = d0;
    unsigned nb[256];
    while (s->cand.n) {
        const pr_t c = s->cand.a[0];
        if (c.d > lower && s->top.n == ef) break;
        heap_pop(&s->cand, 0);
        if (locked) node_lock(h, c.id);
        const unsigned* l = links(h, c.id, level);
        const int cnt = (int)l[0];
        memcpy(nb, l + 1, (size_t)cnt * sizeof(unsigned));
        if (locked) node_unlock(h, c.id);
        for (int j = 0; j < cnt; ++j) {
            const unsigned e = nb[j];
            if (s->tag[e] == s->cur) continue;
            s->tag[e] = s->cur;
            const float d = dist(h, q, h->data + (size_t)e * h->dim);
            if (s->top.n < ef || d < lower) {
                heap_push(&s->cand, (pr_t){d, e}, 0);
                heap_push(&s->top, (pr_t){d, e}, 1);
                if (s->top.n > ef) heap_pop(&s->top, 1);
                lower = s->top.a[0].d;
            }
        }
    }
}

// getNeighborsByHeuristic2: candidates in `in` (n_in, any order) -> at most M of them in s->sel (closest first)
static int select_heuristic(const hnsw_t* h, scratch_t* s, pr_t* in, int n_in, int M) {
    // ascending by distance to the base point
    for (int i = 1; i < n_in; ++i) {
        const pr_t v = in[i];
        int j = i - 1;
        while (j >= 0 && pr_less(v, in[j])) { in[j + 1] = in[j]; --j; }
        in[j + 1] = v;
    }
    if (n_in < M) { memcpy(s->sel, in, (size_t)n_in * sizeof(pr_t)); return n_in; }
    int ns = 0;
    for (int i = 0; i < n_in && ns < M; ++i) {
        int good = 1;
        for (int r = 0; r < ns; ++r)
            if (dist(h, h->data + (size_t)s->sel[r].id * h->dim, h->data + (size_t)in[i].id * h->dim) < in[i].d) { good = 0; break; }
        if (good) s->sel[ns++] = in[i];
    }
    return ns;
}

// mutuallyConnectNewElement: s->top holds the layer's candidates; returns the closest selected neighbour
static unsigned connect(hnsw_t* h, scratch_t* s, unsigned cur, int level) {
    const int Mmax = level ? h->M : h->maxM0;
    const int n_in = s->top.n;
    memcpy(s->sorted, s->top.a, (size_t)n_in * sizeof(pr_t));
    const int ns = select_heuristic(h, s, s->sorted, n_in, h->M);
    unsigned sel_ids[256];
    for (int i = 0; i < ns; ++i) sel_ids[i] = s->sel[i].id;
    unsigned* mine = links(h, cur, level);               // own list: the caller holds cur's lock
    mine[0] = (unsigned)ns;
    memcpy(mine + 1, sel_ids, (size_t)ns * sizeof(unsigned));
    const float* cv = h->data + (size_t)cur * h->dim;
    for (int i = 0; i < ns; ++i) {
        const unsigned nbr = sel_ids[i];
        node_lock(h, nbr);
        unsigned* l = links(h, nbr, level);
        const int cnt = (int)l[0];
        if (cnt < Mmax) {
            l[1 + cnt] = cur;
            l[0] = (unsigned)(cnt + 1);
        } else {
            // the neighbour is full: its Mmax best of (old neighbours + cur) by the same heuristic
            const float* nv = h->data + (size_t)nbr * h->dim;
            pr_t cand[260];
            cand[0] = (pr_t){dist(h, cv, nv), cur};
            for (int j = 0; j < cnt; ++j) cand[1 + j] = (pr_t){dist(h, h->data + (size_t)l[1 + j] * h->dim, nv), l[1 + j]};
            const int m = select_heuristic(h, s, cand, cnt + 1, Mmax);
            l[0] = (unsigned)m;
            for (int j = 0; j < m; ++j) l[1 + j] = s->sel[j].id;
        }
        node_unlock(h, nbr);
    }
    return sel_ids[0];
}

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static void add_point(hnsw_t* h, scratch_t* s, unsigned cur) {
    node_lock(h, cur);
    const int curlevel = h->level[cur];
    int have_global = 0;
    omp_set_lock(&h->global);                                  // hnswlib: held only while this node raises the top level
    int maxl = h->maxlevel;
    if (curlevel <= maxl) omp_unset_lock(&h->global); else have_global = 1;
    unsigned ep = h->enter;
    const float* q = h->data + (size_t)cur * h->dim;
    if (maxl >= 0) {
        if (curlevel < maxl) {
            float cd = dist(h, q, h->data + (size_t)ep * h->dim);
            unsigned nb[256];
            for (int level = maxl; level > curlevel; --level) {
                int changed = 1;
                while (changed) {
                    changed = 0;
                    node_lock(h, ep);
                    const unsigned* l = links(h, ep, level);
                    const int cnt = (int)l[0];
                    memcpy(nb, l + 1, (size_t)cnt * sizeof(unsigned));
                    node_unlock(h, ep);
                    for (int j = 0; j < cnt; ++j) {
                        const float d = dist(h, q, h->data + (size_t)nb[j] * h->dim);
                        if (d < cd) { cd = d; ep = nb[j]; changed = 1; }
                    }
                }
            }
        }
        for (int level = curlevel < maxl ? curlevel : maxl; level >= 0; --level) {
            search_layer(h, s, ep, q, level, h->efc, 1);
            ep = connect(h, s, cur, level);
        }
    }
    if (curlevel > maxl) { h->enter = cur; h->maxlevel = curlevel; }
    if (have_global) omp_unset_lock(&h->global);
    node_unlock(h, cur);
}

static void scratch_init(scratch_t* s, const hnsw_t* h) {
    memset(s, 0, sizeof(*s));
    s->tag = (unsigned short*)calloc(h->n, sizeof(unsigned short));
    s->sel = (pr_t*)malloc(1024 * sizeof(pr_t));
    s->sorted = (pr_t*)malloc((size_t)(h->efc + 1024) * sizeof(pr_t));
}
static void scratch_free(scratch_t* s) { free(s->tag); free(s->sel); free(s->sorted); free(s->cand.a); free(s->top.a); free(s->tmp.a); }

// ---- C interface (oracle/c_ref.py) ----------------------------------------------------------------------------------
void* hnsw_build(const float* data, size_t n, int dim, int metric, int M, int ef_construction, uint64_t seed, int nthreads) {
    if (M < 2 || M > 64 || n == 0 || n > 0xFFFFFFF0ull) return NULL;
    hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
    h->dim = dim; h->M = M; h->maxM0 = 2 * M; h->efc = ef_construction > M ? ef_construction : M; h->metric = metric;
    h->n = n; h->data = data; h->mult = 1.0 / log((double)M);
    h->level = (int*)malloc(n * sizeof(int));
    h->link0 = (unsigned*)calloc(n * (size_t)(h->maxM0 + 1), sizeof(unsigned));
    h->linkU = (unsigned**)calloc(n, sizeof(unsigned*));
    h->lock = (volatile int*)calloc(n, sizeof(int));
    omp_init_lock(&h->global);
    h->maxlevel = -1; h->enter = 0;
    for (size_t i = 0; i < n; ++i) {                            // getRandomLevel: floor(-ln(U) * mult)
        const double u = ((splitmix64(seed + i) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        const int lv = (int)(-log(u) * h->mult);
        h->level[i] = lv;
        if (lv > 0) h->linkU[i] = (unsigned*)calloc((size_t)lv * (M + 1), sizeof(unsigned));
    }
    if (nthreads < 1) nthreads = omp_get_num_procs();
    {   // the first point alone, then everybody (as hnswlib's ParallelFor over add_items does after item 0)
        scratch_t s; scratch_init(&s, h);
        add_point(h, &s, 0);
        scratch_free(&s);
    }
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s; scratch_init(&s, h);
#pragma omp for schedule(dynamic, 64)
        for (long i = 1; i < (long)n; ++i) add_point(h, &s, (unsigned)i);
        scratch_free(&s);
    }
    return h;
}

// searchKnn for nq queries (rows as the index holds them: normalised for cosine): labels [nq][k] (-1 padded), dist [nq][k]
void hnsw_search(void* hv, const float* q, size_t nq, int k, int ef, int nthreads, int64_t* labels, float* dists) {
    hnsw_t* h = (hnsw_t*)hv;
    if (ef < k) ef = k;
    if (nthreads < 1) nthreads = omp_get_num_procs();
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s; scratch_init(&s, h);
        unsigned nb[256];
#pragma omp for schedule(dynamic, 4)
        for (long qi = 0; qi < (long)nq; ++qi) {
            const float* qv = q + (size_t)qi * h->dim;
            unsigned ep = h->enter;
            float cd = dist(h, qv, h->data + (size_t)ep * h->dim);
            for (int level = h->maxlevel; level > 0; --level) {
                int changed = 1;
                while (changed) {
                    changed = 0;
                    const unsigned* l = links(h, ep, level);
                    const int cnt = (int)l[0];
                    memcpy(nb, l + 1, (size_t)cnt * sizeof(unsigned));
                    for (int j = 0; j < cnt; ++j) {
                        const float d = dist(h, qv, h->data + (size_t)nb[j] * h->dim);
                        if (d < cd) { cd = d; ep = nb[j]; changed = 1; }
                    }
                }
            }
            search_layer(h, &s, ep, qv, 0, ef, 0);
            while (s.top.n > k) heap_pop(&s.top, 1);
            const int m = s.top.n;
            for (int j = m; j < k; ++j) { labels[qi * k + j] = -1; dists[qi * k + j] = INFINITY; }
            for (int j = m - 1; j >= 0; --j) {
                const pr_t p = heap_pop(&s.top, 1);
                labels[qi * k + j] = (int64_t)p.id;
                dists[qi * k + j] = p.d;
            }
        }
        scratch_free(&s);
    }
}

int hnsw_max_level(void* hv) { return ((hnsw_t*)hv)->maxlevel; }
double hnsw_mean_degree0(void* hv) {
    hnsw_t* h = (hnsw_t*)hv;
    double s = 0;
    for (size_t i = 0; i < h->n; ++i) s += h->link0[i * (size_t)(h->maxM0 + 1)];
    return s / (double)h->n;
}
void hnsw_free(void* hv) {
    hnsw_t* h = (hnsw_t*)hv;
    if (!h) return;
    for (size_t i = 0; i < h->n; ++i) free(h->linkU[i]);
    free(h->linkU); free(h->link0); free(h->level); free((void*)h->lock);
    omp_destroy_lock(&h->global);
    free(h);
}
