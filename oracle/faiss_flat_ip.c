/*
 * oracle/faiss_flat_ip.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the four faiss-cpu calls that form MiniVectorDB's hot
 * path (ref: minivectordb/vector_database.py:43-46, 475, 497, 511-514 and the
 * identical call sites at minivectordb/sharded_vector_database.py:80-83, 604,
 * 626, 639-642):
 *
 *     faiss.normalize_L2(x)        -> orc_renorm_L2
 *     faiss.IndexFlatIP(d).add(x)  -> the caller's row-major float32 matrix
 *     index.search(q, k)           -> orc_search_flat_ip
 *     gather + temp index + search -> orc_search_gathered   (VDB:510-514)
 *
 * PARITY UNPINNED: the arithmetic of this path lives in `faiss-cpu`
 * (ref: requirements.txt:7, pyproject.toml:22 -- un-vendored, un-pinned PyPI
 * dependency; PyPI's current release when the reference pinned numpy<2 was
 * faiss-cpu 1.8.0).  faiss is neither in /root/reference nor installable in
 * this image, and the reference's tests hold no numeric golden vector for
 * the scan.  This file therefore restates faiss's
 * published algorithm (faiss/utils/distances.cpp: fvec_renorm_L2,
 * exhaustive_inner_product_seq; faiss/utils/Heap.h: CMin heap with id
 * tie-break; faiss/impl/ResultHandler.h: Top1 / Heap / Reservoir handlers with
 * faiss's reservoir capacity (2k+15)&~15 and partition_fuzzy; exhaustive_inner_product_blas:
 * 4096 x 1024 sgemm blocks feeding the block result handlers for nq >= 20)
 * from its documented behaviour.  It is anchored by (a) a float64 numpy gold
 * scorer (oracle/oracle.py) and (b) fixtures produced by running the
 * reference's own Python classes on top of this file (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 -fopenmp -shared -fPIC).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* elementary vector kernels (faiss/utils/distances_simd.cpp)          */
/* ------------------------------------------------------------------ */

/* fvec_inner_product: fp32 multiply-accumulate.  faiss compiles this loop
 * with "imprecise loop" pragmas so the compiler may re-associate into SIMD
 * lanes; the summation order is therefore build dependent.  We allow the
 * same freedom through `omp simd reduction`. */
static inline float ip_f32(const float* a, const float* b, size_t d) {
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (size_t i = 0; i < d; i++) s += a[i] * b[i];
    return s;
}

ORC_API float orc_inner_product(const float* a, const float* b, int64_t d) {
    return ip_f32(a, b, (size_t)d);
}

/* fvec_renorm_L2: per row nr = sum x^2 (fp32); if nr > 0 scale the row by
 * (float)(1.0 / sqrtf(nr)).  Zero rows are left untouched.
 * (ref call sites: VDB:45 whole matrix in place, VDB:475 the query) */
ORC_API void orc_renorm_L2(int64_t d, int64_t n, float* x) {
#pragma omp parallel for schedule(static) if (n > 4096)
    for (int64_t i = 0; i < n; i++) {
        float* row = x + i * d;
        float nr = ip_f32(row, row, (size_t)d);
        if (nr > 0) {
            const float inv = (float)(1.0 / sqrtf(nr));
            for (int64_t j = 0; j < d; j++) row[j] *= inv;
        }
    }
}

/* ------------------------------------------------------------------ */
/* result handlers                                                     */
/* ------------------------------------------------------------------ */

/* CMin<float,int64> ordering with the id tie-break faiss uses since 1.7.3:
 * (v1,i1) precedes (v2,i2) iff v1 < v2, or v1 == v2 and i1 < i2. */
static inline int pair_lt(float v1, int64_t i1, float v2, int64_t i2) {
    return (v1 < v2) || (v1 == v2 && i1 < i2);
}

/* Binary min-heap over k slots (root = weakest kept result).  Replace the
 * root by (v,id) and sift down. */
static void heap_replace_root(int64_t k, float* hv, int64_t* hi, float v, int64_t id) {
    int64_t pos = 0;
    for (;;) {
        int64_t l = 2 * pos + 1, r = l + 1, c;
        if (l >= k) break;
        c = (r < k && pair_lt(hv[r], hi[r], hv[l], hi[l])) ? r : l;
        if (pair_lt(v, id, hv[c], hi[c])) break;
        hv[pos] = hv[c];
        hi[pos] = hi[c];
        pos = c;
    }
    hv[pos] = v;
    hi[pos] = id;
}

/* heap_reorder: pop the root repeatedly into the tail, which leaves the
 * array sorted best-first; unfilled slots (id -1, value -FLT_MAX) end last. */
static void heap_drain_sorted(int64_t k, float* hv, int64_t* hi) {
    for (int64_t live = k; live > 1; live--) {
        float v0 = hv[0];
        int64_t i0 = hi[0];
        float vl = hv[live - 1];
        int64_t il = hi[live - 1];
        heap_replace_root(live - 1, hv, hi, vl, il);
        hv[live - 1] = v0;
        hi[live - 1] = i0;
    }
}

/* heap_push: append (v,id) as element number `k` (1-based) of a heap that holds k-1 and sift it up
 * (faiss/utils/Heap.h heap_push, min-heap flavour with the cmp2 id tie-break). */
static void heap_push_up(int64_t k, float* hv, int64_t* hi, float v, int64_t id) {
    int64_t i = k - 1;
    while (i > 0) {
        int64_t f = (i - 1) / 2;
        if (!pair_lt(v, id, hv[f], hi[f])) break;
        hv[i] = hv[f];
        hi[i] = hi[f];
        i = f;
    }
    hv[i] = v;
    hi[i] = id;
}

/* ------------------------------------------------------------------ */
/* ReservoirTopN (faiss/impl/ResultHandler.h) for C = CMin<float,int64>: */
/* "better" = larger.  Storage of `cap` = (2k+15)&~15 slots; a value is  */
/* stored iff it is STRICTLY better than the threshold; when the storage */
/* is full, partition_fuzzy keeps between k and (cap+k)/2 of the best and */
/* raises the threshold to the partition value.                          */
/* ------------------------------------------------------------------ */
ORC_API int64_t orc_reservoir_capacity(int64_t k) { return (2 * k + 15) & ~(int64_t)15; }

static float median3f(float a, float b, float c) {
    if (a > b) { float t = a; a = b; b = t; }
    if (c > b) return b;
    if (c > a) return c;
    return a;
}

/* faiss/utils/partitioning.cpp: partition_fuzzy_median3 restated for CMin (threshold found by
 * bisection over medians of three samples, then the array is compressed in place, order kept).
 * Keeps every value > thresh plus enough values == thresh to reach q in [q_min, q_max]. */
static float partition_fuzzy_cmin(float* vals, int64_t* ids, int64_t n, int64_t q_min, int64_t q_max, int64_t* q_out) {
    if (q_min == 0) { *q_out = 0; return FLT_MAX; }
    if (q_max >= n) { *q_out = q_max; return -FLT_MAX; }
    float thresh_inf = FLT_MAX;     /* C::Crev::neutral(): nothing is better than it */
    float thresh_sup = -FLT_MAX;    /* C::neutral(): everything is better than it   */
    float thresh = median3f(vals[0], vals[n / 2], vals[n - 1]);
    int64_t n_eq = 0, n_lt = 0, q = 0;
    for (int it = 0; it < 200; it++) {
        n_eq = n_lt = 0;
        for (int64_t i = 0; i < n; i++) {
            if (thresh < vals[i]) n_lt++;            /* C::cmp(thresh, v): v is better */
            else if (vals[i] == thresh) n_eq++;
        }
        if (n_lt <= q_min) {
            if (n_lt + n_eq >= q_min) { q = q_min; break; }
            thresh_inf = thresh;
        } else if (n_lt <= q_max) {
            q = n_lt;
            break;
        } else {
            thresh_sup = thresh;
        }
        /* sample_threshold_median3: three values strictly between the bounds, array walked with a prime stride */
        float v3[3];
        int vi = 0;
        for (int64_t i = 0; i < n; i++) {
            float v = vals[(uint64_t)(i * 6700417ULL) % (uint64_t)n];
            if (v < thresh_inf && thresh_sup < v) {
                v3[vi++] = v;
                if (vi == 3) break;
            }
        }
        float nt = vi == 3 ? median3f(v3[0], v3[1], v3[2]) : vi != 0 ? v3[0] : thresh_inf;
        if (nt == thresh_inf) break;   /* nothing between the bounds */
        thresh = nt;
    }
    int64_t n_eq_1 = q - n_lt;
    if (n_eq_1 < 0) {   /* more than q values strictly better even at the bound */
        q = q_min;
        thresh = nextafterf(thresh, HUGE_VALF);
        n_eq_1 = q;
    }
    int64_t wp = 0;     /* compress_array */
    for (int64_t i = 0; i < n; i++) {
        if (thresh < vals[i]) { vals[wp] = vals[i]; ids[wp] = ids[i]; wp++; }
        else if (n_eq_1 > 0 && vals[i] == thresh) { vals[wp] = vals[i]; ids[wp] = ids[i]; wp++; n_eq_1--; }
    }
    *q_out = wp;
    return thresh;
}

typedef struct { float* vals; int64_t* ids; int64_t i, n, cap; float thr; } reservoir_t;

static inline void reservoir_add(reservoir_t* r, float v, int64_t id) {
    if (r->thr < v) {
        if (r->i == r->cap) r->thr = partition_fuzzy_cmin(r->vals, r->ids, r->cap, r->n, (r->cap + r->n) / 2, &r->i);
        r->vals[r->i] = v;
        r->ids[r->i] = id;
        r->i++;
    }
}

/* ReservoirTopN::to_result: the first min(i,n) stored values are pushed into a heap, the rest replace its
 * root when strictly better, heap_reorder sorts best-first; missing results are (-FLT_MAX, -1). */
static void reservoir_to_result(const reservoir_t* r, float* D, int64_t* I) {
    const int64_t n = r->n, m = r->i < n ? r->i : n;
    for (int64_t j = 0; j < m; j++) heap_push_up(j + 1, D, I, r->vals[j], r->ids[j]);
    if (r->i < n) {
        heap_drain_sorted(m, D, I);
        for (int64_t j = m; j < n; j++) { D[j] = -FLT_MAX; I[j] = -1; }
    } else {
        for (int64_t j = n; j < r->i; j++)
            if (D[0] < r->vals[j]) heap_replace_root(n, D, I, r->vals[j], r->ids[j]);
        heap_drain_sorted(n, D, I);
    }
}

/* One query against rows [0,n).  Handler choice mirrors faiss:
 * k == 1 -> running maximum, k < 100 -> heap, k >= 100 -> ReservoirTopN
 * (distance_compute_min_k_reservoir = 100).  All three admit a row only if
 * its score is STRICTLY greater than the current threshold, so at an exact
 * tie on the boundary the earlier row is kept. */
static void search_one(const float* x, int64_t n, int64_t d, const float* q,
                       int64_t k, float* D, int64_t* I) {
    for (int64_t j = 0; j < k; j++) { D[j] = -FLT_MAX; I[j] = -1; }
    if (k == 1) {
        float best = -FLT_MAX;
        int64_t arg = -1;
        for (int64_t r = 0; r < n; r++) {
            float s = ip_f32(q, x + r * d, (size_t)d);
            if (s > best) { best = s; arg = r; }
        }
        D[0] = best; I[0] = arg;
        return;
    }
    if (k < 100) {
        for (int64_t r = 0; r < n; r++) {
            float s = ip_f32(q, x + r * d, (size_t)d);
            if (s > D[0]) heap_replace_root(k, D, I, s, r);
        }
        heap_drain_sorted(k, D, I);
        return;
    }
    reservoir_t res;
    res.n = k;
    res.cap = orc_reservoir_capacity(k);
    res.i = 0;
    res.thr = -FLT_MAX;
    res.vals = (float*)malloc((size_t)res.cap * sizeof(float));
    res.ids = (int64_t*)malloc((size_t)res.cap * sizeof(int64_t));
    for (int64_t r = 0; r < n; r++) reservoir_add(&res, ip_f32(q, x + r * d, (size_t)d), r);
    reservoir_to_result(&res, D, I);
    free(res.vals);
    free(res.ids);
}

/* ------------------------------------------------------------------ */
/* Block result handlers for the nq >= 20 path                         */
/* (faiss/utils/distances.cpp exhaustive_inner_product_blas: blocks of  */
/* 4096 queries x 1024 rows through sgemm_, every block of inner        */
/* products handed to Top1 / Heap / Reservoir BlockResultHandler::      */
/* add_results(j0, j1, ip_block), end_multiple() after the last block). */
/* The handler state lives in caller-owned arrays so that the sgemm can */
/* be the host BLAS (oracle.py calls numpy's sgemm, as faiss calls      */
/* sgemm_): vals [nq][cap], ids [nq][cap], cnt [nq], thr [nq].          */
/*   kind 0 Top1 (cap 1), 1 Heap (cap k), 2 Reservoir (cap from        */
/*   orc_reservoir_capacity).  Initialise cnt = 0, thr = -FLT_MAX,      */
/*   vals = -FLT_MAX, ids = -1.                                         */
/* ------------------------------------------------------------------ */
ORC_API int orc_block_add(int kind, int64_t k, int64_t cap, int64_t nq, int64_t j0, int64_t j1, const float* ip,
                          float* vals, int64_t* ids, int64_t* cnt, float* thr) {
    if (kind < 0 || kind > 2 || k <= 0 || cap < 1 || j1 < j0) return -1;
    const int64_t nj = j1 - j0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; i++) {
        const float* row = ip + i * nj;
        float* v = vals + i * cap;
        int64_t* id = ids + i * cap;
        if (kind == 0) {
            for (int64_t j = 0; j < nj; j++)
                if (row[j] > v[0]) { v[0] = row[j]; id[0] = j0 + j; }
        } else if (kind == 1) {
            for (int64_t j = 0; j < nj; j++)
                if (row[j] > v[0]) heap_replace_root(k, v, id, row[j], j0 + j);
        } else {
            reservoir_t r = {v, id, cnt[i], k, cap, thr[i]};
            for (int64_t j = 0; j < nj; j++) reservoir_add(&r, row[j], j0 + j);
            cnt[i] = r.i;
            thr[i] = r.thr;
        }
    }
    return 0;
}

ORC_API int orc_block_end(int kind, int64_t k, int64_t cap, int64_t nq, float* vals, int64_t* ids, const int64_t* cnt,
                          const float* thr, float* D, int64_t* I) {
    if (kind < 0 || kind > 2 || k <= 0) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; i++) {
        float* v = vals + i * cap;
        int64_t* id = ids + i * cap;
        float* Di = D + i * k;
        int64_t* Ii = I + i * k;
        if (kind == 0) {
            Di[0] = v[0]; Ii[0] = id[0];
        } else if (kind == 1) {
            heap_drain_sorted(k, v, id);
            memcpy(Di, v, (size_t)k * sizeof(float));
            memcpy(Ii, id, (size_t)k * sizeof(int64_t));
        } else {
            reservoir_t r = {v, id, cnt[i], k, cap, thr[i]};
            for (int64_t j = 0; j < k; j++) { Di[j] = -FLT_MAX; Ii[j] = -1; }
            reservoir_to_result(&r, Di, Ii);
        }
    }
    return 0;
}

/* IndexFlatIP.search for nq queries (exhaustive_inner_product_seq: OpenMP
 * over QUERIES only, each query scans all rows sequentially on one thread).
 * nthreads <= 0 -> OpenMP default.  D: nq*k float32 best-first, I: nq*k int64
 * row numbers; missing results are (-FLT_MAX, -1).
 * (ref call sites: VDB:497, SVDB:626) */
ORC_API int orc_search_flat_ip(const float* x, int64_t n, int64_t d, const float* q,
                               int64_t nq, int64_t k, float* D, int64_t* I, int nthreads) {
    if (k <= 0 || d <= 0 || n < 0 || nq < 0) return -1;
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
#endif
    for (int64_t i = 0; i < nq; i++)
        search_one(x, n, d, q + i * d, k, D + i * k, I + i * k);
    return 0;
}

/* The reference's filtered branch (VDB:508-523 / SVDB:634-649): copy the
 * admissible rows, in the order given, into a fresh matrix ("temp index"),
 * search it, and return positions INTO THE GATHERED LIST (the caller maps
 * them back exactly as VDB:521 does).  `scratch` must hold m*d floats. */
ORC_API int orc_search_gathered(const float* x, int64_t n, int64_t d, const int64_t* rows,
                                int64_t m, const float* q, int64_t nq, int64_t k,
                                float* D, int64_t* I, float* scratch, int nthreads) {
    if (k <= 0 || d <= 0 || m < 0) return -1;
    for (int64_t j = 0; j < m; j++) {
        if (rows[j] < 0 || rows[j] >= n) return -2;
        memcpy(scratch + j * d, x + rows[j] * d, (size_t)d * sizeof(float));
    }
    return orc_search_flat_ip(scratch, m, d, q, nq, k, D, I, nthreads);
}

/* ------------------------------------------------------------------ */
/* counter-based synthetic data                 */
/* ------------------------------------------------------------------ */

/* Integer-only generator so that CPU and GPU produce IDENTICAL bits:
 * z = splitmix64-finalise((row<<20 | col) + seed*golden).
 *   dist 0 ("bell"):    (sum of the four 16-bit fields of z - 131070) / 65536
 *                       (Irwin-Hall(4), zero mean, var 1/3; exact in fp32)
 *   dist 1 ("uniform"): top 24 bits of z * 2^-24  in [0,1)   (what the
 *                       reference tests use: np.random.rand, ref
 *                       tests/test_multithreaded_operations.py:13) */
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* two clones: AVX-512 (native 64-bit multiplies, 2-3x faster on the hosts that have it) and the
 * build's baseline; the resolver picks at load time.  Integer arithmetic only: same bits either way. */
__attribute__((target_clones("arch=x86-64-v4", "default")))
ORC_API void orc_synth_rows(uint64_t seed, int64_t row0, int64_t n, int64_t d, int dist, float* out) {
#pragma omp parallel for schedule(static) if (n * d > (1 << 16))
    for (int64_t r = 0; r < n; r++) {
        uint64_t base = ((uint64_t)(row0 + r) << 20) + seed * 0x9E3779B97F4A7C15ULL;
        for (int64_t c = 0; c < d; c++) {
            uint64_t z = mix64(base + (uint64_t)c);
            float v;
            if (dist == 0) {
                int32_t s = (int32_t)(z & 0xFFFF) + (int32_t)((z >> 16) & 0xFFFF) +
                            (int32_t)((z >> 32) & 0xFFFF) + (int32_t)(z >> 48);
                v = (float)(s - 131070) * (1.0f / 65536.0f);
            } else {
                v = (float)(z >> 40) * (1.0f / 16777216.0f);
            }
            out[r * d + c] = v;
        }
    }
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
