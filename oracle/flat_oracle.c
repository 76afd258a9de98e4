/*
 * flat_oracle.c -- CPU restatement of the flat (exact) vector-search path that
 * luzbetak/rag-faiss-embedding delegates to faiss-cpu IndexFlatL2 / IndexFlatIP.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (rag-faiss-embedding_b200/) never calls into this file and has no CPU fallback.
 *
 * PARITY STATUS: "parity unpinned" for search results.  The arithmetic lives in the
 * third-party wheel faiss-cpu (requirements.txt:13, version unpinned), which is absent from
 * /root/reference and cannot be installed here (no wheel, no network).  The reference has
 * no tests and no golden search outputs.  What IS pinned:
 *   - the on-disk layout, byte for byte, by the reference's own FAISS-written file
 *     data/faiss_index.bin (IxF2, d=384, ntotal=23) -> orc_read_index/orc_write_index;
 *   - search answers on that file against float64 brute force (tests/golden/).
 * The algorithm below restates upstream faiss as published (IndexFlat.cpp,
 * utils/distances.cpp, utils/Heap.h, impl/ResultHandler.h, impl/index_write.cpp):
 *   - call sites it must serve: faiss_store.py:29 (ctor), :46 (add), :64 (search),
 *     :91 (write_index), :106 (read_index); rag_datastore_manager.py:138,173,186,205,218.
 *   - nq <  20: exhaustive_L2sqr_seq / exhaustive_inner_product_seq -- per (query,row)
 *               exact-difference sum((x-q)^2) in fp32, results pushed into a per-query heap;
 *   - nq >= 20: exhaustive_L2sqr_blas -- blocks of 4096 queries x 1024 rows, inner products
 *               by sgemm, dis = |q|^2 + |x|^2 - 2 ip, negative values clamped to 0;
 *   - k < 100 : binary max-heap (L2) / min-heap (IP), replace top only if strictly better;
 *     k >= 100: reservoir of capacity 2k with threshold shrink (same result set);
 *   - output ascending (L2) / descending (IP); missing results label -1, distance
 *     +FLT_MAX (L2) / -FLT_MAX (IP); NaN distances never enter.
 * Ordering among exactly equal distances is normalised to ascending id (faiss does not
 * define it; BASELINE.json exempts exact ties).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_METRIC_IP 0
#define ORC_METRIC_L2 1
#define ORC_BLAS_THRESHOLD 20 /* faiss distance_compute_blas_threshold */
#define ORC_BS_Q 4096         /* faiss distance_compute_blas_query_bs */
#define ORC_BS_B 1024         /* faiss distance_compute_blas_database_bs */

#if defined(__x86_64__) && defined(__GNUC__) && !defined(ORC_NO_CLONES)
#define ORC_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define ORC_CLONES
#endif

/* ------------------------------------------------------------------------------------------
 * Synthetic data: counter-based, bit-identical on host and device (no transcendentals).
 * element(seed, idx) = (sum of 12 hashed 16-bit uniforms - 393210) / 65536  ~ N(0,1) (Irwin-Hall).
 * The CUDA twin is rag-faiss-embedding_b200/csrc/synth.cuh.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t orc_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
    z ^= z >> 27; z *= 0x94d049bb133111ebULL;
    z ^= z >> 31; return z;
}
static inline int32_t orc_synth_int(uint64_t seed, uint64_t idx) {
    uint64_t a = orc_mix64(seed + 0x9E3779B97F4A7C15ULL * (2 * idx + 1));
    uint64_t b = orc_mix64(a ^ 0xD1B54A32D192ED03ULL);
    uint64_t c = orc_mix64(b + idx);
    uint32_t s = 0;
    for (int i = 0; i < 4; i++) {
        s += (uint32_t)((a >> (16 * i)) & 0xffff);
        s += (uint32_t)((b >> (16 * i)) & 0xffff);
        s += (uint32_t)((c >> (16 * i)) & 0xffff);
    }
    return (int32_t)s - 393210;
}
/* rows [row0, row0+nrows) of the synthetic matrix with `d` columns; normalize!=0 divides each
 * row by its L2 norm (sum of squares exact in int64, division in double, rounded once). */
void orc_synth_rows(uint64_t seed, int64_t row0, int64_t nrows, int d, int normalize, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++) {
        uint64_t base = (uint64_t)(row0 + r) * (uint64_t)d;
        float* o = out + r * (int64_t)d;
        if (!normalize) {
            for (int j = 0; j < d; j++) o[j] = (float)orc_synth_int(seed, base + j) * (1.0f / 65536.0f);
        } else {
            int64_t ss = 0;
            for (int j = 0; j < d; j++) { int64_t v = orc_synth_int(seed, base + j); ss += v * v; }
            double nrm = sqrt((double)ss);
            for (int j = 0; j < d; j++) {
                double v = (double)orc_synth_int(seed, base + j);
                o[j] = (ss > 0) ? (float)(v / nrm) : 0.0f;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Distance primitives (fvec_L2sqr / fvec_inner_product / fvec_norm_L2sqr restated):
 * 16 fp32 partial sums (the AVX-512 lane structure), fixed pairwise reduction -> the result
 * does not depend on which clone the CPU dispatches to.
 * ---------------------------------------------------------------------------------------- */
static inline float orc_reduce16(const float* a) {
    float b8[8], b4[4];
    for (int i = 0; i < 8; i++) b8[i] = a[i] + a[i + 8];
    for (int i = 0; i < 4; i++) b4[i] = b8[i] + b8[i + 4];
    return (b4[0] + b4[2]) + (b4[1] + b4[3]);
}
ORC_CLONES
static float orc_l2sqr(const float* x, const float* y, int d) {
    float acc[16] = {0};
    int j = 0;
    for (; j + 16 <= d; j += 16)
        for (int l = 0; l < 16; l++) { float t = x[j + l] - y[j + l]; acc[l] += t * t; }
    for (int l = 0; j < d; j++, l++) { float t = x[j] - y[j]; acc[l] += t * t; }
    return orc_reduce16(acc);
}
ORC_CLONES
static float orc_ip(const float* x, const float* y, int d) {
    float acc[16] = {0};
    int j = 0;
    for (; j + 16 <= d; j += 16)
        for (int l = 0; l < 16; l++) acc[l] += x[j + l] * y[j + l];
    for (int l = 0; j < d; j++, l++) acc[l] += x[j] * y[j];
    return orc_reduce16(acc);
}
float orc_fvec_L2sqr(const float* x, const float* y, int d) { return orc_l2sqr(x, y, d); }
float orc_fvec_inner_product(const float* x, const float* y, int d) { return orc_ip(x, y, d); }

/* ------------------------------------------------------------------------------------------
 * Result handlers.  Keys are "smaller is better": key = dis (L2) or -ip (IP).
 * Total order (key, id) so ties resolve to the lower id, as faiss's strict '<' does for a
 * sequential scan.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float key; int64_t id; } orc_ent;

static inline int orc_less(float ka, int64_t ia, float kb, int64_t ib) {
    return ka < kb || (ka == kb && (uint64_t)ia < (uint64_t)ib); /* id -1 sorts last */
}
/* max-heap on (key,id): root = worst kept element (faiss CMax heap) */
static void orc_heap_sift_down(orc_ent* h, int64_t k, int64_t i) {
    orc_ent v = h[i];
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m;
        if (l >= k) break;
        m = (r < k && orc_less(h[l].key, h[l].id, h[r].key, h[r].id)) ? r : l;
        if (!orc_less(v.key, v.id, h[m].key, h[m].id)) break;
        h[i] = h[m]; i = m;
    }
    h[i] = v;
}
static int orc_ent_cmp(const void* a, const void* b) {
    const orc_ent* x = (const orc_ent*)a; const orc_ent* y = (const orc_ent*)b;
    if (orc_less(x->key, x->id, y->key, y->id)) return -1;
    if (orc_less(y->key, y->id, x->key, x->id)) return 1;
    return 0;
}

typedef struct {
    int64_t k;
    int use_reservoir;   /* k >= 100 in faiss */
    orc_ent* buf;        /* heap: k entries; reservoir: 2k entries */
    int64_t n;           /* reservoir fill */
    int64_t cap;
    float thr; int64_t thr_id; /* reservoir admission threshold */
} orc_handler;

static void orc_handler_init(orc_handler* h, int64_t k, orc_ent* storage) {
    h->k = k; h->use_reservoir = (k >= 100); h->buf = storage;
    h->n = 0; h->cap = 2 * k; h->thr = FLT_MAX; h->thr_id = -1;
    if (h->use_reservoir) { /* filled lazily */ }
    else for (int64_t i = 0; i < k; i++) { storage[i].key = FLT_MAX; storage[i].id = -1; }
}
/* quickselect: after return the k smallest (by (key,id)) occupy buf[0..k) */
static void orc_select_k(orc_ent* a, int64_t n, int64_t k) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        orc_ent p = a[lo + (hi - lo) / 2];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (orc_less(a[i].key, a[i].id, p.key, p.id)) i++;
            while (orc_less(p.key, p.id, a[j].key, a[j].id)) j--;
            if (i <= j) { orc_ent t = a[i]; a[i] = a[j]; a[j] = t; i++; j--; }
        }
        if (k - 1 <= j) hi = j; else if (k - 1 >= i) lo = i; else break;
    }
}
static inline void orc_handler_add(orc_handler* h, float key, int64_t id) {
    if (h->use_reservoir) {
        if (!orc_less(key, id, h->thr, h->thr_id)) return; /* NaN fails too */
        h->buf[h->n].key = key; h->buf[h->n].id = id; h->n++;
        if (h->n == h->cap) { /* shrink to k, new threshold = current k-th */
            orc_select_k(h->buf, h->n, h->k);
            orc_ent w = h->buf[0];
            for (int64_t i = 1; i < h->k; i++) if (orc_less(w.key, w.id, h->buf[i].key, h->buf[i].id)) w = h->buf[i];
            h->thr = w.key; h->thr_id = w.id; h->n = h->k;
        }
    } else {
        if (!orc_less(key, id, h->buf[0].key, h->buf[0].id)) return;
        h->buf[0].key = key; h->buf[0].id = id;
        orc_heap_sift_down(h->buf, h->k, 0);
    }
}
static void orc_handler_finish(orc_handler* h, int metric, float* D, int64_t* I) {
    int64_t n = h->use_reservoir ? h->n : h->k;
    qsort(h->buf, (size_t)n, sizeof(orc_ent), orc_ent_cmp);
    for (int64_t i = 0; i < h->k; i++) {
        if (i < n && h->buf[i].id >= 0) {
            D[i] = metric == ORC_METRIC_L2 ? h->buf[i].key : -h->buf[i].key; I[i] = h->buf[i].id;
        } else { D[i] = metric == ORC_METRIC_L2 ? FLT_MAX : -FLT_MAX; I[i] = -1; }
    }
}

/* ------------------------------------------------------------------------------------------
 * exhaustive_*_seq: parallel over queries only (nq=1 is single-threaded, as in faiss).
 * ---------------------------------------------------------------------------------------- */
static void orc_search_seq(const float* xb, int64_t nb, int d, int metric, const float* xq, int64_t nq,
                           int64_t k, float* D, int64_t* I) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t q = 0; q < nq; q++) {
        int64_t cap = k >= 100 ? 2 * k : k;
        orc_ent* st = (orc_ent*)malloc(sizeof(orc_ent) * (size_t)cap);
        orc_handler h; orc_handler_init(&h, k, st);
        const float* qv = xq + q * (int64_t)d;
        for (int64_t j = 0; j < nb; j++) {
            const float* xv = xb + j * (int64_t)d;
            float key = metric == ORC_METRIC_L2 ? orc_l2sqr(qv, xv, d) : -orc_ip(qv, xv, d);
            orc_handler_add(&h, key, j);
        }
        orc_handler_finish(&h, metric, D + q * k, I + q * k);
        free(st);
    }
}

/* ------------------------------------------------------------------------------------------
 * sgemm restated: C[i,j] = sum_l A[i,l] * B[j,l]  (both row-major, K contiguous) --
 * 4x4 register block over 16-lane partial sums, parallel over the query rows of the block.
 * ---------------------------------------------------------------------------------------- */
ORC_CLONES
static void orc_gemm_nt_rows(const float* A, const float* B, float* C, int64_t i0, int64_t i1, int64_t nbj,
                             int d, int64_t ldc) {
    for (int64_t i = i0; i < i1; i++) {
        const float* a = A + i * (int64_t)d;
        int64_t j = 0;
        for (; j + 4 <= nbj; j += 4) {
            const float *b0 = B + j * (int64_t)d, *b1 = b0 + d, *b2 = b1 + d, *b3 = b2 + d;
            float s0[16] = {0}, s1[16] = {0}, s2[16] = {0}, s3[16] = {0};
            int l = 0;
            for (; l + 16 <= d; l += 16)
                for (int t = 0; t < 16; t++) {
                    float av = a[l + t];
                    s0[t] += av * b0[l + t]; s1[t] += av * b1[l + t];
                    s2[t] += av * b2[l + t]; s3[t] += av * b3[l + t];
                }
            for (int t = 0; l < d; l++, t++) {
                float av = a[l];
                s0[t] += av * b0[l]; s1[t] += av * b1[l]; s2[t] += av * b2[l]; s3[t] += av * b3[l];
            }
            C[i * ldc + j] = orc_reduce16(s0); C[i * ldc + j + 1] = orc_reduce16(s1);
            C[i * ldc + j + 2] = orc_reduce16(s2); C[i * ldc + j + 3] = orc_reduce16(s3);
        }
        for (; j < nbj; j++) C[i * ldc + j] = orc_ip(a, B + j * (int64_t)d, d);
    }
}

/* exhaustive_L2sqr_blas / exhaustive_inner_product_blas */
static void orc_search_blas(const float* xb, int64_t nb, int d, int metric, const float* xq, int64_t nq,
                            int64_t k, float* D, int64_t* I) {
    int64_t cap = k >= 100 ? 2 * k : k;
    orc_ent* st = (orc_ent*)malloc(sizeof(orc_ent) * (size_t)(cap * nq));
    orc_handler* hs = (orc_handler*)malloc(sizeof(orc_handler) * (size_t)nq);
    float* qn = (float*)malloc(sizeof(float) * (size_t)nq);
    float* xn = (float*)malloc(sizeof(float) * (size_t)(nb > 0 ? nb : 1));
    float* ipb = (float*)malloc(sizeof(float) * (size_t)ORC_BS_Q * ORC_BS_B);
    for (int64_t q = 0; q < nq; q++) orc_handler_init(&hs[q], k, st + q * cap);
    if (metric == ORC_METRIC_L2) {
#pragma omp parallel for schedule(static)
        for (int64_t q = 0; q < nq; q++) qn[q] = orc_ip(xq + q * (int64_t)d, xq + q * (int64_t)d, d);
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < nb; j++) xn[j] = orc_ip(xb + j * (int64_t)d, xb + j * (int64_t)d, d);
    }
    for (int64_t i0 = 0; i0 < nq; i0 += ORC_BS_Q) {
        int64_t i1 = i0 + ORC_BS_Q < nq ? i0 + ORC_BS_Q : nq;
        for (int64_t j0 = 0; j0 < nb; j0 += ORC_BS_B) {
            int64_t j1 = j0 + ORC_BS_B < nb ? j0 + ORC_BS_B : nb;
#pragma omp parallel for schedule(dynamic, 8)
            for (int64_t i = i0; i < i1; i++) {
                orc_gemm_nt_rows(xq, xb + j0 * (int64_t)d, ipb - i0 * ORC_BS_B, i, i + 1, j1 - j0, d, ORC_BS_B);
                const float* row = ipb + (i - i0) * ORC_BS_B;
                for (int64_t j = j0; j < j1; j++) {
                    float key;
                    if (metric == ORC_METRIC_L2) {
                        key = qn[i] + xn[j] - 2.0f * row[j - j0];
                        if (key < 0) key = 0; /* faiss: negative values can occur for identical vectors */
                    } else key = -row[j - j0];
                    orc_handler_add(&hs[i], key, j);
                }
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < nq; q++) orc_handler_finish(&hs[q], metric, D + q * k, I + q * k);
    free(ipb); free(xn); free(qn); free(hs); free(st);
}

/* IndexFlat::search.  algo: 0 = faiss's own choice (blas iff nq >= 20), 1 = force seq, 2 = force blas.
 * nthreads <= 0 keeps the OpenMP default.  Returns 0, or -1 on bad arguments (faiss asserts k > 0). */
int orc_search(const float* xb, int64_t nb, int d, int metric, const float* xq, int64_t nq, int64_t k,
               float* D, int64_t* I, int algo, int nthreads) {
    if (k <= 0 || d <= 0 || nq < 0 || nb < 0 || (metric != ORC_METRIC_L2 && metric != ORC_METRIC_IP)) return -1;
#ifdef _OPENMP
    int saved = omp_get_max_threads();
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    int blas = algo == 2 || (algo == 0 && nq >= ORC_BLAS_THRESHOLD);
    if (blas) orc_search_blas(xb, nb, d, metric, xq, nq, k, D, I);
    else orc_search_seq(xb, nb, d, metric, xq, nq, k, D, I);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(saved);
#endif
    return 0;
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * write_index / read_index for IndexFlat (layout verified against the reference's own
 * data/faiss_index.bin):  fourcc | i32 d | i64 ntotal | i64 2^20 | i64 2^20 | u8 is_trained |
 * i32 metric_type | u64 nfloats | fp32 rows.
 * ---------------------------------------------------------------------------------------- */
int orc_write_index(const char* path, const float* xb, int64_t nb, int d, int metric) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    const char* cc = metric == ORC_METRIC_L2 ? "IxF2" : "IxFI";
    int32_t d32 = d, mt = metric; int64_t nt = nb, dummy = 1 << 20; uint8_t tr = 1; uint64_t sz = (uint64_t)nb * d;
    int ok = fwrite(cc, 1, 4, f) == 4 && fwrite(&d32, 4, 1, f) == 1 && fwrite(&nt, 8, 1, f) == 1 &&
             fwrite(&dummy, 8, 1, f) == 1 && fwrite(&dummy, 8, 1, f) == 1 && fwrite(&tr, 1, 1, f) == 1 &&
             fwrite(&mt, 4, 1, f) == 1 && fwrite(&sz, 8, 1, f) == 1 &&
             (sz == 0 || fwrite(xb, 4, sz, f) == sz);
    fclose(f);
    return ok ? 0 : -2;
}
/* two-step read: header first (out pointers may be NULL), then payload into caller memory */
int orc_read_index_header(const char* path, int* d, int64_t* ntotal, int* metric) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char cc[4]; int32_t d32, mt; int64_t nt, dm[2]; uint8_t tr; uint64_t sz;
    int ok = fread(cc, 1, 4, f) == 4 && fread(&d32, 4, 1, f) == 1 && fread(&nt, 8, 1, f) == 1 &&
             fread(dm, 8, 2, f) == 2 && fread(&tr, 1, 1, f) == 1 && fread(&mt, 4, 1, f) == 1;
    if (ok && mt > 1) { float arg; ok = fread(&arg, 4, 1, f) == 1; }
    ok = ok && fread(&sz, 8, 1, f) == 1;
    fclose(f);
    if (!ok) return -2;
    if (memcmp(cc, "IxF2", 4) && memcmp(cc, "IxFI", 4) && memcmp(cc, "IxFl", 4)) return -3;
    if (sz != (uint64_t)nt * (uint64_t)d32) return -4;
    if (d) *d = d32;
    if (ntotal) *ntotal = nt;
    if (metric) *metric = mt;
    return 0;
}
int orc_read_index_rows(const char* path, float* out, int64_t nfloats) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    int32_t mt;
    if (fseek(f, 33, SEEK_SET) || fread(&mt, 4, 1, f) != 1) { fclose(f); return -2; }
    long off = 45 + (mt > 1 ? 4 : 0);
    int ok = fseek(f, off, SEEK_SET) == 0 && (nfloats == 0 || fread(out, 4, (size_t)nfloats, f) == (size_t)nfloats);
    fclose(f);
    return ok ? 0 : -2;
}
