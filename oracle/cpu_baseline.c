/*
 * oracle/cpu_baseline.c -- timed CPU baseline for bench.py (cpu_baseline leg and
 * --impl reference).  TEST / MEASUREMENT INFRASTRUCTURE ONLY; never linked into
 * the product library.
 *
 * Compiled with FMA contraction enabled (unlike vs_oracle.c) because this file is
 * timed, not used as the bit-level checker; its results are themselves verified
 * against vs_oracle.c in tests/test_oracle_golden.py.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_L2 0
#define ORC_IP 1

typedef struct {
    int k, n;
    float *key;
    int64_t *id;
} orc_topk;

static inline int orc_better(float ka, int64_t ia, float kb, int64_t ib) {
    return ka < kb || (ka == kb && ia < ib);
}

static inline void orc_topk_push(orc_topk *t, float key, int64_t id) {
    if (t->n == t->k) {
        if (!orc_better(key, id, t->key[t->k - 1], t->id[t->k - 1])) return;
    } else {
        t->n++;
    }
    int j = t->n - 1;
    while (j > 0 && orc_better(key, id, t->key[j - 1], t->id[j - 1])) {
        t->key[j] = t->key[j - 1];
        t->id[j] = t->id[j - 1];
        j--;
    }
    t->key[j] = key;
    t->id[j] = id;
}

static inline float orc_ip(const float *x, const float *y, int d) {
    float s = 0;
    for (int j = 0; j < d; j++) s += x[j] * y[j];
    return s;
}

/* ------------------------------------------------------------------------- */
/* Timed CPU baseline: the reference's threading model -- one thread per part,*/
/* Faiss single-threaded inside a part (omp_set_num_threads(1),               */
/* Common/VIWithDataPart.h:350; ThreadPool over parts,                        */
/* Storages/MergeTreeSelectWithHybridSearchProcessor.cpp:1212-1241) -- with a  */
/* blocked SIMD inner-product kernel standing in for Faiss's sgemm path        */
/* (knn_inner_product / knn_L2sqr with nx >= 20 go through BLAS blocks of      */
/* 4096 x 1024).  Results are merged with the global tie rule.                */
/* ------------------------------------------------------------------------- */
#define ORC_BQ 4
#define ORC_BY 4

__attribute__((target_clones("avx512f", "avx2,fma", "default"))) static void orc_ip_block(
    const float *x, int nq, const float *y, int nyb, int d, float *out /* [nq][nyb] */) {
    int qi = 0;
    for (; qi + ORC_BQ <= nq; qi += ORC_BQ) {
        int yi = 0;
        for (; yi + ORC_BY <= nyb; yi += ORC_BY) {
            float acc[ORC_BQ][ORC_BY][16];
            memset(acc, 0, sizeof(acc));
            int j = 0;
            for (; j + 16 <= d; j += 16) {
                for (int a = 0; a < ORC_BQ; a++)
                    for (int b = 0; b < ORC_BY; b++) {
                        const float *xa = x + (size_t)(qi + a) * d + j;
                        const float *yb = y + (size_t)(yi + b) * d + j;
#pragma GCC unroll 16
                        for (int l = 0; l < 16; l++) acc[a][b][l] += xa[l] * yb[l];
                    }
            }
            for (int a = 0; a < ORC_BQ; a++)
                for (int b = 0; b < ORC_BY; b++) {
                    float s = 0;
                    for (int l = 0; l < 16; l++) s += acc[a][b][l];
                    for (int jj = j; jj < d; jj++) s += x[(size_t)(qi + a) * d + jj] * y[(size_t)(yi + b) * d + jj];
                    out[(size_t)(qi + a) * nyb + yi + b] = s;
                }
        }
        for (; yi < nyb; yi++)
            for (int a = 0; a < ORC_BQ; a++) out[(size_t)(qi + a) * nyb + yi] = orc_ip(x + (size_t)(qi + a) * d, y + (size_t)yi * d, d);
    }
    for (; qi < nq; qi++)
        for (int yi = 0; yi < nyb; yi++) out[(size_t)qi * nyb + yi] = orc_ip(x + (size_t)qi * d, y + (size_t)yi * d, d);
}

typedef struct {
    int metric;
    const float *x;
    int64_t nx;
    const float *y;
    int64_t y0, y1; /* this part's row range */
    int d, k;
    float *dis;   /* [nx][k] per part */
    int64_t *ids; /* global row ids */
} orc_part_job;

static void *orc_part_worker(void *arg) {
    orc_part_job *jb = (orc_part_job *)arg;
    const int k = jb->k, d = jb->d;
    const int64_t nx = jb->nx;
    const int YB = 256, QB = 64;
    float *keys = (float *)malloc(sizeof(float) * (size_t)k * (size_t)nx);
    int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)k * (size_t)nx);
    int *cnt = (int *)calloc((size_t)nx, sizeof(int));
    float *blk = (float *)malloc(sizeof(float) * (size_t)QB * YB);
    float *xn = NULL, *yn = (float *)malloc(sizeof(float) * YB);
    if (jb->metric == ORC_L2) {
        xn = (float *)malloc(sizeof(float) * (size_t)nx);
        for (int64_t q = 0; q < nx; q++) xn[q] = orc_ip(jb->x + q * d, jb->x + q * d, d);
    }
    for (int64_t y0 = jb->y0; y0 < jb->y1; y0 += YB) {
        int nyb = (int)(jb->y1 - y0 < YB ? jb->y1 - y0 : YB);
        const float *yb = jb->y + y0 * d;
        if (jb->metric == ORC_L2)
            for (int i = 0; i < nyb; i++) yn[i] = orc_ip(yb + (size_t)i * d, yb + (size_t)i * d, d);
        for (int64_t q0 = 0; q0 < nx; q0 += QB) {
            int nqb = (int)(nx - q0 < QB ? nx - q0 : QB);
            orc_ip_block(jb->x + q0 * d, nqb, yb, nyb, d, blk);
            for (int a = 0; a < nqb; a++) {
                int64_t q = q0 + a;
                orc_topk t = {k, cnt[q], keys + q * k, ids + q * k};
                for (int i = 0; i < nyb; i++) {
                    float s = blk[(size_t)a * nyb + i];
                    float key;
                    if (jb->metric == ORC_L2) {
                        /* faiss BLAS path: ||x||^2 + ||y||^2 - 2 x.y clamped at 0 */
                        key = xn[q] + yn[i] - 2 * s;
                        if (key < 0) key = 0;
                    } else {
                        key = -s;
                    }
                    if (t.n == k && !(key <= t.key[k - 1])) continue;
                    orc_topk_push(&t, key, y0 + i);
                }
                cnt[q] = t.n;
            }
        }
    }
    for (int64_t q = 0; q < nx; q++)
        for (int j = 0; j < k; j++) {
            if (j < cnt[q]) {
                jb->dis[q * k + j] = jb->metric == ORC_L2 ? keys[q * k + j] : -keys[q * k + j];
                jb->ids[q * k + j] = ids[q * k + j];
            } else {
                jb->dis[q * k + j] = jb->metric == ORC_L2 ? FLT_MAX : -FLT_MAX;
                jb->ids[q * k + j] = -1;
            }
        }
    free(keys);
    free(ids);
    free(cnt);
    free(blk);
    free(xn);
    free(yn);
    return NULL;
}

/* Multi-part batched brute force (metric L2 or IP; cosine callers pre-normalise).
 * The corpus is cut into n_parts equal row ranges, one pthread each. */
void orc_knn_flat_parts(int metric, const float *x, int64_t nx, const float *y, int64_t ny, int d, int k, int n_parts,
                        float *dis, int64_t *ids) {
    if (n_parts < 1) n_parts = 1;
    if (n_parts > ny && ny > 0) n_parts = (int)ny;
    orc_part_job *jobs = (orc_part_job *)calloc((size_t)n_parts, sizeof(orc_part_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_parts, sizeof(pthread_t));
    int64_t per = (ny + n_parts - 1) / n_parts;
    for (int p = 0; p < n_parts; p++) {
        jobs[p].metric = metric;
        jobs[p].x = x;
        jobs[p].nx = nx;
        jobs[p].y = y;
        jobs[p].y0 = per * p < ny ? per * p : ny;
        jobs[p].y1 = per * (p + 1) < ny ? per * (p + 1) : ny;
        jobs[p].d = d;
        jobs[p].k = k;
        jobs[p].dis = (float *)malloc(sizeof(float) * (size_t)k * (size_t)nx);
        jobs[p].ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)k * (size_t)nx);
        pthread_create(&th[p], NULL, orc_part_worker, &jobs[p]);
    }
    for (int p = 0; p < n_parts; p++) pthread_join(th[p], NULL);
    /* global merge (score, then smaller id) */
    float *key = (float *)malloc(sizeof(float) * (size_t)k);
    int64_t *id = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    for (int64_t q = 0; q < nx; q++) {
        orc_topk t = {k, 0, key, id};
        for (int p = 0; p < n_parts; p++)
            for (int j = 0; j < k; j++) {
                int64_t rid = jobs[p].ids[q * k + j];
                if (rid < 0) continue;
                float s = jobs[p].dis[q * k + j];
                orc_topk_push(&t, metric == ORC_L2 ? s : -s, rid);
            }
        for (int j = 0; j < k; j++) {
            if (j < t.n) {
                dis[q * k + j] = metric == ORC_L2 ? key[j] : -key[j];
                ids[q * k + j] = id[j];
            } else {
                dis[q * k + j] = metric == ORC_L2 ? FLT_MAX : -FLT_MAX;
                ids[q * k + j] = -1;
            }
        }
    }
    for (int p = 0; p < n_parts; p++) {
        free(jobs[p].dis);
        free(jobs[p].ids);
    }
    free(key);
    free(id);
    free(jobs);
    free(th);
}

/* ------------------------------------------------------------------------- */
/* BLAS form of the same baseline.  Faiss routes knn_inner_product/knn_L2sqr  */
/* with nx >= 20 through sgemm in blocks of 4096 queries x 1024 base rows     */
/* (published behaviour of faiss/utils/distances.cpp); the reference reaches  */
/* it from BruteForceSearch.h:77-88 with omp_set_num_threads(1), i.e. one     */
/* single-threaded sgemm stream per part.  The sgemm comes from the OpenBLAS  */
/* that numpy bundles (dlopen'ed; ILP64 cblas interface), so the CPU arm gets  */
/* a vendor-tuned AVX-512/AMX kernel rather than this file's portable loops.  */
/* ------------------------------------------------------------------------- */
#include <dlfcn.h>

typedef void (*sgemm_fn)(int order, int transa, int transb, int64_t m, int64_t n, int64_t k, float alpha, const float *a,
                         int64_t lda, const float *b, int64_t ldb, float beta, float *c, int64_t ldc);
typedef void (*setthr_fn)(int);
static sgemm_fn g_sgemm = NULL;

int orc_blas_load(const char *path) {
    if (g_sgemm) return 0;
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    g_sgemm = (sgemm_fn)dlsym(h, "scipy_cblas_sgemm64_");
    setthr_fn st = (setthr_fn)dlsym(h, "scipy_openblas_set_num_threads64_");
    if (st) st(1); /* threads come from the parts, not from BLAS (omp_set_num_threads(1)) */
    return g_sgemm ? 0 : -2;
}

static void *orc_part_worker_blas(void *arg) {
    orc_part_job *jb = (orc_part_job *)arg;
    const int k = jb->k, d = jb->d;
    const int64_t nx = jb->nx;
    const int64_t YB = 1024, QB = 4096; /* faiss distance_compute_blas_{database,query}_bs */
    float *keys = (float *)malloc(sizeof(float) * (size_t)k * (size_t)nx);
    int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)k * (size_t)nx);
    int *cnt = (int *)calloc((size_t)nx, sizeof(int));
    const int64_t qb_max = nx < QB ? nx : QB;
    float *blk = (float *)malloc(sizeof(float) * (size_t)qb_max * YB);
    float *xn = NULL, *yn = (float *)malloc(sizeof(float) * YB);
    if (jb->metric == ORC_L2) {
        xn = (float *)malloc(sizeof(float) * (size_t)nx);
        for (int64_t q = 0; q < nx; q++) xn[q] = orc_ip(jb->x + q * d, jb->x + q * d, d);
    }
    for (int64_t q0 = 0; q0 < nx; q0 += QB) {
        const int64_t nqb = nx - q0 < QB ? nx - q0 : QB;
        for (int64_t y0 = jb->y0; y0 < jb->y1; y0 += YB) {
            const int64_t nyb = jb->y1 - y0 < YB ? jb->y1 - y0 : YB;
            const float *yb = jb->y + y0 * d;
            /* blk[nqb][nyb] = X[nqb][d] * Y[nyb][d]^T  (row major = 101, NoTrans = 111, Trans = 112) */
            g_sgemm(101, 111, 112, nqb, nyb, d, 1.0f, jb->x + q0 * d, d, yb, d, 0.0f, blk, nyb);
            if (jb->metric == ORC_L2)
                for (int64_t i = 0; i < nyb; i++) yn[i] = orc_ip(yb + (size_t)i * d, yb + (size_t)i * d, d);
            for (int64_t a = 0; a < nqb; a++) {
                const int64_t q = q0 + a;
                orc_topk t = {k, cnt[q], keys + q * k, ids + q * k};
                const float *row = blk + (size_t)a * nyb;
                for (int64_t i = 0; i < nyb; i++) {
                    float key;
                    if (jb->metric == ORC_L2) {
                        key = xn[q] + yn[i] - 2 * row[i];
                        if (key < 0) key = 0;
                    } else {
                        key = -row[i];
                    }
                    if (t.n == k && !(key <= t.key[k - 1])) continue;
                    orc_topk_push(&t, key, y0 + i);
                }
                cnt[q] = t.n;
            }
        }
    }
    for (int64_t q = 0; q < nx; q++)
        for (int j = 0; j < k; j++) {
            if (j < cnt[q]) {
                jb->dis[q * k + j] = jb->metric == ORC_L2 ? keys[q * k + j] : -keys[q * k + j];
                jb->ids[q * k + j] = ids[q * k + j];
            } else {
                jb->dis[q * k + j] = jb->metric == ORC_L2 ? FLT_MAX : -FLT_MAX;
                jb->ids[q * k + j] = -1;
            }
        }
    free(keys);
    free(ids);
    free(cnt);
    free(blk);
    free(xn);
    free(yn);
    return NULL;
}

/* same contract as orc_knn_flat_parts; returns -1 if orc_blas_load() has not succeeded */
int orc_knn_flat_parts_blas(int metric, const float *x, int64_t nx, const float *y, int64_t ny, int d, int k, int n_parts,
                            float *dis, int64_t *ids) {
    if (!g_sgemm) return -1;
    if (n_parts < 1) n_parts = 1;
    if (n_parts > ny && ny > 0) n_parts = (int)ny;
    orc_part_job *jobs = (orc_part_job *)calloc((size_t)n_parts, sizeof(orc_part_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_parts, sizeof(pthread_t));
    int64_t per = (ny + n_parts - 1) / n_parts;
    for (int p = 0; p < n_parts; p++) {
        jobs[p].metric = metric;
        jobs[p].x = x;
        jobs[p].nx = nx;
        jobs[p].y = y;
        jobs[p].y0 = per * p < ny ? per * p : ny;
        jobs[p].y1 = per * (p + 1) < ny ? per * (p + 1) : ny;
        jobs[p].d = d;
        jobs[p].k = k;
        jobs[p].dis = (float *)malloc(sizeof(float) * (size_t)k * (size_t)nx);
        jobs[p].ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)k * (size_t)nx);
        pthread_create(&th[p], NULL, orc_part_worker_blas, &jobs[p]);
    }
    for (int p = 0; p < n_parts; p++) pthread_join(th[p], NULL);
    float *key = (float *)malloc(sizeof(float) * (size_t)k);
    int64_t *id = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    for (int64_t q = 0; q < nx; q++) {
        orc_topk t = {k, 0, key, id};
        for (int p = 0; p < n_parts; p++)
            for (int j = 0; j < k; j++) {
                int64_t rid = jobs[p].ids[q * k + j];
                if (rid < 0) continue;
                float s = jobs[p].dis[q * k + j];
                orc_topk_push(&t, metric == ORC_L2 ? s : -s, rid);
            }
        for (int j = 0; j < k; j++) {
            if (j < t.n) {
                dis[q * k + j] = metric == ORC_L2 ? key[j] : -key[j];
                ids[q * k + j] = id[j];
            } else {
                dis[q * k + j] = metric == ORC_L2 ? FLT_MAX : -FLT_MAX;
                ids[q * k + j] = -1;
            }
        }
    }
    for (int p = 0; p < n_parts; p++) {
        free(jobs[p].dis);
        free(jobs[p].ids);
    }
    free(key);
    free(id);
    free(jobs);
    free(th);
    return 0;
}


/* ------------------------------------------------------------------------- */
/* Small-batch CPU arm (BASELINE configs[0]: FLAT L2 distance(), 10k x 128, one query): faiss' nx < 20 path computes     */
/* exact squared differences row by row with SIMD (fvec_L2sqr_ny) and keeps a heap -- single-threaded inside a part        */
/* (omp_set_num_threads(1), VIWithDataPart.h:350).  16-lane accumulators; target_clones picks AVX-512 / AVX2 at run time.  */
/* Results are verified against vs_oracle.c in tests/test_oracle_golden.py.                                               */
/* ------------------------------------------------------------------------- */
__attribute__((target_clones("avx512f", "avx2,fma", "default"))) static float orc_l2_simd(const float *x, const float *y, int d) {
    float acc[16] = {0};
    int j = 0;
    for (; j + 16 <= d; j += 16) {
#pragma GCC unroll 16
        for (int l = 0; l < 16; l++) {
            const float t = x[j + l] - y[j + l];
            acc[l] += t * t;
        }
    }
    float s = 0;
    for (int l = 0; l < 16; l++) s += acc[l];
    for (; j < d; j++) {
        const float t = x[j] - y[j];
        s += t * t;
    }
    return s;
}
__attribute__((target_clones("avx512f", "avx2,fma", "default"))) static float orc_ip_simd(const float *x, const float *y, int d) {
    float acc[16] = {0};
    int j = 0;
    for (; j + 16 <= d; j += 16) {
#pragma GCC unroll 16
        for (int l = 0; l < 16; l++) acc[l] += x[j + l] * y[j + l];
    }
    float s = 0;
    for (int l = 0; l < 16; l++) s += acc[l];
    for (; j < d; j++) s += x[j] * y[j];
    return s;
}

void orc_knn_flat_simd(int metric, const float *x, int64_t nx, const float *y, int64_t ny, int d, int k, float *dis, int64_t *ids) {
    for (int64_t q = 0; q < nx; q++) {
        orc_topk t = {k, 0, dis + q * k, ids + q * k};
        for (int64_t i = 0; i < ny; i++) {
            const float key = metric == ORC_L2 ? orc_l2_simd(x + q * d, y + i * d, d) : -orc_ip_simd(x + q * d, y + i * d, d);
            if (t.n == k && !(key < t.key[k - 1])) continue;
            orc_topk_push(&t, key, i);
        }
        for (int j = 0; j < k; j++) {
            if (j < t.n) {
                if (metric != ORC_L2) dis[q * k + j] = -dis[q * k + j];
            } else {
                dis[q * k + j] = metric == ORC_L2 ? FLT_MAX : -FLT_MAX;
                ids[q * k + j] = -1;
            }
        }
    }
}
