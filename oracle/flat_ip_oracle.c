/*
 * flat_ip_oracle.c — CPU restatement of the exact inner-product top-k search.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kirag_b200/ imports, links or
 * executes this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg do, and only as the checker / the timed
 * CPU baseline — never as the product.
 *
 * PARITY UNPINNED for the search: the reference's implementation of this path
 * is the third-party wheel faiss-cpu==1.8.0.post1 (/root/reference/requirements.txt:10,
 * call sites /root/reference/retriever/index.py:13,32,47).  It is not vendored
 * under /root/reference, is not installed in this image, cannot be fetched (no
 * network), and the reference ships no tests, golden vectors or fixtures for
 * it.  What follows restates FAISS 1.8.0's published
 * algorithm (IndexFlat::search -> knn_inner_product):
 *
 *   - a score is the plain fp32 inner product <q_i, x_j>;
 *   - for each query the corpus is scanned in ascending j and a size-k
 *     min-heap is maintained; a row enters only if it beats the heap top
 *     (HeapBlockResultHandler::add_results, `if (C::cmp(thresh, dis))`);
 *   - n >= 20 queries go through blocked sgemm (4096 queries x 1024 rows per
 *     block, exhaustive_inner_product_blas) followed by the same heap update
 *     (oracle_heap_add_block below is that update);
 *   - output rows are sorted by descending score; unfilled slots hold
 *     (-FLT_MAX, -1) (heap_heapify's neutral element for CMin<float,int64>).
 *
 * Tie rule adopted by this build (BASELINE.json north_star: "ties broken by
 * lower id"): the total order is (score descending, id ascending).  Because
 * rows are scanned in ascending id and admission is strict, this keeps the
 * same SET of rows as FAISS whenever the k-th and (k+1)-th scores differ; it
 * can differ from FAISS only in which of several bit-identical scores
 * straddling rank k is kept and in the order among bit-identical scores.
 *
 * Pinned parts: the call-site glue (index.py:26-53) is exercised by running
 * the reference's own Indexer unmodified on top of this oracle
 * (tests/test_reference_indexer.py), and the aligner-scoring call site
 * (knowledge_graph/models.py:1532-1538, torch.matmul + torch.topk) is pinned
 * by golden vectors generated with those exact torch calls
 * (tests/golden/make_golden.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* a "beats" b in the (score desc, id asc) order */
static inline int beats(float sa, int64_t ia, float sb, int64_t ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

/* min-heap on the order above: root = the worst kept element */
static void heap_sift_down(float* hs, int64_t* hi, int k, int pos) {
    for (;;) {
        int l = 2 * pos + 1, r = l + 1, w = pos;
        if (l < k && beats(hs[w], hi[w], hs[l], hi[l])) w = l;
        if (r < k && beats(hs[w], hi[w], hs[r], hi[r])) w = r;
        if (w == pos) return;
        float ts = hs[w]; hs[w] = hs[pos]; hs[pos] = ts;
        int64_t ti = hi[w]; hi[w] = hi[pos]; hi[pos] = ti;
        pos = w;
    }
}

/* empty slots are (-FLT_MAX, INT64_MAX): they lose against every real row whose
 * score is >= -FLT_MAX, exactly FAISS's neutral element for CMin<float,int64>
 * (a row scoring below -FLT_MAX, i.e. -inf, is never admitted by FAISS either);
 * they are rewritten to id -1 on output */
static void heap_init(float* hs, int64_t* hi, int k) {
    for (int i = 0; i < k; ++i) { hs[i] = -FLT_MAX; hi[i] = INT64_MAX; }
}

static void heap_push(float* hs, int64_t* hi, int k, float s, int64_t id) {
    if (!(s == s)) return;                         /* NaN: `thresh < dis` is false */
    if (!beats(s, id, hs[0], hi[0])) return;       /* strict admission against the heap top */
    hs[0] = s; hi[0] = id;
    heap_sift_down(hs, hi, k, 0);
}

typedef struct { float s; int64_t id; } pair_t;

static int pair_cmp(const void* a, const void* b) {
    const pair_t* pa = (const pair_t*)a; const pair_t* pb = (const pair_t*)b;
    const int ea = (pa->id < 0 || pa->id == INT64_MAX), eb = (pb->id < 0 || pb->id == INT64_MAX);
    if (ea && eb) return 0;
    if (ea) return 1;
    if (eb) return -1;
    if (beats(pa->s, pa->id, pb->s, pb->id)) return -1;
    if (beats(pb->s, pb->id, pa->s, pa->id)) return 1;
    return 0;
}

static void heap_to_sorted(float* hs, int64_t* hi, int k) {
    pair_t* tmp = (pair_t*)malloc(sizeof(pair_t) * (size_t)k);
    for (int i = 0; i < k; ++i) { tmp[i].s = hs[i]; tmp[i].id = hi[i]; }
    qsort(tmp, (size_t)k, sizeof(pair_t), pair_cmp);
    for (int i = 0; i < k; ++i) {
        if (tmp[i].id < 0 || tmp[i].id == INT64_MAX) { hs[i] = -FLT_MAX; hi[i] = -1; }
        else { hs[i] = tmp[i].s; hi[i] = tmp[i].id; }
    }
    free(tmp);
}

/* fvec_inner_product restated: fp32 accumulate in element order */
static float dot_f32(const float* a, const float* b, int d) {
    float acc = 0.f;
    for (int i = 0; i < d; ++i) acc += a[i] * b[i];
    return acc;
}
/* same contraction with an fp64 accumulator, rounded once (error-budget reference) */
static float dot_f64(const float* a, const float* b, int d) {
    double acc = 0.0;
    for (int i = 0; i < d; ++i) acc += (double)a[i] * (double)b[i];
    return (float)acc;
}

/*
 * IndexFlatIP.search — the n < 20 code path (per-query scan) applied to any n.
 * accum: 0 = fp32 accumulation, 1 = fp64 accumulation rounded to fp32.
 * D [nq,k], I [nq,k] as FAISS returns them.  id_offset is added to the ids.
 */
int oracle_flat_ip_search(const float* xb, int64_t n, int d, const float* xq, int64_t nq, int k,
                          float* D, int64_t* I, int accum, int64_t id_offset) {
    if (k <= 0 || d <= 0 || n < 0 || nq < 0) return 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
        float* hs = D + qi * k;
        int64_t* hi = I + qi * k;
        heap_init(hs, hi, k);
        const float* q = xq + qi * (int64_t)d;
        for (int64_t j = 0; j < n; ++j) {
            const float s = accum ? dot_f64(q, xb + j * (int64_t)d, d) : dot_f32(q, xb + j * (int64_t)d, d);
            heap_push(hs, hi, k, s, j);
        }
        heap_to_sorted(hs, hi, k);
        if (id_offset)
            for (int i = 0; i < k; ++i) if (hi[i] >= 0) hi[i] += id_offset;
    }
    return 0;
}

/*
 * The heap update of the blocked path (exhaustive_inner_product_blas):
 * scores is a row-major [nq, nb] block of inner products of the queries with
 * rows j0 .. j0+nb-1 (computed by the caller's sgemm); heaps are D/I [nq,k],
 * initialised with oracle_heap_init and finished with oracle_heap_finish.
 */
void oracle_heap_init(float* D, int64_t* I, int64_t nq, int k) {
    for (int64_t qi = 0; qi < nq; ++qi) heap_init(D + qi * k, I + qi * k, k);
}
void oracle_heap_add_block(const float* scores, int64_t nq, int64_t nb, int64_t ld, int64_t j0,
                           float* D, int64_t* I, int k) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        float* hs = D + qi * k;
        int64_t* hi = I + qi * k;
        const float* row = scores + qi * ld;
        for (int64_t j = 0; j < nb; ++j) {
            const float s = row[j];
            if (s < hs[0]) continue; /* cheap reject against the heap top */
            heap_push(hs, hi, k, s, j0 + j);
        }
    }
}
void oracle_heap_finish(float* D, int64_t* I, int64_t nq, int k) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) heap_to_sorted(D + qi * k, I + qi * k, k);
}

/* merge of G per-shard results [G,nq,k] -> [nq,k] in the same total order */
int oracle_merge_topk(const float* D_all, const int64_t* I_all, int G, int64_t nq, int k,
                      float* D_out, int64_t* I_out) {
    if (G <= 0 || k <= 0) return 1;
    for (int64_t qi = 0; qi < nq; ++qi) {
        pair_t* tmp = (pair_t*)malloc(sizeof(pair_t) * (size_t)G * (size_t)k);
        for (int g = 0; g < G; ++g)
            for (int j = 0; j < k; ++j) {
                const int64_t src = ((int64_t)g * nq + qi) * k + j;
                tmp[g * k + j].s = D_all[src];
                tmp[g * k + j].id = I_all[src];
                if (!(D_all[src] == D_all[src])) tmp[g * k + j].id = -1;
            }
        qsort(tmp, (size_t)G * (size_t)k, sizeof(pair_t), pair_cmp);
        for (int j = 0; j < k; ++j) {
            if (tmp[j].id < 0) { D_out[qi * k + j] = -FLT_MAX; I_out[qi * k + j] = -1; }
            else { D_out[qi * k + j] = tmp[j].s; I_out[qi * k + j] = tmp[j].id; }
        }
        free(tmp);
    }
    return 0;
}

/*
 * average_pool + F.normalize restated for fp32 hidden states and int64 masks
 * (/root/reference/retriever/encoders.py:56-58,76; mode 1 = BGE CLS tail :115-117).
 * Sums in fp64 and rounds once: this is the error-budget reference; the
 * bit-level reference is torch itself (tests/golden/make_golden.py).
 */
int oracle_pool_normalize(const float* hidden, const int64_t* mask, float* out, int64_t B, int64_t S,
                          int64_t H, int mode, int normalize) {
    for (int64_t b = 0; b < B; ++b) {
        double denom = 1.0;
        if (mode == 0) {
            long long ms = 0;
            for (int64_t s = 0; s < S; ++s) ms += mask[b * S + s];
            denom = (double)ms;
        }
        double ss = 0.0;
        for (int64_t h = 0; h < H; ++h) {
            double acc = 0.0;
            if (mode == 0) {
                for (int64_t s = 0; s < S; ++s)
                    if (mask[b * S + s] != 0) acc += (double)hidden[(b * S + s) * H + h];
            } else {
                acc = (double)hidden[(b * S) * H + h];
            }
            const float pooled = (float)(acc / denom);
            out[b * H + h] = pooled;
            ss += (double)pooled * (double)pooled;
        }
        if (normalize) {
            double nrm = sqrt(ss);
            if (!(nrm > 1e-12)) nrm = (nrm == nrm) ? 1e-12 : nrm;
            for (int64_t h = 0; h < H; ++h) out[b * H + h] = (float)((double)out[b * H + h] / nrm);
        }
    }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
