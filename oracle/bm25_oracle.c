/* TEST INFRASTRUCTURE - plain-C restatement of the reference's BM25 hot path (CPU).
 *
 * This file is the checker and the CPU baseline timed by bench.py; it is never linked into or
 * called from the product library.  Parity pinning: validated against the numpy restatement
 * (oracle/bm25_oracle.py) and, through the committed fixtures in tests/golden/ (made by
 * tests/golden/make_golden.py from the UNMODIFIED reference functions), against the reference
 * itself - see tests/test_oracle.py.
 *
 * What it restates (all arithmetic float64, the reference's own operation order):
 *   orc_build      BM25.__init__/build          bm25_ranking.ipynb:167-189
 *                  (tf per doc, df, postings appended in doc order, avgdl = sum(len)/N, idf)
 *   orc_get_scores BM25.get_scores              bm25_ranking.ipynb:191-204   (dedup = 1)
 *                  scoring loop of score_documents_for_query  team_run1.py:183-194 (dedup = 0)
 *   orc_topk_*     BM25.retrieve_top_n          bm25_ranking.ipynb:206-213, ties canonicalised
 *                  to (score desc, doc id asc); positive_only = heapq.nlargest over touched docs
 *                  team_run1.py:196
 * Variants: 0 "notebook" (norm = 1-b+dl/avgdl), 1 "okapi" (norm = 1-b+b*dl/avgdl, same idf),
 *           2 "okapi_no_plus1" (idf without +1, cosine_similarity_bm25_reranking.py:179).
 *
 * Threading mirrors the reference's only parallel scheme: query-parallel workers sharing one
 * read-only index (process_map(score_documents_for_query, ...), team_run1.py:102,202).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC; no -ffast-math, no FMA
 * contraction: -ffp-contract=off keeps a*b+c as two roundings like CPython).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct orc_index {
    int64_t n_docs;
    int32_t vocab;
    int64_t nnz;
    int64_t *row_ptr; /* [V+1] */
    int32_t *doc;     /* [nnz] ascending inside a term */
    int32_t *tf;      /* [nnz] */
    int32_t *dl;      /* [N] */
    int64_t *df;      /* [V] */
    double *idf;      /* [V] (NaN where df == 0) */
    double avgdl;
    double n_stat;    /* N used in idf (global N of a sharded corpus) */
    double k1, b;
    int variant;
} orc_index;

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

static double idf_of(double n, double df, int variant) {
    double x = (n - df + 0.5) / (df + 0.5);
    if (variant == 2) return log(x);
    return log(1 + x); /* ln(1+x) == ln(x+1): float add commutes */
}

void orc_free(orc_index *ix) {
    if (!ix) return;
    free(ix->row_ptr); free(ix->doc); free(ix->tf); free(ix->dl); free(ix->df); free(ix->idf);
    free(ix);
}

/* Recompute idf from (possibly overridden) statistics. */
void orc_set_stats(orc_index *ix, double n_stat, double avgdl, const int64_t *df_override) {
    ix->n_stat = n_stat;
    ix->avgdl = avgdl;
    for (int32_t t = 0; t < ix->vocab; ++t) {
        int64_t d = df_override ? df_override[t] : ix->df[t];
        ix->idf[t] = d > 0 ? idf_of(n_stat, (double)d, ix->variant) : NAN;
    }
}

orc_index *orc_build(const int32_t *token_ids, const int64_t *doc_offsets, int64_t n_docs,
                     int32_t vocab, double k1, double b, int variant, int n_threads) {
    if (n_docs <= 0 || vocab <= 0) return NULL;
    int64_t total = doc_offsets[n_docs];
    for (int64_t i = 0; i < total; ++i)
        if (token_ids[i] < 0 || token_ids[i] >= vocab) return NULL;
    orc_index *ix = (orc_index *)calloc(1, sizeof(orc_index));
    ix->n_docs = n_docs; ix->vocab = vocab; ix->k1 = k1; ix->b = b; ix->variant = variant;
    ix->dl = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_docs);
    ix->df = (int64_t *)calloc((size_t)vocab, sizeof(int64_t));
    ix->idf = (double *)malloc(sizeof(double) * (size_t)vocab);
    ix->row_ptr = (int64_t *)calloc((size_t)vocab + 1, sizeof(int64_t));
    /* pass 1: per-doc sorted copy -> distinct (term, tf); df counts */
    int32_t *sorted = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
    memcpy(sorted, token_ids, sizeof(int32_t) * (size_t)total);
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(n_threads)
    for (int64_t d = 0; d < n_docs; ++d) {
        int64_t lo = doc_offsets[d], hi = doc_offsets[d + 1];
        ix->dl[d] = (int32_t)(hi - lo);
        qsort(sorted + lo, (size_t)(hi - lo), sizeof(int32_t), cmp_i32);
    }
    for (int64_t d = 0; d < n_docs; ++d) {
        int64_t lo = doc_offsets[d], hi = doc_offsets[d + 1];
        for (int64_t i = lo; i < hi; ++i)
            if (i == lo || sorted[i] != sorted[i - 1]) ix->df[sorted[i]]++;
    }
    for (int32_t t = 0; t < vocab; ++t) ix->row_ptr[t + 1] = ix->row_ptr[t] + ix->df[t];
    ix->nnz = ix->row_ptr[vocab];
    ix->doc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ix->nnz > 0 ? ix->nnz : 1));
    ix->tf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ix->nnz > 0 ? ix->nnz : 1));
    /* pass 2: append postings in doc order (inverted_index[word].append(doc_id), :186) */
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)vocab);
    memcpy(cursor, ix->row_ptr, sizeof(int64_t) * (size_t)vocab);
    for (int64_t d = 0; d < n_docs; ++d) {
        int64_t lo = doc_offsets[d], hi = doc_offsets[d + 1];
        int64_t i = lo;
        while (i < hi) {
            int64_t j = i + 1;
            while (j < hi && sorted[j] == sorted[i]) ++j;
            int64_t p = cursor[sorted[i]]++;
            ix->doc[p] = (int32_t)d;
            ix->tf[p] = (int32_t)(j - i);
            i = j;
        }
    }
    free(cursor); free(sorted);
    orc_set_stats(ix, (double)n_docs, (double)total / (double)n_docs, NULL);
    return ix;
}

int64_t orc_nnz(const orc_index *ix) { return ix->nnz; }
double orc_avgdl(const orc_index *ix) { return ix->avgdl; }

void orc_export(const orc_index *ix, int64_t *row_ptr, int32_t *doc, int32_t *tf, int32_t *dl,
                int64_t *df, double *idf) {
    if (row_ptr) memcpy(row_ptr, ix->row_ptr, sizeof(int64_t) * ((size_t)ix->vocab + 1));
    if (doc) memcpy(doc, ix->doc, sizeof(int32_t) * (size_t)ix->nnz);
    if (tf) memcpy(tf, ix->tf, sizeof(int32_t) * (size_t)ix->nnz);
    if (dl) memcpy(dl, ix->dl, sizeof(int32_t) * (size_t)ix->n_docs);
    if (df) memcpy(df, ix->df, sizeof(int64_t) * (size_t)ix->vocab);
    if (idf) memcpy(idf, ix->idf, sizeof(double) * (size_t)ix->vocab);
}

static double norm_of(const orc_index *ix, double dl) {
    if (ix->variant == 0) return 1 - ix->b + dl / ix->avgdl;      /* bm25_ranking.ipynb:202 */
    return 1 - ix->b + ix->b * dl / ix->avgdl;                    /* team_run1.py:193 */
}

/* scores[N] must be zeroed by the caller.  terms_tmp: scratch of n_terms int32. */
static void accumulate(const orc_index *ix, const int32_t *q_terms, int32_t n_terms, int dedup,
                       double *scores, int32_t *terms_tmp) {
    int32_t m = 0;
    for (int32_t i = 0; i < n_terms; ++i) {
        int32_t t = q_terms[i];
        if (t < 0 || t >= ix->vocab || ix->row_ptr[t + 1] == ix->row_ptr[t]) continue; /* OOV :195 */
        if (isnan(ix->idf[t])) continue;
        terms_tmp[m++] = t;
    }
    if (dedup) { /* set(query), canonical ascending-term order */
        qsort(terms_tmp, (size_t)m, sizeof(int32_t), cmp_i32);
        int32_t w = 0;
        for (int32_t i = 0; i < m; ++i)
            if (i == 0 || terms_tmp[i] != terms_tmp[i - 1]) terms_tmp[w++] = terms_tmp[i];
        m = w;
    }
    const double k1 = ix->k1;
    for (int32_t i = 0; i < m; ++i) {
        int32_t t = terms_tmp[i];
        double idf = ix->idf[t];
        for (int64_t p = ix->row_ptr[t]; p < ix->row_ptr[t + 1]; ++p) {
            int32_t d = ix->doc[p];
            double tf = (double)ix->tf[p];
            double dl = (double)ix->dl[d];
            double score = idf * ((tf * (k1 + 1)) / (tf + k1 * norm_of(ix, dl)));
            scores[d] += score;
        }
    }
}

void orc_get_scores(const orc_index *ix, const int32_t *q_terms, int32_t n_terms, int dedup,
                    double *scores) {
    memset(scores, 0, sizeof(double) * (size_t)ix->n_docs);
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_terms > 0 ? n_terms : 1));
    accumulate(ix, q_terms, n_terms, dedup, scores, tmp);
    free(tmp);
}

/* (score desc, id asc) "better than" */
static int better(double sa, int32_t ia, double sb, int32_t ib) {
    return sa > sb || (sa == sb && ia < ib);
}

/* top-k of scores[N] by a size-k binary heap whose root is the worst kept element. */
static int32_t topk_heap(const double *scores, int64_t n, int32_t k, int positive_only,
                         int32_t *out_ids, double *out_scores) {
    int32_t cnt = 0;
    for (int64_t d = 0; d < n; ++d) {
        double s = scores[d];
        if (positive_only && s == 0.0) continue;
        if (cnt < k) {
            int32_t i = cnt++;
            out_ids[i] = (int32_t)d; out_scores[i] = s;
            while (i > 0) { /* sift up: parent must be worse-or-equal (root = worst) */
                int32_t p = (i - 1) / 2;
                if (better(out_scores[p], out_ids[p], out_scores[i], out_ids[i])) {
                    double ts = out_scores[p]; out_scores[p] = out_scores[i]; out_scores[i] = ts;
                    int32_t ti = out_ids[p]; out_ids[p] = out_ids[i]; out_ids[i] = ti;
                    i = p;
                } else break;
            }
        } else if (better(s, (int32_t)d, out_scores[0], out_ids[0])) {
            out_scores[0] = s; out_ids[0] = (int32_t)d;
            int32_t i = 0;
            for (;;) { /* sift down: move the worst child up */
                int32_t l = 2 * i + 1, r = l + 1, w = i;
                if (l < cnt && better(out_scores[w], out_ids[w], out_scores[l], out_ids[l])) w = l;
                if (r < cnt && better(out_scores[w], out_ids[w], out_scores[r], out_ids[r])) w = r;
                if (w == i) break;
                double ts = out_scores[w]; out_scores[w] = out_scores[i]; out_scores[i] = ts;
                int32_t ti = out_ids[w]; out_ids[w] = out_ids[i]; out_ids[i] = ti;
                i = w;
            }
        }
    }
    /* insertion sort into canonical order */
    for (int32_t i = 1; i < cnt; ++i) {
        double s = out_scores[i]; int32_t id = out_ids[i]; int32_t j = i - 1;
        while (j >= 0 && better(s, id, out_scores[j], out_ids[j])) {
            out_scores[j + 1] = out_scores[j]; out_ids[j + 1] = out_ids[j]; --j;
        }
        out_scores[j + 1] = s; out_ids[j + 1] = id;
    }
    return cnt;
}

/* out_ids/out_scores are [nq, k] (unused tail: id -1, score 0); out_counts[nq] optional. */
void orc_topk_batch(const orc_index *ix, const int32_t *q_terms, const int32_t *q_offsets,
                    int32_t nq, int32_t k, int dedup, int positive_only, int n_threads,
                    int32_t *out_ids, double *out_scores, int32_t *out_counts) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        double *scores = (double *)malloc(sizeof(double) * (size_t)ix->n_docs);
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * 4096);
        int32_t tmp_cap = 4096;
#pragma omp for schedule(dynamic, 1)
        for (int32_t q = 0; q < nq; ++q) {
            int32_t n_terms = q_offsets[q + 1] - q_offsets[q];
            if (n_terms > tmp_cap) { tmp_cap = n_terms; tmp = (int32_t *)realloc(tmp, sizeof(int32_t) * (size_t)tmp_cap); }
            memset(scores, 0, sizeof(double) * (size_t)ix->n_docs);
            accumulate(ix, q_terms + q_offsets[q], n_terms, dedup, scores, tmp);
            int32_t *ids = out_ids + (int64_t)q * k;
            double *sc = out_scores + (int64_t)q * k;
            int32_t kk = k < ix->n_docs ? k : (int32_t)ix->n_docs;
            int32_t c = topk_heap(scores, ix->n_docs, kk, positive_only, ids, sc);
            for (int32_t i = c; i < k; ++i) { ids[i] = -1; sc[i] = 0.0; }
            if (out_counts) out_counts[q] = c;
        }
        free(scores); free(tmp);
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
