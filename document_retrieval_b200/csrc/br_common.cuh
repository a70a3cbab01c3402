// br_common.cuh - shared declarations of the br_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/br_b200.h"

// docs per skip-table sub-range = docs per warp of the tile kernel (2^kSubShift)
#ifndef BR_TILE_SHIFT
#define BR_TILE_SHIFT 11
#endif

namespace br {

constexpr int kSubShift = BR_TILE_SHIFT;

void set_error(const std::string& msg);

#define BR_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            br::set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " (" +       \
                          __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
            return BR_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define BR_TRY(call)                  \
    do {                              \
        int _s = (call);              \
        if (_s != BR_OK) return _s;   \
    } while (0)

#define BR_REQUIRE(cond, status, msg) \
    do {                              \
        if (!(cond)) {                \
            br::set_error(msg);       \
            return status;            \
        }                             \
    } while (0)

// Grow-only device scratch buffer (never shrinks; freed with the handle).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return BR_OK;
        if (p) BR_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        BR_CUDA(cudaMalloc(&p, want));
        cap = want;
        return BR_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// Stream-ordered scratch comes from a PRIVATE memory pool of the library, one per device: with a release threshold of 0
// every stream synchronisation would hand freed memory back to the driver and the next call pay for mapping it again
// (tens of ms for the larger buffers), so the pool keeps up to kPoolKeepBytes; anything above that is released at the
// next synchronisation.  The process's default pool (and with it the host application's own stream-ordered allocations)
// is left alone; br_trim_scratch() returns the retained memory on request.
constexpr uint64_t kPoolKeepBytes = 8ull << 30;
inline cudaMemPool_t& scratch_pool_slot(int dev) {
    static cudaMemPool_t pools[64] = {};
    return pools[dev & 63];
}
inline int scratch_pool(cudaMemPool_t* out) {
    int dev = 0;
    BR_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t& pool = scratch_pool_slot(dev);
    if (!pool) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        BR_CUDA(cudaMemPoolCreate(&pool, &props));
        uint64_t keep = kPoolKeepBytes;
        BR_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = pool;
    return BR_OK;
}
inline int scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    cudaMemPool_t pool;
    BR_TRY(scratch_pool(&pool));
    BR_CUDA(cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, st));
    return BR_OK;
}

constexpr int kNumSMs = 148;          // B200
constexpr float kBandRel = 1e-5f;     // fp32 candidate band kept below the k-th fp32 score

}  // namespace br

// One posting: doc id + precomputed fp32 BM25 weight, 8 bytes, read as one 64-bit word
// (two postings per 128-bit load).
struct __align__(8) br_posting {
    uint32_t doc;
    float w;
};

struct br_index {
    int device = 0;
    int64_t n_docs = 0, doc_base = 0, nnz = 0, sum_dl = 0;
    int32_t vocab = 0;
    bool finalized = false;
    int variant = 0;
    double k1 = 1.5, b = 0.75, n_stat = 0, sum_dl_stat = 0, avgdl = 0;

    // CSR by term (device)
    int64_t* row_ptr = nullptr;     // [V+1]
    br_posting* post = nullptr;     // [nnz]   doc ids ascending inside a term
    br_posting* post_cos = nullptr; // [nnz]   optional TF-IDF cosine weights tf*idf^2/||d|| (br_index_enable_tfidf)
    double* cos_inv_norm = nullptr; // [N]     1/||d|| of the doc's float32 tf*idf vector, float64 (with post_cos)
    uint16_t* tf = nullptr;         // [nnz]   raw term frequency (for the float64 re-score)
    uint32_t* dl = nullptr;         // [N]
    uint32_t* df = nullptr;         // [V]     shard-local document frequency
    double* idf = nullptr;          // [V]     statistics in force (NaN where df_stat == 0)
    float* ub = nullptr;            // [V]     largest posting weight of the term on this shard (0 when none / negative):
                                    //         the per-term score upper bound of the tiled scorer's MaxScore deferral

    // hot-term skip tables for the tiled scorer: term t is "hot" when hot_slot[t] >= 0; then
    // skip[hot_slot[t] * (n_sub+1) + j] = offset (relative to row_ptr[t]) of the first posting
    // with doc >= j << sub_shift.
    int sub_shift = 0;
    int32_t n_sub = 0, n_hot = 0;
    uint32_t hot_min = 0;           // smallest df with a skip table
    int32_t* hot_slot = nullptr;    // [V]
    uint32_t* skip = nullptr;       // [n_hot, n_sub+1]
    // dense rows: w[row][doc] as plain fp32 rows (0 where the doc lacks the term), most frequent term first.
    // Rows [0, n_srows) - terms present in >= ~20% of the docs - are STREAMED by the tiled scorer (accumulators are
    // initialised from them instead of walking their postings); all n_rows rows (terms with df >= N/64) serve as
    // LOOK-UP tables: a term the scorer has deferred (MaxScore) is added to a candidate doc with one 4-byte load.
    float* dense_rows = nullptr;    // [n_rows, n_pad]
    int16_t* row_slot = nullptr;    // [V] row index or -1
    int32_t n_rows = 0, n_srows = 0;
    int64_t n_pad = 0;
    int8_t* sig_bit = nullptr;      // [V] rank (0 = largest df) among the 32 most frequent terms, -1 otherwise

    // host mirrors
    std::vector<uint32_t> h_df;     // shard-local
    std::vector<int64_t> h_df_stat; // statistics in force
    std::vector<double> h_idf;

    // query-time scratch
    br::DevBuf ws_prep, ws_dense, ws_sel, ws_cand, ws_misc, ws_tile, ws_sort, ws_cold, ws_rec;
    br_query_stats stats{};

    bool allow_fused = true, allow_fused_bigk = true, allow_fused_long = true;
    int tile_g = 0;
    // doc-sharded callers: callback that max-reduces thr[Q] over the shards (br_set_thr_exchange), user pointer, and the
    // number of tile launches after which it is called (plus once after seeding); rounds < 0: off
    br_thr_exchange_fn thr_exchange = nullptr;
    void* thr_exchange_user = nullptr;
    int thr_exchange_rounds = -1;
    int thr_exchange_world = 1;
    bool seed_thr = true;           // threshold seeding before the first launch of the tiled scorer
    int defer_pm = 650;             // MaxScore deferral budget of the tiled scorer, per mille of the threshold (0 = off)
    int tile_growth = 2;            // every launch covers this many times the tiles of the one before
    int sparse_mode = 1;            // sparse phase of the tile kernel: 0 next non-empty slice pulled over by shuffles, 1 term list walked from shared memory
    int tile_tpb = 16;              // consecutive tiles per CTA in the large launches
    int tile_dense_min = 32;        // average postings of a term per sub-range from which its slices are walked
                                    // term by term (whole warp, pipelined) instead of concatenated with the sparse ones

    // optional event timing of the scoring kernel
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;   // pairs (start, stop)
    size_t ev_used = 0;
    void prof_begin(cudaStream_t st) {
        if (!profiling) return;
        if (ev_used + 2 > ev_pool.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            ev_pool.push_back(a); ev_pool.push_back(b);
        }
        cudaEventRecord(ev_pool[ev_used], st);
    }
    void prof_end(cudaStream_t st) {
        stats.score_launches += 1;
        if (!profiling) return;
        cudaEventRecord(ev_pool[ev_used + 1], st);
        ev_used += 2;
    }
    void prof_collect() {   // call after the stream is synchronised
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]) == cudaSuccess) stats.score_ms += ms;
        }
        ev_used = 0;
    }
};

namespace br {
// br_build.cu
int index_build(const int32_t* token_ids, const int64_t* doc_offsets, int64_t n_docs, int32_t vocab,
                int64_t doc_base, cudaStream_t st, br_index** out);
int index_finalize(br_index* ix, double k1, double b, int variant, double n_stat, double sum_dl_stat,
                   const int64_t* df_stat_host, cudaStream_t st);
int index_import(const int64_t* row_ptr, const int32_t* doc, const int32_t* tf, const int32_t* dl,
                 int64_t n_docs, int32_t vocab, int64_t doc_base, cudaStream_t st, br_index** out);
int index_import_dev(const int64_t* row_ptr, const int32_t* doc, const uint16_t* tf, const int32_t* dl, int64_t n_docs,
                     int32_t vocab, int64_t nnz, int64_t doc_base, cudaStream_t st, br_index** out);
void index_free(br_index* ix);
// br_query.cu
int score_batch(br_index* ix, const int32_t* q_terms, const int32_t* q_offsets, int32_t nq, int dedup,
                float* out_scores, cudaStream_t st);
int topk_batch(br_index* ix, const int32_t* q_terms, const int32_t* q_offsets, int32_t nq, int32_t n_terms, int32_t k,
               int dedup, int positive_only, int32_t* out_ids, double* out_scores, int32_t* out_counts,
               br_record* out_recs, cudaStream_t st);
int rescore_docs(br_index* ix, const int32_t* q_terms, const int32_t* q_offsets, int32_t nq, int dedup,
                 const int32_t* cand_ids, const int64_t* cand_off, double* out_scores, cudaStream_t st);
int enable_tfidf(br_index* ix, cudaStream_t st);
int tfidf_topk(br_index* ix, const int32_t* q_terms, const int32_t* q_offsets, int32_t nq, int32_t k, int32_t* out_ids,
               double* out_scores, int32_t* out_counts, cudaStream_t st);
int rerank_v3(br_index* ix, const int32_t* q_terms, const int32_t* q_offsets, int32_t nq, const int32_t* cand_ids,
              const int64_t* cand_off, double* out_scores, cudaStream_t st);
int topk_merge(const int64_t* ids, const double* scores, int32_t n_parts, int32_t nq, int32_t k,
               int64_t* out_ids, double* out_scores, cudaStream_t st);
int topk_merge_records(const br_record* recs, int32_t n_parts, int32_t nq, int32_t k, int64_t* out_ids,
                       double* out_scores, cudaStream_t st);
}  // namespace br
