// br_tile.cu - the fused BM25 path: doc-range tiled term-at-a-time scoring with the accumulators in shared memory
// and the top-k selection fused into the tile epilogue.
//
// Why: a dense [Q, N] fp32 accumulator costs 8 B x N of HBM traffic per query (zero + read back) and a global atomic
// per posting; at N = 8.8M that is more than the postings themselves (hard part 2).  Here a CTA owns (a tile
// of W x S consecutive docs) x (a group of G queries); each of its W = 8 warps owns one sub-range of S = 2^kSubShift
// docs and walks, for the terms the queries' plans stream, the slice of the term's posting list that falls in its
// sub-range (slice bounds come from the per-term skip table built with the index) and adds the packed fp32 weights
// into its private [G, S] accumulator rows: plain LDS/FFMA/STS, no atomics, no block barrier after the prologue,
// deterministic order.  The accumulators are 64 KB per CTA (3 CTAs per SM) whatever the shape: G x S = 2048 -
// 4 queries x 512 docs in round 1, 1 query x 2048 docs now (measured on the C4 workload: 134k q/s at 4 x 512, 164k at
// 2 x 1024, 191k at 1 x 2048: a (query, sub-range) costs a fixed ~100 warp instructions of bookkeeping whatever it
// holds, and a 2048-doc slice of a rare term fills the 32 lanes of the walk where a 512-doc slice held 2-3 postings).
// The CTAs of all groups visit a tile back to back (blockIdx.x = group), so the tile's posting slices come from HBM
// once per batch and from L2 afterwards.
//
// Terms too rare for a skip table ("cold": df < hot_min) are bucketed per batch instead: a counting sort of their
// postings by (tile, group) (k_cold_pass / scan); the CTA of a bucket adds those few entries with shared-memory atomics
// before its epilogue.
//
// Epilogue + selection: thr[q] is a lower bound of the k-th best fp32 score of query q over the shard (over the whole
// corpus with the cross-shard exchange).  A doc is emitted as a candidate iff score >= thr*(1-1e-5) (the band that
// makes the float64 re-score exact, see br_query.cu).  thr rises two ways: inside the kernel, when a sub-range holds k
// better docs in k different lanes (atomicMax); and between launches - tiles are processed in chunks of doubling size,
// and after each chunk k_tighten sets thr[q] to the k-th best score among the candidates emitted so far and compacts
// the list.  With doubling chunks every chunk emits about k new candidates per query, so the candidate list stays ~k
// long and the float64 re-score costs almost nothing.
#include <math_constants.h>

#include <algorithm>

#include <cub/cub.cuh>

#include "br_common.cuh"
#include "br_kernels.cuh"
#include "br_query.cuh"

namespace br {

constexpr int TILE_W = 8;            // warps per CTA = sub-ranges per tile
constexpr int TILE_SHIFT = kSubShift;         // docs per sub-range = 2^kSubShift (br_common.cuh; equals br_index::sub_shift)
constexpr int TILE_S = 1 << TILE_SHIFT;
constexpr int TILE_DOCS_SHIFT = TILE_SHIFT + 3;   // log2(W * S)
constexpr int TILE_QT = 20;          // max distinct hot terms of one query in the regular pass
constexpr int TILE_QT_LONG = 40;     // the same in the long-query pass (queries with up to 64 terms, e.g.
                                     // bigram-expanded queries, bm25_ranking.ipynb:105-107); more -> dense path
constexpr int TILE_CAP = 1024;       // candidates kept per query between tighten rounds (k <= 32)
// the same for 32 < k <= 1024: everything passes before the first threshold exists (the first launch covers TILE_CAP_BIG docs),
// later chunks add about k each
constexpr int TILE_CAP_BIG = (4 << kSubShift) > 8192 ? (4 << kSubShift) : 8192;
static_assert((1 << 3) == TILE_W, "TILE_DOCS_SHIFT assumes W == 8");
// queries per CTA: the fp32 accumulators [G][W][S] stay at 64 KB so that 3 CTAs share an SM
// tiles of the first launch (thresholds come from the seeding only): at most 32k docs, so that the docs passing a loose
// threshold fit the candidate list
constexpr int TILE_CHUNK0 = (32768 >> TILE_DOCS_SHIFT) >= 1 ? (32768 >> TILE_DOCS_SHIFT) : 1;
constexpr int TILE_GMAX = (64 * 1024) / (TILE_W * TILE_S * 4) >= 4 ? 4 : ((64 * 1024) / (TILE_W * TILE_S * 4) >= 2 ? 2 : 1);

struct __align__(16) TileEntry {
    int32_t term;
    int32_t slot;
    uint8_t mult[4];     // multiplicity of the term in each query of the group (0: the query lacks it)
    uint8_t pos[4];      // index of the term in that query's unique-term list (bit of the query's deferral mask)
};

// a deferred term of one query: look-up row and multiplicity
struct __align__(8) NeEntry {
    int32_t row;
    float mult;
};
constexpr int NE_MAX = 32;           // deferred terms per query (a query on this path has <= 32 terms)

// cold entry: doc offset inside the tile (12 bits) | query slot in the group (3 bits) << 12, weight
struct __align__(8) ColdEntry {
    uint32_t key;
    float w;
};

// ------------------------------------------------------------------------------------------
// group preparation: union of the hot terms of G consecutive queries, per-query multiplicities
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tile_prep(PrepView pv, const int32_t* __restrict__ q_off,
                                                   const int32_t* __restrict__ hot_slot,
                                                   const int64_t* __restrict__ row_ptr, int32_t nq, int G, int umax,
                                                   TileEntry* __restrict__ entries, int32_t* __restrict__ n_entries,
                                                   int32_t* __restrict__ elig, unsigned long long* __restrict__ cold_total,
                                                   int64_t dense_df_min, const int32_t* __restrict__ perm,
                                                   const int16_t* __restrict__ row_slot, int n_srows, int qt_max,
                                                   int max_terms) {
    __shared__ int32_t s_term[4][256];
    __shared__ uint8_t s_q[4][256], s_m[4][256], s_p[4][256];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int g = blockIdx.x * 4 + wl;
    const int n_groups = (nq + G - 1) / G;
    if (g >= n_groups) return;
    int n = 0;
    unsigned long long cold_sum = 0;
    for (int i = 0; i < G; ++i) {
        if (g * G + i >= nq) break;
        const int q = perm[g * G + i];
        const int32_t off = q_off[q], nu = pv.u_cnt[q];
        int hot = 0, bad = 0;
        unsigned long long cold = 0;
        for (int base = 0; base < nu; base += 32) {
            const int j = base + lane;
            const int32_t t = j < nu ? pv.u_terms[off + j] : -1;
            const bool is_hot = t >= 0 && hot_slot[t] >= 0;
            if (t >= 0 && !is_hot) cold += (unsigned long long)(row_ptr[t + 1] - row_ptr[t]);
            if (t >= 0 && pv.u_mult[off + j] > 255) bad = 1;
            hot += __popc(__ballot_sync(0xffffffffu, is_hot));
        }
        for (int o = 16; o > 0; o >>= 1) cold += __shfl_xor_sync(0xffffffffu, cold, o);
        bad = __any_sync(0xffffffffu, bad);
        const bool ok = hot <= qt_max && !bad && pv.o_cnt[q] <= max_terms;   // longer queries: long pass / dense path
        if (lane == 0) elig[q] = ok ? 1 : 0;
        if (!ok) continue;
        cold_sum += cold;
        for (int base = 0; base < nu; base += 32) {
            const int j = base + lane;
            const int32_t t = j < nu ? pv.u_terms[off + j] : -1;
            const bool is_hot = t >= 0 && hot_slot[t] >= 0;
            const unsigned m = __ballot_sync(0xffffffffu, is_hot);
            if (is_hot) {
                const int p = n + __popc(m & ((1u << lane) - 1));
                s_term[wl][p] = t;
                s_q[wl][p] = (uint8_t)i;
                s_m[wl][p] = (uint8_t)pv.u_mult[off + j];
                s_p[wl][p] = (uint8_t)j;
            }
            n += __popc(m);
        }
    }
    __syncwarp();
    // union of the group's hot terms, terms that are dense everywhere (df >= dense_df_min) first
    int U = 0, Ud = 0, Ur = 0;
    for (int pass = -1; pass < 2; ++pass) {
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            bool first = i < n;
            if (first) {
                const int32_t t = s_term[wl][i];
                const bool is_row = row_slot[t] >= 0 && row_slot[t] < n_srows;      // streamed row
                const bool dense = row_ptr[t + 1] - row_ptr[t] >= dense_df_min;
                const int cls = is_row ? -1 : (dense ? 0 : 1);          // rows first, then dense postings, then sparse
                if (cls != pass) first = false;
                for (int j = 0; first && j < i; ++j)
                    if (s_term[wl][j] == t) first = false;
            }
            const unsigned m = __ballot_sync(0xffffffffu, first);
            if (first) {
                TileEntry e;
                e.term = s_term[wl][i];
                e.slot = pass < 0 ? (int32_t)row_slot[e.term] : hot_slot[e.term];
#pragma unroll
                for (int x = 0; x < 4; ++x) e.mult[x] = e.pos[x] = 0;
                for (int j = i; j < n; ++j)
                    if (s_term[wl][j] == e.term) { e.mult[s_q[wl][j]] = s_m[wl][j]; e.pos[s_q[wl][j]] = s_p[wl][j]; }
                entries[(int64_t)g * umax + U + __popc(m & ((1u << lane) - 1))] = e;
            }
            U += __popc(m);
        }
        if (pass < 0) Ur = U;
        if (pass == 0) Ud = U;
    }
    if (lane == 0) {
        n_entries[g] = U | (Ud << 10) | (Ur << 20);
        if (cold_sum) atomicAdd(cold_total, cold_sum);
    }
}

// ------------------------------------------------------------------------------------------
// cold postings -> (tile, group) buckets: counting sort
// ------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void k_cold_pass(PrepView pv, const int32_t* __restrict__ q_off, const int32_t* __restrict__ hot_slot,
                            const int64_t* __restrict__ row_ptr, const br_posting* __restrict__ post, int32_t nq, int G,
                            int n_groups, const int32_t* __restrict__ elig, const int32_t* __restrict__ inv_perm,
                            uint32_t* __restrict__ counter, ColdEntry* __restrict__ out, uint32_t out_cap) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= nq || !elig[q]) return;
    const int32_t off = q_off[q], nu = pv.u_cnt[q];
    const int pos = inv_perm[q];
    const int g = pos / G;
    const uint32_t slot = (uint32_t)(pos - g * G);
    for (int j = 0; j < nu; ++j) {
        const int32_t t = pv.u_terms[off + j];
        if (hot_slot[t] >= 0) continue;
        const float mult = (float)pv.u_mult[off + j];
        const int64_t lo = row_ptr[t], n = row_ptr[t + 1] - lo;
        for (int64_t i = lane; i < n; i += 32) {
            const br_posting p = post[lo + i];
            const uint32_t tile = p.doc >> TILE_DOCS_SHIFT;
            const int64_t b = (int64_t)tile * n_groups + g;
            const uint32_t pos = atomicAdd(counter + b, 1u);
            if (SCATTER && pos < out_cap) out[pos] = ColdEntry{(p.doc & ((1u << TILE_DOCS_SHIFT) - 1)) | (slot << TILE_DOCS_SHIFT), p.w * mult};
        }
    }
}

// three-kernel exclusive scan of uint32 counts (n up to ~2^31): out[i] = sum_{j<i} in[j], out[n] = total
constexpr int SCAN_T = 256, SCAN_I = 16, SCAN_TILE = SCAN_T * SCAN_I;
__global__ void __launch_bounds__(SCAN_T) k_scan_reduce(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ part) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) {
        const int64_t i = base + j * SCAN_T + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const uint32_t* __restrict__ in, int64_t n,
                                                       const int64_t* __restrict__ part_off, uint32_t* __restrict__ out,
                                                       uint32_t* __restrict__ out_copy) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I], s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) { v[j] = base + j < n ? in[base + j] : 0u; s += v[j]; }
    uint32_t total;
    uint32_t run = block_excl_scan(s, &total) + (uint32_t)part_off[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) {
        if (base + j < n) { out[base + j] = run; out_copy[base + j] = run; }
        run += v[j];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_T - 1) { out[n] = run; }
}

// ------------------------------------------------------------------------------------------
// the tile kernel
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sort_desc(float v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? fmaxf(v, o) : fminf(v, o);
        }
    }
    return v;   // lane i holds the i-th largest
}


// ---- dense slice of one term, specialised on the set of queries (MASK) of the group that contain it ----
// cur[] holds the first chunk (128 postings, 4 per lane) already loaded; chunks are processed with the
// next chunk (of this term, or the first chunk of the next dense term: np/nstart/nhi) in flight.
// Full chunks carry no predicates at all; only the last, partial chunk of a slice is predicated per lane.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}

// Read-modify-write of up to 4 postings per lane into the accumulator rows of the queries in MASK
// (straight-line: all LDS, then the adds, then all STS; ok[x] predicates the lanes of a partial chunk).
template <int G, unsigned MASK, bool MULT>
__device__ __forceinline__ void rmw_chunk(const uint32_t (&a)[4], const bool (&ok)[4], const uint2 (&cur)[4],
                                          const float (&fm)[G]) {
    constexpr uint32_t ROWB = TILE_W * TILE_S * 4;                        // bytes between query rows
    float r[4][G];
#pragma unroll
    for (int q = 0; q < G; ++q)
        if (MASK & (1u << q)) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if (ok[x]) r[x][q] = lds_f32(a[x] + q * ROWB);
        }
#pragma unroll
    for (int q = 0; q < G; ++q)
        if (MASK & (1u << q)) {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float wt = __uint_as_float(cur[x].y);
                if (ok[x]) sts_f32(a[x] + q * ROWB, MULT ? fmaf(wt, fm[q], r[x][q]) : r[x][q] + wt);
            }
        }
}

template <int G, bool MULT, unsigned MASK = 1>
__device__ __forceinline__ void rmw_dispatch(unsigned mask, const uint32_t (&a)[4], const bool (&ok)[4],
                                             const uint2 (&cur)[4], const float (&fm)[G]) {
    if constexpr (MASK < (1u << G)) {
        if (mask == MASK) rmw_chunk<G, MASK, MULT>(a, ok, cur, fm);
        else rmw_dispatch<G, MULT, MASK + 1>(mask, a, ok, cur, fm);
    }
}

struct TileArgs {
    const br_posting* post;
    const int64_t* row_ptr;
    const uint32_t* skip;
    int32_t n_sub;
    int64_t n_docs;
    const TileEntry* entries;
    const int32_t* n_entries;
    int umax;
    int32_t nq;
    int n_groups;
    const int32_t* elig;
    const uint32_t* cold_off;     // [n_tiles * n_groups + 1] or null
    const ColdEntry* cold;
    uint32_t cold_cap;            // entries the cold buffer holds (a batch that needs more is repeated with a larger one)
    const float* rows;            // dense rows [n_rows, n_pad]
    int64_t n_pad;
    float* thr;
    int32_t* cand_cnt;
    int32_t* cand;                // [nq, cap]
    float* cand_h;                // [nq, cap]
    int cap;                      // candidate slots per query (TILE_CAP for k <= 32, TILE_CAP_BIG above)
    int K;
    int tile0;
    int tile_end;                 // tiles [tile0, tile_end) in this launch
    int sub_begin, sub_end;       // sub-ranges [sub_begin, sub_end) of those tiles (the large-k path starts with part of a tile)
    int tpb;                      // consecutive tiles per CTA
    int has_mult;                 // 0: every multiplicity is 1 (set(query) semantics)
    const int32_t* perm;          // group g, slot i -> query perm[g*G+i] (queries sorted by frequent-term signature)
    const uint32_t* defer_mask;   // [nq] bit j: the j-th unique term of the query is deferred in this launch
    const float* ne_ub;           // [nq] sum of the deferred terms' upper bounds (rounded up)
    const int16_t* row_slot;      // [V] look-up row of a term (every deferred term has one)
};

// One fp32 select out of G registers by a run-time index (keeps per-query scalars out of local memory).
template <int G>
__device__ __forceinline__ float sel_q(const float (&v)[G], int q) {
    float r = v[0];
#pragma unroll
    for (int i = 1; i < G; ++i) r = q == i ? v[i] : r;
    return r;
}

#ifndef BR_LIST_CAP
#define BR_LIST_CAP 64
#endif
constexpr int LIST_CAP = BR_LIST_CAP;      // crossing-list slots per warp and sub-range (overflow -> full scan of the sub-range)

// MaxScore deferral (exact): between launches every query gets a plan (make_plan): the terms with the smallest upper
// bounds ub[t] = max posting weight, as long as their sum stays below a fraction of the query's threshold, are
// "deferred" - not streamed at all.  A doc whose partial score p over the streamed ("essential") terms satisfies
// p + sum(deferred ub) < thr*(1-band) cannot reach the threshold and is never looked at; the few others are completed
// by direct look-ups before the usual filter: every deferrable term (df >= N/64) has a look-up row w[row][doc], so a
// completion is a handful of independent 4-byte loads, no search.  The densest terms have the smallest bounds, so in
// the steady state neither the streamed rows nor the dense posting slices are read.
// Without streamed rows every accumulator starts at zero and only grows, so the lane that has just updated one knows
// whether it has reached t1 = thr*(1-band) - sum(deferred ub): it records (query, doc) in its warp's crossing list,
// and after the posting phases the warp completes and filters just the listed docs (an atomic exchange on the
// accumulator removes duplicates).  The S accumulators of a (query, sub-range) are scanned only for queries that
// still stream rows / have no threshold yet, or when a crossing list overflows.
// After the CTA prologue there is no block barrier: every warp fetches the slice bounds of its own sub-range (one tile
// ahead, into registers) and its own copy of the thresholds.
template <int G, int QT, int SPM>
__global__ void __launch_bounds__(TILE_W * 32, 3) k_tile_score(TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* acc = reinterpret_cast<float*>(smem_raw);                                     // [G][W][S]
    TileEntry* ent = reinterpret_cast<TileEntry*>(smem_raw + sizeof(float) * G * TILE_W * TILE_S);   // streamed entries
    int64_t* s_base = reinterpret_cast<int64_t*>(ent + a.umax);                          // [umax]
    float4* s_fm = reinterpret_cast<float4*>(s_base + a.umax);                           // [umax] multiplicities as floats
    uint32_t* s_skiprow = reinterpret_cast<uint32_t*>(s_fm + a.umax);                    // [umax] slot * (n_sub+1)
    NeEntry* s_ne = reinterpret_cast<NeEntry*>(s_skiprow + a.umax);                      // [G][NE_MAX] deferred terms per query
    uint2* s_bnd_all = reinterpret_cast<uint2*>(s_ne + G * NE_MAX);                      // [W][umax] slice (lo, hi) per warp
    uint16_t* s_list_all = reinterpret_cast<uint16_t*>(s_bnd_all + TILE_W * a.umax);     // [W][LIST_CAP]
    // small per-CTA / per-warp state, carved from the same buffer (static __shared__ arrays would have their window
    // addresses rematerialised all over the kernel)
    int* s_qi = reinterpret_cast<int*>(s_list_all + TILE_W * LIST_CAP);                  // [4] query of each group slot (-1: none)
    int* s_cnt = s_qi + 4;                                                               // Ur, Ud, U (streamed), rows mask
    int* s_nne = s_cnt + 4;                                                              // [4] deferred terms of each query (padded to x4)
    float* s_neub = reinterpret_cast<float*>(s_nne + 4);                                 // [4]
    float* s_thq_all = s_neub + 4;                                                       // [W][4] per-warp thresholds of the tile
    float* s_t1_all = s_thq_all + TILE_W * 4;                                            // [W][4]
    int* s_lcnt_all = reinterpret_cast<int*>(s_t1_all + TILE_W * 4);                     // [W]
    int* s_nql = s_lcnt_all + TILE_W;                                                    // [4] sparse streamed terms of each query
    uint8_t* s_ql = reinterpret_cast<uint8_t*>(s_nql + 4);                               // [G][QT] their entry indices

    const int g = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile_first = a.tile0 + blockIdx.y * a.tpb;
    const int tile_last = min(tile_first + a.tpb, a.tile_end);          // exclusive

    if (threadIdx.x < G) {
        const int pos = g * G + threadIdx.x;
        const int qi = pos < a.nq ? a.perm[pos] : -1;
        const bool ok = qi >= 0 && a.elig[qi];
        s_qi[threadIdx.x] = ok ? qi : -1;
        s_neub[threadIdx.x] = ok ? a.ne_ub[qi] : 0.f;
    }
    __syncthreads();
    if (w == 0) {
        // split the group's entries into streamed and deferred ones (stable: rows, dense postings, sparse)
        const int U0 = a.n_entries[g] & 0x3ff, Ud0 = (a.n_entries[g] >> 10) & 0x3ff, Ur0 = a.n_entries[g] >> 20;
        uint32_t dm[G];
#pragma unroll
        for (int i = 0; i < G; ++i) dm[i] = s_qi[i] >= 0 ? a.defer_mask[s_qi[i]] : 0u;
        int na = 0, ur = 0, ud = 0;
        int nne[G], nql[G];
#pragma unroll
        for (int i = 0; i < G; ++i) nne[i] = nql[i] = 0;
        unsigned rows_q = 0;
        for (int base = 0; base < U0; base += 32) {
            const int u = base + lane;
            TileEntry e{};
            float act[4] = {0.f, 0.f, 0.f, 0.f}, def[4] = {0.f, 0.f, 0.f, 0.f};
            bool isa = false;
            if (u < U0) {
                e = a.entries[(int64_t)g * a.umax + u];
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    if (e.mult[i] && s_qi[i] >= 0) {
                        if (e.pos[i] < 32 && ((dm[i] >> e.pos[i]) & 1u)) def[i] = (float)e.mult[i];
                        else { act[i] = (float)e.mult[i]; isa = true; if (u < Ur0) rows_q |= 1u << i; }
                    }
                }
            }
            const unsigned ma = __ballot_sync(0xffffffffu, isa);
            const unsigned lt = (1u << lane) - 1;
#pragma unroll
            for (int i = 0; i < G; ++i) {                                        // per-query lists of streamed sparse terms
                const bool in = isa && u >= Ud0 && act[i] != 0.f;
                const unsigned mq = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const int p = nql[i] + __popc(mq & lt);
                    if (p < QT) s_ql[i * QT + p] = (uint8_t)(na + __popc(ma & lt));
                }
                nql[i] += __popc(mq);
            }
            if (isa) {
                const int p = na + __popc(ma & lt);
                ent[p] = e;
                s_base[p] = a.row_ptr[e.term];
                s_fm[p] = make_float4(act[0], act[1], act[2], act[3]);
                s_skiprow[p] = (uint32_t)e.slot * (uint32_t)(a.n_sub + 1);      // row terms: slot = row index (unused here)
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {                                        // per-query lists of deferred terms
                const unsigned md = __ballot_sync(0xffffffffu, def[i] != 0.f);
                if (def[i] != 0.f) {
                    const int p = nne[i] + __popc(md & lt);
                    if (p < NE_MAX) s_ne[i * NE_MAX + p] = NeEntry{(int32_t)a.row_slot[e.term], def[i]};
                }
                nne[i] += __popc(md);
            }
            ur += __popc(ma & __ballot_sync(0xffffffffu, u < Ur0));
            ud += __popc(ma & __ballot_sync(0xffffffffu, u < Ud0));
            na += __popc(ma);
        }
        rows_q = __reduce_or_sync(0xffffffffu, rows_q);
        if (lane == 0) { s_cnt[0] = ur; s_cnt[1] = ud; s_cnt[2] = na; s_cnt[3] = (int)rows_q; }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int n = min(nne[i], NE_MAX), n4 = (n + 3) & ~3;
            if (lane >= n && lane < n4) s_ne[i * NE_MAX + lane] = NeEntry{0, 0.f};     // padding: row 0, multiplicity 0
            if (lane == 0) { s_nne[i] = n4; s_nql[i] = min(nql[i], QT); }
        }
    }
    __syncthreads();                                   // the only block barrier: from here on the warps run independently
    const int Ur = s_cnt[0], Ud = s_cnt[1], U = s_cnt[2];
    const unsigned rows_q = (unsigned)s_cnt[3];
    uint2* s_bnd = s_bnd_all + w * a.umax;
    uint16_t* s_list = s_list_all + w * LIST_CAP;
    float* s_thq = s_thq_all + w * 4;
    float* s_t1 = s_t1_all + w * 4;
    int* s_lcnt = s_lcnt_all + w;

    // slice bounds of this warp's sub-range: lane covers entries lane, lane+32, ...; fetched one tile ahead
    constexpr int NR = (QT * G + 31) / 32;
    uint32_t nlo[NR], nhi[NR];
    auto load_bounds = [&](int sub) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int u = lane + 32 * r;
            nlo[r] = nhi[r] = 0;
            if (u >= Ur && u < U) {
                const uint32_t* sk = a.skip + s_skiprow[u] + sub;
                nlo[r] = __ldg(sk);
                nhi[r] = __ldg(sk + 1);
            }
        }
    };
    auto store_bounds = [&]() {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int u = lane + 32 * r;
            if (u < U) s_bnd[u] = make_uint2(nlo[r], nhi[r]);     // row terms keep empty slices
        }
    };
    // (large-k path: a launch may start inside its first tile - the warps below sub_begin start one tile later)
    const int tile_begin_w = tile_first + ((tile_first * TILE_W + w < a.sub_begin) ? 1 : 0);
    if (tile_begin_w * TILE_W + w < a.n_sub) load_bounds(tile_begin_w * TILE_W + w);
#pragma unroll 1
    for (int tile = tile_begin_w; tile < tile_last; ++tile) {
    const int sub = tile * TILE_W + w;
    if (sub >= a.sub_end) break;                                      // warp-uniform; the later tiles are out of range too
    store_bounds();
    if (tile + 1 < tile_last && sub + TILE_W < a.n_sub) load_bounds(sub + TILE_W);     // in flight during this tile
    // thresholds of this tile (a stale, lower value is always valid); t1: partial scores below it cannot reach thr.
    // Queries that are scanned anyway (streamed rows / no threshold yet) never push: their t1 is +inf in shared memory.
    if (lane < G) {
        const int qi = s_qi[lane];
        const float th = qi >= 0 ? __ldcg(a.thr + qi) : 0.f;
        const float t = th * (1.f - kBandRel) * (1.f - kBandRel) - s_neub[lane];
        const bool scan = ((rows_q >> lane) & 1u) || !(t > 0.f);
        s_thq[lane] = th;
        s_t1[lane] = (scan || qi < 0) ? CUDART_INF_F : t;
    }
    if (lane == 0) *s_lcnt = 0;
    // bucket bounds of this (tile, group)'s cold postings: requested here, used after the posting phases
    uint32_t cold0 = 0, cold1 = 0;
    if (a.cold_off) {
        const uint32_t* co = a.cold_off + ((int64_t)tile * a.n_groups + g);
        cold0 = __ldg(co);
        cold1 = __ldg(co + 1);
    }
    __syncwarp();
    float t1[G];
    unsigned scan_q = 0;
#pragma unroll
    for (int q = 0; q < G; ++q) {
        t1[q] = s_t1[q];
        if (t1[q] == CUDART_INF_F && s_qi[q] >= 0) scan_q |= 1u << q;
    }
    const uint32_t doc0 = (uint32_t)sub << TILE_SHIFT;
    float* my = acc + w * TILE_S;                       // + q * TILE_W * TILE_S per query
    constexpr int ROW = TILE_W * TILE_S;
    auto push = [&](uint32_t q, uint32_t idx) {          // record a crossing (rare)
        const int p = atomicAdd(s_lcnt, 1);
        if (p < LIST_CAP) s_list[p] = (uint16_t)((q << TILE_SHIFT) | idx);
    };
    if (Ur == 0) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < G; ++q)
#pragma unroll
            for (int s4 = 0; s4 < TILE_S / 128; ++s4) reinterpret_cast<float4*>(my + q * ROW)[s4 * 32 + lane] = z;
    } else {
        // each warp initialises its own accumulator rows: zero + the weights of the streamed dense-row terms of each
        // query (terms present in >= ~20% of the docs are stored as plain fp32 rows, 0 where absent), 4 docs per lane
#pragma unroll 1
        for (int seg = 0; seg < TILE_S / 128; ++seg) {          // 128 docs per pass
            float4 v[G];
#pragma unroll
            for (int q = 0; q < G; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int64_t off = (int64_t)doc0 + seg * 128 + lane * 4;
            for (int u0 = 0; u0 < Ur; u0 += 2) {             // two rows in flight
                float4 d[2];
                float ff[2][4];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int u = min(u0 + k, Ur - 1);
                    d[k] = __ldg(reinterpret_cast<const float4*>(a.rows + (int64_t)ent[u].slot * a.n_pad + off));
                    const float4 f4 = u0 + k < Ur ? s_fm[u] : make_float4(0.f, 0.f, 0.f, 0.f);
                    ff[k][0] = f4.x; ff[k][1] = f4.y; ff[k][2] = f4.z; ff[k][3] = f4.w;
                }
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        if (ff[k][q] != 0.f) {               // warp-uniform
                            v[q].x = fmaf(d[k].x, ff[k][q], v[q].x);
                            v[q].y = fmaf(d[k].y, ff[k][q], v[q].y);
                            v[q].z = fmaf(d[k].z, ff[k][q], v[q].z);
                            v[q].w = fmaf(d[k].w, ff[k][q], v[q].w);
                        }
                    }
            }
#pragma unroll
            for (int q = 0; q < G; ++q) reinterpret_cast<float4*>(my + q * ROW)[seg * 32 + lane] = v[q];
        }
    }
    __syncwarp();

    if constexpr (SPM == 0) {
    // Phase A - sparse slices (most (term, sub-range) pairs hold 0-8 postings).  The lanes are split into G groups of
    // LQ = 32/G, one per query; a group walks the non-empty slices of its own query's terms, one term at a time, LQ
    // postings per step.  The lanes of a group then touch distinct docs of one term, and different groups different
    // accumulator rows: no two lanes of an instruction ever meet in one accumulator, so there is nothing to detect,
    // and neither a prefix scan nor an owner search is needed.  Lane lq of a group first fetches the slice of the
    // group's lq-th term; the walk then pulls the non-empty ones over by shuffle.
    {
        constexpr int LQ = 32 / G;
        constexpr unsigned GMASK = LQ == 32 ? 0xffffffffu : ((1u << LQ) - 1u);
        const int qg = lane / LQ, lq = lane % LQ;
        float* myq = my + qg * ROW;
        const float tq = s_t1[qg];
        const int nt = s_nql[qg];
        int nt_max = nt;
#pragma unroll
        for (int o = LQ; o < 32; o <<= 1) nt_max = max(nt_max, __shfl_xor_sync(0xffffffffu, nt_max, o));
#pragma unroll 1
        for (int r0 = 0; r0 < nt_max; r0 += LQ) {
            uint32_t t_lo = 0, t_hi = 0, t_pb = 0;
            float t_mu = 1.f;
            if (r0 + lq < nt) {
                const int u = s_ql[qg * QT + r0 + lq];
                const uint2 bd = s_bnd[u];
                t_lo = bd.x; t_hi = bd.y;
                t_pb = (uint32_t)s_base[u];                         // posting index of the term's list (nnz < 2^32)
                if (a.has_mult) t_mu = reinterpret_cast<const float*>(s_fm + u)[qg];
            }
            unsigned gm = (__ballot_sync(0xffffffffu, t_lo < t_hi) >> (qg * LQ)) & GMASK;     // non-empty slices of my group
            // software pipeline of depth 1: the posting of step i+1 is requested before the one of step i is added (the walk is
            // a chain of L2 round trips otherwise - ncu: 15 % of all stall samples on the first use of the loaded posting)
            uint32_t c = 0, h = 0, pb = 0;
            float mu = 1.f;
            bool done = false;
            uint2 pv = make_uint2(0u, 0u);
            float pmu = 1.f;
            bool pvalid = false;
            while (true) {
                const bool need = c >= h;                           // group-uniform: current slice exhausted
                const bool take = need && gm != 0;
                const int src = qg * LQ + (take ? __ffs(gm) - 1 : lq);
                const uint32_t n_lo = __shfl_sync(0xffffffffu, t_lo, src);
                const uint32_t n_hi = __shfl_sync(0xffffffffu, t_hi, src);
                const uint32_t n_pb = __shfl_sync(0xffffffffu, t_pb, src);
                if (a.has_mult) { const float n_mu = __shfl_sync(0xffffffffu, t_mu, src); if (take) mu = n_mu; }
                if (take) { c = n_lo; h = n_hi; pb = n_pb; gm &= gm - 1; }
                else if (need) done = true;
                uint2 nv = make_uint2(0u, 0u);
                bool nvalid = false;
                if (!done) {
                    const uint32_t pos = c + lq;
                    if (pos < h) { nv = __ldg(reinterpret_cast<const uint2*>(a.post) + pb + pos); nvalid = true; }
                    c += LQ;
                }
                if (pvalid) {                                       // the step requested one iteration ago
                    const uint32_t idx = pv.x - doc0;
                    const float val = fmaf(__uint_as_float(pv.y), pmu, myq[idx]);
                    myq[idx] = val;
                    if (val >= tq) push(qg, idx);
                }
                __syncwarp();
                pv = nv; pmu = mu; pvalid = nvalid;
                if (__all_sync(0xffffffffu, done && !pvalid)) break;
            }
        }
    }
    __syncwarp();
    } else {
    // Phase A, second form (sparse_mode 1): every lane group walks the term list of its own query straight from shared
    // memory - slice bounds, list base and multiplicity of term i are broadcast reads, the posting of term i+1 is
    // requested before term i is added - instead of pulling the next non-empty slice over from the lane that fetched its
    // bounds (three shuffles, a find-first-set and two votes per step in the first form; ncu put 42 % of the stall
    // samples of the 1024-tile launch on that loop, most of them branch-resolving / short-scoreboard, not memory).
    {
        constexpr int LQ = 32 / G;
        const int qg = lane / LQ, lq = lane % LQ;
        float* myq = my + qg * ROW;
        const float tq = s_t1[qg];
        const int nt = s_nql[qg];
        int nt_max = nt;
#pragma unroll
        for (int o = LQ; o < 32; o <<= 1) nt_max = max(nt_max, __shfl_xor_sync(0xffffffffu, nt_max, o));
        const uint2* const post2 = reinterpret_cast<const uint2*>(a.post);
        uint32_t c_hi = 0, c_lo = 0, c_pb = 0;
        float c_mu = 1.f;
        uint2 c_p = make_uint2(0u, 0u);
        auto fetch = [&](int it, uint32_t& lo, uint32_t& hi, uint32_t& pb, float& mu, uint2& p) {
            lo = hi = 0;
            if (it < nt) {
                const int u = s_ql[qg * QT + it];
                const uint2 bd = s_bnd[u];
                lo = bd.x; hi = bd.y;
                pb = (uint32_t)s_base[u];                           // posting index of the term's list (nnz < 2^32)
                if (a.has_mult) mu = reinterpret_cast<const float*>(s_fm + u)[qg];
                if (lo + lq < hi) p = __ldg(post2 + pb + lo + lq);
            }
        };
        fetch(0, c_lo, c_hi, c_pb, c_mu, c_p);
#pragma unroll 1
        for (int it = 0; it < nt_max; ++it) {
            uint32_t n_lo, n_hi, n_pb = 0;
            float n_mu = 1.f;
            uint2 n_p = make_uint2(0u, 0u);
            fetch(it + 1, n_lo, n_hi, n_pb, n_mu, n_p);
            if (c_lo + lq < c_hi) {
                const uint32_t idx = c_p.x - doc0;
                const float val = fmaf(__uint_as_float(c_p.y), c_mu, myq[idx]);
                myq[idx] = val;
                if (val >= tq) push(qg, idx);
            }
            // a slice longer than the lane group (rare for a sparse-class term)
            for (uint32_t c = c_lo + LQ; __any_sync(0xffffffffu, c < c_hi); c += LQ) {
                __syncwarp();
                if (c + lq < c_hi) {
                    const uint2 p = __ldg(post2 + c_pb + c + lq);
                    const uint32_t idx = p.x - doc0;
                    const float val = fmaf(__uint_as_float(p.y), c_mu, myq[idx]);
                    myq[idx] = val;
                    if (val >= tq) push(qg, idx);
                }
            }
            __syncwarp();
            c_lo = n_lo; c_hi = n_hi; c_pb = n_pb; c_mu = n_mu; c_p = n_p;
        }
    }
    __syncwarp();
    }
    // Phase B - dense slices (terms with >= tile_dense_min postings per sub-range on average that are still streamed:
    // rare once the deferral plans exist), whole warp per term, 64 postings per step.  Inside a term all docs are
    // distinct, so the read-modify-writes of a step are independent.  No atomics.
#pragma unroll 1
    for (int u = Ur; u < Ud; ++u) {
        const uint2 bd = s_bnd[u];
        if (bd.x >= bd.y) continue;
        const float4 f4 = s_fm[u];
        const float ff[4] = {f4.x, f4.y, f4.z, f4.w};
        const uint2* p = reinterpret_cast<const uint2*>(a.post) + s_base[u];
        // the next 64 postings are requested before the current 64 are added (ncu: 10 % of the stall samples of the
        // 2048-doc version sat on the first use of an unprefetched chunk)
        uint2 nv2[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            nv2[x] = make_uint2(0xffffffffu, 0u);
            if (bd.x + lane + 32 * x < bd.y) nv2[x] = __ldg(p + bd.x + lane + 32 * x);
        }
        for (uint32_t c = bd.x; c < bd.y; c += 64) {
            uint2 v[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                v[x] = nv2[x];
                nv2[x] = make_uint2(0xffffffffu, 0u);
                if (c + 64 + lane + 32 * x < bd.y) nv2[x] = __ldg(p + c + 64 + lane + 32 * x);
            }
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                if (v[x].x != 0xffffffffu) {
                    const uint32_t idx = v[x].x - doc0;
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        if (ff[q] != 0.f) {                    // warp-uniform
                            const float nv = fmaf(__uint_as_float(v[x].y), ff[q], my[q * ROW + idx]);
                            my[q * ROW + idx] = nv;
                            if (nv >= t1[q]) push(q, idx);
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    // cold postings of this (tile, group) bucket that fall in this warp's sub-range
    if (a.cold_off) {
        const uint32_t c0 = cold0, c1 = min(cold1, a.cold_cap);
        for (uint32_t i = c0 + lane; i < c1; i += 32) {
            const ColdEntry ce = a.cold[i];
            const uint32_t l = ce.key & ((1u << TILE_DOCS_SHIFT) - 1), q = ce.key >> TILE_DOCS_SHIFT;
            if ((l >> TILE_SHIFT) == (uint32_t)w) {
                const uint32_t idx = l & (TILE_S - 1);
                const float nv = atomicAdd(acc + (q * TILE_W + w) * TILE_S + idx, ce.w) + ce.w;
                if (nv >= s_t1[q]) push(q, idx);
            }
        }
    }
    __syncwarp();

    // crossing list: complete (deferred terms by look-up row) and filter exactly the listed docs, one per lane
    const int n_list = *s_lcnt;
    if (n_list > LIST_CAP) {
        scan_q = (1u << G) - 1;                                        // overflow: scan every query's sub-range instead
    } else {
        for (int i0 = 0; i0 < n_list; i0 += 32) {
            const int i = i0 + lane;
            if (i < n_list) {
                const uint32_t en = s_list[i];
                const uint32_t q = en >> TILE_SHIFT, idx = en & (TILE_S - 1);
                // a doc is listed once per update at or above t1: the first taker marks the accumulator
                float full = atomicExch(my + q * ROW + idx, -1.f);
                if (full > 0.f) {
                    const int nne = s_nne[q];
                    const NeEntry* ne = s_ne + q * NE_MAX;
                    const float* col = a.rows + (doc0 + idx);
                    for (int x0 = 0; x0 < nne; x0 += 4) {          // lists are padded to a multiple of 4
                        float wv[4], mm[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const NeEntry e2 = ne[x0 + x];
                            mm[x] = e2.mult;
                            wv[x] = __ldg(col + (int64_t)e2.row * a.n_pad);
                        }
#pragma unroll
                        for (int x = 0; x < 4; ++x) full = fmaf(wv[x], mm[x], full);
                    }
                    if (full >= s_thq[q] * (1.f - kBandRel)) {
                        const int qi = s_qi[q];
                        const int pos = atomicAdd(a.cand_cnt + qi, 1);
                        if (pos < a.cap) {
                            a.cand[(int64_t)qi * a.cap + pos] = (int32_t)(doc0 + idx);
                            a.cand_h[(int64_t)qi * a.cap + pos] = full;
                        }
                    }
                }
            }
        }
    }
    __syncwarp();

    // scan epilogue: threshold filter over all S accumulators of a query (streamed rows, no threshold yet, overflow)
    if (scan_q)
#pragma unroll 1
    for (int q = 0; q < G; ++q) {
        const int qi = s_qi[q];
        if (qi < 0 || !((scan_q >> q) & 1u)) continue;
        float* myq = my + q * ROW;
        float th = s_thq[q];
        const float t1q = th * (1.f - kBandRel) * (1.f - kBandRel) - s_neub[q];
        float v[TILE_S / 32];                                   // v[4*s + e] = doc s*128 + lane*4 + e
        float mx = 0.f;
#pragma unroll
        for (int s = 0; s < TILE_S / 128; ++s) {
            const float4 f = reinterpret_cast<const float4*>(myq)[s * 32 + lane];
            v[4 * s + 0] = f.x; v[4 * s + 1] = f.y; v[4 * s + 2] = f.z; v[4 * s + 3] = f.w;
            mx = fmaxf(fmaxf(mx, fmaxf(f.x, f.y)), fmaxf(f.z, f.w));
        }
        if (!__any_sync(0xffffffffu, mx >= t1q && mx > 0.f)) continue;        // nothing can reach the threshold here
        float lo_thr = th * (1.f - kBandRel);
        const int nne = s_nne[q];
        if (nne > 0) {
            // (overflow path) complete the partial scores that may still reach the threshold, in place
            const NeEntry* ne = s_ne + q * NE_MAX;
            static_assert(TILE_S / 32 <= 64, "completion mask of the scan epilogue: one bit per doc of a lane");
            unsigned long long pm = 0;
#pragma unroll
            for (int j = 0; j < TILE_S / 32; ++j) pm |= (v[j] >= t1q && v[j] > 0.f) ? (1ull << j) : 0ull;
            while (pm) {
                const int j = __ffsll((long long)pm) - 1;
                pm &= pm - 1;
                const uint32_t idx = (uint32_t)((j >> 2) * 128 + lane * 4 + (j & 3));
                const float* col = a.rows + (doc0 + idx);
                float full = myq[idx];
                for (int i = 0; i < nne; ++i) full = fmaf(__ldg(col + (int64_t)ne[i].row * a.n_pad), ne[i].mult, full);
                myq[idx] = full;
            }
            __syncwarp();
            mx = 0.f;
#pragma unroll
            for (int s = 0; s < TILE_S / 128; ++s) {
                const float4 f = reinterpret_cast<const float4*>(myq)[s * 32 + lane];
                v[4 * s + 0] = f.x; v[4 * s + 1] = f.y; v[4 * s + 2] = f.z; v[4 * s + 3] = f.w;
                mx = fmaxf(fmaxf(mx, fmaxf(f.x, f.y)), fmaxf(f.z, f.w));
            }
        }
        int c = 0;
#pragma unroll
        for (int j = 0; j < TILE_S / 32; ++j) c += (v[j] >= lo_thr && v[j] > 0.f) ? 1 : 0;
        int tot = c;
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (tot == 0) continue;
        if (tot >= a.K && a.K <= 32) {
            // K distinct docs (one per lane) score >= the K-th largest lane maximum: a valid lower
            // bound of the K-th best score of this query over the whole shard (a lane maximum that is only a
            // partial score is a lower bound of that doc's score, so the bound stays valid)
            const float srt = warp_sort_desc(mx, lane);
            const float kth = __shfl_sync(0xffffffffu, srt, a.K - 1);
            if (kth > th) {
                th = kth;
                if (lane == 0) atomicMax(reinterpret_cast<int*>(a.thr + qi), __float_as_int(kth));   // scores >= 0
                lo_thr = th * (1.f - kBandRel);
                c = 0;
#pragma unroll
                for (int j = 0; j < TILE_S / 32; ++j) c += (v[j] >= lo_thr && v[j] > 0.f) ? 1 : 0;
            }
        }
        // warp-exclusive offsets, one global atomic per (query, sub-range)
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(a.cand_cnt + qi, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        int pos = base + incl - c;
        int32_t* out = a.cand + (int64_t)qi * a.cap;
        float* out_h = a.cand_h + (int64_t)qi * a.cap;
#pragma unroll
        for (int j = 0; j < TILE_S / 32; ++j) {
            if (v[j] >= lo_thr && v[j] > 0.f) {
                if (pos < a.cap) { out[pos] = (int32_t)(doc0 + (j >> 2) * 128 + lane * 4 + (j & 3)); out_h[pos] = v[j]; }
                ++pos;
            }
        }
    }
    __syncwarp();
    }   // tile loop
}

// Deferral plan of one query for the next launch (one warp; see k_tile_score): greedily defer the hot terms with the
// smallest upper bound mult * ub[t] while their sum stays below frac * thr * (1 - band).  Cold terms (no skip table)
// are never deferred.  ne_ub is rounded up so that "partial + ne_ub < t" really excludes the doc.
struct PlanArgs {
    const int32_t* q_off;
    const int32_t* u_terms;
    const int32_t* u_mult;
    const int32_t* u_cnt;
    const float* ub;
    const int16_t* row_slot;
    int n_srows;
    float frac;
    uint32_t* defer_mask;
    float* ne_ub;
};
// Only terms with a look-up row can be deferred.  The streamed-row terms (the densest, smallest bounds) go first and
// all of them or nothing: a query whose accumulators start from rows needs the full epilogue scan anyway.
__device__ __forceinline__ void make_plan(const PlanArgs& pa, int q, float thr, int lane) {
    const int32_t off = pa.q_off[q], nu = pa.u_cnt[q];
    uint32_t mask = 0;
    float sum = 0.f;
    if (thr > 0.f && pa.frac > 0.f && nu <= 32) {
        float v = CUDART_INF_F;
        bool srow = false;
        if (lane < nu) {
            const int32_t t = pa.u_terms[off + lane];
            const int rs = pa.row_slot[t];
            if (rs >= 0) { v = pa.ub[t] * (float)pa.u_mult[off + lane]; srow = rs < pa.n_srows; }
        }
        const float budget = thr * (1.f - kBandRel) * pa.frac;
        float ssum = srow ? v : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        if (ssum < budget) {
            mask = __ballot_sync(0xffffffffu, srow);
            sum = ssum;
            if (srow) v = CUDART_INF_F;
            for (int it = 0; it < nu; ++it) {
                float best = v;
                int who = lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int ow = __shfl_xor_sync(0xffffffffu, who, o);
                    if (ob < best || (ob == best && ow < who)) { best = ob; who = ow; }
                }
                if (!(sum + best < budget)) break;         // also stops at +inf (nothing deferrable left)
                sum += best;
                mask |= 1u << who;
                if (lane == who) v = CUDART_INF_F;
            }
        }
    }
    if (lane == 0) {
        pa.defer_mask[q] = mask;
        pa.ne_ub[q] = sum * (1.f + 1e-6f);
    }
}

// Threshold seeding (K <= 32): before any tile is scored, thr[q] := the largest, over the query's terms with df >= K,
// of the K-th largest weight among the term's first SEED_MAX postings - K distinct docs score at least that much, so it
// is a valid lower bound of the K-th best score, and the first launches already run with a deferral plan instead of
// streaming every term against a zero threshold.  One warp per query; a running top-32 is kept sorted across the lanes
// (each batch of 32 weights is sorted, reversed and merged bitonically).
constexpr int SEED_MAX = 256;
__global__ void __launch_bounds__(128) k_seed_thr(const br_posting* __restrict__ post, const int64_t* __restrict__ row_ptr,
                                                  int32_t nq, int K, const int32_t* __restrict__ elig, float* __restrict__ thr,
                                                  PlanArgs pa) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= nq || !elig[q]) return;
    const int32_t off = pa.q_off[q], nu = pa.u_cnt[q];
    float best = 0.f;
    for (int j = nu - 1; j >= 0; --j) {         // descending term id: in a frequency-ranked vocabulary the rare (high-bound)
        const int32_t t = pa.u_terms[off + j];  // terms come first and the frequent ones are then skipped by the ub test
        const float m = (float)pa.u_mult[off + j];
        const int64_t lo = row_ptr[t], df = row_ptr[t + 1] - lo;
        if (df < K || !(pa.ub[t] * m > best)) continue;            // cannot raise the bound
        const int n = (int)min(df, (int64_t)SEED_MAX);
        float run = 0.f;                                            // lane i: i-th largest weight so far
        for (int base = 0; base < n; base += 32) {
            const float wv = base + lane < n ? post[lo + base + lane].w * m : 0.f;
            const float srt = warp_sort_desc(fmaxf(wv, 0.f), lane);
            float x = fmaxf(run, __shfl_sync(0xffffffffu, srt, 31 - lane));     // bitonic: top 32 of the union
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float y = __shfl_xor_sync(0xffffffffu, x, o);
                x = (lane & o) ? fminf(x, y) : fmaxf(x, y);
            }
            run = x;
        }
        best = fmaxf(best, __shfl_sync(0xffffffffu, run, K - 1));
    }
    const float th = best * (1.f - 4e-6f);      // fp32 accumulation of the doc's other terms can only add; slack for rounding
    if (lane == 0) thr[q] = th;
    make_plan(pa, q, th, lane);
}

// Between chunks: thr[q] = max(thr[q], K-th best fp32 score emitted so far); keep only the candidates
// inside the band of the new threshold (sorted by score); sticky overflow flag.
constexpr int TG_T = 256, TGS_T = 64;
__global__ void __launch_bounds__(TGS_T) k_tighten(float* __restrict__ thr, int32_t* __restrict__ cand_cnt,
                                                  int32_t* __restrict__ prev_cnt, int32_t* __restrict__ cand,
                                                  float* __restrict__ cand_h, int K, int32_t* __restrict__ overflow,
                                                  PlanArgs pa) {
    __shared__ float s_h[TILE_CAP];
    __shared__ int32_t s_id[TILE_CAP];
    __shared__ int s_keep;
    const int q = blockIdx.x;
    int n = cand_cnt[q];
    if (n == prev_cnt[q]) return;                      // nothing emitted since the last round
    if (n > TILE_CAP) {
        if (threadIdx.x == 0) overflow[q] = 1;
        n = TILE_CAP;
    }
    int32_t* ids = cand + (int64_t)q * TILE_CAP;
    float* hs = cand_h + (int64_t)q * TILE_CAP;
    int n_sort = 32;                                   // smallest power of two covering the list
    while (n_sort < n) n_sort <<= 1;
    for (int i = threadIdx.x; i < n_sort; i += TGS_T) {
        s_h[i] = i < n ? hs[i] : -1.f;
        s_id[i] = i < n ? ids[i] : -1;
    }
    if (threadIdx.x == 0) s_keep = 0;
    for (int size = 2; size <= n_sort; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n_sort / 2; i += TGS_T) {
                const int x = 2 * i - (i & (stride - 1)), y = x + stride;
                const bool up = (x & size) == 0;
                const float hx = s_h[x], hy = s_h[y];
                if ((hy > hx) == up) {
                    s_h[x] = hy; s_h[y] = hx;
                    const int32_t t = s_id[x]; s_id[x] = s_id[y]; s_id[y] = t;
                }
            }
        }
    }
    __syncthreads();
    float th = thr[q];
    if (n >= K && s_h[K - 1] > th) th = s_h[K - 1];
    const float lo = th * (1.f - kBandRel);
    int keep = 0;
    for (int i = threadIdx.x; i < n; i += TGS_T) keep += (s_h[i] >= lo) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    if ((threadIdx.x & 31) == 0 && keep) atomicAdd(&s_keep, keep);
    __syncthreads();
    keep = s_keep;
    for (int i = threadIdx.x; i < n; i += TGS_T) {
        ids[i] = i < keep ? s_id[i] : -1;
        hs[i] = i < keep ? s_h[i] : 0.f;
    }
    if (threadIdx.x == 0) {
        thr[q] = th;
        cand_cnt[q] = keep;
        prev_cnt[q] = keep;
    }
    if (threadIdx.x < 32) make_plan(pa, q, th, threadIdx.x);
}

// The same for long candidate lists (32 < K <= 1024, up to TILE_CAP_BIG entries): the K-th largest score by a 4 x 8-bit
// radix select over the list in shared memory (scores are positive floats, so their bit patterns order like integers)
// instead of a sort; survivors (score >= new threshold x (1 - band)) are compacted to the head of the query's region in
// arbitrary order - every survivor is re-scored in float64 and ordered by k_final_select afterwards.
__global__ void __launch_bounds__(TG_T) k_tighten_big(float* __restrict__ thr, int32_t* __restrict__ cand_cnt,
                                                      int32_t* __restrict__ prev_cnt, int32_t* __restrict__ cand,
                                                      float* __restrict__ cand_h, int K, int cap, int scap,
                                                      int32_t* __restrict__ overflow, PlanArgs pa) {
    // scap <= cap: list entries the shared-memory copy holds in this round (the whole region after the first launch,
    // which can pass a full tile; half of it later, when the lists are ~K long - a longer one goes to the dense path)
    extern __shared__ __align__(16) unsigned char tb_smem[];
    uint32_t* s_h = reinterpret_cast<uint32_t*>(tb_smem);             // [scap] score bits
    int32_t* s_id = reinterpret_cast<int32_t*>(s_h + scap);           // [scap]
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_prefix, s_need;
    const int q = blockIdx.x;
    int n = cand_cnt[q];
    if (n == prev_cnt[q]) return;                      // nothing emitted since the last round
    if (n > scap) {
        if (threadIdx.x == 0) overflow[q] = 1;
        n = scap;
    }
    int32_t* ids = cand + (int64_t)q * cap;
    float* hs = cand_h + (int64_t)q * cap;
    for (int i = threadIdx.x; i < n; i += TG_T) {
        s_h[i] = __float_as_uint(hs[i]);
        s_id[i] = ids[i];
    }
    float th = thr[q];
    if (n >= K) {
        if (threadIdx.x == 0) { s_prefix = 0; s_need = (uint32_t)K; }
        for (int shift = 24; shift >= 0; shift -= 8) {
            s_hist[threadIdx.x] = 0;                  // TG_T == 256 bins
            __syncthreads();
            const uint32_t prefix = s_prefix, need = s_need;
            const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int i = threadIdx.x; i < n; i += TG_T) {
                const uint32_t b = s_h[i];
                if ((b & himask) == prefix) atomicAdd(&s_hist[(b >> shift) & 0xff], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {                   // walk the bins from the top until `need` elements are covered
                uint32_t acc = 0;
                int b = 255;
                for (; b > 0; --b) {
                    if (acc + s_hist[b] >= need) break;
                    acc += s_hist[b];
                }
                s_prefix = prefix | ((uint32_t)b << shift);
                s_need = need - acc;
            }
            __syncthreads();
        }
        const float kth = __uint_as_float(s_prefix);  // exactly the K-th largest score of the list
        if (kth > th) th = kth;
    }
    __syncthreads();
    const uint32_t lo = __float_as_uint(th * (1.f - kBandRel));
    // stable compaction in rounds of TG_T elements
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += TG_T) {
        const int i = i0 + threadIdx.x;
        const bool keep = i < n && s_h[i] >= lo;
        uint32_t total;
        const uint32_t ex = block_excl_scan(keep ? 1u : 0u, &total);
        const int base = s_base;
        if (keep) {
            ids[base + ex] = s_id[i];
            hs[base + ex] = __uint_as_float(s_h[i]);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + (int)total;
        __syncthreads();
    }
    const int kept = s_base;
    for (int i = kept + threadIdx.x; i < n; i += TG_T) {
        ids[i] = -1;
        hs[i] = 0.f;
    }
    if (threadIdx.x == 0) {
        thr[q] = th;
        cand_cnt[q] = kept;
        prev_cnt[q] = kept;
    }
    if (threadIdx.x < 32) make_plan(pa, q, th, threadIdx.x);
}

__global__ void k_fused_flags(const int32_t* __restrict__ elig, const int32_t* __restrict__ overflow,
                              const int32_t* __restrict__ out_cnt, int32_t nq, int32_t need, int positive_only,
                              int32_t* __restrict__ flags) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    flags[q] = (!elig[q] || overflow[q] || (!positive_only && out_cnt[q] < need)) ? 1 : 0;
}

// signature sort on the device: keys ~sig ascending (= signature descending; the radix sort is stable, so equal
// signatures keep query order exactly like the host stable_sort it replaces), values = query index
__global__ void k_sig_keys(const uint32_t* __restrict__ sig, int32_t nq, uint32_t* __restrict__ key, int32_t* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { key[i] = ~sig[i]; val[i] = i; }
}
__global__ void k_invert_perm(const int32_t* __restrict__ perm, int32_t nq, int32_t* __restrict__ inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) inv[perm[i]] = i;
}

__global__ void k_fill_offsets(int64_t* __restrict__ off, int32_t nq, int64_t stride) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= nq) off[q] = (int64_t)q * stride;
}

// all plans again from the current thresholds (after a cross-shard threshold exchange)
__global__ void __launch_bounds__(128) k_replan(const float* __restrict__ thr, int32_t nq, const int32_t* __restrict__ elig,
                                                PlanArgs pa) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= nq || !elig[q]) return;
    make_plan(pa, q, thr[q], lane);
}

// Cross-shard threshold exchange (doc-sharded callers, br_set_thr_exchange).  After the seeding the shards' thr[Q] are
// gathered and every shard takes the maximum; after a tile launch the shards' current K best fp32 scores of every query
// (the heads of the candidate lists k_tighten has just sorted) are gathered and thr[q] is raised to the K-th largest of the
// world x K values: K distinct docs of the whole corpus score at least that much, so it is a valid lower bound of the
// K-th best score of the WHOLE corpus - the bound a single index would have after the same fraction of its docs - and the
// plans follow.
__global__ void k_xchg_heads(const float* __restrict__ cand_h, const int32_t* __restrict__ cand_cnt, int cap, int32_t nq, int K,
                             float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nq * K) return;
    const int q = (int)(i / K), j = (int)(i - (int64_t)q * K);
    out[i] = j < min(cand_cnt[q], cap) ? cand_h[(int64_t)q * cap + j] : 0.f;
}
constexpr int XCHG_PER = 16;          // world x K <= 32 x XCHG_PER values per query
__global__ void __launch_bounds__(128) k_xchg_union(const float* __restrict__ gathered, int world, int32_t nq, int n_per,
                                                    int K, float* __restrict__ thr) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= nq) return;
    const int n = world * n_per;        // n_per = 1: the shards' thresholds (maximum); n_per = K: their K best scores
    float v[XCHG_PER];
#pragma unroll
    for (int j = 0; j < XCHG_PER; ++j) {
        const int i = lane + 32 * j;
        v[j] = -1.f;
        if (i < n) { const int r = i / n_per, c = i - r * n_per; v[j] = gathered[((int64_t)r * nq + q) * n_per + c]; }
    }
    const int rounds = n_per == 1 ? 1 : K;
    float kth = 0.f;
    for (int r = 0; r < rounds; ++r) {
        float m = -1.f;
#pragma unroll
        for (int j = 0; j < XCHG_PER; ++j) m = fmaxf(m, v[j]);
        float M = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
        const int winner = __ffs(__ballot_sync(0xffffffffu, m == M)) - 1;
        if (lane == winner) {
            bool removed = false;
#pragma unroll
            for (int j = 0; j < XCHG_PER; ++j)
                if (!removed && v[j] == M) { v[j] = -1.f; removed = true; }
        }
        kth = M;
    }
    if (lane == 0 && kth > thr[q]) thr[q] = kth;
}
static int exchange_thr(br_index* ix, float* thr, int32_t nq, const int32_t* elig, const PlanArgs& pa, cudaStream_t st,
                        const TileArgs* a, float* x_local, float* x_gathered) {
    const int world = ix->thr_exchange_world;
    const int n_per = a ? a->K : 1;
    const float* local = thr;
    if (a) {
        k_xchg_heads<<<blocks_for((int64_t)nq * a->K, 256), 256, 0, st>>>(a->cand_h, a->cand_cnt, a->cap, nq, a->K, x_local);
        BR_CUDA(cudaGetLastError());
        local = x_local;
    }
    const int rc = ix->thr_exchange(local, x_gathered, (int64_t)nq * n_per, (void*)st, ix->thr_exchange_user);
    BR_REQUIRE(rc == 0, BR_ERR_STATE, "br_topk_batch: the threshold exchange callback failed");
    k_xchg_union<<<blocks_for((int64_t)nq * 32, 128), 128, 0, st>>>(x_gathered, world, nq, n_per, a ? a->K : 1, thr);
    BR_CUDA(cudaGetLastError());
    k_replan<<<blocks_for((int64_t)nq * 32, 128), 128, 0, st>>>(thr, nq, elig, pa);
    BR_CUDA(cudaGetLastError());
    ix->stats.kernel_launches += a ? 3 : 2;
    return BR_OK;
}

// launches of the tile kernel for an index of n_tiles tiles and top-k (doubling chunks)
static int tile_launch_count(int n_tiles, int n_sub, int k, int growth) {
    const bool big = k > 32;
    const int unit = big ? 1 : TILE_W, n_units = big ? n_sub : n_tiles;
    int t0 = 0, chunk = big ? std::max(1, TILE_CAP_BIG / TILE_S) : std::max(1, std::min(TILE_CHUNK0, TILE_CAP / (TILE_W * 3 * k))), n = 0;
    while (t0 < n_units) { t0 += std::min(std::min(chunk, n_units - t0), 32768 * unit); chunk *= growth; ++n; }
    return n;
}

template <int G, int QT, int SPM>
static int launch_tiles_m(const TileArgs& a0, int n_groups, int n_tiles, size_t smem, cudaStream_t st, br_index* ix,
                        int32_t* prev_cnt, int32_t* overflow, const PlanArgs& pa, int exchange_rounds, float* x_local,
                        float* x_gathered) {
    BR_CUDA(cudaFuncSetAttribute(k_tile_score<G, QT, SPM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // per device
    // chunks of doubling size: thresholds learnt on the first 2^c tiles filter the next 2^c
    // first chunk: as many tiles as the candidate buffer certainly holds with thresholds still at zero
    // (every sub-range can emit up to ~3k docs before its first tightening)
    const bool big = a0.cap > TILE_CAP;
    if (big) BR_CUDA(cudaFuncSetAttribute(k_tighten_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * TILE_CAP_BIG));
    // k <= 32: whole tiles.  Large k: sub-ranges - before the first threshold exists everything passes, so the first launch
    // covers only as many docs as a candidate region holds (TILE_CAP_BIG / S sub-ranges, half a tile), then doubles.
    const int unit = big ? 1 : TILE_W;                  // sub-ranges per scheduling unit
    const int n_units = big ? a0.n_sub : n_tiles;
    int t0 = 0, chunk = big ? std::max(1, TILE_CAP_BIG / TILE_S) : std::max(1, std::min(TILE_CHUNK0, TILE_CAP / (TILE_W * 3 * a0.K))), round = 0;
    while (t0 < n_units) {
        const int nu = std::min(std::min(chunk, n_units - t0), 32768 * unit);
        TileArgs a = a0;
        a.sub_begin = t0 * unit;
        a.sub_end = std::min((t0 + nu) * unit, a0.n_sub);
        a.tile0 = a.sub_begin / TILE_W;
        a.tile_end = (a.sub_end + TILE_W - 1) / TILE_W;
        const int ny = a.tile_end - a.tile0;
        // consecutive tiles per CTA: amortises the CTA prologue in the large launches, keeps the small ones wide
        a.tpb = 1;
        while (a.tpb * 2 <= ix->tile_tpb && ny >= 4 * a.tpb) a.tpb *= 2;
        ix->prof_begin(st);
        k_tile_score<G, QT, SPM><<<dim3((unsigned)n_groups, (unsigned)((ny + a.tpb - 1) / a.tpb)), TILE_W * 32, smem, st>>>(a);
        BR_CUDA(cudaGetLastError());
        ix->prof_end(st);
        if (big) k_tighten_big<<<a0.nq, TG_T, 8 * (size_t)a0.cap, st>>>(a0.thr, a0.cand_cnt, prev_cnt, a0.cand, a0.cand_h, a0.K, a0.cap, a0.cap, overflow, pa);
        else k_tighten<<<a0.nq, TGS_T, 0, st>>>(a0.thr, a0.cand_cnt, prev_cnt, a0.cand, a0.cand_h, a0.K, overflow, pa);
        BR_CUDA(cudaGetLastError());
        ix->stats.kernel_launches += 2;
        t0 += nu;
        chunk *= ix->tile_growth;
        if (round < exchange_rounds) BR_TRY(exchange_thr(ix, a0.thr, a0.nq, a0.elig, pa, st, &a0, x_local, x_gathered));   // same count on every shard
        ++round;
    }
    return BR_OK;
}

template <int G, int QT>
static int launch_tiles(const TileArgs& a0, int n_groups, int n_tiles, size_t smem, cudaStream_t st, br_index* ix,
                        int32_t* prev_cnt, int32_t* overflow, const PlanArgs& pa, int exchange_rounds, float* x_local,
                        float* x_gathered) {
    switch (ix->sparse_mode) {
        case 0: return launch_tiles_m<G, QT, 0>(a0, n_groups, n_tiles, smem, st, ix, prev_cnt, overflow, pa, exchange_rounds, x_local, x_gathered);
        default: return launch_tiles_m<G, QT, 1>(a0, n_groups, n_tiles, smem, st, ix, prev_cnt, overflow, pa, exchange_rounds, x_local, x_gathered);
    }
}

int fused_launch_count(const br_index* ix, int32_t k) {
    if (!fused_supported(ix, k, 1)) return 0;
    return tile_launch_count((ix->n_sub + TILE_W - 1) / TILE_W, ix->n_sub, k, ix->tile_growth);
}

bool fused_supported(const br_index* ix, int32_t k, int32_t nq, bool cos) {
    const bool disabled = false, no_big = !ix->allow_fused_bigk;
    // 32 < k <= 1024: candidate regions of TILE_CAP_BIG slots per query (16 B each) - bounded to 6 GB of scratch
    const bool k_ok = k <= 32 || (!no_big && k <= 1024 && (int64_t)nq * TILE_CAP_BIG * 16 <= (6LL << 30));
    return !disabled && ix->allow_fused && k_ok && (cos || ix->variant != BR_OKAPI_NO_PLUS1) && ix->n_hot > 0 &&
           ix->sub_shift == TILE_SHIFT && ix->skip != nullptr;
}

// Fused path over the whole prepared batch.  h_flags[q] != 0 afterwards -> query q must be served by
// the dense path (not eligible, candidate overflow, or fewer than k docs with a positive score).
// long_pass: the second pass for queries with up to 64 terms / 40 hot terms (groups of at most 2, no deferral plans).
int topk_fused(br_index* ix, const int32_t* q_off, const PrepView& pv, int32_t nq, int32_t k, int dedup,
               int positive_only, int32_t* out_ids, double* out_scores, int32_t* out_counts, cudaStream_t st,
               std::vector<int32_t>* h_flags, bool long_pass, const br_posting* post_table) {
    // post_table != nullptr: the TF-IDF cosine stage (br_tfidf_cosine_topk) - same kernels over the weight table
    // tf*idf^2/||d|| (all weights >= 0 whatever the idf variant), no deferral (the look-up rows and upper bounds hold
    // BM25 weights), candidates re-scored by k_rescore_cos
    const bool cos = post_table != nullptr;
    const br_posting* const post = cos ? post_table : ix->post;
    const int n_srows = cos ? 0 : ix->n_srows;
    int G = nq >= 4 * kNumSMs ? 4 : (nq >= 2 * kNumSMs ? 2 : 1);
    if (ix->tile_g) G = ix->tile_g;
    if (G > TILE_GMAX) G = TILE_GMAX;
    if (long_pass && G > 2) G = 2;
    const int QT = long_pass ? TILE_QT_LONG : TILE_QT;
    const int n_groups = (nq + G - 1) / G, umax = G * QT;
    const int cap = k <= 32 ? TILE_CAP : TILE_CAP_BIG;
    const int n_tiles = (ix->n_sub + TILE_W - 1) / TILE_W;
    const int64_t n_buckets = (int64_t)n_tiles * n_groups;
    const int64_t n_scan_blocks = (n_buckets + SCAN_TILE - 1) / SCAN_TILE;
    size_t bytes = 0;
    auto carve = [&](size_t n) { size_t o = bytes; bytes += (n + 255) & ~(size_t)255; return o; };
    const size_t Q = (size_t)nq + 2;
    const size_t o_ent = carve(sizeof(TileEntry) * (size_t)n_groups * umax), o_ne = carve(4 * (size_t)n_groups),
                 o_el = carve(4 * Q), o_thr = carve(4 * Q), o_cnt = carve(4 * Q), o_prev = carve(4 * Q), o_ovf = carve(4 * Q),
                 o_dm = carve(4 * Q), o_nu = carve(4 * Q),
                 o_off = carve(8 * Q), o_fl = carve(4 * Q), o_oc = carve(4 * Q), o_ct = carve(16),
                 o_bc = carve(4 * (size_t)(n_buckets + 1)), o_bo = carve(4 * (size_t)(n_buckets + 1)),
                 o_cur = carve(4 * (size_t)(n_buckets + 1)), o_part = carve(4 * (size_t)(n_scan_blocks + 1)),
                 o_poff = carve(8 * (size_t)(n_scan_blocks + 2)),
                 o_cand = carve(4 * (size_t)nq * cap), o_ch = carve(4 * (size_t)nq * cap),
                 o_cs = carve(8 * (size_t)nq * cap), o_perm = carve(4 * Q), o_inv = carve(4 * Q), o_sk = carve(12 * Q),
                 o_xl = carve(4 * Q * (size_t)std::min<int32_t>(k, 32)), o_xg = carve(4 * Q * (size_t)std::min<int32_t>(k, 32) * (size_t)std::max(1, ix->thr_exchange_world));
    BR_TRY(ix->ws_tile.reserve(bytes));
    char* p = ix->ws_tile.as<char>();
    TileEntry* entries = (TileEntry*)(p + o_ent);
    int32_t* n_entries = (int32_t*)(p + o_ne);
    int32_t* elig = (int32_t*)(p + o_el);
    float* thr = (float*)(p + o_thr);
    int32_t* cand_cnt = (int32_t*)(p + o_cnt);
    int32_t* prev_cnt = (int32_t*)(p + o_prev);
    int32_t* overflow = (int32_t*)(p + o_ovf);
    uint32_t* defer_mask = (uint32_t*)(p + o_dm);
    float* ne_ub = (float*)(p + o_nu);
    int64_t* cand_off = (int64_t*)(p + o_off);
    int32_t* flags = (int32_t*)(p + o_fl);
    int32_t* cnt_tmp = out_counts ? out_counts : (int32_t*)(p + o_oc);
    unsigned long long* cold_total = (unsigned long long*)(p + o_ct);
    uint32_t* b_cnt = (uint32_t*)(p + o_bc);
    uint32_t* b_off = (uint32_t*)(p + o_bo);
    uint32_t* b_cur = (uint32_t*)(p + o_cur);
    uint32_t* part = (uint32_t*)(p + o_part);
    int64_t* part_off = (int64_t*)(p + o_poff);
    int32_t* cand = (int32_t*)(p + o_cand);
    float* cand_h = (float*)(p + o_ch);
    double* cand_sc = (double*)(p + o_cs);
    int32_t* perm = (int32_t*)(p + o_perm);
    int32_t* inv_perm = (int32_t*)(p + o_inv);
    float* x_local = (float*)(p + o_xl);
    float* x_gathered = (float*)(p + o_xg);

    // group queries that share the most frequent terms: sort by signature (bit 31 = most frequent term), on the device
    {
        uint32_t* key_in = (uint32_t*)(p + o_sk);
        uint32_t* key_out = key_in + Q;
        int32_t* val_in = (int32_t*)(key_out + Q);
        k_sig_keys<<<blocks_for(nq, 256), 256, 0, st>>>(pv.sig, nq, key_in, val_in);
        BR_CUDA(cudaGetLastError());
        size_t tmp_bytes = 0;
        BR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out, val_in, perm, nq, 0, 32, st));
        BR_TRY(ix->ws_sort.reserve(tmp_bytes + 16));
        BR_CUDA(cub::DeviceRadixSort::SortPairs(ix->ws_sort.p, tmp_bytes, key_in, key_out, val_in, perm, nq, 0, 32, st));
        k_invert_perm<<<blocks_for(nq, 256), 256, 0, st>>>(perm, nq, inv_perm);
        BR_CUDA(cudaGetLastError());
        ix->stats.kernel_launches += 2;
    }
    if (ix->ws_cold.cap == 0) BR_TRY(ix->ws_cold.reserve(sizeof(ColdEntry) * std::max<size_t>(1u << 20, (size_t)nq * 2048)));
    for (int attempt = 0; attempt < 2; ++attempt) {
    // one memset covers thr / cand_cnt / prev_cnt / overflow / defer_mask / ne_ub (contiguous carve)
    BR_CUDA(cudaMemsetAsync(p + o_thr, 0, o_off - o_thr, st));
    BR_CUDA(cudaMemsetAsync(cold_total, 0, 16, st));
    k_tile_prep<<<(n_groups + 3) / 4, 128, 0, st>>>(pv, q_off, ix->hot_slot, ix->row_ptr, nq, G, umax, entries, n_entries,
                                                    elig, cold_total, std::max<int64_t>(1, (ix->n_docs * (int64_t)ix->tile_dense_min) >> TILE_SHIFT), perm, ix->row_slot, n_srows, QT, long_pass ? 64 : 32);
    BR_CUDA(cudaGetLastError());
    k_fill_offsets<<<blocks_for(nq + 1, 256), 256, 0, st>>>(cand_off, nq, cap);
    BR_CUDA(cudaGetLastError());
    ix->stats.kernel_launches += 2;
    // Cold postings -> (tile, group) buckets.  Their number is known on the device only; reading it back here would
    // stall the stream in the middle of the batch, so the scatter writes into the buffer kept from earlier batches
    // (bounded by its capacity) and the total is checked after the batch: a batch that needed more is repeated once
    // with a larger buffer.
    const ColdEntry* cold = nullptr;
    const uint32_t* cold_off = nullptr;
    const uint32_t cold_cap = (uint32_t)std::min<size_t>(ix->ws_cold.cap / sizeof(ColdEntry), 0xffffffffu);
    {
        ColdEntry* d_cold = ix->ws_cold.as<ColdEntry>();
        BR_CUDA(cudaMemsetAsync(b_cnt, 0, 4 * (size_t)(n_buckets + 1), st));
        const unsigned qb = blocks_for((int64_t)nq * 32, 128);
        k_cold_pass<false><<<qb, 128, 0, st>>>(pv, q_off, ix->hot_slot, ix->row_ptr, post, nq, G, n_groups, elig, inv_perm,
                                              b_cnt, nullptr, 0u);
        BR_CUDA(cudaGetLastError());
        k_scan_reduce<<<(unsigned)n_scan_blocks, SCAN_T, 0, st>>>(b_cnt, n_buckets, part);
        BR_CUDA(cudaGetLastError());
        k_exscan<uint32_t><<<1, 1024, 0, st>>>(part, n_scan_blocks, part_off);
        BR_CUDA(cudaGetLastError());
        k_scan_apply<<<(unsigned)n_scan_blocks, SCAN_T, 0, st>>>(b_cnt, n_buckets, part_off, b_off, b_cur);
        BR_CUDA(cudaGetLastError());
        k_cold_pass<true><<<qb, 128, 0, st>>>(pv, q_off, ix->hot_slot, ix->row_ptr, post, nq, G, n_groups, elig, inv_perm,
                                             b_cur, d_cold, cold_cap);
        BR_CUDA(cudaGetLastError());
        ix->stats.kernel_launches += 5;
        cold = d_cold;
        cold_off = b_off;
    }
    TileArgs a{post, ix->row_ptr, ix->skip, ix->n_sub, ix->n_docs, entries, n_entries, umax, nq, n_groups, elig,
               cold_off, cold, cold_cap, ix->dense_rows, ix->n_pad, thr, cand_cnt, cand, cand_h, cap, (int)k, 0, 0, 0, 0, 1, dedup ? 0 : 1, perm, defer_mask, ne_ub, ix->row_slot};
    const PlanArgs pa{q_off, pv.u_terms, pv.u_mult, pv.u_cnt, ix->ub, ix->row_slot, n_srows, cos ? 0.f : (float)ix->defer_pm * 1e-3f, defer_mask, ne_ub};
    const size_t smem = sizeof(float) * G * TILE_W * TILE_S + (size_t)umax * (sizeof(TileEntry) + sizeof(int64_t) +
                                                                            sizeof(float4) + sizeof(uint32_t) + sizeof(uint2) * TILE_W) +
                        sizeof(NeEntry) * G * NE_MAX + sizeof(uint16_t) * TILE_W * LIST_CAP + 4 * (20 + 9 * TILE_W) + G * QT;
    // threshold exchange with the other shards: only in the regular pass of the first attempt, so that every shard makes
    // exactly 1 + thr_exchange_rounds calls per batch
    const bool exchange = !cos && ix->thr_exchange != nullptr && ix->thr_exchange_rounds >= 0 && !long_pass && attempt == 0 && k <= 32 &&
                          ix->thr_exchange_world * (int)k <= 32 * XCHG_PER;
    const int xr = exchange ? ix->thr_exchange_rounds : 0;
    if (k <= 32 && ix->seed_thr) {
        k_seed_thr<<<blocks_for((int64_t)nq * 32, 128), 128, 0, st>>>(post, ix->row_ptr, nq, (int)k, elig, thr, pa);
        BR_CUDA(cudaGetLastError());
        ix->stats.kernel_launches += 1;
    }
    if (exchange) BR_TRY(exchange_thr(ix, thr, nq, elig, pa, st, nullptr, x_local, x_gathered));
    if (long_pass) {
        if (G == 1) BR_TRY((launch_tiles<1, TILE_QT_LONG>(a, n_groups, n_tiles, smem, st, ix, prev_cnt, overflow, pa, xr, x_local, x_gathered)));
        else BR_TRY((launch_tiles<2, TILE_QT_LONG>(a, n_groups, n_tiles, smem, st, ix, prev_cnt, overflow, pa, xr, x_local, x_gathered)));
    } else {
        switch (G) {
            case 1: BR_TRY((launch_tiles<1, TILE_QT>(a, n_groups, n_tiles, smem, st, ix, prev_cnt, overflow, pa, xr, x_local, x_gathered))); break;
            case 2: BR_TRY((launch_tiles<2, TILE_QT>(a, n_groups, n_tiles, smem, st, ix, prev_cnt, overflow, pa, xr, x_local, x_gathered))); break;
            default: BR_TRY((launch_tiles<TILE_GMAX, TILE_QT>(a, n_groups, n_tiles, smem, st, ix, prev_cnt, overflow, pa, xr, x_local, x_gathered))); break;
        }
    }
    BR_TRY(launch_rescore_heads(ix, q_off, pv, dedup, nq, cap, cand_cnt, cand, cand_sc, st, cos));
    BR_TRY(launch_final_select(cand, cand_sc, cand_off, 0, nq, k, positive_only, out_ids, out_scores, cnt_tmp, st, cand_cnt));
    // with shared thresholds a shard legitimately returns fewer than k docs (the rest of the corpus holds better ones)
    const int32_t need = exchange ? 0 : (int32_t)std::min<int64_t>(k, ix->n_docs);
    k_fused_flags<<<blocks_for(nq, 256), 256, 0, st>>>(elig, overflow, cnt_tmp, nq, need, positive_only, flags);
    BR_CUDA(cudaGetLastError());
    ix->stats.kernel_launches += 3;
    h_flags->resize((size_t)nq);
    unsigned long long h_cold = 0;
    BR_CUDA(cudaMemcpyAsync(h_flags->data(), flags, 4 * (size_t)nq, cudaMemcpyDeviceToHost, st));
    BR_CUDA(cudaMemcpyAsync(&h_cold, cold_total, sizeof(h_cold), cudaMemcpyDeviceToHost, st));
    BR_CUDA(cudaStreamSynchronize(st));                    // the only host synchronisation of the fused path
    BR_REQUIRE(h_cold < (1ull << 32), BR_ERR_UNSUPPORTED, "br_topk_batch: more than 2^32 cold postings in one batch");
    if (h_cold > cold_cap) {                               // the cold buffer was too small: grow it and repeat the batch
        BR_REQUIRE(attempt == 0, BR_ERR_CUDA, "br_topk_batch: cold buffer still too small after growing it");
        BR_TRY(ix->ws_cold.reserve(sizeof(ColdEntry) * (size_t)h_cold + 256));
        continue;
    }
    break;
    }   // attempt
    return BR_OK;
}

}  // namespace br
