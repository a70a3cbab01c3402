// br_kernels.cuh - device helpers shared by the build and query kernels.
#pragma once
#include "br_common.cuh"

namespace br {

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (blockDim.x <= 1024); returns exclusive
// prefix, writes the block total to *total (valid for all threads after the call).
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nw ? s_warp[lane] : 0;
        const uint32_t wi = warp_incl_scan(w, lane);
        s_warp[lane] = wi - w;
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    const uint32_t r = s_warp[wid] + incl - v;
    *total = s_total;
    __syncthreads();
    return r;
}

// single-block exclusive scan: out[i] = sum_{j<i} in[j], out[n] = total.
template <class TIn>
__global__ void __launch_bounds__(1024) k_exscan(const TIn* __restrict__ in, int64_t n,
                                                 int64_t* __restrict__ out) {
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    constexpr int I = 4;
    for (int64_t base = 0; base < n; base += 1024 * I) {
        uint32_t v[I];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < I; ++j) {
            const int64_t i = base + (int64_t)threadIdx.x * I + j;
            v[j] = i < n ? (uint32_t)in[i] : 0u;
            sum += v[j];
        }
        uint32_t total;
        uint32_t ex = block_excl_scan(sum, &total);
        unsigned long long run = s_carry + ex;
#pragma unroll
        for (int j = 0; j < I; ++j) {
            const int64_t i = base + (int64_t)threadIdx.x * I + j;
            if (i < n) out[i] = (int64_t)run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = (int64_t)s_carry;
}

// The reference formula, float64, one rounding per operation (no FMA contraction):
//   idf * ((tf * (k1 + 1)) / (tf + k1 * (1 - b + dl / avgdl)))            bm25_ranking.ipynb:202
//   idf * ((tf * (k1 + 1)) / (tf + k1 * (1 - b + b * dl / avgdl)))        team_run1.py:193
__device__ __forceinline__ double bm25_contrib(double idf, double tf, double dl, double avgdl, double k1,
                                               double b, int variant) {
    const double r = variant == BR_NOTEBOOK ? __ddiv_rn(dl, avgdl) : __ddiv_rn(__dmul_rn(b, dl), avgdl);
    const double norm = __dadd_rn(__dsub_rn(1.0, b), r);
    const double den = __dadd_rn(tf, __dmul_rn(k1, norm));
    const double num = __dmul_rn(tf, __dadd_rn(k1, 1.0));
    return __dmul_rn(idf, __ddiv_rn(num, den));
}


static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace br
