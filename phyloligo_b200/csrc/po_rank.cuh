// Rank transform of one profile by one CTA (the scipy.stats.spearmanr / rankdata step of
// phylodist.SC, reference core/phylodist.py:82-85), shared by the CUDA-core and the
// tensor-core Spearman paths.
//
// The row is sorted in shared memory as (value, position) pairs with a bitonic network
// (dim padded to a power of two with +inf), then every sorted entry finds its tie group with
// two binary searches: less = entries strictly smaller, eq = entries equal.  The centred
// doubled average rank is the integer 2 * rank - (dim + 1) = 2 * less + eq - dim.
// O(dim log^2 dim) per row instead of the O(dim^2) all-against-all count.
#pragma once
#include <math_constants.h>
#include "po_common.cuh"

namespace po {

inline int64_t rank_pad(int64_t dim) {
    int64_t p = 2;
    while (p < dim) p <<= 1;
    return p;
}
// shared memory of rank_transform_row: keys (double) + positions (int)
inline size_t rank_smem_bytes(int64_t dim) { return (size_t)rank_pad(dim) * 12; }

// Calls emit(position, value) once for every position < dim, and returns (to every thread) this
// thread's share of sum(value^2).  `smem` holds rank_smem_bytes(dim) bytes, 8-byte aligned.
template <typename T, typename EMIT>
__device__ __forceinline__ unsigned long long rank_transform_row(const T* __restrict__ xrow, int dim, int dpad,
                                                                 unsigned char* smem, EMIT emit) {
    double* key = reinterpret_cast<double*>(smem);
    int* pos = reinterpret_cast<int*>(smem + (size_t)dpad * 8);
    for (int e = threadIdx.x; e < dpad; e += blockDim.x) {
        key[e] = e < dim ? (double)xrow[e] : CUDART_INF;
        pos[e] = e;
    }
    __syncthreads();
    for (int k = 2; k <= dpad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < dpad; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const double a = key[i], b = key[l];
                    const bool up = (i & k) == 0;
                    if (up ? (a > b) : (a < b)) {
                        key[i] = b;
                        key[l] = a;
                        const int t = pos[i];
                        pos[i] = pos[l];
                        pos[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    unsigned long long ss = 0ull;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        const double v = key[i];
        int lo = 0, hi = i;  // lower bound: first entry not smaller than v (it is <= i)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (key[mid] < v) lo = mid + 1;
            else hi = mid;
        }
        const int less = lo;
        lo = i + 1;
        hi = dim;            // upper bound: first entry greater than v (it is > i)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (key[mid] > v) hi = mid;
            else lo = mid + 1;
        }
        const int eq = lo - less;
        const int val = 2 * less + eq - dim;
        ss += (unsigned long long)((long long)val * (long long)val);
        emit(pos[i], val);
    }
    return ss;
}

}  // namespace po
