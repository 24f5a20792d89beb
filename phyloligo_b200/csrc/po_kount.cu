// Kount.py's sliding-window stage on the device (sm_100a): every window of every contig is a
// virtual record of po_profile_batch; this file adds the two small per-window kernels around it.
//
//   * window_count_byte_kernel: occurrences of one byte value (the upper-case 'N' of the
//     reference's  seq.count('N') / len(seq) <= n_max_freq_in_windows  filter,
//     reference bin/Kount.py:294) in every window;
//   * window_distance_kernel: distance of every window profile to ONE reference profile, the
//     1-D forms of the reference: KL (bin/Kount.py:71-86), Eucl (:88-92, x1000),
//     JSD (:94-123, x1000), as dispatched by compute_distance_joblib (:317-324).  float64 like
//     the reference; NaN / Inf terms are zeroed term by term (posdef_check_value, :67-69).
// One warp per window, lanes stride over the bytes / dimensions, fixed-order shuffle reduction
// (deterministic).
#include <math.h>
#include "po_common.cuh"

namespace po {

__global__ void __launch_bounds__(256) window_count_byte_kernel(const uint8_t* __restrict__ text,
                                                                const int64_t* __restrict__ begin,
                                                                const int64_t* __restrict__ end, int64_t n,
                                                                unsigned value, int64_t* __restrict__ out) {
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int64_t b = begin[w], e = end[w];
    unsigned c = 0;
    for (int64_t i = b + lane; i < e; i += 32) c += (text[i] == value);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (lane == 0) out[w] = (int64_t)c;
}

__device__ __forceinline__ double scrub(double v) { return (isnan(v) || isinf(v)) ? 0.0 : v; }

template <int METRIC>  // 0 JSD, 1 KL, 2 Eucl
__global__ void __launch_bounds__(256) window_distance_kernel(const double* __restrict__ F, int64_t n, int64_t dim,
                                                              int64_t ld, const double* __restrict__ ref,
                                                              double* __restrict__ out) {
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const double* a = F + w * ld;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t k = lane; k < dim; k += 32) {
        const double x = a[k], y = ref[k];
        if (METRIC == 0) {
            const double h = 0.5 * (x + y);
            s1 += scrub(x * log(x / h));
            s2 += scrub(y * log(y / h));
        } else if (METRIC == 1) {
            s1 += scrub(x * log(x / y));
        } else {
            const double d = x - y;
            s1 += scrub(d * d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
        s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    }
    if (lane == 0) {
        if (METRIC == 0) out[w] = 0.5 * (s1 + s2) * 1000.0;
        else if (METRIC == 1) out[w] = s1;
        else out[w] = sqrt(s1) * 1000.0;
    }
}

}  // namespace po

using namespace po;

extern "C" int po_window_count_byte(const uint8_t* d_text, const int64_t* d_begin, const int64_t* d_end, int64_t n,
                                    int value, int64_t* d_counts, po_stream_t stream) {
    if (n < 0 || value < 0 || value > 255 || (n > 0 && (!d_text || !d_begin || !d_end || !d_counts))) {
        set_error("po_window_count_byte: bad arguments");
        return PO_ERR_ARG;
    }
    if (n == 0) return PO_OK;
    window_count_byte_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(d_text, d_begin, d_end, n,
                                                                                       (unsigned)value, d_counts);
    count_launch(0);
    PO_LAUNCH_CHECK("window_count_byte_kernel");
    return PO_OK;
}

extern "C" int po_window_distances(int metric, const double* d_freq, int64_t n, int64_t dim, int64_t ld,
                                   const double* d_ref, double* d_out, po_stream_t stream) {
    if (metric < 0 || metric > 2 || n < 0 || dim < 1 || ld < dim || (n > 0 && (!d_freq || !d_ref || !d_out))) {
        set_error("po_window_distances: bad arguments (metric %d, n %lld, dim %lld)", metric, (long long)n, (long long)dim);
        return PO_ERR_ARG;
    }
    if (n == 0) return PO_OK;
    const unsigned grid = (unsigned)((n + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (metric == 0) window_distance_kernel<0><<<grid, 256, 0, st>>>(d_freq, n, dim, ld, d_ref, d_out);
    else if (metric == 1) window_distance_kernel<1><<<grid, 256, 0, st>>>(d_freq, n, dim, ld, d_ref, d_out);
    else window_distance_kernel<2><<<grid, 256, 0, st>>>(d_freq, n, dim, ld, d_ref, d_out);
    count_launch(1);
    PO_LAUNCH_CHECK("window_distance_kernel");
    return PO_OK;
}
