// Profiling kernels for patterns of width <= 16 (32-bit rolling window registers).
#include "po_profile_kernel.cuh"
namespace po {
int launch_profile_narrow(const uint8_t* d_text, const int64_t* d_begin, const int64_t* d_end, int64_t n,
                          const PatternGeom& g, int strand, uint32_t* d_counts, uint64_t* d_totals,
                          double* d_freq64, float* d_freq32, cudaStream_t stream) {
    return launch_profile_width<false>(d_text, d_begin, d_end, n, g, strand, d_counts, d_totals, d_freq64, d_freq32, stream);
}
}  // namespace po
