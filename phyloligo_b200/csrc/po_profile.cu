// Dispatcher of the profiling kernels (see po_profile_kernel.cuh for the algorithm).
#include <stdlib.h>
#include "po_common.cuh"

namespace po {

int launch_profile_narrow(const uint8_t*, const int64_t*, const int64_t*, int64_t, const PatternGeom&, int,
                          uint32_t*, uint64_t*, double*, float*, cudaStream_t);
int launch_profile_wide(const uint8_t*, const int64_t*, const int64_t*, int64_t, const PatternGeom&, int,
                        uint32_t*, uint64_t*, double*, float*, cudaStream_t);

bool profile_seg_supported(const PatternGeom& g, int strand, int64_t dim);
int launch_profile_seg(const uint8_t*, const int64_t*, const int64_t*, int64_t, const PatternGeom&, int, uint32_t*,
                       uint64_t*, double*, float*, cudaStream_t);

int launch_profile(const uint8_t* d_text, const int64_t* d_begin, const int64_t* d_end, int64_t n,
                   const PatternGeom& g, int strand, uint32_t* d_counts, uint64_t* d_totals,
                   double* d_freq64, float* d_freq32, cudaStream_t stream) {
    if (n == 0) return PO_OK;
    if (n > 0x7FFFFFFFll) {
        set_error("too many records in one batch (%lld)", (long long)n);
        return PO_ERR_UNSUPPORTED;
    }
    // PO_PROFILE_KERNEL=general forces the general kernel (the parity tests cover both)
    const char* force = getenv("PO_PROFILE_KERNEL");
    const bool general_only = force && !strcmp(force, "general");
    if (!general_only && profile_seg_supported(g, strand, (int64_t)1 << (2 * g.k)))
        return launch_profile_seg(d_text, d_begin, d_end, n, g, strand, d_counts, d_totals, d_freq64, d_freq32, stream);
    if (g.width > 16)
        return launch_profile_wide(d_text, d_begin, d_end, n, g, strand, d_counts, d_totals, d_freq64, d_freq32, stream);
    return launch_profile_narrow(d_text, d_begin, d_end, n, g, strand, d_counts, d_totals, d_freq64, d_freq32, stream);
}

}  // namespace po
