// Composition profiling: spaced-word counting of FASTA records on sm_100a.
//
// Replaces select_strand + .upper() + cut_sequence_and_count_pattern + count2freq
// (reference bin/phyloligo.py:124-149, 683, 601-631, 633-661).
//
// Layout / algorithm
//   * the raw FASTA text sits in HBM once; a record is a byte range whose
//     '\n', '\r', ' ' bytes are transparent (Biopython strips them).
//   * one CTA per record.  Each thread walks 64-byte, 16-byte-aligned chunks
//     (4 x 128-bit coalesced loads), keeping the last `width` bases in rolling
//     registers:  fw  = 2-bit codes, first base of the window in the top field
//                 rc  = the same window reverse-complemented (code ^ 1, reversed)
//                 inv = one bit per base, set when the base is not ACGT
//     A window is counted iff none of its `width` bases is invalid -- exactly the
//     re.split('[^ACGT]+') + len(run) >= len(pattern) rule.
//   * the word code is gathered from the runs of '1' of the pattern and counted
//     in a shared-memory histogram (C=0,G=1,A=2,T=3 so complement = code ^ 1 and
//     the bin order is the reference's product(("C","G","A","T")) order).
//   * strand "both" in the reference is seq + revcomp(seq) with NO separator, so
//     up to width-1 chimeric windows straddle the junction; they are replayed
//     from the last width-1 bases of the record.
//   * the CTA then writes counts, total and count/total (IEEE float64 divide =
//     Python's int/int true division for operands < 2^53; float32 is the cast of
//     that quotient, as the reference's memmap/h5py paths do).
#pragma once
#include <type_traits>
#include "po_common.cuh"

namespace po {

constexpr int PROFILE_THREADS = 128;
constexpr int CHUNK = 64;  // bytes per thread step

__device__ __forceinline__ bool is_skip(uint32_t c) { return c == 10u || c == 13u || c == 32u; }

// returns code in bits 0-1 and "invalid" in bit 2
__device__ __forceinline__ uint32_t classify(uint32_t c) {
    uint32_t up = c & 0xDFu;
    uint32_t x = (up >> 1) & 3u;                // A=0 C=1 T=2 G=3
    uint32_t code = (0x72u >> (2u * x)) & 3u;   // -> A=2 C=0 T=3 G=1
    bool ok = (up == 0x41u) | (up == 0x43u) | (up == 0x47u) | (up == 0x54u);
    return code | (ok ? 0u : 4u);
}

template <bool WIDE>
struct Window {
    using reg_t = typename std::conditional<WIDE, uint64_t, uint32_t>::type;
    reg_t fw, rc;
    uint32_t inv;
    __device__ __forceinline__ void reset() { fw = 0; rc = 0; inv = 0xFFFFFFFFu; }
    __device__ __forceinline__ void push(uint32_t cls, int top_shift) {
        uint32_t code = cls & 3u;
        fw = (fw << 2) | (reg_t)code;
        rc = (rc >> 2) | ((reg_t)(code ^ 1u) << top_shift);
        inv = (inv << 1) | (cls >> 2);
    }
};

template <bool WIDE, int NRUNS>
__device__ __forceinline__ uint32_t gather(typename Window<WIDE>::reg_t r, const PatternGeom& g) {
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < NRUNS; ++i) {
        if (i < g.nruns) w |= ((uint32_t)(r >> g.shift[i]) & g.mask[i]) << g.dst[i];
    }
    return w;
}

template <bool WIDE, int NRUNS>
__device__ __forceinline__ void emit(const Window<WIDE>& win, const PatternGeom& g, uint32_t pmask,
                                     int strand, uint32_t* hist) {
    if ((win.inv & pmask) == 0u) {
        if (strand != PO_STRAND_MINUS) atomicAdd(&hist[gather<WIDE, NRUNS>(win.fw, g)], 1u);
        if (strand != PO_STRAND_PLUS) atomicAdd(&hist[gather<WIDE, NRUNS>(win.rc, g)], 1u);
    }
}

template <bool WIDE, int NRUNS, bool GLOBAL_HIST>
__global__ void __launch_bounds__(PROFILE_THREADS)
profile_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ rec_begin,
               const int64_t* __restrict__ rec_end, const PatternGeom g, int strand, int64_t dim,
               uint32_t* __restrict__ counts, uint64_t* __restrict__ totals,
               double* __restrict__ freq64, float* __restrict__ freq32) {
    extern __shared__ uint32_t smem_hist[];
    __shared__ unsigned long long s_total;
    const int64_t rec = blockIdx.x;
    const int64_t begin = rec_begin[rec];
    const int64_t end = rec_end[rec];
    const int tid = threadIdx.x;
    uint32_t* hist = GLOBAL_HIST ? (counts + rec * dim) : smem_hist;

    if (!GLOBAL_HIST) {
        for (int64_t b = tid; b < dim; b += PROFILE_THREADS) hist[b] = 0u;
    }
    if (tid == 0) s_total = 0ull;
    __syncthreads();

    const int P = g.width;
    const uint32_t pmask = (P >= 32) ? 0xFFFFFFFFu : ((1u << P) - 1u);
    const int top_shift = 2 * (P - 1);

    if (end > begin && P >= 1) {
        const int64_t base = begin & ~(int64_t)15;
        const int64_t nchunks = (end - base + CHUNK - 1) / CHUNK;
        for (int64_t ch = tid; ch < nchunks; ch += PROFILE_THREADS) {
            const int64_t cstart = base + ch * CHUNK;
            Window<WIDE> win;
            win.reset();
            // warm-up: replay the width-1 bases that precede this chunk
            if (cstart > begin && P > 1) {
                int need = P - 1;
                int64_t q = cstart;
                while (q > begin && need > 0) {
                    --q;
                    if (!is_skip(__ldg(text + q))) --need;
                }
                for (; q < cstart; ++q) {
                    uint32_t c = __ldg(text + q);
                    if (!is_skip(c)) win.push(classify(c), top_shift);
                }
            }
            const int lo = (int)max((int64_t)0, begin - cstart);
            const int hi = (int)min((int64_t)CHUNK, end - cstart);
            const uint4* src = reinterpret_cast<const uint4*>(text + cstart);
            uint4 v[CHUNK / 16];
#pragma unroll
            for (int i = 0; i < CHUNK / 16; ++i) {
                // only touch 16-byte words that intersect the record
                if (i * 16 < hi && i * 16 + 16 > lo) v[i] = __ldg(src + i);
                else v[i] = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
            }
            const bool full = (lo == 0) && (hi == CHUNK);
#pragma unroll
            for (int i = 0; i < CHUNK / 16; ++i) {
                const uint32_t wds[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int idx = i * 16 + j * 4 + b;
                        const uint32_t c = (wds[j] >> (8 * b)) & 0xFFu;
                        const bool inside = full || (idx >= lo && idx < hi);
                        if (inside && !is_skip(c)) {
                            win.push(classify(c), top_shift);
                            emit<WIDE, NRUNS>(win, g, pmask, strand, hist);
                        }
                    }
                }
            }
        }
        // junction windows of seq + revcomp(seq)  (bin/phyloligo.py:141)
        if (strand == PO_STRAND_BOTH && tid == 0 && P > 1) {
            uint8_t tail[PO_MAX_PATTERN];
            int m = 0;
            int64_t q = end;
            while (q > begin && m < P - 1) {
                --q;
                uint32_t c = __ldg(text + q);
                if (!is_skip(c)) tail[m++] = (uint8_t)classify(c);  // tail[0] = last base
            }
            if (2 * m >= P) {
                Window<WIDE> win;
                win.reset();
                for (int t = m - 1; t >= 0; --t) win.push(tail[t], top_shift);  // forward order
                for (int t = 0; t < m; ++t) {
                    // revcomp(tail): last base first, complemented; validity unchanged
                    uint32_t cls = tail[t];
                    win.push((cls & 4u) | ((cls & 3u) ^ 1u), top_shift);
                    if ((win.inv & pmask) == 0u) atomicAdd(&hist[gather<WIDE, NRUNS>(win.fw, g)], 1u);
                }
            }
        }
    } else if (P == 0 && end > begin) {
        // pattern without any position: every placement of the empty window counts
        // (len(subseq) >= 0 always holds); handled on the host side -- not reachable.
    }
    __syncthreads();

    // total number of words and the outputs
    unsigned long long part = 0ull;
    for (int64_t b = tid; b < dim; b += PROFILE_THREADS) part += hist[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if ((tid & 31) == 0 && part) atomicAdd(&s_total, part);
    __syncthreads();
    const unsigned long long total = s_total;
    if (tid == 0 && totals) totals[rec] = total;
    const double dt = (double)total;
    for (int64_t b = tid; b < dim; b += PROFILE_THREADS) {
        const uint32_t c = hist[b];
        if (!GLOBAL_HIST && counts) counts[rec * dim + b] = c;
        const double f = total ? (double)c / dt : 0.0;
        if (freq64) freq64[rec * dim + b] = f;
        if (freq32) freq32[rec * dim + b] = (float)f;
    }
}

template <bool WIDE, int NRUNS>
static int launch_profile_t(const uint8_t* d_text, const int64_t* d_begin, const int64_t* d_end, int64_t n,
                            const PatternGeom& g, int strand, int64_t dim, uint32_t* d_counts,
                            uint64_t* d_totals, double* d_freq64, float* d_freq32, cudaStream_t stream) {
    const size_t hist_bytes = (size_t)dim * sizeof(uint32_t);
    const bool global_hist = hist_bytes > 160 * 1024;
    if (global_hist) {
        if (!d_counts) {
            set_error("patterns with more than 7 ones need d_counts (global histogram)");
            return PO_ERR_UNSUPPORTED;
        }
        PO_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, (size_t)n * hist_bytes, stream));
        LaunchTimer t(0, stream);
        profile_kernel<WIDE, NRUNS, true><<<(unsigned)n, PROFILE_THREADS, 0, stream>>>(
            d_text, d_begin, d_end, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32);
        count_launch(0);
    } else {
        auto kern = profile_kernel<WIDE, NRUNS, false>;
        if (hist_bytes > 48 * 1024) {
            PO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
        }
        LaunchTimer t(0, stream);
        kern<<<(unsigned)n, PROFILE_THREADS, hist_bytes, stream>>>(
            d_text, d_begin, d_end, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32);
        count_launch(0);
    }
    PO_LAUNCH_CHECK("profile_kernel");
    return PO_OK;
}

template <bool WIDE>
static int launch_profile_width(const uint8_t* d_text, const int64_t* d_begin, const int64_t* d_end, int64_t n,
                                const PatternGeom& g, int strand, uint32_t* d_counts, uint64_t* d_totals,
                                double* d_freq64, float* d_freq32, cudaStream_t stream) {
    const int64_t dim = (int64_t)1 << (2 * g.k);
    if (g.nruns <= 1)
        return launch_profile_t<WIDE, 1>(d_text, d_begin, d_end, n, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32, stream);
    if (g.nruns <= 4)
        return launch_profile_t<WIDE, 4>(d_text, d_begin, d_end, n, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32, stream);
    return launch_profile_t<WIDE, 16>(d_text, d_begin, d_end, n, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32, stream);
}

}  // namespace po
