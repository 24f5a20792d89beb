// Composition profiling: spaced-word counting of FASTA records on sm_100a.
//
// Replaces select_strand + .upper() + cut_sequence_and_count_pattern + count2freq
// (reference bin/phyloligo.py:124-149, 683, 601-631, 633-661).
//
// Layout / algorithm
//   * the raw FASTA text sits in HBM once; a record is a byte range whose
//     '\n', '\r', ' ' and the other ASCII blanks (TAB, VT, FF) are transparent: Biopython drops blanks and CR
//     anywhere and strips every kind of trailing white space from a line.
//   * one CTA per record.  Each thread walks 64-byte, 16-byte-aligned chunks with
//     128-bit coalesced loads.  Every byte is classified through a 256-entry table in
//     shared memory (2-bit code, "not ACGT" flag, "skip" flag; C=0,G=1,A=2,T=3 so the
//     complement is code ^ 1 and the bin order is the reference's
//     product(("C","G","A","T")) order) and pushed into rolling registers:
//         fw  = 2-bit codes of the last `width` bases, first base in the top field
//         rc  = the same window reverse-complemented (only when it is needed)
//         inv = one bit per base, set when the base is not ACGT
//     A window is counted iff none of its `width` bases is invalid -- exactly the
//     re.split('[^ACGT]+') + len(run) >= len(pattern) rule.  A thread first replays
//     the width-1 bases that precede its chunk.
//   * the word code is gathered from the runs of '1' of the pattern and counted with
//     shared-memory atomics; small histograms (<= 1024 bins) are replicated per warp.
//   * palindromic patterns (every contiguous k-mer): the minus-strand count of word w
//     equals the plus-strand count of revcomp(w), so only plus-strand words are
//     counted and the histogram is folded (both) or permuted (minus) at the end --
//     one atomic per window instead of two.
//   * strand "both" in the reference is seq + revcomp(seq) with NO separator, so up
//     to width-1 chimeric windows straddle the junction; they are replayed from the
//     last width-1 bases of the record, after the fold.
//   * the CTA then writes counts, total and count/total (IEEE float64 divide =
//     Python's int/int true division for operands < 2^53; float32 is the cast of
//     that quotient, as the reference's memmap/h5py paths do).
#pragma once
#include <type_traits>
#include "po_common.cuh"

namespace po {

constexpr int PROFILE_THREADS = 128;
constexpr int PROFILE_WARPS = PROFILE_THREADS / 32;
constexpr int CHUNK = 64;  // bytes per thread step

// class byte: bits 0-1 code, bit 2 invalid (not ACGT after upper-casing), bit 3 skip
__device__ __forceinline__ uint32_t classify_byte(uint32_t c) {
    if (c == 32u || (c >= 9u && c <= 13u)) return 8u;  // space, TAB, LF, VT, FF, CR
    const uint32_t up = c & 0xDFu;
    if (up == 0x43u) return 0u;  // C
    if (up == 0x47u) return 1u;  // G
    if (up == 0x41u) return 2u;  // A
    if (up == 0x54u) return 3u;  // T
    return 4u;
}

template <bool WIDE>
struct Window {
    using reg_t = typename std::conditional<WIDE, uint64_t, uint32_t>::type;
    reg_t fw, rc;
    uint32_t inv;
    __device__ __forceinline__ void reset() { fw = 0; rc = 0; inv = 0xFFFFFFFFu; }
    template <bool NEED_RC>
    __device__ __forceinline__ void push(uint32_t cls, int top_shift) {
        const uint32_t code = cls & 3u;
        fw = (fw << 2) | (reg_t)code;
        if (NEED_RC) rc = (rc >> 2) | ((reg_t)(code ^ 1u) << top_shift);
        inv = (inv << 1) | ((cls >> 2) & 1u);
    }
};

template <bool WIDE, int NRUNS>
__device__ __forceinline__ uint32_t gather(typename Window<WIDE>::reg_t r, const PatternGeom& g) {
    if (NRUNS == 1) return (uint32_t)(r >> g.shift[0]) & g.mask[0];  // contiguous k-mer (dst = 0)
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < NRUNS; ++i) {
        if (i < g.nruns) w |= ((uint32_t)(r >> g.shift[i]) & g.mask[i]) << g.dst[i];
    }
    return w;
}

// digits reversed and complemented: the bin of the reverse-complement word
__device__ __forceinline__ uint32_t revcomp_bin(uint32_t w, int k) {
    if (k == 0) return 0u;
    uint32_t r = __brev(w) >> (32 - 2 * k);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    const uint32_t m = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    return (r ^ 0x55555555u) & m;
}

// MODE 0: count plus-strand words only (strand plus, or a folded palindromic pattern)
// MODE 1: count minus-strand words only
// MODE 2: count both words of every window
template <bool WIDE, int NRUNS, bool GLOBAL_HIST>
__global__ void __launch_bounds__(PROFILE_THREADS)
profile_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ rec_begin,
               const int64_t* __restrict__ rec_end, const PatternGeom g, int strand, int mode, int fold,
               int ncopy, int64_t dim, uint32_t* __restrict__ counts, uint64_t* __restrict__ totals,
               double* __restrict__ freq64, float* __restrict__ freq32) {
    extern __shared__ uint32_t smem_hist[];
    __shared__ unsigned long long s_total;
    __shared__ uint8_t s_lut[256];
    const int64_t rec = blockIdx.x;
    const int64_t begin = rec_begin[rec];
    const int64_t end = rec_end[rec];
    const int tid = threadIdx.x;
    uint32_t* hist0 = GLOBAL_HIST ? (counts + rec * dim) : smem_hist;
    uint32_t* hist = GLOBAL_HIST ? hist0 : (smem_hist + (ncopy > 1 ? (tid >> 5) * dim : 0));

    for (int c = tid; c < 256; c += PROFILE_THREADS) s_lut[c] = (uint8_t)classify_byte((uint32_t)c);
    if (!GLOBAL_HIST) {
        for (int64_t b = tid; b < dim * ncopy; b += PROFILE_THREADS) smem_hist[b] = 0u;
    }
    if (tid == 0) s_total = 0ull;
    __syncthreads();

    const int P = g.width;
    const uint32_t pmask = (P >= 32) ? 0xFFFFFFFFu : ((1u << P) - 1u);
    const int top_shift = 2 * (P - 1);

    // one base: classify, roll, count
    auto step = [&](Window<WIDE>& win, uint32_t c) {
        const uint32_t cls = s_lut[c];
        if (!(cls & 8u)) {
            if (mode == 0) win.template push<false>(cls, top_shift);
            else win.template push<true>(cls, top_shift);
            if ((win.inv & pmask) == 0u) {
                if (mode != 1) atomicAdd(&hist[gather<WIDE, NRUNS>(win.fw, g)], 1u);
                if (mode != 0) atomicAdd(&hist[gather<WIDE, NRUNS>(win.rc, g)], 1u);
            }
        }
    };

    if (end > begin) {
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
                    if (!(s_lut[__ldg(text + q)] & 8u)) --need;
                }
                for (; q < cstart; ++q) {
                    const uint32_t cls = s_lut[__ldg(text + q)];
                    if (!(cls & 8u)) {
                        if (mode == 0) win.template push<false>(cls, top_shift);
                        else win.template push<true>(cls, top_shift);
                    }
                }
            }
#pragma unroll 1
            for (int wi = 0; wi < CHUNK / 16; ++wi) {
                const int64_t w0 = cstart + 16 * wi;
                if (w0 >= end) break;
                if (w0 >= begin && w0 + 16 <= end) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + w0));
                    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) step(win, __byte_perm(wds[j], 0u, 0x4440u + b));
                    }
                } else {  // a word that straddles the start or the end of the record
                    const int64_t lo = max(w0, begin), hi = min(w0 + 16, end);
                    for (int64_t q = lo; q < hi; ++q) step(win, (uint32_t)__ldg(text + q));
                }
            }
        }
    }
    __syncthreads();

    if (!GLOBAL_HIST) {
        // merge the per-warp copies into copy 0
        if (ncopy > 1) {
            for (int64_t b = tid; b < dim; b += PROFILE_THREADS) {
                uint32_t s = smem_hist[b];
                for (int c = 1; c < ncopy; ++c) s += smem_hist[c * dim + b];
                smem_hist[b] = s;
            }
            __syncthreads();
        }
        // palindromic pattern: minus-strand counts are the plus-strand counts of revcomp(w)
        if (fold) {
            for (int64_t b = tid; b < dim; b += PROFILE_THREADS) {
                const uint32_t r = revcomp_bin((uint32_t)b, g.k);
                if ((uint32_t)b <= r) {
                    const uint32_t x = smem_hist[b], y = smem_hist[r];
                    if (strand == PO_STRAND_BOTH) {
                        smem_hist[b] = x + y;
                        smem_hist[r] = x + y;
                    } else {  // minus
                        smem_hist[b] = y;
                        smem_hist[r] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    // junction windows of seq + revcomp(seq)  (bin/phyloligo.py:141)
    if (strand == PO_STRAND_BOTH && tid == 0 && P > 1 && end > begin) {
        uint8_t tail[PO_MAX_PATTERN];
        int m = 0;
        int64_t q = end;
        while (q > begin && m < P - 1) {
            --q;
            const uint32_t cls = s_lut[__ldg(text + q)];
            if (!(cls & 8u)) tail[m++] = (uint8_t)cls;  // tail[0] = last base
        }
        if (2 * m >= P) {
            Window<WIDE> win;
            win.reset();
            for (int t = m - 1; t >= 0; --t) win.template push<false>(tail[t], top_shift);  // forward order
            for (int t = 0; t < m; ++t) {
                // revcomp(tail): last base first, complemented; validity unchanged
                const uint32_t cls = tail[t];
                win.template push<false>((cls & 4u) | ((cls & 3u) ^ 1u), top_shift);
                if ((win.inv & pmask) == 0u) atomicAdd(&hist0[gather<WIDE, NRUNS>(win.fw, g)], 1u);
            }
        }
    }
    __syncthreads();

    // total number of words and the outputs
    unsigned long long part = 0ull;
    for (int64_t b = tid; b < dim; b += PROFILE_THREADS) part += hist0[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if ((tid & 31) == 0 && part) atomicAdd(&s_total, part);
    __syncthreads();
    const unsigned long long total = s_total;
    if (tid == 0 && totals) totals[rec] = total;
    const double dt = (double)total;
    for (int64_t b = tid; b < dim; b += PROFILE_THREADS) {
        const uint32_t c = hist0[b];
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
    // which words are counted per window
    const int fold = (!global_hist && g.palindromic && strand != PO_STRAND_PLUS) ? 1 : 0;
    const int mode = (strand == PO_STRAND_PLUS || fold) ? 0 : (strand == PO_STRAND_MINUS ? 1 : 2);
    if (global_hist) {
        if (!d_counts) {
            set_error("patterns with more than 7 ones need d_counts (global histogram)");
            return PO_ERR_UNSUPPORTED;
        }
        PO_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, (size_t)n * hist_bytes, stream));
        LaunchTimer t(0, stream);
        profile_kernel<WIDE, NRUNS, true><<<(unsigned)n, PROFILE_THREADS, 0, stream>>>(
            d_text, d_begin, d_end, g, strand, mode, 0, 1, dim, d_counts, d_totals, d_freq64, d_freq32);
        count_launch(0);
    } else {
        auto kern = profile_kernel<WIDE, NRUNS, false>;
        const int ncopy = (dim <= 1024) ? PROFILE_WARPS : 1;
        const size_t smem = hist_bytes * ncopy;
        if (smem > 48 * 1024) {
            PO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        LaunchTimer t(0, stream);
        kern<<<(unsigned)n, PROFILE_THREADS, smem, stream>>>(
            d_text, d_begin, d_end, g, strand, mode, fold, ncopy, dim, d_counts, d_totals, d_freq64, d_freq32);
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
    if (g.nruns == 1 && g.dst[0] == 0)
        return launch_profile_t<WIDE, 1>(d_text, d_begin, d_end, n, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32, stream);
    if (g.nruns <= 4)
        return launch_profile_t<WIDE, 4>(d_text, d_begin, d_end, n, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32, stream);
    return launch_profile_t<WIDE, 16>(d_text, d_begin, d_end, n, g, strand, dim, d_counts, d_totals, d_freq64, d_freq32, stream);
}

}  // namespace po
