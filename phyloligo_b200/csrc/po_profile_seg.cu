// Composition profiling, fast path: line-segment streaming with SIMD-in-register
// base packing (sm_100a).
//
// Replaces select_strand + .upper() + cut_sequence_and_count_pattern + count2freq
// (reference bin/phyloligo.py:124-149, 683, 601-631, 633-661) for patterns of width
// 4..16.  Strand plus, or any strand with a palindromic pattern (every contiguous k-mer):
// plus-strand words only, the minus strand is folded in at the end (RCMODE 0).  A spaced
// pattern that is not its own mirror image (111010011) with strand minus / both: the minus
// strand's word of a window is read from a second rolling register that holds the reverse
// complement of the last 16 bases (RCMODE 1: that word only, 2: both words, two atomics per
// window, no fold).  Everything else (wider patterns, more than 7 ones) runs in
// profile_kernel (po_profile_kernel.cuh), which is also the semantic model of this one.
//
// Two launch shapes share the code.  Warp-per-record (histograms of <= 4096 bins): every
// warp owns one record, its private histogram and staging ring, and never meets a CTA
// barrier.  CTA-per-record (larger histograms): four warps split the record's tiles and
// share one histogram.  The record's bytes are cut into segments of S bytes,
// S = length of the record's first line + 1 (so that in a regularly wrapped FASTA file
// every segment is one line and its newline is the segment's last byte; any other
// content is still handled exactly, only slower).  A warp takes tiles of 32 consecutive
// segments; lane 0 stages a tile with one cp.async.bulk into a per-warp, double-buffered
// shared-memory ring (mbarrier completion), then every lane streams through its own
// segment:
//   * a 32-bit word = 4 bases.  (w & 0x06060606) * 0x00820820 packs the four 2-bit codes
//     (internal code = ASCII bits 1-2: A=0, C=1, T=2, G=3; complement = code ^ 2) into the
//     top byte, first base lowest.  The same word is validated with five logic/multiply
//     operations: rebuild the upper-case ASCII byte each code would have and compare.
//   * a 32-bit rolling register holds the last 16 bases; one PRMT appends 4 bases.  (RCMODE 1, 2: a
//     second multiply, ((w >> 1) & 0x03030303) * 0x40100401, packs the same four codes in reverse
//     order; complemented, it is appended at the LOW end of the reverse-complement register.)
//     Every window is one funnel shift + one mask (contiguous k-mers) and one
//     shared-memory atomic into the warp's private histogram copy.
//   * groups of 4 words that contain anything but ACGT/acgt (N runs, the newline at the
//     end of a line, stray blanks) fall back to a per-byte path with the same window
//     semantics as profile_kernel (a window counts iff all its bases are ACGT).
//   * windows that straddle two segments are counted afterwards from the previous
//     lane's last 16 bases (warp shuffle) and the segment's first 16 bases.
// The histogram lives in "internal" bin order (first base = lowest digit, codes A,C,T,G);
// the epilogue folds the reverse strand, replays the seq+revcomp(seq) junction windows
// (bin/phyloligo.py:141) and permutes the bins into the reference's
// product(("C","G","A","T")) order while writing counts / totals / frequencies.
#include <stdlib.h>
#include "po_common.cuh"

namespace po {

constexpr int SEG_THREADS = 128;
constexpr int SEG_WARPS = 4;
constexpr int SEG_MAX_S = 128;          // longest line (segment unit) in bytes
#ifndef PO_SEG_LANE_BYTES
#define PO_SEG_LANE_BYTES 168
#endif
#ifndef PO_SEG_NSTAGE
#define PO_SEG_NSTAGE 1
#endif
constexpr int SEG_LANE_BYTES = PO_SEG_LANE_BYTES;      // most bytes one lane streams per tile (whole lines)
constexpr int SEG_NSTAGE = PO_SEG_NSTAGE;              // staging buffers per warp
constexpr int SEG_STAGE_BYTES = 32 * SEG_LANE_BYTES + 64;  // one tile (+ alignment slack)

struct SegGeom {
    int width, k, nruns;
    unsigned shift[8];  // bit offset of run r inside the window (first base = bit 0)
    unsigned dst[8];    // bit offset of run r inside the word code (first '1' = bit 0)
    unsigned mask[8];
};

__device__ __forceinline__ unsigned sg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sg_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sg_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sg_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SG_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_WAIT_DONE;\n"
        "bra SG_WAIT_LOOP;\n"
        "SG_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sg_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// class byte of the per-byte path: bits 0-1 internal code, bit 2 invalid, bit 3 skip
__device__ __forceinline__ uint32_t seg_classify(uint32_t c) {
    if (c == 32u || (c >= 9u && c <= 13u)) return 8u;  // space, TAB, LF, VT, FF, CR: transparent
    const uint32_t up = c & 0xDFu;
    if (up == 0x41u) return 0u;  // A
    if (up == 0x43u) return 1u;  // C
    if (up == 0x54u) return 2u;  // T
    if (up == 0x47u) return 3u;  // G
    return 4u;
}

// reference bin (first base highest digit, codes C0 G1 A2 T3) -> internal bin (first base lowest
// digit, codes A0 C1 T2 G3).  Per digit (r1 r0) -> (i1 i0) = (r0, ~r1); a full bit reversal reverses
// the digit order and swaps the two bits of every digit, so only the inversion is left.
__device__ __forceinline__ uint32_t seg_ref_to_internal(uint32_t r, int k) {
    const uint32_t m = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    return ((__brev(r) >> (32 - 2 * k)) ^ 0x55555555u) & m;
}
// the inverse map: reversing 0101...01 gives 1010...10
__device__ __forceinline__ uint32_t seg_internal_to_ref(uint32_t i, int k) {
    const uint32_t m = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    return ((__brev(i) >> (32 - 2 * k)) ^ 0xAAAAAAAAu) & m;
}
// bin of the reverse-complement word, internal order (complement = code ^ 2)
__device__ __forceinline__ uint32_t seg_revcomp_bin(uint32_t w, int k) {
    if (k == 0) return 0u;
    uint32_t r = __brev(w) >> (32 - 2 * k);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    const uint32_t m = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    return (r ^ 0xAAAAAAAAu) & m;
}

template <int NRUNS>
__device__ __forceinline__ uint32_t seg_word(uint32_t e, const SegGeom& g) {
    if (NRUNS == 1) return e & g.mask[0];
    uint32_t w = 0;
#pragma unroll
    for (int r = 0; r < NRUNS; ++r)
        if (r < g.nruns) w |= ((e >> g.shift[r]) & g.mask[r]) << g.dst[r];
    return w;
}

// SWZ (histograms of >= 1024 bins): bin w lives at w ^ (top five bits of w).  The epilogue reads the
// histogram in bit-reversed order (the reference's bin order reverses the digits) and the fold pairs
// every bin with its reverse complement: with the plain layout all 32 lanes of those accesses hit one
// bank (ncu, 4096 bins: 83 % of the kernel's shared-memory wavefronts were bank conflicts, 12 000 of
// the 14 600 wavefronts per record came from the epilogue); with the swizzle both the natural and the
// bit-reversed order are conflict free, for two more logic operations per counted window.
template <int NRUNS, bool WARP_REC, int RCMODE, bool SWZ>
__global__ void __launch_bounds__(SEG_THREADS)
profile_seg_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ rec_begin,
                   const int64_t* __restrict__ rec_end, int64_t nrec, const SegGeom g, int strand, int fold,
                   int ncopy, int64_t dim, uint32_t* __restrict__ counts, uint64_t* __restrict__ totals,
                   double* __restrict__ freq64, float* __restrict__ freq32) {
    extern __shared__ __align__(128) unsigned char seg_smem[];
    __shared__ __align__(8) unsigned long long s_bars[SEG_WARPS * 2];
    __shared__ unsigned long long s_total;
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_jw[SEG_WARPS][16];  // junction words (internal bins) of the record(s) of this CTA
    __shared__ int s_nj[SEG_WARPS];
    // WARP_REC: [warp][hist dim] then the staging rings; else [ncopy][hist dim] then the rings
    const int nhist = WARP_REC ? SEG_WARPS : ncopy;
    uint32_t* smem_hist = reinterpret_cast<uint32_t*>(seg_smem);
    unsigned char* stage_base = seg_smem + (((size_t)dim * nhist * 4 + 127) & ~(size_t)127);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rec = WARP_REC ? (int64_t)blockIdx.x * SEG_WARPS + warp : (int64_t)blockIdx.x;
    const bool rec_ok = rec < nrec;
    const int64_t begin = rec_ok ? rec_begin[rec] : 0;
    const int64_t end = rec_ok ? rec_end[rec] : 0;
    // the histogram this warp counts into (byte offset from the start of dynamic shared memory)
    const uint32_t hoff = (uint32_t)((nhist > 1 ? warp : 0) * dim * 4);
    uint32_t* hist0 = WARP_REC ? smem_hist + (size_t)warp * dim : smem_hist;  // the record's final histogram
    const int nthr = WARP_REC ? 32 : SEG_THREADS;   // threads that cooperate on one record
    const int rtid = WARP_REC ? lane : tid;
    auto rec_sync = [&]() {
        if (WARP_REC) __syncwarp();
        else __syncthreads();
    };

    for (int c = tid; c < 256; c += SEG_THREADS) s_lut[c] = (uint8_t)seg_classify((uint32_t)c);
    for (int64_t b = tid; b < dim * nhist; b += SEG_THREADS) smem_hist[b] = 0u;
    if (tid == 0) {
        s_total = 0ull;
#pragma unroll
        for (int i = 0; i < SEG_WARPS * 2; ++i) sg_mbar_init(sg_smem_u32(&s_bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // length of the record's first line (position of the first '\n' within 128 bytes)
    int first_nl = 1 << 20;
    {
        uint32_t hit = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t pos = begin + 4 * lane + q;
            if (pos < end && __ldg(text + pos) == 10u) hit |= 1u << q;
        }
        const uint32_t any = __ballot_sync(0xFFFFFFFFu, hit != 0u);
        if (any) {
            const int l0 = __ffs(any) - 1;
            const uint32_t h0 = __shfl_sync(0xFFFFFFFFu, hit, l0);
            first_nl = 4 * l0 + __ffs(h0) - 1;
        }
    }

    const int P = g.width;
    const uint32_t pmask = (1u << P) - 1u;           // P <= 16
    const uint32_t tailmask = (1u << (P - 1)) - 1u;  // validity bits of the last P-1 bases
    const int wshift = 32 - 2 * P;                   // per-byte path: window = W >> wshift

    // histogram update at byte offset `o4` of this warp's histogram: one ATOMS with an immediate base
    const int swsh = 2 * g.k - 5;  // SWZ: the bin's top five bits, as a byte offset: (o4 >> swsh) & 0x7C
    auto bump = [&](uint32_t o4) {
        if (SWZ) o4 ^= (o4 >> swsh) & 0x7Cu;
        atomicAdd(reinterpret_cast<uint32_t*>(seg_smem + (o4 | hoff)), 1u);
    };
    auto SW = [&](uint32_t b) -> uint32_t { return SWZ ? (b ^ ((b >> swsh) & 31u)) : b; };
    auto count_window = [&](uint32_t e) { bump(seg_word<NRUNS>(e, g) << 2); };
    // one base through the per-byte path (semantics of profile_kernel).  R = reverse complement of the
    // last 16 bases, the newest base (complemented: code ^ 2) in the lowest digit.
    auto push_byte = [&](uint32_t& W, uint32_t& R, uint32_t& inv, uint32_t c, bool do_count) {
        const uint32_t cls = s_lut[c];
        if (!(cls & 8u)) {
            W = (W >> 2) | ((cls & 3u) << 30);
            if (RCMODE != 0) R = (R << 2) | ((cls & 3u) ^ 2u);
            inv = (inv << 1) | ((cls >> 2) & 1u);
            if (do_count && (inv & pmask) == 0u) {
                if (RCMODE != 1) count_window(W >> wshift);
                if (RCMODE != 0) count_window(R);
            }
        }
    };

    uint32_t lastW = 0, last_inv = 0xFFFFFFFFu;
    if (end > begin) {
        const int S = (first_nl + 1 >= 17 && first_nl + 1 <= SEG_MAX_S) ? first_nl + 1 : SEG_MAX_S;
        // a lane streams through L consecutive lines per tile; L balances the lanes over the fewest tiles
        const int64_t nlines = (end - begin + S - 1) / S;
        const int Lmax = SEG_LANE_BYTES / S;
        const int64_t tiles_min = (nlines + 32 * Lmax - 1) / (32 * Lmax);
        const int L = (int)((nlines + 32 * tiles_min - 1) / (32 * tiles_min));
        const int LS = L * S;                       // bytes per lane per tile
        const int64_t nseg = (nlines + L - 1) / L;  // lane segments in the record
        const int64_t ntiles = (nseg + 31) / 32;
        const int64_t t_lo = WARP_REC ? 0 : warp * ntiles / SEG_WARPS;
        const int64_t t_hi = WARP_REC ? ntiles : (warp + 1) * ntiles / SEG_WARPS;
        unsigned char* wstage = stage_base + (size_t)warp * SEG_NSTAGE * SEG_STAGE_BYTES;
        const unsigned bar0 = sg_smem_u32(&s_bars[warp * 2]);

        const int64_t tile_stride = (int64_t)32 * LS;
        auto issue = [&](int64_t tb, int it) {  // stage the tile that starts at byte tb
            const int64_t te = min(end, tb + tile_stride);
            const int64_t tb_al = tb & ~(int64_t)15;
            const unsigned bytes = (unsigned)(((te + 15) & ~(int64_t)15) - tb_al);
            const unsigned bar = bar0 + 8 * (it % SEG_NSTAGE);
            sg_mbar_expect_tx(bar, bytes);
            sg_bulk_g2s(sg_smem_u32(wstage + (size_t)(it % SEG_NSTAGE) * SEG_STAGE_BYTES), text + tb_al, bytes, bar);
        };
        int64_t tb = begin + t_lo * tile_stride;
        if (lane == 0 && t_lo < t_hi) issue(tb, 0);

        // per-window shifts of the funnel {rolling register, new byte}: window ending at the
        // j-th base of a new word starts at bit 2 (17 + j - P) of that 40-bit value
        int sh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[j] = 2 * (17 + j - P) - (NRUNS == 1 ? 2 : 0);
        const uint32_t mask4 = g.mask[0] << 2;

        uint32_t carryW = 0, carryR = 0, carry_inv = 0xFFFFFFFFu;
        const int last_lane = (int)((nseg - 1) & 31);
        int it = 0;
        for (int64_t t = t_lo; t < t_hi; ++t, ++it, tb += tile_stride) {
            if (SEG_NSTAGE > 1 && lane == 0 && t + 1 < t_hi) issue(tb + tile_stride, it + 1);
            const int64_t te = min(end, tb + tile_stride);
            const int64_t tb_al = tb & ~(int64_t)15;
            sg_mbar_wait(bar0 + 8 * (it % SEG_NSTAGE), (unsigned)((it / SEG_NSTAGE) & 1));
            const unsigned char* buf = wstage + (size_t)(it % SEG_NSTAGE) * SEG_STAGE_BYTES;

            // byte offsets inside the staged tile
            const unsigned toff = (unsigned)(tb - tb_al);
            const unsigned tend = (unsigned)(te - tb_al);
            const unsigned seg0 = toff + (unsigned)lane * (unsigned)LS;
            const bool active = seg0 < tend;
            uint32_t W = 0, R = 0, inv = 0xFFFFFFFFu, headW = 0, headR = 0;
            bool head_ok = false;
            if (active) {
                const unsigned seg1 = min(tend, seg0 + (unsigned)LS);
                for (unsigned off = seg0; off < seg1; off += (unsigned)S) {  // one line
                    const bool first = off == seg0;
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(buf + (off & ~3u));
                    const unsigned sel = 0x3210u + 0x1111u * (off & 3u);
                    const unsigned char* bp = buf + off;
                    const int nbytes = (int)(min(seg1, off + (unsigned)S) - off);
                    const int nwords = nbytes >> 2;
                    uint32_t lo = wp[0];
                    uint32_t hiw[4];  // the next four aligned words, loaded one group ahead
#pragma unroll
                    for (int i = 0; i < 4; ++i) hiw[i] = wp[i + 1];
                    int wi = 0;
                    for (; wi + 4 <= nwords; wi += 4) {
                        uint32_t w[4], prod[4], rcp4[4], bad = 0u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t hi = hiw[i];
                            hiw[i] = wp[wi + i + 5];  // may run past the line: stays inside the stage buffer
                            w[i] = __byte_perm(lo, hi, sel);
                            lo = hi;
                            const uint32_t y = w[i] & 0x06060606u;
                            prod[i] = y * 0x00820820u;
                            // the same four codes in reverse order (last base lowest), complemented, in the top byte
                            if (RCMODE != 0) rcp4[i] = ((y >> 1) * 0x40100401u) ^ 0xAA000000u;
                            const uint32_t tb0 = (w[i] >> 2) & ~(w[i] >> 1) & 0x01010101u;  // 1 where the code is T
                            const uint32_t t11 = tb0 * 0x11u;
                            bad |= ((0x41414141u | y) ^ w[i] ^ t11) & 0xDFDFDFDFu;
                        }
                        if (bad == 0u && (inv & tailmask) == 0u) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t nb = prod[i] >> 24;
                                const uint32_t Rhi = R >> 24;
                                if (RCMODE != 0) R = __byte_perm(R, rcp4[i], 0x2107);  // R << 8 | reversed complemented codes
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (RCMODE != 1) {
                                        const uint32_t e = __funnelshift_rc(W, nb, sh[j]);
                                        if (NRUNS == 1) bump(e & mask4);
                                        else count_window(e);
                                    }
                                    // the register as it was after base j of this word: its low 2P bits are the window
                                    if (RCMODE != 0) count_window(j == 3 ? R : __funnelshift_r(R, Rhi, 2 * (3 - j)));
                                }
                                W = __byte_perm(W, prod[i], 0x7321);
                            }
                            inv <<= 16;
                        } else if (bad == 0u && first && wi == 0) {
                            // first group of the lane's segment: windows that need bases before it
                            // are left to the boundary pass
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t nb = prod[i] >> 24;
                                const uint32_t Rhi = R >> 24;
                                if (RCMODE != 0) R = __byte_perm(R, rcp4[i], 0x2107);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (4 * i + j >= P - 1) {
                                        if (RCMODE != 1) {
                                            const uint32_t e = __funnelshift_rc(W, nb, sh[j]);
                                            if (NRUNS == 1) bump(e & mask4);
                                            else count_window(e);
                                        }
                                        if (RCMODE != 0) count_window(j == 3 ? R : __funnelshift_r(R, Rhi, 2 * (3 - j)));
                                    }
                                }
                                W = __byte_perm(W, prod[i], 0x7321);
                            }
                            inv = 0xFFFF0000u;
                            headW = W;
                            headR = R;
                            head_ok = true;
                        } else {
                            for (int q = 0; q < 16; ++q) push_byte(W, R, inv, bp[4 * wi + q], true);
                        }
                    }
                    // what is left of the line: in a wrapped FASTA file exactly its newline
                    if (!(nbytes - 4 * wi == 1 && bp[4 * wi] == 10u))
                        for (int q = 4 * wi; q < nbytes; ++q) push_byte(W, R, inv, bp[q], true);
                }
            }
            // ---- windows that straddle the start of this lane's segment ----
            uint32_t prevW = __shfl_up_sync(0xFFFFFFFFu, W, 1);
            uint32_t prevR = RCMODE != 0 ? __shfl_up_sync(0xFFFFFFFFu, R, 1) : 0u;
            uint32_t prev_inv = __shfl_up_sync(0xFFFFFFFFu, inv, 1);
            bool have_prev = lane > 0;
            if (lane == 0) {
                prevW = carryW;
                prevR = carryR;
                prev_inv = carry_inv;
                have_prev = t > t_lo;
            }
            carryW = __shfl_sync(0xFFFFFFFFu, W, 31);
            if (RCMODE != 0) carryR = __shfl_sync(0xFFFFFFFFu, R, 31);
            carry_inv = __shfl_sync(0xFFFFFFFFu, inv, 31);
            if (WARP_REC && t == ntiles - 1) {  // state after the record's last base, for the junction
                lastW = __shfl_sync(0xFFFFFFFFu, W, last_lane);
                last_inv = __shfl_sync(0xFFFFFFFFu, inv, last_lane);
            }
            const int64_t s0 = tb_al + seg0;
            if (active && P > 1 && s0 > begin) {
                if (have_prev && head_ok && (prev_inv & tailmask) == 0u) {
                    for (int j = 0; j < P - 1; ++j) {
                        if (RCMODE != 1) count_window(__funnelshift_r(prevW, headW, 2 * (17 + j - P)));
                        // {prevR : headR} is the register after the segment's first 16 bases; after base j it was 2 (15 - j) bits shorter
                        if (RCMODE != 0) count_window(__funnelshift_r(headR, prevR, 2 * (15 - j)));
                    }
                } else {
                    // generic: rebuild the P-1 bases before s0 from memory, then replay this
                    // segment's first P-1 bases
                    const int64_t s1 = min(te, s0 + LS);
                    int need = P - 1;
                    int64_t q = s0;
                    while (q > begin && need > 0) {
                        --q;
                        if (!(s_lut[__ldg(text + q)] & 8u)) --need;
                    }
                    uint32_t W2 = 0, R2 = 0, inv2 = 0xFFFFFFFFu;
                    for (; q < s0; ++q) push_byte(W2, R2, inv2, __ldg(text + q), false);
                    int nb = 0;
                    for (q = s0; q < s1 && nb < P - 1; ++q) {
                        const uint32_t c = __ldg(text + q);
                        if (!(s_lut[c] & 8u)) {
                            push_byte(W2, R2, inv2, c, true);
                            ++nb;
                        }
                    }
                }
            }
            __syncwarp();  // every lane is done with this stage before it is refilled
            if (SEG_NSTAGE == 1 && lane == 0 && t + 1 < t_hi) issue(tb + tile_stride, it + 1);
        }
    }
    rec_sync();

    // ---- junction words, total, outputs ----
    // The minus strand of a folded run is not written back into the histogram: the output loop reads
    // bin w and its reverse complement (both conflict free with the swizzle) and adds them.  The
    // P - 1 junction windows of seq + revcomp(seq) (bin/phyloligo.py:141) are replayed by one thread
    // from the record's last P - 1 bases into a short list; they count towards the total, and their
    // bins are rewritten after the output loop.
    const int slot = WARP_REC ? warp : 0;
    if (rtid == 0) {
        int nj = 0;
        if (WARP_REC && strand == PO_STRAND_BOTH && P > 1 && end > begin && (last_inv & tailmask) == 0u) {
            // the record's last P-1 bases are valid and still in the rolling register
            uint32_t W = lastW;
            for (int t = 0; t < P - 1; ++t) {
                W = (W >> 2) | ((((lastW >> (30 - 2 * t)) & 3u) ^ 2u) << 30);
                s_jw[slot][nj++] = seg_word<NRUNS>(W >> wshift, g);
            }
        } else if (strand == PO_STRAND_BOTH && P > 1 && end > begin) {
            uint8_t tail[16];
            int m = 0;
            int64_t q = end;
            while (q > begin && m < P - 1) {
                --q;
                const uint32_t cls = s_lut[__ldg(text + q)];
                if (!(cls & 8u)) tail[m++] = (uint8_t)cls;  // tail[0] = last base
            }
            if (2 * m >= P) {
                uint32_t W = 0, inv = 0xFFFFFFFFu;
                for (int t = m - 1; t >= 0; --t) {  // forward order
                    W = (W >> 2) | ((uint32_t)(tail[t] & 3u) << 30);
                    inv = (inv << 1) | ((tail[t] >> 2) & 1u);
                }
                for (int t = 0; t < m; ++t) {  // revcomp(tail): last base first, complemented
                    W = (W >> 2) | ((uint32_t)((tail[t] & 3u) ^ 2u) << 30);
                    inv = (inv << 1) | ((tail[t] >> 2) & 1u);
                    if ((inv & pmask) == 0u) s_jw[slot][nj++] = seg_word<NRUNS>(W >> wshift, g);
                }
            }
        }
        s_nj[slot] = nj;
    }
    unsigned long long part = 0ull;
    for (int64_t b = rtid; b < dim; b += nthr) part += hist0[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if (!WARP_REC && lane == 0 && part) atomicAdd(&s_total, part);
    rec_sync();
    const int nj = s_nj[slot];
    const bool both_folded = fold && strand == PO_STRAND_BOTH;
    const unsigned long long total = (WARP_REC ? part : s_total) * (both_folded ? 2ull : 1ull) + (unsigned long long)nj;
    // count of reference bin r before the junction words
    auto bin_count = [&](uint32_t i) -> uint32_t {  // i: internal bin
        if (!fold) return hist0[SW(i)];
        const uint32_t y = hist0[SW(seg_revcomp_bin(i, g.k))];
        return both_folded ? hist0[SW(i)] + y : y;
    };
    // count / total exactly as Python's int / int (= the correctly rounded quotient): q = c * RN(1/t)
    // followed by two residual corrections r = c - q t (exact in an FMA), q += r * RN(1/t)
    // (Markstein; every quotient of integers below 2^53 comes out correctly rounded) -- five FP64
    // operations instead of the ~50 instructions of a generic division per bin
    const double dt = (double)total;
    const double rt = total ? 1.0 / dt : 0.0;
    auto write_bin = [&](int64_t r, uint32_t c) {
        if (counts) counts[rec * dim + r] = c;
        const double dc = (double)c;
        double f = dc * rt;
        f = fma(fma(-dt, f, dc), rt, f);
        f = fma(fma(-dt, f, dc), rt, f);
        if (freq64) freq64[rec * dim + r] = f;
        if (freq32) freq32[rec * dim + r] = (float)f;
    };
    if (rec_ok) {
        if (rtid == 0 && totals) totals[rec] = total;
        for (int64_t r = rtid; r < dim; r += nthr) write_bin(r, bin_count(seg_ref_to_internal((uint32_t)r, g.k)));
    }
    if (nj > 0) {  // uniform over the record's threads
        rec_sync();  // the bins below were written by other threads a moment ago
        if (rtid == 0 && rec_ok) {
            for (int a = 0; a < nj; ++a) {
                const uint32_t w = s_jw[slot][a];
                uint32_t mult = 0;
                bool seen = false;
                for (int b = 0; b < nj; ++b) {
                    if (s_jw[slot][b] == w) {
                        ++mult;
                        seen = seen || b < a;
                    }
                }
                if (!seen) write_bin((int64_t)seg_internal_to_ref(w, g.k), bin_count(w) + mult);
            }
        }
    }
}

// true when the segment kernel covers this request
bool profile_seg_supported(const PatternGeom& g, int strand, int64_t dim) {
    if (g.width < 4 || g.width > 16 || g.k < 1 || g.nruns > 4) return false;
    return (size_t)dim * 4 <= 96 * 1024;
}

int launch_profile_seg(const uint8_t* d_text, const int64_t* d_begin, const int64_t* d_end, int64_t n,
                       const PatternGeom& g, int strand, uint32_t* d_counts, uint64_t* d_totals,
                       double* d_freq64, float* d_freq32, cudaStream_t stream) {
    const int64_t dim = (int64_t)1 << (2 * g.k);
    SegGeom sg;
    memset(&sg, 0, sizeof(sg));
    sg.width = g.width;
    sg.k = g.k;
    sg.nruns = g.nruns;
    for (int r = 0; r < g.nruns; ++r) {
        // PatternGeom counts from the last base (shift = 2 (width-1-o2), dst = 2 (k-1-j2))
        int len = 0;
        for (unsigned m = g.mask[r]; m; m >>= 2) ++len;
        const int o2 = g.width - 1 - g.shift[r] / 2, j2 = g.k - 1 - g.dst[r] / 2;
        sg.shift[r] = 2u * (unsigned)(o2 - len + 1);
        sg.dst[r] = 2u * (unsigned)(j2 - len + 1);
        sg.mask[r] = g.mask[r];
    }
    // minus-strand words: folded in at the end when the pattern is its own mirror image, else read
    // from the reverse-complement register (rcmode 1: only those, 2: both words of every window)
    const int fold = (strand != PO_STRAND_PLUS && g.palindromic) ? 1 : 0;
    const int rcmode = (strand == PO_STRAND_PLUS || g.palindromic) ? 0 : (strand == PO_STRAND_MINUS ? 1 : 2);
    const bool contiguous = (g.nruns == 1 && sg.shift[0] == 0);
    // One warp per record with a private histogram up to 1024 bins (4 KB: 24 warps per SM).  With 4096
    // bins a private histogram leaves 8 warps per SM and the kernel latency bound, so the four warps of a
    // CTA share one histogram and one record (20 warps per SM; measured at 50 000 x 15 kb, 6-mers:
    // 1.58 ms against 1.98 ms) -- except when every window costs two multi-run words (the reverse-
    // complement register mode), where splitting a 15 kb record's few tiles over four warps loses more
    // than the occupancy gains (4.17 ms against 3.74 ms).  PO_SEG_WARPREC_MAXDIM overrides the limit.
    int64_t warp_rec_max = rcmode != 0 ? 4096 : 1024;
    if (const char* e = getenv("PO_SEG_WARPREC_MAXDIM")) warp_rec_max = atoll(e);
    const bool warp_rec = dim <= warp_rec_max;
    const int ncopy = warp_rec ? 1 : 1;
    const int nhist = warp_rec ? SEG_WARPS : ncopy;
    const size_t hist_bytes = ((size_t)dim * nhist * 4 + 127) & ~(size_t)127;
    const size_t smem = hist_bytes + (size_t)SEG_WARPS * SEG_NSTAGE * SEG_STAGE_BYTES;
    const unsigned grid = warp_rec ? (unsigned)((n + SEG_WARPS - 1) / SEG_WARPS) : (unsigned)n;
    LaunchTimer t(0, stream);
#define PO_SEG_LAUNCH1(NR, WR, RC, SZ)                                                                               \
    do {                                                                                                             \
        auto kern = profile_seg_kernel<NR, WR, RC, SZ>;                                                              \
        PO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        kern<<<grid, SEG_THREADS, smem, stream>>>(d_text, d_begin, d_end, n, sg, strand, fold, ncopy, dim, d_counts, \
                                                  d_totals, d_freq64, d_freq32);                                     \
    } while (0)
#define PO_SEG_LAUNCH(NR, WR, RC)                  \
    do {                                           \
        if (swz) PO_SEG_LAUNCH1(NR, WR, RC, true); \
        else PO_SEG_LAUNCH1(NR, WR, RC, false);    \
    } while (0)
    const bool swz = dim >= 1024;
    if (contiguous && warp_rec) PO_SEG_LAUNCH(1, true, 0);
    else if (contiguous) PO_SEG_LAUNCH(1, false, 0);
    else if (warp_rec && rcmode == 0) PO_SEG_LAUNCH(4, true, 0);
    else if (warp_rec && rcmode == 1) PO_SEG_LAUNCH(4, true, 1);
    else if (warp_rec) PO_SEG_LAUNCH(4, true, 2);
    else if (rcmode == 0) PO_SEG_LAUNCH(4, false, 0);
    else if (rcmode == 1) PO_SEG_LAUNCH(4, false, 1);
    else PO_SEG_LAUNCH(4, false, 2);
#undef PO_SEG_LAUNCH
#undef PO_SEG_LAUNCH1
    count_launch(0);
    PO_LAUNCH_CHECK("profile_seg_kernel");
    return PO_OK;
}

}  // namespace po
