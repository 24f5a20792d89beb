// Host-side writer of the reference's text matrix format (no device code).
//
// Replaces np.savetxt(path, M, delimiter="\t") at reference bin/phyloligo.py:1059-1066:
// every entry as "%.18e" (the float64 value; a float32 entry is widened first, as numpy
// does), fields separated by one tab, rows ended by '\n', no header.  Readers:
// bin/phyloligo_comparemat.py:7-14, bin/phyloselect.py:616-622, bin/phyloselect.R:324.
// numpy's writer is a per-entry Python loop (~1 us per entry); here row blocks are
// formatted by a pool of threads into private buffers (a printf-free, correctly rounded
// '%.18e': ~20 M entries/s per core, ten times glibc's) and written in order by a writer
// thread that overlaps the formatting of the next wave.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>
#include "po_common.cuh"
#include "po_pow10_table.h"

namespace po {

// ---- '%.18e' without printf -----------------------------------------------------------------
// 19 significant digits, correctly rounded (half to even on the exact binary value), as glibc and
// Python print them.  x = m * 2^e is scaled by a power of ten from a table of 128-bit mantissas
// (po_pow10_table.h) so that V = x * 10^(18-k) lies in [10^18, 10^19); the 192-bit product
// m * mantissa gives floor(V) and the exact remainder.  For table entries that are exact
// (0 <= p <= 38, i.e. 1e-20 <= |x| < 1e19: every frequency and distance) the rounding -- ties
// included -- is decided exactly.  Elsewhere the truncated mantissa leaves V uncertain by less than
// 2^-63; if the remainder is that close to a rounding boundary the entry goes through snprintf
// (about once in 2^60 entries), so the output is always the correctly rounded one.
static const char DIGIT_PAIRS[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
    "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
    "8081828384858687888990919293949596979899";

static inline void put9(char* dst, uint32_t v) {  // exactly 9 digits
    const uint32_t a = v / 10000000u;             // 2 digits
    uint32_t r = v - a * 10000000u;               // 7 digits
    const uint32_t b = r / 100000u;
    r -= b * 100000u;                             // 5 digits
    const uint32_t c = r / 1000u;
    r -= c * 1000u;                               // 3 digits
    const uint32_t d = r / 10u;
    const uint32_t e = r - d * 10u;
    memcpy(dst, DIGIT_PAIRS + 2 * a, 2);
    memcpy(dst + 2, DIGIT_PAIRS + 2 * b, 2);
    memcpy(dst + 4, DIGIT_PAIRS + 2 * c, 2);
    memcpy(dst + 6, DIGIT_PAIRS + 2 * d, 2);
    dst[8] = (char)('0' + e);
}

// floor(V), and how the remainder compares with one half: -1 below, 0 exactly half, +1 above,
// 2 = cannot be decided with this table entry, 3 = V >= 2^64 (k is too small)
static inline int scaled_digits(uint64_t m, int e, int p, uint64_t* out) {
    const Pow10Entry& t = POW10_TABLE[p - POW10_MIN];
    typedef unsigned __int128 u128;
    const u128 lo = (u128)m * t.lo, hi = (u128)m * t.hi;
    uint64_t w[3];
    w[0] = (uint64_t)lo;
    const u128 mid = (lo >> 64) + (uint64_t)hi;
    w[1] = (uint64_t)mid;
    w[2] = (uint64_t)(hi >> 64) + (uint64_t)(mid >> 64);
    const int r = -(t.exp2 + e);  // V = (w2 w1 w0) / 2^r
    if (r < 66 || r > 191) return 2;
    // integer part and remainder of the 192-bit value shifted right by r
    const int ws = r / 64, bs = r % 64;
    uint64_t ipart;
    if (ws == 1) {
        if (bs == 0) { if (w[2]) return 3; ipart = w[1]; }
        else { if (w[2] >> bs) return 3; ipart = (w[1] >> bs) | (w[2] << (64 - bs)); }
    } else {  // ws == 2
        ipart = bs ? (w[2] >> bs) : w[2];
    }
    // remainder = low r bits; compare with half = 2^(r-1)
    uint64_t rem[3] = {w[0], w[1], w[2]};
    if (ws == 1) { rem[2] = 0; rem[1] = bs ? (w[1] & ((1ull << bs) - 1)) : 0; }
    else { rem[2] = bs ? (w[2] & ((1ull << bs) - 1)) : 0; }
    const int hb = r - 1;  // bit index of one half
    uint64_t half[3] = {0, 0, 0};
    half[hb / 64] = 1ull << (hb % 64);
    int cmp = 0;
    for (int i = 2; i >= 0 && cmp == 0; --i) cmp = rem[i] > half[i] ? 1 : (rem[i] < half[i] ? -1 : 0);
    *out = ipart;
    if (p >= 0 && p <= 38) return cmp;  // exact entry: exact decision
    // truncated entry: the true remainder lies in [rem, rem + 2^(r-62)); undecidable near one half and near a carry
    uint64_t mg[3] = {0, 0, 0};
    mg[(r - 62) / 64] = 1ull << ((r - 62) % 64);
    uint64_t up[3];
    unsigned carry = 0;
    for (int i = 0; i < 3; ++i) {
        const u128 sum = (u128)rem[i] + mg[i] + carry;
        up[i] = (uint64_t)sum;
        carry = (unsigned)(sum >> 64);
    }
    uint64_t full[3] = {0, 0, 0};  // 2^r
    if (r < 192) full[r / 64] = 1ull << (r % 64);
    int cfull = 0;
    for (int i = 2; i >= 0 && cfull == 0; --i) cfull = up[i] > full[i] ? 1 : (up[i] < full[i] ? -1 : 0);
    if (carry || cfull >= 0) return 2;  // might carry into the integer part
    if (cmp > 0) return 1;
    int cup = 0;
    for (int i = 2; i >= 0 && cup == 0; --i) cup = up[i] > half[i] ? 1 : (up[i] < half[i] ? -1 : 0);
    return cup < 0 ? -1 : 2;
}

static inline size_t format_e18_fast(char* dst, double v) {
    uint64_t bits;
    memcpy(&bits, &v, 8);
    char* p0 = dst;
    if (bits >> 63) *dst++ = '-';
    const uint64_t frac = bits & 0xFFFFFFFFFFFFFull;
    const int bexp = (int)((bits >> 52) & 0x7FF);
    if (bexp == 0 && frac == 0) {
        memcpy(dst, "0.000000000000000000e+00", 24);
        return (size_t)(dst + 24 - p0);
    }
    const uint64_t m = bexp ? (frac | (1ull << 52)) : frac;
    const int e = bexp ? bexp - 1075 : -1074;
    // k = floor(log10 |x|), from the position of the leading bit (may be one too small)
    const int lead = 63 - __builtin_clzll(m) + e;  // floor(log2 |x|)
    int k = (int)(((long long)lead * 1292913987ll) >> 32);  // floor(lead * log10(2)), 2^32 * log10(2) = 1292913986.5
    uint64_t D = 0;
    int how = 2;
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int p = 18 - k;
        if (p < POW10_MIN || p > POW10_MAX) return 0;
        D = 0;
        how = scaled_digits(m, e, p, &D);
        if (how == 3) { ++k; how = 2; continue; }
        if (how == 2 && D == 0) return 0;
        if (D >= 10000000000000000000ull) { ++k; how = 2; continue; }
        if (D < 1000000000000000000ull) { --k; how = 2; continue; }
        break;
    }
    if (how == 2 || D < 1000000000000000000ull || D >= 10000000000000000000ull) return 0;
    if (how > 0 || (how == 0 && (D & 1))) {
        if (++D == 10000000000000000000ull) {
            D = 1000000000000000000ull;
            ++k;
        }
    }
    const uint64_t first = D / 1000000000000000000ull;
    const uint64_t rest = D - first * 1000000000000000000ull;
    dst[0] = (char)('0' + first);
    dst[1] = '.';
    put9(dst + 2, (uint32_t)(rest / 1000000000ull));
    put9(dst + 11, (uint32_t)(rest % 1000000000ull));
    dst += 20;
    *dst++ = 'e';
    *dst++ = k < 0 ? '-' : '+';
    const unsigned ak = (unsigned)(k < 0 ? -k : k);
    if (ak >= 100) {
        *dst++ = (char)('0' + ak / 100);
        memcpy(dst, DIGIT_PAIRS + 2 * (ak % 100), 2);
    } else {
        memcpy(dst, DIGIT_PAIRS + 2 * ak, 2);
    }
    dst += 2;
    return (size_t)(dst - p0);
}

// one entry, exactly as Python's '%.18e' % float(x) prints it
static inline size_t format_entry(char* dst, double v) {
    if (isnan(v)) {
        memcpy(dst, "nan", 3);
        return 3;
    }
    if (isinf(v)) {
        if (v < 0) {
            memcpy(dst, "-inf", 4);
            return 4;
        }
        memcpy(dst, "inf", 3);
        return 3;
    }
    const size_t n = format_e18_fast(dst, v);
    if (n) return n;
    return (size_t)snprintf(dst, 32, "%.18e", v);
}

template <typename T>
static void format_rows(const T* data, int64_t ld, int64_t cols, int64_t r0, int64_t r1, std::vector<char>& buf) {
    // "-d.dddddddddddddddddde+ddd" is 26 characters, plus the separator
    buf.resize((size_t)(r1 - r0) * (size_t)(cols > 0 ? cols : 1) * 28 + 64);
    char* p = buf.data();
    for (int64_t r = r0; r < r1; ++r) {
        const T* row = data + r * ld;
        for (int64_t c = 0; c < cols; ++c) {
            p += format_entry(p, (double)row[c]);
            *p++ = (c + 1 < cols) ? '\t' : '\n';
        }
        if (cols == 0) *p++ = '\n';
    }
    buf.resize((size_t)(p - buf.data()));
}

}  // namespace po

using namespace po;

extern "C" int po_savetxt_host(const char* path, const void* h_data, int64_t rows, int64_t cols, int64_t ld,
                               int dtype, int threads) {
    if (!path || rows < 0 || cols < 0 || ld < cols || (rows > 0 && cols > 0 && !h_data)) {
        set_error("po_savetxt_host: bad arguments");
        return PO_ERR_ARG;
    }
    if (dtype != PO_F32 && dtype != PO_F64) {
        set_error("po_savetxt_host: dtype must be PO_F32 or PO_F64");
        return PO_ERR_ARG;
    }
    FILE* fh = fopen(path, "wb");
    if (!fh) {
        set_error("po_savetxt_host: cannot open %s: %s", path, strerror(errno));
        return PO_ERR_ARG;
    }
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (threads > 256) threads = 256;
    // a wave = `threads` row blocks of about 4 MB of text each, formatted in parallel; wave k is
    // written (in order, by one writer thread) while wave k+1 is being formatted into the other
    // set of buffers
    const int64_t row_text = (cols > 0 ? cols : 1) * 26;
    int64_t block_rows = (4 << 20) / row_text;
    if (block_rows < 1) block_rows = 1;
    std::vector<std::vector<char>> bufs[2];
    bufs[0].resize((size_t)threads);
    bufs[1].resize((size_t)threads);
    int rc = PO_OK;
    bool write_failed = false;
    int write_errno = 0;
    std::thread writer;
    int which = 0;
    for (int64_t wave0 = 0; wave0 < rows; wave0 += block_rows * threads, which ^= 1) {
        std::vector<std::vector<char>>& cur = bufs[which];
        std::vector<std::thread> pool;
        int used = 0;
        for (int t = 0; t < threads; ++t) {
            const int64_t r0 = wave0 + (int64_t)t * block_rows;
            if (r0 >= rows) break;
            const int64_t r1 = r0 + block_rows < rows ? r0 + block_rows : rows;
            ++used;
            if (dtype == PO_F32)
                pool.emplace_back(format_rows<float>, (const float*)h_data, ld, cols, r0, r1, std::ref(cur[(size_t)t]));
            else
                pool.emplace_back(format_rows<double>, (const double*)h_data, ld, cols, r0, r1, std::ref(cur[(size_t)t]));
        }
        for (auto& th : pool) th.join();
        if (writer.joinable()) writer.join();  // the previous wave is on its way to the file
        if (write_failed) break;
        writer = std::thread([&cur, used, fh, &write_failed, &write_errno]() {
            for (int t = 0; t < used && !write_failed; ++t) {
                const std::vector<char>& b = cur[(size_t)t];
                if (!b.empty() && fwrite(b.data(), 1, b.size(), fh) != b.size()) {
                    write_failed = true;
                    write_errno = errno;
                }
            }
        });
    }
    if (writer.joinable()) writer.join();
    if (write_failed) {
        set_error("po_savetxt_host: write to %s failed: %s", path, strerror(write_errno));
        rc = PO_ERR_ARG;
    }
    if (fclose(fh) != 0 && rc == PO_OK) {
        set_error("po_savetxt_host: closing %s failed: %s", path, strerror(errno));
        rc = PO_ERR_ARG;
    }
    return rc;
}
