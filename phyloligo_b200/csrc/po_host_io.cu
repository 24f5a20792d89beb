// Host-side writer of the reference's text matrix format (no device code).
//
// Replaces np.savetxt(path, M, delimiter="\t") at reference bin/phyloligo.py:1059-1066:
// every entry as "%.18e" (the float64 value; a float32 entry is widened first, as numpy
// does), fields separated by one tab, rows ended by '\n', no header.  Readers:
// bin/phyloligo_comparemat.py:7-14, bin/phyloselect.py:616-622, bin/phyloselect.R:324.
// numpy's writer is a per-entry Python loop (~1 us per entry); here row blocks are
// formatted by a pool of threads into private buffers and written in order.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>
#include "po_common.cuh"

namespace po {

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
    // a wave = `threads` row blocks of about 4 MB of text each, formatted in parallel, written in order
    const int64_t row_text = (cols > 0 ? cols : 1) * 26;
    int64_t block_rows = (4 << 20) / row_text;
    if (block_rows < 1) block_rows = 1;
    std::vector<std::vector<char>> bufs((size_t)threads);
    int rc = PO_OK;
    for (int64_t wave0 = 0; wave0 < rows && rc == PO_OK; wave0 += block_rows * threads) {
        std::vector<std::thread> pool;
        int used = 0;
        for (int t = 0; t < threads; ++t) {
            const int64_t r0 = wave0 + (int64_t)t * block_rows;
            if (r0 >= rows) break;
            const int64_t r1 = r0 + block_rows < rows ? r0 + block_rows : rows;
            ++used;
            if (dtype == PO_F32)
                pool.emplace_back(format_rows<float>, (const float*)h_data, ld, cols, r0, r1, std::ref(bufs[(size_t)t]));
            else
                pool.emplace_back(format_rows<double>, (const double*)h_data, ld, cols, r0, r1, std::ref(bufs[(size_t)t]));
        }
        for (auto& th : pool) th.join();
        for (int t = 0; t < used; ++t) {
            const std::vector<char>& b = bufs[(size_t)t];
            if (!b.empty() && fwrite(b.data(), 1, b.size(), fh) != b.size()) {
                set_error("po_savetxt_host: write to %s failed: %s", path, strerror(errno));
                rc = PO_ERR_ARG;
                break;
            }
        }
    }
    if (fclose(fh) != 0 && rc == PO_OK) {
        set_error("po_savetxt_host: closing %s failed: %s", path, strerror(errno));
        rc = PO_ERR_ARG;
    }
    return rc;
}
