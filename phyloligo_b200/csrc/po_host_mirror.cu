// Host half of "ship the upper triangle only" (no device code).
//
// The distance matrix is symmetric: the block left of the diagonal in the rows below a panel is the
// transpose of the panel's part right of the diagonal.  The reference writes every entry of the
// N x N result (np.memmap rows, bin/phyloligo.py:413-425; the in-RAM matrix of :536-553); over PCIe
// the 4 N^2 bytes of that result are what bounds the end-to-end step (40 GB at 52 GB/s against
// 0.5 s of kernels at 100 000 contigs).  So only the part on and right of the diagonal (plus a
// tunable share of the rest) crosses the link, and the host builds the remainder from what has
// already arrived while the next panels compute and copy:
//   po_host_transpose_f32   one block, synchronous, on `threads` threads
//   po_host_mirror_*        a pool of host threads fed in stream order: a job submitted on a
//                           stream is released by a host callback (cudaLaunchHostFunc) once the
//                           copies enqueued before it have landed, and is cut into strips that the
//                           pool's threads take from an atomic counter
// Transposition kernel: strips of 64 source columns; inside, 16 x 8 source blocks are transposed
// in registers (AVX2 8 x 8 shuffles) and leave as full 64-byte lines with non-temporal stores (the
// destination is written once and not read again: no write-allocate traffic); the rows of the
// next 64-row tile are prefetched, since a strided walk over 64 pages defeats the hardware
// prefetcher.  Without AVX2 (or with a pitch that breaks the 32-byte alignment) the 4 x 4 SSE form
// runs instead.
#include <immintrin.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "po_common.cuh"
#include "po_host_threads.h"

namespace po {

constexpr int64_t MT = 64;  // strip width (source columns) and row-tile height

static void strip_sse(float* dst, int64_t ldd, const float* src, int64_t lds, int64_t rows, int64_t c0, int64_t c1) {
    for (int64_t r0 = 0; r0 < rows; r0 += MT) {
        const int64_t r1 = std::min(rows, r0 + MT);
        int64_t r = r0;
        for (; r + 4 <= r1; r += 4) {
            int64_t c = c0;
            for (; c + 4 <= c1; c += 4) {
                __m128 a0 = _mm_loadu_ps(src + (r + 0) * lds + c);
                __m128 a1 = _mm_loadu_ps(src + (r + 1) * lds + c);
                __m128 a2 = _mm_loadu_ps(src + (r + 2) * lds + c);
                __m128 a3 = _mm_loadu_ps(src + (r + 3) * lds + c);
                _MM_TRANSPOSE4_PS(a0, a1, a2, a3);
                _mm_storeu_ps(dst + (c + 0) * ldd + r, a0);
                _mm_storeu_ps(dst + (c + 1) * ldd + r, a1);
                _mm_storeu_ps(dst + (c + 2) * ldd + r, a2);
                _mm_storeu_ps(dst + (c + 3) * ldd + r, a3);
            }
            for (; c < c1; ++c)
                for (int64_t rr = r; rr < r + 4; ++rr) dst[c * ldd + rr] = src[rr * lds + c];
        }
        for (; r < r1; ++r)
            for (int64_t c = c0; c < c1; ++c) dst[c * ldd + r] = src[r * lds + c];
    }
}

__attribute__((target("avx2"))) static inline void transpose8x8(__m256 (&r)[8]) {
    const __m256 t0 = _mm256_unpacklo_ps(r[0], r[1]), t1 = _mm256_unpackhi_ps(r[0], r[1]);
    const __m256 t2 = _mm256_unpacklo_ps(r[2], r[3]), t3 = _mm256_unpackhi_ps(r[2], r[3]);
    const __m256 t4 = _mm256_unpacklo_ps(r[4], r[5]), t5 = _mm256_unpackhi_ps(r[4], r[5]);
    const __m256 t6 = _mm256_unpacklo_ps(r[6], r[7]), t7 = _mm256_unpackhi_ps(r[6], r[7]);
    const __m256 u0 = _mm256_shuffle_ps(t0, t2, 0x44), u1 = _mm256_shuffle_ps(t0, t2, 0xEE);
    const __m256 u2 = _mm256_shuffle_ps(t1, t3, 0x44), u3 = _mm256_shuffle_ps(t1, t3, 0xEE);
    const __m256 u4 = _mm256_shuffle_ps(t4, t6, 0x44), u5 = _mm256_shuffle_ps(t4, t6, 0xEE);
    const __m256 u6 = _mm256_shuffle_ps(t5, t7, 0x44), u7 = _mm256_shuffle_ps(t5, t7, 0xEE);
    r[0] = _mm256_permute2f128_ps(u0, u4, 0x20);
    r[1] = _mm256_permute2f128_ps(u1, u5, 0x20);
    r[2] = _mm256_permute2f128_ps(u2, u6, 0x20);
    r[3] = _mm256_permute2f128_ps(u3, u7, 0x20);
    r[4] = _mm256_permute2f128_ps(u0, u4, 0x31);
    r[5] = _mm256_permute2f128_ps(u1, u5, 0x31);
    r[6] = _mm256_permute2f128_ps(u2, u6, 0x31);
    r[7] = _mm256_permute2f128_ps(u3, u7, 0x31);
}

// dst[c][r] = src[r][c] for r in [0, rows), c in [c0, c1).  Requires ldd % 8 == 0 and r_al = the first
// row whose destination address is 32-byte aligned (the same for every c); rows below r_al and the
// ragged ends go through the SSE / scalar path.
template <bool NT>
__attribute__((target("avx2"))) static void strip_avx2(float* dst, int64_t ldd, const float* src, int64_t lds, int64_t rows,
                                                       int64_t c0, int64_t c1, int64_t r_al) {
    if (r_al > rows) r_al = rows;
    if (r_al > 0) strip_sse(dst, ldd, src, lds, r_al, c0, c1);
    const int64_t c8 = c0 + (c1 - c0) / 8 * 8;
    const int64_t full = r_al + (rows - r_al) / 16 * 16;  // rows [r_al, full) in blocks of 16
    for (int64_t r0 = r_al; r0 < full; r0 += MT) {
        const int64_t r1 = std::min(full, r0 + MT);
        const int64_t p1 = std::min(rows, r1 + MT);
        for (int64_t r = r1; r < p1; ++r)
            for (int64_t c = c0; c < c1; c += 16) _mm_prefetch((const char*)(src + r * lds + c), _MM_HINT_T0);
        for (int64_t r = r0; r < r1; r += 16) {
            for (int64_t c = c0; c < c8; c += 8) {
                __m256 x[8], y[8];
                for (int k = 0; k < 8; ++k) {
                    x[k] = _mm256_loadu_ps(src + (r + k) * lds + c);
                    y[k] = _mm256_loadu_ps(src + (r + 8 + k) * lds + c);
                }
                transpose8x8(x);
                transpose8x8(y);
                for (int k = 0; k < 8; ++k) {
                    float* d = dst + (c + k) * ldd + r;
                    if (NT) {
                        _mm256_stream_ps(d, x[k]);
                        _mm256_stream_ps(d + 8, y[k]);
                    } else {
                        _mm256_store_ps(d, x[k]);
                        _mm256_store_ps(d + 8, y[k]);
                    }
                }
            }
        }
    }
    if (NT) _mm_sfence();
    if (full < rows) strip_sse(dst + full, ldd, src + full * lds, lds, rows - full, c0, c8);
    if (c8 < c1) strip_sse(dst + r_al, ldd, src + r_al * lds, lds, rows - r_al, c8, c1);
}

static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
// PO_HOST_MIRROR_NT=0: ordinary stores instead of streaming stores (a tuning switch)
static bool use_nt_stores() {
    static const bool v = [] {
        const char* e = getenv("PO_HOST_MIRROR_NT");
        return !(e && e[0] == '0');
    }();
    return v;
}

// one strip of a block; picks the AVX2 form when the destination rows keep a common 32-byte phase
static void transpose_strip(float* dst, int64_t ldd, const float* src, int64_t lds, int64_t rows, int64_t c0, int64_t c1) {
    if (have_avx2() && ldd % 8 == 0 && ((uintptr_t)dst & 3) == 0 && rows >= 24) {
        // first row whose destination address is aligned (the same for every c): to a 64-byte line when
        // the pitch keeps that phase too, else to the 32 bytes the streaming stores need
        const uintptr_t a = (uintptr_t)(dst + c0 * ldd);
        const uintptr_t m = (ldd % 16 == 0) ? 63 : 31;
        const int64_t r_al = (int64_t)(((m + 1 - (a & m)) & m) / 4);
        if (use_nt_stores()) strip_avx2<true>(dst, ldd, src, lds, rows, c0, c1, r_al);
        else strip_avx2<false>(dst, ldd, src, lds, rows, c0, c1, r_al);
    } else {
        strip_sse(dst, ldd, src, lds, rows, c0, c1);
    }
}

// A block is cut into pieces of (at most job_rows source rows) x (64 source columns), walked row block by
// row block: the threads of the pool work inside one row block at a time.  Short row blocks matter -- with
// whole 4096-row panels as one piece per strip the pool's strided reads (one 256-byte run out of every
// 400 KB row) slowed the DMA that runs beside them from 52 to 39 GB/s, with 256-row blocks to 50 GB/s
// (profiles/r02e_sink_job_rows_probe.log).
struct MirrorJob {
    float* dst;
    int64_t ldd;
    const float* src;
    int64_t lds, rows, cols, job_rows, cstrips, nstrips;
    std::atomic<int64_t> next{0}, done{0};
};

static int64_t mirror_job_rows() {
    static const int64_t v = [] {
        const char* e = getenv("PO_HOST_MIRROR_JOB_ROWS");
        const long long r = e ? atoll(e) : 0;
        return (int64_t)(r >= 16 ? r : 256);
    }();
    return v;
}

struct MirrorPool {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<std::shared_ptr<MirrorJob>> ready;
    std::vector<std::thread> threads;
    int64_t submitted = 0, completed = 0;
    bool stop = false;

    void release(std::shared_ptr<MirrorJob> j) {
        {
            std::lock_guard<std::mutex> lk(mu);
            ready.push_back(std::move(j));
        }
        cv_work.notify_all();
    }
    void worker() {
        for (;;) {
            std::shared_ptr<MirrorJob> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !ready.empty(); });
                if (ready.empty()) return;  // stop requested and nothing left
                j = ready.front();
            }
            for (;;) {
                const int64_t s = j->next.fetch_add(1, std::memory_order_relaxed);
                if (s >= j->nstrips) break;
                const int64_t rb = s / j->cstrips, cs = s - rb * j->cstrips;
                const int64_t r0 = rb * j->job_rows, nr = std::min(j->job_rows, j->rows - r0);
                const int64_t c0 = cs * MT, c1 = std::min(j->cols, c0 + MT);
                transpose_strip(j->dst + r0, j->ldd, j->src + r0 * j->lds, j->lds, nr, c0, c1);
                if (j->done.fetch_add(1, std::memory_order_acq_rel) + 1 == j->nstrips) {
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        ++completed;
                    }
                    cv_done.notify_all();
                }
            }
            std::lock_guard<std::mutex> lk(mu);
            if (!ready.empty() && ready.front() == j) ready.pop_front();
        }
    }
};

// what the stream callback carries: the pool and the job it releases
struct MirrorTicket {
    MirrorPool* pool;
    std::shared_ptr<MirrorJob> job;
};

static void CUDART_CB mirror_release_cb(void* p) {
    MirrorTicket* t = static_cast<MirrorTicket*>(p);
    t->pool->release(std::move(t->job));
    delete t;
}

}  // namespace po

using namespace po;

extern "C" {

int po_host_transpose_f32(float* h_dst, int64_t ld_dst, const float* h_src, int64_t ld_src, int64_t rows,
                          int64_t cols, int threads) {
    if (rows < 0 || cols < 0 || ld_src < cols || ld_dst < rows) {
        set_error("po_host_transpose_f32: bad geometry");
        return PO_ERR_ARG;
    }
    if (rows == 0 || cols == 0) return PO_OK;
    if (!h_dst || !h_src) {
        set_error("po_host_transpose_f32: NULL pointer");
        return PO_ERR_ARG;
    }
    const int64_t nstrips = (cols + MT - 1) / MT;
    threads = pick_threads(threads, rows * cols * 4);
    if ((int64_t)threads > nstrips) threads = (int)nstrips;
    std::atomic<int64_t> next{0};
    run_threads(threads, [&](int) {
        for (;;) {
            const int64_t s = next.fetch_add(1, std::memory_order_relaxed);
            if (s >= nstrips) break;
            transpose_strip(h_dst, ld_dst, h_src, ld_src, rows, s * MT, std::min(cols, (s + 1) * MT));
        }
    });
    return PO_OK;
}

void* po_host_mirror_open(int threads) {
    if (threads <= 0) {
        threads = (int)std::thread::hardware_concurrency();
        if (threads <= 0) threads = 1;
    }
    if (threads > 256) threads = 256;
    MirrorPool* pool = new (std::nothrow) MirrorPool();
    if (!pool) {
        set_error("po_host_mirror_open: out of memory");
        return nullptr;
    }
    try {
        for (int t = 0; t < threads; ++t) pool->threads.emplace_back([pool] { pool->worker(); });
    } catch (...) {
        {
            std::lock_guard<std::mutex> lk(pool->mu);
            pool->stop = true;
        }
        pool->cv_work.notify_all();
        for (auto& th : pool->threads) th.join();
        delete pool;
        set_error("po_host_mirror_open: cannot start %d threads", threads);
        return nullptr;
    }
    return pool;
}

int po_host_mirror_submit(void* h_pool, po_stream_t stream, int after_stream, float* h_dst, int64_t ld_dst,
                          const float* h_src, int64_t ld_src, int64_t rows, int64_t cols) {
    MirrorPool* pool = static_cast<MirrorPool*>(h_pool);
    if (!pool) {
        set_error("po_host_mirror_submit: NULL pool");
        return PO_ERR_ARG;
    }
    if (rows < 0 || cols < 0 || ld_src < cols || ld_dst < rows) {
        set_error("po_host_mirror_submit: bad geometry");
        return PO_ERR_ARG;
    }
    if (rows == 0 || cols == 0) return PO_OK;
    if (!h_dst || !h_src) {
        set_error("po_host_mirror_submit: NULL pointer");
        return PO_ERR_ARG;
    }
    // nothing below may throw across the C boundary
    std::shared_ptr<MirrorJob> job;
    MirrorTicket* ticket = nullptr;
    try {
        job = std::make_shared<MirrorJob>();
        if (after_stream) ticket = new MirrorTicket{pool, job};
    } catch (...) {
        set_error("po_host_mirror_submit: out of memory");
        return PO_ERR_ARG;
    }
    job->dst = h_dst;
    job->ldd = ld_dst;
    job->src = h_src;
    job->lds = ld_src;
    job->rows = rows;
    job->cols = cols;
    job->job_rows = mirror_job_rows();
    job->cstrips = (cols + MT - 1) / MT;
    job->nstrips = job->cstrips * ((rows + job->job_rows - 1) / job->job_rows);
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        ++pool->submitted;
    }
    if (!after_stream) {
        pool->release(std::move(job));
        return PO_OK;
    }
    const cudaError_t e = cudaLaunchHostFunc(static_cast<cudaStream_t>(stream), mirror_release_cb, ticket);
    if (e != cudaSuccess) {
        delete ticket;
        {
            std::lock_guard<std::mutex> lk(pool->mu);
            --pool->submitted;
        }
        set_error("cudaLaunchHostFunc failed: %s", cudaGetErrorString(e));
        return PO_ERR_CUDA;
    }
    return PO_OK;
}

int po_host_mirror_wait(void* h_pool) {
    MirrorPool* pool = static_cast<MirrorPool*>(h_pool);
    if (!pool) {
        set_error("po_host_mirror_wait: NULL pool");
        return PO_ERR_ARG;
    }
    std::unique_lock<std::mutex> lk(pool->mu);
    pool->cv_done.wait(lk, [&] { return pool->completed == pool->submitted; });
    return PO_OK;
}

int po_host_mirror_close(void* h_pool) {
    MirrorPool* pool = static_cast<MirrorPool*>(h_pool);
    if (!pool) return PO_OK;
    {
        // jobs still waiting for their stream callback must run first: the callback dereferences the pool
        std::unique_lock<std::mutex> lk(pool->mu);
        pool->cv_done.wait(lk, [&] { return pool->completed == pool->submitted; });
        pool->stop = true;
    }
    pool->cv_work.notify_all();
    for (auto& th : pool->threads) th.join();
    delete pool;
    return PO_OK;
}

}  // extern "C"
