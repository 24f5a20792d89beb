// Host side of the output path (no device code): where finished block rows of the distance
// matrix land.
//
// The reference's --large workers write their block row straight into the caller's mapping:
// output[s] = ... into an np.memmap of the N x N float32 file (bin/phyloligo.py:202-222,
// 413-425) or into the data region of the HDF5 file (:471-478).  Here the block rows come off
// the device by DMA, and the same mapping is the DMA target when the kernel lets us page-lock
// it (po_host_register on a tmpfs / anonymous mapping); otherwise they go through a pinned
// ring and a pool of host threads moves them (po_host_copy2d into the mapping, or
// po_host_pwrite2d through the page cache, which never takes a page fault).  Fresh pages of
// the output file are instantiated ahead of the copies by po_host_prefault.
// The host half of "ship the upper triangle only" (the part of the matrix left of the diagonal is
// the transpose of what has already arrived) lives in po_host_mirror.cu.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/vfs.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#include "po_common.cuh"
#include "po_host_threads.h"

#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
#ifndef MADV_POPULATE_READ
#define MADV_POPULATE_READ 22
#endif

using namespace po;

extern "C" {

// madvise(advice) over [h_ptr, h_ptr + bytes) on `threads` threads, 64 MB per call; falls back to touching
// every page (read or read-modify-write) on kernels without MADV_POPULATE_* (before 5.14)
static int populate_range(const char* who, void* h_ptr, int64_t bytes, int threads, int advice) {
    if (bytes < 0 || (bytes > 0 && !h_ptr)) {
        set_error("%s: bad arguments", who);
        return PO_ERR_ARG;
    }
    if (bytes == 0) return PO_OK;
    const long page = sysconf(_SC_PAGESIZE);
    uintptr_t lo = (uintptr_t)h_ptr & ~(uintptr_t)(page - 1);
    const uintptr_t hi = ((uintptr_t)h_ptr + (uintptr_t)bytes + page - 1) & ~(uintptr_t)(page - 1);
    threads = pick_threads(threads, (int64_t)(hi - lo) / 16);
    const uintptr_t span = hi - lo;
    std::atomic<int> failed{0};
    run_threads(threads, [&](int t) {
        uintptr_t a = lo + (span / page * t / threads) * page;
        const uintptr_t b = (t == threads - 1) ? hi : lo + (span / page * (t + 1) / threads) * page;
        // 64 MB at a time: a populate call holds mmap_lock for reading throughout
        const uintptr_t step = (uintptr_t)64 << 20;
        bool use_madvise = true;
        while (a < b) {
            const uintptr_t e = std::min(b, a + step);
            if (use_madvise && madvise((void*)a, e - a, advice) != 0) {
                if (errno == EINVAL || errno == ENOSYS) {
                    use_madvise = false;  // kernel older than 5.14: touch every page instead
                } else {
                    failed.store(errno);
                    return;
                }
            }
            if (!use_madvise) {
                for (uintptr_t q = a; q < e; q += page) {
                    volatile unsigned char* c = (volatile unsigned char*)q;
                    if (advice == MADV_POPULATE_WRITE) *c = *c;
                    else (void)*c;
                }
            }
            a = e;
        }
    });
    if (failed.load()) {
        set_error("%s: madvise failed: %s", who, strerror(failed.load()));
        return PO_ERR_ARG;
    }
    return PO_OK;
}

int po_host_prefault(void* h_ptr, int64_t bytes, int threads) {
    return populate_range("po_host_prefault", h_ptr, bytes, threads, MADV_POPULATE_WRITE);
}

int po_host_premap(int fd, void* h_ptr, int64_t bytes, int threads) {
    // A shared mapping of a tmpfs file needs no write notification: its page-table entries are writable from
    // the first (read) fault on, and a read fault maps the 16 pages around it in one go (fault-around) when
    // they exist and are up to date.  Populating for reading therefore maps an EXISTING file ten times
    // faster than populating for writing (one fault per page), and later writes fault no more.  On other file
    // systems (dirty tracking: entries would come up read-only) populate for writing as before.
    struct statfs sfs;
    const bool tmpfs = fd >= 0 && fstatfs(fd, &sfs) == 0 && (unsigned long)sfs.f_type == 0x01021994ul;  // TMPFS_MAGIC
    return populate_range("po_host_premap", h_ptr, bytes, threads, tmpfs ? MADV_POPULATE_READ : MADV_POPULATE_WRITE);
}

int po_host_register(void* h_ptr, int64_t bytes) {
    if (!h_ptr || bytes <= 0) {
        set_error("po_host_register: bad arguments");
        return PO_ERR_ARG;
    }
    PO_CUDA_CHECK(cudaHostRegister(h_ptr, (size_t)bytes, cudaHostRegisterPortable));
    return PO_OK;
}

int po_host_unregister(void* h_ptr) {
    if (!h_ptr) return PO_OK;
    PO_CUDA_CHECK(cudaHostUnregister(h_ptr));
    return PO_OK;
}

int po_host_copy2d(void* h_dst, int64_t dst_pitch, const void* h_src, int64_t src_pitch, int64_t width,
                   int64_t rows, int threads) {
    if (width < 0 || rows < 0 || dst_pitch < width || src_pitch < width) {
        set_error("po_host_copy2d: bad geometry");
        return PO_ERR_ARG;
    }
    if (width == 0 || rows == 0) return PO_OK;
    if (!h_dst || !h_src) {
        set_error("po_host_copy2d: NULL pointer");
        return PO_ERR_ARG;
    }
    threads = pick_threads(threads, width * rows);
    if ((int64_t)threads > rows) threads = (int)rows;
    run_threads(threads, [&](int t) {
        const int64_t r0 = rows * t / threads, r1 = rows * (t + 1) / threads;
        if (dst_pitch == width && src_pitch == width) {
            memcpy((char*)h_dst + r0 * width, (const char*)h_src + r0 * width, (size_t)((r1 - r0) * width));
            return;
        }
        for (int64_t r = r0; r < r1; ++r)
            memcpy((char*)h_dst + r * dst_pitch, (const char*)h_src + r * src_pitch, (size_t)width);
    });
    return PO_OK;
}

int po_host_pwrite2d(int fd, int64_t file_offset, int64_t file_pitch, const void* h_src, int64_t src_pitch,
                     int64_t width, int64_t rows, int threads) {
    if (fd < 0 || width < 0 || rows < 0 || file_pitch < width || src_pitch < width || file_offset < 0) {
        set_error("po_host_pwrite2d: bad arguments");
        return PO_ERR_ARG;
    }
    if (width == 0 || rows == 0) return PO_OK;
    if (!h_src) {
        set_error("po_host_pwrite2d: NULL pointer");
        return PO_ERR_ARG;
    }
    threads = pick_threads(threads, width * rows);
    if ((int64_t)threads > rows) threads = (int)rows;
    std::atomic<int> failed{0};
    run_threads(threads, [&](int t) {
        const int64_t r0 = rows * t / threads, r1 = rows * (t + 1) / threads;
        const bool dense = file_pitch == width && src_pitch == width;
        const int64_t pieces = dense ? 1 : r1 - r0;
        for (int64_t k = 0; k < pieces && !failed.load(std::memory_order_relaxed); ++k) {
            const int64_t r = r0 + k;
            const char* src = (const char*)h_src + r * src_pitch;
            int64_t off = file_offset + r * file_pitch;
            int64_t left = dense ? (r1 - r0) * width : width;
            while (left > 0) {
                const ssize_t w = pwrite(fd, src, (size_t)std::min<int64_t>(left, (int64_t)1 << 30), (off_t)off);
                if (w < 0) {
                    if (errno == EINTR) continue;
                    failed.store(errno);
                    return;
                }
                src += w;
                off += w;
                left -= w;
            }
        }
    });
    if (failed.load()) {
        set_error("po_host_pwrite2d: pwrite failed: %s", strerror(failed.load()));
        return PO_ERR_ARG;
    }
    return PO_OK;
}

int po_host_pread(int fd, int64_t file_offset, void* h_dst, int64_t bytes, int threads) {
    if (fd < 0 || file_offset < 0 || bytes < 0 || (bytes > 0 && !h_dst)) {
        set_error("po_host_pread: bad arguments");
        return PO_ERR_ARG;
    }
    if (bytes == 0) return PO_OK;
    threads = pick_threads(threads, bytes);
    std::atomic<int> failed{0};
    run_threads(threads, [&](int t) {
        int64_t a = bytes * t / threads;
        const int64_t b = bytes * (t + 1) / threads;
        while (a < b) {
            const ssize_t r = pread(fd, (char*)h_dst + a, (size_t)(b - a), (off_t)(file_offset + a));
            if (r < 0) {
                if (errno == EINTR) continue;
                failed.store(errno);
                return;
            }
            if (r == 0) {  // shorter file than announced
                failed.store(EIO);
                return;
            }
            a += r;
        }
    });
    if (failed.load()) {
        set_error("po_host_pread: pread failed: %s", strerror(failed.load()));
        return PO_ERR_ARG;
    }
    return PO_OK;
}

}  // extern "C"
