// Library-level entry points: error reporting, version, launch accounting.
#include <atomic>
#include <stdarg.h>
#include "common.cuh"
#include "../../include/s3b200.h"

namespace s3 {

static thread_local char g_error[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void configure_mempool_once() {
    static std::atomic<uint32_t> done_mask{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return;
    const uint32_t bit = 1u << dev;
    if (done_mask.load(std::memory_order_acquire) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done_mask.fetch_or(bit, std::memory_order_release);
}

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace s3

extern "C" {

const char* s3_last_error(void) { return s3::g_error; }

int s3_version(void) { return 100; }

int64_t s3_launch_count(void) { return s3::g_launches.load(std::memory_order_relaxed); }

int s3_copy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
                    int kind, void* stream) {
    S3_REQUIRE(dst && src, "s3_copy2d_async: NULL argument");
    S3_REQUIRE(width >= 0 && height >= 0 && dst_pitch >= width && src_pitch >= width, "s3_copy2d_async: bad geometry");
    S3_REQUIRE(kind == 0 || kind == 1, "s3_copy2d_async: kind must be 0 (host to device) or 1 (device to host)");
    if (width == 0 || height == 0) return S3_OK;
    S3_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)height,
                              kind == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return S3_OK;
}

}  // extern "C"
