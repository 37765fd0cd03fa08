// Shared helpers for the s3b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define S3_OK 0
#define S3_ERR_INVALID (-1)
#define S3_ERR_CUDA (-2)
#define S3_ERR_UNSUPPORTED (-3)

namespace s3 {

void set_error(const char* fmt, ...);
void note_launch(int n);
// keep stream-ordered scratch memory cached in the pool across synchronisations (default: released)
void configure_mempool_once();

#define S3_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) {                                                        \
            s3::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                          cudaGetErrorString(_e));                                      \
            return S3_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define S3_REQUIRE(cond, ...)                                                           \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            s3::set_error(__VA_ARGS__);                                                 \
            return S3_ERR_INVALID;                                                      \
        }                                                                               \
    } while (0)

#define S3_TRY(expr)                                                                    \
    do {                                                                                \
        int _rc = (expr);                                                               \
        if (_rc != S3_OK) return _rc;                                                   \
    } while (0)

#define S3_LAUNCH_CHECK() S3_CUDA(cudaGetLastError())

constexpr int kNumSMs = 148;  // B200

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Stream-ordered scratch allocation (cudaMallocAsync on the caller's stream).
struct Scratch {
    cudaStream_t stream;
    void* ptrs[32];
    int n = 0;
    explicit Scratch(cudaStream_t s) : stream(s) { configure_mempool_once(); }
    ~Scratch() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], stream);
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e != cudaSuccess) {
            set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return S3_ERR_CUDA;
        }
        if (n < 32) ptrs[n++] = p;
        *out = reinterpret_cast<T*>(p);
        return S3_OK;
    }
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ double shfl_up_d(double v, int delta) {
    return __shfl_up_sync(0xffffffffu, v, delta);
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// Order-preserving map fp64 -> u64 (ascending doubles map to ascending unsigned keys).
__device__ __forceinline__ uint64_t f64_to_ordered(double v) {
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

}  // namespace s3
