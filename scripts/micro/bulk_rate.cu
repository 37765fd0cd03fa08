// Microbenchmark: how fast can one SM pull scattered rows of S bytes from L2/HBM into shared memory with 1-D bulk
// copies (cp.async.bulk.shared::cluster.global.mbarrier), issued by all 32 lanes of a producer warp, NSTAGE stages in
// flight?  Prints bytes/clk/SM for S = 256..4096 and for the LDG.128 path on the same row set.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_rate bulk_rate.cu && ./bulk_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// one CTA per SM; warp 0 = producer (32 lanes, one row each per stage), the other warps only wait (consumers read nothing)
template <int NSTAGE>
__global__ void bulk_kernel(const char* __restrict__ data, int64_t n_rows, int64_t pitch, int row_bytes, int iters,
                            long long* cycles, int* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[NSTAGE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
    __syncthreads();
    const long long t0 = clock64();
    uint32_t rng = blockIdx.x * 7919u + lane * 104729u + 1u;
    if (warp == 0) {
        for (int it = 0; it < iters; ++it) {
            const int s = it % NSTAGE;
            if (it >= NSTAGE) mbar_wait(&full[s], ((it / NSTAGE) - 1) & 1);       // stage free again (nobody consumes)
            if (lane == 0) mbar_expect(&full[s], 32u * row_bytes);
            __syncwarp();
            rng = rng * 1664525u + 1013904223u;
            const int64_t row = (int64_t)(rng >> 8) % n_rows;
            bulk_g2s(smem + ((size_t)s * 32 + lane) * row_bytes, data + row * pitch, row_bytes, &full[s]);
        }
        for (int it = iters > NSTAGE ? iters - NSTAGE : 0; it < iters; ++it) mbar_wait(&full[it % NSTAGE], (it / NSTAGE) & 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) { cycles[blockIdx.x] = clock64() - t0; sink[blockIdx.x] = smem[lane]; }
}

// LDG.128 on the same scattered rows: 8 warps, each lane a float4, rows of row_bytes swept 512 B per warp step
__global__ void ldg_kernel(const char* __restrict__ data, int64_t n_rows, int64_t pitch, int row_bytes, int iters,
                           long long* cycles, float* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t rng = blockIdx.x * 7919u + warp * 104729u + 1u;
    float4 acc = make_float4(0, 0, 0, 0);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        rng = rng * 1664525u + 1013904223u;
        const int64_t row = (int64_t)(rng >> 8) % n_rows;
        const char* p = data + row * pitch;
        for (int o = lane * 16; o < row_bytes; o += 512) {
            const float4 v = *reinterpret_cast<const float4*>(p + o);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (acc.x == 123.456f) sink[0] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
    const int sms = 148;
    long long* d_cycles; int* d_sink; float* d_fsink;
    cudaMalloc(&d_cycles, sms * sizeof(long long)); cudaMalloc(&d_sink, sms * sizeof(int)); cudaMalloc(&d_fsink, 4);
    for (int big = 0; big < 2; ++big) {
        const size_t bytes = big ? (size_t)4 << 30 : (size_t)64 << 20;             // HBM-resident vs L2-resident row set
        char* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
        for (int row_bytes : {256, 512, 1024, 2048, 4096}) {
            const int64_t pitch = 4096, n_rows = bytes / pitch;
            auto run = [&](auto kern, int nstage, const char* name) {
                const int iters = 2000;
                const size_t smem = (size_t)nstage * 32 * row_bytes;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                kern<<<sms, 64, smem>>>(d, n_rows, pitch, row_bytes, iters, d_cycles, d_sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
                long long h[sms]; cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
                printf("%s set=%s row=%4d B stages=%d: %.1f B/clk/SM, %.0f clk per row request\n", name, big ? "4GB" : "64MB",
                       row_bytes, nstage, 32.0 * row_bytes * iters / avg, avg / (32.0 * iters));
            };
            if ((size_t)2 * 32 * row_bytes <= 200 * 1024) run(bulk_kernel<2>, 2, "bulk");
            if ((size_t)4 * 32 * row_bytes <= 200 * 1024) run(bulk_kernel<4>, 4, "bulk");
            if ((size_t)8 * 32 * row_bytes <= 200 * 1024) run(bulk_kernel<8>, 8, "bulk");
            {
                const int iters = 4000;
                ldg_kernel<<<sms * 8, 256>>>(d, n_rows, pitch, row_bytes, iters, d_cycles, d_fsink);
                cudaDeviceSynchronize();
                long long h[sms]; cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
                printf("ldg  set=%s row=%4d B (8 CTAs x 8 warps per SM): %.1f B/clk/SM\n", big ? "4GB" : "64MB", row_bytes,
                       8.0 * 8 * row_bytes * iters / avg);
            }
        }
        cudaFree(d);
    }
    return 0;
}
