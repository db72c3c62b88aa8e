// micro-benchmark (B200): which operations share the shared-memory data pipe of an SM?
//   1. LDS.32 alone, SHFL alone, both interleaved: does a shuffle cost a shared-memory wavefront?
//   2. STG.256 with the H phase's address pattern (32 B per lane, lanes 64 B apart, two rows) against the same bytes
//      written contiguously (lane i writes bytes 32 i ..): cycles per instruction while the lines stay in L2.
//   3. the same flow rows through shared memory and a bulk (TMA) store, next to a stream of LDS.128.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/pipe_bench.cu -o scratch/pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode bit 0: LDS.32, bit 1: SHFL, bit 2: LDS.128
__global__ void k_pipe(int mode, int iters, unsigned long long *out, int *sink)
{
    extern __shared__ uint4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const uint32_t a32 = smem_u32(reinterpret_cast<uint32_t *>(sm) + threadIdx.x);
    const uint32_t a128 = smem_u32(sm + (threadIdx.x & 255));
    unsigned acc = threadIdx.x, shv[4] = {threadIdx.x * 7, threadIdx.x * 5, threadIdx.x * 3, threadIdx.x};
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (mode & 1) {
                unsigned v;
                asm volatile("{\n\t.reg .u32 a;\n\tadd.u32 a, %1, %2;\n\tld.shared.u32 %0, [a];\n\t}" : "=r"(v) : "r"(a32), "r"(u * 2048) : "memory");
                acc += v;
            }
            if (mode & 2) shv[u & 3] = __shfl_xor_sync(0xffffffffu, shv[u & 3], 1 + (u & 3)) + 1;
            if (mode & 4) {
                uint4 v;
                asm volatile("{\n\t.reg .u32 a;\n\tadd.u32 a, %4, %5;\n\tld.shared.v4.u32 {%0,%1,%2,%3}, [a];\n\t}" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a128), "r"(u * 4096) : "memory");
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    unsigned sh = shv[0] + shv[1] + shv[2] + shv[3];
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc + sh == 0x12345) *sink = acc;
}

// pattern 0: the H phase (lane = 8-pixel segment: 32 B at 64 B stride, lanes 16-31 one row (pitch) further, second half
// in a second instruction); 1: contiguous (lane i writes 32 B at 32 i, 1 KB per instruction); each warp owns a 64 KB slice
// of a per-SM region that stays in L2.
__global__ void k_stg(int pattern, int iters, float *buf, unsigned long long *out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *base = buf + ((size_t)blockIdx.x * (blockDim.x / 32) + warp) * 16384; // 64 KB per warp
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int slot = (it & 15) * 1024; // floats: 4 KB per iteration, wraps inside the slice
#pragma unroll
        for (int half = 0; half < 2; half++) {
            size_t off;
            if (pattern == 0) off = slot + (lane >> 4) * 512 + (lane & 15) * 16 + half * 8;
            else off = slot + half * 256 + lane * 8;
            const float v = (float)it;
            asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(base + off), "f"(v) : "memory");
        }
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// mode 0: LDS.128 stream only.  1: + every 32 loads the warp-group's flow rows go out with STG.256 (H-phase pattern).
// 2: + the same bytes are first written to shared memory (STS.128) and then leave with one bulk store per 960-byte row.
__global__ void k_tma(int mode, int iters, float *buf, unsigned long long *out, int *sink)
{
    extern __shared__ uint4 sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    uint4 *stage = sm + 2048; // 8 rows x 960 B = 7680 B = 480 chunks
    const int lane = threadIdx.x & 31;
    const uint32_t a128 = smem_u32(sm + threadIdx.x);
    const uint32_t a32 = smem_u32(reinterpret_cast<uint32_t *>(sm) + threadIdx.x);
    float *base = buf + (size_t)blockIdx.x * 65536; // 256 KB per CTA
    const int hi = threadIdx.x >> 4, hseg = threadIdx.x & 15; // 128 threads: 8 rows x 16 segments (15 live)
    unsigned acc = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 24; u++) {
            unsigned v;
            asm volatile("{\n\t.reg .u32 a;\n\tadd.u32 a, %1, %2;\n\tld.shared.u32 %0, [a];\n\t}" : "=r"(v) : "r"(a128), "r"(u * 512 + 4) : "memory");
            acc += v;
        }
#pragma unroll
        for (int u = 0; u < 40; u++) {
            asm volatile("{\n\t.reg .u32 a;\n\tadd.u32 a, %0, %1;\n\tst.shared.u32 [a], %2;\n\t}" ::"r"(a32), "r"(u * 512), "r"(acc) : "memory");
        }
        float *rows = base + (it & 7) * 8 * 1920; // 8 rows of 240 floats at pitch 1920
        if (mode >= 2 && (threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 20; u++) {
            uint4 v;
            asm volatile("{\n\t.reg .u32 a;\n\tadd.u32 a, %4, %5;\n\tld.shared.v4.u32 {%0,%1,%2,%3}, [a];\n\t}" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a128), "r"(u * 1024 + 16) : "memory");
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        if (mode == 1 && hseg < 15) {
            const float v = (float)acc;
#pragma unroll
            for (int half = 0; half < 2; half++)
                asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(rows + hi * 1920 + hseg * 16 + half * 8), "f"(v) : "memory");
        }
        if (mode >= 2) {
            if (hseg < 15) {
                const uint4 v = make_uint4(acc, acc, acc, acc);
#pragma unroll
                for (int q = 0; q < 4; q++) stage[hi * 60 + hseg * 4 + ((q + hseg) & 3)] = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (mode == 2 && threadIdx.x == 0) {
#pragma unroll
            for (int r = 0; r < 8; r++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 960;" ::"l"(rows + r * 1920),
                             "r"(smem_u32(stage + r * 60))
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (mode == 3 && (threadIdx.x & 31) == 0) { // one issuing lane per warp, two rows each
            const int w = threadIdx.x >> 5;
#pragma unroll
            for (int r = 0; r < 2; r++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 960;" ::"l"(rows + (2 * w + r) * 1920),
                             "r"(smem_u32(stage + (2 * w + r) * 60))
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (mode >= 2 && (threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345 + lane) *sink = acc;
}

static double avg148(unsigned long long *d)
{
    unsigned long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += h[i];
    return avg / 148;
}

int main()
{
    unsigned long long *d;
    int *s;
    float *buf;
    cudaMalloc(&d, 1024 * 8);
    cudaMalloc(&s, 4);
    cudaMalloc(&buf, (size_t)148 * 16 * 65536 * 4);
    cudaFuncSetAttribute(k_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    const int iters = 2000, threads = 512;
    const char *names[8] = {"", "LDS.32", "SHFL", "LDS.32 + SHFL", "LDS.128", "LDS.32 + LDS.128", "SHFL + LDS.128", "all three"};
    for (int m = 1; m <= 7; m++) {
        for (int r = 0; r < 2; r++) {
            k_pipe<<<148, threads, 65536>>>(m, iters, d, s);
            cudaDeviceSynchronize();
        }
        printf("pipe %-18s: %.2f cycles per unrolled step per warp (per SM, %d warps)\n", names[m], avg148(d) / ((double)iters * 8 * (threads / 32)), threads / 32);
    }
    for (int p = 0; p < 2; p++) {
        for (int r = 0; r < 2; r++) {
            k_stg<<<148, threads>>>(p, iters, buf, d);
            cudaDeviceSynchronize();
        }
        printf("stg pattern %d (%s): %.2f cycles per warp-level STG.256 (per SM)\n", p, p ? "contiguous" : "H phase", avg148(d) / ((double)iters * 2 * (threads / 32)));
    }
    for (int m = 0; m < 4; m++) {
        for (int r = 0; r < 2; r++) {
            k_tma<<<148 * 4, 128, 49152>>>(m, iters, buf, d, s);
            cudaDeviceSynchronize();
        }
        printf("tma mode %d (%s): %.1f cycles per iteration of 20 LDS.128 + 24 LDS.32 + 40 STS.32 per thread, 4 CTAs of 128 per SM\n", m,
               m == 0 ? "loads only" : m == 1 ? "+ STG.256 rows" : m == 2 ? "+ STS.128 + 8 bulk stores by thread 0" : "+ STS.128 + 2 bulk stores per warp", avg148(d) / (double)iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
