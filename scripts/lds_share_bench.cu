// micro-benchmark (B200): the H phase's LDS.128 patterns.  Today lane = segment s reads chunks 2s+k (k = 0..3), every
// column sum twice per row; candidate: adjacent lanes read the SAME chunk in the same instruction (even lanes in the order
// k = 2,3,0,1, odd lanes 0,1,2,3), which the hardware may serve with half the wavefronts.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/lds_share_bench.cu -o scratch/lds_share_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int swz_old(int c) { return c ^ ((c >> 3) & 1); }
__device__ __forceinline__ int swz_new(int c) { return c ^ ((c >> 3) & 3); }
__global__ void k(int pattern, int iters, unsigned long long *out, int *sink)
{
    extern __shared__ uint4 sm[]; // rows of 32 chunks (128 columns x 4 B)
    for (int i = threadIdx.x; i < 2560; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const int lane = threadIdx.x & 31, seg = lane & 15, row = (threadIdx.x >> 4) & 7;
    unsigned addr[4];
    for (int j = 0; j < 4; j++) {
        int chunk;
        switch (pattern) {
        case 0: chunk = swz_old(2 * seg + j); break;                                       // today
        case 1: chunk = swz_new(2 * seg + ((seg & 1) ? j : (j + 2) & 3)); break;           // shared, new swizzle
        case 2: chunk = swz_old(2 * seg + ((seg & 1) ? j : (j + 2) & 3)); break;           // shared, old swizzle
        case 3: chunk = 2 * seg + ((seg & 1) ? j : (j + 2) & 3); break;                    // shared, no swizzle
        case 4: chunk = swz_new(2 * seg + j); break;                                       // today's order, new swizzle
        default: chunk = seg; break;
        }
        if (chunk > 31) chunk = 31;
        addr[j] = (unsigned)__cvta_generic_to_shared(sm + row * 32 + chunk);
    }
    unsigned acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 5; q++) // five planes, 4 KB apart
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint4 v;
                asm volatile("{\n\t.reg .u32 a;\n\tadd.u32 a, %4, %5;\n\tld.shared.v4.u32 {%0,%1,%2,%3}, [a];\n\t}" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr[j]), "r"(q * 4096 + (it & 7) * 512) : "memory");
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345) *sink = acc;
}
int main()
{
    unsigned long long *d;
    int *s;
    cudaMalloc(&d, 1024 * 8);
    cudaMalloc(&s, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    const int iters = 2000, threads = 128;
    const char *names[5] = {"today (lane = segment, chunks 2s+k)", "pair-shared, swizzle c^((c>>3)&3)", "pair-shared, today's swizzle", "pair-shared, no swizzle", "today's order, new swizzle"};
    for (int p = 0; p < 5; p++) {
        for (int r = 0; r < 2; r++) {
            k<<<148 * 4, threads, 40960>>>(p, iters, d, s);
            cudaDeviceSynchronize();
        }
        unsigned long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; i++) avg += h[i];
        avg /= 148;
        printf("%-40s: %.2f cycles per warp-level LDS.128 (4 CTAs of 128 threads per SM)\n", names[p], avg / ((double)iters * 20 * 16));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
