// micro-benchmark: shared-memory wavefronts per LDS.128 for several address patterns (is a 16-byte chunk that two lanes
// of a warp read in the same instruction fetched once?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int pattern, int iters, unsigned long long *out, int *sink)
{
    extern __shared__ uint4 sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    int base;
    switch (pattern) {
    case 0: base = lane; break;                 // 32 distinct contiguous chunks
    case 1: base = lane >> 1; break;            // adjacent lane pairs share a chunk (16 distinct)
    case 2: base = lane >> 2; break;            // groups of four share (8 distinct)
    case 3: base = 0; break;                    // all lanes the same chunk
    case 4: base = 2 * lane; break;             // H phase today: lane = segment, chunk 2*seg + k (stride 32 B)
    case 5: base = 2 * (lane >> 1); break;      // pairs share, stride 32 B between pairs
    case 6: base = (lane & 15); break;          // half-warps share (lanes l and l+16 the same chunk)
    default: base = lane; break;
    }
    unsigned acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint4 v;
            const unsigned addr = (unsigned)__cvta_generic_to_shared(sm + ((base + u * 64 + (it & 3) * 512) & 4095));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
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
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 2000, threads = 512;
    for (int p = 0; p <= 6; p++) {
        k<<<148, threads, 65536>>>(p, iters, d, s);
        cudaDeviceSynchronize();
        k<<<148, threads, 65536>>>(p, iters, d, s);
        cudaDeviceSynchronize();
        unsigned long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; i++) avg += h[i];
        avg /= 148;
        const double instr = (double)iters * 8 * (threads / 32);
        printf("pattern %d: %.2f cycles per warp-level LDS.128 (per SM)\n", p, avg / instr);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
