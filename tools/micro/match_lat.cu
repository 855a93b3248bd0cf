// Microbenchmark: latency / throughput of MATCH.ANY vs number of distinct values, on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(unsigned distinct, unsigned long long* out, unsigned* sink)
{
    unsigned lane = threadIdx.x & 31;
    unsigned v = lane % distinct + 1000;
    unsigned acc = 0;
    // dependent chain: latency
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 256; i++) { unsigned m = __match_any_sync(0xffffffffu, v); v = (v ^ (m & 0)) ; acc += m; v += (m >> 31) * 0; }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / 256;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void thr(unsigned distinct, unsigned long long* out, unsigned* sink)
{
    unsigned lane = threadIdx.x & 31;
    unsigned v0 = lane % distinct + 1000, v1 = v0 + 7, v2 = v0 + 13, v3 = v0 + 29;
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) {
        unsigned a = __match_any_sync(0xffffffffu, v0);
        unsigned b = __match_any_sync(0xffffffffu, v1);
        unsigned c = __match_any_sync(0xffffffffu, v2);
        unsigned d = __match_any_sync(0xffffffffu, v3);
        acc += a + b + c + d;
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / 256;   // cycles per MATCH per warp with blockDim/32 warps competing
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void ballots(unsigned long long* out, unsigned* sink)
{
    unsigned lane = threadIdx.x & 31; unsigned v = lane * 2654435761u; unsigned acc = 0;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 256; i++) {
        unsigned m = 0xffffffffu;
        #pragma unroll
        for (int b = 0; b < 12; b++) { unsigned bal = __ballot_sync(0xffffffffu, (v >> b) & 1); m &= ((v >> b) & 1) ? bal : ~bal; }
        acc += m; v += m & 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / 256;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    unsigned long long* out; unsigned* sink;
    cudaMallocManaged(&out, 8); cudaMalloc(&sink, 4 * 1024 * 1024);
    unsigned ds[] = {1, 2, 4, 8, 16, 32};
    for (unsigned d : ds) {
        lat<<<1, 32>>>(d, out, sink); cudaDeviceSynchronize();
        printf("match.any latency, 1 warp, %2u distinct: %llu cycles\n", d, out[0]);
    }
    for (unsigned d : ds) for (int warps : {4, 11, 16}) {
        thr<<<148, 32 * warps>>>(d, out, sink); cudaDeviceSynchronize();
        printf("match.any throughput, %2d warps/SM, %2u distinct: %llu cycles per MATCH per warp\n", warps, d, out[0]);
    }
    ballots<<<1, 32>>>(out, sink); cudaDeviceSynchronize();
    printf("12-ballot software match, 1 warp: %llu cycles\n", out[0]);
    return 0;
}
