// Throughput/latency of shared-memory atomicExch (ATOMS.EXCH) with random addresses, and the order in
// which lanes of one warp instruction are served when they hit the same address.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void thr(int mode, unsigned long long* out, unsigned* sink)
{
    __shared__ unsigned tab[8192];
    unsigned lane = threadIdx.x & 31, acc = 0;
    unsigned x = threadIdx.x * 2654435761u + 12345u;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 512; i++) {
        x = x * 1664525u + 1013904223u;
        unsigned slot = (x >> 19);                        // 13 bits, random
        if (mode == 0) acc += atomicExch(&tab[slot], x);
        else if (mode == 1) { unsigned o = tab[slot]; __syncwarp(); tab[slot] = x; __syncwarp(); acc += o + tab[slot]; }
        else acc += tab[slot];
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / 512;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + lane;
}
__global__ void order(unsigned* hist)
{
    __shared__ unsigned tab[64];
    unsigned lane = threadIdx.x & 31;
    for (int it = 0; it < 1000; it++) {
        unsigned slot = (lane * 7 + it) % 5;              // 5 slots -> 6-7 lanes per slot
        if (lane < 5) tab[lane] = 1000;
        __syncwarp();
        unsigned old = atomicExch(&tab[slot], lane);
        __syncwarp();
        // ascending service order <=> every lane sees either 1000 or a LOWER lane of its group
        if (old != 1000 && old > lane) atomicAdd(&hist[1], 1u); else atomicAdd(&hist[0], 1u);
        __syncwarp();
    }
}
int main()
{
    unsigned long long* out; unsigned* sink; unsigned* hist;
    cudaMallocManaged(&out, 8); cudaMalloc(&sink, 4 << 20); cudaMallocManaged(&hist, 8); hist[0] = hist[1] = 0;
    const char* names[] = {"atomicExch", "load+store+load (2 syncs)", "load only"};
    for (int mode = 0; mode < 3; mode++) for (int warps : {1, 4, 12, 24}) {
        thr<<<148, 32 * warps>>>(mode, out, sink); cudaDeviceSynchronize();
        printf("%-28s %2d warps/SM: %4llu cycles per step per warp  (%.1f cycles per warp-step SM-wide)\n", names[mode], warps, out[0], (double)out[0] / warps);
    }
    order<<<1, 32>>>(hist); cudaDeviceSynchronize();
    printf("atomicExch same-address service order: ascending-consistent %u, violations %u\n", hist[0], hist[1]);
    return 0;
}
