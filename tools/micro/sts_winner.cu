// Which lane wins when several lanes of a warp store to the same shared-memory address (STS.U16)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* hist, unsigned seed)
{
    __shared__ unsigned short tab[4096];
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = seed + blockIdx.x * 7919u + warp * 104729u;
    for (int it = 0; it < 4096; it++) {
        x = x * 1664525u + 1013904223u;
        unsigned groups = 1u + (x >> 28);                  // 1..16 distinct slots in this batch
        unsigned y = (x ^ (lane * 2654435761u)) * 2246822519u;
        unsigned slot = (warp * 512u) + ((y >> 20) % groups) * 31u % 512u;
        __syncwarp();
        tab[slot] = (unsigned short)lane;
        __syncwarp();
        unsigned w = tab[slot];
        unsigned grp = __match_any_sync(0xffffffffu, slot);
        unsigned last = 31 - __clz(grp), first = __ffs(grp) - 1;
        if (lane == last) {   // one report per group
            if (w == last) atomicAdd(&hist[0], 1u); else if (w == first) atomicAdd(&hist[1], 1u); else atomicAdd(&hist[2], 1u);
        }
    }
}
int main()
{
    unsigned* h; cudaMallocManaged(&h, 16); h[0] = h[1] = h[2] = 0;
    k<<<148 * 4, 256>>>(h, 12345u); cudaDeviceSynchronize();
    printf("groups where the winner was: last lane %u, first lane %u, another lane %u\n", h[0], h[1], h[2]);
    return 0;
}
