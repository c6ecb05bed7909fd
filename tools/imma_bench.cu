// Microbenchmark (dev tool): does legacy mma.sync int8 (IMMA) run at a useful rate on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ACC>
__global__ void k(int iters, int *out, uint32_t seed)
{
        uint32_t a[4], b[2];
        int d[ACC][4];
        for (int i = 0; i < 4; i++) a[i] = seed * (threadIdx.x + i + 1);
        for (int i = 0; i < 2; i++) b[i] = seed ^ (threadIdx.x * 7 + i);
        for (int j = 0; j < ACC; j++) for (int i = 0; i < 4; i++) d[j][i] = 0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int j = 0; j < ACC; j++) imma(d[j], a, b);
        }
        int s = 0;
        for (int j = 0; j < ACC; j++) for (int i = 0; i < 4; i++) s += d[j][i];
        if (s == 0x7fffffff) out[0] = s;
}

int main()
{
        int *out; cudaMalloc(&out, 64);
        const int iters = 4096, ACC = 8;
        for (int warps : {4, 8, 16, 32}) {
                int blocks = 148 * 2;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k<ACC><<<blocks, warps * 32>>>(iters, out, 12345u);
                cudaEventRecord(e0);
                k<ACC><<<blocks, warps * 32>>>(iters, out, 12345u);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                double macs = (double)blocks * warps * iters * ACC * 16 * 8 * 32;
                printf("warps/block %2d: %.3f ms  %.1f int8 TMAC/s (%.1f TOPS)\n", warps, ms, macs / ms / 1e9, 2 * macs / ms / 1e9);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
}
