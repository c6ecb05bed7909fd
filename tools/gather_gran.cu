// Microbenchmark (dev tool): cost of random W-byte gathers from a table much larger than L2.
// Tells the DRAM fill granularity: if 64 B gathers run at the same gathers/s as 128 B gathers,
// a 64 B miss costs a full 128 B line.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
        return x;
}

// each group of LANES lanes (16 B per lane) gathers one W = 16*LANES byte record
template <int LANES, int U>
__global__ void k_gather(const uint4 *__restrict__ tab, uint64_t nrec, uint64_t per_group, uint4 *out)
{
        uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
        int sub = threadIdx.x % LANES;
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (uint64_t i = 0; i < per_group; i += U) {
                uint4 v[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                        uint64_t r = mix(gid * per_group + i + u) & (nrec - 1);   // nrec is a power of two
                        v[u] = __ldg(tab + r * LANES + sub);
                }
#pragma unroll
                for (int u = 0; u < U; u++) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
        }
        if (acc.x == 0x12345678u && acc.y == 42) out[0] = acc;
}

template <int LANES>
void run(const uint4 *tab, size_t bytes, uint4 *out)
{
        const int W = 16 * LANES;
        uint64_t nrec = bytes / W;
        uint64_t groups = 148ull * 2048 / LANES * 8;       // plenty of threads
        uint64_t per_group = 256;
        uint64_t threads = groups * LANES;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        k_gather<LANES, 8><<<(unsigned)(threads / 256), 256>>>(tab, nrec, per_group, out);
        cudaEventRecord(a);
        for (int rep = 0; rep < 5; rep++) k_gather<LANES, 8><<<(unsigned)(threads / 256), 256>>>(tab, nrec, per_group, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
        double g = (double)groups * per_group;
        printf("W=%4d B: %.3f ms  %.2f G gathers/s  useful %.0f GB/s  (as 128B lines: %.0f GB/s)\n", W, ms, g / ms / 1e6,
               g * W / ms / 1e6, g * (W < 128 ? 128 : W) / ms / 1e6);
}

int main(int argc, char **argv)
{
        if (argc > 1) {
                cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
                printf("set limit %s -> %s\n", argv[1], cudaGetErrorString(e));
        }
        size_t lim = 0;
        cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity = %zu\n", lim);
        size_t maxb = 8ull << 30;
        uint4 *tab, *out;
        cudaMalloc(&tab, maxb); cudaMalloc(&out, 64);
        cudaMemset(tab, 1, maxb);
        for (size_t bytes : {32ull << 20, 64ull << 20, 256ull << 20, 1ull << 30, 8ull << 30}) {
                printf("--- table %zu MB\n", bytes >> 20);
                run<1>(tab, bytes, out);
                run<2>(tab, bytes, out);
                run<4>(tab, bytes, out);
                run<8>(tab, bytes, out);
                run<16>(tab, bytes, out);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
}
