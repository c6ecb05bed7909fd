// band_probe.cu -- dev tool (round-2 experiment): two numbers that decide whether a "row band resident in L2,
// entries sorted by column inside the band" form of the product can beat the random-line rate that bounds k_spmv
// on uniform matrices (profiles/r01_gather_granularity.txt: 46.5 G random 128-byte lines/s = 5.95 TB/s).
//
//   1. sweep:  64-byte gathers whose addresses INCREASE (a random subset of the rows of an 8 GB table, density d),
//              i.e. what the gathers of one band look like when its entries are sorted by column.  If the memory
//              system serves monotone-with-gaps traffic near its streaming rate, and both halves of a line are
//              used when both rows are wanted, the lines fetched per gather drop from 1 to (1 - exp(-2 lam)) / (2 lam) ...
//              the probe simply reports gathers/s per density.
//   1b. lockstep: the same with many sweepers advancing through the table TOGETHER, each taking one row per window --
//              the access pattern of a band whose rows are spread over all warps (accumulators in shared memory).
//   2. red:    throughput of red.global.add.u64 on an L2-resident accumulator block (rows of 16 u64 = 128 bytes,
//              random rows), i.e. the price of accumulating y in L2 instead of in registers.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
        return x;
}

// Warp w of the grid sweeps rows [w * span, (w + 1) * span) of the table in increasing order; row r is gathered
// when hash(r) < density.  4 lanes x 16 bytes per row, 8 groups per warp take 8 consecutive candidate rows.
__global__ void __launch_bounds__(256)
k_sweep(const uint4 *__restrict__ tab, uint64_t nrows, uint64_t span, uint32_t thresh, unsigned long long *count, uint4 *out)
{
        const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int lane = threadIdx.x & 31, g = lane >> 2, sub = lane & 3;
        uint64_t lo = warp * span, hi = lo + span;
        if (hi > nrows) hi = nrows;
        uint4 acc = make_uint4(0, 0, 0, 0);
        unsigned long long mine = 0;
        for (uint64_t r0 = lo; r0 < hi; r0 += 64) {
                uint4 v[8];
                bool take[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                        const uint64_t r = r0 + u * 8 + g;
                        take[u] = r < hi && (uint32_t)mix(r) < thresh;
                        v[u] = make_uint4(0, 0, 0, 0);
                        if (take[u]) v[u] = __ldg(tab + r * 4 + sub);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; mine += take[u] && sub == 0; }
        }
        if (mine) atomicAdd(count, mine);
        if (acc.x == 0x12345678u && acc.y == 42) out[0] = acc;
}

// "Lockstep" model of the banded product: S sweepers (4 lanes = one 64-byte row each) advance through the table
// together; in window k (S * gap consecutive rows) every sweeper gathers exactly one row, at a hashed position, so
// each sweeper's addresses increase, the union of all sweepers touches 1/gap of the rows, and the two rows of a
// 128-byte line are wanted by DIFFERENT sweepers at about the same time -- L2 has to provide the sharing.
// Nothing synchronises the sweepers (drift is part of what is measured).
__global__ void __launch_bounds__(256)
k_lockstep(const uint4 *__restrict__ tab, uint64_t nrows, uint64_t S, uint32_t gap, uint64_t steps, uint4 *out)
{
        const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
        const int sub = threadIdx.x & 3;
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (uint64_t k0 = 0; k0 < steps; k0 += 8) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                        const uint64_t k = k0 + u;
                        const uint64_t slot = (s + mix(k)) % S;                        // a permutation of the sweepers per window
                        uint64_t r = (k * S + slot) * gap + mix(k * S + slot) % gap;
                        if (r >= nrows) r = nrows - 1;
                        v[u] = __ldg(tab + r * 4 + sub);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
        }
        if (acc.x == 0x12345678u && acc.y == 42) out[0] = acc;
}

// every thread adds to one u64 of a random row (16 consecutive lanes = one 128-byte row)
__global__ void __launch_bounds__(256)
k_red(unsigned long long *acc, uint64_t rows_mask, uint64_t per_thread)
{
        const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const uint64_t grp = tid >> 4;
        const int col = threadIdx.x & 15;
        for (uint64_t i = 0; i < per_thread; i++) {
                const uint64_t r = mix(grp * per_thread + i) & rows_mask;
                asm volatile("red.global.add.u64 [%0], %1;" :: "l"(acc + r * 16 + col), "l"((unsigned long long)(i + 1)) : "memory");
        }
}

int main()
{
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        // ---- 1. monotone sweeps with gaps
        const size_t bytes = 8ull << 30;
        const uint64_t nrows = bytes / 64;
        uint4 *tab, *out; unsigned long long *count;
        CK(cudaMalloc(&tab, bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&count, 8));
        CK(cudaMemset(tab, 1, bytes));
        const uint64_t warps = 148ull * 64;                       // 64 warps per SM, each its own contiguous span
        const uint64_t span = (nrows + warps - 1) / warps;
        printf("monotone 64-byte gathers over an 8 GB table (%llu warps, each sweeping its own span upwards)\n", (unsigned long long)warps);
        const double dens[] = {1.0, 0.75, 0.5, 0.3, 0.15, 0.05};
        for (double d : dens) {
                const uint32_t thresh = d >= 1.0 ? 0xffffffffu : (uint32_t)(d * 4294967296.0);
                float best = 1e9f;
                unsigned long long h = 0;
                for (int rep = 0; rep < 3; rep++) {
                        CK(cudaMemset(count, 0, 8));
                        CK(cudaEventRecord(a));
                        k_sweep<<<(unsigned)(warps * 32 / 256), 256>>>(tab, nrows, span, thresh, count, out);
                        CK(cudaEventRecord(b));
                        CK(cudaEventSynchronize(b));
                        float ms; CK(cudaEventElapsedTime(&ms, a, b));
                        if (ms < best) best = ms;
                        CK(cudaMemcpy(&h, count, 8, cudaMemcpyDeviceToHost));
                }
                printf("  density %.2f: %.3f ms  %7.2f G gathers/s  (%.0f GB/s of useful 64 B rows; random gathers: 46.5 G/s)\n", d, best,
                       h / best / 1e6, h * 64.0 / best / 1e6);
        }
        // ---- 1b. lockstep sweepers
        {
                const uint64_t S = 148ull * 2048 / 4;                     // one sweeper per 4 resident threads
                printf("lockstep sweepers (%llu lane groups advance through the table together; 1 row in `gap` is gathered)\n", (unsigned long long)S);
                for (uint32_t gap : {1u, 2u, 3u, 4u, 8u}) {
                        const uint64_t steps = (nrows / gap / S) & ~7ull;
                        float best = 1e9f;
                        for (int rep = 0; rep < 3; rep++) {
                                CK(cudaEventRecord(a));
                                k_lockstep<<<(unsigned)(S * 4 / 256), 256>>>(tab, nrows, S, gap, steps, out);
                                CK(cudaEventRecord(b));
                                CK(cudaEventSynchronize(b));
                                float ms; CK(cudaEventElapsedTime(&ms, a, b));
                                if (ms < best) best = ms;
                        }
                        const double gathers = (double)steps * S;
                        printf("  gap %u: %.3f ms  %7.2f G gathers/s  (%.0f GB/s of useful rows; random gathers: 46.5 G/s)\n", gap, best, gathers / best / 1e6,
                               gathers * 64.0 / best / 1e6);
                }
        }
        CK(cudaFree(tab));
        // ---- 2. L2-resident u64 reductions
        printf("red.global.add.u64 on random 128-byte rows of an accumulator block\n");
        for (size_t mb : {16, 32, 64, 128}) {
                const uint64_t rows = (mb << 20) / 128;
                unsigned long long *acc;
                CK(cudaMalloc(&acc, rows * 128)); CK(cudaMemset(acc, 0, rows * 128));
                const uint64_t threads = 148ull * 2048 * 4, per = 64;
                float best = 1e9f;
                for (int rep = 0; rep < 3; rep++) {
                        CK(cudaEventRecord(a));
                        k_red<<<(unsigned)(threads / 256), 256>>>(acc, rows - 1, per);
                        CK(cudaEventRecord(b));
                        CK(cudaEventSynchronize(b));
                        float ms; CK(cudaEventElapsedTime(&ms, a, b));
                        if (ms < best) best = ms;
                }
                printf("  block %4zu MB: %.3f ms  %7.1f G reds/s  = %.2f G rows of 16 columns/s (a product needs 1.47 G per 30 ms)\n", mb, best,
                       threads * per / best / 1e6, threads * per / 16.0 / best / 1e6);
                CK(cudaFree(acc));
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
}
