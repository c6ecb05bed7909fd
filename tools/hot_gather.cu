// Microbenchmark (dev tool): do L2 eviction-priority hints keep a hot set resident under a
// stream of cold random gathers?  30% of the 64 B gathers go to a hot prefix of the table.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
        return x;
}

template <int MODE, int U>
__global__ void k(const uint4 *__restrict__ tab, uint64_t nrec, uint64_t hot, uint32_t hot_permille,
                  uint64_t per_group, uint4 *out)
{
        constexpr int LANES = 4;
        uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
        int sub = threadIdx.x % LANES;
        uint64_t pol_last, pol_first;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (uint64_t i = 0; i < per_group; i += U) {
                uint4 v[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                        uint64_t h = mix(gid * per_group + i + u);
                        bool is_hot = (uint32_t)(h >> 54) < hot_permille;   // threshold on 10 bits
                        uint64_t r = is_hot ? (((h & 0xffffffffull) * hot) >> 32) : (h & (nrec - 1));
                        const uint4 *p = tab + r * LANES + sub;
                        if (MODE == 0) {
                                v[u] = __ldg(p);
                        } else {
                                uint64_t pol = is_hot ? pol_last : pol_first;
                                asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p), "l"(pol));
                        }
                }
#pragma unroll
                for (int u = 0; u < U; u++) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
        }
        if (acc.x == 0x12345678u && acc.y == 42) out[0] = acc;
}

template <int MODE>
void run(const uint4 *tab, uint64_t nrec, uint64_t hot, uint32_t permille, uint4 *out)
{
        uint64_t groups = 148ull * 2048 / 4 * 8, per_group = 256, threads = groups * 4;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        k<MODE, 8><<<(unsigned)(threads / 256), 256>>>(tab, nrec, hot, permille, per_group, out);
        cudaEventRecord(a);
        for (int rep = 0; rep < 5; rep++) k<MODE, 8><<<(unsigned)(threads / 256), 256>>>(tab, nrec, hot, permille, per_group, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
        double g = (double)groups * per_group;
        printf("mode %d hot=%llu rows (%.0f MB) hot share %.0f%%: %.3f ms  %.2f G gathers/s\n", MODE,
               (unsigned long long)hot, hot * 64 / 1e6, permille / 10.24, ms, g / ms / 1e6);
}

int main(int argc, char **argv)
{
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, 0);
        printf("L2 %d MB, persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB\n", prop.l2CacheSize >> 20,
               prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
        if (argc > 1) {
                size_t want = (size_t)atoi(argv[1]) << 20;
                cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
                size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
                printf("persisting L2 set-aside: asked %zu MB -> %s, now %zu MB\n", want >> 20, cudaGetErrorString(e), got >> 20);
        }
        uint64_t nrec = 1ull << 26;            // 64M rows x 64 B = 4 GB
        uint4 *tab, *out;
        cudaMalloc(&tab, nrec * 64); cudaMalloc(&out, 64);
        cudaMemset(tab, 1, nrec * 64);
        for (uint64_t hot : {500000ull, 1000000ull, 1500000ull}) {
                for (uint32_t pm : {0u, 307u, 512u}) {
                        run<0>(tab, nrec, hot, pm, out);
                        run<1>(tab, nrec, hot, pm, out);
                }
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
}
