// gather_policy.cu -- dev tool (round-2 experiment): does ANY load flavour make a 64-byte gather from a table
// much larger than L2 cost less than a 128-byte line of DRAM traffic on B200?  k_spmv's floor on BASELINE
// config 4 is set by exactly that (profiles/r01_gather_granularity.txt: W = 16..128 B all run at ~46.5 G
// gathers/s, and ncu counts ~128 B of DRAM reads per gather).  Flavours: plain ld.global.nc, the PTX
// prefetch-size qualifiers .L2::64B / .L2::128B / .L2::256B, L1::no_allocate, .cg, .lu, an L2 evict_first /
// no_allocate cache policy, cp.async (16 B pieces to shared memory) and cp.async.bulk (one 64 B copy).
// Run plain for gathers/s, and under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:k_gather
// for bytes per gather (each launch does GROUPS * PER_GROUP gathers, printed below).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

enum Flavour { NC = 0, L2_64, L2_128, L2_256, NOALLOC_64, CG, LU, EVICT_FIRST, NO_ALLOCATE_POLICY, CP_ASYNC, BULK, NFLAV };
static const char *names[NFLAV] = {"ld.global.nc", "ld.global.nc.L2::64B", "ld.global.nc.L2::128B", "ld.global.nc.L2::256B",
                                   "ld.global.nc.L1::no_allocate.L2::64B", "ld.global.cg", "ld.global.lu",
                                   "ld.global.nc + L2::evict_first policy", "ld.global.nc + evict_first on 1/16, unchanged on the rest",
                                   "cp.async.cg 4 x 16 B -> smem", "cp.async.bulk 64 B -> smem"};

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
        return x;
}

template <int F> __device__ __forceinline__ uint4 load16(const uint4 *p, uint64_t pol)
{
        uint4 v;
        if (F == NC) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        else if (F == L2_64) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        else if (F == L2_128) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        else if (F == L2_256) asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        else if (F == NOALLOC_64) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        else if (F == CG) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        else if (F == LU) asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        else asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
        return v;
}

constexpr int LANES = 4;          // 4 x 16 B = one 64-byte record (a row of a block at n = 16)
constexpr int U = 8;              // gathers in flight per lane group

template <int F>
__global__ void __launch_bounds__(256)
k_gather(const uint4 *__restrict__ tab, uint64_t nrec, uint64_t per_group, uint4 *out)
{
        const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
        const int sub = threadIdx.x % LANES;
        uint64_t pol = 0;
        if (F == EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        if (F == NO_ALLOCATE_POLICY) asm volatile("createpolicy.fractional.L2::evict_first.L2::evict_unchanged.b64 %0, 0.0625;" : "=l"(pol));
        uint4 acc = make_uint4(0, 0, 0, 0);
        if (F == CP_ASYNC || F == BULK) {
                __shared__ __align__(128) uint4 stage[256 / LANES][U][LANES];        // 64 groups x 8 x 64 B = 32 KB
                __shared__ __align__(8) uint64_t bar[256 / LANES];
                const int g = threadIdx.x / LANES;
                const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[g]);
                if (F == BULK && sub == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
                __syncthreads();
                uint32_t phase = 0;
                for (uint64_t i = 0; i < per_group; i += U) {
                        if (F == CP_ASYNC) {
#pragma unroll
                                for (int u = 0; u < U; u++) {
                                        const uint64_t r = mix(gid * per_group + i + u) & (nrec - 1);
                                        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[g][u][sub]);
                                        asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" :: "r"(dst), "l"(tab + r * LANES + sub) : "memory");
                                }
                                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                                __syncwarp();
                        } else {
                                if (sub == 0) {
                                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(U * 64) : "memory");
#pragma unroll
                                        for (int u = 0; u < U; u++) {
                                                const uint64_t r = mix(gid * per_group + i + u) & (nrec - 1);
                                                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[g][u][0]);
                                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
                                                             :: "r"(dst), "l"(tab + r * LANES), "r"(bar_a) : "memory");
                                        }
                                }
                                uint32_t ok = 0;
                                while (!ok)
                                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                                     : "=r"(ok) : "r"(bar_a), "r"(phase) : "memory");
                                phase ^= 1;
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) { uint4 v = stage[g][u][sub]; acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w; }
                        __syncwarp();
                }
        } else {
                for (uint64_t i = 0; i < per_group; i += U) {
                        uint4 v[U];
#pragma unroll
                        for (int u = 0; u < U; u++) {
                                const uint64_t r = mix(gid * per_group + i + u) & (nrec - 1);      // nrec is a power of two
                                v[u] = load16<F>(tab + r * LANES + sub, pol);
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
                }
        }
        if (acc.x == 0x12345678u && acc.y == 42) out[0] = acc;
}

template <int F> void run(const uint4 *tab, size_t bytes, uint4 *out)
{
        const uint64_t nrec = bytes / 64;
        const uint64_t groups = 148ull * 2048 / LANES * 8, per_group = 256, threads = groups * LANES;
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        k_gather<F><<<(unsigned)(threads / 256), 256>>>(tab, nrec, per_group, out);
        CK(cudaEventRecord(a));
        for (int rep = 0; rep < 3; rep++) k_gather<F><<<(unsigned)(threads / 256), 256>>>(tab, nrec, per_group, out);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 3;
        const double g = (double)groups * per_group;
        printf("  %-52s %8.3f ms  %6.2f G gathers/s  (%.0f gathers per launch)\n", names[F], ms, g / ms / 1e6, g);
}

int main()
{
        const size_t bytes = 8ull << 30;
        uint4 *tab, *out;
        CK(cudaMalloc(&tab, bytes)); CK(cudaMalloc(&out, 64));
        CK(cudaMemset(tab, 1, bytes));
        printf("64-byte gathers, table %zu MB\n", bytes >> 20);
        run<NC>(tab, bytes, out); run<L2_64>(tab, bytes, out); run<L2_128>(tab, bytes, out); run<L2_256>(tab, bytes, out);
        run<NOALLOC_64>(tab, bytes, out); run<CG>(tab, bytes, out); run<LU>(tab, bytes, out);
        run<EVICT_FIRST>(tab, bytes, out); run<NO_ALLOCATE_POLICY>(tab, bytes, out);
        run<CP_ASYNC>(tab, bytes, out); run<BULK>(tab, bytes, out);
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
}
