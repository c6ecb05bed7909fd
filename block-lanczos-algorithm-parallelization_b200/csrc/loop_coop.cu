// loop_coop.cu -- the block-Lanczos loop (sequential/lanczos_modp.c:631-659) as ONE persistent cooperative kernel.
//
// For problems whose operators and vector blocks live in L2 (BASELINE configs 1-3: <= 26 MB) an iteration is not
// bound by bandwidth but by latency: as six kernel nodes of a CUDA graph it costs about 4 us per node (launch ramp,
// first dependent loads, tail), 24 / 41 / 73 us per iteration on configs 1 / 2 / 3.  Here the whole loop is one
// kernel: one thread block per SM stays resident, the phases of an iteration are separated by grid barriers
// (one reduction to L2 + polling, see grid_sync) instead of kernel boundaries, and up to `max_iters` iterations run without
// the host:
//
//     tmp <- S1 v | barrier | rows crossing tiles | barrier | Av <- S2 tmp | barrier | rows crossing tiles | barrier |
//     dots | barrier whose last arriver runs the n x n stage (small_body) | orthogonalize | barrier
//
// (the two fix-up phases and their barriers are skipped for operators in which no row crosses a tile border).
// The device code of every phase is the code of the stand-alone kernels (spmv_body.cuh, dense_body.cuh,
// small_body.cuh) instantiated with COH = 1: vectors written by other blocks earlier in the same launch are read
// through L2 (ld.global.cg), never through the non-coherent path.  The loop flags (DevSmall) work as in the chain of
// kernels: the n x n stage raises `halt` (no pivot, or iteration limit reached) and the loop ends before the next
// product, so v, tmp, Av, p are left exactly as the reference leaves them.
#include "blk_internal.cuh"
#include "spmv_body.cuh"
#include "dense_body.cuh"
#include "small_body.cuh"

namespace {

constexpr int LOOP_TB = 512;          // one block of 16 warps per SM (<= 128 registers per thread)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_flag(const int *p)
{
        int v;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
}

// Grid barriers.  bar[0] counts arrivals and only grows (barrier number k of a launch is complete when it reaches
// k * nblocks; zero at launch), bar[1] is the release word of the barriers that carry a serial section.
//
// grid_sync: every block's thread 0 adds 1 with release semantics (cumulative over the block's writes, which
// __syncthreads has ordered before it) and polls the counter with acquire loads: the last arrival IS the release, so
// the critical path is one reduction to L2 plus one poll round trip.  Everything the phases exchange is read at L2
// (ld.global.cg / volatile, see COH in spmv_body.cuh), so no L1 invalidation is needed after the barrier.
__device__ __forceinline__ void grid_sync(unsigned *bar, const unsigned nblocks, unsigned &gen)
{
        gen++;
        __syncthreads();
        if (threadIdx.x == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                const unsigned target = nblocks * gen;
                while ((int)(ld_acquire_u32(bar) - target) < 0) { }
        }
        __syncthreads();
}

// grid_sync_serial: the last block to arrive runs serial() (all of its threads) before anybody is released.
template <class F>
__device__ __forceinline__ void grid_sync_serial(unsigned *bar, const unsigned nblocks, unsigned &gen, F serial)
{
        __shared__ int s_last;
        gen++;
        __syncthreads();
        if (threadIdx.x == 0) {
                unsigned ticket;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(bar) : "memory");
                s_last = ticket == nblocks * gen - 1;
        }
        __syncthreads();
        if (s_last) {
                serial();
                __syncthreads();
                if (threadIdx.x == 0) st_release_u32(bar + 1, gen);
        } else if (threadIdx.x == 0) {
                while (ld_acquire_u32(bar + 1) != gen) { }
        }
        __syncthreads();
}

template <int L, int V, int FOLD>
__global__ void __launch_bounds__(LOOP_TB, 1)
k_loop(const LoopArgs a)
{
        constexpr int NP = L * V;
        constexpr int WPB = LOOP_TB / 32;
        extern __shared__ u32 sm_loop[];
        const ModP m = a.m;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const unsigned nblocks = gridDim.x;
        const int64_t gwarp = (int64_t)blockIdx.x * WPB + warp, nwarps = (int64_t)nblocks * WPB;
        const int64_t ggroup = ((int64_t)blockIdx.x * LOOP_TB + threadIdx.x) / L, ngroups = (int64_t)nblocks * LOOP_TB / L;
        const int sub = threadIdx.x % L;
        const PushTargets none;
        unsigned gen = 0;
        // optional phase clocks (block 0, SM cycles): [0] product 1, [1] its fix-up, [2] product 2, [3] its fix-up,
        // [4] dots + n x n stage, [5] orthogonalize -- each including the barrier that ends it
        const bool prof = a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        long long tprev = prof ? clock64() : 0;
        auto lap = [&](int i) {
                if (prof) {
                        const long long t = clock64();
                        a.prof[i] += (unsigned long long)(t - tprev);
                        tprev = t;
                }
        };

        for (int it = 0; it < a.max_iters; it++) {
                if (ld_flag(&a.state->halt)) break;               // uniform over the grid: written only inside a barrier

                // ---- tmp <- S1 v,  Av <- S2 tmp
#pragma unroll 1
                for (int which = 0; which < 2; which++) {
                        const LoopOp &op = which ? a.s2 : a.s1;
                        const u32 *x = which ? a.tmp : a.v;
                        u32 *y = which ? a.Av : a.tmp;
                        for (int64_t t = gwarp; t < op.ntiles; t += nwarps)
                                spmv_tile<L, V, FOLD, 0, 0, 1>(op.ent, op.chunk_row, op.whead, t, op.Q, op.rows, x, y, m, 0, 0, none,
                                                               nullptr, nullptr, lane);
                        grid_sync(a.bar, nblocks, gen);
                        lap(2 * which);
                        if (op.crossing) {
                                for (int64_t gid = ggroup; gid < op.ntiles; gid += ngroups)
                                        spmv_fix_row<L, V, 1>(op.tail_row, op.span, op.whead, gid, sub, 0, op.ntiles, y, m, none);
                                grid_sync(a.bar, nblocks, gen);
                                lap(2 * which + 1);
                        }
                }

                // ---- block dot products; the last block to arrive runs the n x n stage
                dots_block<NP, FOLD, 1, LOOP_TB>(a.N, a.v, a.Av, a.sums, m, blockIdx.x, nblocks);
                grid_sync_serial(a.bar, nblocks, gen, [&]() { small_body(a.n, NP, a.sums, a.mats, a.state, 0, m, sm_loop); });
                lap(4);

                // ---- orthogonalize (skipped when the n x n stage found no pivot: v stays, the loop has halted)
                if (ld_flag(&a.state->do_ortho)) {
                        u32 *C = sm_loop, *D = C + NP * NP, *Wm = D + NP * NP, *dm = Wm + NP * NP;
                        for (int e = threadIdx.x; e < NP * NP; e += LOOP_TB) {
                                C[e] = __ldcg(a.mats + MAT_C * NP * NP + e);
                                D[e] = __ldcg(a.mats + MAT_VTAVD * NP * NP + e);
                                Wm[e] = __ldcg(a.mats + MAT_WINV * NP * NP + e);
                        }
                        for (int e = threadIdx.x; e < NP; e += LOOP_TB) dm[e] = __ldcg(a.mats + MAT_D * NP * NP + e);
                        __syncthreads();
                        // one thread per row (JT = NP); whole warps iterate together
                        for (int64_t base = gwarp * 32; base < a.N; base += nwarps * 32)
                                ortho_slot<NP, NP, FOLD, 1>(base + lane, a.N, a.v, a.Av, a.p, a.v, a.p, C, D, Wm, dm, m);
                }
                grid_sync(a.bar, nblocks, gen);
                lap(5);
        }
}

template <int L, int V>
const void *loop_kernel(int fold_every)
{
        switch (fold_every) {
        case 0: return (const void *)k_loop<L, V, 0>;
        case 8: return (const void *)k_loop<L, V, 8>;
        default: return (const void *)k_loop<L, V, 2>;
        }
}

const void *loop_kernel_for(int np, int fold_every)
{
        switch (np) {
        case 1: return loop_kernel<1, 1>(fold_every);
        case 2: return loop_kernel<1, 2>(fold_every);
        case 4: return loop_kernel<1, 4>(fold_every);
        case 8: return loop_kernel<2, 4>(fold_every);
        case 16: return loop_kernel<4, 4>(fold_every);
        }
        return nullptr;
}

size_t loop_smem_bytes(int n, int np)
{
        size_t small = sizeof(u32) * small_smem_words(n), ortho = sizeof(u32) * ((size_t)3 * np * np + np);
        return small > ortho ? small : ortho;
}

}  // namespace

bool loop_coop_supported(int np) { return loop_kernel_for(np, 0) != nullptr; }

int loop_coop_grid(int n, int np, const ModP &m, std::string *why)
{
        const void *f = loop_kernel_for(np, m.fold_every);
        if (!f) { *why = "no persistent loop kernel for this n"; return 0; }
        int dev = 0, coop = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) {
                *why = "device does not support cooperative launches";
                return 0;
        }
        const size_t smem = loop_smem_bytes(n, np);
        if (smem > 48 * 1024 && cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                *why = "shared memory of the persistent loop kernel";
                cudaGetLastError();
                return 0;
        }
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, LOOP_TB, smem) != cudaSuccess || per_sm < 1) {
                *why = "the persistent loop kernel does not fit an SM";
                cudaGetLastError();
                return 0;
        }
        return blk_sm_count();          // one block per SM: fewer arrivals per barrier than a fuller grid
}

int launch_loop_coop(const LoopArgs &args, int n, int np, int grid, cudaStream_t st, std::string *err)
{
        const void *f = loop_kernel_for(np, args.m.fold_every);
        if (!f) { *err = "no persistent loop kernel for this n"; return 1; }
        cudaError_t e = cudaMemsetAsync(args.bar, 0, 2 * sizeof(unsigned), st);
        if (e == cudaSuccess) {
                void *params[1] = {(void *)&args};
                e = cudaLaunchCooperativeKernel(f, dim3((unsigned)grid), dim3(LOOP_TB), params, loop_smem_bytes(n, np), st);
        }
        if (e != cudaSuccess) { *err = std::string("persistent loop kernel: ") + cudaGetErrorString(e); return 1; }
        return 0;
}
