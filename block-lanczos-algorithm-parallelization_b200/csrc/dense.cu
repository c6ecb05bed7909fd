// dense.cu -- the tall-skinny GF(p) phases of one block-Lanczos iteration:
//
//   k_dots    vtAv = v^T Av, vtAAv = Av^T Av                (block_dot_products, sequential/lanczos_modp.c:443-453,
//                                                           built there from matmul_CpAtB :305-315)
//   k_small   sum of partials, semi_inverse, c / vtAvd      (semi_inverse :342-438; the n x n part of orthogonalize :460-475)
//   k_ortho   next v and p, row by row                      (orthogonalize :478-491 built from matmul_CpAB :292-302,
//                                                           and the v <- tmp copy :655-656, here a no-op: v is updated in place)
//
// The reference does a 64-bit `%` per multiply-add; these kernels accumulate in u64 with the
// lazy fold of modp.cuh and reduce once per output.  Blocks are stored with leading dimension
// n_pad (power of two >= n, extra columns are zero and stay zero through every phase).
#include "blk_internal.cuh"
#include "small_body.cuh"
#include "dense_body.cuh"

namespace {

constexpr int DOTS_TB = 256;

// ------------------------------------------------------------------------------------------
// dots: a team of T = (NP/TI)^2 threads owns the NP x NP outputs (TI x TI register tile per
// thread, for both matrices); teams stride over the rows.  Block results are added to sums[2][NP*NP].
// ------------------------------------------------------------------------------------------
template <int NP, int FOLD>
__global__ void __launch_bounds__(DOTS_TB)
k_dots(int64_t rows, const u32 *__restrict__ v, const u32 *__restrict__ Av, unsigned long long *sums,
       ModP m, const DevSmall *state, SmallFuse fuse)
{
        pdl_prologue();
        if (state && state->halt) {
                // a halted iteration must not re-run orthogonalize (k_small would have cleared the flag)
                if (fuse.counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) fuse.state->do_ortho = 0;
                return;
        }

        dots_block<NP, FOLD, 0, DOTS_TB>(rows, v, Av, sums, m, blockIdx.x, gridDim.x);
        if (fuse.counter && last_block_done(fuse.counter, gridDim.x)) {
                extern __shared__ u32 sm_fused[];
                small_body(fuse.n, NP, sums, fuse.mats, fuse.state, 0, m, sm_fused);
        }
}

// ------------------------------------------------------------------------------------------
// small: one block (see small_body.cuh)
// ------------------------------------------------------------------------------------------
constexpr int SMALL_TB = 256;

__global__ void __launch_bounds__(SMALL_TB)
k_small(int n, int np, unsigned long long *__restrict__ sums, u32 *__restrict__ mats,
        DevSmall *__restrict__ state, int mode, ModP m)
{
        pdl_prologue();
        extern __shared__ u32 sm_small[];
        small_body(n, np, sums, mats, state, mode, m, sm_small);
}

// ------------------------------------------------------------------------------------------
// ortho: each row independently,
//   v'[r,:] = (d ? Av[r,:] : v[r,:]) + v[r,:] c + p[r,:] vtAvd
//   p'[r,:] = (d ? 0 : p[r,:])       + v[r,:] winv
// NP/JT threads per row, JT output columns each; the three n x n matrices sit in shared memory
// and are read with warp-broadcast LDS.128.
// ------------------------------------------------------------------------------------------
constexpr int ORTHO_TB = 128;

template <int NP, int JT, int FOLD>
__global__ void __launch_bounds__(ORTHO_TB)
k_ortho(int64_t rows, const u32 *v, const u32 *__restrict__ Av, const u32 *p, u32 *v_out, u32 *p_out,
        const u32 *__restrict__ mats, ModP m, const DevSmall *__restrict__ state, int force)
{
        pdl_prologue();
        extern __shared__ u32 sm[];
        u32 *C = sm, *D = C + NP * NP, *Wm = D + NP * NP, *dm = Wm + NP * NP;
        if (!force && !state->do_ortho) return;
        for (int e = threadIdx.x; e < NP * NP; e += ORTHO_TB) {
                C[e] = mats[MAT_C * NP * NP + e];
                D[e] = mats[MAT_VTAVD * NP * NP + e];
                Wm[e] = mats[MAT_WINV * NP * NP + e];
        }
        for (int e = threadIdx.x; e < NP; e += ORTHO_TB) dm[e] = mats[MAT_D * NP * NP + e];
        __syncthreads();

        ortho_slot<NP, JT, FOLD, 0>((int64_t)blockIdx.x * ORTHO_TB + threadIdx.x, rows, v, Av, p, v_out, p_out, C, D, Wm, dm, m);
}

// host-layout rows (n per row) -> device rows (np per row, zero padded).  Row r of src goes to device row
// map[r0 + r] when a relabelling is in force (scatter: src is read in order), else to row r of dst.
__global__ void k_pad_rows(const u32 *__restrict__ src, u32 *__restrict__ dst, int64_t rows, int n, int np,
                           const u32 *__restrict__ map, int64_t r0)
{
        int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= rows * np) return;
        int64_t r = e / np;
        int j = (int)(e - r * np);
        int64_t dr = map ? (int64_t)map[r0 + r] : r;
        dst[dr * np + j] = j < n ? src[r * n + j] : 0u;
}
// device rows -> host-layout rows: row r of dst comes from device row map[r0 + r] (gather), else row r of src
__global__ void k_unpad_rows(const u32 *__restrict__ src, u32 *__restrict__ dst, int64_t rows, int n, int np,
                             const u32 *__restrict__ map, int64_t r0)
{
        int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= rows * n) return;
        int64_t r = e / n;
        int j = (int)(e - r * n);
        int64_t sr = map ? (int64_t)map[r0 + r] : r;
        dst[e] = src[sr * np + j];
}

template <int NP>
int dots_fold(const ModP &m, int64_t rows, const u32 *v, const u32 *Av, u64 *sums, int nblocks,
              const DevSmall *state, const SmallFuse &fuse, cudaStream_t st)
{
        size_t smem = fuse.counter ? sizeof(u32) * small_smem_words(fuse.n) : 0;
        switch (m.fold_every) {
        case 0: launch_k(k_dots<NP, 0>, nblocks, DOTS_TB, smem, st, rows, v, Av, (unsigned long long *)sums, m, state, fuse); break;
        case 8: launch_k(k_dots<NP, 8>, nblocks, DOTS_TB, smem, st, rows, v, Av, (unsigned long long *)sums, m, state, fuse); break;
        default: launch_k(k_dots<NP, 2>, nblocks, DOTS_TB, smem, st, rows, v, Av, (unsigned long long *)sums, m, state, fuse); break;
        }
        return 1;
}

template <int NP, int JT, int FOLD>
void ortho_go(int64_t rows, u32 *v, const u32 *Av, u32 *p, u32 *v_out, u32 *p_out, const u32 *mats,
              const ModP &m, const DevSmall *state, int force, cudaStream_t st)
{
        size_t smem = sizeof(u32) * (3 * NP * NP + NP);
        if (rows < 0) {         // prepare only: opt in to > 48 KB of dynamic shared memory (per device)
                if (smem > 48 * 1024)
                        cudaFuncSetAttribute(k_ortho<NP, JT, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                return;
        }
        int64_t threads = rows * (NP / JT);
        unsigned blocks = (unsigned)((threads + ORTHO_TB - 1) / ORTHO_TB);
        launch_k(k_ortho<NP, JT, FOLD>, blocks, ORTHO_TB, smem, st, rows, v, Av, p, v_out, p_out, mats, m, state, force);
}

template <int NP, int JT>
int ortho_fold(int64_t rows, u32 *v, const u32 *Av, u32 *p, u32 *v_out, u32 *p_out, const u32 *mats,
               const ModP &m, const DevSmall *state, int force, cudaStream_t st)
{
        switch (m.fold_every) {
        case 0: ortho_go<NP, JT, 0>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st); break;
        case 8: ortho_go<NP, JT, 8>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st); break;
        default: ortho_go<NP, JT, 2>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st); break;
        }
        return 1;
}

}  // namespace

int dots_num_blocks(int64_t rows, int np)
{
        int ti = np < 4 ? np : 4;
        int team = (np / ti) * (np / ti);
        int teams = DOTS_TB / team;
        int64_t want = (rows + (int64_t)teams * 4 - 1) / ((int64_t)teams * 4);     // >= 4 rows per team
        if (want < 1) want = 1;
        if (want > blk_sm_count() * 4) want = blk_sm_count() * 4;
        return (int)want;
}

int launch_dots(const Geometry &geo, const ModP &m, int64_t rows, const u32 *v, const u32 *Av,
                u64 *sums, int nblocks, const DevSmall *state, const SmallFuse &fuse, cudaStream_t st)
{
        if (rows <= 0 && !fuse.counter) return 0;       // an empty shard adds nothing to the sums
        if (dense_umma_supported(geo.np, rows)) {
                int k = launch_dots_umma(geo.np, m, rows, v, Av, sums, state, fuse, st);
                if (k > 0) return k;          // (a tensor map that cannot be encoded falls through to mma.sync)
        }
        if (dense_mma_supported(geo.np)) return launch_dots_mma(geo.np, m, rows, v, Av, sums, state, fuse, st);
        switch (geo.np) {
        case 1: return dots_fold<1>(m, rows, v, Av, sums, nblocks, state, fuse, st);
        case 2: return dots_fold<2>(m, rows, v, Av, sums, nblocks, state, fuse, st);
        case 4: return dots_fold<4>(m, rows, v, Av, sums, nblocks, state, fuse, st);
        case 8: return dots_fold<8>(m, rows, v, Av, sums, nblocks, state, fuse, st);
        case 16: return dots_fold<16>(m, rows, v, Av, sums, nblocks, state, fuse, st);
        case 32: return dots_fold<32>(m, rows, v, Av, sums, nblocks, state, fuse, st);
        case 64: return dots_fold<64>(m, rows, v, Av, sums, nblocks, state, fuse, st);
        }
        return -1;
}

int launch_small(const Geometry &geo, const ModP &m, u64 *sums, u32 *mats, DevSmall *state, int mode,
                 cudaStream_t st)
{
        size_t smem = sizeof(u32) * small_smem_words(geo.n);
        launch_k(k_small, 1, SMALL_TB, smem, st, geo.n, geo.np, (unsigned long long *)sums, mats, state, mode, m);
        return 1;
}

int launch_ortho(const Geometry &geo, const ModP &m, int64_t rows, u32 *v, const u32 *Av, u32 *p,
                 u32 *v_out, u32 *p_out, const u32 *mats, const DevSmall *state, int force, cudaStream_t st)
{
        if (rows == 0) return 0;          // (rows < 0 is the prepare-only call of dense_prepare)
        if (rows > 0 && dense_umma_supported(geo.np, rows)) {
                int k = launch_ortho_umma(geo.np, m, rows, v, Av, p, v_out, p_out, mats, state, force, st);
                if (k > 0) return k;
        }
        if (rows >= 0 && dense_mma_supported(geo.np))
                return launch_ortho_mma(geo.np, m, rows, v, Av, p, v_out, p_out, mats, state, force, st);
        switch (geo.np) {
        case 1: return ortho_fold<1, 1>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 2: return ortho_fold<2, 2>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 4: return ortho_fold<4, 4>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 8: return ortho_fold<8, 8>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 16: return ortho_fold<16, 16>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 32: return ortho_fold<32, 16>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 64: return ortho_fold<64, 16>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        }
        return -1;
}

void dense_prepare(const Geometry &geo, const ModP &m)
{
        cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        launch_ortho(geo, m, -1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
        dense_mma_prepare(geo.np);
        dense_umma_prepare(geo.np);
}

int launch_pad_rows(const u32 *src, u32 *dst, int64_t rows, int n, int np, const u32 *map, int64_t r0, cudaStream_t st)
{
        int64_t tot = rows * np;
        if (tot == 0) return 0;
        k_pad_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, dst, rows, n, np, map, r0);
        return 1;
}
int launch_unpad_rows(const u32 *src, u32 *dst, int64_t rows, int n, int np, const u32 *map, int64_t r0, cudaStream_t st)
{
        int64_t tot = rows * n;
        if (tot == 0) return 0;
        k_unpad_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, dst, rows, n, np, map, r0);
        return 1;
}
