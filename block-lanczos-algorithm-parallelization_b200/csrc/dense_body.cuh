// dense_body.cuh -- device code of the CUDA-core dense phases (block_dot_products and the row loops of
// orthogonalize), shared by the stand-alone kernels (dense.cu: k_dots, k_ortho) and by the persistent loop
// kernel (loop_coop.cu).  COH = 1: blocks are re-read inside one launch after other thread blocks rewrote them,
// so they are loaded with ld.global.cg (L2) -- see spmv_body.cuh.
#pragma once
#include "blk_internal.cuh"

template <int V> struct VecD;
template <> struct VecD<1> { typedef unsigned int T; };
template <> struct VecD<2> { typedef uint2 T; };
template <> struct VecD<4> { typedef uint4 T; };

template <int V> __device__ __forceinline__ void ldv(u32 (&o)[V], const u32 *p)
{
        typename VecD<V>::T t = *reinterpret_cast<const typename VecD<V>::T *>(p);
        const u32 *w = reinterpret_cast<const u32 *>(&t);
#pragma unroll
        for (int k = 0; k < V; k++) o[k] = w[k];
}
template <int V> __device__ __forceinline__ void stv(u32 *p, const u32 (&o)[V])
{
        typename VecD<V>::T t;
        u32 *w = reinterpret_cast<u32 *>(&t);
#pragma unroll
        for (int k = 0; k < V; k++) w[k] = o[k];
        *reinterpret_cast<typename VecD<V>::T *>(p) = t;
}

template <int V, int COH> __device__ __forceinline__ void ldv_c(u32 (&o)[V], const u32 *p)
{
        if (COH) {
                typename VecD<V>::T t = __ldcg(reinterpret_cast<const typename VecD<V>::T *>(p));
                const u32 *w = reinterpret_cast<const u32 *>(&t);
#pragma unroll
                for (int k = 0; k < V; k++) o[k] = w[k];
        } else {
                ldv<V>(o, p);
        }
}

// dots: a team of T = (NP/TI)^2 threads owns the NP x NP outputs (TI x TI register tile per thread, for both
// matrices); teams stride over the rows (block `block` of `nblocks`, TB threads each).  The block's results are
// added to sums[2][NP*NP] (u64 integer atomics: order-free, hence deterministic).  Ends with the block's atomics issued.
template <int NP, int FOLD, int COH, int TB>
__device__ __forceinline__ void dots_block(const int64_t rows, const u32 *v, const u32 *Av, unsigned long long *sums, const ModP &m,
                                           const int64_t block, const int64_t nblocks)
{
        constexpr int TI = NP < 4 ? NP : 4;
        constexpr int PER = NP / TI;          // tiles per dimension
        constexpr int T = PER * PER;          // threads per team
        constexpr int TEAMS = TB / T;
        constexpr int FE = FOLD ? FOLD : 64;  // rows between folds
        const int tid = threadIdx.x;
        const int team = tid / T, tt = tid % T;
        const int i0 = (tt / PER) * TI, j0 = (tt % PER) * TI;
        u64 a1[TI][TI], a2[TI][TI];
#pragma unroll
        for (int a = 0; a < TI; a++)
#pragma unroll
                for (int b = 0; b < TI; b++) { a1[a][b] = 0; a2[a][b] = 0; }

        const int64_t stride = (int64_t)nblocks * TEAMS;
        int since = 0;
        for (int64_t r = (int64_t)block * TEAMS + team; r < rows; r += stride) {
                u32 vi[TI], ai[TI], aj[TI];
                ldv_c<TI, COH>(vi, v + r * NP + i0);
                ldv_c<TI, COH>(ai, Av + r * NP + i0);
                ldv_c<TI, COH>(aj, Av + r * NP + j0);
#pragma unroll
                for (int a = 0; a < TI; a++)
#pragma unroll
                        for (int b = 0; b < TI; b++) {
                                mp_mac(a1[a][b], vi[a], aj[b]);
                                mp_mac(a2[a][b], ai[a], aj[b]);
                        }
                if (++since == FE) {
                        since = 0;
#pragma unroll
                        for (int a = 0; a < TI; a++)
#pragma unroll
                                for (int b = 0; b < TI; b++) { mp_fold(a1[a][b], m); mp_fold(a2[a][b], m); }
                }
        }

        // block result -> global u64 sums (integer addition: order-free, hence deterministic).
        // Every addend is a canonical residue < 2^31, so 2^33 blocks could not overflow.
        // Teams inside a warp are combined with shuffles first (mod-p adds on u32).
        constexpr int TPW = T < 32 ? 32 / T : 1;       // teams per warp
        u32 r1[TI][TI], r2[TI][TI];
#pragma unroll
        for (int a = 0; a < TI; a++)
#pragma unroll
                for (int b = 0; b < TI; b++) {
                        u32 x = mp_reduce(a1[a][b], m), y = mp_reduce(a2[a][b], m);
#pragma unroll
                        for (int off = T; off < T * TPW; off <<= 1) {
                                x = mp_add(x, __shfl_xor_sync(0xffffffffu, x, off), m);
                                y = mp_add(y, __shfl_xor_sync(0xffffffffu, y, off), m);
                        }
                        r1[a][b] = x; r2[a][b] = y;
                }
        const bool writer = T >= 32 || (tid & 31) < T;  // one team per warp carries the warp's result
        if (TEAMS == 1) {
#pragma unroll
                for (int a = 0; a < TI; a++)
#pragma unroll
                        for (int b = 0; b < TI; b++) {
                                atomicAdd(&sums[(i0 + a) * NP + j0 + b], (unsigned long long)r1[a][b]);
                                atomicAdd(&sums[NP * NP + (i0 + a) * NP + j0 + b], (unsigned long long)r2[a][b]);
                        }
        } else {
                __shared__ unsigned long long acc[TEAMS == 1 ? 1 : 2 * NP * NP];
                for (int e = tid; e < 2 * NP * NP; e += TB) acc[e] = 0;
                __syncthreads();
                if (writer) {
#pragma unroll
                        for (int a = 0; a < TI; a++)
#pragma unroll
                                for (int b = 0; b < TI; b++) {
                                        atomicAdd(&acc[(i0 + a) * NP + j0 + b], (unsigned long long)r1[a][b]);
                                        atomicAdd(&acc[NP * NP + (i0 + a) * NP + j0 + b], (unsigned long long)r2[a][b]);
                                }
                }
                __syncthreads();
                for (int e = tid; e < 2 * NP * NP; e += TB)
                        atomicAdd(&sums[e], (unsigned long long)mp_reduce(acc[e], m));
        }
}

// ortho: one (row, JT-column slice) per call; gid = row * (NP/JT) + slice.  C, D, Wm (NP x NP each) and dm (NP) are
// the n x n operands in shared memory.  All lanes of a warp must call it together (in-place update: __syncwarp).
template <int NP, int JT, int FOLD, int COH>
__device__ __forceinline__ void ortho_slot(const int64_t gid, const int64_t rows, const u32 *v, const u32 *Av, const u32 *p,
                                           u32 *v_out, u32 *p_out, const u32 *C, const u32 *D, const u32 *Wm, const u32 *dm,
                                           const ModP &m)
{
        constexpr int TPR = NP / JT;
        constexpr int KV = NP < 4 ? NP : 4;
        constexpr int JV = JT < 4 ? JT : 4;
        constexpr int FV = FOLD ? FOLD / 2 : 32;     // k-steps between folds of accV (2 products per k)
        constexpr int FP = FOLD ? FOLD : 64;         // ... of accP (1 product per k)
        const int64_t r = gid / TPR;
        const int j0 = (int)(gid % TPR) * JT;
        const bool active = r < rows;
        u64 accV[JT], accP[JT];
#pragma unroll
        for (int j = 0; j < JT; j++) { accV[j] = 0; accP[j] = 0; }
        u32 nv[JT], npw[JT];
        if (active) {
                const u32 *vr = v + r * NP, *pr = p + r * NP;
#pragma unroll
                for (int k0 = 0; k0 < NP; k0 += KV) {
                        u32 vk[KV], pk[KV];
                        ldv_c<KV, COH>(vk, vr + k0);
                        ldv_c<KV, COH>(pk, pr + k0);
#pragma unroll
                        for (int kk = 0; kk < KV; kk++) {
                                const int k = k0 + kk;
#pragma unroll
                                for (int jj = 0; jj < JT; jj += JV) {
                                        u32 cc[JV], dd[JV], ww[JV];
                                        ldv<JV>(cc, C + k * NP + j0 + jj);
                                        ldv<JV>(dd, D + k * NP + j0 + jj);
                                        ldv<JV>(ww, Wm + k * NP + j0 + jj);
#pragma unroll
                                        for (int j = 0; j < JV; j++) {
                                                mp_mac(accV[jj + j], vk[kk], cc[j]);
                                                mp_mac(accV[jj + j], pk[kk], dd[j]);
                                                mp_mac(accP[jj + j], vk[kk], ww[j]);
                                        }
                                }
                                if ((k + 1) % FV == 0) {
#pragma unroll
                                        for (int j = 0; j < JT; j++) mp_fold(accV[j], m);
                                }
                                if ((k + 1) % FP == 0) {
#pragma unroll
                                        for (int j = 0; j < JT; j++) mp_fold(accP[j], m);
                                }
                        }
                }
#pragma unroll
                for (int jj = 0; jj < JT; jj += JV) {
                        u32 av[JV], vb[JV], pb[JV];
                        ldv_c<JV, COH>(av, Av + r * NP + j0 + jj);
                        ldv_c<JV, COH>(vb, vr + j0 + jj);
                        ldv_c<JV, COH>(pb, pr + j0 + jj);
#pragma unroll
                        for (int j = 0; j < JV; j++) {
                                bool dj = dm[j0 + jj + j] != 0;
                                nv[jj + j] = mp_add(mp_reduce(accV[jj + j], m), dj ? av[j] : vb[j], m);
                                npw[jj + j] = mp_add(mp_reduce(accP[jj + j], m), dj ? 0u : pb[j], m);
                        }
                }
        }
        if (TPR > 1) __syncwarp();      // in place: every lane of the row has read v, p before anyone writes
        if (active) {
#pragma unroll
                for (int jj = 0; jj < JT; jj += JV) {
                        u32 a[JV], b[JV];
#pragma unroll
                        for (int j = 0; j < JV; j++) { a[j] = nv[jj + j]; b[j] = npw[jj + j]; }
                        stv<JV>(v_out + r * NP + j0 + jj, a);
                        stv<JV>(p_out + r * NP + j0 + jj, b);
                }
        }
}
