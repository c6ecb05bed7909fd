// dense_mma.cu -- the tall-skinny GF(p) phases on the int8 tensor-core path (n_pad in {8,16,32}).
//
// ncu on the CUDA-core kernels of dense.cu (config 4, n = 16) shows both phases bound by the
// integer pipe: 11.2 ms each against 1.0 ms (dots) and 2.5 ms (orthogonalize) of HBM time.
// BASELINE.json's north_star allows tensor cores exactly in that case.  A 31-bit residue is four
// byte limbs, x = sum_a x_a 2^(8a), so a GF(p) product of matrices becomes 8-bit integer GEMMs
// with s32 accumulation plus a recombination mod p:
//
//   orthogonalize (sequential/lanczos_modp.c:478-491):  out[r,:] = base[r,:] + v[r,:] C  (+ p[r,:] D)
//       A operand = the raw bytes of row r of v (K = 4 n_pad: byte kappa = 4k + a is limb a of
//       v[r,k]) -- no repacking at all; B operand = limb b of (2^(8a) C[k,j] mod p), prepared once
//       per iteration by k_small in fragment order (`bfrag`); out = sum_b 2^(8b) S_b mod p.
//   block_dot_products (:443-453):  C[i,j] = sum_r X[r,i] Y[r,j]
//       K runs over rows, so both operands are limb planes transposed on the fly (4x4 byte
//       transposes with PRMT); the four warps of a block take one limb b of Y each.
//
// mma.sync.m16n8k32.u8.u8.s32 (SASS IMMA.16832.U8.U8) sustains 573 T int8 MAC/s on B200
// (tools/imma_bench.cu), so the MMA rate is not the bound here; the operand preparation on the
// SM's instruction stream is (3x / 2x the HBM time on config 4).  n_pad = 16 therefore runs on
// the TMA-fed tcgen05 kernels of dense_umma.cu by default; these kernels serve n_pad = 8 and 32
// and BLK_DENSE=mma.
#include <cstdlib>
#include "blk_internal.cuh"
#include "small_body.cuh"

namespace {

__device__ __forceinline__ void imma(int (&d)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1)
{
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int W> __device__ __forceinline__ void load_words(u32 (&o)[W], const u32 *p, bool ok)
{
        if (!ok) {
#pragma unroll
                for (int k = 0; k < W; k++) o[k] = 0;
                return;
        }
        if (W == 1) {
                o[0] = *p;
        } else if (W == 2) {
                uint2 t = *reinterpret_cast<const uint2 *>(p);
                o[0] = t.x; o[1 % W] = t.y;
        } else {
#pragma unroll
                for (int k = 0; k < W; k += 4) {
                        uint4 t = *reinterpret_cast<const uint4 *>(p + k);
                        o[k] = t.x; o[(k + 1) % W] = t.y; o[(k + 2) % W] = t.z; o[(k + 3) % W] = t.w;
                }
        }
}

// ------------------------------------------------------------------------------------------
// orthogonalize.  One warp per 16 rows; thread (g = lane>>2, tig = lane&3) holds words
// [tig*NP/4, (tig+1)*NP/4) of rows r0+g and r0+g+8 of v and p: these ARE its A fragments
// (k-slot (s, half h, tig, byte) <-> word tig*NP/4 + 2s + h; bfrag uses the same map).
// ------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(128)
k_ortho_mma(int64_t rows, const u32 *v, const u32 *__restrict__ Av, const u32 *p, u32 *v_out, u32 *p_out,
            const u32 *__restrict__ mats, ModP m, const DevSmall *__restrict__ state, int force)
{
        pdl_prologue();
        constexpr int T = NP / 8, S = NP / 8, WPT = NP / 4;
        constexpr int FRAG = 12 * T * S * 64;
        extern __shared__ u32 sm[];
        u32 *dsm = sm + FRAG;
        if (!force && !state->do_ortho) return;
        const u32 *bfrag = mats + MAT_COUNT * NP * NP;
        for (int e = threadIdx.x; e < FRAG / 4; e += 128)
                reinterpret_cast<uint4 *>(sm)[e] = reinterpret_cast<const uint4 *>(bfrag)[e];
        for (int e = threadIdx.x; e < NP; e += 128) dsm[e] = mats[MAT_D * NP * NP + e];
        __syncthreads();

        const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
        const int64_t warps = (int64_t)gridDim.x * 4;
        const int64_t ntile = (rows + 15) / 16;
        const uint2 *fr = reinterpret_cast<const uint2 *>(sm);
        for (int64_t tile = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5); tile < ntile; tile += warps) {
                const int64_t rA = tile * 16 + g, rB = rA + 8;
                const bool okA = rA < rows, okB = rB < rows;
                u32 av[2][WPT], ap[2][WPT];
                load_words<WPT>(av[0], v + rA * NP + tig * WPT, okA);
                load_words<WPT>(av[1], v + rB * NP + tig * WPT, okB);
                load_words<WPT>(ap[0], p + rA * NP + tig * WPT, okA);
                load_words<WPT>(ap[1], p + rB * NP + tig * WPT, okB);
                u32 resV[T][4], resP[T][4];
#pragma unroll
                for (int t = 0; t < T; t++) {
                        u64 oV[4] = {0, 0, 0, 0}, oP[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                                int cV[4] = {0, 0, 0, 0}, cP[4] = {0, 0, 0, 0};
#pragma unroll
                                for (int s = 0; s < S; s++) {
                                        const uint2 bc = fr[(((0 * 4 + b) * T + t) * S + s) * 32 + lane];
                                        const uint2 bd = fr[(((1 * 4 + b) * T + t) * S + s) * 32 + lane];
                                        const uint2 bw = fr[(((2 * 4 + b) * T + t) * S + s) * 32 + lane];
                                        imma(cV, av[0][2 * s], av[1][2 * s], av[0][2 * s + 1], av[1][2 * s + 1], bc.x, bc.y);
                                        imma(cV, ap[0][2 * s], ap[1][2 * s], ap[0][2 * s + 1], ap[1][2 * s + 1], bd.x, bd.y);
                                        imma(cP, av[0][2 * s], av[1][2 * s], av[0][2 * s + 1], av[1][2 * s + 1], bw.x, bw.y);
                                }
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                        oV[i] += (u64)(u32)cV[i] << (8 * b);
                                        oP[i] += (u64)(u32)cP[i] << (8 * b);
                                }
                        }
                        // this thread's outputs: rows rA (i = 0,1) and rB (i = 2,3), columns c0, c0+1
                        const int c0 = 8 * t + 2 * tig;
                        const bool d0 = dsm[c0] != 0, d1 = dsm[c0 + 1] != 0;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                                const int64_t r = h ? rB : rA;
                                u32 ba[2] = {0, 0}, bv[2] = {0, 0}, bp[2] = {0, 0};
                                if (h ? okB : okA) {
                                        load_words<2>(ba, Av + r * NP + c0, true);
                                        load_words<2>(bv, v + r * NP + c0, true);
                                        load_words<2>(bp, p + r * NP + c0, true);
                                }
                                resV[t][2 * h] = mp_add(mp_reduce(oV[2 * h], m), d0 ? ba[0] : bv[0], m);
                                resV[t][2 * h + 1] = mp_add(mp_reduce(oV[2 * h + 1], m), d1 ? ba[1] : bv[1], m);
                                resP[t][2 * h] = mp_add(mp_reduce(oP[2 * h], m), d0 ? 0u : bp[0], m);
                                resP[t][2 * h + 1] = mp_add(mp_reduce(oP[2 * h + 1], m), d1 ? 0u : bp[1], m);
                        }
                }
                __syncwarp();          // in place: every lane holds its inputs in registers before anyone writes
#pragma unroll
                for (int t = 0; t < T; t++) {
                        const int c0 = 8 * t + 2 * tig;
                        if (okA) {
                                *reinterpret_cast<uint2 *>(v_out + rA * NP + c0) = make_uint2(resV[t][0], resV[t][1]);
                                *reinterpret_cast<uint2 *>(p_out + rA * NP + c0) = make_uint2(resP[t][0], resP[t][1]);
                        }
                        if (okB) {
                                *reinterpret_cast<uint2 *>(v_out + rB * NP + c0) = make_uint2(resV[t][2], resV[t][3]);
                                *reinterpret_cast<uint2 *>(p_out + rB * NP + c0) = make_uint2(resP[t][2], resP[t][3]);
                        }
                }
        }
}

// ------------------------------------------------------------------------------------------
// dots.  A block = 4 warps working on the same 32-row steps; warp w owns limb b = w of the
// right operand.  CB = min(NP,16) columns per block (NP = 32: gridDim.y = 4 column-block pairs).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void limb_planes(u32 (&L)[4], u32 x0, u32 x1, u32 x2, u32 x3)
{
        // L[a] = {byte a of x0, byte a of x1, byte a of x2, byte a of x3}
        u32 t0 = __byte_perm(x0, x1, 0x5140), t1 = __byte_perm(x2, x3, 0x5140);
        u32 t2 = __byte_perm(x0, x1, 0x7362), t3 = __byte_perm(x2, x3, 0x7362);
        L[0] = __byte_perm(t0, t1, 0x5410); L[1] = __byte_perm(t0, t1, 0x7632);
        L[2] = __byte_perm(t2, t3, 0x5410); L[3] = __byte_perm(t2, t3, 0x7632);
}

// Limb planes of columns col, col+1 (.. col+NC-1) over the 8 rows this thread covers in a 32-row
// step.  The k-slot -> row map is free as long as both operands use it: slot (half h, tig, beta)
// is row r0 + 16h + 4 beta + tig, so one load instruction (fixed h, beta) touches 4 consecutive
// rows = 2 full 128-byte lines at n_pad = 16 instead of 4 half-used ones.
template <int NC, bool FULL>
__device__ __forceinline__ void column_planes(u32 (&lo)[NC][4], u32 (&hi)[NC][4], const u32 *__restrict__ X, int NP,
                                              int64_t r0, int col, int tig, int64_t rows)
{
        u32 x[8][NC];
        // all 8 rows of this thread sit at compile-time offsets from one base pointer; the row-bound
        // tests exist only in the last (partial) step of the block
        const u32 *base = X + (r0 + tig) * NP + col;
#pragma unroll
        for (int q = 0; q < 8; q++) {
                const int dr = (q < 4 ? 0 : 16) + 4 * (q & 3);
                if (FULL || r0 + tig + dr < rows) load_words<NC>(x[q], base + (int64_t)dr * NP, true);
                else {
#pragma unroll
                        for (int c = 0; c < NC; c++) x[q][c] = 0;
                }
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
                limb_planes(lo[c], x[0][c], x[1][c], x[2][c], x[3][c]);
                limb_planes(hi[c], x[4][c], x[5][c], x[6][c], x[7][c]);
        }
}

template <int NP>
__global__ void __launch_bounds__(128)
k_dots_mma(int64_t rows, const u32 *__restrict__ v, const u32 *__restrict__ Av,
           unsigned long long *sums, ModP m, const DevSmall *state, SmallFuse fuse)
{
        pdl_prologue();
        constexpr int CB = NP < 16 ? NP : 16;          // columns per block (8 or 16)
        constexpr int NB = NP / CB;                    // column blocks per dimension
        constexpr int MT = CB == 16 ? 4 : 2;           // m-tiles: CB=16 one per limb a; CB=8 two limbs per tile
        constexpr int NT = CB == 16 ? 2 : 1;           // n-tiles of this warp's limb
        constexpr int FLUSH = 1000;                    // 32 000 rows * 255^2 < 2^31
        if (state && state->halt) {
                // a halted iteration must not re-run orthogonalize (k_small would have cleared the flag)
                if (fuse.counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) fuse.state->do_ortho = 0;
                return;
        }
        const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3, b = threadIdx.x >> 5;
        const int ib = blockIdx.y / NB, jb = blockIdx.y % NB;
        int accVA[MT][NT][4], accAA[MT][NT][4];
#pragma unroll
        for (int a = 0; a < MT; a++)
#pragma unroll
                for (int t = 0; t < NT; t++)
#pragma unroll
                        for (int i = 0; i < 4; i++) { accVA[a][t][i] = 0; accAA[a][t][i] = 0; }

        // block-level accumulators: the four warps (limbs b) add here first, so that each block
        // issues one global atomic per output instead of one per warp and flush
        __shared__ unsigned long long acc[2 * CB * CB];
        for (int e = threadIdx.x; e < 2 * CB * CB; e += 128) acc[e] = 0;
        __syncthreads();
        const u32 pw = mp_reduce(1ull << (8 * b), m);              // 2^(8b) mod p
        auto flush = [&]() {
#pragma unroll
                for (int t = 0; t < NT; t++)
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                                // C fragment element i: row = g + 8*(i>>1), col = 2*tig + (i&1)
                                u64 sVA = 0, sAA = 0;
                                // CB = 16: tile rows g / g+8 hold left columns 2g / 2g+1, n-tile t holds
                                // right columns 2n + t (n = index inside the tile)
                                int oi, oj = CB == 16 ? 2 * (2 * tig + (i & 1)) + t : 2 * tig + (i & 1);
                                if constexpr (CB == 16) {
                                        oi = 2 * g + (i >> 1);
#pragma unroll
                                        for (int a = 0; a < 4; a++) {
                                                sVA += (u64)(u32)accVA[a][t][i] << (8 * a);
                                                sAA += (u64)(u32)accAA[a][t][i] << (8 * a);
                                        }
                                } else {
                                        // rows g: limb 2q, rows g+8: limb 2q+1, both for column g: elements i and i^2
                                        // belong to the same output; let the i < 2 element collect both
                                        oi = g;
                                        if (i >= 2) continue;
#pragma unroll
                                        for (int q = 0; q < 2; q++) {
                                                sVA += ((u64)(u32)accVA[q][t][i] << (16 * q)) + ((u64)(u32)accVA[q][t][i + 2] << (16 * q + 8));
                                                sAA += ((u64)(u32)accAA[q][t][i] << (16 * q)) + ((u64)(u32)accAA[q][t][i + 2] << (16 * q + 8));
                                        }
                                }
                                atomicAdd(&acc[oi * CB + oj], (unsigned long long)mp_mul(mp_reduce(sVA, m), pw, m));
                                atomicAdd(&acc[CB * CB + oi * CB + oj], (unsigned long long)mp_mul(mp_reduce(sAA, m), pw, m));
                        }
#pragma unroll
                for (int a = 0; a < MT; a++)
#pragma unroll
                        for (int t = 0; t < NT; t++)
#pragma unroll
                                for (int i = 0; i < 4; i++) { accVA[a][t][i] = 0; accAA[a][t][i] = 0; }
        };

        const int64_t nstep = (rows + 31) / 32;
        int since = 0;
        for (int64_t step = blockIdx.x; step < nstep; step += gridDim.x) {
                const int64_t r0 = step * 32;
                // left operands: columns ib*CB + 2g, 2g+1 (CB = 16) or ib*CB + g (CB = 8) of v and Av
                u32 vlo[NT][4], vhi[NT][4], alo[NT][4], ahi[NT][4], blo[NT][4], bhi[NT][4];
                if (r0 + 32 <= rows) {
                        column_planes<NT, true>(vlo, vhi, v, NP, r0, ib * CB + NT * g, tig, rows);
                        column_planes<NT, true>(alo, ahi, Av, NP, r0, ib * CB + NT * g, tig, rows);
                        if (NB > 1) column_planes<NT, true>(blo, bhi, Av, NP, r0, jb * CB + NT * g, tig, rows);
                } else {
                        column_planes<NT, false>(vlo, vhi, v, NP, r0, ib * CB + NT * g, tig, rows);
                        column_planes<NT, false>(alo, ahi, Av, NP, r0, ib * CB + NT * g, tig, rows);
                        if (NB > 1) column_planes<NT, false>(blo, bhi, Av, NP, r0, jb * CB + NT * g, tig, rows);
                }
                if (NB == 1) {
#pragma unroll
                        for (int c = 0; c < NT; c++)
#pragma unroll
                                for (int a = 0; a < 4; a++) { blo[c][a] = alo[c][a]; bhi[c][a] = ahi[c][a]; }
                }
#pragma unroll
                for (int a = 0; a < MT; a++) {
                        u32 v0, v1, v2, v3, a0, a1, a2, a3;
                        if constexpr (CB == 16) {         // rows g -> column 2g, rows g+8 -> column 2g+1, limb a
                                v0 = vlo[0][a]; v1 = vlo[NT - 1][a]; v2 = vhi[0][a]; v3 = vhi[NT - 1][a];
                                a0 = alo[0][a]; a1 = alo[NT - 1][a]; a2 = ahi[0][a]; a3 = ahi[NT - 1][a];
                        } else {                // rows g -> limb 2a, rows g+8 -> limb 2a+1, column g
                                v0 = vlo[0][2 * a]; v1 = vlo[0][2 * a + 1]; v2 = vhi[0][2 * a]; v3 = vhi[0][2 * a + 1];
                                a0 = alo[0][2 * a]; a1 = alo[0][2 * a + 1]; a2 = ahi[0][2 * a]; a3 = ahi[0][2 * a + 1];
                        }
#pragma unroll
                        for (int t = 0; t < NT; t++) {
                                imma(accVA[a][t], v0, v1, v2, v3, blo[t][b], bhi[t][b]);
                                imma(accAA[a][t], a0, a1, a2, a3, blo[t][b], bhi[t][b]);
                        }
                }
                if (++since == FLUSH) { flush(); since = 0; }
        }
        flush();
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * CB * CB; e += 128) {
                int which = e / (CB * CB), r = e % (CB * CB);
                int oi = ib * CB + r / CB, oj = jb * CB + r % CB;
                atomicAdd(&sums[which * NP * NP + oi * NP + oj], (unsigned long long)mp_reduce(acc[e], m));
        }
        if (fuse.counter && last_block_done(fuse.counter, gridDim.x * gridDim.y)) {
                extern __shared__ u32 sm_fused[];
                small_body(fuse.n, NP, sums, fuse.mats, fuse.state, 0, m, sm_fused);
        }
}

bool use_mma()
{
        static int on = -1;
        if (on < 0) {
                const char *e = getenv("BLK_DENSE");
                on = !(e && e[0] == 'c');          // BLK_DENSE=cuda selects the CUDA-core kernels
        }
        return on != 0;
}

template <int NP> size_t ortho_smem() { return sizeof(u32) * (12 * (NP / 8) * (NP / 8) * 64 + NP); }

template <int NP>
int ortho_go(int64_t rows, u32 *v, const u32 *Av, u32 *p, u32 *v_out, u32 *p_out, const u32 *mats, const ModP &m,
             const DevSmall *state, int force, cudaStream_t st)
{
        if (rows < 0) {
                if (ortho_smem<NP>() > 48 * 1024)
                        cudaFuncSetAttribute(k_ortho_mma<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem<NP>());
                return 0;
        }
        int64_t tiles = (rows + 15) / 16;
        int64_t blocks = (tiles + 3) / 4;
        int64_t cap = (int64_t)blk_sm_count() * (NP == 32 ? 4 : 8);
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        launch_k(k_ortho_mma<NP>, (unsigned)blocks, 128, ortho_smem<NP>(), st, rows, v, Av, p, v_out, p_out, mats, m, state, force);
        return 1;
}

template <int NP>
int dots_go(int64_t rows, const u32 *v, const u32 *Av, u64 *sums, const ModP &m, const DevSmall *state,
            const SmallFuse &fuse, cudaStream_t st)
{
        constexpr int NB = NP / (NP < 16 ? NP : 16);
        int64_t steps = (rows + 31) / 32;
        int64_t bx = (steps + 7) / 8;                    // >= 8 steps per block
        int64_t cap = (int64_t)blk_sm_count() * 8 / (NB * NB);
        if (bx > cap) bx = cap;
        if (bx < 1) bx = 1;
        dim3 grid((unsigned)bx, NB * NB);
        size_t smem = fuse.counter ? sizeof(u32) * small_smem_words(fuse.n) : 0;
        launch_k(k_dots_mma<NP>, grid, 128, smem, st, rows, v, Av, (unsigned long long *)sums, m, state, fuse);
        return 1;
}

}  // namespace

bool dense_mma_supported(int np) { return use_mma() && (np == 8 || np == 16 || np == 32); }

void dense_mma_prepare(int np)
{
        switch (np) {
        case 8: ortho_go<8>(-1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ModP(), nullptr, 0, nullptr); break;
        case 16: ortho_go<16>(-1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ModP(), nullptr, 0, nullptr); break;
        case 32: ortho_go<32>(-1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ModP(), nullptr, 0, nullptr); break;
        }
}

int launch_ortho_mma(int np, const ModP &m, int64_t rows, u32 *v, const u32 *Av, u32 *p, u32 *v_out, u32 *p_out,
                     const u32 *mats, const DevSmall *state, int force, cudaStream_t st)
{
        switch (np) {
        case 8: return ortho_go<8>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 16: return ortho_go<16>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        case 32: return ortho_go<32>(rows, v, Av, p, v_out, p_out, mats, m, state, force, st);
        }
        return -1;
}

int launch_dots_mma(int np, const ModP &m, int64_t rows, const u32 *v, const u32 *Av, u64 *sums,
                    const DevSmall *state, const SmallFuse &fuse, cudaStream_t st)
{
        switch (np) {
        case 8: return dots_go<8>(rows, v, Av, sums, m, state, fuse, st);
        case 16: return dots_go<16>(rows, v, Av, sums, m, state, fuse, st);
        case 32: return dots_go<32>(rows, v, Av, sums, m, state, fuse, st);
        }
        return -1;
}
