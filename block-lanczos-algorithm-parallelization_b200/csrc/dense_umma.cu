// dense_umma.cu -- block_dot_products and orthogonalize on tcgen05 (5th-generation tensor cores),
// n_pad = 16.
//
// The IMMA kernel of dense_mma.cu spends its time preparing operands (PRMT byte transposes, 8 row
// loads per thread and step): 3.1 ms on config 4 against 1.0 ms of HBM time.  Here nothing of
// that is left on the SM's instruction stream:
//
//   * a row of a block is 64 contiguous bytes (column i, limb a) -> byte 4i + a.  A TMA box of
//     {64 bytes, 256 rows} with SWIZZLE_64B lands in shared memory as the Major-MN canonical
//     layout ((16,4),(8,k)) : ((1,16),(64,512)), which tcgen05.mma kind::i8 takes for A and for B
//     (tools/umma_i8_probe.cu is the known-answer test of exactly this),
//   * so   D[(i,a),(j,b)] = sum_r X_bytes[r][4i+a] * Av_bytes[r][4j+b]   (u8 x u8 -> s32, K = rows)
//     is 32 rows per instruction with no data movement by threads at all.  X = v gives vtAv,
//     X = Av gives vtAAv (sequential/lanczos_modp.c:443-453); in WIDE mode the two are one
//     M = 128 instruction whose A operand spans the v tile and the Av tile (two 64-byte MN atoms,
//     LBO = distance of the tiles),
//   * accumulators live in TMEM, double buffered; every 32 000 rows (255^2 * 32 000 < 2^31) four
//     epilogue warps read them back (tcgen05.ld) and recombine  sum_ab 2^(8(a+b)) D mod p  while
//     the tensor core already works on the next epoch.
//
// Dots: one persistent CTA per SM: warp 4 lane 0 = TMA producer, warp 5 lane 0 = MMA issuer,
// warps 0-3 = epilogue (TMEM lanes 32w..32w+31).
//
// orthogonalize (sequential/lanczos_modp.c:456-491) is the transposed shape: M = rows, and a
// 128-row TMA tile of v (or p) is the K-major A operand as it stands (K = the 64 bytes of a row);
// the B operands, limb b of (2^(8a) X[i][j] mod p) at [K = 4i + a][N = 4j + b] for X in
// {c, vtAvd, winv}, are three 4 KB tiles every CTA derives from `mats` when it starts.  Six
// instructions per 128 rows leave D_v = v c + p vtAvd and D_p = v winv in TMEM; sixteen epilogue
// warps recombine the limbs, add the base term, overwrite the v and p tiles in shared memory and
// a store warp sends them back with TMA.
//
// Every mbarrier wait is bounded and traps, so a broken pipeline is a loud launch failure, not a
// hang.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "blk_internal.cuh"
#include "small_body.cuh"

namespace {

constexpr int ROW_BYTES = 64;                         // n_pad = 16
constexpr int TILE_ROWS = 256;                        // rows per TMA box (box dimension limit)
constexpr int TILE_BYTES = TILE_ROWS * ROW_BYTES;     // 16 KB per operand
constexpr int STAGES = 4;                             // 4 x 32 KB in flight per SM
constexpr int EPOCH_TILES = 125;                      // 32 000 rows between accumulator flushes
constexpr int KSTEP_ROWS = 32;                        // K of one kind::i8 instruction
constexpr int THREADS = 192;
#ifndef BLK_UMMA_TIMEOUT_CYCLES
#define BLK_UMMA_TIMEOUT_CYCLES 20000000000ll          // ~10 s of SM clock
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}

// HINT: pass a suspend-time hint (ns) so that a waiting warp is parked by the hardware instead of
// re-issuing try_wait from its instruction stream (experiment: variants 4-7 of the orthogonalize kernel)
template <bool HINT = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
        const long long t0 = clock64();
        do {
                uint32_t ok;
                if (HINT)
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
                else
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                if (ok) return;
        } while (clock64() - t0 < BLK_UMMA_TIMEOUT_CYCLES);
        __trap();          // no progress for seconds: fail the launch instead of hanging the GPU
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_rows(uint32_t dst, const CUtensorMap *map, int row0, uint32_t bar)
{
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(dst), "l"(map), "r"(0), "r"(row0), "r"(bar) : "memory");
}

// shared-memory matrix descriptor, Major-MN, SWIZZLE_64B: 8 rows (K) per 512-byte group (SBO);
// LBO = byte distance to the next 64-byte MN atom (only used when the operand is 128 wide)
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo)
{
        uint64_t d = 0;
        d |= (uint64_t)((addr & 0x3FFFF) >> 4);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)(512 >> 4) << 32;
        d |= 1ull << 46;
        d |= 4ull << 61;
        return d;
}

// D = s32, A = B = u8, both Major-MN
__host__ __device__ constexpr uint32_t i8_idesc(int M, int N)
{
        return (2u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
                     :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// this thread's TMEM lane, 64 consecutive columns
__device__ __forceinline__ void tmem_ld64(uint32_t (&r)[64], uint32_t taddr)
{
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                     "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
                       "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                       "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
                       "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
                       "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One accumulator row m = 4i + a (this thread's TMEM lane), columns 4j + b in r[]:
// acc[i][j] += 2^(8a) * sum_b 2^(8b) r[4j+b]  (mod p), the four limbs a sit in four adjacent lanes.
__device__ __forceinline__ void fold_row(const uint32_t (&r)[64], int m_row, bool valid, unsigned long long *acc, const ModP &m)
{
        const int i = m_row >> 2, a = m_row & 3;
        const u32 pw = mp_reduce(1ull << (8 * a), m);
#pragma unroll
        for (int j = 0; j < 16; j++) {
                u64 t = (u64)r[4 * j] + ((u64)r[4 * j + 1] << 8) + ((u64)r[4 * j + 2] << 16) + ((u64)r[4 * j + 3] << 24);
                u32 x = mp_mul(mp_reduce(t, m), pw, m);
                x = mp_add(x, __shfl_xor_sync(0xffffffffu, x, 1), m);
                x = mp_add(x, __shfl_xor_sync(0xffffffffu, x, 2), m);
                if (valid && a == 0) atomicAdd(&acc[i * 16 + j], (unsigned long long)x);
        }
}

template <bool WIDE>
__global__ void __launch_bounds__(THREADS, 1)
k_dots_umma(const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_av, int64_t ntile,
            unsigned long long *sums, ModP m, const DevSmall *state, SmallFuse fuse)
{
        pdl_prologue();
        constexpr int NP = 16;
        constexpr uint32_t BUF_COLS = WIDE ? 64 : 128;         // TMEM columns of one accumulator set
        extern __shared__ uint8_t dyn_raw[];
        __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
        __shared__ uint32_t tmem_slot;
        __shared__ unsigned long long acc[2 * NP * NP];

        if (state && state->halt) {
                // a halted iteration must not re-run orthogonalize (k_small would have cleared the flag)
                if (fuse.counter && blockIdx.x == 0 && threadIdx.x == 0) fuse.state->do_ortho = 0;
                return;
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const uint32_t stage0 = (smem_u32(dyn_raw) + 1023u) & ~1023u;
        const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[STAGES]);
        const uint32_t bar_acc_full = smem_u32(&bars[2 * STAGES]), bar_acc_empty = smem_u32(&bars[2 * STAGES + 2]);

        const int64_t t_lo = ntile * blockIdx.x / gridDim.x, t_hi = ntile * (blockIdx.x + 1) / gridDim.x;
        const int my = (int)(t_hi - t_lo);
        const int nepoch = (my + EPOCH_TILES - 1) / EPOCH_TILES;

        for (int e = threadIdx.x; e < 2 * NP * NP; e += THREADS) acc[e] = 0;
        if (threadIdx.x == 0) {
                for (int s = 0; s < STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
                for (int b = 0; b < 2; b++) { mbar_init(bar_acc_full + 8 * b, 1); mbar_init(bar_acc_empty + 8 * b, 128); }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 4) {
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(2 * BUF_COLS));
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem = tmem_slot;

        if (warp == 4) {
                if (lane == 0) {                                   // ---- TMA producer
                        for (int t = 0; t < my; t++) {
                                const int s = t % STAGES;
                                const uint32_t ph = (uint32_t)(t / STAGES) & 1u;
                                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                             :: "r"(bar_full + 8 * s), "r"(2 * TILE_BYTES) : "memory");
                                const int row0 = (int)((t_lo + t) * TILE_ROWS);
                                tma_rows(stage0 + s * 2 * TILE_BYTES, &map_v, row0, bar_full + 8 * s);
                                tma_rows(stage0 + s * 2 * TILE_BYTES + TILE_BYTES, &map_av, row0, bar_full + 8 * s);
                        }
                }
                __syncwarp();
        } else if (warp == 5) {
                if (lane == 0) {                                   // ---- MMA issuer
                        int t = 0;
                        for (int e = 0; e < nepoch; e++) {
                                const int b = e & 1;
                                mbar_wait(bar_acc_empty + 8 * b, (((uint32_t)e >> 1) & 1u) ^ 1u);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                const uint32_t d = tmem + b * BUF_COLS;
                                uint32_t accumulate = 0;
                                const int t_end = min(my, (e + 1) * EPOCH_TILES);
                                for (; t < t_end; t++) {
                                        const int s = t % STAGES;
                                        mbar_wait(bar_full + 8 * s, (uint32_t)(t / STAGES) & 1u);
                                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                        const uint32_t sv = stage0 + s * 2 * TILE_BYTES, sa = sv + TILE_BYTES;
#pragma unroll
                                        for (int k = 0; k < TILE_ROWS / KSTEP_ROWS; k++) {
                                                const uint32_t off = k * KSTEP_ROWS * ROW_BYTES;
                                                const uint64_t db = mn_desc(sa + off, ROW_BYTES);
                                                if (WIDE) {
                                                        // rows 0-63 of D: v^T Av, rows 64-127: Av^T Av
                                                        umma_i8(d, mn_desc(sv + off, TILE_BYTES), db, i8_idesc(128, 64), accumulate);
                                                } else {
                                                        umma_i8(d, mn_desc(sv + off, ROW_BYTES), db, i8_idesc(64, 64), accumulate);
                                                        umma_i8(d + 64, db, db, i8_idesc(64, 64), accumulate);
                                                }
                                                accumulate = 1;
                                        }
                                        umma_commit(bar_empty + 8 * s);    // the stage is free once these MMAs have read it
                                }
                                umma_commit(bar_acc_full + 8 * b);
                        }
                }
                __syncwarp();
        } else {                                                   // ---- epilogue warps 0-3
                for (int e = 0; e < nepoch; e++) {
                        const int b = e & 1;
                        mbar_wait(bar_acc_full + 8 * b, ((uint32_t)e >> 1) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + b * BUF_COLS;
                        uint32_t r[64];
                        if (WIDE) {
                                // M = 128: row m of D in TMEM lane m
                                tmem_ld64(r, taddr);
                                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                                mbar_arrive(bar_acc_empty + 8 * b);
                                const int row = warp * 32 + lane;
                                fold_row(r, row & 63, true, acc + (row >> 6) * NP * NP, m);
                        } else {
                                // M = 64: row m of D in TMEM lane 32 (m / 16) + m % 16
                                tmem_ld64(r, taddr);
                                fold_row(r, warp * 16 + (lane & 15), lane < 16, acc, m);
                                tmem_ld64(r, taddr + 64);
                                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                                mbar_arrive(bar_acc_empty + 8 * b);
                                fold_row(r, warp * 16 + (lane & 15), lane < 16, acc + NP * NP, m);
                        }
                }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(2 * BUF_COLS));
        for (int e = threadIdx.x; e < 2 * NP * NP; e += THREADS)
                atomicAdd(&sums[e], (unsigned long long)mp_reduce(acc[e], m));
        if (fuse.counter && last_block_done(fuse.counter, gridDim.x)) {
                // every load of this block has completed: the stage buffers are free scratch
                u32 *scratch = reinterpret_cast<u32 *>(dyn_raw + (stage0 - smem_u32(dyn_raw)));
                small_body(fuse.n, NP, sums, fuse.mats, fuse.state, 0, m, scratch);
        }
}


// ------------------------------------------------------------------------------------------
// orthogonalize
// ------------------------------------------------------------------------------------------
constexpr int OT_ROWS = 128;                          // rows per tile = M of one instruction
constexpr int OT_BYTES = OT_ROWS * ROW_BYTES;         // 8 KB per operand
constexpr int OSTAGE_BYTES = 3 * OT_BYTES;            // v, Av, p
constexpr int OB_BYTES = 64 * ROW_BYTES;              // one B operand: K = 64 rows of 64 bytes

// K-major, SWIZZLE_64B: rows (M) 64 bytes apart, 8-row groups 512 bytes apart (SBO); `addr` may
// point 32 bytes into the row for the second K step
__device__ __forceinline__ uint64_t k_desc(uint32_t addr)
{
        uint64_t d = 0;
        d |= (uint64_t)((addr & 0x3FFFF) >> 4);
        d |= (uint64_t)1 << 16;
        d |= (uint64_t)(512 >> 4) << 32;
        d |= 1ull << 46;
        d |= 4ull << 61;
        return d;
}

// D = s32, A = u8 K-major, B = u8 Major-MN
__host__ __device__ constexpr uint32_t i8_idesc_kmaj_a(int M, int N)
{
        return (2u << 4) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// this thread's TMEM lane, CW (32 or 64) consecutive columns
template <int CW> __device__ __forceinline__ void tmem_ld_cols(uint32_t (&r)[CW], uint32_t taddr);
template <> __device__ __forceinline__ void tmem_ld_cols<64>(uint32_t (&r)[64], uint32_t taddr) { tmem_ld64(r, taddr); }
template <> __device__ __forceinline__ void tmem_ld_cols<32>(uint32_t (&r)[32], uint32_t taddr)
{
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
        return v;
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v)
{
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void tma_store_rows(const CUtensorMap *map, uint32_t src, int row0)
{
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     :: "l"(map), "r"(0), "r"(row0), "r"(src) : "memory");
}

// sum_b 2^(8b) r_b mod p for limb sums r_b <= 128 * 255^2 < 2^23 (K = 128 bytes per output):
// both halves fit 32 bits, so the recombination is two shifts-and-adds and one wide multiply-add
__device__ __forceinline__ u32 recombine23(u32 r0, u32 r1, u32 r2, u32 r3, const ModP &m)
{
        const u32 lo = r0 + (r1 << 8), hi = r2 + (r3 << 8);
        return mp_reduce((u64)hi * 65536ull + lo, m);
}

// CW = accumulator columns per epilogue warp (64: 8 warps, 32: 16 warps); S = tiles in flight;
// SH = epilogue warps wait with a suspend-time hint
template <int CW, int S, bool SH = false>
__global__ void __launch_bounds__((2 * 4 * 64 / CW + 3) * 32, 1)
k_ortho_umma(const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_av,
             const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_vout,
             const __grid_constant__ CUtensorMap map_pout, int64_t ntile, const u32 *__restrict__ mats, ModP m,
             const DevSmall *__restrict__ state, int force)
{
        pdl_prologue();
        constexpr int NP = 16;
        constexpr int EPI = 2 * 4 * 64 / CW;                        // epilogue warps
        constexpr int NTHR = (EPI + 3) * 32;                        // + producer, MMA issuer, store warp
        extern __shared__ uint8_t dyn_raw[];
        __shared__ __align__(8) uint64_t bars[3 * S + 4];
        __shared__ uint32_t tmem_slot;
        if (!force && !state->do_ortho) return;

        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const uint32_t bmat0 = (smem_u32(dyn_raw) + 1023u) & ~1023u;       // Bc, Bd, Bw
        const uint32_t stage0 = bmat0 + 3 * OB_BYTES;
        const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[S]);
        const uint32_t bar_out = smem_u32(&bars[2 * S]);
        const uint32_t bar_acc_full = smem_u32(&bars[3 * S]), bar_acc_empty = smem_u32(&bars[3 * S + 2]);

        const int64_t t_lo = ntile * blockIdx.x / gridDim.x, t_hi = ntile * (blockIdx.x + 1) / gridDim.x;
        const int my = (int)(t_hi - t_lo);

        // B operands: row k = 4i + a holds the 16 words (2^(8a) X[i][j] mod p), j = 0..15, with the
        // 16-byte chunks of a row XOR-swizzled like TMA SWIZZLE_64B does (chunk ^= (k >> 1) & 3)
        for (int e = threadIdx.x; e < 3 * 64 * NP; e += NTHR) {
                const int j = e & 15, k = (e >> 4) & 63, X = e >> 10;
                const int which = X == 0 ? MAT_C : (X == 1 ? MAT_VTAVD : MAT_WINV);
                const u32 x = mats[which * NP * NP + (k >> 2) * NP + j];
                const u32 w = mp_mul(x, mp_reduce(1ull << (8 * (k & 3)), m), m);
                const uint32_t off = X * OB_BYTES + k * ROW_BYTES + ((((uint32_t)j >> 2) ^ (((uint32_t)k >> 1) & 3u)) << 4) + (j & 3) * 4;
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(bmat0 + off), "r"(w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        u32 dmask = 0;                                              // bit j: d[j] != 0 (:478)
        for (int j = 0; j < NP; j++) dmask |= (mats[MAT_D * NP * NP + j] != 0 ? 1u : 0u) << j;

        if (threadIdx.x == 0) {
                for (int s = 0; s < S; s++) {
                        mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_out + 8 * s, EPI * 32);
                }
                for (int b = 0; b < 2; b++) { mbar_init(bar_acc_full + 8 * b, 1); mbar_init(bar_acc_empty + 8 * b, EPI * 32); }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == EPI) {
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_slot)));
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem = tmem_slot;

        if (warp == EPI) {
                if (lane == 0) {                                   // ---- TMA producer
                        for (int t = 0; t < my; t++) {
                                const int s = t % S;
                                mbar_wait(bar_empty + 8 * s, ((uint32_t)(t / S) & 1u) ^ 1u);
                                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                             :: "r"(bar_full + 8 * s), "r"(OSTAGE_BYTES) : "memory");
                                const int row0 = (int)((t_lo + t) * OT_ROWS);
                                const uint32_t sb = stage0 + s * OSTAGE_BYTES;
                                tma_rows(sb, &map_v, row0, bar_full + 8 * s);
                                tma_rows(sb + OT_BYTES, &map_av, row0, bar_full + 8 * s);
                                tma_rows(sb + 2 * OT_BYTES, &map_p, row0, bar_full + 8 * s);
                        }
                }
                __syncwarp();
        } else if (warp == EPI + 1) {
                if (lane == 0) {                                   // ---- MMA issuer
                        constexpr uint32_t idesc = i8_idesc_kmaj_a(128, 64);
                        for (int t = 0; t < my; t++) {
                                const int s = t % S, b = t & 1;
                                mbar_wait(bar_full + 8 * s, (uint32_t)(t / S) & 1u);
                                mbar_wait(bar_acc_empty + 8 * b, (((uint32_t)t >> 1) & 1u) ^ 1u);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                const uint32_t sv = stage0 + s * OSTAGE_BYTES, sp = sv + 2 * OT_BYTES;
                                const uint32_t dv = tmem + b * 128, dp = dv + 64;
                                const uint32_t Bc = bmat0, Bd = bmat0 + OB_BYTES, Bw = bmat0 + 2 * OB_BYTES;
                                // D_v = v c + p vtAvd ; D_p = v winv   (K = 64 bytes of a row = 2 steps of 32)
                                umma_i8(dv, k_desc(sv), mn_desc(Bc, ROW_BYTES), idesc, 0);
                                umma_i8(dv, k_desc(sv + 32), mn_desc(Bc + 32 * ROW_BYTES, ROW_BYTES), idesc, 1);
                                umma_i8(dv, k_desc(sp), mn_desc(Bd, ROW_BYTES), idesc, 1);
                                umma_i8(dv, k_desc(sp + 32), mn_desc(Bd + 32 * ROW_BYTES, ROW_BYTES), idesc, 1);
                                umma_i8(dp, k_desc(sv), mn_desc(Bw, ROW_BYTES), idesc, 0);
                                umma_i8(dp, k_desc(sv + 32), mn_desc(Bw + 32 * ROW_BYTES, ROW_BYTES), idesc, 1);
                                umma_commit(bar_acc_full + 8 * b);
                        }
                }
                __syncwarp();
        } else if (warp == EPI + 2) {
                if (lane == 0) {                                   // ---- TMA store of finished tiles
                        for (int t = 0; t < my; t++) {
                                const int s = t % S;
                                mbar_wait(bar_out + 8 * s, (uint32_t)(t / S) & 1u);
                                const int row0 = (int)((t_lo + t) * OT_ROWS);
                                const uint32_t sb = stage0 + s * OSTAGE_BYTES;
                                tma_store_rows(&map_vout, sb, row0);
                                tma_store_rows(&map_pout, sb + 2 * OT_BYTES, row0);
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                                mbar_arrive(bar_empty + 8 * s);
                        }
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
                __syncwarp();
        } else {                                                   // ---- epilogue warps
                // warp -> TMEM lane quarter q (hardware rule: warp % 4), output half (0: new v, 1: new p)
                // and, for CW = 32, which 32 of the 64 accumulator columns
                const int q = warp & 3, part = warp >> 2;
                const int half = part / (64 / CW), col0 = (part % (64 / CW)) * CW;
                const int row = q * 32 + lane;
                const uint32_t rsw = ((uint32_t)row >> 1) & 3u;
                for (int t = 0; t < my; t++) {
                        const int s = t % S, b = t & 1;
                        mbar_wait<SH>(bar_full + 8 * s, (uint32_t)(t / S) & 1u);           // the tile's bytes (TMA)
                        const uint32_t sb = stage0 + s * OSTAGE_BYTES + row * ROW_BYTES;
                        // base terms first: they do not depend on the products
                        uint4 base[CW / 16], alt[CW / 16];
#pragma unroll
                        for (int c = 0; c < CW / 16; c++) {
                                const uint32_t off = ((uint32_t)(col0 / 16 + c) ^ rsw) << 4;
                                if (half == 0) { base[c] = lds128(sb + off); alt[c] = lds128(sb + OT_BYTES + off); }
                                else { base[c] = lds128(sb + 2 * OT_BYTES + off); alt[c] = make_uint4(0, 0, 0, 0); }
                        }
                        mbar_wait<SH>(bar_acc_full + 8 * b, ((uint32_t)t >> 1) & 1u);      // its products
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        uint32_t r[CW];
                        tmem_ld_cols<CW>(r, tmem + ((uint32_t)(q * 32) << 16) + b * 128 + half * 64 + col0);
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(bar_acc_empty + 8 * b);
#pragma unroll
                        for (int c = 0; c < CW / 16; c++) {
                                // new v = v c + p vtAvd + (d ? Av : v);  new p = v winv + (d ? 0 : p)      (:484-489)
                                const u32 dm = dmask >> (col0 / 4 + 4 * c);
                                uint4 o;
                                o.x = mp_add(recombine23(r[16 * c + 0], r[16 * c + 1], r[16 * c + 2], r[16 * c + 3], m), (dm & 1u) ? alt[c].x : base[c].x, m);
                                o.y = mp_add(recombine23(r[16 * c + 4], r[16 * c + 5], r[16 * c + 6], r[16 * c + 7], m), (dm & 2u) ? alt[c].y : base[c].y, m);
                                o.z = mp_add(recombine23(r[16 * c + 8], r[16 * c + 9], r[16 * c + 10], r[16 * c + 11], m), (dm & 4u) ? alt[c].z : base[c].z, m);
                                o.w = mp_add(recombine23(r[16 * c + 12], r[16 * c + 13], r[16 * c + 14], r[16 * c + 15], m), (dm & 8u) ? alt[c].w : base[c].w, m);
                                const uint32_t off = ((uint32_t)(col0 / 16 + c) ^ rsw) << 4;
                                sts128(sb + (half ? 2 * OT_BYTES : 0) + off, o);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // tile writes -> TMA store
                        mbar_arrive(bar_out + 8 * s);
                }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == EPI) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem));
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encode_fn()
{
        // resolved once; initialisation of a function-local static is thread-safe (one context per thread
        // in a multi-GPU process, tools/mg_check.cu)
        static const EncodeTiled fn = []() -> EncodeTiled {
                void *p = nullptr;
                cudaDriverEntryPointQueryResult qr;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
                    qr == cudaDriverEntryPointSuccess)
                        return (EncodeTiled)p;
                return nullptr;
        }();
        return fn;
}

// rows x 64 bytes, box = TILE_ROWS rows; rows past the end read as zero (they add nothing)
bool row_map(CUtensorMap *map, const void *base, int64_t rows, int box_rows = TILE_ROWS)
{
        EncodeTiled fn = encode_fn();
        if (!fn) return false;
        cuuint64_t dims[2] = {(cuuint64_t)ROW_BYTES, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ROW_BYTES};
        cuuint32_t box[2] = {(cuuint32_t)ROW_BYTES, (cuuint32_t)box_rows}, estr[2] = {1, 1};
        return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr size_t DOTS_SMEM = (size_t)STAGES * 2 * TILE_BYTES + 1024;
constexpr int ORTHO_DEFAULT_VARIANT = 3;          // 16 epilogue warps, 7 tiles in flight: 2.63 ms on config 4 (variant 0: 3.5 ms)
constexpr size_t ortho_smem(int stages) { return (size_t)stages * OSTAGE_BYTES + 3 * OB_BYTES + 1024; }

int umma_mode()
{
        // default: M = 128 (one instruction for both products).  BLK_DENSE=umma64: two M = 64 instructions;
        // BLK_DENSE=mma: the IMMA kernels of dense_mma.cu; BLK_DENSE=cuda: the CUDA-core kernels of dense.cu
        static int mode = -1;
        if (mode < 0) {
                const char *e = getenv("BLK_DENSE");
                mode = 2;
                if (e && (e[0] == 'm' || e[0] == 'c')) mode = 0;
                else if (e && e[0] == 'u' && strlen(e) > 4 && e[4] == '6') mode = 1;
        }
        return mode;
}

}  // namespace

bool dense_umma_supported(int np, int64_t rows)
{
        return umma_mode() != 0 && np == 16 && rows > 0 && rows < (1ll << 31) - TILE_ROWS && encode_fn() != nullptr;
}

void dense_umma_prepare(int np)
{
        if (np != 16) return;
        cudaFuncSetAttribute(k_dots_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DOTS_SMEM);
        cudaFuncSetAttribute(k_dots_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DOTS_SMEM);
        cudaFuncSetAttribute(k_ortho_umma<64, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(5));
        cudaFuncSetAttribute(k_ortho_umma<64, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(7));
        cudaFuncSetAttribute(k_ortho_umma<32, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(5));
        cudaFuncSetAttribute(k_ortho_umma<32, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(7));
        cudaFuncSetAttribute(k_ortho_umma<64, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(5));
        cudaFuncSetAttribute(k_ortho_umma<64, 7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(7));
        cudaFuncSetAttribute(k_ortho_umma<32, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(5));
        cudaFuncSetAttribute(k_ortho_umma<32, 7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem(7));
}

// wide < 0: follow BLK_DENSE
int launch_dots_umma(int np, const ModP &m, int64_t rows, const u32 *v, const u32 *Av, u64 *sums,
                     const DevSmall *state, const SmallFuse &fuse, cudaStream_t st, int wide)
{
        if (np != 16) return -1;
        CUtensorMap mv, ma;
        if (!row_map(&mv, v, rows) || !row_map(&ma, Av, rows)) return -1;
        const int64_t ntile = (rows + TILE_ROWS - 1) / TILE_ROWS;
        const unsigned grid = (unsigned)(ntile < blk_sm_count() ? ntile : blk_sm_count());
        if (wide < 0) wide = umma_mode() == 2;
        if (wide)
                launch_k(k_dots_umma<true>, grid, THREADS, DOTS_SMEM, st, mv, ma, ntile, (unsigned long long *)sums, m, state, fuse);
        else
                launch_k(k_dots_umma<false>, grid, THREADS, DOTS_SMEM, st, mv, ma, ntile, (unsigned long long *)sums, m, state, fuse);
        return 1;
}

// variant: 0..3 = (8 epilogue warps, 5 tiles in flight), (8, 7), (16, 5), (16, 7); 4..7 = the same with a
// suspend-time hint on the epilogue waits (not yet measured); < 0: the default
int launch_ortho_umma(int np, const ModP &m, int64_t rows, u32 *v, const u32 *Av, u32 *p, u32 *v_out, u32 *p_out,
                      const u32 *mats, const DevSmall *state, int force, cudaStream_t st, int variant)
{
        if (np != 16) return -1;
        CUtensorMap mv, ma, mp, mvo, mpo;
        if (!row_map(&mv, v, rows, OT_ROWS) || !row_map(&ma, Av, rows, OT_ROWS) || !row_map(&mp, p, rows, OT_ROWS) ||
            !row_map(&mvo, v_out, rows, OT_ROWS) || !row_map(&mpo, p_out, rows, OT_ROWS))
                return -1;
        const int64_t ntile = (rows + OT_ROWS - 1) / OT_ROWS;
        const unsigned grid = (unsigned)(ntile < blk_sm_count() ? ntile : blk_sm_count());
        if (variant < 0) variant = ORTHO_DEFAULT_VARIANT;
        switch (variant) {
        case 0: launch_k(k_ortho_umma<64, 5>, grid, 11 * 32, ortho_smem(5), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        case 1: launch_k(k_ortho_umma<64, 7>, grid, 11 * 32, ortho_smem(7), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        case 2: launch_k(k_ortho_umma<32, 5>, grid, 19 * 32, ortho_smem(5), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        case 3: launch_k(k_ortho_umma<32, 7>, grid, 19 * 32, ortho_smem(7), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        case 4: launch_k(k_ortho_umma<64, 5, true>, grid, 11 * 32, ortho_smem(5), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        case 5: launch_k(k_ortho_umma<64, 7, true>, grid, 11 * 32, ortho_smem(7), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        case 6: launch_k(k_ortho_umma<32, 5, true>, grid, 19 * 32, ortho_smem(5), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        default: launch_k(k_ortho_umma<32, 7, true>, grid, 19 * 32, ortho_smem(7), st, mv, ma, mp, mvo, mpo, ntile, mats, m, state, force); break;
        }
        return 1;
}
