// umma_i8_probe.cu -- known-answer probe for the planned tcgen05 dense kernels (dev tool, round-2 prep).
//
// Claim under test ("Next round"): a block of raw rows -- 64 bytes of
// (column, limb) per row, rows 64 B apart -- loaded by TMA with SWIZZLE_64B is a valid Major-MN
// operand of tcgen05.mma kind::i8, for A and for B, so that
//        D[m][n] = sum_k A_rows[k][m] * B_rows[k][n]          (u8 x u8 -> s32, M = N = 64, K = 32)
// needs no byte transposes.  The probe loads 32 rows of A and B, issues ONE UMMA, reads all 128
// TMEM lanes x 64 columns back and reports where (which lane/column) every expected value landed.
// All waits are bounded, so a wrong descriptor ends in a report, not in a hang.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
{
        for (int spin = 0; spin < 2000000; spin++) {
                uint32_t ok;
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                if (ok) return true;
        }
        return false;
}

__global__ void __launch_bounds__(128)
k_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int *out, int *status)
{
        __shared__ __align__(1024) uint8_t sA[32 * 64];
        __shared__ __align__(1024) uint8_t sB[32 * 64];
        __shared__ __align__(8) uint64_t bar_tma, bar_mma;
        __shared__ uint32_t tmem_base;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

        if (warp == 0) {
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base)));
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        if (threadIdx.x == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_tma)));
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_mma)));
                asm volatile("fence.mbarrier_init.release.cluster;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t tmem = tmem_base;

        if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar_tma)), "r"(2 * 32 * 64));
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             :: "r"(smem_u32(sA)), "l"(&mapA), "r"(0), "r"(0), "r"(smem_u32(&bar_tma)) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             :: "r"(smem_u32(sB)), "l"(&mapB), "r"(0), "r"(0), "r"(smem_u32(&bar_tma)) : "memory");
        }
        bool ok = mbar_wait(smem_u32(&bar_tma), 0);
        if (!ok) { if (threadIdx.x == 0) status[0] = 1; }          // TMA never completed
        __syncthreads();

        if (ok && threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                // shared-memory matrix descriptors: Major-MN, SWIZZLE_64B canonical layout
                //   start address >> 4 | LBO >> 4 at bit 16 | SBO >> 4 at bit 32 | version 1 at bit 46 | layout 4 at bit 61
                auto desc = [](uint32_t addr) {
                        uint64_t d = 0;
                        d |= (uint64_t)((addr & 0x3FFFF) >> 4);
                        d |= (uint64_t)(64 >> 4) << 16;            // LBO: next 64-byte MN atom (unused, a single atom)
                        d |= (uint64_t)(512 >> 4) << 32;           // SBO: next group of 8 rows (K)
                        d |= 1ull << 46;
                        d |= 4ull << 61;
                        return d;
                };
                const uint64_t da = desc(smem_u32(sA)), db = desc(smem_u32(sB));
                // instruction descriptor: D = s32, A/B = unsigned 8 bit, both Major-MN, N = 64, M = 64
                const uint32_t idesc = (2u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             :: "r"(smem_u32(&bar_mma)) : "memory");
        }
        bool ok2 = ok && mbar_wait(smem_u32(&bar_mma), 0);
        if (ok && !ok2 && threadIdx.x == 0) status[0] = 2;        // MMA never committed
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (ok2) {
                // warp w may read TMEM lanes 32w .. 32w+31; thread `lane` gets one lane, 64 columns
                uint32_t r[64];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
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
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                for (int c = 0; c < 64; c++) out[(warp * 32 + lane) * 64 + c] = (int)r[c];
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem));
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
        const int K = 32, MN = 64;
        uint8_t hA[K * MN], hB[K * MN];
        srand(7);
        for (int i = 0; i < K * MN; i++) { hA[i] = (uint8_t)(rand() & 0xff); hB[i] = (uint8_t)(rand() & 0xff); }
        int expect[MN][MN];
        for (int m = 0; m < MN; m++)
                for (int n = 0; n < MN; n++) {
                        int s = 0;
                        for (int k = 0; k < K; k++) s += (int)hA[k * MN + m] * (int)hB[k * MN + n];
                        expect[m][n] = s;
                }
        uint8_t *dA, *dB; int *dout, *dstat;
        CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB)));
        CK(cudaMalloc(&dout, 128 * 64 * 4)); CK(cudaMalloc(&dstat, 4));
        CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
        CK(cudaMemset(dout, 0xee, 128 * 64 * 4)); CK(cudaMemset(dstat, 0, 4));

        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
        CUtensorMap mA, mB;
        cuuint64_t dims[2] = {(cuuint64_t)MN, (cuuint64_t)K}, strides[1] = {(cuuint64_t)MN};
        cuuint32_t box[2] = {(cuuint32_t)MN, (cuuint32_t)K}, estr[2] = {1, 1};
        for (int w = 0; w < 2; w++) {
                CUresult r = ((EncodeTiled)fn)(w ? &mB : &mA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, w ? (void *)dB : (void *)dA, dims, strides, box, estr,
                                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
        }
        k_probe<<<1, 128>>>(mA, mB, dout, dstat);
        cudaError_t e = cudaDeviceSynchronize();
        printf("kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        static int hout[128 * 64]; int hstat = 0;
        CK(cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&hstat, dstat, 4, cudaMemcpyDeviceToHost));
        printf("status %d (0 ok, 1 TMA timeout, 2 MMA timeout)\n", hstat);
        if (hstat) return 1;
        // where did D land?
        // M = 64 accumulators use 16 lanes of each of the four 32-lane sub-partitions:
        //   row m -> lane 32*(m/16) + m%16
        int direct = 0, transposed = 0, found_anywhere = 0, quad = 0;
        for (int m = 0; m < MN; m++)
                for (int n = 0; n < MN; n++) {
                        direct += hout[m * 64 + n] == expect[m][n];
                        transposed += hout[n * 64 + m] == expect[m][n];
                        quad += hout[(32 * (m / 16) + m % 16) * 64 + n] == expect[m][n];
                }
        printf("D[m][n] at (lane m, column n): %d / 4096   at (lane n, column m): %d / 4096   at (lane 32*(m/16)+m%%16, column n): %d / 4096\n",
               direct, transposed, quad);
        for (int m = 15; m < 18; m++)
                for (int n = 0; n < 1; n++) {
                        for (int l = 0; l < 128; l++)
                                for (int c = 0; c < 64; c++)
                                        if (hout[l * 64 + c] == expect[m][n]) { printf("  expect[%d][%d] = %d found at lane %d column %d\n", m, n, expect[m][n], l, c); found_anywhere++; }
                }
        printf("lanes holding data (not 0xeeeeeeee): ");
        for (int l = 0; l < 128; l++) if (hout[l * 64] != (int)0xeeeeeeee) printf("%d ", l);
        printf("\n%s\n", quad == 4096 ? "PASS: raw rows + SWIZZLE_64B are valid Major-MN i8 operands; D row m sits in TMEM lane 32*(m/16) + m%16" :
                                        "MISMATCH: see the placement report above");
        return 0;
}
