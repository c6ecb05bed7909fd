// umma_dense_test.cu -- dev tool: the tcgen05 kernels of csrc/dense_umma.cu against the IMMA kernels of
// csrc/dense_mma.cu (the validated default path) on the same device buffers, bit for bit, plus timings.
// Starts in milliseconds (no Python), so it is what a short GPU slot is spent on.
//
//      umma_dense_test dots64|dots128|ortho [rows ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "blk_internal.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_fill(u32 *x, int64_t count, u32 p, u64 seed)
{
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
                u64 z = (u64)i * 0x9E3779B97F4A7C15ull + seed;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                x[i] = (u32)(z % p);
        }
}

// mismatch statistics of two blocks: total, by column, by row mod 8, first position
__global__ void k_diff(const u32 *a, const u32 *b, int64_t rows, unsigned long long *stat)
{
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * 16; i += (int64_t)gridDim.x * blockDim.x)
                if (a[i] != b[i]) {
                        atomicAdd(&stat[0], 1ull);
                        atomicAdd(&stat[1 + (i & 15)], 1ull);
                        atomicAdd(&stat[17 + ((i >> 4) & 7)], 1ull);
                        atomicMin(&stat[25], (unsigned long long)i);
                }
}

static u32 *dalloc(int64_t rows) { u32 *p; CK(cudaMalloc(&p, (size_t)rows * 64 + 64)); return p; }

static void fill(u32 *x, int64_t rows, u32 p, u64 seed)
{
        k_fill<<<148 * 8, 256>>>(x, rows * 16, p, seed);
        CK(cudaGetLastError());
}

static bool diff(const char *what, const u32 *a, const u32 *b, int64_t rows)
{
        unsigned long long *st, h[26];
        CK(cudaMalloc(&st, sizeof(h)));
        CK(cudaMemset(st, 0, sizeof(h)));
        CK(cudaMemset(st + 25, 0xff, 8));
        k_diff<<<148 * 8, 256>>>(a, b, rows, st);
        CK(cudaMemcpy(h, st, sizeof(h), cudaMemcpyDeviceToHost));
        CK(cudaFree(st));
        if (!h[0]) { printf("  %-22s identical (%lld rows)\n", what, (long long)rows); return true; }
        printf("  %-22s MISMATCH %llu of %lld words; first at row %llu col %llu\n     by column:", what, h[0], (long long)rows * 16, h[25] >> 4, h[25] & 15);
        for (int j = 0; j < 16; j++) printf(" %llu", h[1 + j]);
        printf("\n     by row mod 8:");
        for (int j = 0; j < 8; j++) printf(" %llu", h[17 + j]);
        printf("\n");
        u32 ra[16], rb[16];
        int64_t r = (int64_t)(h[25] >> 4);
        CK(cudaMemcpy(ra, a + r * 16, 64, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rb, b + r * 16, 64, cudaMemcpyDeviceToHost));
        printf("     want:"); for (int j = 0; j < 16; j++) printf(" %08x", ra[j]);
        printf("\n     got: "); for (int j = 0; j < 16; j++) printf(" %08x", rb[j]);
        printf("\n");
        return false;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

static int test_dots(int wide, const std::vector<int64_t> &sizes)
{
        int bad = 0;
        const u32 primes[3] = {2147483647u, 65537u, 1073741789u};
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        u64 *sums; CK(cudaMalloc(&sums, 3 * 512 * 8));
        for (int64_t rows : sizes) {
                u32 *v = dalloc(rows), *Av = dalloc(rows);
                for (int pi = 0; pi < 3; pi++) {
                        ModP m; modp_make(&m, primes[pi]);
                        fill(v, rows, m.p, 1 + pi); fill(Av, rows, m.p, 77 + pi);
                        CK(cudaMemset(sums, 0, 3 * 512 * 8));
                        SmallFuse none;
                        launch_dots_mma(16, m, rows, v, Av, sums, nullptr, none, 0);
                        CK(cudaDeviceSynchronize());
                        if (launch_dots_umma(16, m, rows, v, Av, sums + 512, nullptr, none, 0, wide) != 1) { printf("FAIL launch_dots_umma refused\n"); return 1; }
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("FAIL dots kernel (rows %lld): %s\n", (long long)rows, cudaGetErrorString(e)); return 1; }
                        u64 h[1024];
                        CK(cudaMemcpy(h, sums, sizeof(h), cudaMemcpyDeviceToHost));
                        int wrong = 0, first = -1;
                        for (int i = 0; i < 512; i++)
                                if (h[i] % m.p != h[512 + i] % m.p) { if (first < 0) first = i; wrong++; }
                        printf("  dots %s rows %lld p %u: %s", wide ? "M=128" : "2xM=64", (long long)rows, m.p, wrong ? "MISMATCH" : "identical");
                        if (wrong) {
                                printf(" (%d of 512; first: matrix %d i %d j %d want %llu got %llu)", wrong, first >> 8, (first >> 4) & 15, first & 15,
                                       (unsigned long long)(h[first] % m.p), (unsigned long long)(h[512 + first] % m.p));
                                int by_m[2] = {0, 0};
                                for (int i = 0; i < 512; i++) by_m[i >> 8] += h[i] % m.p != h[512 + i] % m.p;
                                printf(" vtAv %d vtAAv %d", by_m[0], by_m[1]);
                                bad++;
                        }
                        printf("\n");
                        if (pi == 0 && rows >= 1000000) {
                                for (int which = 0; which < 2; which++) {
                                        float best = 1e9f;
                                        for (int rep = 0; rep < 5; rep++) {
                                                CK(cudaEventRecord(e0));
                                                if (which) launch_dots_umma(16, m, rows, v, Av, sums + 1024, nullptr, none, 0, wide);
                                                else launch_dots_mma(16, m, rows, v, Av, sums + 1024, nullptr, none, 0);
                                                CK(cudaEventRecord(e1));
                                                float ms = time_ms(e0, e1);
                                                if (rep && ms < best) best = ms;
                                        }
                                        printf("    %-8s %.3f ms  (%.0f GB/s of v + Av)\n", which ? "tcgen05" : "IMMA", best, rows * 128.0 / best * 1e-6);
                                }
                        }
                }
                CK(cudaFree(v)); CK(cudaFree(Av));
        }
        // fused n x n stage: the last block runs small_body; outputs must match the IMMA kernel's
        {
                const int64_t rows = 200000;
                ModP m; modp_make(&m, 2147483647u);
                u32 *v = dalloc(rows), *Av = dalloc(rows);
                fill(v, rows, m.p, 5); fill(Av, rows, m.p, 6);
                u32 *mats[2]; DevSmall *st[2]; unsigned *ctr; u64 *s2;
                CK(cudaMalloc(&ctr, 8)); CK(cudaMemset(ctr, 0, 8));
                CK(cudaMalloc(&s2, 2 * 512 * 8)); CK(cudaMemset(s2, 0, 2 * 512 * 8));
                const size_t words = mats_words(16);
                std::vector<u32> hm[2]; DevSmall hs[2];
                for (int w = 0; w < 2; w++) {
                        CK(cudaMalloc(&mats[w], words * 4)); CK(cudaMemset(mats[w], 0, words * 4));
                        CK(cudaMalloc(&st[w], sizeof(DevSmall))); CK(cudaMemset(st[w], 0, sizeof(DevSmall)));
                        SmallFuse f; f.counter = ctr + w; f.mats = mats[w]; f.state = st[w]; f.n = 16;
                        if (w) launch_dots_umma(16, m, rows, v, Av, s2 + 512, st[w], f, 0, wide);
                        else launch_dots_mma(16, m, rows, v, Av, s2, st[w], f, 0);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("FAIL fused dots kernel %d: %s\n", w, cudaGetErrorString(e)); return 1; }
                        hm[w].resize(words);
                        CK(cudaMemcpy(hm[w].data(), mats[w], words * 4, cudaMemcpyDeviceToHost));
                        CK(cudaMemcpy(&hs[w], st[w], sizeof(DevSmall), cudaMemcpyDeviceToHost));
                }
                bool same = hm[0] == hm[1] && !memcmp(&hs[0], &hs[1], sizeof(DevSmall));
                printf("  fused n x n stage (mats, bfrag, loop state; iters %d npiv %d do_ortho %d): %s\n", hs[1].iters, hs[1].npiv, hs[1].do_ortho,
                       same ? "identical" : "MISMATCH");
                bad += !same;
        }
        return bad;
}

static int test_ortho(const std::vector<int64_t> &sizes)
{
        int bad = 0;
        const u32 primes[3] = {2147483647u, 65537u, 1073741789u};
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const size_t words = mats_words(16);
        u32 *mats; CK(cudaMalloc(&mats, words * 4));
        for (int64_t rows : sizes) {
                u32 *v = dalloc(rows), *Av = dalloc(rows), *p = dalloc(rows);
                u32 *vo[2] = {dalloc(rows), dalloc(rows)}, *po[2] = {dalloc(rows), dalloc(rows)};
                for (int pi = 0; pi < 3; pi++) {
                        ModP m; modp_make(&m, primes[pi]);
                        fill(v, rows, m.p, 11 + pi); fill(Av, rows, m.p, 22 + pi); fill(p, rows, m.p, 33 + pi);
                        // coefficient matrices: random residues; d = 1 on columns j % 3 != 0; bfrag for the IMMA kernel
                        std::vector<u32> hm(words, 0);
                        u64 z = 12345 + pi;
                        for (int X = 0; X < MAT_COUNT; X++)
                                for (int e = 0; e < 256; e++) { z = z * 6364136223846793005ull + 1442695040888963407ull; hm[X * 256 + e] = (u32)((z >> 20) % m.p); }
                        for (int j = 0; j < 16; j++) hm[MAT_D * 256 + j] = j % 3 != 0;
                        u32 pw[4];
                        for (int be = 0; be < 4; be++) pw[be] = (u32)((1ull << (8 * be)) % m.p);
                        u32 *bf = hm.data() + MAT_COUNT * 256;
                        for (int idx = 0; idx < 12 * 2 * 2 * 64; idx++) {          // same index map as small_body.cuh
                                int h = idx & 1, lane = (idx >> 1) & 31, rest = idx >> 6;
                                int s = rest % 2; rest /= 2;
                                int t = rest % 2; rest /= 2;
                                int b = rest & 3, X = rest >> 2;
                                int j = 8 * t + (lane >> 2), w = (lane & 3) * 4 + 2 * s + h;
                                int which = X == 0 ? MAT_C : (X == 1 ? MAT_VTAVD : MAT_WINV);
                                u32 x = hm[which * 256 + w * 16 + j], word = 0;
                                for (int be = 0; be < 4; be++) word |= ((u32)(((u64)x * pw[be]) % m.p >> (8 * b)) & 0xffu) << (8 * be);
                                bf[idx] = word;
                        }
                        CK(cudaMemcpy(mats, hm.data(), words * 4, cudaMemcpyHostToDevice));
                        for (int w = 0; w < 2; w++) { CK(cudaMemset(vo[w], 0xcd, rows * 64)); CK(cudaMemset(po[w], 0xcd, rows * 64)); }
                        launch_ortho_mma(16, m, rows, v, Av, p, vo[0], po[0], mats, nullptr, 1, 0);
                        CK(cudaDeviceSynchronize());
                        printf(" ortho rows %lld p %u\n", (long long)rows, m.p);
                        for (int var = 0; var < 8; var++) {
                                CK(cudaMemset(vo[1], 0xcd, rows * 64)); CK(cudaMemset(po[1], 0xcd, rows * 64));
                                if (launch_ortho_umma(16, m, rows, v, Av, p, vo[1], po[1], mats, nullptr, 1, 0, var) != 1) { printf("FAIL launch_ortho_umma refused\n"); return 1; }
                                cudaError_t e = cudaDeviceSynchronize();
                                if (e != cudaSuccess) { printf("FAIL ortho kernel variant %d (rows %lld): %s\n", var, (long long)rows, cudaGetErrorString(e)); return 1; }
                                char what[64];
                                snprintf(what, sizeof what, "variant %d new v", var); bad += !diff(what, vo[0], vo[1], rows);
                                snprintf(what, sizeof what, "variant %d new p", var); bad += !diff(what, po[0], po[1], rows);
                        }
                        if (pi == 0 && rows >= 1000000) {
                                for (int which = -1; which < 8; which++) {
                                        float best = 1e9f;
                                        for (int rep = 0; rep < 5; rep++) {
                                                CK(cudaEventRecord(e0));
                                                if (which >= 0) launch_ortho_umma(16, m, rows, v, Av, p, vo[1], po[1], mats, nullptr, 1, 0, which);
                                                else launch_ortho_mma(16, m, rows, v, Av, p, vo[0], po[0], mats, nullptr, 1, 0);
                                                CK(cudaEventRecord(e1));
                                                float ms = time_ms(e0, e1);
                                                if (rep && ms < best) best = ms;
                                        }
                                        const char *names[9] = {"IMMA", "tcgen05 8 epilogue warps, 5 tiles", "tcgen05 8 warps, 7 tiles", "tcgen05 16 warps, 5 tiles", "tcgen05 16 warps, 7 tiles",
                                                                 "8 warps, 5 tiles, suspend hint", "8 warps, 7 tiles, suspend hint", "16 warps, 5 tiles, suspend hint", "16 warps, 7 tiles, suspend hint"};
                                        printf("    %-36s %.3f ms  (%.0f GB/s of 3 reads + 2 writes)\n", names[which + 1], best, rows * 320.0 / best * 1e-6);
                                }
                        }
                        // sustained, as the loop runs it: 40 launches back to back, average (not best-of), separate outputs vs in place
                        if (pi == 0 && rows >= 1000000) {
                                for (int inplace = 0; inplace < 2; inplace++) {
                                        launch_ortho_umma(16, m, rows, v, Av, p, inplace ? v : vo[1], inplace ? p : po[1], mats, nullptr, 1, 0);
                                        CK(cudaEventRecord(e0));
                                        for (int rep = 0; rep < 40; rep++)
                                                launch_ortho_umma(16, m, rows, v, Av, p, inplace ? v : vo[1], inplace ? p : po[1], mats, nullptr, 1, 0);
                                        CK(cudaEventRecord(e1));
                                        float ms = time_ms(e0, e1) / 40;
                                        printf("    default variant, 40 launches back to back, %-16s %.3f ms average  (%.0f GB/s)\n",
                                               inplace ? "in place:" : "separate outputs:", ms, rows * 320.0 / ms * 1e-6);
                                }
                        }
                        // in place, as the iteration calls it (v_out = v, p_out = p)
                        if (pi == 2) {
                                launch_ortho_umma(16, m, rows, v, Av, p, v, p, mats, nullptr, 1, 0);
                                CK(cudaDeviceSynchronize());
                                bad += !diff("new v (in place)", vo[0], v, rows);
                                bad += !diff("new p (in place)", po[0], p, rows);
                        }
                }
                CK(cudaFree(v)); CK(cudaFree(Av)); CK(cudaFree(p));
                for (int w = 0; w < 2; w++) { CK(cudaFree(vo[w])); CK(cudaFree(po[w])); }
        }
        return bad;
}

int main(int argc, char **argv)
{
        if (argc < 2) { printf("usage: %s dots64|dots128|ortho [rows ...]\n", argv[0]); return 2; }
        std::vector<int64_t> sizes;
        for (int i = 2; i < argc; i++) sizes.push_back(atoll(argv[i]));
        if (sizes.empty()) sizes = {37, 1000, 256 * 148 * 2 + 19, 50000000};
        setenv("BLK_DENSE", "umma", 1);
        if (!dense_umma_supported(16, 1000)) { printf("FAIL: cuTensorMapEncodeTiled not available\n"); return 1; }
        dense_mma_prepare(16);
        dense_umma_prepare(16);
        int bad;
        if (!strcmp(argv[1], "dots64")) bad = test_dots(0, sizes);
        else if (!strcmp(argv[1], "dots128")) bad = test_dots(1, sizes);
        else bad = test_ortho(sizes);
        printf("%s: %s\n", argv[1], bad ? "FAILED" : "PASS");
        return bad != 0;
}
