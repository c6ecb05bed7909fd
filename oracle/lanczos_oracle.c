/*
 * oracle/lanczos_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the block-Lanczos mod-p hot path of the
 * reference (T-amairi/block-lanczos-algorithm-parallelization,
 * sequential/lanczos_modp.c).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file; the
 * CUDA product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks every function
 * here (a) against the committed golden vectors in tests/golden/ that were
 * produced by executing the unmodified reference (tests/golden/make_golden.py)
 * and (b) when oracle/_ref/libref_seq.so is present, directly against the
 * reference's own object code on fresh random inputs.
 *
 * The reference keeps `n` (blocking factor) and `prime` as globals
 * (sequential/lanczos_modp.c:39-40); here they are explicit arguments.  All
 * values are canonical residues in [0,p) held in uint32_t, exactly like the
 * reference (every store there is `... % prime`).  Because of that, the
 * order of summation is free: any mathematically correct evaluation mod p is
 * bit-identical to the reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t u32;
typedef uint64_t u64;

/* ------------------------------------------------------------------ RNG
 * xoshiro256+ with the reference's fixed seed
 * (sequential/lanczos_modp.c:64-87).  The Lanczos start block is
 * v[i] = next() % p for i row-major over N*n (ibid. :624-625).          */
typedef struct { u64 s[4]; } orc_rng_t;

static inline u64 rot_left(u64 w, unsigned k) { return (w << k) | (w >> (64u - k)); }

void orc_rng_seed(orc_rng_t *g)
{
        g->s[0] = 0x1415926535ull; g->s[1] = 0x8979323846ull;
        g->s[2] = 0x2643383279ull; g->s[3] = 0x5028841971ull;
}

u64 orc_rng_next(orc_rng_t *g)
{
        u64 *s = g->s;
        u64 out = rot_left(s[0] + s[3], 23) + s[0];
        u64 shifted = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= shifted;
        s[3] = rot_left(s[3], 45);
        return out;
}

/* fill dst[0..count) with the reference's start values for prime p */
void orc_start_block(u32 *dst, long count, u64 p)
{
        orc_rng_t g;
        orc_rng_seed(&g);
        for (long t = 0; t < count; t++)
                dst[t] = (u32)(orc_rng_next(&g) % p);
}

/* ------------------------------------------------------------------ SpMV
 * y <- M x (transpose == 0) or y <- M^T x (transpose != 0) on n-wide
 * row-major blocks; COO input in file order.
 * Restates sparse_matrix_vector_product, sequential/lanczos_modp.c:266-287:
 * only the first rows*n entries of y are cleared and written.            */
void orc_spmv(u32 *y, int nrows, int ncols, long nnz,
              const int *Mi, const int *Mj, const u32 *Mx,
              const u32 *x, int transpose, int n, u64 p)
{
        long out_rows = transpose ? ncols : nrows;
        memset(y, 0, sizeof(u32) * (size_t)out_rows * (size_t)n);
        for (long e = 0; e < nnz; e++) {
                long r = transpose ? Mj[e] : Mi[e];
                long c = transpose ? Mi[e] : Mj[e];
                u64 a = Mx[e];
                u32 *yr = y + r * n;
                const u32 *xc = x + c * n;
                for (int l = 0; l < n; l++)
                        yr[l] = (u32)((yr[l] + a * xc[l]) % p);
        }
}

/* ------------------------------------------------------------------ n x n
 * C <- C + A B and C <- C + A^T B (sequential/lanczos_modp.c:292-315).   */
static void small_acc_AB(u32 *C, const u32 *A, const u32 *B, int n, u64 p)
{
        for (int r = 0; r < n; r++)
                for (int c = 0; c < n; c++) {
                        u64 s = C[r * n + c];
                        for (int k = 0; k < n; k++)
                                s = (s + (u64)A[r * n + k] * B[k * n + c]) % p;
                        C[r * n + c] = (u32)s;
                }
}

void orc_matmul_CpAB(u32 *C, const u32 *A, const u32 *B, int n, u64 p)
{
        small_acc_AB(C, A, B, n, p);
}

void orc_matmul_CpAtB(u32 *C, const u32 *A, const u32 *B, int n, u64 p)
{
        for (int r = 0; r < n; r++)
                for (int c = 0; c < n; c++) {
                        u64 s = C[r * n + c];
                        for (int k = 0; k < n; k++)
                                s = (s + (u64)A[k * n + r] * B[k * n + c]) % p;
                        C[r * n + c] = (u32)s;
                }
}

/* a^-1 mod m by the extended Euclidean algorithm
 * (invmod, sequential/lanczos_modp.c:318-336); result canonical in [0,m). */
u32 orc_invmod(u32 a, u32 m)
{
        int64_t r0 = m, r1 = a % m, t0 = 0, t1 = 1;
        while (r1) {
                int64_t q = r0 / r1, w;
                w = r0 - q * r1; r0 = r1; r1 = w;
                w = t0 - q * t1; t0 = t1; t1 = w;
        }
        return (u32)(t0 < 0 ? t0 + m : t0);
}

/* ------------------------------------------------------------------ dots
 * vtAv <- v^T Av, vtAAv <- Av^T Av  (block_dot_products,
 * sequential/lanczos_modp.c:443-453).  The reference walks n rows at a time
 * up to ceil(N/n)*n and relies on zero padding; summing rows [0,N) is the
 * same value mod p.                                                       */
void orc_block_dot_products(u32 *vtAv, u32 *vtAAv, long N,
                            const u32 *Av, const u32 *v, int n, u64 p)
{
        for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                        u64 s1 = 0, s2 = 0;
                        for (long r = 0; r < N; r++) {
                                u64 avj = Av[r * n + j];
                                s1 = (s1 + (u64)v[r * n + i] * avj) % p;
                                s2 = (s2 + (u64)Av[r * n + i] * avj) % p;
                        }
                        vtAv[i * n + j] = (u32)s1;
                        vtAAv[i * n + j] = (u32)s2;
                }
}

/* ------------------------------------------------------------------ semi-inverse
 * One Gauss-Jordan sweep as done twice by semi_inverse
 * (sequential/lanczos_modp.c:351-382 and :393-436): for each column j take
 * the FIRST row i >= j with a non-zero entry, scale that row to make the
 * pivot 1, swap it into row j, clear column j elsewhere.  `W` (may be NULL)
 * receives the same row operations.  Returns the number of pivots and sets
 * d[j] = 1 exactly for pivot columns.                                      */
static int gauss_jordan_sweep(u32 *M, u32 *W, u32 *d, int n, u64 p)
{
        int found = 0;
        memset(d, 0, sizeof(u32) * (size_t)n);
        for (int j = 0; j < n; j++) {
                int piv = -1;
                for (int i = j; i < n && piv < 0; i++)
                        if (M[i * n + j]) piv = i;
                if (piv < 0) continue;
                d[j] = 1; found++;
                u64 inv = orc_invmod(M[piv * n + j], (u32)p);
                for (int k = 0; k < n; k++) {
                        u32 a = (u32)(((u64)M[piv * n + k] * inv) % p);
                        M[piv * n + k] = M[j * n + k];
                        M[j * n + k] = a;
                        if (W) {
                                u32 b = (u32)(((u64)W[piv * n + k] * inv) % p);
                                W[piv * n + k] = W[j * n + k];
                                W[j * n + k] = b;
                        }
                }
                for (int i = 0; i < n; i++) {
                        if (i == j) continue;
                        u64 neg = p - M[i * n + j];       /* in [1,p] like the reference */
                        for (int k = 0; k < n; k++) {
                                M[i * n + k] = (u32)((M[i * n + k] + neg * M[j * n + k]) % p);
                                if (W)
                                        W[i * n + k] = (u32)((W[i * n + k] + neg * W[j * n + k]) % p);
                        }
                }
        }
        return found;
}

/* semi_inverse, sequential/lanczos_modp.c:342-438.  Returns #pivots of the
 * second sweep; winv and d as in the reference.                           */
int orc_semi_inverse(const u32 *U, u32 *winv, u32 *d, int n, u64 p)
{
        u32 *work = malloc(sizeof(u32) * (size_t)n * (size_t)n);
        memcpy(work, U, sizeof(u32) * (size_t)n * (size_t)n);
        gauss_jordan_sweep(work, NULL, d, n, p);                  /* phase 1: d */
        for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                        int keep = d[i] && d[j];
                        work[i * n + j] = keep ? U[i * n + j] : 0;
                        winv[i * n + j] = (i == j && d[i]) ? 1 : 0;
                }
        int npiv = gauss_jordan_sweep(work, winv, d, n, p);       /* phase 2 */
        free(work);
        return npiv;
}

/* ------------------------------------------------------------------ orthogonalize
 * orthogonalize, sequential/lanczos_modp.c:456-492.  Next v goes to tmp
 * rows [0,N), p is updated in place.  c and vtAvd are formed as in the
 * reference (entries may equal p; harmless as multipliers).               */
void orc_orthogonalize(const u32 *v, u32 *tmp, u32 *pp, const u32 *d,
                       const u32 *vtAv, const u32 *vtAAv, const u32 *winv,
                       long N, const u32 *Av, int n, u64 p)
{
        size_t nn = (size_t)n * (size_t)n;
        u32 *c = calloc(nn, sizeof(u32));
        u32 *spl = malloc(nn * sizeof(u32));
        u32 *nvd = malloc(nn * sizeof(u32));
        for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                        spl[i * n + j] = d[j] ? vtAAv[i * n + j] : vtAv[i * n + j];
                        nvd[i * n + j] = d[j] ? (u32)(p - vtAv[i * n + j]) : 0;
                }
        small_acc_AB(c, winv, spl, n, p);
        for (size_t t = 0; t < nn; t++) c[t] = (u32)(p - c[t]);

        for (long r = 0; r < N; r++) {
                const u32 *vr = v + r * n, *ar = Av + r * n;
                u32 *pr = pp + r * n, *tr = tmp + r * n;
                for (int j = 0; j < n; j++) {
                        u64 nv = d[j] ? ar[j] : vr[j];
                        u64 np = d[j] ? 0 : pr[j];
                        for (int k = 0; k < n; k++) {
                                nv = (nv + (u64)vr[k] * c[k * n + j]) % p;
                                nv = (nv + (u64)pr[k] * nvd[k * n + j]) % p;
                                np = (np + (u64)vr[k] * winv[k * n + j]) % p;
                        }
                        tr[j] = (u32)nv;
                        /* p row is read for every j: stage the new row */
                        spl[j] = (u32)np;   /* spl is free from here on (n <= n*n) */
                }
                for (int j = 0; j < n; j++) pr[j] = spl[j];
        }
        free(c); free(spl); free(nvd);
}

/* ------------------------------------------------------------------ whole loop
 * block_lanczos main loop, sequential/lanczos_modp.c:631-659, on caller-
 * allocated blocks of `pad` u32 each (pad = max(ceil(N/n)n, ceil(Mc/n)n)*n,
 * ibid. :594-597) that already hold the state (fresh start: all zero except
 * v = orc_start_block).  Runs until `stop_after` total iterations are
 * reached (if > 0) or the semi-inverse finds no pivot.  *iters is in/out.
 * Returns 1 if it stopped on "no pivot", else 0.                           */
int orc_lanczos_run(int nrows, int ncols, long nnz,
                    const int *Mi, const int *Mj, const u32 *Mx,
                    int n, u64 p, int right_kernel, int stop_after,
                    u32 *v, u32 *tmp, u32 *Av, u32 *pp, int *iters)
{
        long N = right_kernel ? ncols : nrows;
        size_t nn = (size_t)n * (size_t)n;
        u32 *vtAv = malloc(nn * 4), *vtAAv = malloc(nn * 4), *winv = malloc(nn * 4);
        u32 *d = malloc((size_t)n * 4);
        int no_pivot = 0;
        for (;;) {
                if (stop_after > 0 && *iters == stop_after) break;
                orc_spmv(tmp, nrows, ncols, nnz, Mi, Mj, Mx, v, !right_kernel, n, p);
                orc_spmv(Av, nrows, ncols, nnz, Mi, Mj, Mx, tmp, right_kernel, n, p);
                orc_block_dot_products(vtAv, vtAAv, N, Av, v, n, p);
                if (orc_semi_inverse(vtAv, winv, d, n, p) == 0) { no_pivot = 1; break; }
                orc_orthogonalize(v, tmp, pp, d, vtAv, vtAAv, winv, N, Av, n, p);
                memcpy(v, tmp, sizeof(u32) * (size_t)N * (size_t)n);
                ++*iters;
        }
        free(vtAv); free(vtAAv); free(winv); free(d);
        return no_pivot;
}
