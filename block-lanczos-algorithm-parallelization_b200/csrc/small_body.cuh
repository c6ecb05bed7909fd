// small_body.cuh -- the n x n stage of one iteration (device code shared by k_small and by the
// dot-product kernels, whose last block to finish runs it: one launch less per iteration).
#pragma once
#include "blk_internal.cuh"

__host__ __device__ static inline size_t small_smem_words(int n) { return 6 * (size_t)n * n + 4 * (size_t)n + 8; }

// "last block done": every block calls this after its results are globally visible; exactly one
// block (the last to arrive) gets true and may consume everybody's results.
__device__ __forceinline__ bool last_block_done(unsigned *counter, unsigned nblocks)
{
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
                unsigned ticket = atomicAdd(counter, 1u);
                s_last = ticket == nblocks - 1;
                if (s_last) *counter = 0;              // re-arm for the next launch
        }
        __syncthreads();
        if (s_last) __threadfence();
        return s_last != 0;
}

// ------------------------------------------------------------------------------------------
// small: one block.  Working matrices use the true n as leading dimension in shared memory.
// ------------------------------------------------------------------------------------------

struct GjBuf { u32 *M, *W, *scal; };

// One Gauss-Jordan sweep of semi_inverse (sequential/lanczos_modp.c:351-382 and :393-436): for
// each column j the FIRST row i >= j with a non-zero entry is the pivot; the reference scales that
// row by the inverse of the pivot, swaps it into row j and clears column j in every other row.
//
// Done here fraction-free so that no modular inverse sits on the serial path: row i of the
// working matrices is kept as scal[i] times the reference's row (scal[i] != 0), the update is
//      row_i <- pv * row_i - row_i[j] * row_piv        scal[i] <- scal[i] * pv     (i != j)
//      row_j <- row_piv                                scal[j] <- pv
// Zero patterns -- and therefore every pivot decision -- are identical to the reference's, and
// dividing row i by scal[i] at the end (n independent inverses, one per thread) gives exactly
// the reference's W.  Double buffered in shared memory: one barrier per pivot column.
// Returns the number of pivots; d[j] = 1 on pivot columns; *cur = buffer holding the result.
static __device__ __noinline__ int gj_sweep(GjBuf *buf, int *cur, bool carry_w, u32 *d, int n, const ModP &m)
{
        const int tid = threadIdx.x, SMALL_TB = blockDim.x;
        int found = 0, s = *cur;
        for (int j = 0; j < n; j++) {
                const u32 *Ms = buf[s].M, *Ws = buf[s].W, *ss = buf[s].scal;
                u32 *Md = buf[s ^ 1].M, *Wd = buf[s ^ 1].W, *sd = buf[s ^ 1].scal;
                int piv = -1;
                for (int i = j; i < n; i++)
                        if (Ms[i * n + j] != 0) { piv = i; break; }       // same answer in every thread
                if (tid == 0) d[j] = piv >= 0;
                if (piv < 0) continue;
                found++;
                const u32 pv = Ms[piv * n + j];
                for (int e = tid; e < n * n; e += SMALL_TB) {
                        const int i = e / n, k = e - i * n;
                        if (i == j) {
                                Md[e] = Ms[piv * n + k];
                                if (carry_w) Wd[e] = Ws[piv * n + k];
                        } else {
                                const int r = (i == piv) ? j : i;           // row i after the swap
                                const u32 f = mp_neg(Ms[r * n + j], m);
                                Md[e] = mp_reduce((u64)pv * Ms[r * n + k] + (u64)f * Ms[piv * n + k], m);
                                if (carry_w) Wd[e] = mp_reduce((u64)pv * Ws[r * n + k] + (u64)f * Ws[piv * n + k], m);
                        }
                }
                if (carry_w)
                        for (int i = tid; i < n; i += SMALL_TB) {
                                const int r = (i == piv) ? j : i;
                                sd[i] = (i == j) ? pv : mp_mul(ss[r], pv, m);
                        }
                __syncthreads();
                s ^= 1;
        }
        *cur = s;
        return found;
}

// mode 0: full step; 1: reduce dots only; 2: semi_inverse of mats[VTAV]; 3: coefficients only
// Runs on ONE thread block of any size (all of its threads must call it); `sm` is a shared-memory
// scratch of small_smem_words(n) u32.
static __device__ __noinline__ void small_body(int n, int np, unsigned long long *sums, u32 *mats, DevSmall *state_,
                                        int mode, const ModP &m, u32 *sm)
{
        const int tid = threadIdx.x, SMALL_TB = blockDim.x;
        const int nn = n * n, npp = np * np;
        // (in the persistent loop kernel successive iterations may run this stage on different thread blocks: the few
        // flag words are read and written as volatile, i.e. at L2)
        volatile DevSmall *state = state_;
        u32 *A = sm;               // vtAv   (n x n)
        u32 *B = A + nn;           // vtAAv
        GjBuf buf[2];
        buf[0].M = B + nn;   buf[0].W = buf[0].M + nn;
        buf[1].M = buf[0].W + nn; buf[1].W = buf[1].M + nn;
        buf[0].scal = buf[1].W + nn; buf[1].scal = buf[0].scal + n;
        u32 *d = buf[1].scal + n;  // n
        u32 *d1 = d + n;           // n (phase-1 pivots)

        if (mode == 0 && state->halt) {
                if (tid == 0) state->do_ortho = 0;
                return;
        }

        // ---- gather the dot products (and clear the accumulators for the next iteration)
        if (mode <= 1) {
                for (int e = tid; e < 2 * npp; e += SMALL_TB) {
                        int which = e / npp, r = e - which * npp;
                        int i = r / np, j = r - i * np;
                        u32 val = mp_reduce(__ldcg(&sums[e]), m);        // (added by other blocks' atomics: read at L2)
                        sums[e] = 0;
                        if (i < n && j < n) (which ? B : A)[i * n + j] = val;
                        mats[(which ? MAT_VTAAV : MAT_VTAV) * npp + r] = val;
                }
                if (mode == 1) return;
        } else {
                for (int e = tid; e < nn; e += SMALL_TB) {
                        int i = e / n, j = e - i * n;
                        A[e] = mats[MAT_VTAV * npp + i * np + j];
                        B[e] = mats[MAT_VTAAV * npp + i * np + j];
                }
        }
        __syncthreads();

        int npiv = 0;
        u32 *W = buf[0].W;
        if (mode == 0 || mode == 2) {
                // ---- semi_inverse, phase 1: which columns carry a pivot
                int cur = 0;
                for (int e = tid; e < nn; e += SMALL_TB) buf[0].M[e] = A[e];
                __syncthreads();
                gj_sweep(buf, &cur, false, d1, n, m);
                __syncthreads();
                // ---- phase 2 on the d x d restriction, carrying winv along
                for (int e = tid; e < nn; e += SMALL_TB) {
                        int i = e / n, j = e - i * n;
                        bool keep = d1[i] && d1[j];
                        buf[cur].M[e] = keep ? A[e] : 0u;
                        buf[cur].W[e] = (i == j && d1[i]) ? 1u : 0u;
                }
                for (int i = tid; i < n; i += SMALL_TB) buf[cur].scal[i] = 1u;
                __syncthreads();
                npiv = gj_sweep(buf, &cur, true, d, n, m);
                __syncthreads();
                // undo the row scalings: n independent inverses, one per thread
                for (int i = tid; i < n; i += SMALL_TB) buf[cur].scal[i] = mp_inv(buf[cur].scal[i], m);
                __syncthreads();
                W = buf[cur].W;
                for (int e = tid; e < nn; e += SMALL_TB) W[e] = mp_mul(W[e], buf[cur].scal[e / n], m);
                __syncthreads();
                for (int e = tid; e < npp; e += SMALL_TB) {
                        int i = e / np, j = e - i * np;
                        mats[MAT_WINV * npp + e] = (i < n && j < n) ? W[i * n + j] : 0u;
                }
                for (int j = tid; j < np; j += SMALL_TB) mats[MAT_D * npp + j] = j < n ? d[j] : 0u;
                if (tid == 0) state->npiv = npiv;
                if (mode == 2) return;
        } else {
                for (int e = tid; e < nn; e += SMALL_TB) {
                        int i = e / n, j = e - i * n;
                        W[e] = mats[MAT_WINV * npp + i * np + j];
                }
                for (int j = tid; j < n; j += SMALL_TB) d[j] = mats[MAT_D * npp + j];
                __syncthreads();
                npiv = 1;
        }

        // ---- correctness_tests (sequential/lanczos_modp.c:532-557), on request (BLK_CHECK=1): vtAv, vtAAv, winv
        // symmetric; winv supported on the pivot rows/columns; winv * (vtAv D) == D.  The reference asserts them
        // on every iteration; here a violation halts the loop before anything is updated and blk_iterate fails.
        if (mode == 0 && state->check) {
                __shared__ int s_bad;
                if (tid == 0) {
                        s_bad = 0;
                        if (state->fault_iter > 0 && state->iters + 1 == state->fault_iter) A[n > 1 ? 1 : 0] ^= 1u;   // fault injection
                }
                __syncthreads();
                int bad = 0;
                for (int e = tid; e < nn; e += SMALL_TB) {
                        const int i = e / n, j = e - i * n;
                        if (A[e] != A[j * n + i]) bad |= 1;
                        if (B[e] != B[j * n + i]) bad |= 2;
                        if (W[e] != W[j * n + i]) bad |= 4;
                        if (W[e] != 0 && !d[i] && !d[j]) bad |= 8;
                        u64 s = 0;
                        if (d[j])
                                for (int k = 0; k < n; k++) {
                                        s += (u64)W[i * n + k] * A[k * n + j];
                                        mp_fold(s, m);
                                }
                        if (mp_reduce(s, m) != (i == j ? d[i] : 0u)) bad |= 16;
                }
                if (bad) atomicOr(&s_bad, bad);
                __syncthreads();
                if (s_bad) {
                        if (tid == 0) { state->check_failed = s_bad; state->halt = 1; state->do_ortho = 0; }
                        return;
                }
        }

        // ---- coefficients of orthogonalize (sequential/lanczos_modp.c:460-475)
        //   c     = -(winv * spliced), spliced[:,j] = d[j] ? vtAAv[:,j] : vtAv[:,j]
        //   vtAvd = d[j] ? -vtAv[:,j] : 0
        // (the reference stores p - x, which may equal p; canonical here, same value mod p)
        for (int e = tid; e < npp; e += SMALL_TB) {
                int i = e / np, j = e - i * np;
                u32 cval = 0, dval = 0;
                if (i < n && j < n) {
                        const u32 *S = d[j] ? B : A;
                        u64 s = 0;
                        for (int k = 0; k < n; k++) {
                                s += (u64)W[i * n + k] * S[k * n + j];
                                mp_fold(s, m);
                        }
                        cval = mp_neg(mp_reduce(s, m), m);
                        dval = d[j] ? mp_neg(A[i * n + j], m) : 0u;
                }
                mats[MAT_C * npp + e] = cval;
                mats[MAT_VTAVD * npp + e] = dval;
        }

        // ---- right-hand operands of the tensor-core orthogonalize, in MMA fragment order
        if (np >= 8 && np <= 32) {
                __syncthreads();                         // mats[C, VTAVD, WINV] written above by this block
                const int T = np / 8, WPT = np / 4;
                const int total = 12 * T * T * 64;
                u32 pw[4];
                for (int be = 0; be < 4; be++) pw[be] = mp_reduce(1ull << (8 * be), m);
                u32 *bfrag = mats + MAT_COUNT * npp;
                for (int idx = tid; idx < total; idx += SMALL_TB) {
                        int h = idx & 1, lane = (idx >> 1) & 31, rest = idx >> 6;
                        int s = rest % T; rest /= T;
                        int t = rest % T; rest /= T;
                        int b = rest & 3, X = rest >> 2;
                        int j = 8 * t + (lane >> 2), w = (lane & 3) * WPT + 2 * s + h;
                        const int which = X == 0 ? MAT_C : (X == 1 ? MAT_VTAVD : MAT_WINV);
                        u32 x = mats[which * npp + w * np + j];
                        u32 word = 0;
                        for (int be = 0; be < 4; be++)
                                word |= ((mp_mul(x, pw[be], m) >> (8 * b)) & 0xffu) << (8 * be);
                        bfrag[idx] = word;
                }
        }

        if (mode == 0 && tid == 0) {
                if (npiv == 0) {
                        state->stopped = 1; state->halt = 1; state->do_ortho = 0;
                } else {
                        int it = state->iters + 1;
                        state->iters = it;
                        state->do_ortho = 1;
                        if (state->limit > 0 && it >= state->limit) state->halt = 1;
                }
        }
}

