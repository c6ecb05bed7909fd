// modp.cuh -- GF(p) arithmetic for word-size primes on sm_100a.
//
// The reference does one 64-bit `%` per multiply-accumulate
// (sequential/lanczos_modp.c:284,300,313).  All of its state is canonical
// (every store is `% prime`), so any exact evaluation mod p is bit-identical
// (SURVEY.md F8).  Here products are accumulated lazily in u64:
//
//   mac     acc += a*b                       one IMAD.WIDE.U32
//   fold    acc  = lo32(acc) + hi32(acc)*c   c = 2^32 mod p; keeps acc < (2^32-1)*p + 2^32
//   reduce  Barrett with mu = floor(2^64/p)  -> canonical u32
//
// `fold_every` (0, 2 or 8) says how many products may be added between two
// folds without overflowing u64; 0 means 64 or more (kernels never chain more
// than 64 products without folding).  It is derived from p on the host
// (modp_make) and selects a kernel template instance.
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;

struct ModP {
        u32 p;        // modulus, 2 <= p < 2^31
        u32 c32;      // 2^32 mod p
        u64 mu;       // floor(2^64 / p)
        int fold_every;
        // short Barrett after one fold (see mp_reduce): valid when fast != 0
        u32 fast;     // 1: (fold(x) >> s1) and mu2 fit 32 bits
        u32 s1, t2;   // shifts
        u32 mu2;      // floor(2^(s1+t2) / p)
};

// host: derive the constants; returns false if p is out of range
static inline bool modp_make(ModP *m, u64 p)
{
        if (p < 2 || p >= (1ull << 31)) return false;
        m->p = (u32)p;
        m->c32 = (u32)((1ull << 32) % p);
        // floor(2^64 / p) without 128-bit types: 2^64 = q*p + r
        u64 q = (~0ull) / p, r = (~0ull) % p;     // 2^64 - 1 = q*p + r
        if (r + 1 == p) q += 1;                   // 2^64 = (q+1)*p
        m->mu = q;
        // after a fold acc <= (2^32-1) + (2^32-1)*c32; each product <= (p-1)^2
        long double after_fold = 4294967295.0L + 4294967295.0L * (long double)m->c32;
        long double room = 18446744073709551615.0L - after_fold;
        long double prod = (long double)(p - 1) * (long double)(p - 1);
        long double k = prod > 0 ? room / prod : 1e30L;
        if (k >= 64.0L) m->fold_every = 0;
        else if (k >= 8.0L) m->fold_every = 8;
        else if (k >= 2.0L) m->fold_every = 2;
        else return false;
        // Short reduction: x1 = fold(x) <= B1 = (2^32-1)(1+c32) < 2^XB.  With k = bits(p), s1 = k-3,
        // t2 = XB+2-s1, mu2 = floor(2^(XB+2)/p):  qh = ((x1 >> s1) * mu2) >> t2 satisfies
        // floor(x1/p) - 1 <= qh <= floor(x1/p)  (truncation errors 2^s1/p <= 1/4 and x1/2^(XB+2) < 1/4),
        // so x1 - qh*p lies in [0, 2p) and can be formed in 32-bit arithmetic (2p <= 2^32).
        // Usable when both factors fit 32 bits, i.e. XB - s1 <= 32: primes 2^k +- small, e.g.
        // 2^31-1, 2^30-35, 65537.
        m->fast = 0; m->s1 = m->t2 = 0; m->mu2 = 0;
        {
                long double b1 = 4294967295.0L * (1.0L + (long double)m->c32);
                int xb = 1;
                while (xb < 64 && b1 >= (long double)(1ull << xb) ) xb++;
                if (b1 >= 18446744073709551616.0L) xb = 65;
                int k = 0;
                while ((p >> k) != 0) k++;
                int s1 = k > 3 ? k - 3 : 0;
                if (xb <= 64 && xb - s1 <= 32) {
                        int e = xb + 2;
                        u64 q = 0, r = 1;
                        bool ok = true;
                        for (int i = 0; i < e; i++) {          // q = floor(2^e / p) by long division
                                r <<= 1; q <<= 1;
                                if (r >= p) { r -= p; q += 1; }
                                if (q >> 32) { ok = false; break; }
                        }
                        if (ok) { m->fast = 1; m->s1 = (u32)s1; m->t2 = (u32)(e - s1); m->mu2 = (u32)q; }
                }
        }
        return true;
}

#ifdef __CUDACC__
__device__ __forceinline__ void mp_mac(u64 &acc, u32 a, u32 b)
{
        acc += (u64)a * (u64)b;
}

__device__ __forceinline__ void mp_fold(u64 &acc, const ModP &m)
{
        acc = (u64)(u32)acc + (u64)(u32)(acc >> 32) * (u64)m.c32;
}

// canonical residue of any u64
__device__ __forceinline__ u32 mp_reduce(u64 acc, const ModP &m)
{
        if (m.fast) {           // uniform branch: one fold, 32-bit Barrett, one conditional subtract
                u64 x1 = (u64)(u32)acc + (u64)(u32)(acc >> 32) * (u64)m.c32;
                u32 y = (u32)(x1 >> m.s1);
                u32 qh = (u32)(((u64)y * (u64)m.mu2) >> m.t2);
                u32 r = (u32)x1 - qh * m.p;
                u32 r2 = r - m.p;
                return r2 < r ? r2 : r;      // r < p: r - p wraps around
        }
        u64 q = __umul64hi(acc, m.mu);
        u64 r = acc - q * (u64)m.p;
        if (r >= m.p) r -= m.p;
        if (r >= m.p) r -= m.p;
        return (u32)r;
}

__device__ __forceinline__ u32 mp_add(u32 a, u32 b, const ModP &m)
{
        u32 s = a + b;                 // < 2^32 because p < 2^31
        return s >= m.p ? s - m.p : s;
}

__device__ __forceinline__ u32 mp_neg(u32 a, const ModP &m)
{
        return a ? m.p - a : 0u;
}

__device__ __forceinline__ u32 mp_mul(u32 a, u32 b, const ModP &m)
{
        return mp_reduce((u64)a * (u64)b, m);
}

// a^-1 mod p, extended Euclid (same value as invmod, sequential/lanczos_modp.c:318-336:
// the canonical inverse is unique).  a must be non-zero mod p.
__device__ __forceinline__ u32 mp_inv(u32 a, const ModP &m)
{
        long long t0 = 0, t1 = 1;
        u32 r0 = m.p, r1 = a % m.p;
        while (r1) {
                u32 q = r0 / r1;
                u32 r2 = r0 - q * r1; r0 = r1; r1 = r2;
                long long t2 = t0 - (long long)q * t1; t0 = t1; t1 = t2;
        }
        if (t0 < 0) t0 += m.p;
        return (u32)t0;
}
#endif
