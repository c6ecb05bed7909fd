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
