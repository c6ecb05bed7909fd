"""CPU suite, part 4: the GF(p) arithmetic header of the CUDA kernels (csrc/modp.cuh) is plain
C++ on the host side; compile its constant derivation with g++ and check the reductions the
kernels rely on (fold, 64-bit Barrett, the short 32-bit Barrett) against `%` on many primes and
on worst-case inputs."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#include "modp.cuh"
// host replicas of the device functions (same expressions as in modp.cuh)
static u64 umul64hi(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }
static u32 reduce_long(u64 acc, const ModP &m) { u64 q = umul64hi(acc, m.mu); u64 r = acc - q * (u64)m.p; if (r >= m.p) r -= m.p; if (r >= m.p) r -= m.p; return (u32)r; }
static u32 reduce_short(u64 acc, const ModP &m) {
    u64 x1 = (u64)(u32)acc + (u64)(u32)(acc >> 32) * (u64)m.c32; u32 y = (u32)(x1 >> m.s1);
    u32 qh = (u32)(((u64)y * (u64)m.mu2) >> m.t2); u32 r = (u32)x1 - qh * m.p; u32 r2 = r - m.p; return r2 < r ? r2 : r; }
static u64 rnd() { return ((u64)rand() << 42) ^ ((u64)rand() << 21) ^ (u64)rand(); }
int main() {
    u64 primes[] = {2, 3, 7, 251, 65521, 65537, 1048583, 4294967ull, 16777259ull, 268435459ull, 536870923ull, 1073741789ull,
                    1073741827ull, 1610612741ull, 2147483629ull, 2147483647ull};
    long bad = 0; int nfast = 0;
    for (u64 p : primes) {
        ModP m; if (!modp_make(&m, p)) { printf("rejected %llu\n", (unsigned long long)p); return 2; }
        nfast += m.fast;
        // fold keeps the value congruent and the chain bound holds
        int chain = m.fold_every ? m.fold_every : 64;
        for (int t = 0; t < 200000; t++) {
            u64 acc = 0; unsigned __int128 exact = 0;
            for (int k = 0; k < 3; k++) {                       // three fold periods
                for (int i = 0; i < chain; i++) {
                    u64 a = (t % 3 == 0) ? p - 1 : rnd() % p, b = (t % 5 == 0) ? p - 1 : rnd() % p;
                    unsigned __int128 wide = (unsigned __int128)acc + (unsigned __int128)a * b;
                    if (wide >> 64) { bad++; printf("overflow p=%llu\n", (unsigned long long)p); break; }
                    acc = (u64)wide; exact += (unsigned __int128)a * b;
                }
                acc = (u64)(u32)acc + (u64)(u32)(acc >> 32) * (u64)m.c32;     // mp_fold
            }
            if (reduce_long(acc, m) != (u32)(exact % p)) bad++;
            if (m.fast && reduce_short(acc, m) != (u32)(exact % p)) bad++;
        }
        for (int t = 0; t < 2000000; t++) {
            u64 x = rnd(); if (t % 7 == 0) x = ~0ull - (u64)(t % 4096); if (t % 11 == 0) x = (u64)t;
            if (reduce_long(x, m) != (u32)(x % p)) bad++;
            if (m.fast && reduce_short(x, m) != (u32)(x % p)) bad++;
        }
    }
    ModP m; if (modp_make(&m, 1ull << 31) || modp_make(&m, 1) || modp_make(&m, 4294967311ull)) { printf("range check\n"); return 3; }
    printf("bad=%ld fast=%d\n", bad, nfast);
    return bad ? 1 : 0;
}
'''


def test_modp_header_reductions(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    inc = os.path.join(ROOT, "block-lanczos-algorithm-parallelization_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-I", inc, str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad=0" in r.stdout
    assert int(r.stdout.split("fast=")[1]) >= 6          # 2^31-1, 2^30-35, 65537, ... take the short path
