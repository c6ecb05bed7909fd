// spmv_body.cuh -- device code of the sparse product on the interleaved chunk stream, shared by the
// stand-alone kernels (spmv.cu: k_spmv, k_spmv_fix) and by the persistent loop kernel (loop_coop.cu).
//
// COH = 1 (persistent kernel): the vectors are rewritten by other blocks of the SAME launch between
// two grid barriers, so x rows are gathered with ld.global.cg (L2, the coherence point) instead of the
// non-coherent read-only path; the matrix stream itself never changes and keeps __ldg.
#pragma once
#include "blk_internal.cuh"

constexpr int SPMV_U = 8;         // unroll: gathers in flight per lane

template <int V> struct Vec;
template <> struct Vec<1> { typedef unsigned int T; };
template <> struct Vec<2> { typedef uint2 T; };
template <> struct Vec<4> { typedef uint4 T; };

// Gather of one x row segment.  HOT = 1: x rows below `hot` (the high-degree prefix of a
// degree-sorted dimension) are loaded with an L2 evict_last policy, everything else with
// evict_first, so that the part of x that is gathered over and over stays resident while the
// once-only traffic streams through (tools/hot_gather.cu).
template <int V, int HOT> __device__ __forceinline__ void load_vec(u32 (&o)[V], const u32 *p, u64 pol)
{
        if (V == 4 && HOT) {
                u32 a, b, c, d;
                asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                             : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p), "l"(pol));
                o[0] = a; o[1 % V] = b; o[2 % V] = c; o[3 % V] = d;
        } else {
                typename Vec<V>::T t = __ldg(reinterpret_cast<const typename Vec<V>::T *>(p));
                const u32 *w = reinterpret_cast<const u32 *>(&t);
#pragma unroll
                for (int k = 0; k < V; k++) o[k] = w[k];
        }
}
template <int V> __device__ __forceinline__ void load_vec_rw(u32 (&o)[V], const u32 *p)
{
        typename Vec<V>::T t = *reinterpret_cast<const typename Vec<V>::T *>(p);
        const u32 *w = reinterpret_cast<const u32 *>(&t);
#pragma unroll
        for (int k = 0; k < V; k++) o[k] = w[k];
}
template <int V> __device__ __forceinline__ void store_vec(u32 *p, const u32 (&o)[V])
{
        typename Vec<V>::T t;
        u32 *w = reinterpret_cast<u32 *>(&t);
#pragma unroll
        for (int k = 0; k < V; k++) w[k] = o[k];
        *reinterpret_cast<typename Vec<V>::T *>(p) = t;
}

template <int V> __device__ __forceinline__ void load_vec_cg(u32 (&o)[V], const u32 *p)
{
        typename Vec<V>::T t = __ldcg(reinterpret_cast<const typename Vec<V>::T *>(p));
        const u32 *w = reinterpret_cast<const u32 *>(&t);
#pragma unroll
        for (int k = 0; k < V; k++) o[k] = w[k];
}

__device__ __forceinline__ u32 ld_acquire(const u32 *p)
{
        u32 v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
}

// Store of a finished (or partial) output row.  ACC = 1 (compact column-band operators, SpOp::rowmap): output row `row` of the
// operator is row rowmap[row] of y and the result is ADDED to what y holds -- no other warp of the launch touches that row.
template <int V, int ACC>
__device__ __forceinline__ void store_row(u32 *ys, const u32 *__restrict__ rowmap, const u32 row, const int np, const u32 (&r)[V], const ModP &m)
{
        if (ACC) {
                u32 *dst = ys + (size_t)__ldg(rowmap + row) * np;
                u32 old[V], o[V];
                load_vec_rw<V>(old, dst);
#pragma unroll
                for (int k = 0; k < V; k++) o[k] = mp_add(old[k], r[k], m);
                store_vec<V>(dst, o);
        } else {
                store_vec<V>(ys + (size_t)row * np, r);
        }
}

// One tile (one warp): lane-group g walks chunk g of tile t.  back / ready (nullable together): complete the
// rows that cross tile borders by look-back (see SpOp).
template <int L, int V, int FOLD, int HOT, int PUSH, int COH, int ACC = 0>
__device__ __forceinline__ void spmv_tile(const uint2 *__restrict__ ent, const u32 *__restrict__ chunk_row, u32 *whead,
                                          const int64_t t, const int Q, const u32 rows, const u32 *x, u32 *y, const ModP &m,
                                          const u64 pol_hot, const u64 pol_cold, const PushTargets &push,
                                          const u32 *__restrict__ back, u32 *ready, const int lane,
                                          const u32 *__restrict__ rowmap = nullptr)
{
        constexpr int G = 32 / L;
        constexpr int NP = L * V;
        constexpr int U = SPMV_U;
        const int g = lane / L, sub = lane % L;

        const u32 cr = __ldg(chunk_row + t * G + g);
        u32 row = cr & 0x7fffffffu;
        bool head_open = (cr >> 31) != 0;
        const u32 first_started = row + (head_open ? 1u : 0u);      // (group 0's value is the tile's)
        const uint2 *e = ent + t * G * Q + g;
        const u32 *xs = x + sub * V;
        u32 *ys = y + sub * V;

        u64 acc[V];
        u32 headv[V];
#pragma unroll
        for (int k = 0; k < V; k++) { acc[k] = 0; headv[k] = 0; }
        int head_type = 0;              // 0 none, 1 row ended inside the chunk, 2 whole chunk inside one row
        bool pending = false;

        for (int i0 = 0; i0 < Q; i0 += U) {
                uint2 ee[U];
#pragma unroll
                for (int u = 0; u < U; u++) ee[u] = __ldg(e + (size_t)(i0 + u) * G);
                u32 xv[U][V];
#pragma unroll
                for (int u = 0; u < U; u++)
                {
                        // (HOT: bit 30 of the column word, set at layout-build time, marks the L2-resident x rows)
                        const u32 col = ee[u].x & (HOT ? 0x3fffffffu : 0x7fffffffu);
                        if (COH) load_vec_cg<V>(xv[u], xs + (size_t)col * NP);
                        else load_vec<V, HOT>(xv[u], xs + (size_t)col * NP, (ee[u].x & 0x40000000u) ? pol_hot : pol_cold);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
#pragma unroll
                        for (int k = 0; k < V; k++) mp_mac(acc[k], ee[u].y, xv[u][k]);
                        if (FOLD != 0 && (u % (FOLD ? FOLD : 1)) == (FOLD ? FOLD : 1) - 1) {
#pragma unroll
                                for (int k = 0; k < V; k++) mp_fold(acc[k], m);
                        }
                        pending = true;
                        if (ee[u].x & 0x80000000u) {            // last entry of its row
                                u32 r[V];
#pragma unroll
                                for (int k = 0; k < V; k++) { r[k] = mp_reduce(acc[k], m); acc[k] = 0; }
                                if (head_open) {
#pragma unroll
                                        for (int k = 0; k < V; k++) headv[k] = r[k];
                                        head_type = 1;
                                        head_open = false;
                                } else {
                                        store_row<V, ACC>(ys, rowmap, row, NP, r, m);
                                }
                                row++;
                                pending = false;
                        }
                }
        }

        u32 tailv[V];
#pragma unroll
        for (int k = 0; k < V; k++) tailv[k] = 0;
        bool has_tail = false;
        if (pending) {
                if (head_open) {
#pragma unroll
                        for (int k = 0; k < V; k++) headv[k] = mp_reduce(acc[k], m);
                        head_type = 2;
                } else if (row < rows) {            // row == rows: trailing padding only
#pragma unroll
                        for (int k = 0; k < V; k++) tailv[k] = mp_reduce(acc[k], m);
                        has_tail = true;
                }
        }

        // stitch rows that cross chunk borders inside the warp: suffix scan over the groups.
        // S = sum of chunk heads from this chunk up to (and including) the chunk where the
        // row ends; closed = that chunk lies inside the warp.
        int closed = head_type != 2;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
                u32 sp[V];
#pragma unroll
                for (int k = 0; k < V; k++) sp[k] = __shfl_down_sync(0xffffffffu, headv[k], d * L);
                int cp = __shfl_down_sync(0xffffffffu, closed, d * L);
                if (!closed && g + d < G) {
#pragma unroll
                        for (int k = 0; k < V; k++) headv[k] = mp_add(headv[k], sp[k], m);
                        closed = cp;
                }
        }
        u32 nx[V];
#pragma unroll
        for (int k = 0; k < V; k++) nx[k] = __shfl_down_sync(0xffffffffu, headv[k], L);
        if (has_tail) {
                // complete if the row ends inside this warp, else a partial finished by k_spmv_fix
                u32 o[V];
#pragma unroll
                for (int k = 0; k < V; k++) o[k] = (g < G - 1) ? mp_add(tailv[k], nx[k], m) : tailv[k];
                store_row<V, ACC>(ys, rowmap, row, NP, o, m);
        }
        if (g == 0 && head_type != 0) store_vec<V>(whead + t * NP + sub * V, headv);

        if (back) {
                // ---- rows that cross tile borders, finished here instead of by a second kernel.
                // (1) A row still open at the END of this tile needs this tile's part later: it is in memory now
                //     (the partial row in y if the row started here, whead[t] if the whole tile lies inside it).
                const int open_end = __shfl_sync(0xffffffffu, (int)(pending && (head_type == 2 || has_tail)), 31);
                __syncwarp();
                if (open_end && lane == 0) {
                        __threadfence();
                        *(volatile u32 *)(ready + t) = 1u;
                }
                // (2) The row open at the START of this tile ends in it and started bk tiles back: wait for those
                //     tiles (they have lower indices, so they are running or done), add the partial of the first, the
                //     whead of the ones in between and this tile's own head, and store the finished row.
                const u32 bk = __ldg(back + t);
                if (bk) {
                        for (u32 j = 1 + lane; j <= bk; j += 32)
                                while (ld_acquire(ready + (t - j)) == 0) { }
                        __syncwarp();
                        const u32 hrow = __shfl_sync(0xffffffffu, cr & 0x7fffffffu, 0);      // group 0 started on that row
                        u64 sum[V];
#pragma unroll
                        for (int k = 0; k < V; k++) sum[k] = 0;
                        for (u32 j = 1 + g; j < bk; j += G) {
                                u32 h[V];
                                load_vec_cg<V>(h, whead + (size_t)(t - j) * NP + sub * V);
#pragma unroll
                                for (int k = 0; k < V; k++) sum[k] += h[k];
                        }
                        if (g == 0) {
                                u32 h[V];
                                load_vec_cg<V>(h, ys + (size_t)hrow * NP);
#pragma unroll
                                for (int k = 0; k < V; k++) sum[k] += (u64)h[k] + headv[k];
                        }
#pragma unroll
                        for (int d = L; d < 32; d <<= 1) {
#pragma unroll
                                for (int k = 0; k < V; k++) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], d);
                        }
                        if (g == 0) {
                                u32 o[V];
#pragma unroll
                                for (int k = 0; k < V; k++) o[k] = mp_reduce(sum[k], m);
                                store_vec<V>(ys + (size_t)hrow * NP, o);
                                if (PUSH) {
#pragma unroll
                                        for (int q = 0; q < PushTargets::MAX; q++)
                                                if (q < push.n) store_vec<V>(push.y[q] + (size_t)hrow * NP + sub * V, o);
                                }
                        }
                        for (u32 j = 1 + lane; j <= bk; j += 32) ready[t - j] = 0u;          // consumed: re-arm for the next launch
                }
        }

        if (PUSH) {
                // Rows this warp has finalised: every row that STARTS in the tile, except the one still open at
                // its end (that one is finished -- and pushed -- by k_spmv_fix).  They are consecutive:
                // [first row started by group 0, the row the last group stands on).
                const u32 ra = __shfl_sync(0xffffffffu, first_started, 0);
                u32 rb = __shfl_sync(0xffffffffu, row, 31);
                if (rb > rows) rb = rows;
                __syncwarp();                                    // the warp's own stores to y, made by other lanes
                for (u32 r = ra + g; r < rb; r += G) {
                        u32 o[V];
                        load_vec_cg<V>(o, ys + (size_t)r * NP);
#pragma unroll
                        for (int q = 0; q < PushTargets::MAX; q++)
                                if (q < push.n) store_vec<V>(push.y[q] + (size_t)r * NP + sub * V, o);
                }
        }
}

// One row that crosses tile borders: y[row] (partial left by tile `gid`, where the row starts) plus the heads of the
// `span` following tiles.  Finishes the rows that END in tiles [tile_lo, tile_hi).  COH as above.
template <int L, int V, int COH>
__device__ __forceinline__ void spmv_fix_row(const u32 *__restrict__ tail_row, const u32 *__restrict__ span, const u32 *whead,
                                             const int64_t gid, const int sub, const int64_t tile_lo, const int64_t tile_hi,
                                             u32 *y, const ModP &m, const PushTargets &push, const u32 *__restrict__ rowmap = nullptr)
{
        constexpr int NP = L * V;
        u32 sp = __ldg(span + gid);
        if (sp == 0) return;
        if (gid + sp >= tile_hi || gid + sp < tile_lo) return;      // ends in a later / an earlier piece
        u32 r = __ldg(tail_row + gid);
        if (rowmap) r = __ldg(rowmap + r);          // compact band operator: the partial row lives (accumulated) in y[rowmap[r]]
        u32 cur[V];
        if (COH) load_vec_cg<V>(cur, y + (size_t)r * NP + sub * V);
        else load_vec_rw<V>(cur, y + (size_t)r * NP + sub * V);
        u64 s[V];
#pragma unroll
        for (int k = 0; k < V; k++) s[k] = cur[k];
        const u32 *w = whead + (size_t)(gid + 1) * NP + sub * V;
#pragma unroll 4
        for (u32 j = 0; j < sp; j++) {
                u32 h[V];
                if (COH) load_vec_cg<V>(h, w + (size_t)j * NP);
                else load_vec_rw<V>(h, w + (size_t)j * NP);
#pragma unroll
                for (int k = 0; k < V; k++) s[k] += h[k];
        }
#pragma unroll
        for (int k = 0; k < V; k++) cur[k] = mp_reduce(s[k], m);
        store_vec<V>(y + (size_t)r * NP + sub * V, cur);
        for (int q = 0; q < push.n; q++) store_vec<V>(push.y[q] + (size_t)r * NP + sub * V, cur);
}
