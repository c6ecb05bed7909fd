"""CPU suite, part 7: an executable model of the "interleaved chunk stream" (csrc/blk_internal.cuh) and of the
way k_spmv / k_spmv_fix (csrc/spmv.cu) walk it, in plain numpy, checked against the oracle.  The GPU tests prove
the kernels; this file is the specification they implement, small enough to read:

  * entries sorted by (row, col); an empty row owns one dummy entry (col 0, val 0); every entry carries LAST =
    "final entry of its row"; the stream is padded with (0, 0, not LAST) entries to a whole number of tiles;
  * chunk = Q consecutive entries, walked by one lane group; tile = G chunks = one warp; chunk_row = row of the
    chunk's first entry, plus HEAD_OPEN when that row started in an earlier chunk;
  * a group stores the rows that start AND end inside its chunk; the piece of a row it inherits (head) and the row
    it leaves unfinished (tail) are stitched inside the warp by a suffix scan over the groups; what crosses tile
    borders goes through whead[tile] and is finished either by k_spmv_fix using tail_row[tile] and span[tile]
    (the default: measured faster), or -- BLK_SPMV_FIX=lookback -- inside k_spmv by look-back: the tile in which the row ENDS
    (back[tile] = number of tiles back to the one where it started) waits for the `ready` flags of those tiles,
    adds the partial row the first one left in y, the whead of the ones in between and its own head, and clears
    the flags again;
  * column bands (SpOp::bands, small n_pad): the operator cut into K column ranges, each an ordinary chunk stream; either every
    band covers all rows and the K partial results are added mod p (k_band_combine), or every band is a compact stream over
    its non-empty rows (rowmap) whose results are added into a cleared y.
"""
import numpy as np
import pytest


class ChunkStream:
    def __init__(self, M, transpose, Q, G):
        rows, cols = (M.ncols, M.nrows) if transpose else (M.nrows, M.ncols)
        ri, ci = (M.j, M.i) if transpose else (M.i, M.j)
        order = np.lexsort((ci, ri))
        r, c, x = ri[order].astype(np.int64), ci[order].astype(np.int64), M.x[order].astype(np.int64)
        cnt = np.bincount(r, minlength=rows)
        er, ec, ex, last = [], [], [], []
        pos = 0
        for row in range(rows):
            k = int(cnt[row])
            if k == 0:                                   # dummy entry: the row still gets written (as zero)
                er.append(row); ec.append(0); ex.append(0); last.append(True)
            for t in range(k):
                er.append(row); ec.append(int(c[pos + t])); ex.append(int(x[pos + t])); last.append(t == k - 1)
            pos += k
        stored = len(er)
        tile_len = Q * G
        ntiles = max(1, -(-stored // tile_len))
        pad = ntiles * tile_len - stored
        er += [rows] * pad; ec += [0] * pad; ex += [0] * pad; last += [False] * pad
        self.rows, self.cols, self.Q, self.G, self.ntiles, self.stored = rows, cols, Q, G, ntiles, stored
        self.er, self.ec, self.ex, self.last = map(np.array, (er, ec, ex, last))
        nch = ntiles * G
        self.chunk_row = np.zeros(nch, np.int64)
        self.head_open = np.zeros(nch, bool)
        for ch in range(nch):
            s = ch * Q
            self.chunk_row[ch] = self.er[s]
            # the first entry continues a row begun earlier iff the previous stored entry is not LAST
            self.head_open[ch] = s > 0 and s < stored and not self.last[s - 1]
        # per tile: the row left open at its end and how many following tiles it takes to finish it
        self.tail_row = np.zeros(ntiles, np.int64)
        self.span = np.zeros(ntiles, np.int64)
        for t in range(ntiles):
            e = (t + 1) * tile_len - 1                  # last entry of the tile
            if e < stored - 1 and not self.last[e] and self._starts_in_tile(e, t):
                row = self.er[e]
                end = e + 1
                while not self.last[end]:
                    end += 1
                self.tail_row[t] = row
                self.span[t] = end // tile_len - t

    def _starts_in_tile(self, e, t):
        """the row of entry e begins inside tile t (a row that began earlier is carried by an earlier tile's span)"""
        s = e
        while s > 0 and not self.last[s - 1]:
            s -= 1
        return s >= t * self.Q * self.G

    def interleaved(self):
        """the storage order of the device array: entry i of chunk g of tile t at t*G*Q + i*G + g"""
        Q, G = self.Q, self.G
        idx = np.arange(self.ntiles * G * Q).reshape(self.ntiles, G, Q)          # [tile][chunk][i] -> stream position
        return idx.transpose(0, 2, 1).reshape(-1)                                # position in memory -> stream position

    def spmv(self, xblk, n, p, lookback=False):
        """y <- S x exactly as the kernels do it (python integers: no overflow questions here)"""
        Q, G = self.Q, self.G
        back = np.zeros(self.ntiles, np.int64)                   # layout_build.cu k_tile_tails: back[t + span] = span
        for t in range(self.ntiles):
            if self.span[t]:
                assert back[t + self.span[t]] == 0               # a tile finishes at most one row
                back[t + self.span[t]] = self.span[t]
        ready = np.zeros(self.ntiles, bool)
        X = xblk.reshape(self.cols, n).astype(object)
        y = np.zeros((self.rows + 1, n), dtype=object)           # row `rows` = the padding's sink, never read
        whead = np.zeros((self.ntiles, n), dtype=object)
        pending_last = np.zeros(self.ntiles, bool)               # k_spmv: pending && (head_type == 2 || has_tail) of the last group
        for t in range(self.ntiles):                              # ---- k_spmv: one warp per tile
            headv = np.zeros((G, n), dtype=object)
            head_type = np.zeros(G, int)
            tailv = np.zeros((G, n), dtype=object)
            has_tail = np.zeros(G, bool)
            tail_row = np.zeros(G, np.int64)
            for g in range(G):                                    # one lane group per chunk
                ch = t * G + g
                row, open_ = int(self.chunk_row[ch]), bool(self.head_open[ch])
                acc, pending = np.zeros(n, dtype=object), False
                for i in range(Q):
                    s = ch * Q + i
                    acc = acc + int(self.ex[s]) * X[self.ec[s]]
                    pending = True
                    if self.last[s]:
                        r = acc % p
                        if open_:
                            headv[g], head_type[g], open_ = r, 1, False
                        else:
                            y[row] = r
                        row += 1
                        acc, pending = np.zeros(n, dtype=object), False
                if pending:
                    if open_:
                        headv[g], head_type[g] = acc % p, 2       # the whole chunk lies inside one row
                    elif row < self.rows:
                        tailv[g], has_tail[g], tail_row[g] = acc % p, True, row
                    if g == G - 1:
                        pending_last[t] = head_type[g] == 2 or has_tail[g]
            # suffix scan: S_g = heads of chunks g, g+1, ... up to the chunk in which the row ends
            S = headv.copy()
            closed = head_type != 2
            for g in range(G - 2, -1, -1):
                if not closed[g]:
                    S[g] = (S[g] + S[g + 1]) % p
                    closed[g] = closed[g + 1]
            for g in range(G):
                if has_tail[g]:
                    y[tail_row[g]] = (tailv[g] + (S[g + 1] if g < G - 1 else 0)) % p
            if head_type[0] != 0:
                whead[t] = S[0]
            if lookback:
                # (1) the row open at the end of the tile: its part is in memory now
                if pending_last[t] :
                    ready[t] = True
                # (2) the row open at the start ends here and began back[t] tiles earlier (lower tiles: already done)
                bk = int(back[t])
                if bk:
                    assert self.head_open[t * G] and closed[0] and all(ready[t - j] for j in range(1, bk + 1))
                    hrow = int(self.chunk_row[t * G])
                    y[hrow] = (y[hrow] + sum(whead[t - j] for j in range(1, bk)) + S[0]) % p
                    for j in range(1, bk + 1):
                        ready[t - j] = False
        if lookback:
            assert not ready.any()                                # every flag was consumed and cleared: re-armed for the next launch
            return np.array(y[:self.rows].tolist(), dtype=np.uint32).ravel()
        for t in range(self.ntiles):                              # ---- k_spmv_fix: rows crossing tile borders
            if self.span[t]:
                y[self.tail_row[t]] = (y[self.tail_row[t]] + sum(whead[t + 1 + j] for j in range(int(self.span[t])))) % p
        return np.array(y[:self.rows].tolist(), dtype=np.uint32).ravel()


def matrices(B):
    s = B.synth
    giant = s.powerlaw_rows(60, 300, mean=4, seed=4)
    gi = np.concatenate([giant.i, np.full(700, 17, np.int32), np.arange(60, dtype=np.int32)])
    gj = np.concatenate([giant.j, np.random.default_rng(0).integers(0, 300, 700).astype(np.int32), np.full(60, 5, np.int32)])
    gx = np.concatenate([giant.x, np.arange(700, dtype=np.uint32) + 1, np.arange(60, dtype=np.uint32) + 7])
    return {
        "uniform": s.uniform_rows(90, 70, 5, seed=1),
        "powerlaw_empty": s.powerlaw_rows(150, 120, mean=6, seed=2, with_empty_rows=25, order="file"),
        "giant_row_and_column": B.SparseCOO(60, 300, gi, gj, gx),       # rows spanning many chunks and many tiles
        "single": B.SparseCOO(1, 1, np.zeros(1, np.int32), np.zeros(1, np.int32), np.array([3], np.uint32)),
        "empty": B.SparseCOO(5, 4, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint32)),
    }


@pytest.mark.parametrize("Q,G", [(8, 8), (8, 2), (16, 4), (64, 32)])
def test_chunk_stream_model_matches_oracle(lib, oracle, Q, G):
    p, n = 2147483647, 3
    rng = np.random.default_rng(Q * 100 + G)
    for name, M in matrices(lib).items():
        Mp = M.reduced(p)
        for transpose in (False, True):
            cs = ChunkStream(Mp, transpose, Q, G)
            cols = M.nrows if transpose else M.ncols
            x = rng.integers(0, p, size=cols * n).astype(np.uint32)
            x[::3] = p - 1
            want = oracle.sparse_matrix_vector_product(Mp, x, transpose, n, p)
            assert np.array_equal(cs.spmv(x, n, p), want), (name, transpose)
            assert np.array_equal(cs.spmv(x, n, p, lookback=True), want), (name, transpose, "look-back")
            # layout invariants: every row owns >= 1 stored entry, exactly one LAST per row, padding after the data
            assert cs.stored == Mp.nnz + int((np.bincount((Mp.j if transpose else Mp.i), minlength=cs.rows) == 0).sum())
            assert int(cs.last.sum()) == cs.rows and not cs.last[cs.stored:].any()
            perm = cs.interleaved()
            assert sorted(perm.tolist()) == list(range(cs.ntiles * Q * G))
            # one step of a warp (fixed i) reads G consecutive memory positions: entry i of each of its chunks
            t, i = cs.ntiles - 1, Q // 2
            mem = np.empty_like(perm); mem[perm] = np.arange(perm.size)                   # stream position -> memory position
            where = [mem[(t * G + g) * Q + i] for g in range(G)]
            assert where == list(range(where[0], where[0] + G))


@pytest.mark.parametrize("K", [2, 3, 7])
def test_column_band_model_matches_oracle(lib, oracle, K):
    """csrc/context.cu build_bands + csrc/spmv.cu launch_spmv: both forms of the banded product."""
    p, n, Q, G = 2147483647, 2, 8, 4
    rng = np.random.default_rng(K)
    for name, M in matrices(lib).items():
        Mp = M.reduced(p)
        for transpose in (False, True):
            rows, cols = (M.ncols, M.nrows) if transpose else (M.nrows, M.ncols)
            ri, ci = (Mp.j, Mp.i) if transpose else (Mp.i, Mp.j)
            x = rng.integers(0, p, size=cols * n).astype(np.uint32)
            want = oracle.sparse_matrix_vector_product(Mp, x, transpose, n, p)
            band_col = [cols * b // K for b in range(K + 1)]
            partial = np.zeros((K, rows * n), dtype=np.uint64)
            acc = np.zeros((rows, n), dtype=np.uint64)                           # k_zero_block
            seen = 0
            for b in range(K):
                sel = (ci >= band_col[b]) & (ci < band_col[b + 1])
                seen += int(sel.sum())
                # form (a): the band over all rows (rows without entries in it get a dummy entry and produce zero)
                band = lib.SparseCOO(rows, cols, ri[sel], ci[sel], Mp.x[sel])
                partial[b] = ChunkStream(band, False, Q, G).spmv(x, n, p)
                # form (b): compact rows -- compact_rows() renumbers the rows that occur, rowmap maps them back
                rowmap = np.unique(ri[sel])
                if rowmap.size:
                    renum = np.searchsorted(rowmap, ri[sel]).astype(np.int32)
                    compact = lib.SparseCOO(int(rowmap.size), cols, renum, ci[sel], Mp.x[sel])
                    cs = ChunkStream(compact, False, Q, G)
                    assert cs.stored == int(sel.sum())                           # no dummy entries in a compact band
                    part = cs.spmv(x, n, p).reshape(rowmap.size, n).astype(np.uint64)
                    acc[rowmap] = (acc[rowmap] + part) % p                       # store_row<ACC>: y[rowmap[r]] += result
            assert seen == Mp.nnz                                                # the bands partition the entries
            assert np.array_equal((partial.sum(axis=0) % p).astype(np.uint32), want), (name, transpose, "partial + combine")
            assert np.array_equal(acc.astype(np.uint32).ravel(), want), (name, transpose, "accumulate")
