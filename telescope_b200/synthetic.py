# -*- coding: utf-8 -*-
"""Deterministic synthetic read x locus score matrices in the regime of real Telescope runs.

The matrix is defined chunk by chunk (CHUNK_ROWS reads per chunk, each chunk seeded by (seed, chunk index)), so
any contiguous row range -- in particular one GPU rank's shard -- can be generated without generating the rest,
and the same (N, K, avg, skew, seed) always yields the same bytes whatever the sharding.

Shape of the data (mirrors what `Telescope._mapping_to_matrix`, reference model.py:287-362, produces):
  * canonical CSR: column indices strictly increasing within a row, int32; scores uint16 (> 0)
  * ~20 % of reads are unique (one alignment, Y=0: exercises pisum0 / `unique`), the rest are multi-mapped with
    Poisson row lengths (or a Zipf tail truncated at 200 when skew=True), overall mean ~ `avg`
  * a read's loci are a run of near neighbours around a "home" locus drawn from a log-normal abundance over the
    K loci (TE subfamilies), so EM has real work to do instead of collapsing in two iterations
  * scores live where the bundled data's do (139..211): best hit 200+U{0..11}, the others lower by a geometric
    amount (floor 140), with a small share of exact ties with the best hit (initial `binmax` ties, as in real data)
"""
import concurrent.futures
import os

import numpy as np

CHUNK_ROWS = 1 << 20

_GEOM_P = 0.15
_TIE_P = 0.05


def _delta_table():
    """256-entry table mapping a uniform byte to the score deficit of a non-best alignment."""
    u = (np.arange(256) + 0.5) / 256.0
    tie = u < _TIE_P
    v = (u - _TIE_P) / (1.0 - _TIE_P)
    d = np.floor(np.log1p(-np.clip(v, 0, 1 - 1e-12)) / np.log1p(-_GEOM_P)).astype(np.int64) + 1
    d[tie] = 0
    return np.minimum(d, 80).astype(np.int16)


_DELTA = _delta_table()
_GAP = np.array([1, 1, 1, 2, 2, 2, 3, 4], dtype=np.int32)


def _abundance_cdf(n_cols, seed):
    rng = np.random.default_rng([seed, 0xAB])
    a = np.exp(1.5 * rng.standard_normal(n_cols))
    c = np.cumsum(a)
    return c / c[-1]


def _chunk_lengths(rng, rows, avg, skew, n_cols):
    uniq = rng.random(rows) < 0.2
    if skew:
        # Zipf(1.6) tail truncated to [2, 200]; the scale keeps the overall mean near `avg`
        lens = np.minimum(np.ceil(rng.zipf(1.6, rows) * (avg / 7.0)), 200).astype(np.int64)
        lens = np.maximum(lens, 2)
    else:
        lens = np.maximum(rng.poisson((avg - 0.2) / 0.8, rows), 2).astype(np.int64)
    lens[uniq] = 1
    return np.minimum(lens, max(1, n_cols // 8))


def _chunk(seed, c, rows, n_cols, avg, skew, cdf, want_data):
    rng = np.random.default_rng([seed, 1 + c])
    lens = _chunk_lengths(rng, rows, avg, skew, n_cols)
    if not want_data:
        return lens, None, None
    nnz = int(lens.sum())
    bits = rng.integers(0, 1 << 16, size=nnz, dtype=np.uint16)
    starts = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=starts[1:])
    row = np.repeat(np.arange(rows, dtype=np.int32), lens)
    # columns: home - span/2 + running sum of small positive gaps (strictly increasing => canonical CSR)
    gap = _GAP[(bits >> 8) & 7].astype(np.int64)
    run = np.cumsum(gap)
    first = run[starts[:-1]] - gap[starts[:-1]]
    span = run[starts[1:] - 1] - first
    home = np.searchsorted(cdf, rng.random(rows)).astype(np.int64)
    col0 = np.clip(home - span // 2, 0, n_cols - 1 - span)
    cols = (run - first[row] + col0[row] - gap[starts[:-1]][row]).astype(np.int32)
    # scores
    best = (200 + rng.integers(0, 12, size=rows)).astype(np.int16)
    score = np.maximum(best[row] - _DELTA[bits & 0xFF], 140).astype(np.uint16)
    bestpos = starts[:-1] + (rng.random(rows) * lens).astype(np.int64)
    score[bestpos] = best.astype(np.uint16)
    return lens, cols, score


def synth_csr(n_rows, n_cols, avg=20, skew=False, seed=1001, row_start=0, row_stop=None, threads=None):
    """Rows [row_start, row_stop) of the synthetic matrix -> (indptr int64 (local, starts at 0), indices int32, raw uint16)."""
    row_stop = n_rows if row_stop is None else row_stop
    assert 0 <= row_start <= row_stop <= n_rows
    cdf = _abundance_cdf(n_cols, seed)
    c0, c1 = row_start // CHUNK_ROWS, (max(row_stop, 1) - 1) // CHUNK_ROWS + 1
    chunks = list(range(c0, c1))
    rows_of = lambda c: min(CHUNK_ROWS, n_rows - c * CHUNK_ROWS)
    threads = threads or min(len(chunks), os.cpu_count() or 1, 32)

    def gen(c):
        lens, cols, score = _chunk(seed, c, rows_of(c), n_cols, avg, skew, cdf, True)
        lo = max(row_start - c * CHUNK_ROWS, 0)
        hi = min(row_stop - c * CHUNK_ROWS, rows_of(c))
        if lo > 0 or hi < rows_of(c):
            st = np.zeros(lens.size + 1, dtype=np.int64)
            np.cumsum(lens, out=st[1:])
            cols, score, lens = cols[st[lo]:st[hi]], score[st[lo]:st[hi]], lens[lo:hi]
        return lens, cols, score

    if threads > 1 and len(chunks) > 1:
        with concurrent.futures.ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(gen, chunks))
    else:
        parts = [gen(c) for c in chunks]
    if not parts:
        return np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.uint16)
    lens = np.concatenate([p[0] for p in parts])
    indptr = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    raw = np.empty(int(indptr[-1]), dtype=np.uint16)
    o = 0
    for _, cols, score in parts:
        indices[o:o + cols.size] = cols
        raw[o:o + cols.size] = score
        o += cols.size
    return indptr, indices, raw


def synth_row_nnz(n_rows, n_cols, avg=20, skew=False, seed=1001):
    """Per-row entry counts of the whole matrix (cheap; lets ranks pick nnz-balanced shard boundaries)."""
    out = []
    for c in range((n_rows + CHUNK_ROWS - 1) // CHUNK_ROWS):
        rows = min(CHUNK_ROWS, n_rows - c * CHUNK_ROWS)
        out.append(_chunk(seed, c, rows, n_cols, avg, skew, None, False)[0])
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


def shard_bounds(n_rows, n_shards):
    """Contiguous row ranges for the shards, aligned to generator chunks where possible so that no rank generates
    rows it does not own.  Row lengths are i.i.d., so equal row counts are nnz-balanced to well under 1 %."""
    if n_shards <= 1:
        return [(0, n_rows)]
    edges = [int(round(n_rows * i / n_shards)) for i in range(n_shards + 1)]
    return [(edges[i], edges[i + 1]) for i in range(n_shards)]
