# -*- coding: utf-8 -*-
"""`csr_matrix_plus`: the sparse-matrix type that crosses the TelescopeLikelihood boundary.

The reference's class of the same name (telescope/utils/sparse_plus.py:24-174) is what `Telescope` hands to
`TelescopeLikelihood` (raw_scores) and what comes back (`Q`, `z`, `reassign()` results), so the type and its helper
methods are part of the drop-in surface.  This is a fresh, vectorised host-side implementation with the same
results; the reference's versions loop over rows / entries in Python (sparse_plus.py:122-125,147-152,163-164).
The EM loop itself never calls these helpers -- it runs in libtelescope_b200.so.
"""
import numpy as np
import scipy.sparse


def _recip0(v):
    """Reciprocal with 1/0 -> 0 (reference sparse_plus.py:16-22)."""
    with np.errstate(divide="ignore"):
        out = 1.0 / v
    out[np.isinf(out)] = 0
    return out


class csr_matrix_plus(scipy.sparse.csr_matrix):

    def _rows(self):
        return np.repeat(np.arange(self.shape[0]), np.diff(self.indptr))

    def norm(self, axis=None):
        """Normalise the whole matrix (axis=None) or each row (axis=1) to sum 1; all-zero rows stay zero."""
        if axis is None:
            return type(self)(self.multiply(1.0 / self.sum()))
        if axis == 1:
            return type(self)(self.multiply(_recip0(self.sum(1))))
        raise NotImplementedError

    def scale(self, axis=None):
        """Divide by the global maximum (axis=None) or by each row's maximum (axis=1)."""
        if axis is None:
            return type(self)(self.multiply(1.0 / self.max()))
        if axis == 1:
            return type(self)(self.multiply(_recip0(self.max(1).toarray())))
        raise NotImplementedError

    def binmax(self, axis=None):
        """1 where an entry equals its row's maximum, everything else dropped (int8)."""
        if axis != 1:
            raise NotImplementedError
        rowmax = self.max(1).toarray().ravel()
        hit = (self.data == rowmax[self._rows()]).astype(np.int8)
        ret = type(self)((hit, self.indices.copy(), self.indptr.copy()), shape=self.shape)
        ret.eliminate_zeros()
        return ret

    def count(self, axis=None):
        """Stored entries per row as an N x 1 array."""
        if axis != 1:
            raise NotImplementedError
        return np.array(np.diff(self.indptr), ndmin=2).T

    def choose_random(self, axis=None):
        """Keep one stored entry per row, chosen with the global numpy RNG; rows with one entry draw nothing.

        Stream-compatible with the reference (sparse_plus.py:146-153): one `np.random.choice(range(a, b))` per row
        with more than one entry, in row order, is the same draw as `a + np.random.randint(0, b - a)`.
        """
        if axis != 1:
            raise NotImplementedError
        ret = self.copy()
        lens = np.diff(ret.indptr)
        multi = np.flatnonzero(lens > 1)
        if multi.size:
            picks = draw_picks(lens[multi])
            keep = np.ones(ret.data.size, dtype=bool)
            rows = ret._rows()
            keep[np.isin(rows, multi)] = False
            keep[ret.indptr[multi] + picks] = True
            ret.data[~keep] = 0
        ret.eliminate_zeros()
        return ret

    def check_equal(self, other):
        if self.shape != other.shape:
            return False
        return (self != other).nnz == 0

    def apply_func(self, func):
        ret = self.copy()
        ret.data = np.fromiter((func(v) for v in self.data), self.data.dtype, count=len(self.data))
        return ret

    def save(self, filename):
        np.savez(filename, data=self.data, indices=self.indices, indptr=self.indptr, shape=self.shape)

    @classmethod
    def load(cls, filename):
        loader = np.load(filename)
        return cls((loader["data"], loader["indices"], loader["indptr"]), shape=loader["shape"])


def draw_picks(counts):
    """One uniform draw in [0, n) per element of `counts`, consuming the legacy global numpy RNG exactly as a
    sequence of `np.random.choice(range(n))` calls would (one bounded draw each, in order)."""
    counts = np.asarray(counts, dtype=np.int64)
    if counts.size == 0:
        return counts.copy()
    # np.random.choice(range(a, b)) is one bounded draw randint(0, b-a); randint with an array of bounds walks the
    # same generator state element by element (checked in tests/test_host.py against the sequential form).  For long
    # arrays (25 M tie reads at 50 M reads) the library walks the same MT19937 state in C, ~5x faster.
    if counts.size >= 4096:
        picks = _draw_native(counts, rows=False)
        if picks is not None:
            return picks
    return np.random.randint(0, counts)


def draw_row_picks(nbest):
    """Tie-breaks of reassign('choose') for every read at once: `nbest` holds the number of best hits per read; reads
    with more than one get a draw in [0, nbest), in read order -- the draws `draw_picks(nbest[nbest > 1])` makes, without
    the gather / scatter through the index of the tie reads -- the others get 0 and consume nothing.  int32 array."""
    nbest = np.ascontiguousarray(nbest, dtype=np.int32)
    if nbest.size >= 4096:
        picks = _draw_native(nbest, rows=True)
        if picks is not None:
            return picks
    picks = np.zeros(nbest.size, dtype=np.int32)
    t = np.flatnonzero(nbest > 1)
    if t.size:
        picks[t] = np.random.randint(0, nbest[t].astype(np.int64))
    return picks


def _draw_native(counts, rows):
    """tsc_mt19937_draw_picks / _rows on numpy's global legacy generator (stream-identical to np.random.randint with an
    array of bounds, tests/test_host.py); None when the library or the generator state is not usable."""
    import ctypes as C
    try:
        from . import _abi
        lib = _abi.load()
        fn = lib.tsc_mt19937_draw_rows if rows else lib.tsc_mt19937_draw_picks
    except Exception:
        return None
    st = np.random.get_state()
    if st[0] != 'MT19937':
        return None
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(st[2]))
    ctype, dtype = (C.c_int32, np.int32) if rows else (C.c_int64, np.int64)
    counts = np.ascontiguousarray(counts, dtype=dtype)
    if not rows and counts.size and counts.min() < 1:
        return None                                   # numpy raises for an empty range: let it
    picks = np.empty(counts.size, dtype=np.int32)
    fn.restype = C.c_int
    rc = fn(key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(pos), counts.ctypes.data_as(C.POINTER(ctype)),
            C.c_int64(counts.size), picks.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        return None                                   # (a bound of 2^31 or more: numpy's own path handles it)
    np.random.set_state(('MT19937', key, pos.value, st[3], st[4]))
    return picks
