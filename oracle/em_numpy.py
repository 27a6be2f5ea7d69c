# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Telescope EM path (lean numpy restatement).

This file restates, in flat numpy over the three CSR arrays, what the reference computes in
`telescope/utils/model.py:631-865` (class TelescopeLikelihood) and `telescope/utils/sparse_plus.py:16-165`
(csr_matrix_plus helpers).  It exists so that the CUDA path has something to be checked against on the GPU box,
where `/root/reference` is absent.  It is imported ONLY by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`; the product (`telescope_b200/`) never imports it and has no
CPU fallback.

Parity status: PINNED.  `tests/test_oracle.py` checks this restatement against
  * the reference's own known answers (README.md:70-71 final log-likelihood 95252.596293 on the bundled data;
    `telescope/data/telescope_report.tsv`; `tests/test_sparse_plus.py:24-55`; docstrings sparse_plus.py:33-41,106-115),
  * golden vectors produced by running the unmodified reference class in the build container
    (`tests/golden/make_golden.py` -> `tests/golden/*.npz`), and
  * the live reference whenever `/root/reference` is present.

Operation order follows the reference where it changes bits:
  Q   = expm1((raw * (1/max)) * 100.)                    model.py:652-653, sparse_plus.py:89-91
  n   = Q * (pi*theta)   (ambiguous rows)  |  Q * pi     (unique rows)          model.py:718-720
  z   = n * recip0(rowsum(n))             (multiply by reciprocal, 1/0 -> 0)    sparse_plus.py:16-22,52
  c   = (z * w) * Y ; thetasum = colsum(c)                                       model.py:730-733
Row sums are taken in storage order (as scipy's csr_matvec does); column sums are taken in row order.
"""
from __future__ import annotations

import numpy as np

REASSIGN_METHODS = ("exclude", "choose", "average", "conf", "unique", "all")


def q_lut(max_score, scale_factor=100.0):
    """LUT[s] = expm1((s * (1/max_score)) * 100.)  -- bit-identical to model.py:652-653 applied to uint16 s.

    `raw.scale()` is `raw.multiply(1. / raw.max())` (sparse_plus.py:91): uint16 data times a Python float gives
    float64(data) * float64(1/max); `.multiply(100.)` and `.expm1()` follow elementwise.
    """
    max_score = int(max_score)
    if max_score <= 0:
        return np.zeros(1, dtype=np.float64)
    s = np.arange(max_score + 1, dtype=np.float64)          # uint16 -> float64 is exact
    return np.expm1((s * (1.0 / max_score)) * scale_factor)


def recip0(v):
    """1/v with 1/0 -> 0 (sparse_plus.py:16-22)."""
    with np.errstate(divide="ignore"):
        r = 1.0 / v
    r[np.isinf(r)] = 0
    return r


def _row_ids(indptr):
    lens = np.diff(indptr)
    return np.repeat(np.arange(lens.size, dtype=np.int64), lens), lens


def _seg_sum(values, indptr):
    """Per-row sum in storage order.  np.add.reduceat mishandles empty segments, so mask them."""
    n_rows = indptr.size - 1
    out = np.zeros(n_rows, dtype=np.float64)
    if values.size == 0:
        return out
    lens = np.diff(indptr)
    nz = lens > 0
    starts = indptr[:-1][nz]
    out[nz] = np.add.reduceat(values, starts)
    return out


def _seg_max(values, indptr, n_cols):
    """Per-row max as scipy's sparse .max(1) gives it: implicit zeros count when the row is not full."""
    n_rows = indptr.size - 1
    out = np.zeros(n_rows, dtype=values.dtype)
    if values.size == 0:
        return out
    lens = np.diff(indptr)
    nz = lens > 0
    m = np.maximum.reduceat(values, indptr[:-1][nz])
    full = lens[nz] >= n_cols
    m = np.where(full, m, np.maximum(m, 0))
    out[nz] = m
    return out


class EMOracle(object):
    """Flat-array twin of TelescopeLikelihood (model.py:631-865).

    Attributes mirror the reference: N, K, Q (data array, same CSR structure as the input), Y (N,), pi, theta,
    pi_init, theta_init, z (data array, explicit zeros where the reference would have dropped the entry), lnl.
    """

    def __init__(self, indptr, indices, raw, n_cols, em_epsilon=1e-7, max_iter=100, pi_prior=0,
                 theta_prior=200000, max_score=None, lut=None):
        self.indptr = np.asarray(indptr).astype(np.int64)
        self.indices = np.asarray(indices)
        self.raw = np.asarray(raw)
        self.N = self.indptr.size - 1
        self.K = int(n_cols)
        # model.py:640 -- raw_scores.max() (sparse max: implicit zeros participate, scores are non-negative)
        self.max_score = int(self.raw.max()) if (max_score is None and self.raw.size) else int(max_score or 0)
        self.lut = q_lut(self.max_score) if lut is None else np.asarray(lut, dtype=np.float64)
        self.Q = self.lut[self.raw]                                           # model.py:653
        self.row, self.lens = _row_ids(self.indptr)
        self.epsilon = em_epsilon                                             # model.py:661-662
        self.max_iter = max_iter
        self.pi = np.repeat(1.0 / self.K, self.K)                             # model.py:667
        self.theta = np.repeat(1.0 / self.K, self.K)                          # model.py:673
        self.pi_init = None
        self.theta_init = None
        self.Y = (self.lens > 1).astype(np.uint8)                             # model.py:679
        self.lnl = float("inf")                                               # model.py:683
        self.pi_prior = pi_prior
        self.theta_prior = theta_prior
        self.weights = _seg_max(self.Q, self.indptr, self.K)                  # model.py:690
        self.total_wt = self.weights.sum()                                    # model.py:691
        self.ambig_wt = (self.weights * self.Y).sum()                         # model.py:692
        wmax = self.weights.max() if self.N else 0.0
        self.pi_prior_wt = self.pi_prior * wmax                               # model.py:696
        self.theta_prior_wt = self.theta_prior * wmax                         # model.py:697
        self._yk = self.Y[self.row].astype(bool)                              # per-entry ambiguity flag
        uniq = ~self._yk
        self.pisum0 = np.bincount(self.indices[uniq], weights=self.Q[uniq], minlength=self.K)  # model.py:699
        self.z = None
        self.n_iter = 0
        self.converged = False
        self.diffs = []
        self.lnls = []

    # ---- model.py:702-722
    def _numerator(self, pi, theta):
        pt = pi * theta
        return self.Q * np.where(self._yk, pt[self.indices], pi[self.indices])

    def estep(self, pi, theta):
        n = self._numerator(pi, theta)
        r = recip0(_seg_sum(n, self.indptr))
        return n * r[self.row]

    # ---- model.py:724-742
    def mstep(self, z):
        c = (z * self.weights[self.row]) * self.Y[self.row]
        thetasum = np.bincount(self.indices, weights=c, minlength=self.K)
        theta_hat = (thetasum + self.theta_prior_wt) / (self.ambig_wt + self.theta_prior_wt * self.K)
        pi_hat = ((self.pisum0 + thetasum) + self.pi_prior_wt) / (self.total_wt + self.pi_prior_wt * self.K)
        return pi_hat, theta_hat

    # ---- model.py:744-760
    def calculate_lnl(self, z, pi, theta):
        return float((z * np.log1p(self._numerator(pi, theta))).sum())

    # ---- model.py:762-806
    def em(self, use_likelihood=False):
        inum, converged, reached_max = 0, False, False
        self.diffs, self.lnls = [], []
        while not (converged or reached_max):
            _z = self.estep(self.pi, self.theta)
            _pi, _theta = self.mstep(_z)
            inum += 1
            if inum == 1:
                self.pi_init, self.theta_init = _pi, _theta
            diff_est = np.abs(_pi - self.pi).sum()
            self.diffs.append(float(diff_est))
            if use_likelihood:
                _lnl = self.calculate_lnl(_z, _pi, _theta)
                converged = abs(_lnl - self.lnl) < self.epsilon
                self.lnl = _lnl
                self.lnls.append(float(_lnl))
            else:
                converged = diff_est < self.epsilon
            reached_max = inum >= self.max_iter
            self.z = _z
            self.pi, self.theta = _pi, _theta
        if not use_likelihood:
            self.lnl = self.calculate_lnl(self.z, self.pi, self.theta)
        self.n_iter, self.converged = inum, bool(converged)
        return self

    # ---- model.py:808-865 + sparse_plus.py:99-165
    def initial_z(self):
        """Q.norm(1) (model.py:837)."""
        return self.Q * recip0(_seg_sum(self.Q, self.indptr))[self.row]

    def _binmax(self, z):
        """sparse_plus.py:99-129 on a matrix whose exact-zero entries are not stored."""
        zmax = _seg_max(z, self.indptr, self.K)
        return (z == zmax[self.row]) & (z != 0)

    def reassign_data(self, method, thresh=0.9, initial=False, rng_state=None):
        """Per-entry assignment values in the input CSR order (zeros where the reference stores nothing).

        `choose` consumes the legacy global numpy RNG exactly as sparse_plus.py:146-153 does (one
        np.random.choice(range(d_start, d_end)) per row with more than one best hit, rows in order).
        """
        if method not in REASSIGN_METHODS:
            raise ValueError('Argument "method" should be one of (exclude, choose, average, conf, unique, all)')
        z = self.initial_z() if initial else self.z
        if method in ("exclude", "choose", "average"):
            best = self._binmax(z)
            nbest = np.zeros(self.N, dtype=np.int64)
            np.add.at(nbest, self.row[best], 1)
            if method == "exclude":
                return (best & (nbest[self.row] == 1)).astype(np.int8)
            if method == "average":
                return best * recip0(nbest.astype(np.float64))[self.row]
            out = best.astype(np.int8)
            pos = np.flatnonzero(best)                        # positions of best entries, row-major
            prow = self.row[pos]
            starts = np.flatnonzero(np.r_[True, prow[1:] != prow[:-1]]) if pos.size else np.array([], dtype=np.int64)
            ends = np.r_[starts[1:], pos.size] if pos.size else starts
            for a, b in zip(starts, ends):
                if b - a > 1:
                    chosen = np.random.choice(range(a, b))
                    for j in range(a, b):
                        if j != chosen:
                            out[pos[j]] = 0
            return out
        if method == "conf":
            v = np.where(z >= thresh, z, 0.0)
            return v * recip0(_seg_sum(v, self.indptr))[self.row]
        if method == "unique":
            return np.ceil(z * (1 - self.Y[self.row])).astype(np.uint8)
        return (z > 0).astype(np.uint8)                       # 'all'

    def reassign_colsum(self, method, thresh=0.9, initial=False):
        """`reassign(...).sum(0).A1` (model.py:435-441,457)."""
        d = self.reassign_data(method, thresh, initial)
        if d.dtype.kind in "iu":
            out = np.zeros(self.K, dtype=np.int64 if d.dtype.kind == "i" else np.uint64)
            np.add.at(out, self.indices, d.astype(out.dtype))
            return out
        return np.bincount(self.indices, weights=d, minlength=self.K)
