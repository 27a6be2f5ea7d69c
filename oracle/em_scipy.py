# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- scipy.sparse port of the reference EM loop, used as the timed CPU baseline.

`/root/reference` cannot travel to the GPU box, so `bench.py`'s `cpu_baseline` leg and `--impl reference` arm time
this port instead (`cpu_baseline.kind == "port"`).  Unlike `oracle/em_numpy.py` (a lean single-pass restatement
used as the parity checker), this file deliberately performs the SAME sequence of scipy.sparse / numpy calls as
`telescope/utils/model.py:635-806` and `telescope/utils/sparse_plus.py:16-52`, so that its wall time is the
reference's wall time: four broadcasting `multiply` calls (CSR->COO->CSR round trips), one sparse add and a
`sum(1)` per E-step; two `multiply` and a `sum(0)` per M-step.  All of it is single-threaded inside scipy's
`_sparsetools`, exactly like the reference.  `tests/test_oracle.py` checks it against the live reference (when
present) and the golden vectors; `DESIGN.md` records the measured time ratio port/reference in the build container.

Never imported by the product.
"""
import numpy as np
import scipy.sparse as sp

from .em_numpy import recip0


def _csr(m):
    return sp.csr_matrix(m)


def _row_normalise(m):
    # sparse_plus.py:52 -- multiply by the reciprocal of the row sums, 1/0 -> 0
    return _csr(m.multiply(recip0(m.sum(1))))


class ScipyEM(object):
    """Op-for-op port of TelescopeLikelihood.__init__/estep/mstep/calculate_lnl/em."""

    def __init__(self, raw_scores, em_epsilon=1e-7, max_iter=100, pi_prior=0, theta_prior=200000):
        raw_scores = _csr(raw_scores)
        self.N, self.K = raw_scores.shape
        top = raw_scores.max()                                                       # model.py:640
        self.Q = _csr(raw_scores.multiply(1.0 / top)).multiply(100.0).expm1()        # model.py:652-653
        self.Q = _csr(self.Q)
        self.epsilon, self.max_iter = em_epsilon, max_iter
        self.pi = np.repeat(1.0 / self.K, self.K)
        self.theta = np.repeat(1.0 / self.K, self.K)
        self.pi_init = self.theta_init = None
        nper = np.diff(self.Q.indptr)
        self.Y = (np.array(nper, ndmin=2).T > 1).astype(np.uint8)                    # model.py:679
        self.lnl = float("inf")
        self.w = self.Q.max(1)                                                       # model.py:690 (COO N x 1)
        self.total_wt = self.w.sum()
        self.ambig_wt = self.w.multiply(self.Y).sum()
        self.pi_prior_wt = pi_prior * self.w.max()
        self.theta_prior_wt = theta_prior * self.w.max()
        self.pisum0 = self.Q.multiply(1 - self.Y).sum(0)                             # model.py:699
        self.z = None
        self.diffs = []
        self.n_iter = 0
        self.converged = False

    def _inner(self, pi, theta):
        amb = _csr(self.Q.multiply(self.Y)).multiply(pi * theta)                     # model.py:718 / 755
        uni = _csr(self.Q.multiply(1 - self.Y)).multiply(pi)                         # model.py:719 / 756
        return _csr(amb + uni)

    def estep(self, pi, theta):
        return _row_normalise(self._inner(pi, theta))                                # model.py:720-722

    def mstep(self, z):
        thetasum = z.multiply(self.w).multiply(self.Y).sum(0)                        # model.py:730-733
        theta_hat = (thetasum + self.theta_prior_wt) / (self.ambig_wt + self.theta_prior_wt * self.K)
        pisum = self.pisum0 + thetasum
        pi_hat = (pisum + self.pi_prior_wt) / (self.total_wt + self.pi_prior_wt * self.K)
        return pi_hat.A1, theta_hat.A1

    def calculate_lnl(self, z, pi, theta):
        return z.multiply(self._inner(pi, theta).log1p()).sum()                      # model.py:755-758

    def em(self, use_likelihood=False):
        k, done, capped = 0, False, False
        self.diffs = []
        while not (done or capped):
            z = self.estep(self.pi, self.theta)
            pi, theta = self.mstep(z)
            k += 1
            if k == 1:
                self.pi_init, self.theta_init = pi, theta
            d = abs(pi - self.pi).sum()                                              # model.py:781
            self.diffs.append(float(d))
            if use_likelihood:
                lnl = self.calculate_lnl(z, pi, theta)
                done = abs(lnl - self.lnl) < self.epsilon
                self.lnl = lnl
            else:
                done = d < self.epsilon
            capped = k >= self.max_iter
            self.z, self.pi, self.theta = z, pi, theta
        if not use_likelihood:
            self.lnl = self.calculate_lnl(self.z, self.pi, self.theta)
        self.n_iter, self.converged = k, bool(done)
        return self
