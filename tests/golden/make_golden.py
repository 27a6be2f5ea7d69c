#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Generate the committed golden vectors by running the UNMODIFIED reference (mlbendall/telescope @ 4cf18595,
/root/reference) in the build container.  The GPU box has no /root/reference; tests there compare against these
files.  Re-run with:  python tests/golden/make_golden.py

Outputs (tests/golden/):
  bundled.npz         score matrix of the bundled BAM/GTF (built by telescope_b200.host, the reference's own loader
                      needs pysam) + everything the reference's TelescopeLikelihood produces on it with default
                      options: pi, theta, pi_init, theta_init, lnl, per-iteration diffs, z, the seven reassign column
                      sums of output_report (with the RNG seeded as telescope_assign.run does)
  bundled_run_stats.tsv / bundled_TE_counts.tsv   the reference's Telescope.output_report on that run
  case_*.npz          small seeded synthetic matrices (stored, not regenerated) with the reference's results under
                      different options (priors, use_likelihood, max_iter cut-off, Zipf rows, empty reads,
                      duplicated loci that produce exact posterior ties)
"""
import os
import sys
from collections import Counter, OrderedDict

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import RefOpts, import_reference  # noqa: E402
from telescope_b200.synthetic import synth_csr  # noqa: E402

M, csr_plus = import_reference()

REPORT_CALLS = [("conf", 0.9, False), ("all", 0.9, True), ("unique", 0.9, False), ("exclude", 0.9, True),
                ("choose", 0.9, True), ("average", 0.9, True), ("exclude", 0.9, False)]


def aligned(z, m):
    """z's data in m's entry order, zeros where z stores nothing."""
    z = sp.csr_matrix(z)
    z.sort_indices()
    K = np.int64(m.shape[1])
    full = np.repeat(np.arange(m.shape[0], dtype=np.int64), np.diff(m.indptr)) * K + m.indices
    sub = np.repeat(np.arange(m.shape[0], dtype=np.int64), np.diff(z.indptr)) * K + z.indices
    out = np.zeros(m.nnz)
    out[np.searchsorted(full, sub)] = z.data
    return out


def run_reference(m, opts, use_likelihood=False, seed=0):
    m = csr_plus(m)
    tl = M.TelescopeLikelihood(m, opts)
    lnls, diffs = [], []
    import logging

    class Grab(logging.Handler):
        def emit(self, rec):
            msg = rec.getMessage()
            if msg.startswith("Iteration"):
                parts = dict(p.strip().split("=") for p in msg.split(",")[1:])
                diffs.append(float(parts["diff"]))
    # the log line rounds to 5 significant digits; recompute exact diffs by stepping the reference ourselves
    pi, theta = tl.pi.copy(), tl.theta.copy()
    exact, exact_lnl, lnl_prev = [], [], float("inf")
    for it in range(max(1, opts.max_iter)):
        z = tl.estep(pi, theta)
        npi, ntheta = tl.mstep(z)
        d = abs(npi - pi).sum()
        exact.append(d)
        pi, theta = npi, ntheta
        if use_likelihood:
            l = tl.calculate_lnl(z, pi, theta)
            exact_lnl.append(l)
            if abs(l - lnl_prev) < opts.em_epsilon:
                break
            lnl_prev = l
        elif d < opts.em_epsilon:
            break
    tl.em(use_likelihood=use_likelihood, loglev=logging.DEBUG)
    assert np.array_equal(tl.pi, pi) and np.array_equal(tl.theta, theta), "manual stepping must equal em()"
    np.random.seed(seed)
    colsums = [np.asarray(tl.reassign(meth, th, ini).sum(0)).ravel().astype(np.float64) for meth, th, ini in REPORT_CALLS]
    out = dict(
        indptr=m.indptr.astype(np.int64), indices=m.indices.astype(np.int32), raw=m.data.astype(np.uint16),
        shape=np.array(m.shape, dtype=np.int64),
        em_epsilon=opts.em_epsilon, max_iter=opts.max_iter, pi_prior=opts.pi_prior, theta_prior=opts.theta_prior,
        use_likelihood=int(use_likelihood), seed=seed,
        pi=tl.pi, theta=tl.theta, pi_init=tl.pi_init, theta_init=tl.theta_init, lnl=float(tl.lnl),
        diffs=np.array(exact), lnls=np.array(exact_lnl), n_iter=len(exact),
        z=aligned(tl.z, m), z_init=aligned(tl.Q.norm(1), m),
        Y=tl.Y.ravel(), weights=np.asarray(tl._weights.todense()).ravel(),
        total_wt=float(tl._total_wt), ambig_wt=float(tl._ambig_wt), pisum0=np.asarray(tl._pisum0).ravel(),
        colsums=np.array(colsums),
    )
    return tl, out


def bundled():
    from telescope_b200.host.annotation import Annotation
    from telescope_b200.host.telescope import Telescope

    class O(object):
        samfile = os.path.join(ROOT, "telescope_b200", "data", "alignment.bam")
        gtffile = os.path.join(ROOT, "telescope_b200", "data", "annotation.gtf")
        no_feature_key, overlap_threshold, overlap_mode, stranded_mode, ncpu = "__no_feature", 0.2, "threshold", "None", 1
        version = "GOLDEN"
    ts = Telescope(O)
    ts.load_alignment(Annotation(O.gtffile, "locus", "None"))
    opts = RefOpts()
    seed = ts.get_random_seed()
    tl, out = run_reference(ts.raw_scores, opts, seed=seed)
    assert "%.6f" % tl.lnl == "95252.596293", tl.lnl      # reference README.md:70-71
    fnames = sorted(ts.feat_index, key=ts.feat_index.get)
    out["feat_names"] = np.array(fnames)
    out["feat_lengths"] = np.array([ts.feature_length[f] for f in fnames], dtype=np.int64)
    out["run_info_keys"] = np.array(list(ts.run_info.keys()))
    out["run_info_vals"] = np.array([str(v) for v in ts.run_info.values()])
    np.savez_compressed(os.path.join(HERE, "bundled.npz"), **out)
    # the reference's own report writer on the reference's own model
    rts = M.Telescope.__new__(M.Telescope)
    rts.opts = opts
    rts.run_info = OrderedDict(ts.run_info)
    rts.shape = ts.shape
    rts.feat_index = dict(ts.feat_index)
    rts.feature_length = Counter(ts.feature_length)
    rts.read_index = dict(ts.read_index)
    rts.raw_scores = csr_plus(ts.raw_scores)
    np.random.seed(seed)
    rts.output_report(tl, os.path.join(HERE, "bundled_run_stats.tsv"), os.path.join(HERE, "bundled_TE_counts.tsv"))
    rts.save(os.path.join(HERE, "bundled_checkpoint"))
    print("bundled: %s nnz=%d iters=%d lnl=%.6f" % (ts.shape, ts.raw_scores.nnz, out["n_iter"], tl.lnl))


def synthetic_cases():
    def mat(N, K, avg, skew, seed):
        ip, ix, raw = synth_csr(N, K, avg, skew, seed)
        return sp.csr_matrix((raw, ix, ip), shape=(N, K))
    cases = OrderedDict()
    cases["default"] = (mat(2500, 80, 6, False, 101), RefOpts(max_iter=100), False)
    cases["priors"] = (mat(2000, 64, 8, False, 102), RefOpts(max_iter=40, pi_prior=4, theta_prior=11), False)
    cases["likelihood"] = (mat(2000, 64, 8, False, 103), RefOpts(max_iter=30, em_epsilon=1e-2), True)
    cases["cutoff"] = (mat(3000, 150, 10, False, 104), RefOpts(max_iter=7), False)
    cases["zipf"] = (mat(1500, 900, 20, True, 105), RefOpts(max_iter=15), False)
    # empty reads and a locus nobody maps to
    m = mat(1200, 50, 5, False, 106).tolil()
    for r in (0, 17, 600, 1199):
        m.rows[r], m.data[r] = [], []
    m = sp.csr_matrix(m)
    m = sp.csr_matrix(sp.hstack([m, sp.csr_matrix((m.shape[0], 3), dtype=m.dtype)]))
    cases["empty"] = (m, RefOpts(max_iter=20), False)
    # duplicated loci: locus K+j is an exact copy of locus j for a third of the loci -> exact ties in the final z
    m = mat(2000, 60, 5, False, 107)
    dup = m[:, :20]
    m2 = sp.csr_matrix(sp.hstack([m, dup]))
    m2.sort_indices()
    cases["duploci"] = (m2, RefOpts(max_iter=30), False)
    for name, (m, opts, ul) in cases.items():
        m = sp.csr_matrix(m).astype(np.uint16)
        m.sort_indices()
        _, out = run_reference(m, opts, use_likelihood=ul, seed=12345)
        np.savez_compressed(os.path.join(HERE, "case_%s.npz" % name), **out)
        ties = int((out["colsums"][6].sum() != (out["Y"].size)))
        print("case %-10s shape=%s nnz=%d iters=%d lnl=%.6f final_exclude_sum=%d" % (
            name, m.shape, m.nnz, out["n_iter"], out["lnl"], out["colsums"][6].sum()))


if __name__ == "__main__":
    bundled()
    synthetic_cases()
