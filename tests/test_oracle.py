"""CPU tier: pin the oracle (oracle/em_numpy.py, oracle/em_scipy.py) to the reference.

1. known answers the reference itself publishes: README.md:70-71 (final log-likelihood 95252.596293 on the bundled
   data), telescope/data/telescope_report.tsv, tests/test_sparse_plus.py:24-55, docstrings sparse_plus.py:33-41,106-115
2. golden vectors written by tests/golden/make_golden.py from the unmodified reference
3. the live reference, when /root/reference exists (build container only)
"""
import csv
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, rel_err
from oracle.em_numpy import EMOracle, q_lut, recip0
from oracle.em_scipy import ScipyEM
from oracle.ref_shim import RefOpts, import_reference, reference_available

GOLD = os.path.join(ROOT, "tests", "golden")
REPORT_CALLS = [("conf", 0.9, False), ("all", 0.9, True), ("unique", 0.9, False), ("exclude", 0.9, True),
                ("choose", 0.9, True), ("average", 0.9, True), ("exclude", 0.9, False)]


def load_case(path):
    g = np.load(path)
    return {k: g[k] for k in g.files}


def oracle_for(g):
    return EMOracle(g["indptr"], g["indices"], g["raw"], int(g["shape"][1]), float(g["em_epsilon"]), int(g["max_iter"]),
                    float(g["pi_prior"]), float(g["theta_prior"]))


CASES = sorted(glob.glob(os.path.join(GOLD, "case_*.npz"))) + [os.path.join(GOLD, "bundled.npz")]


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_numpy_oracle_matches_reference_golden(path):
    g = load_case(path)
    o = oracle_for(g)
    assert np.array_equal(o.Y, g["Y"]) and np.array_equal(o.weights, g["weights"])
    assert rel_err(o.pisum0, g["pisum0"]) < 1e-13
    o.em(use_likelihood=bool(g["use_likelihood"]))
    assert o.n_iter == int(g["n_iter"])
    assert rel_err(o.diffs, g["diffs"]) < 1e-9
    assert rel_err(o.pi, g["pi"]) < 1e-10 and rel_err(o.theta, g["theta"]) < 1e-10
    assert rel_err(o.pi_init, g["pi_init"]) < 1e-12 and rel_err(o.theta_init, g["theta_init"]) < 1e-12
    assert abs(o.lnl - float(g["lnl"])) <= 1e-12 * abs(float(g["lnl"]))
    assert rel_err(o.z, g["z"]) < 1e-9
    assert rel_err(o.initial_z(), g["z_init"]) < 1e-13
    if bool(g["use_likelihood"]):
        assert rel_err(o.lnls, g["lnls"]) < 1e-12
    np.random.seed(int(g["seed"]))
    for (meth, th, ini), ref in zip(REPORT_CALLS, g["colsums"]):
        got = o.reassign_colsum(meth, th, ini).astype(np.float64)
        if meth in ("average", "conf"):
            assert rel_err(got, ref) < 1e-9, meth
        else:
            assert np.array_equal(got, ref), meth


def test_bundled_known_answers():
    g = load_case(os.path.join(GOLD, "bundled.npz"))
    o = oracle_for(g).em()
    assert "%.6f" % o.lnl == "95252.596293"                     # reference README.md:70-71
    assert o.n_iter == 16
    trace = [1.3795, 0.7388, 0.24275, 0.065133, 0.017653, 0.0050039, 0.0014807, 0.00045365, 0.00014275, 4.5844e-05,
             1.4953e-05, 4.9364e-06, 1.645e-06, 5.5228e-07, 1.8652e-07, 6.3301e-08]
    assert ["%.5g" % d for d in o.diffs] == ["%.5g" % d for d in trace]
    # reference telescope/data/telescope_report.tsv (v1.0.2 layout): every numeric column except transcript_length
    rows = list(csv.reader(open(os.path.join(ROOT, "telescope_b200", "data", "telescope_report.tsv")), delimiter="\t"))
    hdr, body = rows[1], {r[0]: dict(zip(rows[1], r)) for r in rows[2:]}
    names = [str(n) for n in g["feat_names"]]
    np.random.seed(int(g["seed"]))
    cols = {k: o.reassign_colsum(m, th, ini) for k, (m, th, ini) in zip(
        ["final_conf", "init_aligned", "unique_count", "init_best", "init_best_random", "init_best_avg", "final_count"], REPORT_CALLS)}
    fmt = {"final_count": "%d", "final_conf": "%.2f", "init_aligned": "%d", "unique_count": "%d", "init_best": "%d",
           "init_best_random": "%d", "init_best_avg": "%.2f"}
    assert len(body) == len(names) == 59
    for j, name in enumerate(names):
        for k, f in fmt.items():
            assert float(f % cols[k][j]) == float(body[name][k]), (name, k)
        assert float("%.3g" % o.pi[j]) == float(body[name]["final_prop"])
        assert float("%.3g" % o.pi_init[j]) == float(body[name]["init_prop"])


@pytest.mark.parametrize("name", ["default", "priors", "empty"])
def test_scipy_port_matches_golden(name):
    g = load_case(os.path.join(GOLD, "case_%s.npz" % name))
    m = sp.csr_matrix((g["raw"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    s = ScipyEM(m, float(g["em_epsilon"]), int(g["max_iter"]), float(g["pi_prior"]), float(g["theta_prior"])).em()
    assert s.n_iter == int(g["n_iter"])
    assert np.array_equal(s.pi, g["pi"]) and np.array_equal(s.theta, g["theta"]), "the port runs the reference's exact op sequence"
    assert float(s.lnl) == float(g["lnl"])


def test_q_lut_and_recip0():
    lut = q_lut(211)
    assert lut[0] == 0.0 and lut[211] == np.expm1(100.0)
    assert lut[139] == np.expm1((139 * (1.0 / 211)) * 100.0)
    assert np.array_equal(recip0(np.array([2.0, 0.0, 4.0])), np.array([0.5, 0.0, 0.25]))


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_live_reference_agrees_on_fresh_input():
    M, csr_plus = import_reference()
    from telescope_b200.synthetic import synth_csr
    ip, ix, raw = synth_csr(4000, 120, 7, True, 777)
    m = csr_plus((raw, ix, ip), shape=(4000, 120))
    opts = RefOpts(max_iter=12, pi_prior=1, theta_prior=50)
    tl = M.TelescopeLikelihood(m, opts)
    tl.em()
    o = EMOracle(ip, ix, raw, 120, opts.em_epsilon, opts.max_iter, opts.pi_prior, opts.theta_prior).em()
    assert np.array_equal(o.Q, tl.Q.data)
    assert rel_err(o.pi, tl.pi) < 1e-10 and rel_err(o.theta, tl.theta) < 1e-10
    assert abs(o.lnl - tl.lnl) <= 1e-12 * abs(tl.lnl)
    s = ScipyEM(m, opts.em_epsilon, opts.max_iter, opts.pi_prior, opts.theta_prior).em()
    assert np.array_equal(s.pi, tl.pi) and float(s.lnl) == float(tl.lnl)
    for meth in ("exclude", "average", "conf", "unique", "all"):
        a = np.asarray(tl.reassign(meth, 0.9).sum(0)).ravel()
        b = o.reassign_colsum(meth, 0.9)
        assert rel_err(b.astype(np.float64), a.astype(np.float64)) < 1e-9
