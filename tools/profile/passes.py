# -*- coding: utf-8 -*-
"""Profiling driver: one model on the benchmark matrix (smaller by default), a few EM iterations, then each device
pass once -- a short, predictable launch sequence for `ncu` (see profiles/README.md for the exact commands).

    python tools/profile/passes.py [--reads 10000000] [--skew]
"""
import argparse
import os
import sys

import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from telescope_b200.likelihood import TelescopeLikelihood          # noqa: E402
from telescope_b200.synthetic import synth_csr                      # noqa: E402


class Opts(object):
    em_epsilon, max_iter, pi_prior, theta_prior = -1.0, 6, 0, 200000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--loci", type=int, default=30_000)
    ap.add_argument("--skew", action="store_true")
    ap.add_argument("--kernel", default="auto")
    a = ap.parse_args()
    ip, ix, raw = synth_csr(a.reads, a.loci, 20, a.skew, 1004)
    m = sp.csr_matrix((raw, ix, ip.astype("int32")), shape=(a.reads, a.loci))
    tl = TelescopeLikelihood(m, Opts, devices=[0], kernel=a.kernel)
    tl.em()                                     # 6 x (fused [+ re-partition] [+ residual tiles] + tail) + final lnl
    for name in ("fused", "estep", "lnl", "reassign"):
        print(name, tl.time_pass(name, 1))      # warm-up launch + 1 timed launch each
    print(tl.layout_stats(), tl.lnl)
    tl.close()


if __name__ == "__main__":
    main()
