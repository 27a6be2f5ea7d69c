# -*- coding: utf-8 -*-
"""Where the per-launch fixed cost of the per-iteration kernel goes: start / end time of every one-warp CTA of
k_ell<ELL_FUSED> (debug build with -DTSC_ELL_TRACE, see profiles/README.md).

    python -c "from telescope_b200 import build as b; b.build_library(out='build/variants/libtsc_trace.so', defines=['TSC_ELL_TRACE'])"
    TELESCOPE_B200_LIB=build/variants/libtsc_trace.so python tools/profile/trace_ell.py [--reads 6250000]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from telescope_b200.likelihood import TelescopeLikelihood          # noqa: E402
from telescope_b200.synthetic import synth_csr                      # noqa: E402


class Opts(object):
    em_epsilon, max_iter, pi_prior, theta_prior = -1.0, 6, 0, 200000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=6_250_000)       # one GPU's share of the benchmark matrix at 8 GPUs
    ap.add_argument("--loci", type=int, default=30_000)
    ap.add_argument("--iters", type=int, default=6)
    a = ap.parse_args()
    Opts.max_iter = a.iters
    ip, ix, raw = synth_csr(a.reads, a.loci, 20, False, 1004)
    m = sp.csr_matrix((raw, ix, ip.astype("int32")), shape=(a.reads, a.loci))
    tl = TelescopeLikelihood(m, Opts, devices=[0])
    tl.em()
    n = tl.layout_stats()["stream_ctas"]
    buf = np.zeros(2 * n, dtype=np.uint64)
    fn = tl._lib.tsc_debug_ell_trace
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    rc = fn(tl._h, buf.ctypes.data_as(C.c_void_p), n)
    assert rc == 0
    t = buf.reshape(n, 2).astype(np.int64)
    t0 = t[:, 0].min()
    start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
    q = lambda v: " ".join("%.1f" % x for x in np.percentile(v, [0, 1, 10, 50, 90, 99, 100]))
    print("kernel times (ms):", tl.kernel_times_ms())
    print("CTAs %d; us since the first CTA started, percentiles 0 1 10 50 90 99 100" % n)
    print("  start:", q(start))
    print("  end  :", q(end))
    print("  busy :", q(end - start))
    print("  last end - median end: %.1f us; last end - first start: %.1f us" % (end.max() - np.median(end), end.max()))
    order = np.argsort(end)
    print("  slowest CTAs:", order[-8:], " fastest:", order[:8])
    tl.close()


if __name__ == "__main__":
    main()
