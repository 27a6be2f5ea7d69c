# -*- coding: utf-8 -*-
"""Shared-memory bank conflicts of the long-read kernel (k_ell_long, tsc_ell.cuh) on the Zipf matrix, emulated on the CPU.

In k_ell_long a warp works on one read: lane l owns entries l, l + 32, ... and every entry makes three 8-byte
shared-memory accesses at its locus' row of the warp's window (pi*theta gather, accumulator load, accumulator store).  The
hardware serves an 8-byte access per half-warp; a half-warp's 16 addresses cost as many wavefronts as the fullest of the 16
bank pairs holds distinct words (2 wavefronts per instruction = conflict-free).  ncu on B200 shows 36 % of the kernel's
shared-memory wavefronts to be conflicts and the L1 data pipe at 90 % (profiles/ncu_details_r2d_k_ell_long.csv).

Two re-layouts that cost nothing at run time were candidates: XOR-swizzling the window row (row ^ ((row >> 4) & 15)) and
dealing the 32 entries of a chunk to the lanes by bank pair (ranked by row & 15, alternately to the two half-warps).  This
script counts the wavefronts each would give.  Result (3000 long reads of the 50 M x 30 k Zipf matrix, seed 1004):

    plain 3.81   swizzle 3.99   deal 3.71   swizzle + deal 3.75   wavefronts per 8-byte access

-- the conflicts come from 32 scattered rows meeting 16 bank pairs, not from an unlucky layout; neither was built.

    python tools/analysis/long_read_bank_conflicts.py [--reads 300000] [--sample 3000]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from telescope_b200.synthetic import synth_csr          # noqa: E402

LONG_WIN = 512          # kLongWin
LONG_MIN = 48           # reads with more entries are long (2 * kEllTMax)


def wavefronts(rows):
    """One half-warp phase: distinct words per bank pair, the fullest pair counts."""
    if rows.size == 0:
        return 0
    return int(np.bincount(np.unique(rows) & 15, minlength=16).max())


def deal(rows):
    order = np.argsort(rows & 15, kind="stable")
    rank = np.empty(rows.size, dtype=np.int64)
    rank[order] = np.arange(rows.size)
    return rows[(rank & 1) == 0], rows[(rank & 1) == 1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=300_000)
    ap.add_argument("--sample", type=int, default=3000)
    a = ap.parse_args()
    ip, ix, _ = synth_csr(a.reads, 30_000, 20, True, 1004)
    longs = np.flatnonzero(np.diff(ip) > LONG_MIN)[:a.sample]
    tot = dict.fromkeys(("plain", "swizzle", "deal", "swizzle + deal"), 0)
    chunks = 0
    for r in longs:
        rows = ix[ip[r]:ip[r + 1]] & (LONG_WIN - 1)
        for c0 in range(0, rows.size, 32):
            ch = rows[c0:c0 + 32]
            sw = ch ^ ((ch >> 4) & 15)
            chunks += 1
            tot["plain"] += wavefronts(ch[:16]) + wavefronts(ch[16:])
            tot["swizzle"] += wavefronts(sw[:16]) + wavefronts(sw[16:])
            tot["deal"] += sum(wavefronts(h) for h in deal(ch))
            tot["swizzle + deal"] += sum(wavefronts(h) for h in deal(sw))
    print("%d long reads, %d chunks of 32 entries" % (longs.size, chunks))
    for k, v in tot.items():
        print("%-16s %.2f wavefronts per 8-byte access (2 = conflict-free)" % (k, v / chunks))


if __name__ == "__main__":
    main()
