#!/usr/bin/env python
"""CPU emulation behind DESIGN.md section 8 item 1: what clustering reads by their first locus does to a tile's
locus range on the benchmark matrix, and how often a per-warp 256-locus shared-memory window would have to move.

    python tools/analysis/cluster_window_emulation.py [n_reads_sampled]

Round-1 output (first 4 M reads of the 50 M x 30 k benchmark matrix, first 20 000 tiles):
  original order          : 116 entries/tile, locus span/tile median 21925, distinct loci/tile 115.6, RED sectors/tile 62.2
                            (ncu on the committed kernel: ~63 per tile -- the emulation matches the hardware count)
  clustered by first locus: 116 entries/tile, locus span/tile median 60 (p90 72), never > 256, distinct loci/tile 52.8,
                            RED sectors/tile 49.9; window W=256 moves 6 times in 20 000 tiles, 0 fallback tiles
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from telescope_b200.synthetic import synth_csr  # noqa: E402


def greedy_tiles(ip):
    starts, r, n = [0], 0, len(ip) - 1
    while r < n:
        r2 = int(np.searchsorted(ip, ip[r] + 128, side="right")) - 1
        r = max(r2, r + 1)
        starts.append(r)
    return np.array(starts)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    K = 30000
    ip, ix, _ = synth_csr(50_000_000, K, 20, False, 1004, 0, n)
    lens = np.diff(ip)
    order = np.argsort(ix[ip[:-1]], kind="stable")
    lens_c = lens[order]
    ip_c = np.zeros(n + 1, np.int64)
    np.cumsum(lens_c, out=ip_c[1:])
    src = np.repeat(ip[:-1][order], lens_c) + (np.arange(ip_c[-1]) - np.repeat(ip_c[:-1], lens_c))
    ix_c = ix[src]
    for name, ipx, ixx in (("original order", ip, ix), ("clustered by first locus", ip_c, ix_c)):
        ipx = ipx[:400_001]
        ts = greedy_tiles(ipx)[:20001]
        spans, sect, distinct, lo_hi = [], [], [], []
        for a, b in zip(ts[:-1], ts[1:]):
            c = ixx[ipx[a]:ipx[b]]
            spans.append(c.max() - c.min() + 1)
            distinct.append(len(np.unique(c)))
            sect.append(sum(len(np.unique(c[k:k + 32] >> 2)) for k in range(0, len(c), 32)))
            lo_hi.append((c.min(), c.max()))
        spans = np.array(spans)
        print("%s: %d tiles, locus span/tile median %.0f p90 %.0f, >256: %.3f, distinct loci/tile %.1f, RED sectors/tile %.1f"
              % (name, len(spans), np.median(spans), np.percentile(spans, 90), np.mean(spans > 256), np.mean(distinct), np.mean(sect)))
        if name.startswith("clustered"):
            W, moves, wb, fallback = 256, 0, -10 ** 9, 0
            for lo, hi in lo_hi:
                if hi - lo + 1 > W:
                    fallback += 1
                elif lo < wb or hi >= wb + W:
                    moves, wb = moves + 1, (lo // 32) * 32
            print("   window W=%d: %d moves over %d tiles, %d fallback tiles" % (W, moves, len(lo_hi), fallback))


if __name__ == "__main__":
    main()
