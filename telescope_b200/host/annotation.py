# -*- coding: utf-8 -*-
"""GTF annotation -> per-chromosome locus intervals with overlap queries (host side; numpy only).

Same semantics as the reference's intervaltree-backed class (telescope/utils/_annotation_intervaltree.py:29-102):
`exon` rows carrying the locus attribute become half-open intervals [start, end+1); overlapping intervals of the
same locus are merged; `intersect_blocks` sums, per locus, the overlap of each aligned block [b_start, b_end+1)
with the intervals (optionally only those on the fragment's strand).  Implemented as sorted arrays + bisection
instead of an interval tree (intervaltree is not in this image).
"""
import bisect
import re
from collections import Counter, OrderedDict, defaultdict

_ATTR = re.compile(r'(\w+)\s+"(.+?)";')


class Annotation(object):

    def __init__(self, gtf_file, attribute_name="locus", stranded_mode="None", feature_type="exon"):
        self.key = attribute_name
        self.loci = OrderedDict()                 # locus -> its GTF rows, first-seen order
        self.run_stranded = stranded_mode != "None"
        per_chrom = defaultdict(list)             # chrom -> [begin, end, locus, strand]
        fh = open(gtf_file) if isinstance(gtf_file, str) else gtf_file
        for line in fh:
            if line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 9 or f[2] != feature_type:
                continue
            attr = dict(_ATTR.findall(f[8]))
            if self.key not in attr:
                continue
            locus = attr[self.key]
            self.loci.setdefault(locus, []).append(f)
            b, e = int(f[3]), int(f[4]) + 1
            ivs = per_chrom[f[0]]
            keep = []
            for iv in ivs:                         # merge with every overlapping interval of the same locus
                if iv[2] == locus and iv[0] < e and b < iv[1]:
                    b, e = min(b, iv[0]), max(e, iv[1])
                else:
                    keep.append(iv)
            keep.append([b, e, locus, f[6]])
            per_chrom[f[0]] = keep
        if isinstance(gtf_file, str):
            fh.close()
        self._index = {}
        for chrom, ivs in per_chrom.items():
            ivs.sort(key=lambda iv: iv[0])
            begins = [iv[0] for iv in ivs]
            reach, m = [], 0                       # running maximum of interval ends: bounds the backward scan
            for iv in ivs:
                m = max(m, iv[1])
                reach.append(m)
            self._index[chrom] = (begins, reach, ivs)

    def feature_length(self):
        out = Counter()
        for _, _, ivs in self._index.values():
            for b, e, locus, _ in ivs:
                out[locus] += e - b
        return out

    def intersect_blocks(self, ref, blocks, frag_strand=None):
        out = Counter()
        idx = self._index.get(ref)
        if idx is None:
            return out
        begins, reach, ivs = idx
        for b_start, b_end in blocks:
            qb, qe = b_start, b_end + 1
            i = bisect.bisect_left(begins, qe) - 1            # last interval starting before the query's end
            while i >= 0 and reach[i] > qb:
                b, e, locus, strand = ivs[i]
                if e > qb and (not self.run_stranded or strand == frag_strand):
                    out[locus] += max(0, min(e, qe) - max(b, qb))
                i -= 1
        return out
