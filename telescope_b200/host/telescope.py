# -*- coding: utf-8 -*-
"""Host-side container around the EM path: alignment -> score matrix, checkpoint, reports.

Mirrors the reference's `Telescope` class (telescope/utils/model.py:74-564) where it is part of the drop-in surface:
  * `load_alignment`  -- fragments -> (read, locus, score, length) mappings -> uint16 CSR score matrix
                         (model.py:214-285 sequential loader, 30-63 best alignment per locus, 287-362 matrix)
  * `save` / `load`   -- the NPZ checkpoint, same keys and dtypes (model.py:108-148)
  * `get_random_seed` -- model.py:150-153
  * `output_report`   -- the two TSVs, same columns / rounding / ordering / quirks (model.py:420-477)
  * `update_sam`      -- model.py:479-521 (with the tagging of model.py:30-63 during loading), BAM input only
  * `print_summary`   -- model.py:523-555
BAM/GTF parsing and overlap stay on the host (BASELINE.json north_star); only the EM loop and reassignment run on
the GPU, through `telescope_b200.likelihood.TelescopeLikelihood`.

Differences from the reference, on purpose:
  * the score matrix is assembled with one vectorised sort/dedupe instead of a Python dict-of-keys matrix;
  * `read_index` is rebuilt after reads without an annotated hit are dropped, so `save` -> `load` round-trips in
    that case too (the reference keeps the stale index and its own `load` then fails its shape assertion).
"""
import logging as lg
from collections import Counter, OrderedDict

import numpy as np
import pandas as pd

from ..sparse_plus import csr_matrix_plus as csr_matrix
from . import bam

BIG_INT = 2 ** 32 - 1

# fragment classes, same codes as the reference (alignment.py:22-29)
CODES = [('SU', 'single_unmapped'), ('SM', 'single_mapped'), ('PU', 'pair_unmapped'), ('PM', 'pair_mapped'),
         ('PX', 'pair_mixed'), ('PX*', 'pair_mixed_unmapped')]


def str2int(s):
    for cast in (int, float):
        try:
            return cast(s)
        except ValueError:
            pass
    return s


def merge_blocks(ivs, dist=0):
    """Union of intervals, joining those at most `dist` apart (reference helpers.py:74-104)."""
    if len(ivs) <= 1:
        return list(ivs)
    ivs = sorted(ivs, key=lambda x: x[0])
    out = [ivs[0]]
    for iv in ivs[1:]:
        if iv[0] - out[-1][1] > dist:
            out.append(iv)
        else:
            out[-1] = (out[-1][0], max(iv[1], out[-1][1]))
    return out


class Fragment(object):
    """One alignment of a fragment: a mate pair or a single segment (reference calignment.pyx:16-98)."""
    __slots__ = ("r1", "r2", "_blocks")

    def __init__(self, r1, r2=None):
        self.r1, self.r2, self._blocks = r1, r2, None

    query_id = property(lambda self: self.r1.name)
    is_paired = property(lambda self: self.r2 is not None)
    is_unmapped = property(lambda self: self.r1.is_unmapped)
    r1_is_reversed = property(lambda self: self.r1.is_reverse)

    @property
    def refblocks(self):
        if self._blocks is None:
            b = list(self.r1.blocks) + (list(self.r2.blocks) if self.r2 is not None else [])
            self._blocks = merge_blocks(b, 1)
        return self._blocks

    @property
    def alnlen(self):
        return sum(b[1] - b[0] for b in self.refblocks)

    @property
    def alnscore(self):
        return self.r1.score + (self.r2.score if self.r2 is not None else 0)

    # ---- record editing, only used with --updated_sam (segments then carry their raw BAM records)
    def _editors(self):
        if getattr(self.r1, "_ed", None) is None:
            self.r1._ed = bam.RecordEditor(self.r1)
            if self.r2 is not None:
                self.r2._ed = bam.RecordEditor(self.r2)
        return [self.r1._ed] + ([self.r2._ed] if self.r2 is not None else [])

    def set_tag(self, tag, value):
        for ed in self._editors():
            ed.set_tag(tag, value)

    def set_mapq(self, q):
        for ed in self._editors():
            ed.set_mapq(q)

    def set_flag(self, bits):
        for ed in self._editors():
            ed.set_flag(bits)

    def unset_flag(self, bits):
        for ed in self._editors():
            ed.unset_flag(bits)

    def write(self, out):
        for ed in self._editors():
            out.write(ed)


def _key(a):
    return (a.name, a.is_read1, a.ref_id, a.pos, a.next_ref_id, a.next_pos, abs(a.tlen))


def _mate_key(a):
    return (a.name, not a.is_read1, a.next_ref_id, a.next_pos, a.ref_id, a.pos, abs(a.tlen))


def pair_up(alns):
    """Pair the segments of one bundle by mate coordinates (reference alignment.py:128-145)."""
    waiting, out = {}, []
    for a in alns:
        if not a.is_paired:
            out.append(Fragment(a))
            continue
        mate = waiting.pop(_mate_key(a), None)
        if mate is None:
            waiting[_key(a)] = a
        else:
            out.append(Fragment(a, mate) if a.is_read1 else Fragment(mate, a))
    out.extend(Fragment(a) for a in waiting.values())
    return out


def classify(alns):
    """(code index, fragments) for a bundle (reference alignment.py:148-161)."""
    first = alns[0]
    if not first.is_paired:
        return (0 if first.is_unmapped else 1), [Fragment(a) for a in alns]
    if first.is_proper_pair:
        return 3, pair_up(alns)
    if len(alns) == 2 and all(a.is_unmapped for a in alns):
        return 2, [Fragment(alns[0], alns[1])]
    return 4, [Fragment(a) for a in alns]


class Telescope(object):

    def __init__(self, opts):
        self.opts = opts
        self.single_cell = False
        self.run_info = OrderedDict()
        self.feature_length = None
        self.read_index = {}
        self.feat_index = {}
        self.shape = None
        self.raw_scores = None
        self.run_info['version'] = getattr(opts, 'version', 'unknown')
        # BAM with non overlapping fragments (or unmapped) / with overlapping fragments (model.py:89-92)
        self.other_bam = opts.outfile_path('other.bam') if hasattr(opts, 'outfile_path') else None
        self.tmp_bam = opts.outfile_path('tmp_tele.bam') if hasattr(opts, 'outfile_path') else None
        with bam.AlignmentReader(opts.samfile) as sf:
            self.ref_names, self.ref_lengths = list(sf.references), list(sf.lengths)
        self.has_index = False

    # ------------------------------------------------------------------ checkpoint (model.py:108-148)
    def save(self, filename):
        _feat_list = sorted(self.feat_index, key=self.feat_index.get)
        _flen_list = [self.feature_length[f] for f in _feat_list]
        np.savez(filename,
                 _run_info=list(self.run_info.items()),
                 _flen_list=_flen_list,
                 _feat_list=_feat_list,
                 _read_list=sorted(self.read_index, key=self.read_index.get),
                 _shape=self.shape,
                 _raw_scores_data=self.raw_scores.data,
                 _raw_scores_indices=self.raw_scores.indices,
                 _raw_scores_indptr=self.raw_scores.indptr,
                 _raw_scores_shape=self.raw_scores.shape)

    @classmethod
    def load(cls, filename):
        loader = np.load(filename)
        obj = cls.__new__(cls)
        obj.single_cell = False
        obj.run_info = OrderedDict()
        for k, v in loader['_run_info']:
            obj.run_info[str(k)] = str2int(str(v))
        obj.feature_length = Counter()
        for f, fl in zip(loader['_feat_list'], loader['_flen_list']):
            obj.feature_length[str(f)] = fl
        obj.read_index = {str(n): i for i, n in enumerate(loader['_read_list'])}
        obj.feat_index = {str(n): i for i, n in enumerate(loader['_feat_list'])}
        obj.shape = len(obj.read_index), len(obj.feat_index)
        assert tuple(loader['_shape']) == obj.shape
        obj.raw_scores = csr_matrix((loader['_raw_scores_data'], loader['_raw_scores_indices'],
                                     loader['_raw_scores_indptr']), shape=loader['_raw_scores_shape'])
        return obj

    def get_random_seed(self):
        ret = self.run_info['total_fragments'] % self.shape[0] * self.shape[1]
        return ret % 4294967295

    # ------------------------------------------------------------------ alignment -> matrix
    def _assign_func(self, annotation):
        nf, thresh = self.opts.no_feature_key, self.opts.overlap_threshold
        mode = str(self.opts.stranded_mode)
        if self.opts.overlap_mode != 'threshold':
            raise NotImplementedError('overlap_mode "%s" is a stub in the reference too' % self.opts.overlap_mode)

        def assign(frag):
            # fragment strand by library type (reference model.py:877-889)
            if frag.r1_is_reversed:
                strand = ('+' if mode[-1] == 'F' else '-') if frag.is_paired else ('-' if mode[0] == 'F' else '+')
            else:
                strand = ('-' if mode[-1] == 'F' else '+') if frag.is_paired else ('+' if mode[0] == 'F' else '-')
            hits = annotation.intersect_blocks(self.ref_names[frag.r1.ref_id], frag.refblocks, strand)
            if not hits:
                return nf
            name, overlap = hits.most_common()[0]
            return name if overlap > frag.alnlen * thresh else nf
        return assign

    def load_alignment(self, annotation):
        self.run_info['annotated_features'] = len(annotation.loci)
        self.feature_length = annotation.feature_length().copy()
        if getattr(self.opts, 'ncpu', 1) > 1:
            lg.warning('--ncpu > 1 is not supported (it fails in the reference as well); loading sequentially')
        nf = self.opts.no_feature_key
        assign = self._assign_func(annotation)
        info = Counter()
        reads, feats, scores, lens = [], [], [], []
        min_as, max_as = BIG_INT, -BIG_INT
        update_sam = bool(getattr(self.opts, 'updated_sam', False))
        with bam.AlignmentReader(self.opts.samfile, keep_raw=update_sam) as sf:
            if update_sam and not sf.is_bam:
                raise NotImplementedError('--updated_sam needs BAM input')
            bam_u = bam.BamWriter(self.other_bam, sf.header_text, sf.references, sf.lengths) if update_sam else None
            bam_t = bam.BamWriter(self.tmp_bam, sf.header_text, sf.references, sf.lengths) if update_sam else None
            for alns in bam.bundles(sf):
                info['total_fragments'] += 1
                if info['total_fragments'] % 500000 == 0:
                    msg = '...processed {:.1f}M fragments'.format(info['total_fragments'] / 1e6)
                    lg.info(msg) if info['total_fragments'] % 2500000 == 0 else lg.debug(msg)
                ci, frags = classify(alns)
                code = CODES[ci][0]
                info[code] += 1
                if code in ('SU', 'PU'):
                    if update_sam:
                        frags[0].write(bam_u)
                    continue
                mapped = [f for f in frags if not f.is_unmapped]
                ambig = len(mapped) > 1
                sc = [f.alnscore for f in mapped]
                min_as, max_as = min(min_as, *sc), max(max_as, *sc)
                hit = [assign(f) for f in mapped]
                if all(h == nf for h in hit):
                    info['nofeat_A' if ambig else 'nofeat_U'] += 1
                    if update_sam:
                        for f in frags:
                            f.write(bam_u)
                    continue
                info['feat_A' if ambig else 'feat_U'] += 1
                # best alignment per locus: highest score+length, first one wins ties (model.py:30-47)
                best = OrderedDict()
                for f, h, s in zip(mapped, hit, sc):
                    k = s + f.alnlen
                    if h not in best or k > best[h][0]:
                        best[h] = (k, s, f.alnlen, f)
                ranked = sorted(best.items(), key=lambda kv: kv[1][1], reverse=True)
                for h, (_, s, ln, _f) in ranked:
                    reads.append(alns[0].name); feats.append(h); scores.append(s); lens.append(ln)
                if update_sam:
                    # ZF = locus, ZT = PRI for the locus's best alignment / SEC for the others, ZB = best loci (model.py:48-61)
                    top = ','.join(h for h, v in ranked if v[1] == ranked[0][1][1])
                    for f, h in zip(mapped, hit):
                        f.set_tag('ZF', h)
                        f.set_tag('ZT', 'PRI' if f is best[h][3] else 'SEC')
                    for f in mapped:
                        f.set_tag('ZB', top)
                    for f in frags:
                        f.write(bam_t)
            if update_sam:
                bam_u.close()
                bam_t.close()
        self._mapping_to_matrix(reads, feats, scores, lens, (min_as, max_as), info)
        for f in ('total_fragments', 'pair_mapped', 'pair_mixed', 'single_mapped', 'unmapped', 'unique', 'ambig',
                  'overlap_unique', 'overlap_ambig'):
            self.run_info[f] = info[f]

    def _mapping_to_matrix(self, reads, feats, scores, lens, scorerange, info):
        min_as, max_as = scorerange
        lg.debug('min alignment score: {}'.format(min_as))
        lg.debug('max alignment score: {}'.format(max_as))
        ridx, fidx = {}, {self.opts.no_feature_key: 0}
        rows = np.fromiter((ridx.setdefault(r, len(ridx)) for r in reads), dtype=np.int64, count=len(reads))
        cols = np.fromiter((fidx.setdefault(f, len(fidx)) for f in feats), dtype=np.int64, count=len(feats))
        # rescaled score + aligned length, stored as uint16 like the reference's dok matrix (model.py:294-308)
        vals = (np.asarray(scores, dtype=np.int64) - min_as + 1 + np.asarray(lens, dtype=np.int64)).astype(np.uint16)
        nr, nc = len(ridx), len(fidx)
        # duplicate (read, locus) pairs keep the maximum: sort by (row, col, value), take the last of each run
        order = np.lexsort((vals, cols, rows))
        rows, cols, vals = rows[order], cols[order], vals[order]
        last = np.r_[(rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1]), True] if rows.size else np.zeros(0, bool)
        rows, cols, vals = rows[last], cols[last], vals[last]
        nz = vals != 0
        rows, cols, vals = rows[nz], cols[nz], vals[nz]
        info['unmapped'] = info['SU'] + info['PU']
        info['unique'] = info['nofeat_U'] + info['feat_U']
        info['ambig'] = info['nofeat_A'] + info['feat_A']
        for cs, desc in CODES:
            if cs in info:
                info[desc] = info[cs]
                del info[cs]
        # drop reads whose only hit is the no-feature column (model.py:350-357)
        has_feat = np.zeros(nr, dtype=bool)
        has_feat[rows[cols > 0]] = True
        keep = has_feat[rows]
        newrow = np.cumsum(has_feat) - 1
        rows, cols, vals = newrow[rows[keep]], cols[keep], vals[keep]
        n_keep = int(has_feat.sum())
        indptr = np.zeros(n_keep + 1, dtype=np.int32)
        np.cumsum(np.bincount(rows, minlength=n_keep), out=indptr[1:])
        self.raw_scores = csr_matrix((vals, cols.astype(np.int32), indptr), shape=(n_keep, nc))
        names = np.array(sorted(ridx, key=ridx.get), dtype=object)[has_feat] if nr else np.array([], dtype=object)
        self.read_index = {str(n): i for i, n in enumerate(names)}
        self.feat_index = fidx
        self.shape = (n_keep, nc)
        info['overlap_unique'] = int(np.sum(self.raw_scores.count(1) == 1))
        info['overlap_ambig'] = self.shape[0] - info['overlap_unique']

    # ------------------------------------------------------------------ reports (model.py:420-477)
    def output_report(self, tl, stats_filename, counts_filename):
        _rmethod, _rprob = self.opts.reassign_mode, self.opts.conf_prob
        _fnames = sorted(self.feat_index, key=self.feat_index.get)
        _flens = self.feature_length
        if hasattr(tl, 'report_colsums'):
            # CUDA model: one pass over the matrix for all columns (RNG draws in the reference's order)
            cs = tl.report_colsums(_rprob, _rmethod)
        else:
            # any object with the reference's interface: same call order as model.py:435-441,457
            colsum = getattr(tl, 'reassign_colsum', None) or (lambda *a, **k: tl.reassign(*a, **k).sum(0).A1)
            cs = OrderedDict()
            cs['final_conf'] = colsum('conf', _rprob)
            cs['init_aligned'] = colsum('all', initial=True)
            cs['unique_count'] = colsum('unique')
            cs['init_best'] = colsum('exclude', initial=True)
            cs['init_best_random'] = colsum('choose', initial=True)
            cs['init_best_avg'] = colsum('average', initial=True)
            cs['final'] = colsum(_rmethod, _rprob)
        stats = pd.DataFrame(OrderedDict([
            ('transcript', _fnames),
            ('transcript_length', [_flens[f] for f in _fnames]),
            ('final_conf', cs['final_conf']),
            ('final_prop', tl.pi),
            ('init_aligned', cs['init_aligned']),
            ('unique_count', cs['unique_count']),
            ('init_best', cs['init_best']),
            ('init_best_random', cs['init_best_random']),
            ('init_best_avg', cs['init_best_avg']),
            ('init_prop', tl.pi_init),
        ]))
        stats.sort_values('final_prop', ascending=False, inplace=True)
        stats = stats.round(pd.Series([2, 3, 2, 3], index=['final_conf', 'final_prop', 'init_best_avg', 'init_prop']))
        counts = pd.DataFrame(OrderedDict([('transcript', _fnames), ('count', cs['final'])]))
        counts.sort_values('transcript', inplace=True)
        comment = ["## RunInfo"] + ['{}:{}'.format(*tup) for tup in self.run_info.items()]
        with open(stats_filename, 'w') as outh:
            outh.write('\t'.join(comment))          # no newline: the reference glues the header to it (model.py:471)
            stats.to_csv(outh, sep='\t', index=False)
        with open(counts_filename, 'w') as outh:
            counts.to_csv(outh, sep='\t', index=False)

    def update_sam(self, tl, filename):
        """Re-tag the overlapping fragments with their posterior and assignment (model.py:479-521): XP = posterior in
        percent, mapping quality = phred(posterior), YC = display colour, secondary flag for everything but the
        alignment a fragment is assigned to."""
        import sys
        _rmethod, _rprob = self.opts.reassign_mode, self.opts.conf_prob
        mat = csr_matrix(tl.reassign(_rmethod, _rprob))
        z = csr_matrix(tl.z)
        mat.sort_indices()
        z.sort_indices()

        def row_lookup(m, r):
            """{locus: value} of one read (one CSR slice per fragment instead of a scipy scalar lookup per alignment)"""
            a, b = m.indptr[r], m.indptr[r + 1]
            return dict(zip(m.indices[a:b].tolist(), m.data[a:b].tolist()))
        vermilion, yellow, pale, grey = '217,95,2', '230,171,2', '209,236,228', '248,248,248'

        def phred(p):
            return int(round(-10 * np.log10(1 - p))) if p < 1.0 else 255

        with bam.AlignmentReader(self.tmp_bam, keep_raw=True) as sf:
            pg = '@PG\tID:telescope\tPN:telescope\tCL:%s\tVN:%s\n' % (' '.join(sys.argv), self.run_info['version'])
            with bam.BamWriter(filename, sf.header_text + pg, sf.references, sf.lengths) as out:
                for alns in bam.bundles(sf):
                    _, frags = classify(alns)
                    if not frags:
                        continue
                    ridx = self.read_index[frags[0].query_id]
                    zrow, mrow = row_lookup(z, ridx), row_lookup(mat, ridx)
                    for f in frags:
                        if f.is_unmapped:
                            f.write(out)
                            continue
                        tags = f.r1.tags
                        assert b'ZT' in tags, 'Missing ZT tag'
                        if tags[b'ZT'] == 'SEC':
                            f.set_flag(bam.FSECONDARY)
                            f.set_tag('YC', grey)
                            f.set_mapq(0)
                        else:
                            fidx = self.feat_index[tags[b'ZF']]
                            prob = float(zrow.get(fidx, 0.0))
                            f.set_mapq(phred(prob))
                            f.set_tag('XP', int(round(prob * 100)))
                            if mrow.get(fidx, 0) > 0:
                                f.unset_flag(bam.FSECONDARY)
                                f.set_tag('YC', vermilion)
                            else:
                                f.set_flag(bam.FSECONDARY)
                                f.set_tag('YC', yellow if prob >= 0.2 else pale)
                        f.write(out)

    # ------------------------------------------------------------------ model.py:523-555
    def print_summary(self, loglev=lg.WARNING):
        _d = Counter()
        for k, v in self.run_info.items():
            try:
                _d[k] = int(v)
            except ValueError:
                pass
        if 'mapped_pairs' in _d:
            _d['pair_mapped'] = _d['mapped_pairs']
        if 'mapped_single' in _d:
            _d['single_mapped'] = _d['mapped_single']
        lines = [
            "Alignment Summary:",
            '    {} total fragments.'.format(_d['total_fragments']),
            '        {} mapped as pairs.'.format(_d['pair_mapped']),
            '        {} mapped as mixed.'.format(_d['pair_mixed']),
            '        {} mapped single.'.format(_d['single_mapped']),
            '        {} failed to map.'.format(_d['unmapped']),
            '--',
            '    {} fragments mapped to reference; of these'.format(_d['pair_mapped'] + _d['pair_mixed'] + _d['single_mapped']),
            '        {} had one unique alignment.'.format(_d['unique']),
            '        {} had multiple alignments.'.format(_d['ambig']),
            '--',
            '    {} fragments overlapped annotation; of these'.format(_d['overlap_unique'] + _d['overlap_ambig']),
            '        {} map to one locus.'.format(_d['overlap_unique']),
            '        {} map to multiple loci.'.format(_d['overlap_ambig']),
            '\n',
        ]
        for ln in lines:
            lg.log(loglev, ln)

    def __str__(self):
        if hasattr(self.opts, 'samfile'):
            return '<Telescope samfile=%s, gtffile=%s>' % (self.opts.samfile, self.opts.gtffile)
        if hasattr(self.opts, 'checkpoint'):
            return '<Telescope checkpoint=%s>' % self.opts.checkpoint
        return '<Telescope>'
