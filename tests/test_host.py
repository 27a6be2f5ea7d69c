"""CPU tier: host logic -- the C ABI surface, csr_matrix_plus, the BAM/GTF loader, checkpoint format, report
writers, CLI parsing, synthetic generator.  No compute calls into the CUDA library happen here."""
import ctypes
import io
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, Opts
from oracle.em_numpy import EMOracle

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "telescope_b200", "data")


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    from telescope_b200 import _abi
    header = open(os.path.join(ROOT, "include", "telescope_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tsc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    assert _abi.load().tsc_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    from telescope_b200 import _abi
    from telescope_b200.likelihood import TelescopeLikelihood
    if _abi.device_count() > 0:
        pytest.skip("a GPU is present")
    m = sp.csr_matrix(np.array([[200, 190, 0], [0, 180, 170]], dtype=np.uint16))
    with pytest.raises(_abi.TelescopeCudaError):
        TelescopeLikelihood(m, Opts())


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "telescope_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), f


# ------------------------------------------------------------------------------------------------ csr_matrix_plus
def test_csr_matrix_plus_known_answers():
    """reference tests/test_sparse_plus.py:24-66 and the docstrings of sparse_plus.py:33-41,76-85,106-115."""
    from telescope_b200.sparse_plus import csr_matrix_plus as C
    m = C(np.array([[1, 0, 2], [0, 0, 3], [4, 5, 6]]))
    assert np.allclose(m.norm().toarray(), np.array([[1, 0, 2], [0, 0, 3], [4, 5, 6]]) / 21.0)
    assert np.allclose(m.norm(1).toarray(), [[1 / 3., 0, 2 / 3.], [0, 0, 1], [4 / 15., 5 / 15., 6 / 15.]])
    z = C(np.array([[1, 0, 2], [0, 0, 0], [4, 5, 6]]))
    assert np.allclose(z.norm(1).toarray(), [[1 / 3., 0, 2 / 3.], [0, 0, 0], [4 / 15., 5 / 15., 6 / 15.]])   # zero row stays zero
    s = C([[10, 0, 20], [0, 0, 30], [40, 50, 60]])
    assert np.allclose(s.scale().toarray(), [[1 / 6., 0, 2 / 6.], [0, 0, 0.5], [4 / 6., 5 / 6., 1]])
    assert np.allclose(s.scale(1).toarray(), [[0.5, 0, 1], [0, 0, 1], [4 / 6., 5 / 6., 1]])
    b = C([[6, 0, 2], [0, 0, 3], [4, 5, 6]])
    assert np.array_equal(b.binmax(1).toarray(), [[1, 0, 0], [0, 0, 1], [0, 0, 1]])
    assert np.array_equal(b.count(1).ravel(), [2, 1, 3])
    assert isinstance(m.norm(1), C)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        m.save(os.path.join(d, "m"))
        assert m.check_equal(C.load(os.path.join(d, "m.npz")))


@pytest.mark.parametrize("n,lo,hi", [(5000, 2, 5), (70001, 1, 40), (8192, 2, 3), (20000, 2, 2 ** 31 - 1), (4096, 3, 4),
                                     (300, 2, 9)])
def test_native_tie_draws_walk_numpys_generator(n, lo, hi):
    """tsc_mt19937_draw_picks / _rows (host code of the library, used for long arrays) give the draws of
    np.random.randint(0, counts) -- i.e. of the reference's np.random.choice(range(a, b)) per tie read,
    sparse_plus.py:146-153 -- and leave the global generator in the same state (next uniform and normal draws agree)."""
    from telescope_b200.sparse_plus import draw_picks, draw_row_picks
    counts = np.random.default_rng(n).integers(lo, hi, size=n)
    np.random.seed(1000 + n)
    want = np.random.randint(0, counts)
    after_want = (np.random.random(4), np.random.standard_normal(3), np.random.randint(0, 1000, 5))
    np.random.seed(1000 + n)
    got = draw_picks(counts)
    after_got = (np.random.random(4), np.random.standard_normal(3), np.random.randint(0, 1000, 5))
    assert np.array_equal(got, want)
    assert all(np.array_equal(a, b) for a, b in zip(after_got, after_want))
    if hi < 2 ** 20:
        # per-read form: reads with 0 or 1 best hits in between consume nothing and pick 0
        nbest = np.zeros(3 * n, dtype=np.int32)
        where = np.sort(np.random.default_rng(n + 1).choice(3 * n, size=n, replace=False))
        nbest[where] = counts
        nbest[(where[::7] + 1) % (3 * n)] = np.where(nbest[(where[::7] + 1) % (3 * n)] == 0, 1, nbest[(where[::7] + 1) % (3 * n)])
        ties = np.flatnonzero(nbest > 1)
        np.random.seed(77)
        want_rows = np.zeros(3 * n, dtype=np.int64)
        want_rows[ties] = np.random.randint(0, nbest[ties].astype(np.int64))
        nxt_want = np.random.random(2)
        np.random.seed(77)
        got_rows = draw_row_picks(nbest)
        nxt_got = np.random.random(2)
        assert got_rows.dtype == np.int32 and np.array_equal(got_rows, want_rows) and np.array_equal(nxt_got, nxt_want)


def test_native_tie_draws_after_a_partial_block_and_across_regenerations():
    """The generator position is honoured: draws that start in the middle of a 624-word block and run over several
    regenerations continue numpy's sequence."""
    from telescope_b200.sparse_plus import draw_picks
    counts = np.full(10000, 3)                       # 25 % rejections: ~13 300 words, 21 regenerations
    for burn in (0, 1, 311, 623, 624, 625):
        np.random.seed(9)
        np.random.randint(0, 2 ** 31 - 1, burn)      # one word each
        want = np.random.randint(0, counts)
        tail_want = np.random.randint(0, 10 ** 6, 3)
        np.random.seed(9)
        np.random.randint(0, 2 ** 31 - 1, burn)
        got = draw_picks(counts)
        tail_got = np.random.randint(0, 10 ** 6, 3)
        assert np.array_equal(got, want) and np.array_equal(tail_got, tail_want), burn


def test_choose_random_consumes_rng_like_reference_loop():
    from telescope_b200.sparse_plus import csr_matrix_plus as C
    from telescope_b200.sparse_plus import draw_picks
    counts = np.random.default_rng(0).integers(2, 40, size=500)
    np.random.seed(123)
    seq = np.array([np.random.choice(range(5, 5 + n)) - 5 for n in counts])      # sparse_plus.py:149
    np.random.seed(123)
    assert np.array_equal(draw_picks(counts), seq)
    rng = np.random.default_rng(1)
    d = (rng.random((200, 30)) < 0.2).astype(np.int8)
    m = C(d)
    np.random.seed(7)
    mine = m.choose_random(1)
    np.random.seed(7)
    ref = m.copy()
    for a, b in zip(ref.indptr[:-1], ref.indptr[1:]):                                # the reference's loop, verbatim semantics
        if b - a > 1:
            ch = np.random.choice(range(a, b))
            ref.data[a:ch] = 0
            ref.data[ch + 1:b] = 0
    ref.eliminate_zeros()
    assert (mine != ref).nnz == 0


# ------------------------------------------------------------------------------------------------ loader / checkpoint
class AssignOpts(Opts):
    samfile = os.path.join(DATA, "alignment.bam")
    gtffile = os.path.join(DATA, "annotation.gtf")
    no_feature_key, overlap_threshold, overlap_mode, stranded_mode, ncpu = "__no_feature", 0.2, "threshold", "None", 1
    version = "GOLDEN"
    updated_sam = False


@pytest.fixture(scope="module")
def bundled_ts():
    from telescope_b200.host.annotation import Annotation
    from telescope_b200.host.telescope import Telescope
    opts = AssignOpts()
    ts = Telescope(opts)
    ts.load_alignment(Annotation(opts.gtffile, "locus", "None"))
    return ts


def test_loader_reproduces_bundled_matrix(bundled_ts):
    g = np.load(os.path.join(GOLD, "bundled.npz"))
    ts = bundled_ts
    assert ts.shape == (1000, 59) and ts.raw_scores.nnz == 18471
    assert ts.raw_scores.dtype == np.uint16 and ts.raw_scores.indices.dtype == np.int32
    assert np.array_equal(ts.raw_scores.data, g["raw"]) and np.array_equal(ts.raw_scores.indices, g["indices"])
    assert np.array_equal(ts.raw_scores.indptr, g["indptr"])
    # run info of the bundled report (telescope/data/telescope_report.tsv:1)
    want = dict(annotated_features=99, total_fragments=1000, pair_mapped=1000, pair_mixed=0, single_mapped=0, unmapped=0,
                unique=0, ambig=1000, overlap_unique=0, overlap_ambig=1000)
    for k, v in want.items():
        assert ts.run_info[k] == v, k
    assert ts.get_random_seed() == 0
    assert sorted(ts.feat_index, key=ts.feat_index.get)[0] == "__no_feature"


def test_checkpoint_is_the_reference_format(bundled_ts, tmp_path):
    from telescope_b200.host.telescope import Telescope
    p = str(tmp_path / "ckpt")
    bundled_ts.save(p)
    mine, ref = np.load(p + ".npz"), np.load(os.path.join(GOLD, "bundled_checkpoint.npz"))   # written by the reference's save()
    assert sorted(mine.files) == sorted(ref.files)
    for k in ref.files:
        assert mine[k].dtype.kind == ref[k].dtype.kind and mine[k].shape == ref[k].shape, k
        if k != "_run_info":
            assert np.array_equal(mine[k], ref[k]), k
    back = Telescope.load(os.path.join(GOLD, "bundled_checkpoint.npz"))                      # and we read the reference's file
    assert back.shape == (1000, 59) and back.run_info["total_fragments"] == 1000
    assert (back.raw_scores != bundled_ts.raw_scores).nnz == 0
    assert back.get_random_seed() == 0


def test_mapping_to_matrix_keeps_max_and_drops_unannotated_reads():
    from telescope_b200.host.telescope import Telescope
    from collections import Counter
    ts = Telescope.__new__(Telescope)
    ts.opts = AssignOpts()
    ts.read_index, ts.feat_index = {}, {}
    info = Counter()
    reads = ["r1", "r1", "r1", "r2", "r3", "r3"]
    feats = ["A", "B", "A", "__no_feature", "B", "__no_feature"]
    ts._mapping_to_matrix(reads, feats, [10, 12, 11, 9, 10, 10], [100, 100, 100, 100, 90, 90], (9, 12), info)
    # rescale = score - min + 1, value = rescale + length; duplicates keep the max; r2 hits no feature -> dropped
    assert ts.shape == (2, 3) and ts.read_index == {"r1": 0, "r3": 1}
    assert np.array_equal(ts.raw_scores.toarray(), [[0, 103, 104], [92, 0, 92]])
    assert info["overlap_unique"] == 0 and info["overlap_ambig"] == 2


# ------------------------------------------------------------------------------------------------ reports
class OracleModel(object):
    """Stands in for the GPU class in CPU-tier report tests: same reassign_colsum / pi / pi_init surface."""

    def __init__(self, m, opts):
        self.o = EMOracle(m.indptr, m.indices, m.data, m.shape[1], opts.em_epsilon, opts.max_iter, opts.pi_prior, opts.theta_prior).em()
        self.pi, self.pi_init = self.o.pi, self.o.pi_init

    def reassign_colsum(self, method, thresh=0.9, initial=False):
        return self.o.reassign_colsum(method, thresh, initial)


def strip_version(text):
    return re.sub(r"version:[^\t]*", "version:X", text)


def test_output_report_matches_reference_files(bundled_ts, tmp_path):
    ts = bundled_ts
    ts.opts = AssignOpts()
    np.random.seed(ts.get_random_seed())
    ts.output_report(OracleModel(ts.raw_scores, ts.opts), str(tmp_path / "s.tsv"), str(tmp_path / "c.tsv"))
    assert strip_version(open(tmp_path / "s.tsv").read()) == strip_version(open(os.path.join(GOLD, "bundled_run_stats.tsv")).read())
    assert open(tmp_path / "c.tsv").read() == open(os.path.join(GOLD, "bundled_TE_counts.tsv")).read()
    head = open(tmp_path / "s.tsv").readline()
    assert head.startswith("## RunInfo\tversion:") and "overlap_ambig:1000transcript\t" in head   # glued header, model.py:471


# ------------------------------------------------------------------------------------------------ CLI / misc
def test_cli_options_and_defaults():
    from telescope_b200 import cli
    import argparse
    p = argparse.ArgumentParser()
    cli.add_assign_arguments(p)
    a = p.parse_args(["x.bam", "y.gtf"])
    assert (a.pi_prior, a.theta_prior, a.em_epsilon, a.max_iter, a.reassign_mode, a.conf_prob) == (0, 200000, 1e-7, 100, "exclude", 0.9)
    assert a.overlap_threshold == 0.2 and a.no_feature_key == "__no_feature" and a.attribute == "locus" and not a.skip_em
    p = argparse.ArgumentParser()
    cli.add_resume_arguments(p)
    r = p.parse_args(["ckpt.npz", "--max_iter", "5", "--use_likelihood", "--devices", "0,1"])
    o = cli.Options(r)
    assert o.max_iter == 5 and o.use_likelihood and o.device_list() == [0, 1]
    assert o.outfile_path("run_stats.tsv") == os.path.join(".", "telescope-run_stats.tsv")
    buf = io.StringIO()
    import contextlib
    with contextlib.redirect_stdout(buf):
        cli.main(["test"])
    assert buf.getvalue().startswith("telescope assign ") and buf.getvalue().strip().endswith("annotation.gtf")


def test_synthetic_generator_is_canonical_and_shardable():
    from telescope_b200.synthetic import shard_bounds, synth_csr
    N, K = 2_200_000, 4000
    ip, ix, raw = synth_csr(N, K, 10, False, 9)
    lens = np.diff(ip)
    assert 9.5 < lens.mean() < 10.5 and 0.19 < np.mean(lens == 1) < 0.21
    inner = np.ones(ix.size, bool)
    inner[ip[:-1]] = False
    assert (np.diff(ix.astype(np.int64))[inner[1:]] > 0).all(), "strictly increasing loci inside a read"
    assert ix.min() >= 0 and ix.max() < K and raw.min() >= 140 and raw.max() <= 211
    parts = [synth_csr(N, K, 10, False, 9, lo, hi) for lo, hi in shard_bounds(N, 3)]
    assert np.array_equal(np.concatenate([p[1] for p in parts]), ix) and np.array_equal(np.concatenate([p[2] for p in parts]), raw)
    ipz, _, _ = synth_csr(50000, 3000, 20, True, 9)
    assert np.diff(ipz).max() == 200


def test_bam_reader_basics():
    from telescope_b200.host import bam
    with bam.AlignmentReader(os.path.join(DATA, "alignment.bam")) as sf:
        assert len(sf.references) > 0
        n, names, first = 0, set(), None
        for seg in sf:
            first = first or seg
            n += 1
            names.add(seg.name)
    assert n == 66414 and len(names) == 1000
    assert first.is_paired and first.is_proper_pair and first.score is not None and first.blocks


# ------------------------------------------------------------------------------------------------ --updated_sam
def test_updated_sam_retags_fragments_like_the_reference_rules(tmp_path):
    """telescope assign --updated_sam (reference model.py:30-63 tagging, 479-521 update_sam): every alignment of an
    overlapping fragment gets ZF/ZT/ZB while loading; afterwards the best alignment per locus carries XP (posterior in
    percent), mapq = phred(posterior), a colour, and everything but the assigned alignment is flagged secondary."""
    import scipy.sparse as sp
    from telescope_b200.host import bam
    from telescope_b200.host.annotation import Annotation
    from telescope_b200.host.telescope import Telescope

    class O(AssignOpts):
        updated_sam = True
        outdir = str(tmp_path)
        exp_tag = "t"

        def outfile_path(self, suffix):
            return os.path.join(self.outdir, "%s-%s" % (self.exp_tag, suffix))

    opts = O()
    ts = Telescope(opts)
    ts.load_alignment(Annotation(opts.gtffile, "locus", "None"))
    n_in = 66414
    with bam.AlignmentReader(ts.tmp_bam) as t, bam.AlignmentReader(ts.other_bam) as u:
        tmp, other = list(t), list(u)
    assert len(tmp) + len(other) == n_in and len(other) == 0          # every bundled fragment overlaps the annotation
    assert all(b"ZF" in s.tags and s.tags[b"ZT"] in ("PRI", "SEC") and b"ZB" in s.tags for s in tmp)

    class Model(OracleModel):                                            # CPU stand-in with the reference's interface
        def __init__(self, m, o):
            OracleModel.__init__(self, m, o)
            self.m = m
            self.z = sp.csr_matrix((self.o.z, m.indices.copy(), m.indptr.copy()), shape=m.shape)

        def reassign(self, method, thresh=0.9, initial=False):
            d = self.o.reassign_data(method, thresh, initial)
            return sp.csr_matrix((d, self.m.indices.copy(), self.m.indptr.copy()), shape=self.m.shape)

    tl = Model(ts.raw_scores, opts)
    out = os.path.join(str(tmp_path), "t-updated.bam")
    ts.update_sam(tl, out)
    with bam.AlignmentReader(out) as r:
        upd = list(r)
        assert "@PG\tID:telescope\tPN:telescope" in r.header_text
    assert len(upd) == len(tmp)
    mat, z = tl.reassign("exclude", 0.9), tl.z
    n_assigned = 0
    for a, b in zip(tmp, upd):
        assert a.name == b.name and a.pos == b.pos
        if a.tags[b"ZT"] == "SEC":
            assert b.flag & bam.FSECONDARY and b.tags[b"YC"] == "248,248,248"
            continue
        i, j = ts.read_index[a.name], ts.feat_index[a.tags[b"ZF"]]
        p = z[i, j]
        assert b.tags[b"XP"] == int(round(p * 100))
        if mat[i, j] > 0:
            n_assigned += 1
            assert not (b.flag & bam.FSECONDARY) and b.tags[b"YC"] == "217,95,2"
        else:
            assert b.flag & bam.FSECONDARY and b.tags[b"YC"] == ("230,171,2" if p >= 0.2 else "209,236,228")
    assert n_assigned == 2 * int(mat.sum())                              # both mates of each assigned fragment


# ------------------------------------------------------------------------------------------------ host helpers
def test_merge_blocks_docstring_answers():
    """reference helpers.py:86-93"""
    from telescope_b200.host.telescope import merge_blocks
    assert merge_blocks([]) == []
    assert merge_blocks([(1, 10)]) == [(1, 10)]
    assert merge_blocks([(4, 9), (10, 14), (1, 3)]) == [(1, 3), (4, 9), (10, 14)]
    assert merge_blocks([(4, 9), (10, 14), (1, 3)], dist=1) == [(1, 14)]


def _seg(name, flag, ref_id=0, pos=100, nref=0, npos=300, tlen=250, blocks=((100, 150),), score=40):
    from telescope_b200.host import bam
    s = bam.Segment()
    s.name, s.flag, s.ref_id, s.pos, s.next_ref_id, s.next_pos, s.tlen = name, flag, ref_id, pos, nref, npos, tlen
    s.blocks, s.score, s.tags, s._ed = list(blocks), score, {}, None
    return s


def test_fragment_classification_and_pairing():
    """reference alignment.py:128-161: bundle -> (code, fragments); mates pair up by coordinates."""
    from telescope_b200.host.telescope import CODES, classify
    # proper pair with two alignments: mates pair by (pos, next_pos)
    r1a = _seg("q", 99, pos=100, npos=300, tlen=250, score=40)
    r2a = _seg("q", 147, pos=300, npos=100, tlen=-250, blocks=((300, 350),), score=38)
    r1b = _seg("q", 355, pos=900, npos=1100, tlen=250, blocks=((900, 950),), score=30)
    r2b = _seg("q", 403, pos=1100, npos=900, tlen=-250, blocks=((1100, 1150),), score=31)
    code, frags = classify([r1a, r2a, r1b, r2b])
    assert CODES[code][0] == "PM" and len(frags) == 2
    assert frags[0].r1 is r1a and frags[0].r2 is r2a and frags[0].alnscore == 78
    assert frags[0].refblocks == [(100, 150), (300, 350)] and frags[0].alnlen == 100
    assert frags[1].r1 is r1b and frags[1].alnscore == 61
    # single-end mapped / unmapped, unmapped pair, mixed pair
    assert CODES[classify([_seg("s", 0)])[0]][0] == "SM"
    assert CODES[classify([_seg("s", 4)])[0]][0] == "SU"
    code, frags = classify([_seg("p", 77), _seg("p", 141)])
    assert CODES[code][0] == "PU" and len(frags) == 1 and frags[0].r2 is not None
    code, frags = classify([_seg("x", 73), _seg("x", 133)])
    assert CODES[code][0] == "PX" and len(frags) == 2 and frags[0].r2 is None


def test_annotation_overlap_and_strand(tmp_path):
    """reference _annotation_intervaltree.py:29-102: exon rows -> [start, end+1), same-locus overlaps merged, per-locus
    overlap of blocks [b_start, b_end+1), optional strand filter."""
    from telescope_b200.host.annotation import Annotation
    gtf = tmp_path / "a.gtf"
    rows = [
        ("chr1", 100, 200, "+", "L1"), ("chr1", 150, 300, "+", "L1"),      # overlapping exons of one locus merge
        ("chr1", 250, 400, "-", "L2"), ("chr2", 10, 20, "+", "L3"), ("chr1", 1000, 1100, "+", "L1"),
    ]
    gtf.write_text("".join('%s\tsrc\texon\t%d\t%d\t.\t%s\t.\tgene_id "g"; locus "%s";\n' % r for r in rows) +
                   'chr1\tsrc\tgene\t1\t5000\t.\t+\t.\tlocus "IGNORED";\n# comment\n')
    an = Annotation(str(gtf), "locus", "None")
    assert list(an.loci) == ["L1", "L2", "L3"]
    assert an.feature_length() == {"L1": (301 - 100) + (1101 - 1000), "L2": 401 - 250, "L3": 21 - 10}
    hit = an.intersect_blocks("chr1", [(180, 260)])                        # query [180, 261)
    assert hit == {"L1": 261 - 180, "L2": 261 - 250}
    assert an.intersect_blocks("chrX", [(1, 10)]) == {} and an.intersect_blocks("chr1", [(500, 600)]) == {}
    st = Annotation(str(gtf), "locus", "RF")
    assert st.intersect_blocks("chr1", [(180, 260)], "-") == {"L2": 11}
    assert st.intersect_blocks("chr1", [(180, 260)], "+") == {"L1": 81}


def test_sam_text_input_gives_the_same_matrix(tmp_path):
    """The loader accepts SAM text as well as BAM: dump the first fragments of the bundled BAM to SAM and compare."""
    from telescope_b200.host import bam
    from telescope_b200.host.annotation import Annotation
    from telescope_b200.host.telescope import Telescope
    src = os.path.join(DATA, "alignment.bam")
    ops = "MIDNSHP=X"
    with bam.AlignmentReader(src, keep_raw=True) as r:
        refs, lens = r.references, r.lengths
        lines = ["@HD\tVN:1.0\tSO:unsorted\n"] + ["@SQ\tSN:%s\tLN:%d\n" % t for t in zip(refs, lens)]
        n_names, last = 0, None
        import struct
        for s in r:
            if s.name != last:
                n_names, last = n_names + 1, s.name
                if n_names > 60:
                    break
            n_cig = struct.unpack_from("<H", s.raw, 12)[0]
            l_name = s.raw[8]
            cig = struct.unpack_from("<%dI" % n_cig, s.raw, 32 + l_name)
            cigar = "".join("%d%s" % (c >> 4, ops[c & 15]) for c in cig) or "*"
            rnext = "=" if s.next_ref_id == s.ref_id else (refs[s.next_ref_id] if s.next_ref_id >= 0 else "*")
            lines.append("\t".join([s.name, str(s.flag), refs[s.ref_id], str(s.pos + 1), "255", cigar, rnext,
                                    str(s.next_pos + 1), str(s.tlen), "*", "*", "AS:i:%d" % s.score]) + "\n")
    sam = tmp_path / "first60.sam"
    sam.write_text("".join(lines))

    class O(AssignOpts):
        samfile = str(sam)
    ts = Telescope(O())
    ts.load_alignment(Annotation(O.gtffile, "locus", "None"))
    g = np.load(os.path.join(GOLD, "bundled.npz"))
    assert ts.shape[0] == 60 and ts.run_info["total_fragments"] == 60
    # same 60 reads as the first 60 rows of the BAM-derived matrix (scores are rescaled by the global minimum AS, so
    # compare structure and score differences within reads)
    ip = g["indptr"]
    assert np.array_equal(np.diff(ts.raw_scores.indptr), np.diff(ip[:61]))
    names = sorted(ts.feat_index, key=ts.feat_index.get)
    gold_names = [str(n) for n in g["feat_names"]]
    for r_ in range(60):
        a = {names[j]: int(v) for j, v in zip(ts.raw_scores.indices[ts.raw_scores.indptr[r_]:ts.raw_scores.indptr[r_ + 1]],
                                              ts.raw_scores.data[ts.raw_scores.indptr[r_]:ts.raw_scores.indptr[r_ + 1]])}
        b = {gold_names[j]: int(v) for j, v in zip(g["indices"][ip[r_]:ip[r_ + 1]], g["raw"][ip[r_]:ip[r_ + 1]])}
        assert set(a) == set(b)
        k0 = min(a)
        assert all(a[k] - a[k0] == b[k] - b[k0] for k in a)


# ------------------------------------------------------------------------------------------------ reference's own annotation tests
class TestAnnotationLikeTheReference(object):
    """reference telescope/tests/test_annotation_parsers.py:16-95 (TestAnnotationIntervalTree), same fixture
    (tests/data/annotation_test.2.gtf, copied from the reference's test data) and the same expected answers, against
    the array-based Annotation.  (`subregion` belongs to the reference's broken parallel loader and is not provided.)"""

    gtffile = os.path.join(ROOT, "tests", "data", "annotation_test.2.gtf")

    def setup_method(self):
        from telescope_b200.host.annotation import Annotation
        self.A = Annotation(self.gtffile, "locus")

    def test_annot_created(self):
        assert self.A.key == "locus"

    def test_annot_treesize(self):
        n = {chrom: len(idx[2]) for chrom, idx in self.A._index.items()}
        assert n == {"chr1": 3, "chr2": 4, "chr3": 2}

    def test_empty_lookups(self):
        A = self.A
        assert not A.intersect_blocks("chr1", [(1, 9999)])
        assert not A.intersect_blocks("chr1", [(20001, 39999)])
        assert not A.intersect_blocks("chr1", [(50001, 79999)])
        assert not A.intersect_blocks("chr1", [(90001, 90001)])
        assert not A.intersect_blocks("chr1", [(190000, 590000)])
        assert not A.intersect_blocks("chr2", [(1, 9999)])
        assert not A.intersect_blocks("chr3", [(1, 9999)])
        assert not A.intersect_blocks("chr4", [(1, 1000000000)])
        assert not A.intersect_blocks("chrX", [(1, 1000000000)])

    def test_simple_lookups(self):
        for line in open(self.gtffile):
            f = line.rstrip("\n").split("\t")
            iv = (int(f[3]), int(f[4]))
            loc = f[8].split('"')[1]
            r = self.A.intersect_blocks(f[0], [iv])
            assert loc in r
            assert (r[loc] - 1) == (iv[1] - iv[0])

    def test_overlap_lookups(self):
        A = self.A
        assert A.intersect_blocks("chr1", [(1, 10000)])["locus1"] == 1
        assert A.intersect_blocks("chr2", [(1, 10000)])["locus4"] == 1
        assert A.intersect_blocks("chr3", [(1, 10000)])["locus7"] == 1
        r = A.intersect_blocks("chr1", [(19990, 40000)])
        assert r["locus1"] == 11 and r["locus2"] == 1
        assert A.intersect_blocks("chr2", [(44990, 46010)])["locus5"] == 22
        assert A.intersect_blocks("chr3", [(44990, 46010)])["locus8"] == 1021


def test_header_is_plain_c_and_links(tmp_path):
    """include/telescope_b200.h must compile as C (not only C++) and every declared function must resolve against the
    built library -- what a cgo/cffi/ctypesgen user of the header would need."""
    import shutil
    import subprocess
    from telescope_b200 import _abi
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi_check.c"
    calls = "\n".join("    p[n++] = (void*)%s;" % s for s in _abi.SYMBOLS)
    src.write_text('#include "telescope_b200.h"\n#include <stdio.h>\nint main(void) {\n    void* p[64]; int n = 0;\n%s\n'
                   '    tsc_config c; tsc_config_default(&c);\n'
                   '    int ok = 0; for (int i = 0; i < n; ++i) ok += p[i] != 0;\n'
                   '    printf("%%d %%d %%d\\n", ok, tsc_abi_version(), c.n_local_devices);\n    return 0;\n}\n' % calls)
    exe = tmp_path / "abi_check"
    lib_dir = os.path.dirname(_abi.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-l:libtelescope_b200.so", "-Wl,-rpath," + lib_dir], check=True)
    out = subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, universal_newlines=True).stdout.split()
    assert out == [str(len(_abi.SYMBOLS)), "2", "1"]


def test_missing_library_fails_loudly(tmp_path):
    """No CUDA library -> an exception that says how to build it, never a silent CPU path."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from telescope_b200 import _abi\n"
            "try:\n    _abi.load()\nexcept _abi.TelescopeCudaError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n" % ROOT)
    env = dict(os.environ, TELESCOPE_B200_LIB=str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, universal_newlines=True, check=True).stdout
    assert out.strip() == "RAISED True"


def test_reference_arm_prints_the_contract_line():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--reads", "60000", "--loci", "500", "--cpu-seconds", "1"], stdout=subprocess.PIPE,
                         universal_newlines=True, check=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["metric"] == "em_iterations_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    # the unmodified reference class whenever it is reachable (live tree here, oracle/_ref on the GPU box), else the port
    from oracle import ref_shim
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("synthetic CSR")
    assert set(d["config"]) == {"workload", "n_reads", "n_loci", "nnz", "parallelism", "l2"} and d["config"]["nnz"] > 0


def test_staged_reference_is_the_unmodified_reference(tmp_path):
    """oracle/make_ref.py stages byte-identical copies of the reference's Python files (plus generated stubs) and the
    staged tree imports and reproduces the README's known answer on the bundled matrix."""
    import filecmp
    import subprocess
    import sys
    from oracle import make_ref, ref_shim
    if not os.path.isdir("/root/reference/telescope"):
        pytest.skip("reference tree not present")
    make_ref.make()
    for rel in ("telescope/utils/model.py", "telescope/utils/sparse_plus.py", "telescope/utils/helpers.py"):
        assert filecmp.cmp(os.path.join("/root/reference", rel), os.path.join(make_ref.DEST, rel), shallow=False)
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from oracle import ref_shim\n"
            "assert ref_shim.reference_is_staged_copy()\n"
            "model, csr = ref_shim.import_reference()\n"
            "g = np.load(%r)\n"
            "import scipy.sparse as sp\n"
            "m = csr(sp.csr_matrix((g['raw'], g['indices'], g['indptr']), shape=tuple(g['shape'])))\n"
            "tl = model.TelescopeLikelihood(m, ref_shim.RefOpts()); tl.em()\n"
            "print('LNL %%.6f' %% tl.lnl)\n") % (ROOT, os.path.join(ROOT, "tests", "golden", "bundled.npz"))
    env = dict(os.environ, TELESCOPE_REFERENCE_ROOT=make_ref.DEST)
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, universal_newlines=True, check=True).stdout
    assert "LNL 95252.596293" in out
