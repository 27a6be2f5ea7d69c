# -*- coding: utf-8 -*-
"""`telescope assign | resume | test` -- same sub-commands, options, log lines and output files as the reference
CLI (telescope/__main__.py:49-92, telescope_assign.py:48-185,372-451, telescope_resume.py:29-103,183-232), with the
EM loop running on the GPU through `telescope_b200.likelihood.TelescopeLikelihood`.

Only GPU-selection flags are new: --devices (comma-separated CUDA ordinals; default "0").
"""
from __future__ import print_function

import argparse
import errno
import logging as lg
import os
import sys
from time import time

import numpy as np

from . import __version__

USAGE = ''' %(prog)s <command> [<args>]

The most commonly used commands are:
   assign    Reassign ambiguous fragments that map to repetitive elements
   resume    Resume previous run from checkpoint file
   test      Generate a command line for testing
'''

REASSIGN_HELP = ('Reassignment mode. After EM is complete, each fragment is reassigned according to the expected '
                 'value of its membership weights. "exclude" - fragments with multiple best assignments are excluded '
                 'from the final counts; "choose" - the best assignment is randomly chosen from among the set of '
                 'best assignments; "average" - the fragment is divided evenly among the best assignments; "conf" - '
                 'only assignments that exceed a certain threshold (see --conf_prob) are accepted; "unique" - only '
                 'uniquely aligned reads are included.')


def fmtmins(seconds):
    return '%d minutes and %d secs' % (seconds // 60, seconds % 60)


def _reporting(p):
    g = p.add_argument_group('Reporting Options', '')
    g.add_argument('--quiet', action='store_true', help='Silence (most) output.')
    g.add_argument('--debug', action='store_true', help='Print debug messages.')
    g.add_argument('--logfile', type=argparse.FileType('w'), help='Log output to this file.')
    g.add_argument('--outdir', default='.', help='Output directory.')
    g.add_argument('--exp_tag', default='telescope', help='Experiment tag')
    return g


def _model(p, skip_em):
    g = p.add_argument_group('Model Parameters', '')
    g.add_argument('--pi_prior', type=int, default=0, help='Prior on pi. Equivalent to adding n unique reads.')
    g.add_argument('--theta_prior', type=int, default=200000,
                   help='Prior on theta. Equivalent to adding n non-unique reads.')
    g.add_argument('--em_epsilon', type=float, default=1e-7, help='EM Algorithm Epsilon cutoff')
    g.add_argument('--max_iter', type=int, default=100, help='EM Algorithm maximum iterations')
    g.add_argument('--use_likelihood', action='store_true',
                   help='Use difference in log-likelihood as convergence criteria.')
    if skip_em:
        g.add_argument('--skip_em', action='store_true',
                       help='Exits after loading alignment and saving checkpoint file.')
    d = p.add_argument_group('GPU Options', '')
    d.add_argument('--devices', default='0', help='CUDA device ordinals to shard the reads over, e.g. 0,1,2,3')


def add_assign_arguments(p):
    g = p.add_argument_group('Input Options', '')
    g.add_argument('samfile', help='Path to alignment file (SAM or BAM), collated so that all alignments for a '
                                   'read pair appear sequentially in the file.')
    g.add_argument('gtffile', help='Path to annotation file (GTF format)')
    g.add_argument('--attribute', default='locus', help='GTF attribute that defines a transposable element locus.')
    g.add_argument('--no_feature_key', default='__no_feature', help='Used internally to represent alignments.')
    g.add_argument('--ncpu', default=1, type=int, help='Number of cores to use. (Multiple cores not supported yet).')
    g.add_argument('--tempdir', help='Path to temporary directory.')
    r = _reporting(p)
    r.add_argument('--updated_sam', action='store_true', help='Generate an updated alignment file.')
    m = p.add_argument_group('Run Modes', '')
    m.add_argument('--reassign_mode', default='exclude', choices=['all', 'exclude', 'choose', 'average', 'conf', 'unique'],
                   help=REASSIGN_HELP)
    m.add_argument('--use_every_reassign_mode', action='store_true',
                   help='Whether to output count matrices generated using every reassign mode.')
    m.add_argument('--conf_prob', type=float, default=0.9, help='Minimum probability for high confidence assignment.')
    m.add_argument('--overlap_mode', default='threshold', choices=['threshold', 'intersection-strict', 'union'],
                   help='Overlap mode.')
    m.add_argument('--overlap_threshold', type=float, default=0.2,
                   help='Fraction of fragment that must be contained within a feature to be assigned to that locus.')
    m.add_argument('--annotation_class', default='intervaltree', choices=['intervaltree', 'htseq'],
                   help='Accepted for compatibility; overlaps are computed with sorted arrays.')
    m.add_argument('--stranded_mode', type=str, default='None', choices=['None', 'RF', 'R', 'FR', 'F'],
                   help='Options for considering feature strand when assigning reads.')
    _model(p, skip_em=True)


def add_resume_arguments(p):
    g = p.add_argument_group('Input Options', '')
    g.add_argument('checkpoint', help='Path to checkpoint file.')
    _reporting(p)
    m = p.add_argument_group('Run Modes', '')
    m.add_argument('--reassign_mode', default='exclude', choices=['exclude', 'choose', 'average', 'conf', 'unique'],
                   help=REASSIGN_HELP)
    m.add_argument('--conf_prob', type=float, default=0.9, help='Minimum probability for high confidence assignment.')
    _model(p, skip_em=False)


class Options(object):
    """argparse namespace + the helpers the reference's option classes provide (telescope_assign.py:29-46)."""

    def __init__(self, args):
        for k, v in vars(args).items():
            setattr(self, k, v)
        self.version = __version__
        if getattr(self, 'logfile', None) is None:
            self.logfile = sys.stderr

    def outfile_path(self, suffix):
        return os.path.join(self.outdir, '%s-%s' % (self.exp_tag, suffix))

    def device_list(self):
        return [int(x) for x in str(getattr(self, 'devices', '0')).split(',') if x != '']

    def __str__(self):
        skip = ('func', 'subcommand', 'logfile')
        rows = ['{:34}{}'.format('Version:', self.version)]
        rows += ['    {:30}{}'.format(k + ':', v) for k, v in sorted(vars(self).items()) if k not in skip]
        return '\n'.join(rows)


def configure_logging(opts):
    loglev = lg.WARNING if opts.quiet else lg.INFO
    if opts.debug:
        loglev = lg.DEBUG
    fmt = '%(asctime)s %(levelname)-8s %(message)-60s (from %(funcName)s in %(filename)s:%(lineno)d)'
    lg.basicConfig(level=loglev, format=fmt, datefmt='%Y-%m-%d %H:%M:%S', stream=opts.logfile, force=True)


def _run_em_and_report(ts, opts, total_time, what):
    from .likelihood import TelescopeLikelihood
    seed = ts.get_random_seed()
    lg.debug("Random seed: {}".format(seed))
    np.random.seed(seed)
    ts_model = TelescopeLikelihood(ts.raw_scores, opts, devices=opts.device_list())
    lg.info('Running Expectation-Maximization...')
    stime = time()
    ts_model.em(use_likelihood=opts.use_likelihood, loglev=lg.INFO)
    lg.info("EM completed in %s" % fmtmins(time() - stime))
    lg.info("Generating Report...")
    ts.output_report(ts_model, opts.outfile_path('run_stats.tsv'), opts.outfile_path('TE_counts.tsv'))
    if getattr(opts, 'updated_sam', False):
        lg.info("Creating updated SAM file...")
        ts.update_sam(ts_model, opts.outfile_path('updated.bam'))
    lg.info("telescope %s complete (%s)" % (what, fmtmins(time() - total_time)))
    ts_model.close()


def run_assign(args):
    from .host.annotation import Annotation
    from .host.telescope import Telescope
    opts = Options(args)
    configure_logging(opts)
    lg.info('\n{}\n'.format(opts))
    total_time = time()
    ts = Telescope(opts)
    lg.info('Loading annotation...')
    stime = time()
    annot = Annotation(opts.gtffile, opts.attribute, opts.stranded_mode)
    lg.info("Loaded annotation in {}".format(fmtmins(time() - stime)))
    lg.info('Loaded {} features.'.format(len(annot.loci)))
    lg.info('Loading alignments...')
    stime = time()
    ts.load_alignment(annot)
    lg.info("Loaded alignment in {}".format(fmtmins(time() - stime)))
    ts.print_summary(lg.INFO)
    if ts.run_info['overlap_unique'] + ts.run_info['overlap_ambig'] == 0:
        lg.info("No alignments overlapping annotation")
        lg.info("telescope assign complete (%s)" % fmtmins(time() - total_time))
        return
    annot = None
    ts.save(opts.outfile_path('checkpoint'))
    if opts.skip_em:
        lg.info("Skipping EM...")
        lg.info("telescope assign complete (%s)" % fmtmins(time() - total_time))
        return
    _run_em_and_report(ts, opts, total_time, 'assign')


def run_resume(args):
    from .host.telescope import Telescope
    opts = Options(args)
    configure_logging(opts)
    lg.info('\n{}\n'.format(opts))
    total_time = time()
    lg.info('Loading Telescope object from file...')
    ts = Telescope.load(opts.checkpoint)
    ts.opts = opts
    ts.print_summary(lg.INFO)
    _run_em_and_report(ts, opts, total_time, 'resume')


def generate_test_command(args):
    data = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
    aln, gtf = os.path.join(data, 'alignment.bam'), os.path.join(data, 'annotation.gtf')
    for p in (aln, gtf):
        if not os.path.exists(p):
            raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), p)
    print('telescope assign %s %s' % (aln, gtf), file=sys.stdout)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        argparse.ArgumentParser(description='Tools for analysis of repetitive DNA elements', usage=USAGE).print_help(sys.stderr)
        sys.exit(1)
    parser = argparse.ArgumentParser(prog='telescope', description='Tools for analysis of repetitive DNA elements')
    parser.add_argument('--version', action='version', version=__version__, default=__version__)
    sub = parser.add_subparsers(help='Sub-command help', dest='subcommand')
    p = sub.add_parser('assign', description='Reassign ambiguous fragments that map to repetitive elements',
                       formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_assign_arguments(p)
    p.set_defaults(func=run_assign)
    p = sub.add_parser('resume', description='Resume a previous telescope run',
                       formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_resume_arguments(p)
    p.set_defaults(func=run_resume)
    p = sub.add_parser('test', description='Print a test command', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.set_defaults(func=generate_test_command)
    args = parser.parse_args(argv)
    args.func(args)


if __name__ == '__main__':
    main()
