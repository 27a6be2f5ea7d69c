# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- import the *real* reference implementation from /root/reference.

The reference (mlbendall/telescope @ 4cf18595) is pure Python on scipy.sparse, but `telescope/utils/model.py`
imports four things this image does not have (`past.utils.old_div` model.py:4, `pysam` model.py:17, the compiled
`telescope.utils.calignment` via alignment.py:15, and `future.standard_library` sparse_plus.py:6-7).  None of them
is touched by the EM path (`TelescopeLikelihood`, model.py:631-865; `csr_matrix_plus`, sparse_plus.py:24-174), so
empty stand-in modules are registered before the import.

`/root/reference` exists only in the build container, never on the GPU box; there the unmodified copy staged by
`oracle/make_ref.py` under the git-ignored `oracle/_ref/` is imported instead.  Used by
`tests/golden/make_golden.py` (to generate the committed fixtures), by the `not gpu` tests that cross-check the
oracle restatements, and by `bench.py`'s CPU arm (`cpu_baseline.kind == "reference"`).  Nothing in the product
imports it.
"""
import os
import sys
import types
import warnings

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # written by oracle/make_ref.py


def _pick_root():
    """The live tree in the build container; on the GPU box the unmodified copy staged by oracle/make_ref.py."""
    env = os.environ.get("TELESCOPE_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", _STAGED]:
        if os.path.isfile(os.path.join(cand, "telescope", "utils", "model.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "telescope", "utils", "model.py"))


def reference_is_staged_copy():
    return os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_STAGED)


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Return (model_module, csr_matrix_plus) of the unmodified reference."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    fut = _stub("future")
    fut.standard_library = _stub("future.standard_library", install_aliases=lambda: None)
    past = _stub("past")
    past.utils = _stub("past.utils", old_div=lambda a, b: a // b if isinstance(a, int) and isinstance(b, int) else a / b)
    _stub("pysam")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub("telescope.utils.calignment", AlignedPair=object)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import telescope.utils.model as model
        from telescope.utils.sparse_plus import csr_matrix_plus
    return model, csr_matrix_plus


class RefOpts(object):
    """The four attributes TelescopeLikelihood.__init__ reads (model.py:661-662,686-687) plus report options."""

    def __init__(self, em_epsilon=1e-7, max_iter=100, pi_prior=0, theta_prior=200000,
                 reassign_mode="exclude", conf_prob=0.9):
        self.em_epsilon = em_epsilon
        self.max_iter = max_iter
        self.pi_prior = pi_prior
        self.theta_prior = theta_prior
        self.reassign_mode = reassign_mode
        self.conf_prob = conf_prob
