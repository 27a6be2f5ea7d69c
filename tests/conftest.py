import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The C-ABI library must exist for both tiers (symbol checks on CPU, everything on GPU)."""
    from telescope_b200 import build
    build.build_library()        # rebuilds when a source is newer than the .so; a failed build fails the session --
    return build.LIB             # never test against a stale binary


class Opts(object):
    def __init__(self, em_epsilon=1e-7, max_iter=100, pi_prior=0, theta_prior=200000, reassign_mode="exclude",
                 conf_prob=0.9):
        self.em_epsilon, self.max_iter = em_epsilon, max_iter
        self.pi_prior, self.theta_prior = pi_prior, theta_prior
        self.reassign_mode, self.conf_prob = reassign_mode, conf_prob


@pytest.fixture
def opts_cls():
    return Opts


def rel_err(a, b):
    """max |a-b| / |b| over entries where b != 0, and max |a| where b == 0."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nz = b != 0
    r = np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])) if nz.any() else 0.0
    z = np.max(np.abs(a[~nz])) if (~nz).any() else 0.0
    return max(r, z)
