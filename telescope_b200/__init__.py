# -*- coding: utf-8 -*-
"""telescope_b200 -- Telescope's EM reassignment loop (TelescopeLikelihood) as sm_100a CUDA kernels behind a C ABI,
plus the host-side pieces of the drop-in surface (checkpoint, reports, CLI)."""
__version__ = "1.0.3+b200.r1"
