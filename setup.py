# -*- coding: utf-8 -*-
"""Packaging for telescope-b200.  The CUDA library is built in-tree first (`python -m telescope_b200.build`, needs
nvcc; sm_100a only) and shipped as package data; the console script keeps the reference's name (`telescope`,
reference setup.py:58-62) so `telescope assign | resume | test` work unchanged."""
import os
import re

from setuptools import find_packages, setup

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "telescope_b200", "__init__.py")) as fh:
    VERSION = re.search(r'__version__ = "([^"]+)"', fh.read()).group(1)

setup(
    name="telescope-b200",
    version=VERSION,
    description="Telescope's EM reassignment loop as sm_100a CUDA kernels behind a C ABI (drop-in TelescopeLikelihood)",
    packages=find_packages(include=["telescope_b200", "telescope_b200.*"]),
    package_data={"telescope_b200": ["libtelescope_b200.so", "data/*", "csrc/*"]},
    python_requires=">=3.8",
    install_requires=["numpy", "scipy", "pandas"],
    entry_points={"console_scripts": ["telescope=telescope_b200.cli:main"]},
)
