# -*- coding: utf-8 -*-
"""Build libtelescope_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtelescope_b200.so")
SOURCES = [os.path.join(CSRC, "tsc_api.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("tsc_kernels.cuh", "tsc_tiles.cuh", "tsc_ell.cuh")] + \
          [os.path.join(os.path.dirname(HERE), "include", "telescope_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # keep n = Q*pt and sum += n as separate roundings, like the reference
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def find_nvcc():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC=...)")
    return cand


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False, out=None, defines=()):
    """Compile the CUDA library; returns its path.  Raises on failure -- there is no fallback.
    `out` / `defines` build tuning variants (e.g. -DTSC_TILE_WARPS=8) next to the default library."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out or LIB] + SOURCES + ["-ldl"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d)" % res.returncode)
    if out is None:
        with open(os.path.join(HERE, "libtelescope_b200.ptxas.log"), "w") as fh:
            fh.write(res.stdout)
    return out or LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
