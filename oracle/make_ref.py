# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference package under oracle/_ref/ so that it can
travel to the GPU box (where /root/reference does not exist) and be timed there as the CPU arm.

    python -m oracle.make_ref            # needs /root/reference (the build container); idempotent

What it does
  * copies the reference's pure-Python package files (`telescope/**/*.py`, byte for byte, no edits) from
    $TELESCOPE_REFERENCE_ROOT (default /root/reference) to oracle/_ref/telescope/ -- the timed code is
    telescope/utils/model.py:631-865 and telescope/utils/sparse_plus.py:16-165 themselves;
  * writes four stand-in packages next to it (oracle/_ref/_stubs/): `future`, `past`, `pysam`, and the compiled
    extension `telescope.utils.calignment` -- modules model.py imports at load time but the EM path never calls
    (see oracle/ref_shim.py for the file:line of each import).  The stubs are generated here, not copied.

oracle/_ref/ is git-ignored (never committed: the repository holds no reference sources) but NOT gpurun-ignored,
so the staged copy ships with the snapshot like our own built .so.  `__graft_entry__.build()` runs this recipe
whenever the reference tree is present.  Nothing in the product imports oracle/.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("TELESCOPE_REFERENCE_ROOT", "/root/reference")

_STUBS = {
    "future/__init__.py": "",
    "future/standard_library.py": "def install_aliases():\n    pass\n",
    "past/__init__.py": "",
    "past/utils.py": ("def old_div(a, b):\n"
                      "    if isinstance(a, int) and isinstance(b, int):\n"
                      "        return a // b\n"
                      "    return a / b\n"),
    "pysam/__init__.py": "",
}
# the compiled Cython extension (calignment.pyx) is only used while parsing BAM files
_CALIGNMENT_STUB = "AlignedPair = object\n"


def staged():
    return os.path.isfile(os.path.join(DEST, "telescope", "utils", "model.py"))


def make(verbose=False):
    src_pkg = os.path.join(SRC_ROOT, "telescope")
    if not os.path.isfile(os.path.join(src_pkg, "utils", "model.py")):
        raise RuntimeError("reference tree not present at %s" % SRC_ROOT)
    n = 0
    for dirpath, dirnames, filenames in os.walk(src_pkg):
        dirnames[:] = [d for d in dirnames if d not in ("tests", "data", "__pycache__")]
        rel = os.path.relpath(dirpath, SRC_ROOT)
        os.makedirs(os.path.join(DEST, rel), exist_ok=True)
        for f in filenames:
            if not f.endswith(".py"):
                continue
            s, d = os.path.join(dirpath, f), os.path.join(DEST, rel, f)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
            n += 1
    with open(os.path.join(DEST, "telescope", "utils", "calignment.py"), "w") as fh:
        fh.write(_CALIGNMENT_STUB)
    for rel, body in _STUBS.items():
        p = os.path.join(DEST, "_stubs", rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as fh:
            fh.write(body)
    with open(os.path.join(DEST, "SOURCE"), "w") as fh:
        fh.write("staged from %s by oracle/make_ref.py (unmodified *.py files + generated stubs)\n" % SRC_ROOT)
    if verbose:
        sys.stderr.write("oracle/_ref: %d reference files staged\n" % n)
    return DEST


if __name__ == "__main__":
    print(make(verbose=True))
