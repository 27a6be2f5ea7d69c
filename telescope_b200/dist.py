# -*- coding: utf-8 -*-
"""Bootstrap for one-process-per-GPU runs on one node (e.g. under `python -m torch.distributed.run`), without torch.

All the EM path needs from the launcher is RANK / WORLD_SIZE / LOCAL_RANK and a way to hand rank 0's 128-byte NCCL id
to the other ranks.  The ranks of one launch share a parent process (the launcher's agent), so the id travels through
a file in the temp directory named after that parent's pid and MASTER_PORT; everything afterwards (barriers, max over
ranks, the per-iteration all-reduce) goes over the library's own NCCL communicator.
"""
import os
import tempfile
import time

from . import _abi
from .likelihood import DistInfo


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _parent_start_ticks():
    """Start time of the launcher process (clock ticks since boot): makes the key unique even if a pid is recycled."""
    try:
        with open("/proc/%d/stat" % os.getppid()) as fh:
            return fh.read().rsplit(")", 1)[1].split()[19]
    except Exception:
        return "0"


def _rdzv_path():
    tag = "%d_%s_%s" % (os.getppid(), _parent_start_ticks(), os.environ.get("MASTER_PORT", "0"))
    return os.path.join(tempfile.gettempdir(), "telescope_b200_rdzv_" + tag)


def rendezvous(timeout=600.0):
    """DistInfo for this process (None when WORLD_SIZE is 1)."""
    rank, world, _ = env_world()
    if world <= 1:
        return None
    path = _rdzv_path()
    if rank == 0:
        ident = _abi.nccl_unique_id()
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as fh:
            fh.write(ident)
        os.replace(tmp, path)
    else:
        t0, born = time.time(), time.time() - 3600.0
        while True:
            try:
                st = os.stat(path)
                if st.st_size == 128 and st.st_mtime >= born:
                    with open(path, "rb") as fh:
                        ident = fh.read()
                    if len(ident) == 128:
                        break
            except OSError:
                pass
            if time.time() - t0 > timeout:
                raise RuntimeError("rank %d: no NCCL id from rank 0 at %s after %.0f s" % (rank, path, timeout))
            time.sleep(0.02)
    return DistInfo(world, rank, ident)


def file_barrier(tag, timeout=600.0):
    """Barrier for the ranks of one launch before any communicator exists (same shared-parent file scheme)."""
    rank, world, _ = env_world()
    if world <= 1:
        return
    base = "%s.%s." % (_rdzv_path(), tag)
    with open(base + str(rank), "w"):
        pass
    t0 = time.time()
    while not all(os.path.exists(base + str(r)) for r in range(world)):
        if time.time() - t0 > timeout:
            raise RuntimeError("rank %d: barrier %r timed out" % (rank, tag))
        time.sleep(0.005)


def cleanup():
    rank, world, _ = env_world()
    if world > 1 and rank == 0:
        import glob
        for f in glob.glob(_rdzv_path() + "*"):
            try:
                os.unlink(f)
            except OSError:
                pass
