# -*- coding: utf-8 -*-
"""Bootstrap for one-process-per-GPU runs on one node (e.g. under `python -m torch.distributed.run`), without torch.

All the EM path needs from the launcher is RANK / WORLD_SIZE / LOCAL_RANK and a way to swap a few small blobs between
the ranks before any GPU communication exists: the 64-byte CUDA IPC handles of the peer-memory exchange buffers (or
rank 0's 128-byte NCCL id when the NCCL transport is used).  The ranks of one launch share a parent process (the
launcher's agent), so the blobs travel through files in a private (0700, owner-checked) directory, named after

    parent pid + parent start time + MASTER_PORT + TORCHELASTIC_RUN_ID + TORCHELASTIC_RESTART_COUNT + tag + call number

-- unique per launch ATTEMPT and per call, so a restarted worker group or a second rendezvous in the same launch never
reads a stale blob.  Everything afterwards (barriers, max over ranks, the per-iteration exchange) goes over the
library's own transport.
"""
import os
import stat
import tempfile
import time

from . import _abi
from .likelihood import DistInfo

_calls = {}


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _parent_start_ticks():
    """Start time of the launcher process (clock ticks since boot): makes the key unique even if a pid is recycled."""
    try:
        with open("/proc/%d/stat" % os.getppid()) as fh:
            return fh.read().rsplit(")", 1)[1].split()[19]
    except Exception:
        return "0"


def _private_dir():
    d = os.path.join(tempfile.gettempdir(), "telescope_b200_%d" % os.getuid())
    try:
        os.mkdir(d, 0o700)
    except FileExistsError:
        pass
    st = os.lstat(d)
    if not stat.S_ISDIR(st.st_mode) or st.st_uid != os.getuid() or (st.st_mode & 0o077):
        raise RuntimeError("%s is not a private directory of uid %d; refusing to exchange ids through it" % (d, os.getuid()))
    return d


def _attempt_key():
    e = os.environ
    return "%d_%s_%s_%s_%s" % (os.getppid(), _parent_start_ticks(), e.get("MASTER_PORT", "0"),
                               e.get("TORCHELASTIC_RUN_ID", "x").replace(os.sep, "_"), e.get("TORCHELASTIC_RESTART_COUNT", "0"))


def _base(tag):
    n = _calls.get(tag, 0)
    _calls[tag] = n + 1
    return os.path.join(_private_dir(), "%s.%s.%d." % (_attempt_key(), tag, n))


def allgather(tag, payload=b"", timeout=600.0):
    """Every rank contributes `payload` (bytes); returns the list of all ranks' payloads.  Doubles as a barrier."""
    rank, world, _ = env_world()
    if world <= 1:
        return [payload]
    base = _base(tag)
    tmp = base + "tmp%d" % rank
    with open(tmp, "wb") as fh:
        fh.write(payload)
    os.replace(tmp, base + str(rank))            # atomic: a reader sees the whole blob or nothing
    out, t0, polls = [None] * world, time.time(), 0
    while True:
        for r in range(world):
            if out[r] is None:
                try:
                    with open(base + str(r), "rb") as fh:
                        out[r] = fh.read()
                except OSError:
                    pass
        if all(o is not None for o in out):
            return out
        if time.time() - t0 > timeout:
            missing = [r for r in range(world) if out[r] is None]
            raise RuntimeError("rank %d: no %r blob from ranks %s after %.0f s" % (rank, tag, missing, timeout))
        polls += 1
        if polls > 400:                  # ranks of one launch arrive within a millisecond or two of each other: look again
            time.sleep(0.0005)           # at once for a while before yielding the core between looks


def rendezvous(timeout=600.0, transport="peer"):
    """DistInfo for this process (None when WORLD_SIZE is 1).

    transport "peer": the ranks exchange CUDA IPC handles of their exchange buffers at construction (no NCCL at all);
    "nccl": rank 0's NCCL id is handed to every rank here."""
    rank, world, _ = env_world()
    if world <= 1:
        return None
    ident = None
    if transport == "nccl":
        ident = allgather("ncclid", _abi.nccl_unique_id() if rank == 0 else b"", timeout)[0]
        if len(ident) != 128:
            raise RuntimeError("rank %d: malformed NCCL id from rank 0" % rank)
    return DistInfo(world, rank, ident, allgather=allgather, transport=transport)


def file_barrier(tag, timeout=600.0):
    """Barrier for the ranks of one launch before any communicator exists."""
    allgather("barrier_" + tag, b"", timeout)


def cleanup(timeout=30.0):
    """Remove this launch attempt's files.  Called by EVERY rank as its last rendezvous action: the other ranks
    acknowledge that they will read nothing more, then rank 0 deletes."""
    rank, world, _ = env_world()
    if world <= 1:
        return
    base = _base("cleanup")
    if rank != 0:
        with open(base + str(rank), "wb"):
            pass
        return
    t0 = time.time()
    while not all(os.path.exists(base + str(r)) for r in range(1, world)):
        if time.time() - t0 > timeout:
            return                      # somebody died: leave the files, they are keyed to this attempt only
        time.sleep(0.002)
    import glob
    for f in glob.glob(os.path.join(_private_dir(), _attempt_key() + ".*")):
        try:
            os.unlink(f)
        except OSError:
            pass
