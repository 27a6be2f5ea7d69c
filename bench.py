#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- EM iterations/sec of Telescope's reassignment loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload: the 50 M-read x 30 k-locus synthetic CSR (avg 20 alignments/read, ~1e9 entries, fp64) that
BASELINE.json's metric is quoted on; with N GPUs the same matrix is row-sharded (strong scaling).  A "step" is one
EM iteration (E-step + M-step + parameter update) over the whole matrix; em_epsilon = -1 keeps the loop from
stopping early, exactly as on the CPU arm.

  value    = K / device time of `em(max_iter=K)` (CUDA events on the library's stream, max over ranks), inputs
             resident in HBM.  Like the reference's em(), the call ends with one calculate_lnl pass (model.py:800-801).
  e2e      = K / wall time of the whole job through the public class with HOST (pinned) CSR buffers:
             TelescopeLikelihood(csr, opts) [H2D of the CSR, Q build, clustering] + em(K) + D2H of pi/theta.
             Three such jobs run back to back, each on a fresh model; the line reports the median and lists all
             three (`runs`, `runs_laps_ms`): the first job of a process allocates its device memory (and the staging
             pool / IPC mappings) from the driver, later ones reuse the blocks the destroyed model handed back.
             e2e_pageable = the same from ordinary (pageable) numpy/scipy arrays, e2e_report = e2e plus the one-pass
             output_report column sums (model.py:432-458).
  roofline = the fused E+M kernel(s) of one iteration: ALGORITHMIC bytes nnz*12 + (N+1)*4 / mean
             per-iteration kernel time (CUDA events inside tsc_em), against MEASURED_PEAKS.json's HBM copy bandwidth;
             `traffic` = DRAM bytes one launch really moves (ncu capture, profiles/) -- the clustered stream stores a
             locus in 1 byte and skips unique reads, so it is BELOW the algorithmic bytes; `frac_dram` is the fraction
             on those real bytes.
  parity   = every rank feeds the first rows of its block to a second N-rank model; rank 0 runs the CPU oracle
             (oracle/em_numpy.py) on the concatenated sample: lnl, pi, theta to 1e-6, `exclude` counts bit-exact.
  other_configs.zipf = BASELINE.json config 5 (Zipf rows, max 200 entries/read): value, fused fraction, parity.
  cpu_baseline = the UNMODIFIED reference class (oracle/_ref, staged by oracle/make_ref.py; the scipy port when it
             is absent) on the first 4 % of the reads, single-threaded like scipy, scaled linearly in entries.

`--impl reference` prints the same line for the CPU arm alone (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "em_iterations_per_sec"
UNIT = "iter/s"
SEED = 1004
CPU_SAMPLE_FRACTION = 0.04         # of the reads, for both CPU legs (cpu_baseline and --impl reference)


class Opts(object):
    def __init__(self, max_iter):
        self.em_epsilon = -1.0          # never "<": fixed iteration count on both arms
        self.max_iter = max_iter
        self.pi_prior = 0
        self.theta_prior = 200000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=50_000_000)
    ap.add_argument("--loci", type=int, default=30_000)
    ap.add_argument("--avg", type=int, default=20)
    ap.add_argument("--skew", action="store_true", help="config 5 as the main workload: Zipf reads-per-row (max 200)")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--replicas", type=int, default=0)
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--cpu-seconds", type=float, default=170.0, help="--impl reference: CPU budget for W+K iterations")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip pageable e2e, parity and the Zipf configuration")
    ap.add_argument("--permute", action="store_true", help="renumber loci by descending entry count on the device")
    return ap.parse_args()


def _si(n):
    for div, suf in ((10 ** 6, "M"), (10 ** 3, "k")):
        if n >= div and n % div == 0:
            return "%d%s" % (n // div, suf)
    return str(n)


def workload_name(a, skew=None):
    skew = a.skew if skew is None else skew
    return "synthetic CSR %s reads x %s loci, avg %d alignments/read%s, seed %d" % (
        _si(a.reads), _si(a.loci), a.avg, ", Zipf rows (max 200)" if skew else "", SEED)


def base_config(a, world, total_nnz, alg_bytes_per_gpu):
    """The part of `config` both arms print identically."""
    return {
        "workload": workload_name(a), "n_reads": a.reads, "n_loci": a.loci, "nnz": int(total_nnz),
        "parallelism": ("read-sharded x%d, one exchange of K doubles per iteration" % world) if world > 1 else "1 GPU",
        "l2": "per-iteration inputs (%.1f GB per GPU) are larger than L2, no flush needed" % (alg_bytes_per_gpu / 1e9),
    }


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_engine():
    """('reference', factory) with the unmodified reference class when it is importable, else ('port', factory)."""
    from oracle import ref_shim
    if ref_shim.reference_available():
        model, csr_plus = ref_shim.import_reference()
        import logging as lg

        class RefRun(object):
            def __init__(self, m, max_iter):
                self.tl = model.TelescopeLikelihood(csr_plus(m), ref_shim.RefOpts(em_epsilon=-1.0, max_iter=max_iter))

            def em(self, n):
                self.tl.max_iter = n
                self.tl.em(use_likelihood=False, loglev=lg.DEBUG)
                return self

            lnl = property(lambda self: float(self.tl.lnl))
            pi = property(lambda self: np.asarray(self.tl.pi))
            theta = property(lambda self: np.asarray(self.tl.theta))
        where = "oracle/_ref (staged copy)" if ref_shim.reference_is_staged_copy() else ref_shim.REFERENCE_ROOT
        return "reference", RefRun, "telescope.utils.model.TelescopeLikelihood, unmodified, from " + where
    from oracle.em_scipy import ScipyEM

    class PortRun(object):
        def __init__(self, m, max_iter):
            self.o = ScipyEM(m, em_epsilon=-1.0, max_iter=max_iter)

        def em(self, n):
            self.o.max_iter = n
            self.o.em()
            return self

        lnl = property(lambda self: float(self.o.lnl))
        pi = property(lambda self: np.asarray(self.o.pi))
        theta = property(lambda self: np.asarray(self.o.theta))
    return "port", PortRun, "oracle/em_scipy.py (op-for-op scipy port; the staged reference is missing)"


def cpu_sample_rows(a, n_iters=None, budget_s=None):
    rows = int(max(20_000, min(a.reads, round(a.reads * CPU_SAMPLE_FRACTION))))
    if n_iters and budget_s:       # keep W+K iterations inside the budget (~180 ns per entry and iteration + init)
        rows = int(min(rows, max(20_000, budget_s / (180e-9 * a.avg * (n_iters + 2.0)))))
    return rows


def run_cpu(a, n_warm, n_iters, rows, full_nnz):
    """Time the CPU implementation on the first `rows` reads; value = iter/s scaled linearly to the full workload."""
    import scipy.sparse as sp
    from telescope_b200.synthetic import synth_csr
    kind, Run, what = cpu_engine()
    ip, ix, raw = synth_csr(a.reads, a.loci, a.avg, a.skew, SEED, 0, rows)
    m = sp.csr_matrix((np.asarray(raw), np.asarray(ix), np.asarray(ip)), shape=(rows, a.loci))
    t0 = time.perf_counter()
    run = Run(m, max(1, n_warm))
    t_init = time.perf_counter() - t0
    if n_warm > 0:
        run.em(n_warm)
    t0 = time.perf_counter()
    run.em(n_iters)
    dt = time.perf_counter() - t0
    return {
        "value": (n_iters / dt) * m.nnz / full_nnz,
        "unit": UNIT,
        "cores": 1,
        "kind": kind,
        "sample": "first %d reads (%d entries, %.3g of the workload's %d), %d warm-up + %d timed EM iterations (the call "
                  "ends with the reference's final calculate_lnl) in %.1f s = %.1f ns/entry/iter, construction %.1f s; "
                  "scaled linearly in entries to the full matrix; %s" % (
                      rows, m.nnz, m.nnz / float(full_nnz), full_nnz, n_warm, n_iters, dt, dt / n_iters / m.nnz * 1e9,
                      t_init, what),
        "host_cpus": os.cpu_count(),
        "_lnl": run.lnl, "_pi": run.pi, "_theta": run.theta, "_rows": rows,
    }


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = tempfile.mktemp(prefix="tsc_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------- peaks
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(nnz, skew):
    """DRAM bytes per launch of the fused kernel from the committed ncu capture, scaled per entry (or None)."""
    p = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        key = "dram_bytes_per_entry_zipf" if skew and "dram_bytes_per_entry_zipf" in d else "dram_bytes_per_entry"
        return float(d[key]) * nnz
    except Exception:
        return None


def fused_kernel_name(layout):
    if not layout["slices"]:
        return "k_tiles<TILE_FUSED> (E-step + M-step accumulation, telescope_b200/csrc/tsc_tiles.cuh)"
    s = "k_ell_fused (E-step + M-step accumulation over the clustered slice stream, telescope_b200/csrc/tsc_ell.cuh)"
    if layout["residual_entries"]:
        s += " + k_tiles<TILE_FUSED> on the residual CSR (%.3g of the ambiguous entries)" % (
            layout["residual_entries"] / float(layout["residual_entries"] + layout["stream_entries"]))
    return s


# ----------------------------------------------------------------------------------------------- main arm
def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(1, a.steps), max(0, a.warmup)

    if a.impl == "reference":
        if rank != 0:
            return 0
        from telescope_b200.synthetic import synth_row_nnz
        full_nnz = int(synth_row_nnz(a.reads, a.loci, a.avg, a.skew, SEED).sum())
        rows = cpu_sample_rows(a, W + K, a.cpu_seconds)
        r = run_cpu(a, W, K, rows, full_nnz)
        n = max(1, a.gpus)
        alg = full_nnz * 12.0 / n + (a.reads // n + 1) * 4.0
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 / r["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(a, n, full_nnz, alg),
            "cpu_baseline": {k: v for k, v in r.items() if not k.startswith("_")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    from telescope_b200 import _abi
    from telescope_b200.likelihood import TelescopeLikelihood
    from telescope_b200.synthetic import shard_bounds, synth_csr
    from telescope_b200 import dist as tsc_dist
    import scipy.sparse as sp

    # one process per GPU: the ranks swap their exchange-buffer handles (or the NCCL id) through files shared by the
    # launcher's children; barriers and max-over-ranks go over the library's own transport (no torch in the workers)
    live = {"tl": None}

    def _reduce(x, op):
        if world == 1 or live["tl"] is None:
            return x
        return float(live["tl"].allreduce([x], op)[0])

    def barrier():
        _reduce(0.0, "sum")

    def max_over_ranks(x):
        return _reduce(x, "max")

    def sum_over_ranks(x):
        return _reduce(x, "sum")

    def model(m, n_iter, max_score, **extra):
        """A model over this rank's block `m`, joined with the other ranks' (fresh rendezvous per model)."""
        kw = dict(devices=[local_rank], dist=tsc_dist.rendezvous(transport=a.transport), max_score=max_score,
                  kernel=a.kernel, replicas=a.replicas, permute_columns=a.permute)
        kw.update(extra)
        return TelescopeLikelihood(m, Opts(n_iter), **kw)

    def gen_block(skew, row_lo, row_hi, pinned):
        ip, ix, raw = synth_csr(a.reads, a.loci, a.avg, skew, SEED, row_lo, row_hi)
        keep = None
        if pinned:
            keep = (_abi.PinnedArray(ix.shape, np.int32), _abi.PinnedArray(raw.shape, np.uint16))
            keep[0].array[:] = ix
            keep[1].array[:] = raw
            ix, raw = keep[0].array, keep[1].array
        if ip[-1] < 2 ** 31:
            ip = ip.astype(np.int32)      # scipy keeps int32 indices only next to an int32 indptr
        return sp.csr_matrix((raw, ix, ip), shape=(row_hi - row_lo, a.loci), copy=False), keep

    # the generator's score range is fixed (best hits reach 211 in any block of more than a few thousand reads), so
    # the global maximum needs no exchange before the transport exists; it is re-checked after construction
    GLOBAL_MAX = 211

    # ---- this rank's block of reads, generated straight into page-locked host memory
    lo, hi = shard_bounds(a.reads, world)[rank]
    t_gen = time.perf_counter()
    m, pins = gen_block(a.skew, lo, hi, True)
    t_gen = time.perf_counter() - t_gen
    local_nnz = int(m.nnz)
    max_score = GLOBAL_MAX if world > 1 else int(m.data.max())

    # one-off library warm-up (CUDA module load, first allocations) on a toy matrix, as any long-lived caller has
    wi, wx, wr = synth_csr(4096, 64, 6, False, 1)
    warm = TelescopeLikelihood(sp.csr_matrix((wr, wx, wi), shape=(4096, 64)), Opts(2), devices=[local_rank])
    warm.em()
    warm.close()

    def timed_e2e(mat, tag):
        """The whole job through the public class, host buffers in, parameters out.  All ranks start together."""
        tsc_dist.file_barrier(tag)
        t0 = time.perf_counter()
        tl = model(mat, K, max_score)
        live["tl"] = tl
        t_create = time.perf_counter() - t0
        tl.em()
        _ = tl.pi.copy(), tl.theta.copy()
        t = max_over_ranks(time.perf_counter() - t0)
        return tl, t, t_create

    # three whole jobs back to back (each: fresh model from the host arrays, K iterations, parameters to the host); the
    # line reports the median run and lists all three
    e2e_runs = []
    for rep in range(3):
        tl, t_e2e, t_create = timed_e2e(m, "e2e%d" % rep)
        e2e_runs.append((t_e2e, t_create, tl.create_seconds, tl.create_laps))
        if rep < 2:
            live["tl"] = None
            tl.close()
    t_e2e, t_create, create_s, create_laps = sorted(e2e_runs)[1]
    assert int(max_over_ranks(float(m.data.max()))) == max_score, "score range assumption violated"
    total_nnz = int(sum_over_ranks(float(local_nnz)))
    c_e2e = tl.counters()

    # ---- device-resident: W warm-up iterations, then exactly K timed ones.  Clocks are sampled from the warm-up to
    # the end of the extra kernel timings below (the GPU is under the same kind of load throughout).
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if W > 0:
        tl.max_iter = W
        tl.em()
    tl.max_iter = K
    c0 = tl.counters()
    barrier()
    t0 = time.perf_counter()
    tl.em()
    wall = time.perf_counter() - t0
    dev_ms = max_over_ranks(tl.em_device_ms())
    wall = max_over_ranks(wall)
    barrier()
    c1 = tl.counters()
    kms = tl.kernel_times_ms()
    kern_ms = max_over_ranks(float(np.mean(kms)) if len(kms) else float("nan"))
    tms = tl.tail_times_ms()
    tail_ms = max_over_ranks(float(np.mean(tms)) if len(tms) else float("nan"))
    final_lnl = tl.lnl

    # other device passes of the path, timed alone (no host copies): standalone E-step (writes z), log-likelihood,
    # one reassign mode; then the one-pass report through the public method (device pass + K-vectors to host)
    passes = {}
    for name in ("estep", "lnl", "reassign"):
        try:
            passes[name] = max_over_ranks(tl.time_pass(name, 3))
        except Exception:             # e.g. not enough memory for the 8 B/entry z buffer
            passes[name] = None
    barrier()
    t0 = time.perf_counter()
    tl.report_colsums(0.9, "exclude")
    t_report = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    layout = tl.layout_stats()
    value = K / (dev_ms * 1e-3)
    peak, peak_src = hbm_peak()
    rows_local = hi - lo
    alg_bytes = local_nnz * 12.0 + (rows_local + 1) * 4.0
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = ncu_traffic(local_nnz, a.skew)
    live["tl"] = None
    tl.close()

    line = None
    if rank == 0:
        # (the same expression as the reference arm, so that both arms print an identical `config`)
        cfg = base_config(a, world, total_nnz, total_nnz * 12.0 / world + (a.reads // world + 1) * 4.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "notes": {
                "timing": "CUDA events on the library stream around em(max_iter=K), max over ranks; includes the final "
                          "calculate_lnl pass like the reference's em()",
                "kernel": a.kernel, "transport": a.transport if world > 1 else "none (1 GPU)",
                "wall_ms_per_step": wall * 1e3 / K, "gen_s": round(t_gen, 2),
                "tail_ms_per_step": tail_ms, "tail": "replica sum + exchange between the GPUs + update + loop control (k_tail)",
            },
            "clocks": clocks,
            "e2e": {
                "value": K / t_e2e, "unit": UNIT,
                "h2d_bytes_per_step": c_e2e["h2d_bytes"] / K, "d2h_bytes_per_step": c_e2e["d2h_bytes"] / K,
                "what": "TelescopeLikelihood(host CSR in pinned memory) + em(%d) + pi/theta to host, wall clock, max over ranks; "
                        "construction %.3f s; the library was warmed up once on a 4096-read toy matrix before the timer "
                        "(CUDA module load), nothing of the workload is cached" % (K, t_create),
                "construction_s": t_create, "tsc_create_s": create_s, "tsc_create_laps_ms": create_laps,
                "runs": [K / r[0] for r in e2e_runs],
                "runs_what": "three whole jobs back to back in this process, value = the median run; the first one allocates its "
                             "device memory from the driver, the later ones get the blocks the destroyed model handed back to "
                             "the library's cache (no data is cached: every job uploads and rebuilds everything)",
                "runs_laps_ms": [r[3] for r in e2e_runs],
            },
            "e2e_report": {
                "value": K / (t_e2e + t_report), "unit": UNIT, "report_s": t_report,
                "what": "e2e plus Telescope.output_report's seven column sums in one device pass (model.py:432-458)",
            },
            "gpu_launches": c1["launches"] - c0["launches"],
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": fused_kernel_name(layout),
                "kernel_ms": kern_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "stream_bytes": layout["stream_bytes"],
                "frac_dram": None if not traffic else traffic / (kern_ms * 1e-3) / 1e9 / peak,
                "note": "achieved/frac use the ALGORITHMIC bytes of canonical CSR (12 B per entry, SURVEY 8d); the kernel "
                        "itself moves `traffic` bytes (1-byte loci, unique reads skipped), frac_dram is on those",
            },
            "final_lnl": final_lnl,
            "layout": layout,
            "other_kernels": {
                "estep_z": None if not passes.get("estep") else {
                    "ms": passes["estep"], "algorithmic_bytes": local_nnz * 20.0 + (rows_local + 1) * 4.0,
                    "achieved_gbs": (local_nnz * 20.0 + (rows_local + 1) * 4.0) / (passes["estep"] * 1e-3) / 1e9,
                    "frac": (local_nnz * 20.0 + (rows_local + 1) * 4.0) / (passes["estep"] * 1e-3) / 1e9 / peak,
                    "what": "k_tiles<TILE_Z>: E-step alone, reads Q+locus, writes z (nnz*20 + (N+1)*4 bytes, SURVEY 8d)"},
                "lnl_ms": passes.get("lnl"), "reassign_exclude_ms": passes.get("reassign"),
            },
        }

    if not a.no_extras:
        # ---- e2e again from ordinary pageable arrays (what a scipy caller hands over)
        mp_ = sp.csr_matrix((np.array(m.data), np.array(m.indices), np.array(m.indptr)), shape=m.shape, copy=False)
        tl2, t_pg, t_pg_create = timed_e2e(mp_, "e2e_pageable")
        if rank == 0:
            line["e2e_pageable"] = {"value": K / t_pg, "unit": UNIT, "construction_s": t_pg_create,
                                    "tsc_create_laps_ms": tl2.create_laps,
                                    "what": "as e2e, but the CSR arrays are ordinary pageable numpy arrays"}
        live["tl"] = None
        tl2.close()
        del mp_

    # ---- parity of the N-rank path against the CPU oracle on a sample of every rank's block
    def parity(skew, rows_per_rank, iters, tag):
        from oracle.em_numpy import EMOracle
        tsc_dist.file_barrier("parity_" + tag)
        bounds = shard_bounds(a.reads, world)
        s_rows = [min(rows_per_rank, b[1] - b[0]) for b in bounds]
        sub, _ = gen_block(skew, lo, lo + s_rows[rank], False)
        g = model(sub, iters, GLOBAL_MAX)
        g.em()
        counts = g.reassign_colsum("exclude")
        out = None
        if rank == 0:
            parts = [synth_csr(a.reads, a.loci, a.avg, skew, SEED, b[0], b[0] + s) for b, s in zip(bounds, s_rows)]
            offs = np.cumsum([0] + [int(p[0][-1]) for p in parts[:-1]])
            ip = np.concatenate([np.zeros(1, dtype=np.int64)] + [p[0][1:] + off for p, off in zip(parts, offs)])
            ix, raw = np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])
            assert int(raw.max()) == GLOBAL_MAX
            o = EMOracle(ip, ix, raw, a.loci, -1.0, iters, 0, 200000).em()
            nz = o.pi > 0
            out = {
                "sample_rows": int(sum(s_rows)), "sample_entries": int(ix.size), "ranks": world, "iterations": iters,
                "lnl_rel_err": abs(g.lnl - o.lnl) / abs(o.lnl),
                "pi_max_rel_err": float(np.max(np.abs(g.pi[nz] - o.pi[nz]) / o.pi[nz])),
                "theta_max_rel_err": float(np.max(np.abs(g.theta - o.theta) / o.theta)),
                "counts_bit_exact": bool(np.array_equal(counts, o.reassign_colsum("exclude"))),
                "counts_total": int(counts.sum()), "tolerance": 1e-6,
                "oracle": "oracle/em_numpy.py on the concatenated sample (rank 0), same seeded rows",
            }
        g.close()
        return out

    if not a.no_extras:
        per_rank = cpu_sample_rows(a) if world == 1 else 250_000
        p = parity(a.skew, per_rank, 3, "main")
        if rank == 0:
            line["parity"] = p

    # ---- BASELINE.json config 5: Zipf reads-per-row (max 200), the divergence / long-read stress
    if not a.no_extras and not a.skew:
        del m, pins
        mz, pz = gen_block(True, lo, hi, True)
        tsc_dist.file_barrier("zipf")
        tz = model(mz, 2, GLOBAL_MAX if world > 1 else int(mz.data.max()))
        live["tl"] = tz
        tz.em()
        tz.max_iter = 8
        barrier()
        tz.em()
        z_ms = max_over_ranks(tz.em_device_ms())
        zk = tz.kernel_times_ms()
        z_kern = max_over_ranks(float(np.mean(zk)))
        z_nnz, z_rows = int(mz.nnz), hi - lo
        z_total = int(sum_over_ranks(float(z_nnz)))
        z_layout = tz.layout_stats()
        z_lnl = tz.lnl
        live["tl"] = None
        tz.close()
        del mz, pz
        pzr = parity(True, 250_000, 2, "zipf")
        if rank == 0:
            zb = z_nnz * 12.0 + (z_rows + 1) * 4.0
            line["other_configs"] = {"zipf": {
                "workload": workload_name(a, True), "nnz": z_total, "steps": 8, "value": 8 / (z_ms * 1e-3), "unit": UNIT,
                "ms_per_step": z_ms / 8, "fused_kernel_ms": z_kern, "algorithmic_bytes": zb,
                "achieved_gbs": zb / (z_kern * 1e-3) / 1e9, "frac": zb / (z_kern * 1e-3) / 1e9 / peak,
                "kernel": fused_kernel_name(z_layout), "layout": z_layout, "final_lnl": z_lnl, "parity": pzr,
            }}

    # ---- CPU arm beside it (rank 0, single-GPU run only): the reference on the same 4 % sample the parity used
    if rank == 0 and world == 1 and not a.no_cpu:
        rows = cpu_sample_rows(a)
        cb = run_cpu(a, 0, 2, rows, total_nnz)
        line["cpu_baseline"] = {k: v for k, v in cb.items() if not k.startswith("_")}
        # and the GPU path against THAT run: same rows, same two iterations
        sub, _ = gen_block(a.skew, 0, rows, False)
        g = TelescopeLikelihood(sub, Opts(2), devices=[local_rank], kernel=a.kernel)
        g.em()
        nz = cb["_pi"] > 0
        line["parity_vs_cpu_baseline"] = {
            "kind": cb["kind"], "sample_rows": rows, "iterations": 2,
            "lnl_rel_err": abs(g.lnl - cb["_lnl"]) / abs(cb["_lnl"]),
            "pi_max_rel_err": float(np.max(np.abs(g.pi[nz] - cb["_pi"][nz]) / cb["_pi"][nz])),
            "theta_max_rel_err": float(np.max(np.abs(g.theta - cb["_theta"]) / cb["_theta"])),
            "tolerance": 1e-6,
        }
        g.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        tsc_dist.file_barrier("end")
        tsc_dist.cleanup()
    return 0


if __name__ == "__main__":
    sys.exit(main())
