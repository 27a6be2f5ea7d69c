#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- EM iterations/sec of Telescope's reassignment loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload: the 50 M-read x 30 k-locus synthetic CSR (avg 20 alignments/read, ~1e9 entries, fp64) that
BASELINE.json's metric is quoted on; with N GPUs the same matrix is row-sharded (strong scaling).  A "step" is one
EM iteration (E-step + M-step + parameter update) over the whole matrix; em_epsilon = -1 keeps the loop from
stopping early, exactly as on the CPU arm.

  value  = K / device time of `em(max_iter=K)` (CUDA events on the library's stream, max over ranks), inputs
           resident in HBM.  Like the reference's em(), the call ends with one calculate_lnl pass (model.py:800-801).
  e2e    = K / wall time of the whole job through the public class with HOST (pinned) CSR buffers:
           TelescopeLikelihood(csr, opts) [H2D of the CSR, Q build, tiling] + em(K) + D2H of pi/theta.
  roofline = fused E+M kernel: algorithmic bytes nnz*12 + (N+1)*4 / mean per-iteration kernel time
           measured with CUDA events inside tsc_em, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline = the scipy.sparse port of the reference loop (oracle/em_scipy.py; op-for-op the reference's calls,
           single-threaded like scipy) on the first rows of the same matrix, scaled linearly in nnz to the full
           workload (the reference needs ~100 B/nnz of host RAM; linearity measured in BASELINE.md).

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
    ap.add_argument("--skew", action="store_true", help="config 5: Zipf reads-per-row (max 200)")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--replicas", type=int, default=0)
    ap.add_argument("--smem-table-cols", type=int, default=-1)
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="CPU baseline budget")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--permute", action="store_true", help="renumber loci by descending entry count on the device")
    return ap.parse_args()


def workload_name(a):
    return "synthetic CSR %s reads x %s loci, avg %d alignments/read%s, seed %d" % (
        _si(a.reads), _si(a.loci), a.avg, ", Zipf rows (max 200)" if a.skew else "", SEED)


def _si(n):
    for div, suf in ((10 ** 6, "M"), (10 ** 3, "k")):
        if n >= div and n % div == 0:
            return "%d%s" % (n // div, suf)
    return str(n)


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_sample_rows(a, n_iters):
    """Rows of the workload the CPU arm can finish in the budget (reference: ~153 ns per nnz per iteration, +init)."""
    per_row = 160e-9 * a.avg * (n_iters + 1.5)
    return int(max(20_000, min(a.reads, a.cpu_seconds / per_row)))


def run_cpu(a, n_warm, n_iters, csr=None):
    """Time the scipy port on a bounded sample; returns dict(value=iter/s scaled to the full workload, ...)."""
    import scipy.sparse as sp
    from oracle.em_scipy import ScipyEM
    from telescope_b200.synthetic import synth_csr
    rows = cpu_sample_rows(a, n_warm + n_iters)
    if csr is None:
        ip, ix, raw = synth_csr(a.reads, a.loci, a.avg, a.skew, SEED, 0, rows)
    else:
        ip, ix, raw = csr
        ip = ip[:rows + 1]
        ix, raw = ix[:ip[-1]], raw[:ip[-1]]
    m = sp.csr_matrix((np.asarray(raw), np.asarray(ix), np.asarray(ip)), shape=(rows, a.loci))
    full_nnz = a.reads * float(m.nnz) / rows          # the generator's rows are i.i.d.
    em = ScipyEM(m, em_epsilon=-1.0, max_iter=max(1, n_warm))
    if n_warm > 0:
        em.em()
    em.max_iter = n_iters
    t0 = time.perf_counter()
    em.em()
    dt = time.perf_counter() - t0
    sample_ips = n_iters / dt
    return {
        "value": sample_ips * m.nnz / full_nnz,
        "unit": UNIT,
        "cores": 1,
        "kind": "port",
        "sample": "first %d reads (%d entries, %.3g of the workload), %d timed EM iterations in %.1f s = %.1f ns/entry/iter; "
                  "scaled linearly in entries to the full matrix" % (rows, m.nnz, m.nnz / full_nnz, n_iters, dt,
                                                                     dt / n_iters / m.nnz * 1e9),
        "host_cpus": os.cpu_count(),
        "_lnl": float(em.lnl), "_rows": rows, "_pi": em.pi,
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


def ncu_traffic(nnz):
    """DRAM bytes per launch of the fused kernel from the committed ncu capture, scaled per entry (or None)."""
    p = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        return float(d["dram_bytes_per_entry"]) * nnz
    except Exception:
        return None


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
        r = run_cpu(a, W, K)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 / r["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "n_reads": a.reads, "n_loci": a.loci},
            "cpu_baseline": {k: v for k, v in r.items() if not k.startswith("_")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    from telescope_b200 import _abi
    from telescope_b200.likelihood import TelescopeLikelihood
    from telescope_b200.synthetic import shard_bounds, synth_csr
    import scipy.sparse as sp

    # one process per GPU: the NCCL id travels through a file shared by the launcher's children; barriers and
    # max-over-ranks go over the library's own communicator (no torch in the workers)
    from telescope_b200 import dist as tsc_dist
    dist = tsc_dist.rendezvous()
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

    # ---- this rank's block of reads, generated straight into page-locked host memory
    lo, hi = shard_bounds(a.reads, world)[rank]
    t_gen = time.perf_counter()
    ip, ix, raw = synth_csr(a.reads, a.loci, a.avg, a.skew, SEED, lo, hi)
    pin_ix, pin_raw = _abi.PinnedArray(ix.shape, np.int32), _abi.PinnedArray(raw.shape, np.uint16)
    pin_ix.array[:] = ix
    pin_raw.array[:] = raw
    del ix, raw
    if ip[-1] < 2 ** 31:
        ip = ip.astype(np.int32)      # scipy keeps int32 indices only next to an int32 indptr
    m = sp.csr_matrix((pin_raw.array, pin_ix.array, ip), shape=(hi - lo, a.loci), copy=False)
    t_gen = time.perf_counter() - t_gen
    local_nnz = int(m.nnz)
    # the generator's score range is fixed (best hits reach 211 in any block of more than a few thousand reads), so
    # the global maximum needs no exchange before the communicator exists; it is re-checked after construction
    max_score = 211 if world > 1 else int(pin_raw.array.max())

    kw = dict(devices=[local_rank], dist=dist, max_score=max_score, kernel=a.kernel, replicas=a.replicas,
              smem_table_cols=a.smem_table_cols, permute_columns=a.permute)

    # one-off library warm-up (CUDA module load, first allocations) on a toy matrix, as any long-lived caller has
    wi, wx, wr = synth_csr(4096, 64, 6, False, 1)
    warm = TelescopeLikelihood(sp.csr_matrix((wr, wx, wi), shape=(4096, 64)), Opts(2), devices=[local_rank])
    warm.em()
    warm.close()
    # ---- e2e: the whole job through the public class, host buffers in, parameters out.  All ranks start together
    # (no communicator exists yet, so the barrier goes through the launcher-shared temp files).
    tsc_dist.file_barrier("e2e")
    t0 = time.perf_counter()
    tl = TelescopeLikelihood(m, Opts(K), **kw)
    live["tl"] = tl
    t_create = time.perf_counter() - t0
    create_s, create_laps = tl.create_seconds, tl.create_laps
    tl.em()
    pi_e2e = tl.pi.copy()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    assert int(max_over_ranks(float(pin_raw.array.max()))) == max_score, "score range assumption violated"
    total_nnz = int(sum_over_ranks(float(local_nnz)))
    c_e2e = tl.counters()
    lnl_first = tl.lnl

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

    # other device passes of the path, timed alone (no host copies): standalone E-step (writes z), log-likelihood,
    # one reassign mode
    passes = {}
    for name in ("estep", "lnl", "reassign"):
        try:
            passes[name] = max_over_ranks(tl.time_pass(name, 3))
        except Exception as exc:      # e.g. not enough memory for the 8 B/entry z buffer
            passes[name] = None
    clocks = sampler.stop() if sampler else None
    value = K / (dev_ms * 1e-3)
    peak, peak_src = hbm_peak()
    rows_local = hi - lo
    alg_bytes = local_nnz * 12.0 + (rows_local + 1) * 4.0
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = ncu_traffic(local_nnz)
    layout = tl.layout_stats()
    fused_name = ("k_ell_fused (E-step + M-step accumulation over the clustered slice stream, telescope_b200/csrc/tsc_ell.cuh)"
                  + (" + k_tiles<TILE_FUSED> on the residual CSR" if layout["residual_reads"] else "")
                  if layout["slices"] else
                  "k_tiles<TILE_FUSED> (E-step + M-step accumulation, telescope_b200/csrc/tsc_tiles.cuh)")

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(a), "n_reads": a.reads, "n_loci": a.loci, "nnz": total_nnz,
                "parallelism": "read-sharded x%d, 1 NCCL all-reduce of K doubles per iteration" % world if world > 1 else "1 GPU",
                "l2": "per-iteration inputs (%.1f GB per GPU) are larger than L2, no flush needed" % (alg_bytes / 1e9),
                "timing": "CUDA events on the library stream around em(max_iter=K), max over ranks; includes the final "
                          "calculate_lnl pass like the reference's em()",
                "kernel": a.kernel, "wall_ms_per_step": wall * 1e3 / K, "gen_s": round(t_gen, 2),
            },
            "clocks": clocks,
            "e2e": {
                "value": K / t_e2e, "unit": UNIT,
                "h2d_bytes_per_step": c_e2e["h2d_bytes"] / K, "d2h_bytes_per_step": c_e2e["d2h_bytes"] / K,
                "what": "TelescopeLikelihood(host CSR in pinned memory) + em(%d) + pi/theta to host, wall clock, max over ranks; "
                        "construction %.3f s; the library was warmed up once on a 4096-read toy matrix before the timer "
                        "(CUDA module load), nothing of the workload is cached" % (K, t_create),
                "construction_s": t_create, "tsc_create_s": create_s, "tsc_create_laps_ms": create_laps,
            },
            "gpu_launches": c1["launches"] - c0["launches"],
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": fused_name,
                "kernel_ms": kern_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
            },
            "final_lnl": tl.lnl,
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

    # ---- CPU arm beside it (rank 0, single-GPU run only) + parity of the GPU path on the same sample
    if rank == 0 and world == 1 and not a.no_cpu:
        live["tl"] = None
        tl.close()
        cb = run_cpu(a, 0, 2, csr=(ip, pin_ix.array, pin_raw.array))
        rows = cb["_rows"]
        sub = sp.csr_matrix((pin_raw.array[:ip[rows]], pin_ix.array[:ip[rows]], ip[:rows + 1]), shape=(rows, a.loci))
        o2 = Opts(2)
        g = TelescopeLikelihood(sub, o2, devices=[local_rank], kernel=a.kernel)
        g.em()
        line["parity"] = {
            "sample_rows": rows, "iterations": 2,
            "lnl_rel_err": abs(g.lnl - cb["_lnl"]) / abs(cb["_lnl"]),
            "pi_max_rel_err": float(np.max(np.abs(g.pi - cb["_pi"]) / np.maximum(cb["_pi"], 1e-300))),
            "tolerance": 1e-6,
        }
        g.close()
        line["cpu_baseline"] = {k: v for k, v in cb.items() if not k.startswith("_")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        barrier()
        tsc_dist.cleanup()
    tl.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
