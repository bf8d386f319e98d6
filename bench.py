#!/usr/bin/env python
"""bench.py -- polarized SOS spectral solves/sec on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[2], an O2 A-band-like CKD absorption band on synthetic
atmospheres: P spectral points per GPU, 1-25 CKD terms each (ragged NT 101..600), N=41 angles (40 Gauss +
solar), OS_NB=80, Lambert surface; every term is one SOS/SOS_OS term-solve, terms are CKD-summed per point and
the band sums are synthesised on 13 azimuths (SOS_TRPHI_OPTION view mode 2) = one band-solve.
configs[0]/[1] (single wavelength, 5 terms) are parity-test cases (tests/test_gpu_parity.py), too small to bench.

One "step" = one pass of the hot path over the whole per-rank batch.  value = spectral points solved per second
with inputs resident in HBM; e2e = the same through the host-buffer API (H2D of all inputs, D2H of the
CKD-summed Fourier coefficients inside the timed region).  N>1: terms are sharded round-robin along the CKD-term
axis (weak scaling: P points per GPU), one NCCL reduce forms the band sums on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "radiativetransfer-sos_b200"

POINTS_PER_GPU = int(os.environ.get("SOS_BENCH_POINTS", "96"))
NB_GAUSS, OS_NB = 40, 80
METRIC = "polarized SOS spectral solves/sec"
# DRAM traffic of one full-width launch of the hot kernel on this workload, from the ncu --set full capture under profiles/
NCU_TRAFFIC_BYTES = 5.900181e9 + 3.017455e9
NCU_TRAFFIC_SOURCE = ("dram__bytes_read.sum + dram__bytes_write.sum of one full-width k_sweep launch (148 persistent CTAs, 4 960 "
                      "items, 6.3 ms under ncu) from the committed round-2 capture profiles/r2_ksweep_ncu_full_summary.txt; not "
                      "measured live.  The field itself is 16 B per element and launch (read + write); the rest are the "
                      "per-(layer, angle) weight tables (3 x 8 B per element) and the second read of the field by the other "
                      "direction's work unit")
UNIT = "spectral points/s"


def make_workload(npoints):
    pkg = importlib.import_module(PKG)
    return pkg.synth.config_ckd_band(npoints=npoints, seed=20261021, nb_gauss=NB_GAUSS, os_nb=OS_NB,
                                     surface="lambert", rho=0.1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None):
        """Samples taken at or after t_begin (the start of the timed region).  nvidia-smi needs a few hundred ms to deliver
        its first line, so the sampler is started during the warm-up steps (same kernels, same load); when the timed region is
        too short to hold a sample of its own, the samples of the last warm-up steps stand in and `window` says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if t_begin is None or t >= t_begin]
        window = "timed region"
        if not inside:
            inside, window = [r for _, r in self.rows[-8:]], "last warm-up steps (timed region shorter than one nvidia-smi sample)"
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_baseline(wl, budget_s=20.0, max_threads=None):
    """Times the CPU oracle (C restatement of the reference loops; gfortran is unavailable) on a bounded sample
    of the same workload, one thread per host core (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    from util import oracle_term
    orc.lib()
    cores = max_threads or os.cpu_count() or 1
    # probe one term to size the sample
    t0 = time.perf_counter()
    oracle_term(orc, wl.optics[wl.terms[0].optics], wl.terms[0])
    one = max(time.perf_counter() - t0, 1e-3)
    nterm_target = int(max(cores, min(len(wl.terms), budget_s * cores / one)))
    # whole spectral points only
    pts, ids = [], []
    for i, t in enumerate(wl.terms):
        if t.optics not in pts:
            if len(ids) >= nterm_target:
                break
            pts.append(t.optics)
        ids.append(i)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda i: oracle_term(orc, wl.optics[wl.terms[i].optics], wl.terms[i]), ids))
    dt = time.perf_counter() - t0
    return {"value": len(pts) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d spectral points (%d term-solves) of the same band in %.1f s; C restatement of the "
                      "reference loops (gcc -O2, no FMA), gfortran unavailable" % (len(pts), len(ids), dt),
            "term_solves_per_s": len(ids) / dt}


def cpu_baseline_reference(wl, points_per_core=3):
    """Times the reference's own flow (SOS -> SOS_OS -> SOS_AGGREGATE, with its files) from oracle/_ref/libsosref.so =
    the reference's Fortran sources translated to C by oracle/f77_to_c.py (no Fortran compiler exists here), on a
    bounded sample of whole spectral points.  One process per host core; TERM-solves are distributed over the processes
    (longest first), the workers are warm before the clock starts.  value = wall-clock rate;
    the occupancy rate (cores x points / sum of busy seconds) is reported beside it."""
    from oracle import ref_runner
    if not ref_runner.available():
        raise RuntimeError("oracle/_ref/libsosref.so is missing")
    cores = os.cpu_count() or 1
    npts = len({t.optics for t in wl.terms})
    ids = set(range(min(npts, max(1, cores * points_per_core))))
    r = ref_runner.run_points(wl, ids, cores)
    wall_rate = r["points"] / r["wall"]
    occ_rate = r["workers"] * r["points"] / r["busy_sum"]
    return {"value": wall_rate, "unit": UNIT, "cores": r["workers"], "kind": "reference",
            "sample": "%d spectral points (%d term-solves) of the same band in %.1f s wall through the reference's per-term flow "
                      "(PROFIL_TMP -> SOS -> SOS_OS, then SOS_AGGREGATE, files included); term-solves spread over %d "
                      "single-thread processes (longest first), workers warm before the clock; wall-based %.3f points/s, "
                      "occupancy-based (cores x points / sum of busy seconds) %.3f points/s; reference Fortran translated "
                      "to C by oracle/f77_to_c.py, gcc -O2 (no Fortran compiler available)"
                      % (r["points"], r["terms"], r["wall"], r["workers"], wall_rate, occ_rate),
            "term_solves_per_s": r["terms"] / r["wall"], "occupancy_value": occ_rate, "wall_s": r["wall"]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = make_workload(POINTS_PER_GPU)
    vals = []
    total = args.warmup + args.steps
    ppc = 3                                          # spectral points per host core in one step's sample
    for s in range(total):
        t0 = time.perf_counter()
        try:
            r = cpu_baseline_reference(wl, points_per_core=ppc)
        except Exception as e:                       # no oracle/_ref on this box: time the (bit-identical) port instead
            print("reference library unavailable (%r): timing the oracle port" % (e,), file=sys.stderr)
            r = cpu_baseline(wl, budget_s=8.0)
        if s == 0:                                   # keep the whole run within ~10 minutes on a slow box: smaller samples
            dt = time.perf_counter() - t0
            ppc = int(min(3, max(1, 3 * 600.0 / max(dt * total, 1e-9))))
        if s >= args.warmup:
            vals.append(r)
    v = float(np.mean([x["value"] for x in vals]))
    walls = [x["wall_s"] for x in vals if "wall_s" in x]
    r = vals[-1]
    r["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (1e3 * float(np.mean(walls)) if walls else None), "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "O2 A-band-like CKD band (BASELINE configs[2]), bounded sample per step",
                       "nb_gauss": NB_GAUSS, "os_nb": OS_NB},
            "cpu_baseline": r,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(line))


def dgemm_peak(torch, dev):
    """cuBLAS DGEMM TFLOP/s (the FP64 roofline denominator; MEASURED_PEAKS.json has no FP64 entry)."""
    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


_STDOUT_FD = None


def _emit(text):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(text, flush=True)


def lpt_partition(costs, nparts):
    """Longest-processing-time-first partition of whole spectral points: returns part index per point."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(nparts)
    part = np.zeros(len(costs), dtype=np.int64)
    for p in order:
        k = int(np.argmin(load))
        part[p] = k
        load[k] += costs[p]
    return part


def point_costs(wl):
    """Cost proxy of a spectral point: sum over its CKD terms of levels x (1 + scattering optical depth)."""
    c = np.zeros(len(wl.optics))
    for t in wl.terms:
        dh = np.diff(np.asarray(t.h))
        tau_scat = float(np.sum(dh * (np.asarray(t.pcaer)[1:] + np.asarray(t.pcmol)[1:])))
        c[t.optics] += (t.nt + 1) * (1.0 + tau_scat)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--band-points", type=int, default=0,
                    help="strong scaling: a FIXED band of this many spectral points shared by all GPUs (default: weak "
                         "scaling, %d points per GPU)" % POINTS_PER_GPU)
    ap.add_argument("--shard", default=os.environ.get("SOS_BENCH_SHARD", "wavelengths"), choices=["wavelengths", "terms"],
                    help="wavelengths: whole spectral points per GPU (LPT), no reduce, one NCCL gather of the result tables; "
                         "terms: CKD terms round-robin, one NCCL reduce of the partial band sums")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # Library banners (e.g. "NCCL version ...") go to fd 1: keep stdout for the ONE JSON line by pointing fd 1 at
    # stderr while working and restoring it for the final print.
    sys.stdout.flush()
    global _STDOUT_FD
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    api = importlib.import_module(PKG + ".api")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the SOS hot path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    solver = api.Solver(local)
    if world > 1:
        # torch.distributed is launch plumbing only (barrier, max over ranks, broadcast of the 128-byte NCCL id);
        # the data path uses the communicator owned by libsosgpu.so
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        uid = torch.zeros(api.Solver.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(solver.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, src=0)
        solver.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))

    npoints_total = args.band_points if args.band_points > 0 else POINTS_PER_GPU * world
    scaling = "strong" if args.band_points > 0 else "weak"
    wl = make_workload(npoints_total)
    nterms_total = len(wl.terms)
    if args.shard == "wavelengths":
        part = lpt_partition(point_costs(wl), world)
        my_points = [p for p in range(npoints_total) if part[p] == rank]
        local_group = {p: g for g, p in enumerate(my_points)}
        my_ids = [i for i, t in enumerate(wl.terms) if part[t.optics] == rank]
        groups = [local_group[wl.terms[i].optics] for i in my_ids]
        ngroup = len(my_points)
        groups_of_rank = [int(np.sum(part == r)) for r in range(world)]
    else:
        my_ids = list(range(rank, nterms_total, world))
        groups = [wl.terms[i].optics for i in my_ids]
        ngroup = npoints_total
        groups_of_rank = None
    batch = solver.upload(wl, my_ids, groups=groups, ngroup=ngroup)
    NPHI = 13                                          # SOS_TRPHI_OPTION view mode 2, dphi = 30

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def finish(b, download):
        """CKD band sums -> azimuth synthesis -> complete band result on rank 0."""
        if args.shard == "wavelengths":
            solver.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=False)
            return solver.gather_tables(b, groups_of_rank, NPHI, root=0, download=download)
        solver.reduce_groups(b, root=0)                # ONE ncclReduce of the partial band sums (+ group metadata)
        if rank == 0:
            _, up_t, down_t = solver.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=download)
            return up_t, down_t
        return None, None

    def step_resident():
        solver.run(batch, want_terms=False, want_groups=False, part_only=args.shard == "terms" and world > 1)
        finish(batch, False)

    sampler = ClockSampler(local)
    for w in range(args.warmup):
        if rank == 0 and w == max(0, args.warmup - 2):     # running before the timed region starts (see ClockSampler.stop)
            sampler.start()
        step_resident()
    sync_all()
    if rank == 0 and sampler.proc is None:
        sampler.start()
    l0 = solver.launches
    st_acc = {"flops": 0.0, "step_ms": 0.0, "step_launches": 0, "steps": 0, "bytes": 0.0, "total_ms": 0.0, "useful_flops": 0.0}
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
        s = solver.stats(batch)
        for k in st_acc:
            st_acc[k] += s[k]
    sync_all()
    dt = time.perf_counter() - t0
    launches = solver.launches - l0
    # device time of the steps (CUDA events inside the library, on its stream), max over ranks
    dev_ms = torch.tensor([st_acc["total_ms"], dt * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(dev_ms[0]), float(dev_ms[1])
    # the barrier-to-barrier wall clock bounds the device time from above; report the conservative one
    step_ms = max(dev_ms, wall_ms) / args.steps
    clocks = sampler.stop(t0) if rank == 0 else None

    # ---- e2e: host buffers in, host buffers out, every step ----
    e2e_steps = max(1, min(args.steps, 3))
    h2d = d2h = 0
    t0 = 0.0
    verbose = bool(os.environ.get("SOS_BENCH_VERBOSE"))
    for e2e_it in range(-1, e2e_steps):            # iteration -1 is an untimed warm-up of the host-buffer path
        if e2e_it == 0:
            sync_all()
            t0 = time.perf_counter()
            h2d = d2h = 0
        ta = time.perf_counter()
        b2 = solver.upload(wl, my_ids, groups=groups, ngroup=ngroup)
        tb = time.perf_counter()
        if args.shard == "wavelengths":            # every rank brings back the CKD-summed Fourier coefficients of its wavelengths
            tr, gr = solver.run(b2, want_terms=True, want_groups=True, want_rec=False)
            d2h += gr.rec.nbytes
        else:
            tr, _ = solver.run(b2, want_terms=True, want_groups=False, want_rec=False, part_only=world > 1)
        tc = time.perf_counter()
        up_t, down_t = finish(b2, True)
        if args.shard == "terms" and rank == 0:
            gr = solver.groups(b2)                 # the REDUCED band sums
            d2h += gr.rec.nbytes
        if up_t is not None:
            d2h += up_t.nbytes + down_t.nbytes
        h2d += b2.h2d_bytes
        d2h += tr.n_fourier.nbytes + tr.n_scatter.nbytes
        td = time.perf_counter()
        b2.free()
        if verbose:
            print("rank %d e2e: upload %.1f ms, run+download %.1f ms, finish %.1f ms, free %.1f ms"
                  % (rank, (tb - ta) * 1e3, (tc - tb) * 1e3, (td - tc) * 1e3, (time.perf_counter() - td) * 1e3), file=sys.stderr)
    sync_all()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_dt, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = e2e_t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.SUM)
        e2e_dt, h2d, d2h = float(tmax[0]), float(e2e_t[1]), float(e2e_t[2])
    flops_t = torch.tensor([st_acc["flops"], st_acc["step_ms"]], dtype=torch.float64, device=dev)

    if rank == 0:
        peak = dgemm_peak(torch, dev)
        achieved = st_acc["flops"] / (st_acc["step_ms"] * 1e-3) / 1e12 if st_acc["step_ms"] > 0 else 0.0
        if os.environ.get("SOS_BENCH_NOCPU"):              # kernel-tuning runs: skip the CPU leg
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "skipped (SOS_BENCH_NOCPU)"}
        elif world == 1:                               # the CPU leg is timed at N=1 only
            try:
                cb = cpu_baseline_reference(make_workload(POINTS_PER_GPU))
            except Exception as e1:
                try:
                    cb = cpu_baseline(make_workload(POINTS_PER_GPU), budget_s=15.0)
                    cb["sample"] += " (oracle/_ref unavailable: %r)" % (e1,)
                except Exception as e:  # pragma: no cover
                    cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        else:
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "timed at N=1 only"}
        shard_txt = {"wavelengths": "whole spectral points per GPU (LPT on levels x (1 + scattering depth)), no reduce; one NCCL "
                                    "send/recv gather of the synthesised tables to rank 0 (communicator owned by libsosgpu.so)",
                     "terms": "CKD-term axis round-robin, ONE ncclReduce of the partial band sums + group metadata to rank 0 "
                              "(communicator owned by libsosgpu.so)"}[args.shard]
        line = {
            "metric": METRIC, "value": npoints_total / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "O2 A-band-like CKD band (BASELINE configs[2]): %d spectral points (%s), %d CKD "
                                   "term-solves total, N=41 angles, OS_NB=80, NT 101..600, Lambert rho=0.1"
                                   % (npoints_total, "%d per GPU" % POINTS_PER_GPU if scaling == "weak" else "fixed band",
                                      nterms_total),
                       "points_per_gpu": npoints_total / world, "term_solves": nterms_total,
                       "term_solves_per_s": nterms_total / (step_ms * 1e-3),
                       "sharding": shard_txt if world > 1 else "none",
                       "cache": "field working set (GBs) >> 126 MB L2; no reuse between steps"},
            "e2e": {"value": npoints_total / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d / e2e_steps),
                    "d2h_bytes_per_step": int(d2h / e2e_steps)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "k_sweep (FP64 DMMA source-function contraction fused with the "
                         "layer recurrence; persistent, warp-specialised)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": NCU_TRAFFIC_BYTES,
                         "traffic_source": NCU_TRAFFIC_SOURCE,
                         "peak_source": "cuBLAS DGEMM 6144^3 measured in this run (MEASURED_PEAKS.json has no FP64 "
                                        "entry)",
                         "algorithmic": "2*(6N)^2*(NT+1) FLOP per (term, Fourier order, scattering order>=2) launched; a wave "
                                        "solves 8 Fourier orders at once, useful_frac is the share spent on orders the terms keep "
                                        "(s < n_fourier), the rest is computed past the Fourier stop and discarded",
                         "useful_frac": st_acc["useful_flops"] / st_acc["flops"] if st_acc["flops"] else None,
                         "kernel_ms_per_step": st_acc["step_ms"] / args.steps,
                         "kernel_launches_per_step": st_acc["step_launches"] / args.steps,
                         "kernel_share_of_step": st_acc["step_ms"] / max(st_acc["total_ms"], 1e-9),
                         "hbm_recurrence_GBs": st_acc["bytes"] / (st_acc["step_ms"] * 1e-3) / 1e9 if st_acc["step_ms"] else 0},
            "cpu_baseline": cb,
        }
        _emit(json.dumps(line))
    del flops_t
    batch.free()
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
