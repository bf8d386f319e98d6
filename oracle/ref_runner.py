"""ref_runner.py -- TEST / BENCH INFRASTRUCTURE.  Runs the reference's own per-term flow on host cores from
oracle/_ref/libsosref.so (the reference's Fortran sources translated by oracle/f77_to_c.py): for every CKD term
PROFIL_TMP is written, SOS reads it and calls SOS_OS, SOS_OS writes the term's Fourier records; SOS_AGGREGATE then folds
them into the wavelength's SOS_Result.bin through its temporary file -- files and all, as SOS_PROC.F:3459-3594 does.
The translated code keeps Fortran's static storage and unit table, so parallelism is one PROCESS per core (spawned),
which is also how one would run the Fortran executable.

Work distribution: TERM-solves, not spectral points, are spread over the worker processes
(longest first onto the least loaded worker, cost proxy NT+1), so a 25-term point does not pin one core while the others
idle.  Workers are started, load the library and signal "ready" BEFORE the clock starts.  Two rates come back: the
wall-clock one (slowest worker + the serial SOS_AGGREGATE pass) and the occupancy one (sum of the workers' busy seconds).

solve_terms() is the same machinery returning every term's Fourier records and scalars: the GPU parity tests compare
the CUDA path with the reference's own statements through it."""
import ctypes as C
import os
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsosref.so")
MX, NBM, NTM = 80, 200, 600


def available():
    return os.path.exists(LIB)


def _fs(s):
    return C.create_string_buffer(s.encode().ljust(500), 500)


def _formats():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sos_formats", os.path.join(os.path.dirname(HERE), "radiativetransfer-sos_b200", "formats.py"))
    fm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fm)
    return fm


_ip = lambda v: C.byref(C.c_int(int(v)))
_dp = lambda v: C.byref(C.c_double(float(v)))
_P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
_L = C.c_size_t(500)


def _pad(v, n):
    a = np.zeros(n)
    a[:len(v)] = v
    return a


def _solve_one(lib, fm, tmp, o, term, surf_path, want_trans=False):
    """SOS (SOS.F:340) of the reference for one term -> dict(rec [nf][3][W], scalars)."""
    aik, z, h, xa, ym = term
    N = o["nbmu"]
    fprof, fos = os.path.join(tmp, "PROFIL_TMP"), os.path.join(tmp, "OS_TMP.bin")
    rmu, ga = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    rmu[MX - N:MX + N + 1], ga[MX - N:MX + N + 1] = o["rmu"], o["ga"]
    al, be, gm, ze = (_pad(o[k], NBM + 1) for k in ("alpha", "beta", "gamma", "zeta"))
    fm.write_profile(fprof, z, h, xa, ym)
    if os.path.exists(fos):
        os.remove(fos)
    sc = [C.c_double(0) for _ in range(6)]
    tdg = np.zeros(2 * MX + 1)
    ier = C.c_int(0)
    lib.sos_(_fs(fos), _fs("TRANS" if want_trans else "NO_OUTPUT"), _fs(fprof), _ip(len(h) - 1), _dp(o["zout"]), _ip(o["igmax"]),
             _ip(o["ipolar"]), _dp(o["ron"]), _dp(o["ind_surf"]), _dp(o["rho"]), _ip(1 if surf_path else 0), _ip(o["ifresnel"]),
             _fs(surf_path or "none"), _ip(o["n0"]), _dp(o["piz"]), _dp(o["piztr"]), _dp(o["a_trunc"]), _P(rmu), _P(ga),
             _dp(o["tetas"]), _ip(o["os_nb"]), _ip(N), _P(al), _P(be), _P(gm), _P(ze), C.byref(sc[0]), C.byref(sc[1]),
             C.byref(sc[2]), C.byref(sc[3]), _P(tdg), C.byref(sc[4]), C.byref(sc[5]), _ip(0), _ip(6), C.byref(ier), _L, _L, _L, _L)
    if ier.value != 0:
        raise RuntimeError("reference SOS returned IER=%d" % ier.value)
    rec = fm.read_result_bin(fos, N)
    return dict(rec=rec, ttot_tronc=sc[0].value, ttot_vrai=sc[1].value, tauout=sc[2].value, tdifmus=sc[3].value,
                emoins=sc[4].value, eplus=sc[5].value, tdifmug=tdg[MX - N:MX + N + 1].copy())


def _worker(job_path, out_path):
    """job = [(term id, optics dict, (aik, zprof, h, pcaer, pcmol), surface array or None), ...]"""
    import pickle
    import sys
    fm = _formats()
    lib = C.CDLL(LIB)
    with open(job_path, "rb") as f:
        job = pickle.load(f)
    print("ready", flush=True)
    sys.stdin.readline()                                         # the parent starts the clock, then says go
    t0 = time.perf_counter()
    out = {}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        surf_cache = {}
        for tid, o, term, keep in job:
            sp = None
            if o.get("surf") is not None:
                key = id(o["surf"])
                if key not in surf_cache:
                    surf_cache[key] = os.path.join(tmp, "SURF%d.bin" % len(surf_cache))
                    fm.write_surface_bin(surf_cache[key], o["surf"])
                sp = surf_cache[key]
            r = _solve_one(lib, fm, tmp, o, term, sp)
            if not keep:
                r["rec"] = np.zeros((r["rec"].shape[0], 0, 0))
            out[tid] = r
    busy = time.perf_counter() - t0
    with open(out_path, "wb") as f:
        pickle.dump((out, busy), f)
    print("done %.6f" % busy, flush=True)


def _optics_dict(o):
    return dict(nbmu=o.nbmu, rmu=np.asarray(o.rmu, dtype=np.float64), ga=np.asarray(o.ga, dtype=np.float64), n0=o.n0, tetas=o.tetas,
                os_nb=o.os_nb, alpha=np.asarray(o.alpha), beta=np.asarray(o.beta), gamma=np.asarray(o.gamma), zeta=np.asarray(o.zeta),
                a_trunc=o.a_trunc, piz=o.piz, piztr=o.piztr, ron=o.ron, rho=o.rho, ifresnel=o.ifresnel, ind_surf=o.ind_surf,
                igmax=o.igmax, ipolar=o.ipolar, zout=o.zout, surf=(np.asarray(o.surf) if o.imat_surf == 1 else None))


def solve_terms(workload, term_ids, cores, keep_records=True, timeout_s=1800):
    """Term-solves of a synth.Workload through the reference's SOS on `cores` worker processes (plain subprocesses of this
    file: no fork of a CUDA process).  Returns ({term id: result dict}, wall seconds of the solve phase, [busy seconds])."""
    import pickle
    import subprocess
    import sys
    term_ids = list(term_ids)
    od = {}
    units = []
    for i in term_ids:
        t = workload.terms[i]
        if id(workload.optics[t.optics]) not in od:
            od[id(workload.optics[t.optics])] = _optics_dict(workload.optics[t.optics])
        units.append((i, od[id(workload.optics[t.optics])],
                      (t.aik, np.asarray(t.zprof), np.asarray(t.h), np.asarray(t.pcaer), np.asarray(t.pcmol)), keep_records))
    units.sort(key=lambda u: -len(u[2][2]))                       # longest profile first ...
    nw = max(1, min(cores, len(units)))
    jobs, load = [[] for _ in range(nw)], [0] * nw
    for u in units:                                              # ... onto the least loaded worker
        k = load.index(min(load))
        jobs[k].append(u)
        load[k] += len(u[2][2])
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for i, job in enumerate(jobs):
            jp, op = os.path.join(tmp, "job%d.pkl" % i), os.path.join(tmp, "out%d.pkl" % i)
            with open(jp, "wb") as f:
                pickle.dump(job, f)
            procs.append((subprocess.Popen([sys.executable, os.path.abspath(__file__), "--job", jp, "--out", op], stdin=subprocess.PIPE,
                                           stdout=subprocess.PIPE, text=True), op))
        results, busy = {}, []
        try:
            for pr, _ in procs:                                  # every worker has imported numpy and loaded the library
                if pr.stdout.readline().strip() != "ready":
                    raise RuntimeError("reference worker failed to start")
            t0 = time.perf_counter()
            for pr, _ in procs:
                pr.stdin.write("go\n")
                pr.stdin.flush()
            deadline = t0 + timeout_s
            for pr, op in procs:
                line = pr.stdout.readline().strip()
                if not line.startswith("done") or time.perf_counter() > deadline:
                    raise RuntimeError("reference worker failed (rc=%r)" % pr.poll())
            wall = time.perf_counter() - t0
            for pr, op in procs:
                pr.wait(timeout=60)
                with open(op, "rb") as f:
                    out, b = pickle.load(f)
                results.update(out)
                busy.append(b)
        finally:
            for pr, _ in procs:
                if pr.poll() is None:
                    pr.kill()
    return results, wall, busy


def aggregate_point(lib, fm, tmp, nbmu, terms):
    """SOS_AGGREGATE of the reference over the terms [(aik, result dict), ...] of one spectral point, through its files.
    Returns (records incl. the reference's trailing zero record, dict of aggregated scalars)."""
    fos, fagg, fres = (os.path.join(tmp, n) for n in ("OS_TMP.bin", "AGG_TMP.bin", "SOS_Result.bin"))
    if os.path.exists(fres):
        os.remove(fres)
    acc = [C.c_double(0) for _ in range(6)]
    tdg = np.zeros(2 * MX + 1)
    for aik, r in terms:
        fm.write_result_bin(fos, r["rec"])
        tdg_tmp = np.zeros(2 * MX + 1)
        tdg_tmp[MX - nbmu:MX + nbmu + 1] = r["tdifmug"]
        ier = C.c_int(0)
        lib.sos_aggregate_(_ip(nbmu), _dp(aik), _fs(fos), _dp(r["ttot_tronc"]), _dp(r["ttot_vrai"]), _dp(r["tauout"]), _dp(r["tdifmus"]),
                           _P(tdg_tmp), _dp(r["emoins"]), _dp(r["eplus"]), _fs(fagg), _fs(fres), C.byref(acc[0]), C.byref(acc[1]),
                           C.byref(acc[2]), C.byref(acc[3]), _P(tdg), C.byref(acc[4]), C.byref(acc[5]), C.byref(ier), _L, _L, _L)
        if ier.value != 0:
            raise RuntimeError("reference SOS_AGGREGATE returned IER=%d" % ier.value)
    return fm.read_result_bin(fres, nbmu), dict(ttot_tronc=acc[0].value, ttot_vrai=acc[1].value, tauout=acc[2].value,
                                                tdifmus=acc[3].value, emoins=acc[4].value, eplus=acc[5].value)


def run_points(workload, point_ids, cores, timeout_s=900):
    """Whole spectral points through the reference flow: term-solves spread over `cores` processes, then the serial
    SOS_AGGREGATE pass per point (timed too).  Returns dict(points, terms, wall, busy_sum, agg_s, workers)."""
    ids = [i for i, t in enumerate(workload.terms) if t.optics in point_ids]
    res, wall, busy = solve_terms(workload, ids, cores, keep_records=True, timeout_s=timeout_s)
    fm = _formats()
    lib = C.CDLL(LIB)
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        for p in sorted(point_ids):
            mine = [(workload.terms[i].aik, res[i]) for i in ids if workload.terms[i].optics == p]
            if mine:
                aggregate_point(lib, fm, tmp, workload.optics[p].nbmu, mine)
    agg = time.perf_counter() - t0
    return dict(points=len(set(workload.terms[i].optics for i in ids)), terms=len(ids), wall=wall + agg, busy_sum=sum(busy) + agg,
                agg_s=agg, workers=len(busy))


if __name__ == "__main__":
    import sys
    _worker(sys.argv[sys.argv.index("--job") + 1], sys.argv[sys.argv.index("--out") + 1])
