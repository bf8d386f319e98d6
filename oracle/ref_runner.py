"""ref_runner.py -- TEST / BENCH INFRASTRUCTURE.  Runs the reference's own per-term flow on host cores from
oracle/_ref/libsosref.so (the reference's Fortran sources translated by oracle/f77_to_c.py): for every CKD term
PROFIL_TMP is written, SOS reads it and calls SOS_OS, SOS_OS writes the term's Fourier records, SOS_AGGREGATE folds them
into the wavelength's SOS_Result.bin through its temporary file -- files and all, as SOS_PROC.F:3459-3594 does.
The translated code keeps Fortran's static storage and unit table, so parallelism is one PROCESS per core (spawned),
which is also how one would run the Fortran executable."""
import ctypes as C
import os
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsosref.so")
MX, NBM = 80, 200


def available():
    return os.path.exists(LIB)


def _fs(s):
    return C.create_string_buffer(s.encode().ljust(500), 500)


def _worker(job):
    """job = (optics dict, [(aik, zprof, h, pcaer, pcmol), ...] grouped by spectral point) -> (points, terms, seconds)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sos_formats", os.path.join(os.path.dirname(HERE), "radiativetransfer-sos_b200", "formats.py"))
    fm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fm)
    lib = C.CDLL(LIB)
    ip = lambda v: C.byref(C.c_int(int(v)))
    dp = lambda v: C.byref(C.c_double(float(v)))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L = C.c_size_t(500)
    t0 = time.perf_counter()
    nterm = 0
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        fprof, fos, fagg, fres = (os.path.join(tmp, n) for n in ("PROFIL_TMP", "OS_TMP.bin", "AGG_TMP.bin", "SOS_Result.bin"))
        for o, terms in job:
            N = o["nbmu"]
            rmu, ga = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
            rmu[MX - N:MX + N + 1], ga[MX - N:MX + N + 1] = o["rmu"], o["ga"]
            pad = lambda v: np.concatenate([np.asarray(v, dtype=np.float64), np.zeros(NBM + 1 - len(v))])
            if os.path.exists(fres):
                os.remove(fres)
            acc = [C.c_double(0) for _ in range(6)]
            tdg_tmp, tdg = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
            for aik, z, h, xa, ym in terms:
                al, be, gm, ze = pad(o["alpha"]), pad(o["beta"]), pad(o["gamma"]), pad(o["zeta"])
                fm.write_profile(fprof, z, h, xa, ym)
                sc = [C.c_double(0) for _ in range(6)]
                ier = C.c_int(0)
                r = rmu.copy()
                lib.sos_(_fs(fos), _fs("NO_OUTPUT"), _fs(fprof), ip(len(h) - 1), dp(o["zout"]), ip(o["igmax"]), ip(o["ipolar"]),
                         dp(o["ron"]), dp(o["ind_surf"]), dp(o["rho"]), ip(0), ip(o["ifresnel"]), _fs("none"), ip(o["n0"]),
                         dp(o["piz"]), dp(o["piztr"]), dp(o["a_trunc"]), P(r), P(ga), dp(o["tetas"]), ip(o["os_nb"]), ip(N),
                         P(al), P(be), P(gm), P(ze), C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]), C.byref(sc[3]), P(tdg_tmp),
                         C.byref(sc[4]), C.byref(sc[5]), ip(0), ip(6), C.byref(ier), L, L, L, L)
                if ier.value != 0:
                    raise RuntimeError("reference SOS returned IER=%d" % ier.value)
                lib.sos_aggregate_(ip(N), dp(aik), _fs(fos), C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]), C.byref(sc[3]), P(tdg_tmp),
                                   C.byref(sc[4]), C.byref(sc[5]), _fs(fagg), _fs(fres), C.byref(acc[0]), C.byref(acc[1]),
                                   C.byref(acc[2]), C.byref(acc[3]), P(tdg), C.byref(acc[4]), C.byref(acc[5]), C.byref(ier), L, L, L)
                if ier.value != 0:
                    raise RuntimeError("reference SOS_AGGREGATE returned IER=%d" % ier.value)
                nterm += 1
    return len(job), nterm, time.perf_counter() - t0


def _optics_dict(o):
    if o.imat_surf == 1:
        raise ValueError("surface-matrix workloads are not wired into the reference runner")
    return dict(nbmu=o.nbmu, rmu=np.asarray(o.rmu, dtype=np.float64), ga=np.asarray(o.ga, dtype=np.float64), n0=o.n0, tetas=o.tetas,
                os_nb=o.os_nb, alpha=np.asarray(o.alpha), beta=np.asarray(o.beta), gamma=np.asarray(o.gamma), zeta=np.asarray(o.zeta),
                a_trunc=o.a_trunc, piz=o.piz, piztr=o.piztr, ron=o.ron, rho=o.rho, ifresnel=o.ifresnel, ind_surf=o.ind_surf,
                igmax=o.igmax, ipolar=o.ipolar, zout=o.zout)


def run_points(workload, point_ids, cores, timeout_s=600):
    """Whole spectral points of a synth.Workload through the reference flow on `cores` worker processes (plain
    subprocesses of this file: no fork of a CUDA process, no re-import of the caller's __main__).
    Returns (points, term_solves, wall_seconds)."""
    import json
    import pickle
    import subprocess
    import sys
    by_point = {}
    for t in workload.terms:
        if t.optics in point_ids:
            by_point.setdefault(t.optics, []).append((t.aik, np.asarray(t.zprof), np.asarray(t.h), np.asarray(t.pcaer), np.asarray(t.pcmol)))
    units = [(_optics_dict(workload.optics[p]), terms) for p, terms in sorted(by_point.items())]
    units.sort(key=lambda u: -sum(len(t[2]) for t in u[1]))          # longest first
    jobs = [[] for _ in range(max(1, min(cores, len(units))))]
    load = [0] * len(jobs)
    for u in units:
        k = load.index(min(load))
        jobs[k].append(u)
        load[k] += sum(len(t[2]) for t in u[1])
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i, job in enumerate(jobs):
            paths.append(os.path.join(tmp, "job%d.pkl" % i))
            with open(paths[-1], "wb") as f:
                pickle.dump(job, f)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--job", p], stdout=subprocess.PIPE, text=True) for p in paths]
        out = []
        try:
            for pr in procs:
                so, _ = pr.communicate(timeout=timeout_s)
                if pr.returncode != 0:
                    raise RuntimeError("reference worker failed (rc=%d)" % pr.returncode)
                out.append(json.loads(so.strip().splitlines()[-1]))
        finally:
            for pr in procs:
                if pr.poll() is None:
                    pr.kill()
        wall = time.perf_counter() - t0
    return sum(o[0] for o in out), sum(o[1] for o in out), wall


if __name__ == "__main__":
    import json
    import pickle
    import sys
    with open(sys.argv[sys.argv.index("--job") + 1], "rb") as f:
        print(json.dumps(_worker(pickle.load(f))))
