"""Throughput of the per-term profile chain on the device against the reference's own routines on one host
core: SOS_ABSPROFILE -> SOS_PROFILE -> PROFIL_TMP hop for the CKD terms of a band.   usage: python tools/profile_chain_bench.py [nterm ...]"""
import importlib
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import profile_cases as pc      # noqa: E402
import refdirect                # noqa: E402

api = importlib.import_module("radiativetransfer-sos_b200.api")
s = api.Solver(0)
user, altabs, ro = pc.gas_atmosphere(0)
t = pc.ckd_tables(0)
sizes = [int(a) for a in sys.argv[1:]] or [620, 16384]
for n in sizes:
    terms = s_terms = api.Solver.profile_terms(pc.make_terms(t, n, 0))
    for rep in range(3):
        t0 = time.perf_counter()
        nt, z, h, pa, pm, ier = s.profile_chain(t, user, altabs, ro, terms)
        call = (time.perf_counter() - t0) * 1e3
    print("profile chain, %d terms: call %.2f ms (tables + terms H2D, 3 kernels, %.1f MB D2H), kernels %.3f ms -> %.0f terms/s "
          "(call), mean NT %.0f, errors %d" % (n, call, 4 * n * 601 * 8 / 1e6, s.last_kernel_ms, n / call * 1e3, nt.mean(), int((ier != 0).sum())))
ref = refdirect.lib()
if ref is not None:
    tmp = tempfile.mkdtemp()
    sub = pc.make_terms(t, 200, 0)
    t0 = time.perf_counter()
    for term in sub:
        _, tau = refdirect.absprofile(ref, t, user, altabs, ro, term)
        refdirect.profile(ref, tmp, altabs, tau, term)
    dt = time.perf_counter() - t0
    print("reference routines (translated Fortran, one host core, incl. the PROFIL_TMP file): %d terms in %.2f s -> %.0f terms/s"
          % (len(sub), dt, len(sub) / dt))
