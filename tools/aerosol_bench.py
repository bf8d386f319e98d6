"""Timing of the aerosol chain on the GPU: k_mie alone for Mie tables of growing size-parameter range (second
call of each: the memory pool has grown by then) and sosgpu_aerosols for a hyperspectral-sweep-sized list of models.
usage: python tools/aerosol_bench.py [nwavelengths]   (run on the GPU box; prints lines for profiles/)"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aerosol_cases as ac                                     # input generators only (angles, size distributions)

api = importlib.import_module("radiativetransfer-sos_b200.api")
s = api.Solver(0)
nbmu, xmu, xhr = ac.mie_angles(40, (0.0,))
for af in (100.0, 200.0, 1200.0, 4900.0):
    ms = []
    for rep in range(3):
        t0 = time.time()
        t = s.mie(nbmu, xmu, 1.45, -0.004, 0.0001, af)
        ms.append((s.last_kernel_ms, 1e3 * (time.time() - t0)))
    nsteps = float(np.sum(np.trunc(2.0 * t["rec"][:, 0].astype(np.float64) + 5.0)))
    print("k_mie alphaf=%6.0f: %5d records x %d angles, %.3g series terms: kernel %.2f ms (first call %.2f), call %.1f ms (first %.1f); "
          "%.2f G angle-terms/s" % (af, t["g"].size, 2 * nbmu + 1, nsteps, ms[-1][0], ms[0][0], ms[-1][1], ms[0][1],
                                    nsteps * (2 * nbmu + 1) / ms[-1][0] / 1e6))
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 256
comps, models = [], []
for i, wa in enumerate(np.linspace(0.4, 2.5, nw)):
    n0 = len(comps)
    for rn, in_, ig, v1, v2, v3 in ((1.45 - 0.01 * i / nw, -0.004, 1, 0.40, 0.60, -999.0), (1.42 + 0.01 * i / nw, -0.008, 1, 0.08, 0.45, -999.0)):
        comps.append((rn, in_, 0.0001, ac.alphaf_for(ac.lnd_rmax(v1, v2), float(wa)), ig, v1, v2, v3, float(wa)))
    models.append((2, [n0, n0 + 1], [0.5, 0.5], 1))
for rep in range(2):
    t0 = time.time()
    o = s.aerosols(nbmu, xmu, xhr, comps, models, 80, want_phase=False)
    dt = 1e3 * (time.time() - t0)
assert (o["model_ier"] == 0).all()
nrec = sum(s.mie_count(c[2], c[3]) for c in comps)
print("sosgpu_aerosols: %d wavelengths x 2 log-normal modes, wavelength-dependent index -> %d Mie tables, %d records x %d angles, "
      "expansions to order 80: kernels %.1f ms, call %.1f ms (%.0f wavelengths/s)" % (nw, len(comps), nrec, 2 * nbmu + 1, s.last_kernel_ms, dt,
                                                                                   nw / dt * 1e3))
s.close()
