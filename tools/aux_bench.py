"""Timing of the non-dominant kernels of the path at the demo / bench sizes:
k_glitter (N=41, OS_NB=OS_NS=80, OS_NM=160) against the C restatement on one host core, and the batched azimuth
synthesis k_trphi on the 96 CKD-summed wavelengths of the bench band.   usage: python tools/aux_bench.py [--cpu]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radiativetransfer-sos_b200")
api = importlib.import_module("radiativetransfer-sos_b200.api")
s = api.Solver(0)
rmu, ga, n0, _ = pkg.synth.sos_angles(40, 35.0)
N = (rmu.size - 1) // 2
for rep in range(3):
    t0 = time.perf_counter()
    surf, il = s.glitter(N, rmu, ga, 2.0, 1.34, 80, 80, 160)
    print("glitter N=%d OS_NB=80 OS_NS=80 OS_NM=160: call %.2f ms (H2D, k_mat_fresnel + E15.8 round trip, kernel, D2H of %.1f MB), "
          "k_glitter %.3f ms, mean IL %.1f" % (N, (time.perf_counter() - t0) * 1e3, surf.nbytes / 1e6, s.last_kernel_ms, il.mean()))
if "--cpu" in sys.argv:
    from oracle import oracle as orc
    t0 = time.perf_counter()
    ref, il0 = orc.glitter(N, rmu, ga, 2.0, 1.34, 80, 80, 160)
    dt = time.perf_counter() - t0
    print("C restatement of SOS_GLITTER on one host core: %.2f s; IL identical: %s; REAL*4 bit-identical fraction %.5f"
          % (dt, bool(np.array_equal(il, il0)), float(np.mean(surf.view(np.uint32) == ref.view(np.uint32)))))
wl = pkg.synth.config_ckd_band(npoints=96, seed=20261021, nb_gauss=40, os_nb=80, surface="lambert", rho=0.1)
b = s.upload(wl)
s.run(b, want_terms=False, want_groups=False)
st = s.stats(b)
for dl in (False, False, True):
    t0 = time.perf_counter()
    n, up, dn = s.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=dl)
    print("batch_trphi 96 wavelengths x %d azimuths download=%s: call %.2f ms, k_trphi %.3f ms" %
          (n, dl, (time.perf_counter() - t0) * 1e3, s.last_kernel_ms))
rec_bytes = 96 * 81 * 3 * 83 * 8
out_bytes = 96 * 2 * 7 * 13 * 41 * 8
print("k_trphi algorithmic bytes: %.1f MB read + %.1f MB written" % (rec_bytes / 1e6, out_bytes / 1e6))
