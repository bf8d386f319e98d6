"""Short single-GPU driver for ncu / timing experiments: uploads the bench workload once and runs the hot path.
usage: python tools/prof_run.py [points] [runs]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radiativetransfer-sos_b200")
api = importlib.import_module("radiativetransfer-sos_b200.api")

points = int(sys.argv[1]) if len(sys.argv) > 1 else 24
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = pkg.synth.config_ckd_band(npoints=points, seed=20261021, nb_gauss=40, os_nb=80, surface="lambert", rho=0.1)
s = api.Solver(0)
if os.environ.get("SOS_WAVE"):
    s.set_options(0, int(os.environ["SOS_WAVE"]))
b = s.upload(wl)
for r in range(runs):
    t0 = time.perf_counter()
    s.run(b, want_terms=False, want_groups=False)
    dt = time.perf_counter() - t0
    st = s.stats(b)
    print("run %d: %.1f ms wall, %.1f ms dev, k_step %.1f ms in %d launches, %.2f TFLOP/s (k_step), %d steps, %d terms"
          % (r, dt * 1e3, st["total_ms"], st["step_ms"], st["step_launches"],
             st["flops"] / max(st["step_ms"], 1e-9) / 1e9, st["steps"], len(wl.terms)))
b.free()
s.close()

# useful vs executed contraction steps (Fourier orders computed beyond the Fourier stop are discarded work)
s2 = api.Solver(0)
tr, _ = s2.solve(wl, want_groups=False, want_rec=False)
useful = sum(int(max(tr.n_scatter[i, k] - 1, 0)) for i in range(len(wl.terms)) for k in range(tr.n_fourier[i]))
print("useful steps %d of %d executed (%.1f%%); n_fourier min/mean/max %d/%.1f/%d" %
      (useful, st["steps"], 100.0 * useful / st["steps"], tr.n_fourier.min(), tr.n_fourier.mean(), tr.n_fourier.max()))
