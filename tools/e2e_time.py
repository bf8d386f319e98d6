"""Per-phase wall time of the host-buffer path (upload / run+download / synthesis / free) on the bench workload.
usage: python tools/e2e_time.py [points] [reps]   (SOS_TRACE=1 adds the library's own breakdown)"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("radiativetransfer-sos_b200")
api = importlib.import_module("radiativetransfer-sos_b200.api")
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 96
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
wl = pkg.synth.config_ckd_band(npoints=npts, seed=20261021, nb_gauss=40, os_nb=80, surface="lambert", rho=0.1)
s = api.Solver(0)
resident = s.upload(wl)                                  # like bench.py: a resident batch stays alive meanwhile
s.run(resident, want_terms=False, want_groups=False)
for it in range(reps):
    t = [time.perf_counter()]
    b = s.upload(wl); t.append(time.perf_counter())
    tr, gr = s.run(b, want_terms=True, want_groups=True, want_rec=False); t.append(time.perf_counter())
    s.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=True); t.append(time.perf_counter())
    b.free(); t.append(time.perf_counter())
    print("it %d: upload %.1f  run+download %.1f  trphi %.1f  free %.1f  total %.1f ms" %
          ((it,) + tuple((t[i + 1] - t[i]) * 1e3 for i in range(4)) + ((t[4] - t[0]) * 1e3,)), flush=True)
