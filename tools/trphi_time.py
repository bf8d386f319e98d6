import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radiativetransfer-sos_b200")
api = importlib.import_module("radiativetransfer-sos_b200.api")
wl = pkg.synth.config_ckd_band(npoints=96, seed=20261021, nb_gauss=40, os_nb=80, surface="lambert", rho=0.1)
s = api.Solver(0)
b = s.upload(wl)
s.run(b, want_terms=False, want_groups=False)
for dl in (False, False, True, True):
    t0 = time.perf_counter()
    s.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=dl)
    print("batch_trphi download=%s: %.2f ms" % (dl, (time.perf_counter() - t0) * 1e3))
