#!/usr/bin/env python
"""tools/cfg4_bench.py -- BASELINE configs[3] at full size: a hyperspectral sweep of 2 100 wavelengths (0.4-2.5 um), 1-5 CKD
terms each (B ~ 6 000 term-solves, N = 25, OS_NB = 80), to time kernel (4) where it is bandwidth-type: the CKD aggregation
(k_aggregate, SOS_AGGREGATE) and the azimuth synthesis (k_trphi, SOS_TRPHI_OPTION on 13 azimuths) over thousands of
wavelengths in one launch each.  Prints one JSON line with achieved GB/s against MEASURED_PEAKS.json's HBM figure.
Lambertian ground: the two kernels do not depend on the surface, and 2 100 distinct surface matrices would be 3.8 GB."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radiativetransfer-sos_b200")
api = importlib.import_module("radiativetransfer-sos_b200.api")


def main():
    nwave = int(sys.argv[1]) if len(sys.argv) > 1 else 2100
    syn = pkg.synth
    rng = np.random.default_rng(20261022)
    wl = syn.Workload("hyperspectral_%d" % nwave)
    t0 = time.time()
    for p, nu in enumerate(np.linspace(4000.0, 25000.0, nwave)):
        lam = 1.0e4 / nu
        gfine = 0.60 - 0.05 * (lam - 0.55)
        o = syn.make_optics(nb_gauss=24, tetas=35.0, os_nb=80, surface="lambert", rho=0.1, g_modes=((0.85, 0.35), (gfine, 0.65)), seed=p)
        wl.optics.append(o)
        nt_ = int(rng.integers(1, 6))
        w, tg = syn._ckd_terms(nt_, rng)
        ta = 0.3 * (lam / 0.55) ** -1.3
        for k in range(nt_):
            wl.terms.append(syn.Term(p, float(w[k]), *syn.profile(syn.rayleigh_tau(lam), 8.0, ta, 2.0, tg[k], 3.0)))
    print("workload: %d wavelengths, %d term-solves built in %.1f s" % (nwave, len(wl.terms), time.time() - t0), file=sys.stderr)
    s = api.Solver(0)
    b = s.upload(wl)
    for _ in range(2):
        s.run(b, want_terms=False, want_groups=False)
        s.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=False)
    t0 = time.time()
    tr, gr = s.run(b, want_terms=True, want_groups=True, want_rec=False)
    st = s.stats(b)
    s.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=False)
    trphi_ms = s.last_kernel_ms
    N, W, rs = 25, 51, 81
    per = rs * 3 * W * 8
    agg_bytes = len(wl.terms) * per + nwave * per                 # reads every term's records, writes every group's
    trphi_bytes = float(np.sum(gr.n_rec)) * 3 * W * 8 + nwave * 2 * 7 * 13 * N * 8
    peak = 6542.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"config": "BASELINE configs[3]: %d wavelengths, %d term-solves, N=25, OS_NB=80" % (nwave, len(wl.terms)),
           "solve_ms": st["total_ms"], "wavelengths_per_s": nwave / (st["total_ms"] * 1e-3),
           "sweep_tflops": st["flops"] / (st["step_ms"] * 1e-3) / 1e12,
           "k_aggregate": {"ms": st["aggregate_ms"], "bytes": agg_bytes, "GBs": agg_bytes / (st["aggregate_ms"] * 1e-3) / 1e9,
                           "frac_of_hbm": agg_bytes / (st["aggregate_ms"] * 1e-3) / 1e9 / peak},
           "k_trphi": {"ms": trphi_ms, "bytes": trphi_bytes, "GBs": trphi_bytes / (trphi_ms * 1e-3) / 1e9,
                       "frac_of_hbm": trphi_bytes / (trphi_ms * 1e-3) / 1e9 / peak},
           "hbm_peak_GBs": peak, "n_fourier_mean": float(np.mean(tr.n_fourier))}
    print(json.dumps(out))
    b.free()
    s.close()


if __name__ == "__main__":
    main()
