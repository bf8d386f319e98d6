"""Generates tests/golden/oracle_small.npz from the C oracle.  tests/test_oracle_vs_reference.py
(test_committed_golden_vectors_are_what_the_reference_produces) checks that the committed file equals, bit for bit, what
the reference's own SOS produces when run from oracle/_ref (the Fortran sources translated to C), so the fixtures are
reference-generated vectors in content even though no Fortran compiler exists here.
Run: python tests/make_golden.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases(pkg):
    syn = pkg.synth
    out = {}
    o = syn.make_optics(nb_gauss=8, tetas=35.0, os_nb=16, surface="lambert", rho=0.1)
    out["lambert"] = (o, syn.Term(0, 1.0, *syn.profile(0.03, 8.0, 0.2, 2.0, 0.05)))
    o = syn.make_optics(nb_gauss=8, tetas=50.0, os_nb=16, surface="brdf", rho=0.05)
    out["brdf"] = (o, syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.1, 2.0, 0.5)))
    o = syn.make_optics(nb_gauss=6, tetas=20.0, os_nb=12, surface="fresnel", rho=0.0)
    out["fresnel"] = (o, syn.Term(0, 1.0, *syn.profile(0.1, 8.0, 0.05, 2.0, 0.0)))
    o = syn.make_optics(nb_gauss=8, tetas=35.0, os_nb=16, surface="brdf", rho=0.05, zout=3.0)
    out["zout"] = (o, syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.3, 2.0, 0.1)))
    return out


if __name__ == "__main__":
    from util import oracle_term
    from oracle import oracle as orc
    pkg = importlib.import_module("radiativetransfer-sos_b200")
    d = {}
    for name, (o, t) in cases(pkg).items():
        r = oracle_term(orc, o, t)
        d[name + "_nf"] = r.n_fourier
        d[name + "_nsc"] = r.n_scatter
        d[name + "_rec"] = r.rec
        print(name, r.n_fourier, r.n_scatter)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_small.npz"), **d)
