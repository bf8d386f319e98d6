"""TEST INFRASTRUCTURE: api.Solver's methods as the keyword front end calls them (frontend.py, band.py, aerosols.py), every
device stage replaced by the reference's own routine in oracle/_ref/libsosref.so (aerosol optics: the device functions stepped on
the host, bit-identical to SOS_MIE / SOS_AEROSOLS, tests/test_aerosols_vs_reference.py).  With it the HOST logic of the front end
-- keyword mapping, optical-thickness scaling, the result-file hop of the aerosol coefficients, CKD term lists and modes, which
groups are aggregated, direct-term models, output shapes -- can be compared with the reference's driver SOS_PROC bit for bit
without a GPU (tests/test_sos_proc_vs_reference.py).  Nothing of the product uses this."""
import importlib

import numpy as np

import refdirect
from test_aerosols_vs_reference import HostSolver


class ReferenceFlowSolver(HostSolver):
    def __init__(self, hostlib, ref, tmp, ga, cores=8):
        super().__init__(hostlib)
        self.ref, self.tmp, self.ga, self.cores = ref, str(tmp), np.asarray(ga), cores
        self.fm = importlib.import_module("radiativetransfer-sos_b200.formats")
        self.direct = {}
        self.calls = []

    # ---- surface files (SOS_SURFACE's branches, frontend.surface) ----
    def glitter(self, nbmu, rmu, ga, wind, ind, os_nb, os_ns, os_nm):
        return refdirect.glitter(self.ref, self.fm, self.tmp, nbmu, rmu, ga, wind, ind, os_nb, os_ns, os_nm), None

    def roujean(self, nbmu, rmu, os_nb, k0, k1, k2):
        ier, surf = refdirect.roujean(self.ref, self.fm, self.tmp, nbmu, rmu, self.ga, os_nb, k0, k1, k2)
        assert ier == 0
        return surf

    def surface_bpdf(self, isurf, nbmu, rmu, ga, ind, os_nb, os_ns, os_nm, coef_c=0.0):
        return refdirect.surface_bpdf(self.ref, self.fm, self.tmp, isurf, nbmu, rmu, ga, ind, os_nb, os_ns, os_nm, coef_c=coef_c)

    def bpdf_ajout_brdf(self, a, b):
        return refdirect.bpdf_ajout_brdf(self.ref, self.fm, self.tmp, a, b)

    def set_direct_models(self, roujean=None, irondeaux=0, ibreon=0, nadal=None, maignan=None):
        assert nadal is None
        self.direct = dict(roujean=roujean, bpdf=dict(irondeaux=irondeaux, ibreon=ibreon, imaignan=int(maignan is not None),
                                                      coef=0.0 if maignan is None else maignan))

    def _fresh_library(self):
        import ctypes
        import os
        import shutil
        self.ncopy = getattr(self, "ncopy", 0) + 1
        dst = os.path.join(self.tmp, "libsosref_copy%d.so" % self.ncopy)
        shutil.copy(refdirect.runner().LIB, dst)
        return ctypes.CDLL(dst)

    # ---- profiles (SOS_ABSPROFILE, SOS_PROFILE and the PROFIL_TMP hop) ----
    def profile(self, altabs, tau, terms, text_hop=True):
        n = len(terms)
        nt, ier = np.zeros(n, np.int32), np.zeros(n, np.int32)
        z, h, pa, pm = (np.zeros((n, 601)) for _ in range(4))
        for i, t in enumerate(terms):
            # the layer branch of SOS_PROFILE starts its Rayleigh sums from Hmol(0) without setting it (SOS_PROFIL.F:873, set at
            # :926 only): a second call in one process starts from the previous call's top-of-atmosphere value (3e-7 TR) instead of
            # the zeroed storage the driver's single call -- and the product, csrc/profile_chain.cuh -- start from.  A fresh copy of
            # the library has fresh storage.
            lib = self._fresh_library() if t["iprofil"] == 2 else self.ref
            e, k, _, zz, hh, aa, mm = refdirect.profile(lib, self.tmp, np.asarray(altabs), np.asarray(tau)[i], t)
            ier[i] = e
            if e == 0:
                nt[i] = k
                z[i, :k + 1], h[i, :k + 1], pa[i, :k + 1], pm[i, :k + 1] = zz, hh, aa, mm
        return nt, z, h, pa, pm, ier

    def profile_chain(self, tables, userprofil, altabs, ro, terms, text_hop=True, want_tauabs=False):
        tau = np.zeros((len(terms), 50))
        for i, t in enumerate(terms):
            e, tau[i] = refdirect.absprofile(self.ref, tables, userprofil, altabs, ro, t)
            assert e == 0
        return self.profile(altabs, tau, terms) + ((tau,) if want_tauabs else ())

    # ---- solves, CKD sums, synthesis (SOS, SOS_AGGREGATE, SOS_TRPHI_OPTION) ----
    def upload(self, wl, groups=None, ngroup=None):
        outer = self

        class Batch:
            def __init__(s):
                s.wl, s.groups, s.ngroup, s.direct = wl, list(groups), ngroup, [0] * ngroup
                s.wmax = max(2 * o.nbmu + 1 for o in wl.optics)

            def free(s):
                outer.calls.append("free")
        return Batch()

    def set_group_direct(self, batch, direct):
        batch.direct = [int(d) for d in direct]

    def run(self, batch, want_terms=True, want_groups=True):
        wl = batch.wl
        r, _, _ = refdirect.runner().solve_terms(wl, list(range(len(wl.terms))), self.cores)
        ng, W = batch.ngroup, batch.wmax
        recs, sc = [], []
        for g in range(ng):
            ids = [i for i, gg in enumerate(batch.groups) if gg == g]
            N = wl.optics[wl.terms[ids[0]].optics].nbmu
            if batch.direct[g] and len(ids) == 1:                 # one solve, no SOS_AGGREGATE
                recs.append(r[ids[0]]["rec"])
                sc.append({k: r[ids[0]][k] for k in ("ttot_tronc", "ttot_vrai", "tauout", "emoins", "eplus")})
            else:
                rec, s = refdirect.runner().aggregate_point(self.ref, self.fm, self.tmp, N, [(wl.terms[i].aik, r[i]) for i in ids])
                recs.append(rec)
                sc.append(s)
        batch.recs, batch.sc = recs, sc

        class G:
            pass
        G.n_rec = np.array([int(np.flatnonzero(x.reshape(x.shape[0], -1).any(axis=1))[-1]) + 1 for x in recs], np.int32)
        G.rec = np.zeros((ng, max(x.shape[0] for x in recs), 3, W))
        for g, x in enumerate(recs):
            G.rec[g, :x.shape[0], :, :x.shape[2]] = x
        for k in ("ttot_tronc", "ttot_vrai", "tauout", "emoins", "eplus"):
            setattr(G, k, np.array([s[k] for s in sc]))
        return None, G

    def batch_trphi(self, batch, igli, wind, ind, ifresnel, itrphi, phios, pas_phi, ipolar, download=True):
        wl = batch.wl
        nmax = (batch.wmax - 1) // 2
        ups, downs = [], []
        for g in range(batch.ngroup):
            o = wl.optics[wl.terms[batch.groups.index(g)].optics]
            nphi, pf, th, up, down = refdirect.trphi_option(self.ref, self.fm, self.tmp, batch.recs[g], o.nbmu, o.rmu, o.ga,
                                                            batch.sc[g]["ttot_tronc"], batch.sc[g]["tauout"], igli, o.n0, wind, ind, ifresnel,
                                                            itrphi, phios, pas_phi, ipolar, roujean=self.direct.get("roujean"),
                                                            bpdf=self.direct.get("bpdf"))
            u, d = np.zeros((7, nphi, nmax)), np.zeros((7, nphi, nmax))
            u[:, :, :o.nbmu], d[:, :, :o.nbmu] = up, down
            ups.append(u)
            downs.append(d)
        return nphi, np.array(ups), np.array(downs)
