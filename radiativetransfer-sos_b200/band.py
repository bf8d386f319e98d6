"""Multi-wavelength front end over the C ABI: what SOS_PROC does for ONE wavelength (SOS_PROC.F:3340-3874) --
CKD term list and weights, per-term gas and scattering profiles, one term-solve per CKD term, CKD aggregation, azimuth
synthesis, result files -- done for a LIST of wavelengths with every stage batched on the device:

    CKD terms + AIK      enumerate_ckd_terms          the eight nested loops of SOS_PROC.F:3459-3487, host (a few integers)
    profiles             Solver.profile_chain         SOS_ABSPROFILE -> SOS_PROFILE -> PROFIL_TMP hop, all terms at once
    term-solves + sum    Solver.upload / run          SOS + SOS_OS + SOS_AGGREGATE, all terms of all wavelengths at once
    synthesis            Solver.batch_trphi           SOS_TRPHI_OPTION, all wavelengths at once
    transmissions        Solver.transmissions         the -SOS.Trans solves of SOS.F:605-637 (optional)
    files                write_updown / write_trans / write_flux / formats.write_result_bin, one directory per wavelength

Keywords and the aerosol / surface preparation are frontend.py's (which calls this module): here a wavelength arrives as
the optics SOS_PREPA_OS would hand to SOS plus the scalars of the profile.  No CPU fallback: Solver() raises without a GPU."""
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import api, formats
from .synth import Optics, Term, Workload

NBABS = 8


@dataclass
class Wavelength:
    """One spectral point of the band."""
    optics: Optics                  # outputs of SOS_PREPA_OS for this wavelength + the per-run scalars of the SOS call
    lamb1: int                      # index of its spectral interval in the CKD tables (1-based, SOS_PREPA_ABSPROFILE)
    tr: float                       # Rayleigh optical thickness
    ta: float                       # aerosol optical thickness
    hr: float = 8.0
    ha: float = 2.0
    absprofil: int = 2              # 7: no gaseous absorption
    iprofil: int = 1
    zmin: float = 0.0
    zmax: float = 0.0
    name: str = ""


@dataclass
class BandResult:
    nterm: List[int] = field(default_factory=list)         # CKD terms per wavelength
    nt: Optional[np.ndarray] = None                        # levels per term
    groups: object = None                                  # api.GroupResults (CKD-summed Fourier coefficients, fluxes, taus)
    nphi: int = 0
    up: Optional[np.ndarray] = None                        # [nwave, 7, nphi, N]
    down: Optional[np.ndarray] = None
    tdifmus: Optional[np.ndarray] = None                   # [nwave]
    tdifmug: Optional[np.ndarray] = None                   # [nwave, N]
    dirs: List[str] = field(default_factory=list)
    ind_angout: Optional[np.ndarray] = None                # user-angle flags of the radiance angles (set by frontend.run)


def enumerate_ckd_terms(nexp, kdis_ai, lamb1):
    """The CKD terms of one spectral interval in the reference's loop order (IK1 outermost .. IK8 innermost,
    SOS_PROC.F:3459-3466) with their weights AIK = prod_k KDIS_AI(IKk, k, LAMB1) / sum (:3481-3489).
    nexp: NEXP(8, 50); kdis_ai: KDIS_AI(5, 8, 50), Fortran order.  Raises where the reference stops (|sum - 1| >= 1e-6)."""
    n = [int(nexp[k, lamb1 - 1]) for k in range(NBABS)]
    iks, aik = [], []
    idx = [1] * NBABS
    total = int(np.prod(n))
    for _ in range(total):
        a = kdis_ai[idx[0] - 1, 0, lamb1 - 1]
        for k in range(1, NBABS):                            # left to right as the Fortran product
            a = a * kdis_ai[idx[k] - 1, k, lamb1 - 1]
        iks.append(tuple(idx))
        aik.append(float(a))
        for k in range(NBABS - 1, -1, -1):                   # IK8 fastest
            idx[k] += 1
            if idx[k] <= n[k]:
                break
            idx[k] = 1
    s = 0.0
    for a in aik:
        s = s + a
    if abs(s - 1.0) >= 1e-6:                                 # SOS_PROC.F:3410
        raise ValueError("CKD weights of interval %d sum to %.9f" % (lamb1, s))
    return iks, [a / s for a in aik]


def estimated_absorption(aik, tau):
    """-SOS.AbsModeCKD 2: the absorption profile of ONE solve from the CKD terms of an interval (SOS_PROC.F:3613-3662):
    TRSCKD(I) = sum over the terms, in loop order, of AIK exp(-TAUABS(I)); TAUABS(I) = -log TRSCKD(I), negative values set to 0.
    aik [nterm], tau [nterm, 50] -> [50]."""
    trs = np.zeros(tau.shape[1])
    for a, t in zip(aik, tau):
        trs = trs + a * np.exp(-t)
    est = -np.log(trs)
    est[est < 0.0] = 0.0
    return est


def run_band(solver, tables, kdis_ai, userprofil, altabs, ro, waves, itrphi=1, phios=0.0, pas_phi=30, outdir=None,
             trans=False, flux=False, ckd_mode=1):
    """Runs the band.  tables: CKD tables as READ_CKD_COEFF fills them (dict, Fortran-ordered: nb_temp, nb_pres, nb_conc,
    tab_temp, tab_pres, tab_conc, nexp, ki, kh) -- or a list of such dicts, one per wavelength (the same object for wavelengths of the
    same coefficient file); kdis_ai: KDIS_AI(5,8,50) (or the list of them); userprofil / altabs / ro: the gas atmosphere of
    SOS_PREPA_ABSPROFILE.  With outdir, wavelength w gets outdir/<name or index>/SOS_Up.txt, SOS_Down.txt, SOS_Result.bin and,
    on request, SOS_Trans.txt / SOS_Flux.txt.  ckd_mode: -SOS.AbsModeCKD, 1 = one solve per CKD term and the AIK-weighted sum
    (SOS_PROC.F:3454-3600), 2 = one solve per wavelength on the absorption profile estimated from its CKD terms (:3609-3716);
    a wavelength without gaseous absorption is one solve in either mode (:2366)."""
    if ckd_mode not in (1, 2):
        raise ValueError("-SOS.AbsModeCKD must be 1 or 2 (SOS_PROC error 2515)")
    res = BandResult()
    pterms, aiks, owner = [], [], []
    # tables / kdis_ai: one set for the band, or a list with one entry per wavelength (a band that spans several CKD coefficient
    # files: READ_CKD_COEFF fills the tables with the 50 spectral intervals of one file)
    per_wave = isinstance(tables, (list, tuple))
    tab_of = (lambda w: tables[w]) if per_wave else (lambda w: tables)
    ai_of = (lambda w: kdis_ai[w]) if per_wave else (lambda w: kdis_ai)
    for w, wv in enumerate(waves):
        if wv.absprofil == 7:
            iks, aik = [(1,) * NBABS], [1.0]
        else:
            iks, aik = enumerate_ckd_terms(tab_of(w)["nexp"], ai_of(w), wv.lamb1)
        res.nterm.append(len(iks))
        for ik, a in zip(iks, aik):
            pterms.append(dict(lamb1=wv.lamb1, ik=ik, absprofil=wv.absprofil, iprofil=wv.iprofil, tr=wv.tr, hr=wv.hr, ta=wv.ta,
                               ha=wv.ha, zmin=wv.zmin, zmax=wv.zmax))
            aiks.append(a)
            owner.append(w)
    # ---- per-term profiles (device) ----
    if tables is None:                                        # -AP.AbsProfile.Type 7 for every wavelength: no gas tables needed
        if any(wv.absprofil != 7 for wv in waves):
            raise ValueError("run_band without CKD tables: every wavelength must have absprofil = 7 (no gaseous absorption)")
        altabs = np.linspace(120.0, 0.0, 50) if altabs is None else altabs
        tau = np.zeros((len(pterms), 50))
        nt, z, h, pa, pm, ier = solver.profile(altabs, tau, pterms, text_hop=True)
    elif not per_wave:
        nt, z, h, pa, pm, ier, tau = solver.profile_chain(tables, userprofil, altabs, ro, pterms, text_hop=True, want_tauabs=True)
    else:                                                     # one device call per run of wavelengths that share their tables
        parts, i0 = [], 0
        while i0 < len(pterms):
            t0 = tab_of(owner[i0])
            i1 = i0
            while i1 < len(pterms) and tab_of(owner[i1]) is t0:
                i1 += 1
            if t0 is None:
                raise ValueError("wavelength %d has gaseous absorption but no CKD tables" % owner[i0])
            parts.append(solver.profile_chain(t0, userprofil, altabs, ro, pterms[i0:i1], text_hop=True, want_tauabs=True))
            i0 = i1
        nt, z, h, pa, pm, ier, tau = (np.concatenate([p[k] for p in parts]) for k in range(7))
    if ier.any():
        bad = int(np.flatnonzero(ier)[0])
        raise RuntimeError("profile chain: term %d of wavelength %d failed with code %d" % (bad, owner[bad], int(ier[bad])))
    if ckd_mode == 2 and tables is not None:
        # one solve per wavelength: the per-term absorption profiles (device) -> estimated profile (host, 50 levels per wavelength)
        # -> SOS_PROFILE for one term per wavelength (device).  The per-term profiles of the chain above are not used.
        first = np.cumsum([0] + res.nterm)
        tau = np.array([estimated_absorption(aiks[first[w]:first[w + 1]], tau[first[w]:first[w + 1]]) if wv.absprofil != 7
                        else tau[first[w]] for w, wv in enumerate(waves)])
        pterms = [pterms[first[w]] for w in range(len(waves))]
        aiks, owner, res.nterm = [1.0] * len(waves), list(range(len(waves))), [1] * len(waves)
        nt, z, h, pa, pm, ier = solver.profile(altabs, tau, pterms, text_hop=True)
        if ier.any():
            bad = int(np.flatnonzero(ier)[0])
            raise RuntimeError("profile of wavelength %d (estimated absorption) failed with code %d" % (bad, int(ier[bad])))
    res.nt = nt
    # ---- term-solves, CKD sums, synthesis (device) ----
    wl = Workload("band")
    wl.optics = [wv.optics for wv in waves]
    for i, w in enumerate(owner):
        n = int(nt[i]) + 1
        wl.terms.append(Term(w, aiks[i], z[i, :n].copy(), h[i, :n].copy(), pa[i, :n].copy(), pm[i, :n].copy()))
    batch = solver.upload(wl, groups=owner, ngroup=len(waves))
    try:
        # one solve that SOS_PROC does not aggregate (no gas: SOS_PROC.F:2366; mode 2: :3700-3708): SOS's own optical thicknesses
        direct = [int(n == 1 and (wv.absprofil == 7 or ckd_mode == 2)) for n, wv in zip(res.nterm, waves)]
        if any(direct):
            solver.set_group_direct(batch, direct)
        _, gr = solver.run(batch, want_terms=False, want_groups=True)
        res.groups = gr
        o0 = waves[0].optics
        res.nphi, res.up, res.down = solver.batch_trphi(batch, o0.igli, o0.wind, o0.ind_surf, o0.ifresnel, itrphi, phios, pas_phi,
                                                        o0.ipolar)
    finally:
        batch.free()
    if trans:
        tds, tdg = solver.transmissions(wl)
        nw, nmax = len(waves), max(wv.optics.nbmu for wv in waves)
        res.tdifmus, res.tdifmug = np.zeros(nw), np.zeros((nw, nmax))
        for i, w in enumerate(owner):                         # SOS_AGGREGATE.F:415-436: AIK-weighted sums, term order
            res.tdifmus[w] += aiks[i] * tds[i]
            res.tdifmug[w, :waves[w].optics.nbmu] += aiks[i] * tdg[i, :waves[w].optics.nbmu]
    # ---- files ----
    if outdir is not None:
        first = np.cumsum([0] + res.nterm)
        for w, wv in enumerate(waves):
            o = wv.optics
            d = os.path.join(outdir, wv.name or "wave_%04d" % w)
            os.makedirs(d, exist_ok=True)
            res.dirs.append(d)
            N = o.nbmu
            nrec = int(gr.n_rec[w])
            formats.write_result_bin(os.path.join(d, "SOS_Result.bin"), gr.rec[w, :nrec, :, :2 * N + 1])
            theta = np.degrees(np.arccos(np.asarray(o.rmu)[N + 1:2 * N + 1]))
            phi = np.array([phios, phios + 180.0]) if itrphi == 1 else np.arange(0.0, 360.0 + 0.5 * pas_phi, pas_phi)
            api.write_updown(os.path.join(d, "SOS_Up.txt"), os.path.join(d, "SOS_Down.txt"), N, itrphi, phios, pas_phi, o.zout,
                             phi[:res.nphi], theta, res.up[w, :, :res.nphi, :N], res.down[w, :, :res.nphi, :N])
            if trans:
                api.write_trans(os.path.join(d, "SOS_Trans.txt"), o.tetas, gr.ttot_tronc[w], gr.ttot_vrai[w], res.tdifmus[w],
                                np.asarray(o.rmu)[N + 1:2 * N + 1], res.tdifmug[w, :N])
            if flux:
                if userprofil is None:
                    raise ValueError("the Flux file needs the gas atmosphere's altitudes (userprofil)")
                last = first[w + 1] - 1                       # TAUABS of the last CKD term, as SOS_PROC leaves it (:3860)
                api.write_flux(os.path.join(d, "SOS_Flux.txt"), o.tetas, gr.ttot_tronc[w], gr.ttot_vrai[w], gr.emoins[w], gr.eplus[w],
                               wv.tr, wv.hr, wv.ta, wv.ha, np.asarray(userprofil)[:, 0], tau[last])
    return res
