"""Synthetic inputs of the shapes BASELINE.json names.

Everything produced here is passed through the reference's lossy text / REAL*4 channels
(formats.py) so that the CUDA path and the oracle consume bit-identical rounded inputs.
No network, no CKD tables: absorption is a synthetic k-distribution.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .formats import round_e, round_f

# REAL*4 literals of inc/SOS.h promoted to double
CTE_TCOUCHE = float(np.float32(0.005))                       # SOS.h:208
CTE_TOA_FIRST_LAYER = float(np.float32(0.0002))              # SOS.h:213
CTE_OS_NT = 600                                              # SOS.h:202
CTE_OS_NT_MIN = 100                                          # SOS.h:229
CTE_MDF = float(np.float32(0.0279))                          # SOS.h:373
CTE_SEUIL_ECART_MUS = float(np.float32(0.00001))             # SOS.h:561


def sos_gauss(mm):
    """Restatement of SOS_GAUSS (SOS_ANGLES.F:1022-1103): Gauss-Legendre nodes on (-1,1) for N=2*MM-2 points,
    Newton iteration from the asymptotic root estimate.  Returns (mu[1..MM-1], w[1..MM-1]) in the
    reference's AMU(K), K=1..MM-1 order (ascending mu)."""
    tol = 1.0e-15
    pi = np.arccos(-1.0)
    n = 2 * mm - 2
    aa = 2.0 / pi ** 2
    ab = -62.0 / (3.0 * pi ** 4)
    ac = 15116.0 / (15.0 * pi ** 6)
    ad = -12554474.0 / (105.0 * pi ** 8)
    en = float(n)
    u = 1.0 - (2.0 / pi) ** 2
    d = 1.0 / np.sqrt((en + 0.5) ** 2 + u / 4.0)
    r = np.zeros(n)
    w = np.zeros(n)
    for k in range(1, n + 1):
        az = 4.0 * k - 1.0
        z = 0.25 * pi * (az + aa / az + ab / az ** 3 + ac / az ** 5 + ad / az ** 7)
        x = np.cos(z * d)
        while True:
            pa = [1.0, x]
            for nn in range(3, n + 2):
                enn = nn - 1.0
                pa.append(((2.0 * enn - 1.0) * x * pa[-1] - (enn - 1.0) * pa[-2]) / enn)
            pnp = en * (pa[n - 1] - x * pa[n]) / (1.0 - x * x)
            xi = x - pa[n] / pnp
            if abs(xi - x) - tol <= 0.0:
                r[k - 1] = x
                w[k - 1] = 2.0 * (1.0 - x * x) / (en * pa[n - 1]) ** 2
                break
            x = xi
    amu = np.zeros(mm)
    pmu = np.zeros(mm)
    for i in range(1, mm):
        k = mm - i
        amu[k] = r[i - 1]
        pmu[k] = w[i - 1]
    return amu[1:], pmu[1:]


def sos_angles(nb_gauss, tetas, user_angles_deg=()):
    """SOS_ANGLES + SOS_ANGLES_GAUSS_USER (SOS_ANGLES.F:227-650, 713-...): Gauss angles (descending mu),
    user angles with zero weight, solar angle inserted (weight 0) unless it coincides with a node.
    Returns rmu[2N+1], ga[2N+1] (index j at [j+N], rmu[-j] = -rmu[j]), n0, user flags."""
    mu, w = sos_gauss(nb_gauss + 1)
    mu = list(mu[:nb_gauss]) + [float(np.cos(np.deg2rad(a))) for a in user_angles_deg]
    w = list(w[:nb_gauss]) + [0.0] * len(user_angles_deg)
    order = sorted(range(len(mu)), key=lambda i: -mu[i])
    mu = [mu[i] for i in order]
    w = [w[i] for i in order]
    flag = [1 if x == 0.0 else 0 for x in w]
    xmus = float(np.cos(tetas * (np.arccos(-1.0) / 180.0)))       # XMUS = DCOS(TETAS*CONVDEGRAD), CONVDEGRAD = PI/180 (SOS_ANGLES.F:296, 402)
    imus = -1
    for j, m in enumerate(mu):
        if abs(xmus - m) < CTE_SEUIL_ECART_MUS:
            imus = j + 1
    if imus == -1:
        if xmus > mu[0]:
            imus = 1
        elif xmus < mu[-1]:
            imus = len(mu) + 1
        else:
            for j in range(len(mu) - 1):
                if mu[j + 1] < xmus < mu[j]:
                    imus = j + 2
        mu.insert(imus - 1, xmus)
        w.insert(imus - 1, 0.0)
        flag.insert(imus - 1, 0)
    # D21.14 channel of the angles file
    mu = round_e(np.array(mu), 14)
    w = round_e(np.array(w), 14)
    n = len(mu)
    rmu = np.zeros(2 * n + 1)
    ga = np.zeros(2 * n + 1)
    rmu[n + 1:] = mu
    rmu[:n] = -mu[::-1]
    ga[n + 1:] = w
    ga[:n] = w[::-1]
    return rmu, ga, imus, np.array(flag)


def phase_coefficients(os_nb, g_modes=((0.75, 0.97), (-0.3, 0.03)), pol=0.35, pol_decay=0.5, seed=None):
    """Synthetic truncated-phase-matrix Legendre coefficients alpha, beta, gamma, zeta (0:os_nb).
    beta_l = (2l+1) * sum f_i g_i^l  (mixture of Henyey-Greenstein lobes), beta_0 = 1.
    alpha_l = zeta_l = gamma_l = 0 for l < 2 (the reference reads RSL/TSL(l<2) uninitialised).
    Values go through the E15.8 channel of Aerosols.txt."""
    l = np.arange(os_nb + 1, dtype=np.float64)
    beta = np.zeros(os_nb + 1)
    for g, f in g_modes:
        beta += f * (2 * l + 1) * np.power(g, l)
    beta /= beta[0]
    alpha = 0.92 * beta.copy()
    zeta = 0.88 * beta.copy()
    gamma = -pol * np.sqrt(1.5) * (2 * l + 1) / 5.0 * np.power(pol_decay, np.maximum(l - 2, 0))
    # a small Rayleigh-like l=2 signature
    alpha[:2] = 0.0
    zeta[:2] = 0.0
    gamma[:2] = 0.0
    if seed is not None:
        rng = np.random.default_rng(seed)
        beta[1:] *= 1.0 + 0.02 * rng.standard_normal(os_nb)
    return round_e(alpha, 8), round_e(beta, 8), round_e(gamma, 8), round_e(zeta, 8)


def layering(ttot):
    """Number of layers and per-layer optical thickness per SOS_PROFILE's no-gas rules
    (SOS_PROFIL.F:354-390) extended with the CTE_OS_NT cap.  Returns H levels (0:NT)."""
    if ttot / CTE_OS_NT_MIN <= CTE_TOA_FIRST_LAYER:
        nt = CTE_OS_NT_MIN
        return np.linspace(0.0, ttot, nt + 1)
    if ttot / CTE_OS_NT_MIN < CTE_TCOUCHE:
        nt = CTE_OS_NT_MIN + 1
        t_layer = (ttot - CTE_TOA_FIRST_LAYER) / CTE_OS_NT_MIN
    else:
        nt = int((ttot - CTE_TOA_FIRST_LAYER) / CTE_TCOUCHE)
        nt = min(nt, CTE_OS_NT - 1)
        t_layer = (ttot - CTE_TOA_FIRST_LAYER) / nt
        nt = nt + 1
    h = np.zeros(nt + 1)
    h[1] = CTE_TOA_FIRST_LAYER
    h[2:] = CTE_TOA_FIRST_LAYER + t_layer * np.arange(1, nt)
    return h


def profile(tau_ray, h_ray, tau_aer, h_aer, tau_gas=0.0, h_gas=2.0, z_toa=120.0):
    """Synthetic per-term profile in the PROFIL_TMP layout: z (km), H = total optical depth from TOA,
    PCAER / PCMOL = aerosol-extinction and molecular-scattering fractions of each layer's optical thickness
    (level 0 copies level 1, SOS_PROFIL.F:1081-1085).  Exponential vertical distributions."""
    def t_of_z(z):
        return (tau_ray * np.exp(-z / h_ray) + tau_aer * np.exp(-z / h_aer) + tau_gas * np.exp(-z / h_gas))
    ttot = t_of_z(0.0)
    h = layering(ttot)
    nt = h.size - 1
    z = np.zeros(nt + 1)
    z[0] = z_toa
    if nt > 1:                                   # all levels bisect at once (t_of_z decreasing in z); same arithmetic per level
        lo, hi = np.zeros(nt - 1), np.full(nt - 1, z_toa)
        for _ in range(80):
            mid = 0.5 * (lo + hi)
            above = t_of_z(mid) > h[1:nt]
            lo = np.where(above, mid, lo)
            hi = np.where(above, hi, mid)
        z[1:nt] = 0.5 * (lo + hi)
    z[nt] = 0.0
    tr = tau_ray * np.exp(-z / h_ray)
    ta = tau_aer * np.exp(-z / h_aer)
    tg = tau_gas * np.exp(-z / h_gas)
    tr[0] = ta[0] = tg[0] = 0.0
    dtr, dta, dtg = np.diff(tr), np.diff(ta), np.diff(tg)
    dt = dtr + dta + dtg
    pcaer = np.zeros(nt + 1)
    pcmol = np.zeros(nt + 1)
    pcaer[1:] = dta / dt
    pcmol[1:] = dtr / dt
    pcaer[0], pcmol[0] = pcaer[1], pcmol[1]
    hh = np.concatenate([[0.0], np.cumsum(dt)])
    return round_f(z, 5), round_e(hh, 8), round_e(pcaer, 8), round_e(pcmol, 8)


def synthetic_surface(rmu, nbmu, os_nb, rho_dir=0.05, seed=0):
    """Synthetic BRDF/BPDF Fourier matrices in the surface-file layout (SOS_SURFACE.F:2404-2412):
    surf[s, m, J-1, I-1] = R_m(I,J), REAL*4, I = incidence, J = reflection.  Smooth, reciprocal-like,
    decaying with the Fourier order; used until a cached glitter file is supplied."""
    n = nbmu
    mu = rmu[n + 1:]
    rng = np.random.default_rng(seed)
    surf = np.zeros((os_nb + 1, 9, n, n), dtype=np.float32)
    mi, mj = np.meshgrid(mu, mu, indexing="xy")      # [J, I]
    base = mi * mj
    shape = 1.0 + 0.4 * (1 - mi) * (1 - mj)
    amp = np.array([1.0, -0.12, 0.05, -0.12, 0.35, 0.04, -0.05, -0.04, 0.30])
    jitter = 1.0 + 0.01 * rng.standard_normal((9, 1, 1))
    for s in range(os_nb + 1):
        dec = rho_dir * np.exp(-0.12 * s) / (1.0 + 0.05 * s * s)
        for m in range(9):
            f = shape if m in (0, 4, 8) else (mi - mj) * 0.5 + 0.3 * (1 - mi * mj)
            if s == 0 and m in (2, 5, 6, 7):
                continue                          # sine-type elements vanish for s = 0
            surf[s, m] = (dec * amp[m] * jitter[m] * base * f).astype(np.float32)
    return surf


@dataclass
class Optics:
    """Everything SOS_PREPA_OS hands to SOS for one wavelength (SOS_PREPA_OS.F:328-335) plus the
    per-run scalars of the SOS call (SOS.F:340-345)."""
    nbmu: int
    rmu: np.ndarray
    ga: np.ndarray
    n0: int
    tetas: float
    os_nb: int
    alpha: np.ndarray
    beta: np.ndarray
    gamma: np.ndarray
    zeta: np.ndarray
    a_trunc: float
    piztr: float
    ron: float = CTE_MDF
    rho: float = 0.0
    imat_surf: int = 0
    ifresnel: int = 0
    ind_surf: float = 1.34
    wind: float = 2.0
    igli: int = 0
    surf: Optional[np.ndarray] = None
    igmax: int = 100
    ipolar: int = 1
    zout: float = -1.0

    @property
    def piz(self):
        return self.piztr / (1 + 0.5 * self.a_trunc * (self.piztr - 1))   # SOS_PREPA_OS.F:700


@dataclass
class Term:
    """One (wavelength, CKD term): the per-term profile SOS_PROFILE writes and the CKD weight AIK."""
    optics: int
    aik: float
    zprof: np.ndarray
    h: np.ndarray
    pcaer: np.ndarray
    pcmol: np.ndarray

    @property
    def nt(self):
        return self.h.size - 1


@dataclass
class Workload:
    name: str
    optics: List[Optics] = field(default_factory=list)
    terms: List[Term] = field(default_factory=list)


def _ckd_terms(nterm, rng=None, tau0=(0.003, 0.03, 0.3, 1.5, 6.0)):
    """Gauss-Legendre weights on g in [0,1] and gas optical depths per term."""
    x, w = np.polynomial.legendre.leggauss(nterm)
    w = 0.5 * w
    if rng is None:
        t = np.array(tau0[:nterm])
    else:
        t = np.sort(np.exp(rng.uniform(np.log(1e-3), np.log(30.0), nterm)))
    return w / w.sum(), t


def make_optics(nb_gauss=40, tetas=35.0, os_nb=80, surface="lambert", rho=0.0, a_trunc=0.4, piztr=0.99,
                g_modes=((0.75, 0.97), (-0.3, 0.03)), zout=-1.0, ipolar=1, igmax=100, seed=0, user_angles=()):
    rmu, ga, n0, _ = sos_angles(nb_gauss, tetas, user_angles)
    nbmu = (rmu.size - 1) // 2
    al, be, gm, ze = phase_coefficients(os_nb, g_modes, seed=seed)
    o = Optics(nbmu=nbmu, rmu=rmu, ga=ga, n0=n0, tetas=tetas, os_nb=os_nb, alpha=al, beta=be, gamma=gm, zeta=ze,
               a_trunc=round_f(a_trunc, 5), piztr=round_f(piztr, 5), rho=rho, zout=zout, ipolar=ipolar, igmax=igmax)
    if surface == "brdf":
        o.imat_surf = 1
        o.surf = synthetic_surface(rmu, nbmu, os_nb, seed=seed)
    elif surface == "fresnel":
        o.ifresnel = 1
    elif surface == "glitter":
        o.imat_surf = 1
        o.igli = 1
        o.surf = synthetic_surface(rmu, nbmu, os_nb, seed=seed)   # replaced by the glitter pipeline when available
    return o


def rayleigh_tau(wa_um):
    """CNES Rayleigh optical depth formula (SOS_PROC.F:3333-3334 shape): 0.008569 l^-4 (1 + 0.0113 l^-2 + 0.00013 l^-4)."""
    return 0.008569 * wa_um ** -4 * (1 + 0.0113 * wa_um ** -2 + 0.00013 * wa_um ** -4)


def config_demo(polar_plane=False, nterm=5, nb_gauss=40, os_nb=80, surface="glitter"):
    """configs[0]/[1]: demo-like single wavelength 0.910 um, Rayleigh + aerosol, 5 synthetic H2O CKD terms."""
    wl = Workload("demoPolar" if polar_plane else "demo")
    wl.optics.append(make_optics(nb_gauss=nb_gauss, tetas=35.0, os_nb=os_nb, surface=surface, rho=0.0))
    w, t0 = _ckd_terms(nterm)
    for k in range(nterm):
        z, h, xa, ym = profile(rayleigh_tau(0.910), 8.0, 0.25, 2.0, t0[k], 2.0)
        wl.terms.append(Term(0, float(w[k]), z, h, xa, ym))
    return wl


def config_ckd_band(npoints=25, seed=20261021, nb_gauss=40, os_nb=80, surface="lambert", rho=0.1, max_terms=25):
    """configs[2]: O2 A-band like CKD band: npoints spectral points, 1..25 terms each, log-uniform gas depths."""
    rng = np.random.default_rng(seed)
    wl = Workload("ckd_band_%d" % npoints)
    base = make_optics(nb_gauss=nb_gauss, tetas=35.0, os_nb=os_nb, surface=surface, rho=rho)
    for p in range(npoints):
        wl.optics.append(base)                      # same angles / aerosol model across the narrow band
        n1 = int(rng.choice([1, 2, 5]))
        n2 = int(rng.choice([1, 1, 5])) if max_terms >= 25 else 1
        w1, t1 = _ckd_terms(n1, rng)
        w2, t2 = _ckd_terms(n2, rng)
        for a in range(n1):
            for b in range(n2):
                z, h, xa, ym = profile(rayleigh_tau(0.765), 8.0, 0.20, 2.0, t1[a] + 0.2 * t2[b], 4.0)
                wl.terms.append(Term(p, float(w1[a] * w2[b]), z, h, xa, ym))
    return wl


def config_hyperspectral(nwave=2100, seed=20261022, nb_gauss=24, os_nb=80, max_terms=5):
    """configs[3]: sweep 0.4-2.5 um, bimodal aerosol, BRDF/BPDF surface matrix (synthetic cache file)."""
    rng = np.random.default_rng(seed)
    wl = Workload("hyperspectral_%d" % nwave)
    nus = np.linspace(4000.0, 25000.0, nwave)
    for p, nu in enumerate(nus):
        lam = 1.0e4 / nu
        gfine = 0.60 - 0.05 * (lam - 0.55)
        o = make_optics(nb_gauss=nb_gauss, tetas=35.0, os_nb=os_nb, surface="brdf", rho=0.1,
                        g_modes=((0.85, 0.35), (gfine, 0.65)), seed=p)
        wl.optics.append(o)
        nt_ = int(rng.integers(1, max_terms + 1))
        w, t0 = _ckd_terms(nt_, rng)
        ta = 0.3 * (lam / 0.55) ** -1.3
        for k in range(nt_):
            z, h, xa, ym = profile(rayleigh_tau(lam), 8.0, ta, 2.0, t0[k], 3.0)
            wl.terms.append(Term(p, float(w[k]), z, h, xa, ym))
    return wl


def config_angular_stress(nterm=64, seed=20261023, nb_gauss=79, os_nb=200):
    """configs[4]: N=80 angles, OS_NB=200, forward-peaked coarse mode, ~200 layers, output at altitude."""
    rng = np.random.default_rng(seed)
    wl = Workload("angular_stress")
    o = make_optics(nb_gauss=nb_gauss, tetas=35.0, os_nb=os_nb, surface="lambert", rho=0.05, a_trunc=0.0,
                    g_modes=((0.92, 0.70), (0.60, 0.30)), zout=3.0)
    wl.optics.append(o)
    w, t0 = _ckd_terms(8, rng)
    for k in range(nterm):
        tg = float(t0[k % 8]) * 0.05
        z, h, xa, ym = profile(rayleigh_tau(0.55), 8.0, 0.9, 2.0, tg, 3.0)
        wl.terms.append(Term(0, 1.0 / nterm, z, h, xa, ym))
    return wl
