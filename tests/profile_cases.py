"""Synthetic inputs of the per-term profile chain (SOS_ABSPROFILE -> SOS_PROFILE): a 50-level gas atmosphere, CKD tables in
the storage READ_CKD_COEFF fills (Fortran order, the extents of inc/SOS.h:246-282), term lists, and the same tables written as
files in the format READ_CKD_COEFF parses.  No reference data is involved: every number is generated here from a seed, so
the GPU box (which has no /root/reference) builds identical inputs."""
import os

import numpy as np

NBABS, NLEV, NCOL, NWVL, NAI, NTMAX, NPMAX, NCMAX, NT_MAX = 8, 50, 13, 50, 5, 9, 31, 12, 600
GAS = ["H2O", "CO2", "O3", "N2O", "CO", "CH4", "O2", "NO2"]
VMR = np.array([0.0, 400.0, 0.3, 0.32, 0.1, 1.8, 209000.0, 0.0002]) * 1e-6      # volume mixing ratios (H2O from its profile)


def gas_atmosphere(seed=0):
    """USERPROFIL(50,13) (level 1 = ground), ALTABS(50) (descending), RO(8,50) (particles / cm2 per layer, layer index as
    USERPROFIL's lower level)."""
    rng = np.random.default_rng(seed)
    alt = np.concatenate([np.arange(0, 25.0, 1.0), np.arange(25.0, 50.0, 2.5), np.arange(50.0, 121.0, 5.0)])[:NLEV]
    assert alt.size == NLEV and alt[-1] == 120.0
    p = 1013.0 * np.exp(-alt / (7.2 + 0.3 * rng.random()))
    t = np.where(alt < 11, 288.0 - 6.5 * alt, np.where(alt < 20, 216.5, np.where(alt < 47, 216.5 + 1.9 * (alt - 20), 268.0 - 2.2 * (alt - 47))))
    t = np.maximum(t, 175.0) + rng.normal(0, 1.5, NLEV)
    h2o_ppmv = 12000.0 * np.exp(-alt / 2.2) + 4.0
    user = np.zeros((NLEV, NCOL), order="F")
    user[:, 0], user[:, 1], user[:, 2] = alt, p, t
    user[:, 3] = h2o_ppmv * 1e6                                    # the file unit is ppmv * 1e6 (SOS_ABSPROFILE.F:336)
    for k in range(1, NBABS):
        col = {1: 4, 2: 5, 3: 6, 4: 7, 5: 8, 6: 9, 7: 11}[k]
        user[:, col] = VMR[k] * 1e6 * 1e6
    ro = np.zeros((NBABS, NLEV), order="F")
    air = 2.15e25                                                  # molecules / cm2 of the whole column
    for j in range(NLEV - 1):                                      # layer between levels j and j+1 (0-based, ground first)
        frac = (p[j] - p[j + 1]) / 1013.0
        ro[0, j] = air * frac * 0.5 * (h2o_ppmv[j] + h2o_ppmv[j + 1]) * 1e-6
        for k in range(1, NBABS):
            ro[k, j] = air * frac * VMR[k]
    return user, np.ascontiguousarray(alt[::-1]), ro


def ckd_tables(seed=0, nlamb=NWVL, rough=True):
    """Tables with smooth positive k(P, T[, C]); interval l carries a total column optical thickness between about 1e-3 and 40
    for the strongest exponential so that the weak, intermediate and saturated (> 1.5) branches of SOS_PROFILE all occur."""
    rng = np.random.default_rng(seed)
    tab_temp = np.zeros(NTMAX); tab_temp[:] = 160.0 + 20.0 * np.arange(NTMAX)
    tab_pres = np.zeros(NPMAX); tab_pres[:] = 0.007 * (1100.0 / 0.007) ** (np.arange(NPMAX) / (NPMAX - 1.0))
    tab_conc = np.zeros(NCMAX); tab_conc[:] = np.concatenate([[0.0], 10.0 ** np.linspace(1.0, 4.7, NCMAX - 1)])
    nexp = np.ones((NBABS, NWVL), dtype=np.int32, order="F")
    ai = np.zeros((NAI, NBABS, NWVL), order="F")
    ki = np.zeros((NTMAX, NPMAX, NAI, NBABS, NWVL), order="F")
    kh = np.zeros((NTMAX, NPMAX, NCMAX, NAI, NWVL), order="F")
    col = np.array([2.15e25 * 3e-3, *(2.15e25 * VMR[1:])])         # rough column amounts
    tt, pp = tab_temp[:, None], tab_pres[None, :]
    for l in range(nlamb):
        active = rng.random(NBABS) < 0.45
        active[6] = True                                           # O2 always (the A band)
        tau_max = 10.0 ** rng.uniform(-3.0, 1.6)
        for k in range(NBABS):
            if not active[k]:
                ai[0, k, l] = 1.0
                continue
            n = int(rng.integers(1, NAI + 1))
            nexp[k, l] = n
            w = rng.random(n) + 0.2
            ai[:n, k, l] = w / w.sum()
            for i in range(n):
                amp = tau_max * 10.0 ** (-(n - 1 - i) * rng.uniform(0.5, 1.2)) / col[k] * (1.0 if k == 6 else rng.uniform(0.01, 0.3))
                shape = (0.2 + (pp / 1013.0) ** rng.uniform(0.3, 1.0)) * (tt / 296.0) ** rng.uniform(-1.5, 2.5)
                if rough and rng.random() < 0.2:                   # a kink in T: the spline undershoots below zero somewhere
                    shape = shape * np.where(tt < 160.0 + 20.0 * rng.integers(2, 7), 1e-4, 1.0)
                if k == 0:
                    for c in range(NCMAX):
                        kh[:, :, c, i, l] = amp * shape * (1.0 + 0.2 * c / NCMAX)
                else:
                    ki[:, :, i, k, l] = amp * shape
    return dict(nb_temp=NTMAX, nb_pres=NPMAX, nb_conc=NCMAX, tab_temp=tab_temp, tab_pres=tab_pres, tab_conc=tab_conc, nexp=nexp,
                ai=ai, ki=ki, kh=kh)


def make_terms(tables, n, seed=0, iprofil=1):
    """n (wavelength, CKD term) entries: random interval, random exponential per gas, Rayleigh / aerosol parameters."""
    rng = np.random.default_rng(seed + 17)
    out = []
    for i in range(n):
        l = int(rng.integers(1, NWVL + 1))
        ik = [int(rng.integers(1, tables["nexp"][k, l - 1] + 1)) for k in range(NBABS)]
        # 2 (tr + ta) + tau_gas stays below 3: beyond, the profile with gas needs more than CTE_OS_NT = 600 levels and the
        # reference writes past its arrays (it then loops for ever or crashes; libsosgpu.so returns error 9600 instead)
        ta = 0.0 if rng.random() < 0.2 else float(10.0 ** rng.uniform(-2.0, -0.46))
        out.append(dict(lamb1=l, ik=ik, absprofil=7 if rng.random() < 0.1 else 2, iprofil=iprofil,
                        tr=float(rng.uniform(0.01, 0.25)), hr=8.0, ta=ta, ha=float(rng.uniform(1.0, 4.0)),
                        zmin=float(rng.choice([0.0, 1.0, 2.5])), zmax=float(rng.uniform(3.0, 8.0))))
    return out


def write_ckd_files(root, tables, nustep=10, numax=13500):
    """The tables as $root/fic/COEFF_CKD/<step>cmm1/coef_<GAS>_<numax>_<numin>_<step>cmm1 (SOS_SUB_TRS.F:596-650, body :652-760)."""
    d = os.path.join(root, "fic", "COEFF_CKD", "%dcmm1" % nustep)
    os.makedirs(d, exist_ok=True)
    numin = numax - NWVL * nustep
    for k, g in enumerate(GAS):
        with open(os.path.join(d, "coef_%s_%d_%d_%dcmm1" % (g, numax, numin, nustep)), "w") as f:
            for i in range(21 if k == 0 else 18):
                f.write("header line %d\n" % (i + 1))
            f.write("%d %d %d\n" % (numax, numin, nustep))
            f.write("%d\n" % NTMAX + " ".join("%.2f" % v for v in tables["tab_temp"]) + "\n")
            f.write("%d\n" % NPMAX + " ".join("%.17g" % v for v in tables["tab_pres"]) + "\n")
            if k == 0:
                f.write("%d\n" % NCMAX + " ".join("%.17g" % v for v in tables["tab_conc"]) + "\n")
            for l in range(NWVL):
                nu_hi, nu_lo = numax - l * nustep, numax - (l + 1) * nustep
                n = int(tables["nexp"][k, l])
                none = n == 1 and not (tables["kh"][:, :, :, 0, l].any() if k == 0 else tables["ki"][:, :, 0, k, l].any())
                f.write("%.8f %.8f %.8f %.1f %.1f  %d\n" % (1e4 / nu_hi, 2e4 / (nu_hi + nu_lo), 1e4 / nu_lo, nu_hi, nu_lo, 0 if none else n))
                if none:
                    continue
                f.write(" ".join("%.17g" % v for v in tables["ai"][:n, k, l]) + "\n")
                for i in range(n):
                    if k == 0:
                        for c in range(NCMAX):
                            for p in range(NPMAX):
                                f.write("%d %d %d " % (i + 1, c + 1, p + 1) + " ".join("%.17g" % v for v in tables["kh"][:, p, c, i, l]) + "\n")
                    else:
                        for p in range(NPMAX):
                            f.write("%d %d " % (i + 1, p + 1) + " ".join("%.17g" % v for v in tables["ki"][:, p, i, k, l]) + "\n")
    return d
