"""Keyword-driven front end: what SOS_ABS_MAIN + SOS_PROC do for ONE wavelength from the command-line keywords
(SOS_ABS_MAIN.F:213-912, SOS_PROC.F:1218-3874), for a LIST of wavelengths with every stage on the device:

    angles        SOS_ANGLES            synth.sos_angles (radiance angles + solar angle), mie_angles (phase-function angles)   host
    aerosols      SOS_AEROSOLS          aerosols.run: all wavelengths in one device call
    surface       SOS_SURFACE           Solver.glitter / roujean / surface_bpdf / bpdf_ajout_brdf                              device
    profiles      SOS_PROFIL (+ gas)    band.run_band -> Solver.profile / profile_chain                                        device
    solves        SOS, SOS_OS, SOS_AGGREGATE, SOS_TRPHI_OPTION   band.run_band                                                 device
    files         SOS_Up / SOS_Down / SOS_Result.bin / Trans / Flux, aerosol result file                                       host

Supported keyword values: -SURF.Type 0 1 2 3 4 5 7 (6, Nadal: refused as SOS_PROC.F:2210-2226 refuses it; its surface file is
Solver.surface_nadal); -AER.Model 0 1 2 3 4 5 (5: up to four modes);
-AP.AbsProfile.Type 7 (no gaseous absorption), 0 (user profile file) and 1 .. 6 (predefined atmospheres: absprofile.py reads their
tables from the user's installation of the reference, $SOS_ABS_ROOT/src/SOS_SUB_TRS.F, and the CKD coefficients from
$SOS_ABS_ROOT/fic/COEFF_CKD), or the gas atmosphere and CKD tables handed in by the caller (`gas=`); user angle files
(-ANG.Rad.UserAngFile, -ANG.Aer.UserAngFile) with their output files (-SOS.ResFileUp.UserAng, -SOS.ResFileDown.UserAng); a user
aerosol file (-AER.UserFile) and a user surface matrix file (-SURF.File) in the layouts of the reference's own files.  Unsupported values raise NotImplementedError naming the keyword; nothing is silently replaced.  No CPU fallback."""
import os

import numpy as np

from . import absprofile, aerosols, api, band, keywords, synth
from .formats import round_e

DEFAULT_NBMU_LUM, DEFAULT_NBMU_MIE = 24, 40                      # inc/SOS.h:508-514
DEFAULT_OS_NB, DEFAULT_OS_NS, DEFAULT_OS_NM = 80, 48, 128        # inc/SOS.h:521-535
HT_STD_PSURF = 1013.0                                            # CTE_HT_STD_PSURF (inc/SOS.h:192)


def expansion_orders(nb_gauss_mie, nb_gauss_lum):
    """OS_NB, OS_NS, OS_NM from the numbers of Gauss angles as given on the command line (None = keyword absent)
    (SOS_ANGLES.F:303-329)."""
    os_nb = DEFAULT_OS_NB if nb_gauss_mie is None else 2 * nb_gauss_mie
    if nb_gauss_lum is None:
        os_ns, os_nm = DEFAULT_OS_NS, DEFAULT_OS_NM
    else:
        os_ns = 2 * nb_gauss_lum
        os_nm = os_nb + os_ns
    if os_nm < os_nb + os_ns:
        raise ValueError("OS_NM < OS_NB + OS_NS (SOS_ANGLES error 947)")
    return os_nb, os_ns, os_nm


def read_user_angles(path):
    """A user angle file (-ANG.Rad.UserAngFile / -ANG.Aer.UserAngFile): one zenith angle in degrees per line, 0 <= angle <= 90
    (SOS_ANGLES.F:776-786, list-directed READ of one value per record)."""
    out = []
    with open(path) as f:
        for ln in f:
            if not ln.strip():
                continue
            v = float(ln.replace(",", " ").split()[0].replace("D", "E").replace("d", "e"))
            if v < 0.0 or v > 90.0:
                raise ValueError("user angle %g outside [0, 90] degrees (SOS_ANGLES error 972)" % v)
            out.append(v)
    return out


def mie_angles(nb_gauss, user_deg=()):
    """Angles of the phase-function calculations (SOS_ANGLES_GAUSS_USER with USAGE = 'MIE', SOS_ANGLES.F:713-860): the Gauss angles
    and the user angles (weight 0) in ascending mu, through the D21.14 fields of the angles file SOS_AEROSOLS reads back (format
    300, :639)."""
    mu, w = synth.sos_gauss(nb_gauss + 1)
    mu = list(mu[:nb_gauss]) + [float(np.cos(a * (np.arccos(-1.0) / 180.0))) for a in user_deg]
    w = list(w[:nb_gauss]) + [0.0] * len(user_deg)
    if len(mu) > 100:
        raise ValueError("more than CTE_MIE_NBMU_MAX = 100 angles (SOS_ANGLES error 980)")
    order = np.argsort(mu, kind="stable")
    mu, w = round_e(np.asarray(mu)[order], 14), round_e(np.asarray(w)[order], 14)
    n = len(mu)
    xmu, xhr = np.zeros(2 * n + 1), np.zeros(2 * n + 1)
    xmu[n + 1:], xhr[n + 1:] = mu, w
    xmu[:n], xhr[:n] = -mu[::-1], w[::-1]
    return n, xmu, xhr


def rayleigh_thickness(psurf, wa):
    """CNES formulation (Perbos, 1982) of the molecular optical thickness (SOS_PROC.F:3331-3334)."""
    return (psurf / HT_STD_PSURF) * 1.e-4 * (float(np.float32(84.35)) / wa ** 4 - float(np.float32(1.225)) / wa ** 5
                                             + float(np.float32(1.4)) / wa ** 6)


def aerosol_model(kw, wavelengths=None):
    """-AER.* keywords -> a model of aerosols.py (wavelengths: those of the run, where the reference's rules depend on them)."""
    m = kw.get("-AER.Model")
    waref = kw.get("-AER.Waref")
    only_ref = wavelengths is not None and waref is not None and [float(w) for w in wavelengths] == [float(waref)]
    if m == 0:
        # the refractive index at the reference wavelength (-AER.MMD.MRwaref / MIwaref) is what SOS_PROC hands to SOS_AEROSOLS in its
        # call at that wavelength; required when the simulation wavelength differs (SOS_PROC.F:1702-1711, error 2314)
        rn, in_ = kw["-AER.MMD.MRwa"], kw["-AER.MMD.MIwa"]
        if not only_ref:
            if "-AER.MMD.MRwaref" not in kw or "-AER.MMD.MIwaref" not in kw:
                if wavelengths is not None:
                    raise ValueError("-AER.Model 0 with a simulation wavelength other than -AER.Waref requires -AER.MMD.MRwaref and "
                                     "-AER.MMD.MIwaref (SOS_PROC error 2314)")
            else:
                rn = (lambda wa, r=kw["-AER.MMD.MRwaref"], s_=rn: r if wa == waref else s_)
                in_ = (lambda wa, r=kw["-AER.MMD.MIwaref"], s_=in_: r if wa == waref else s_)
        sd = kw.get("-AER.MMD.SDtype")
        if sd == 1:
            return aerosols.MonoModal(rn, in_, 1, kw["-AER.MMD.LNDradius"], kw["-AER.MMD.LNDvar"])
        if sd == 2:
            return aerosols.MonoModal(rn, in_, 2, kw["-AER.MMD.JD.rmin"], kw["-AER.MMD.JD.slope"],
                                      kw.get("-AER.MMD.JD.rmax", 50.0))             # CTE_DEFAULT_AER_JUNGE_RMAX
        raise ValueError("-AER.MMD.SDtype must be 1 or 2")
    if m == 1:
        root = os.environ.get("SOS_ABS_ROOT", "")
        data = os.path.join(root, "fic", "Data_WMO_cor_2015_12_16")                  # CTE_AER_DATAWMO
        return aerosols.Wmo(data, kw["-AER.WMO.Model"],
                            [kw.get(k, 0.0) for k in ("-AER.WMO.DL", "-AER.WMO.WS", "-AER.WMO.OC", "-AER.WMO.SO")])
    if m == 2:
        return aerosols.ShettleFenn(os.path.join(os.environ.get("SOS_ABS_ROOT", ""), "fic"), kw["-AER.SF.Model"], kw["-AER.SF.RH"])
    if m == 3:
        b = aerosols.BimodalLnd(kw["-AER.BMD.CM.MRwa"], kw["-AER.BMD.CM.MIwa"], kw["-AER.BMD.CM.SDradius"], kw["-AER.BMD.CM.SDvar"],
                                kw["-AER.BMD.FM.MRwa"], kw["-AER.BMD.FM.MIwa"], kw["-AER.BMD.FM.SDradius"], kw["-AER.BMD.FM.SDvar"])
        if kw.get("-AER.BMD.VCdef") == 1:
            b.cv_coarse, b.cv_fine = kw["-AER.BMD.CoarseVC"], kw["-AER.BMD.FineVC"]
        elif kw.get("-AER.BMD.VCdef") == 2:
            b.rtauct = kw["-AER.BMD.RAOT"]
        else:
            raise ValueError("-AER.BMD.VCdef must be 1 or 2")
        # SOS_PROC's call of SOS_AEROSOLS at the reference wavelength passes the indices of that wavelength (-AER.BMD.*.M?waref) as
        # the mode indices, whatever the definition of the mixture (SOS_PROC.F:2896-2907); when the simulation wavelength is the
        # reference one, the indices of the simulation wavelength are used for both (:1815-1822)
        keys = ("-AER.BMD.CM.MRwaref", "-AER.BMD.CM.MIwaref", "-AER.BMD.FM.MRwaref", "-AER.BMD.FM.MIwaref")
        if all(k in kw for k in keys) and not only_ref:
            ref = tuple(kw[k] for k in keys)
            sim = (b.coarse_rn, b.coarse_in, b.fine_rn, b.fine_in)
            b.coarse_rn, b.coarse_in, b.fine_rn, b.fine_in = (
                (lambda wa, r=r, s=s: r if wa == waref else s) for r, s in zip(ref, sim))
        elif b.rtauct is not None and not only_ref:
            raise ValueError("-AER.BMD.VCdef 2 requires the indices at the reference wavelength, -AER.BMD.CM/FM.MRwaref / MIwaref "
                             "(SOS_PROC error 2329)")
        return b
    if m == 5:
        if "-AER.DefMixture" not in kw:
            raise ValueError("-AER.Model 5 requires -AER.DefMixture (SOS_PROC error 2340)")
        return aerosols.read_mixture_file(kw["-AER.DefMixture"], kw["-AER.Waref"])
    raise ValueError("-AER.Model %r: 0 (mono-modal), 1 (WMO), 2 (Shettle & Fenn), 3 (bimodal log-normal), 4 (external phase functions, "
                     "handled by run), 5 (user mixture)" % (m,))


def surface(solver, kw, nbmu, rmu, ga, os_nb, os_ns, os_nm):
    """-SURF.* keywords -> (fields of synth.Optics, direct-term models of SOS_TRPHI).  The surface records are computed on the
    device for this angle set (the reference caches them in -SURF.Dir; a cache is the caller's business)."""
    t = kw["-SURF.Type"]
    o, direct = dict(rho=kw["-SURF.Alb"]), {}
    if t == 0:
        return o, direct
    if t == 2:
        o.update(ifresnel=1, ind_surf=kw["-SURF.Ind"])
        return o, direct
    user = None
    if "-SURF.File" in kw:
        # a surface file of the user replaces the computation of SOS_SURFACE (SOS_PROC.F:3186-3189); the surface type still selects
        # the direct terms of SOS_TRPHI (SOS_PREPA_OS.F:486-497)
        user = read_surface_file(kw["-SURF.File"], nbmu, os_nb)
    if t == 1:
        surf = user if user is not None else solver.glitter(nbmu, rmu, ga, kw["-SURF.Glitter.Wind"], kw["-SURF.Ind"], os_nb, os_ns, os_nm)[0]
        o.update(imat_surf=1, igli=1, surf=surf, ind_surf=kw["-SURF.Ind"], wind=kw["-SURF.Glitter.Wind"])
        return o, direct
    if t == 6:
        # SOS_PROC.F:2210-2226 prints this and stops before its parameter checks; the surface file of SOS_SURFACE_BPDF with ISURF = 6
        # is available below the front end (Solver.surface_nadal, the direct term through set_direct_models(nadal=...))
        raise ValueError("The Nadal's BPDF model is not supported ==> Select another surface model (-SURF.Type 6, as SOS_PROC)")
    if t in (3, 4, 5, 7):
        k = (kw["-SURF.Roujean.K0"], kw["-SURF.Roujean.K1"], kw["-SURF.Roujean.K2"])
        surf = user if user is not None else solver.roujean(nbmu, rmu, os_nb, *k)
        direct["roujean"] = k
        if t in (4, 5, 7):
            if t == 7 and "-SURF.Maignan.C" not in kw:
                raise ValueError("-SURF.Type 7 requires -SURF.Maignan.C (SOS_PROC.F:2238-2240)")
            if user is None:
                bpdf = solver.surface_bpdf(t, nbmu, rmu, ga, kw["-SURF.Ind"], os_nb, os_ns, os_nm, coef_c=kw.get("-SURF.Maignan.C", 0.0))
                surf = solver.bpdf_ajout_brdf(bpdf, surf)
            if t == 7:
                direct["maignan"] = kw["-SURF.Maignan.C"]
            else:
                direct["irondeaux" if t == 4 else "ibreon"] = 1
            o["ind_surf"] = kw["-SURF.Ind"]
        o.update(imat_surf=1, surf=surf)
        return o, direct
    raise ValueError("-SURF.Type %r: 0 .. 7 (SOS_PROC.F:2177)" % (t,))


def read_surface_file(path, nbmu, os_nb):
    """-SURF.File: a surface matrix file in the layout SOS_OS reads (SOS_OS.F:916-925: one unformatted record per Fourier order
    0..OS_NB with nine N x N REAL*4 matrices).  The file must have been made for the angle set of the run."""
    import struct
    from .formats import _read_records
    if not os.path.exists(path):
        raise ValueError("-SURF.File %s does not exist (SOS_PREPA_OS error 1020)" % path)
    recs = _read_records(path)
    want = 9 * nbmu * nbmu * 4
    if len(recs) < os_nb + 1 or any(len(r) != want for r in recs[:os_nb + 1]):
        raise ValueError("-SURF.File %s: %d records of %s bytes, expected %d records of %d bytes (%d angles, OS_NB = %d)"
                         % (path, len(recs), sorted({len(r) for r in recs}), os_nb + 1, want, nbmu, os_nb))
    return np.array([np.frombuffer(r, dtype="<f4").reshape(9, nbmu, nbmu) for r in recs[:os_nb + 1]])


def read_user_aerosols(path, wa, os_nb, aot):
    """-AER.UserFile: an aerosol file in the layout of SOS_AEROSOLS' result file, read the way SOS_PREPA_OS.F:669-693 reads it
    (truncation coefficient and truncated albedo after the ':' of lines 4 and 5, OS_NB + 1 list-directed rows of alpha, beta,
    gamma, zeta from line 9).  The optical thickness is -AER.AOTref itself (SOS_PROC.F:1867-1876: the simulation wavelength must
    be the reference one)."""
    with open(path) as f:
        lines = f.read().splitlines()
    num = lambda t: float(t.replace("D", "E").replace("d", "e"))
    try:
        a, piztr = num(lines[3].split(":", 1)[1].split()[0]), num(lines[4].split(":", 1)[1].split()[0])
        rows = np.array([[num(x) for x in ln.replace(",", " ").split()[:4]] for ln in lines[8:8 + os_nb + 1]])
    except (IndexError, ValueError):
        raise ValueError("-AER.UserFile %s: not in the layout of the aerosol result file (SOS_PREPA_OS error 923)" % path)
    if rows.shape != (os_nb + 1, 4):
        raise ValueError("-AER.UserFile %s: %d coefficient rows, expected OS_NB + 1 = %d" % (path, rows.shape[0], os_nb + 1))
    piz = piztr / (1 + 0.5 * a * (piztr - 1))                     # SOS_PREPA_OS.F:700
    return aerosols.AerosolOptics(wa, 0.0, 0.0, piz, piztr, a, 0.0, 1 if a != 0.0 else 0, rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], aot)


def user_angle_file(src, dst, itrphi, nbmu, flags):
    """-SOS.ResFileUp.UserAng / -SOS.ResFileDown.UserAng: the header of the full file and the records of the user angles only
    (IND_ANGOUT = 1), in the order SOS_ABS_MAIN writes them (SOS_ABS_MAIN.F:2303-2420 for -SOS.View 1: angle index N-1 .. 0 of the
    half plane phi + 180, then 0 .. N-1; :2449-2504 for -SOS.View 2: 0 .. N-1 for every azimuth)."""
    head, rows = [], []
    for ln in open(src).read().split("\n"):
        if ln == "":
            continue
        (head if ln.lstrip().startswith("#") else rows).append(ln)
    if itrphi == 1:
        if len(rows) != 2 * nbmu:
            raise ValueError("%s: %d records, expected %d" % (src, len(rows), 2 * nbmu))
        jj = list(range(nbmu - 1, -1, -1)) + list(range(nbmu))
    else:
        if len(rows) % nbmu:
            raise ValueError("%s: %d records are not a multiple of %d angles" % (src, len(rows), nbmu))
        jj = list(range(nbmu)) * (len(rows) // nbmu)
    with open(dst, "w") as f:
        f.write("\n".join(head + [r for r, j in zip(rows, jj) if flags[j] == 1]) + "\n")


def run_keywords(solver, argv, wavelengths=None, gas=None):
    """Runs the simulation the keywords describe, for `wavelengths` (microns; default: the one of -SOS_Main.Wa).  gas: None with
    -AP.AbsProfile.Type 7, else dict(tables=, kdis_ai=, userprofil=, altabs=, ro=, lamb1=[per wavelength]) as band.run_band
    takes them.  Results go to <-SOS_Main.ResRoot>/SOS/<wavelength>/ under the names of -SOS.ResFileUp / -SOS.ResFileDown /
    -SOS.ResBin (+ -SOS.Trans / -SOS.Flux), the aerosol result files to <ResRoot>/AER.  Returns (band.BandResult, [AerosolOptics])."""
    return run(solver, keywords.parse(argv), wavelengths, gas)


def run(solver, kw, wavelengths=None, gas=None):
    """run_keywords on an already parsed keyword dict (keywords.parse, or sos.sos_proc's arguments)."""
    wl = [float(w) for w in (wavelengths if wavelengths is not None else [kw["-SOS_Main.Wa"]])]
    nb_lum, nb_mie = kw.get("-ANG.Rad.NbGauss"), kw.get("-ANG.Aer.NbGauss")
    os_nb, os_ns, os_nm = expansion_orders(nb_mie, nb_lum)
    user_lum = read_user_angles(kw["-ANG.Rad.UserAngFile"]) if "-ANG.Rad.UserAngFile" in kw else []
    user_mie = read_user_angles(kw["-ANG.Aer.UserAngFile"]) if "-ANG.Aer.UserAngFile" in kw else []
    rmu, ga, n0, flags = synth.sos_angles(nb_lum or DEFAULT_NBMU_LUM, kw["-ANG.Thetas"], user_lum)
    nbmu = (rmu.size - 1) // 2
    if nbmu > 80:
        raise ValueError("more than CTE_OS_NBMU_MAX = 80 radiance angles (SOS_ANGLES error 981)")
    for k in ("-SOS.ResFileUp.UserAng", "-SOS.ResFileDown.UserAng"):
        if k in kw and not user_lum:
            raise ValueError("%s requires -ANG.Rad.UserAngFile" % k)             # SOS_ABS_MAIN.F:100-103
    mie_n, xmu, xhr = mie_angles(nb_mie or DEFAULT_NBMU_MIE, user_mie)
    absprofil = kw["-AP.AbsProfile.Type"]
    if absprofil != 7 and kw["-AP.AerProfile.Type"] == 2:
        raise ValueError("an aerosol layer profile (-AP.AerProfile.Type 2) goes with -AP.AbsProfile.Type 7 only (SOS_PROC error 2513)")
    if absprofil != 7 and kw.get("-SOS.AbsModeCKD") not in (1, 2):
        raise ValueError("-AP.AbsProfile.Type %d requires -SOS.AbsModeCKD 1 or 2 (SOS_PROC error 2515)" % absprofil)
    if absprofil != 7 and gas is None:
        # SOS_PREPA_ABSPROFILE: the user's profile file (type 0) or a predefined atmosphere (1 .. 6, tables read from the user's
        # installation of the reference), the CKD coefficient files of $SOS_ABS_ROOT/fic/COEFF_CKD
        if "-AP.SpectralResol" not in kw:
            raise ValueError("-AP.AbsProfile.Type %d requires -AP.SpectralResol" % absprofil)
        nd = absprofile.NOT_DEFINED
        gas = absprofile.prepare(getattr(solver, "lib", None) or api.load_library(), wl, kw["-AP.SpectralResol"], absprofil,
                                 kw.get("-AP.AbsProfile.UserFile"), kw.get("-AP.Psurf", nd), kw.get("-AP.H2O", nd), kw.get("-AP.O3", nd),
                                 kw.get("-AP.CO2", nd), kw.get("-AP.CH4", nd))
    # ---- aerosols: all wavelengths in one device call ----
    aot_ref = kw.get("-AER.AOTref", 0.0)
    user_aer = aot_ref > 0.0 and "-AER.UserFile" in kw
    if user_aer:                                                  # no aerosol processing (SOS_PROC.F:1867-1876, 2929-2933)
        if wl != [float(kw["-AER.Waref"])]:
            raise ValueError("-AER.UserFile: the simulation wavelength must be the reference wavelength -AER.Waref (SOS_PROC error 2350)")
        if "-AER.ResFile" in kw:
            raise ValueError("-AER.UserFile and -AER.ResFile exclude each other (SOS_PROC error 2351)")
        aer = [read_user_aerosols(kw["-AER.UserFile"], wl[0], os_nb, aot_ref)]
    elif aot_ref > 0.0 and kw.get("-AER.Model") == 4:             # external phase functions (SOS_PROC.F:1837-1847)
        if "-AER.ExtData" not in kw:
            raise ValueError("-AER.Model 4 requires -AER.ExtData (SOS_PROC error 2330)")
        if wl != [float(kw["-AER.Waref"])]:
            raise ValueError("-AER.Model 4: the simulation wavelength must be the reference wavelength -AER.Waref (SOS_PROC error 2331)")
        aer = [aerosols.external_data(solver, kw["-AER.ExtData"], mie_n, xmu, xhr, os_nb, kw["-AER.Tronca"], wl[0], aot_ref)]
    elif aot_ref > 0.0:
        aer = aerosols.run(solver, mie_n, xmu, xhr, os_nb, aerosol_model(kw, wl), wl, waref=kw["-AER.Waref"], aot_ref=aot_ref,
                           itronc=kw["-AER.Tronca"])
    else:                                                         # no aerosols: PIZ = 0, coefficients 0 (SOS_AEROSOLS.F:1136-1139)
        z = np.zeros(os_nb + 1)
        aer = [aerosols.AerosolOptics(w, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, z, z, z, z, 0.0) for w in wl]
    # ---- surface: once (the keywords carry one refractive index for all wavelengths) ----
    sf, direct = surface(solver, kw, nbmu, rmu, ga, os_nb, os_ns, os_nm)
    solver.set_direct_models(**direct)
    itype = kw["-AP.AerProfile.Type"]
    waves = []
    for k, (w, a) in enumerate(zip(wl, aer)):
        # a user file is read as it is (list-directed READ of SOS_PREPA_OS); computed coefficients go through the result file's formats
        f = (dict(alpha=a.alpha, beta=a.beta, gamma=a.gamma, zeta=a.zeta, a_trunc=a.coef_tronca, piztr=a.piztr) if user_aer
             else aerosols.through_result_file(a))
        o = synth.Optics(nbmu=nbmu, rmu=rmu.copy(), ga=ga, n0=n0, tetas=kw["-ANG.Thetas"], os_nb=os_nb, alpha=f["alpha"], beta=f["beta"],
                         gamma=f["gamma"], zeta=f["zeta"], a_trunc=f["a_trunc"], piztr=f["piztr"], igmax=kw["-SOS.IGmax"],
                         ipolar=kw["-SOS.Ipolar"], zout=kw["-SOS.OutputAlt"], **sf)
        tr = kw["-AP.MOT"] if "-AP.MOT" in kw else rayleigh_thickness(kw.get("-AP.Psurf", HT_STD_PSURF), w)
        waves.append(band.Wavelength(optics=o, lamb1=(gas["lamb1"][k] if gas else 1), tr=tr, ta=a.ta, hr=kw.get("-AP.HR", 8.0),
                                     ha=kw.get("-AP.AerHS.HA", 2.0), absprofil=absprofil, iprofil=itype,
                                     zmin=kw.get("-AP.AerLayer.Zmin", 0.0), zmax=kw.get("-AP.AerLayer.Zmax", 0.0), name="%.6f" % w))
    root = kw["-SOS_Main.ResRoot"]
    g = gas or {}
    res = band.run_band(solver, g.get("tables"), g.get("kdis_ai"), g.get("userprofil"), g.get("altabs"), g.get("ro"), waves,
                        itrphi=kw["-SOS.View"], phios=kw.get("-SOS.View.Phi", 0.0), pas_phi=kw.get("-SOS.View.Dphi", 30),
                        outdir=os.path.join(root, "SOS"), trans="-SOS.Trans" in kw, flux="-SOS.Flux" in kw and gas is not None,
                        ckd_mode=(kw.get("-SOS.AbsModeCKD", 1) if absprofil != 7 else 1))
    # ---- the reference's file names ----
    names = {"SOS_Up.txt": kw["-SOS.ResFileUp"], "SOS_Down.txt": kw["-SOS.ResFileDown"], "SOS_Result.bin": kw["-SOS.ResBin"],
             "SOS_Trans.txt": kw.get("-SOS.Trans"), "SOS_Flux.txt": kw.get("-SOS.Flux")}
    res.ind_angout = np.asarray(flags, dtype=np.int32)          # IND_ANGOUT(1:N): 1 for the user angles
    for d in res.dirs:
        for key, full in (("-SOS.ResFileUp.UserAng", "SOS_Up.txt"), ("-SOS.ResFileDown.UserAng", "SOS_Down.txt")):
            if key in kw:
                user_angle_file(os.path.join(d, full), os.path.join(d, kw[key]), kw["-SOS.View"], nbmu, flags)
        for src, dst in names.items():
            p = os.path.join(d, src)
            if dst and dst != src and os.path.exists(p):
                os.replace(p, os.path.join(d, dst))
    if aot_ref > 0.0 and not user_aer:
        os.makedirs(os.path.join(root, "AER"), exist_ok=True)
        base = kw.get("-AER.ResFile", "Aerosols.txt")
        for w, a in zip(wl, aer):
            name = base if len(wl) == 1 else "%s_%.6f" % (base, w)
            api.write_aerosols(os.path.join(root, "AER", name), os_nb, a.kmat1, a.kmat2, a.asym, a.coef_tronca, a.piztr, a.alpha, a.beta,
                               a.gamma, a.zeta)
    return res, aer
