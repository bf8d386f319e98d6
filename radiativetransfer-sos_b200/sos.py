"""`sos.sos_proc`: the entry binding/run_sos.py calls (the f2py wrapper of SUBROUTINE SOS_PROC, SOS_PROC.F:1218-1237) with the same
argument names, the same "not defined" conventions (-999 / -999.0 / 'NO_...' strings) and the same outputs in the same order and
shapes (tables (361, 81) = X_FIN(0:360, 0:CTE_OS_NBMU_MAX)), on top of the device front end (frontend.run).

    from importlib import import_module
    sos = import_module("radiativetransfer-sos_b200.sos")
    nblum, ind_angout, phi, vza, sca_ang_up, i_up, q_up, u_up, pol_ang_up, pol_rate_up, l_pol_up, \\
        sca_ang_down, i_down, q_down, u_down, pol_ang_down, pol_rate_down, l_pol_down, \\
        flux_dir_down, flux_diff_down, flux_tot_down, flux_diff_up, coef_tronca = sos.sos_proc(resroot=..., wa_simu=..., ...)

What frontend.run does not support raises NotImplementedError (the reference sets IER = -1 and prints); no CPU fallback."""
import numpy as np

from . import api, frontend, keywords, synth

NOT_DEFINED = -999
NBMU_MAX = 80                                       # CTE_OS_NBMU_MAX (inc/SOS.h:471)

# argument of SOS_PROC (lower case, as f2py exposes it) -> keyword of SOS_ABS_MAIN; order of SOS_PROC.F:1218-1237
ARGS = (
    ("resroot", "-SOS_Main.ResRoot"), ("ficmain_log", "-SOS_Main.Log"), ("wa_simu", "-SOS_Main.Wa"), ("nbmu_gauss_lum", "-ANG.Rad.NbGauss"),
    ("ficangles_user_lum", "-ANG.Rad.UserAngFile"), ("tetas", "-ANG.Thetas"), ("ficangles_res_lum", "-ANG.Rad.ResFile"),
    ("nbmu_gauss_mie", "-ANG.Aer.NbGauss"), ("ficangles_user_mie", "-ANG.Aer.UserAngFile"), ("ficangles_res_mie", "-ANG.Aer.ResFile"),
    ("ficanglog", "-ANG.Log"), ("waref_aot", "-AER.Waref"), ("aot_ref", "-AER.AOTref"), ("itronc_aer", "-AER.Tronca"),
    ("ficgranu_log", "-AER.Log"), ("ficmie_log", "-AER.MieLog"), ("dir_mie", "-AER.DirMie"), ("ficgranu", "-AER.ResFile"),
    ("imod_aer", "-AER.Model"), ("rn_wa", "-AER.MMD.MRwa"), ("in_wa", "-AER.MMD.MIwa"), ("rn_waref", "-AER.MMD.MRwaref"),
    ("in_waref", "-AER.MMD.MIwaref"), ("igranu", "-AER.MMD.SDtype"), ("lnd_radius_mmd_aer", "-AER.MMD.LNDradius"),
    ("lnd_lnvar_mmd_aer", "-AER.MMD.LNDvar"), ("jd_slope_mmd_aer", "-AER.MMD.JD.slope"), ("jd_rmin_mmd_aer", "-AER.MMD.JD.rmin"),
    ("jd_rmax_mmd_aer", "-AER.MMD.JD.rmax"), ("imodele_wmo", "-AER.WMO.Model"), ("c_wmo_dl", "-AER.WMO.DL"), ("c_wmo_ws", "-AER.WMO.WS"),
    ("c_wmo_oc", "-AER.WMO.OC"), ("c_wmo_so", "-AER.WMO.SO"), ("imodele_sf", "-AER.SF.Model"), ("rh", "-AER.SF.RH"),
    ("mode_param_bilnd", "-AER.BMD.VCdef"), ("user_cv_coarse", "-AER.BMD.CoarseVC"), ("user_cv_fine", "-AER.BMD.FineVC"),
    ("rtauct_waref", "-AER.BMD.RAOT"), ("bmd_cm_mrwa", "-AER.BMD.CM.MRwa"), ("bmd_cm_miwa", "-AER.BMD.CM.MIwa"),
    ("bmd_cm_mrwaref", "-AER.BMD.CM.MRwaref"), ("bmd_cm_miwaref", "-AER.BMD.CM.MIwaref"), ("bmd_cm_rmodal", "-AER.BMD.CM.SDradius"),
    ("bmd_cm_var", "-AER.BMD.CM.SDvar"), ("bmd_fm_mrwa", "-AER.BMD.FM.MRwa"), ("bmd_fm_miwa", "-AER.BMD.FM.MIwa"),
    ("bmd_fm_mrwaref", "-AER.BMD.FM.MRwaref"), ("bmd_fm_miwaref", "-AER.BMD.FM.MIwaref"), ("bmd_fm_rmodal", "-AER.BMD.FM.SDradius"),
    ("bmd_fm_var", "-AER.BMD.FM.SDvar"), ("ficextdata_aer", "-AER.ExtData"), ("ficmixture_aer", "-AER.DefMixture"),
    ("ficuser_aer", "-AER.UserFile"), ("ficprofil_log", "-AP.Log"), ("tr", "-AP.MOT"), ("hr", "-AP.HR"), ("ha", "-AP.AerHS.HA"),
    ("iprofil", "-AP.AerProfile.Type"), ("zmin", "-AP.AerLayer.Zmin"), ("zmax", "-AP.AerLayer.Zmax"), ("psurf", "-AP.Psurf"),
    ("h2o", "-AP.H2O"), ("o3", "-AP.O3"), ("co2", "-AP.CO2"), ("ch4", "-AP.CH4"), ("absprofil", "-AP.AbsProfile.Type"),
    ("ficabsprofil", "-AP.AbsProfile.UserFile"), ("nustep", "-AP.SpectralResol"), ("isurf", "-SURF.Type"), ("dir_surf", "-SURF.Dir"),
    ("ficsurf_log", "-SURF.Log"), ("surf_ind", "-SURF.Ind"), ("wind", "-SURF.Glitter.Wind"), ("k0_roujean", "-SURF.Roujean.K0"),
    ("k1_roujean", "-SURF.Roujean.K1"), ("k2_roujean", "-SURF.Roujean.K2"), ("alpha_nadal", "-SURF.Nadal.Alpha"),
    ("beta_nadal", "-SURF.Nadal.Beta"), ("coef_c_maignan", "-SURF.Maignan.C"), ("rho", "-SURF.Alb"), ("ficsurf", "-SURF.File"),
    ("ficsos_log", "-SOS.Log"), ("ficsos_res_bin", "-SOS.ResBin"), ("fictrans", "-SOS.Trans"), ("ficflux", "-SOS.Flux"), ("zout", "-SOS.OutputAlt"),
    ("igmax", "-SOS.IGmax"), ("ipolar", "-SOS.Ipolar"), ("itrphi", "-SOS.View"), ("phios", "-SOS.View.Phi"), ("pas_phi", "-SOS.View.Dphi"),
    ("imode_ckd_calcul", "-SOS.AbsModeCKD"), ("ier", None), ("trace", None),
)
_ALIASES = {"imodel_wmo": "imodele_wmo", "alpha_nada": "alpha_nadal", "beta_nada": "beta_nadal", "fic_trans": "fictrans", "fic_flux": "ficflux"}
_LOGS = {"-SOS_Main.Log", "-ANG.Log", "-AER.Log", "-AER.MieLog", "-AP.Log", "-SURF.Log", "-SOS.Log", "-ANG.Rad.ResFile", "-ANG.Aer.ResFile"}


def _defined(v):
    if v is None:
        return False
    if isinstance(v, (bytes, str)):
        s = v.decode() if isinstance(v, bytes) else v
        return bool(s.strip()) and not s.strip().startswith("NO_")
    return float(v) != float(NOT_DEFINED)


def to_keywords(*args, **kwargs):
    """SOS_PROC's arguments (positional in the order of ARGS, or by name) -> the keyword dict of keywords.parse."""
    names = [a for a, _ in ARGS]
    if len(args) > len(names):
        raise TypeError("sos_proc takes at most %d arguments" % len(names))
    given = dict(zip(names, args))
    for k, v in kwargs.items():
        k = _ALIASES.get(k, k)
        if k not in names:
            raise TypeError("sos_proc got an unexpected argument %r" % k)
        if k in given:
            raise TypeError("sos_proc got argument %r twice" % k)
        given[k] = v
    kw = dict(keywords.DEFAULTS)
    for name, key in ARGS:
        if key is None or name not in given or not _defined(given[name]):
            continue
        v = given[name]
        t = keywords.KEYWORDS[key]
        kw[key] = (v.decode() if isinstance(v, bytes) else str(v)).strip() if t == keywords.S else (int(v) if t == keywords.I else float(v))
    for k in _LOGS:                                       # log / intermediate angle files are not written by the device front end
        kw.pop(k, None)
    missing = [k for k in keywords.REQUIRED if k not in kw]
    if missing:
        raise ValueError("sos_proc: not defined: " + " ".join(missing))
    return kw


def sos_proc(*args, solver=None, gas=None, **kwargs):
    """SOS_PROC.  solver: an api.Solver to reuse (default: one is created for the call); gas: the gas atmosphere and CKD tables when
    absprofil is not 7 (frontend.run).  Returns the 23 outputs of the f2py wrapper."""
    kw = to_keywords(*args, **kwargs)
    own = solver is None
    s = api.Solver(0) if own else solver
    try:
        res, aer = frontend.run(s, kw, None, gas)
    finally:
        if own:
            s.close()
    nb_lum = kw.get("-ANG.Rad.NbGauss") or frontend.DEFAULT_NBMU_LUM
    user = frontend.read_user_angles(kw["-ANG.Rad.UserAngFile"]) if "-ANG.Rad.UserAngFile" in kw else []
    rmu, ga, n0, flags = synth.sos_angles(nb_lum, kw["-ANG.Thetas"], user)
    n = (rmu.size - 1) // 2
    ind = np.zeros(NBMU_MAX + 1, dtype=np.int32)          # IND_ANGOUT_FIN(0:80): 1 where the angle is a user angle
    ind[:n] = flags
    phi, vza = np.zeros(361), np.zeros(NBMU_MAX + 1)
    tabs = [np.zeros((361, NBMU_MAX + 1)) for _ in range(14)]
    itrphi = kw["-SOS.View"]
    if itrphi == 1:
        phi[0] = kw.get("-SOS.View.Phi", 0.0)             # only PHI_FIN(0) is assigned (SOS_TRPHI.F:445, 504)
    else:
        phi[:res.nphi] = np.arange(res.nphi) * kw.get("-SOS.View.Dphi", 30)
    vza[:n] = np.degrees(np.arccos(np.asarray(rmu)[n + 1:2 * n + 1]))
    for t in range(7):
        tabs[t][:res.nphi, :n] = res.up[0, t, :res.nphi, :n]
        tabs[7 + t][:res.nphi, :n] = res.down[0, t, :res.nphi, :n]
    g = res.groups
    tdir, fdd, fd = api.write_flux("NO_OUTPUT", kw["-ANG.Thetas"], float(g.ttot_tronc[0]), float(g.ttot_vrai[0]), float(g.emoins[0]),
                                   float(g.eplus[0]), 0.0, 8.0, 0.0, 2.0, np.zeros(50), np.zeros(50))
    # COEF_TRONCA is an output of the SOS_AEROSOLS calls (SOS_PROC.F:2912, 3047): with a user aerosol file there is none and the
    # wrapper hands back its initial 0, not the coefficient read from the file
    ct = 0.0 if "-AER.UserFile" in kw else float(aer[0].coef_tronca)
    return (n, ind, phi, vza, *tabs, tdir, fdd, fd, float(g.eplus[0]), ct)
