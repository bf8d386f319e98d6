"""In-process calls into oracle/_ref/libsosref.so (the reference's own Fortran statements, translated to C by
oracle/f77_to_c.py) for the GPU parity tests: SOS_GLITTER, SOS_TRPHI_OPTION, SOS_AGGREGATE, plus the multi-process
term-solve runner.  TEST INFRASTRUCTURE: only tests/ and bench.py's CPU legs come here.

On the GPU box /root/reference does not exist: the library is the prebuilt one that travels with the repository
snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored)."""
import ctypes as C
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MX, NTM, NBM = 80, 600, 200


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def runner():
    return _load("sos_ref_runner", os.path.join(ROOT, "oracle", "ref_runner.py"))


def lib():
    """The reference library: built from /root/reference when that exists (this container), else the prebuilt file."""
    path = os.path.join(ROOT, "oracle", "_ref", "libsosref.so")
    if os.path.isdir("/root/reference"):
        build_ref = _load("sos_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
        path = build_ref.build(verbose=False) or path
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


def _fs(s):
    return C.create_string_buffer(s.encode().ljust(500), 500)


_ip = lambda v: C.byref(C.c_int(int(v)))
_dp = lambda v: C.byref(C.c_double(float(v)))
_P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
_L = C.c_size_t(500)


def _angles(o_rmu, o_ga, N):
    rmu, ga = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    rmu[MX - N:MX + N + 1], ga[MX - N:MX + N + 1] = o_rmu, o_ga
    return rmu, ga


def glitter(ref, fm, tmp, N, rmu, ga, wind, ind, os_nb, os_ns, os_nm):
    """SOS_GLITTER (SOS_GLITTER.F:229) through its files -> REAL*4 records [os_nb+1][9][N][N]."""
    r, g = _angles(rmu, ga, N)
    fgl = os.path.join(tmp, "GLITTER.bin")
    if os.path.exists(fgl):
        os.remove(fgl)
    ier = C.c_int(99)
    ref.sos_glitter_(_ip(N), _P(r), _P(g), _dp(wind), _dp(ind), _ip(os_nb), _ip(os_ns), _ip(os_nm), _fs(os.path.join(tmp, "GSF")),
                     _fs(os.path.join(tmp, "FRESNEL")), _fs(os.path.join(tmp, "MAT_REFLEX")), _fs(fgl), _ip(0), C.byref(ier),
                     _L, _L, _L, _L)
    assert ier.value == 0, "reference SOS_GLITTER IER=%d" % ier.value
    return fm.read_surface_bin(fgl, N)


def roujean(ref, fm, tmp, N, rmu, ga, os_nb, k0, k1, k2):
    """SOS_ROUJEAN (SOS_ROUJEAN.F:212) through its files -> (IER, REAL*4 records [os_nb+1][9][N][N])."""
    r, g = _angles(rmu, ga, N)
    f = os.path.join(tmp, "ROUJEAN.bin")
    if os.path.exists(f):
        os.remove(f)
    ier = C.c_int(99)
    ref.sos_roujean_(_ip(N), _P(r), _P(g), _ip(os_nb), _dp(k0), _dp(k1), _dp(k2), _fs(os.path.join(tmp, "RJ_MAT_REFLEX")), _fs(f),
                     _ip(0), C.byref(ier), _L, _L)
    return ier.value, (fm.read_surface_bin(f, N) if ier.value == 0 else None)


def surface_bpdf(ref, fm, tmp, isurf, N, rmu, ga, ind, os_nb, os_ns, os_nm, coef_c=0.0):
    """SOS_SURFACE_BPDF (SOS_SURFACE_BPDF.F:219) through its files -> REAL*4 records."""
    r, g = _angles(rmu, ga, N)
    f = os.path.join(tmp, "BPDF.bin")
    if os.path.exists(f):
        os.remove(f)
    ier = C.c_int(99)
    ref.sos_surface_bpdf_(_ip(N), _P(r), _P(g), _dp(ind), _ip(isurf), _dp(0.0), _dp(0.0), _dp(coef_c), _ip(os_nb), _ip(os_ns), _ip(os_nm),
                          _fs(os.path.join(tmp, "B_GSF")), _fs(os.path.join(tmp, "B_FRESNEL")), _fs(os.path.join(tmp, "B_MAT_REFLEX")),
                          _fs(f), _ip(0), C.byref(ier), _L, _L, _L, _L)
    assert ier.value == 0, "reference SOS_SURFACE_BPDF IER=%d" % ier.value
    return fm.read_surface_bin(f, N)


def bpdf_ajout_brdf(ref, fm, tmp, surf1, surf2):
    """SOS_BPDF_AJOUT_BRDF (SOS_SURFACE.F:2503) through its files."""
    N, os_nb = surf1.shape[2], surf1.shape[0] - 1
    f1, f2, f3 = (os.path.join(tmp, n) for n in ("AJ1.bin", "AJ2.bin", "AJ3.bin"))
    fm.write_surface_bin(f1, surf1)
    fm.write_surface_bin(f2, surf2)
    if os.path.exists(f3):
        os.remove(f3)
    ier = C.c_int(99)
    ref.sos_bpdf_ajout_brdf_(_fs(f1), _fs(f2), _ip(N), _ip(os_nb), _fs(f3), C.byref(ier), _L, _L, _L)
    assert ier.value == 0, "reference SOS_BPDF_AJOUT_BRDF IER=%d" % ier.value
    return fm.read_surface_bin(f3, N)


def mat_fresnel(ref, tmp, N, rmu, ga, ind, os_ns):
    """SOS_MAT_FRESNEL (SOS_SURFACE.F:1235) -> alpha, beta, gamma, zeta [os_ns+1] read back from its 4(E15.8) text file."""
    r, g = _angles(rmu, ga, N)
    f = os.path.join(tmp, "RES_FRESNEL")
    ier = C.c_int(99)
    ref.sos_mat_fresnel_(_ip(N), _P(r), _P(g), _dp(ind), _ip(os_ns), _fs(f), _ip(0), C.byref(ier), _L)
    assert ier.value == 0, "reference SOS_MAT_FRESNEL IER=%d" % ier.value
    rows = []
    with open(f) as fh:
        for line in fh:
            if line.strip():
                rows.append([float(line[i * 15:(i + 1) * 15]) for i in range(4)])
    a = np.array(rows)
    assert a.shape == (os_ns + 1, 4)
    return a[:, 0], a[:, 1], a[:, 2], a[:, 3]


def trphi_option(ref, fm, tmp, rec, N, rmu, ga, tau, tauout, igli, n0, wind, ind, ifresnel, itrphi, phios, pas, ipolar=1,
                 roujean=None, bpdf=None):
    """SOS_TRPHI_OPTION (SOS_TRPHI.F:285) on a result file -> (nphi, phi, theta, up[7][nphi][N], down[7][nphi][N]).
    roujean = (k0, k1, k2); bpdf = dict(irondeaux=, ibreon=, inadal=, alpha=, beta=, imaignan=, coef=)."""
    fos = os.path.join(tmp, "TRPHI_Result.bin")
    fm.write_result_bin(fos, rec)
    r, g = _angles(rmu, ga, N)
    pf, th = np.zeros(361), np.zeros(MX + 1)
    tabs = [np.zeros((MX + 1, 361)) for _ in range(14)]
    ier = C.c_int(99)
    k0, k1, k2 = roujean if roujean else (0.0, 0.0, 0.0)
    b = dict(irondeaux=0, ibreon=0, inadal=0, alpha=0.0, beta=0.0, imaignan=0, coef=0.0)
    b.update(bpdf or {})
    ref.sos_trphi_option_(_ip(N), _P(r), _P(g), _fs(fos), _dp(tau), _dp(tauout), _dp(-1.0), _ip(igli), _ip(n0), _dp(wind), _dp(ind),
                          _ip(ifresnel), _ip(1 if roujean else 0), _dp(k0), _dp(k1), _dp(k2), _ip(b["irondeaux"]), _ip(b["ibreon"]),
                          _ip(b["inadal"]), _dp(b["alpha"]), _dp(b["beta"]), _ip(b["imaignan"]), _dp(b["coef"]), _ip(itrphi),
                          _dp(phios), _ip(pas), _ip(ipolar), _P(pf), _P(th), *[_P(t) for t in tabs], C.byref(ier), _L)
    assert ier.value == 0, "reference SOS_TRPHI_OPTION IER=%d" % ier.value
    nphi = 2 if itrphi == 1 else 360 // pas + 1
    up = np.array([tabs[t][:N, :nphi].T for t in range(7)])
    down = np.array([tabs[7 + t][:N, :nphi].T for t in range(7)])
    return nphi, pf[:nphi].copy(), th[:N].copy(), up, down


def compare_terms(tr, res, ids, wl, assert_close, what):
    """CUDA term results vs the reference's: returns (count mismatches, terms) after asserting nothing -- the caller
    decides; Stokes are compared only where the counts agree."""
    bad = []
    for n, i in enumerate(ids):
        r = res[i]
        W = 2 * wl.optics[wl.terms[i].optics].nbmu + 1
        nf = r["rec"].shape[0]
        if tr.n_fourier[n] != nf:
            bad.append((i, int(tr.n_fourier[n]), nf))
            continue
        assert_close(tr.rec[n, :nf, :, :W], r["rec"], "%s term %d records" % (what, i))
        assert_close(tr.emoins[n], r["emoins"], "%s term %d EMOINS" % (what, i))
        assert_close(tr.eplus[n], r["eplus"], "%s term %d EPLUS" % (what, i))
        for k in ("ttot_tronc", "ttot_vrai", "tauout"):
            assert getattr(tr, k)[n] == r[k], (what, i, k)
    return bad


# ---- the per-term profile chain ------------------------------------------------------------------------
_IP = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))


def absprofile(ref, tables, user, altabs, ro, term):
    """SOS_ABSPROFILE (SOS_ABSPROFILE.F:184) -> (IER, TAUABSTOT[50]).  INTEGER*2 arguments are int in the translated library."""
    tau = np.zeros(50)
    ier = C.c_int(99)
    iabs = np.ones(8, dtype=np.int32)
    ik = [_ip(v) for v in term["ik"]]
    ref.sos_absprofile_(_ip(term["absprofil"]), _dp(13000.0), _ip(term["lamb1"]), _IP(iabs), _P(user), _P(altabs), _P(ro),
                        _IP(tables["nexp"]), _P(tables["ki"]), _P(tables["kh"]), *ik, _P(tables["tab_pres"]), _ip(tables["nb_pres"]),
                        _P(tables["tab_temp"]), _ip(tables["nb_temp"]), _P(tables["tab_conc"]), _ip(tables["nb_conc"]), _P(tau),
                        _ip(0), _ip(0), C.byref(ier))
    return ier.value, tau


def read_profile_file(path):
    """PROFIL_TMP as SOS reads it (SOS.F:511-516, format 2X,I5,F10.5,3(E15.8)) -> zprof, h, pcaer, pcmol."""
    z, h, pa, pm = [], [], [], []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            z.append(float(line[7:17])); h.append(float(line[17:32])); pa.append(float(line[32:47])); pm.append(float(line[47:62]))
    return np.array(z), np.array(h), np.array(pa), np.array(pm)


def profile(ref, tmp, altabs, tabs, term):
    """SOS_PROFILE (SOS_PROFIL.F:224) through its file -> (IER, NT, text, zprof, h, pcaer, pcmol)."""
    f = os.path.join(tmp, "PROFIL_REF.txt")
    if os.path.exists(f):
        os.remove(f)
    nt, ier = C.c_int(0), C.c_int(99)
    ref.sos_profile_(_ip(term["iprofil"]), _dp(term["tr"]), _dp(term["hr"]), _dp(term["ta"]), _dp(term["ha"]), _dp(term["zmin"]),
                     _dp(term["zmax"]), _ip(term["absprofil"]), _P(altabs), _P(tabs), _ip(0), _ip(0), _fs(f), C.byref(nt), C.byref(ier), _L)
    if ier.value != 0:
        return ier.value, 0, b"", None, None, None, None
    text = open(f, "rb").read()
    return (0, nt.value, text) + read_profile_file(f)


# ---- the whole run: SOS_PROC (SOS_PROC.F:415-481), the entry binding/run_sos.py calls through f2py ----
_SOS_PROC_DEFAULTS = dict(
    resroot="UNDEFINED_REPERTORY", ficmain_log="NO_LOG_FILE", ficangles_user_lum="NO_USER_ANGLES", ficangles_res_lum="SOS_UsedAngles.txt",
    ficangles_user_mie="NO_USER_ANGLES", ficangles_res_mie="Aer_UsedAngles.txt", ficanglog="NO_LOG_FILE", itronc_aer=1, ficgranu_log="NO_LOG_FILE",
    ficmie_log="NO_LOG_FILE", dir_mie="UNDEFINED_REPERTORY", ficgranu="Aerosols.txt", jd_rmax_mmd_aer=50.0,
    ficextdata_aer="NO_USER_AEROSOLS_PHAZE_FCT", ficmixture_aer="NO_USER_AEROSOLS_MIXTURE", ficuser_aer="NO_USER_AEROSOLS",
    ficprofil_log="NO_LOG_FILE", psurf=1013.0, ficabsprofil="NO_USER_ABS_PROFILE_FILE", dir_surf="UNDEFINED_REPERTORY", ficsurf_log="NO_LOG_FILE",
    ficsurf="DEFAULT", ficsos_log="NO_LOG_FILE", ficsos_res_bin="SOS_Result.bin", fictrans="NO_OUTPUT", ficflux="NO_OUTPUT", zout=-1.0, igmax=100,
    ipolar=1, ier=0, trace=0)                                    # the initial values of SOS_ABS_MAIN.F:1280-1470


def sos_proc(ref, **given):
    """SOS_PROC of the reference library with the arguments of the f2py wrapper by name (radiativetransfer-sos_b200/sos.py: ARGS);
    arguments not given keep the values SOS_ABS_MAIN initialises them with ("not defined" = -999).  Returns (IER, the 23 outputs in
    the wrapper's order: tables as [361, 81] = X_FIN(0:360, 0:80) transposed to (azimuth, angle))."""
    import importlib
    sos = importlib.import_module("radiativetransfer-sos_b200.sos")
    kwm = importlib.import_module("radiativetransfer-sos_b200.keywords")
    args, lens, keep = [], [], []
    ier = C.c_int(0)
    for name, key in sos.ARGS:
        v = given.get(name, _SOS_PROC_DEFAULTS.get(name))
        t = kwm.KEYWORDS[key] if key else "i"
        if name == "ier":
            args.append(C.byref(ier))
        elif t == "s":
            b = _fs(str(v) if v is not None else "UNDEFINED")
            keep.append(b)
            args.append(b)
            lens.append(_L)
        elif t == "i":
            c = C.c_int(int(v) if v is not None else -999)
            keep.append(c)
            args.append(C.byref(c))
        else:
            c = C.c_double(float(v) if v is not None else -999.0)
            keep.append(c)
            args.append(C.byref(c))
    nb = C.c_int(0)
    ind = np.zeros(MX + 1, dtype=np.int32)
    phi, theta = np.zeros(361), np.zeros(MX + 1)
    tabs = [np.zeros((MX + 1, 361)) for _ in range(14)]          # Fortran X_FIN(0:360, 0:80): azimuth fastest
    sc = [C.c_double(0) for _ in range(5)]
    ref.sos_proc_(*args, C.byref(nb), ind.ctypes.data_as(C.POINTER(C.c_int)), _P(phi), _P(theta), *[_P(t) for t in tabs],
                  *[C.byref(s) for s in sc], *lens)
    return ier.value, (nb.value, ind, phi, theta, *[t.T.copy() for t in tabs], *[s.value for s in sc])
