"""aerosols.py (what SOS_AEROSOLS decides on the host) against SOS_AEROSOLS itself (SOS_AEROSOLS.F:680-3127, the driver, in
oracle/_ref/libsosref.so): every aerosol model of the keyword set -- mono-modal log-normal and Junge (0), WMO (1), Shettle & Fenn
(2), bimodal log-normal by volume concentrations and by the coarse share of the optical thickness (3), external phase functions (4),
a user mixture (5) -- from the reference's own angles file to the aerosol result file, compared LINE BY LINE, with KMAT1 (which scales
the optical thickness, SOS_PROC.F:3063), the single scattering albedo and the truncation coefficient BIT-IDENTICAL.

The device work of aerosols.py (Solver.aerosols, Solver.decompo_legendre) is done here by the same device functions stepped on the
host (csrc/aerosol_chain.cuh through tests/aerosol_host.cpp, one thread, no barrier -- TEST INFRASTRUCTURE, see
tests/test_aerosol_chain.py), so that the whole comparison runs without a GPU and libm is the same on both sides."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import aerosol_cases as ac
import refdirect
import test_aerosol_chain as tac
import test_aerosol_models as tam
from test_aerosol_chain import host  # noqa: F401  (fixture)

_fs, _L, _ip, _dp = refdirect._fs, refdirect._L, refdirect._ip, refdirect._dp


class HostSolver:
    """Solver.aerosols / Solver.decompo_legendre with the device functions stepped on the host."""
    def __init__(self, lib):
        self.h = lib
        self.tables = {}

    def aerosols(self, nbmu, xmu, xhr, components, models, os_nb, want_phase=True):
        nc, nm, nang = len(components), len(models), 2 * nbmu + 1
        ck, ph, cier = np.zeros((nc, 3)), np.zeros((nc, 3, nang)), np.zeros(nc, np.int32)
        for i, (rn, in_, a0, af, ig, v1, v2, v3, wa) in enumerate(components):
            key = (rn, in_, a0, af)
            if key not in self.tables:
                self.tables[key] = tac.host_mie(self.h, nbmu, xmu, rn, in_, a0, af)
            e, k1, k2, snr, p11, p12, p33 = tac.host_granu(self.h, nbmu, self.tables[key], ig, v1, v2, v3, wa)
            cier[i], ck[i], ph[i] = e, (k1, k2, snr), (p11, p12, p33)
        scal, coef, mier = np.zeros((nm, 8)), np.zeros((nm, 6, os_nb + 1)), np.zeros(nm, np.int32)
        for m, (ncomp, idx, w, itronc) in enumerate(models):
            e, s, c, _ = tac.host_model(self.h, nbmu, xmu, xhr, ck, ph[:, 0], ph[:, 1], ph[:, 2], ncomp, idx, w, itronc, os_nb)
            mier[m], scal[m], coef[m] = e, s, c
        return dict(comp_k=ck, comp_phase=ph, comp_ier=cier, scal=scal, coef=coef, phase=None, model_ier=mier)

    def decompo_legendre(self, itronc, nbmu, xmu, xhr, os_nb, p11, p12, p22, p33):
        one = np.ones((1, 3))
        e, s, c, ph = tac.host_model(self.h, nbmu, xmu, xhr, one, np.array([p11]), np.array([p12]), np.array([p33]), 0, [0], [1.0], itronc, os_nb,
                                     p22c=np.ascontiguousarray(p22, dtype=np.float64))
        return dict(alp=c[0], beta11=c[1], gamma12=c[2], zeta=c[3], beta22=c[4], delta33=c[5], p11=ph[0], ttt=None, coef_tronca=s[4], z1=s[6],
                    itronc=int(s[7]), ier=e)


def ref_aerosols(ref, tmp, ficangles, wa, waref, aot_ref, itronc, imod, **k):
    """SOS_AEROSOLS of the reference library -> (IER, KMAT1, PIZ, COEF_TRONCA, result file path)."""
    nd = -999.0
    g = lambda n, d=nd: k.get(n, d)
    out = os.path.join(tmp, "Aerosols_ref.txt")
    mie_dir, dir_tmp = os.path.join(tmp, "MIE"), os.path.join(tmp, "TMP")     # the reference looks for '/TMP' in the name (:981)
    os.makedirs(mie_dir, exist_ok=True)
    os.makedirs(dir_tmp, exist_ok=True)
    kmat1, piz, ct, ier = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(99)
    ref.sos_aerosols_(_fs(ficangles), _dp(wa), _dp(aot_ref if wa == waref else 0.1), _dp(waref), _dp(aot_ref), _ip(itronc), _ip(imod),
                      _dp(g("rn")), _dp(g("in_")), _ip(g("igranu", -999)), _dp(g("v1")), _dp(g("v2")), _dp(g("v3")),
                      _ip(g("imodele_wmo", -999)), _dp(g("dl")), _dp(g("ws")), _dp(g("oc")), _dp(g("so")), _ip(g("imodele_sf", -999)), _dp(g("rh")),
                      _ip(g("mode_bilnd", -999)), _dp(g("cv_coarse")), _dp(g("cv_fine")), _dp(g("rtauct")),
                      _dp(g("cm_mrwa")), _dp(g("cm_miwa")), _dp(g("cm_mrwaref")), _dp(g("cm_miwaref")), _dp(g("cm_rmodal")), _dp(g("cm_var")),
                      _dp(g("fm_mrwa")), _dp(g("fm_miwa")), _dp(g("fm_mrwaref")), _dp(g("fm_miwaref")), _dp(g("fm_rmodal")), _dp(g("fm_var")),
                      _fs(g("ficextdata", "NO_USER_AEROSOLS_PHAZE_FCT")), _fs(g("ficmixture", "NO_USER_AEROSOLS_MIXTURE")), _fs(dir_tmp), _fs(mie_dir),
                      _fs("NO_LOG_FILE"), _fs(out), _fs("NO_LOG_FILE"), C.byref(kmat1), C.byref(piz), C.byref(ct), C.byref(ier),
                      _L, _L, _L, _L, _L, _L, _L, _L)
    return ier.value, kmat1.value, piz.value, ct.value, out


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_aerosols_") or not hasattr(ref, "sos_angles_"):
        pytest.skip("oracle/_ref/libsosref.so (with SOS_AEROSOLS, SOS_ANGLES) not available")
    tmp = str(tmp_path_factory.mktemp("aer"))
    root = os.path.join(tmp, "abs_root")
    os.makedirs(os.path.join(root, "fic"))
    ac.write_wmo_file(os.path.join(root, "fic", "Data_WMO_cor_2015_12_16"))
    ac.write_sf_files(os.path.join(root, "fic"))
    os.environ["SOS_ABS_ROOT"] = root
    flum, fmie = os.path.join(tmp, "SOS_UsedAngles.txt"), os.path.join(tmp, "Aer_UsedAngles.txt")
    ier = C.c_int(99)
    ref.sos_angles_(_ip(12), _dp(35.0), _fs("NO_USER_ANGLES"), _ip(20), _fs("NO_USER_ANGLES"), _fs("NO_LOG_FILE"), _fs(flum), _fs(fmie),
                    C.byref(ier), _L, _L, _L, _L, _L)
    assert ier.value == 0
    return ref, tmp, root, fmie


def _mine_file(aer, o, os_nb, path):
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    api.write_aerosols(path, os_nb, o.kmat1, o.kmat2, o.asym, o.coef_tronca, o.piztr, o.alpha, o.beta, o.gamma, o.zeta)
    return open(path).read().split("\n")


CASES = [
    ("mono-modal log-normal", 0, dict(rn=1.45, in_=-0.004, igranu=1, v1=0.10, v2=0.46), 1),
    ("mono-modal Junge", 0, dict(rn=1.40, in_=-0.001, igranu=2, v1=0.05, v2=4.0, v3=5.0), 0),
    ("WMO maritime", 1, dict(imodele_wmo=2), 1),
    ("WMO user volumes", 1, dict(imodele_wmo=4, dl=0.2, ws=0.5, oc=0.25, so=0.05), 1),
    ("Shettle & Fenn maritime, RH 70 %", 2, dict(imodele_sf=3, rh=70.0), 1),
    ("bimodal by volumes", 3, dict(mode_bilnd=1, cv_coarse=0.3, cv_fine=0.7, cm_mrwa=1.45, cm_miwa=-0.004, cm_mrwaref=1.46, cm_miwaref=-0.005,
                                    cm_rmodal=0.6, cm_var=0.5, fm_mrwa=1.40, fm_miwa=-0.008, fm_mrwaref=1.41, fm_miwaref=-0.009, fm_rmodal=0.08,
                                    fm_var=0.45), 1),
    ("bimodal by optical-thickness share", 3, dict(mode_bilnd=2, rtauct=0.4, cm_mrwa=1.45, cm_miwa=-0.004, cm_mrwaref=1.46, cm_miwaref=-0.005,
                                                    cm_rmodal=0.6, cm_var=0.5, fm_mrwa=1.40, fm_miwa=-0.008, fm_mrwaref=1.41, fm_miwaref=-0.009,
                                                    fm_rmodal=0.08, fm_var=0.45), 1),
    ("external phase functions", 4, dict(), 1),
    ("user mixture", 5, dict(), 1),
]


@pytest.mark.parametrize("name,imod,par,itronc", CASES, ids=[c[0] for c in CASES])
def test_aerosol_models_vs_sos_aerosols(setup, host, name, imod, par, itronc):
    ref, tmp, root, fmie = setup
    aer = importlib.import_module("radiativetransfer-sos_b200.aerosols")
    fe = importlib.import_module("radiativetransfer-sos_b200.frontend")
    os.environ["SOS_ABS_ROOT"] = root
    wa, waref, aot, os_nb = 0.865, 0.550, 0.3, 40
    n, xmu, xhr = fe.mie_angles(20)
    if imod == 4:
        wa = waref                                               # SOS_PROC error 2331 otherwise
        par = dict(ficextdata=os.path.join(tmp, "ext.txt"))
        tam._write_ext(par["ficextdata"])
    if imod == 5:
        par = dict(ficmixture=os.path.join(tmp, "mix.txt"))
        open(par["ficmixture"], "w").write(tam.MIX)
    solver = HostSolver(host)
    model = {
        0: lambda: aer.MonoModal(par["rn"], par["in_"], par["igranu"], par["v1"], par["v2"], par.get("v3", -999.0)),
        1: lambda: aer.Wmo(os.path.join(root, "fic", "Data_WMO_cor_2015_12_16"), par["imodele_wmo"], [par.get(k, 0.0) for k in ("dl", "ws", "oc", "so")]),
        2: lambda: aer.ShettleFenn(os.path.join(root, "fic"), par["imodele_sf"], par["rh"]),
        3: lambda: _bimodal(aer, par, waref),
        4: lambda: None,
        5: lambda: aer.read_mixture_file(par["ficmixture"], waref),
    }[imod]()
    for w in ([waref, wa] if wa != waref else [wa]):             # SOS_PROC calls SOS_AEROSOLS at the reference wavelength first
        q = dict(par)
        if imod == 3 and w == waref:                             # SOS_PROC's call at the reference wavelength passes the indices of that
            for k in ("cm_mr", "cm_mi", "fm_mr", "fm_mi"):        # wavelength as the "simulation" ones too (SOS_PROC.F:2896-2907)
                q[k + "wa"] = q[k + "waref"]
        ier, k1, piz, ct, fref = ref_aerosols(ref, tmp, fmie, w, waref, aot, itronc, imod, **q)
        assert ier == 0, (name, w)
        if imod == 4:
            o = aer.external_data(solver, par["ficextdata"], n, xmu, xhr, os_nb, itronc, w, aot)
        else:
            o = aer.run(solver, n, xmu, xhr, os_nb, model, [w], waref=waref, aot_ref=aot, itronc=itronc)[0]
        assert o.kmat1 == k1 and o.piz == piz and o.coef_tronca == ct, (name, w, o.kmat1, k1, o.piz, piz, o.coef_tronca, ct)
        mine, theirs = _mine_file(aer, o, os_nb, os.path.join(tmp, "Aerosols_mine.txt")), open(fref).read().split("\n")
        assert len(mine) == len(theirs) and mine == theirs, (name, w, [(a, b) for a, b in zip(mine, theirs) if a != b][:4])


def _bimodal(aer, par, waref):
    b = aer.BimodalLnd(par["cm_mrwa"], par["cm_miwa"], par["cm_rmodal"], par["cm_var"], par["fm_mrwa"], par["fm_miwa"], par["fm_rmodal"], par["fm_var"])
    if par["mode_bilnd"] == 1:
        b.cv_coarse, b.cv_fine = par["cv_coarse"], par["cv_fine"]
    else:
        b.rtauct = par["rtauct"]
    refi = (par["cm_mrwaref"], par["cm_miwaref"], par["fm_mrwaref"], par["fm_miwaref"])
    sim = (b.coarse_rn, b.coarse_in, b.fine_rn, b.fine_in)
    b.coarse_rn, b.coarse_in, b.fine_rn, b.fine_in = ((lambda wa, r=r, s=s: r if wa == waref else s) for r, s in zip(refi, sim))
    return b
