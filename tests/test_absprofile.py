"""The gas atmosphere of a run from the -AP.* keywords (absprofile.py = SOS_PREPA_ABSPROFILE + DATATM on the host; SURVEY 8f N1 rest).

DATATM (SOS_SUB_TRS.F:908-1003) against the routine itself in oracle/_ref/libsosref.so -- BIT-IDENTICAL -- for a user profile and,
where the reference tree is present, for the six predefined atmospheres, whose tables absprofile.py reads from the DATA statements
of the reference's source file at run time (the translated routines TROPICA .. USTAD62 are the check of that reader).  absprofile.prepare as a whole
against SOS_PREPA_ABSPROFILE itself -- USERPROFIL, ALTABS, RO, CKD tables, LAMB1 BIT-IDENTICAL, with and without the scalings, for
a user profile and for predefined atmospheres on the reference's own data files; the host CKD reader of libsosgpu.so against
READ_CKD_COEFF itself on generated files and on the reference's own coefficient files; and the scalings once more by what they
have to achieve (column amounts and surface concentrations equal to the requested ones)."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import profile_cases as pc
import refdirect

_P = refdirect._P
REFROOT = "/root/reference"


def _ab():
    return importlib.import_module("radiativetransfer-sos_b200.absprofile")


def _ref_datatm(ref, iatm, user, psurf):
    ro, p, t, alt, dens = np.zeros((8, 50), order="F"), np.zeros(50), np.zeros(50), np.zeros(50), np.zeros(50)
    du = np.array(user, order="F", dtype=np.float64)
    ref.datatm_(_P(ro), _P(p), _P(t), _P(alt), C.byref(C.c_short(iatm)), _P(dens), _P(du), C.byref(C.c_int(50)), C.byref(C.c_double(psurf)))
    return ro, p, t, alt, du


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def _write_profile(path, user):
    with open(path, "w") as f:
        for i in range(50):
            f.write("%2d " % (i + 1) + " ".join("%.9E" % v for v in user[i]) + "\n")


def _user(seed=3):
    user, _, _ = pc.gas_atmosphere(seed)
    user = np.array(user)
    user[:, 3:10] *= 1e-6                                         # ppmv, the unit of the profile file
    user[:, 10] = 2.5e19 * user[:, 1] / 1013.0
    user[:, 11] = 2.3e-5                                          # NO2
    return np.array([[float("%.9E" % v) for v in row] for row in user])


def test_datatm_bit_identical(tmp_path):
    ab = _ab()
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "datatm_"):
        pytest.skip("oracle/_ref/libsosref.so not available")
    user = _user()
    for psurf in (-999.0, 990.0):
        ro, p, t, alt, du = _ref_datatm(ref, 0, user, psurf)
        ro2, p2, t2, alt2, u2 = ab.datatm(user, 0, psurf)
        assert _same(ro, ro2) and _same(p, p2) and _same(t, t2) and _same(alt, alt2) and _same(du, u2)
    if not (os.path.isdir(REFROOT) and hasattr(ref, "tropica_")):
        return
    for iatm in range(1, 7):                                      # tables parsed from the reference's DATA statements
        u = ab.standard_atmosphere(iatm, REFROOT)
        assert u[0, 0] == 0.0 and u[-1, 0] == 120.0 and 1000.0 < u[0, 1] < 1020.0 and (np.diff(u[:, 1]) < 0).all()
        for psurf in (-999.0, 980.0):
            ro, p, t, alt, du = _ref_datatm(ref, iatm, np.zeros((50, 13)), psurf)
            ro2, p2, t2, alt2, u2 = ab.datatm(u, iatm, psurf)
            assert _same(ro, ro2) and _same(p, p2) and _same(t, t2) and _same(alt, alt2) and _same(du[:, :11], u2[:, :11]), (iatm, psurf)
    with pytest.raises(ValueError, match="SOS_SUB_TRS.F"):
        ab.standard_atmosphere(3, str(tmp_path))


def test_atmosphere_scalings(tmp_path):
    ab = _ab()
    user = _user()
    f = str(tmp_path / "profile.txt")
    _write_profile(f, user)
    assert _same(ab.read_user_profile(f), user)
    u0, altabs, ro0 = ab.atmosphere(0, f)
    assert _same(u0, user) and (np.diff(altabs) < 0).all() and altabs[0] == 120.0 and altabs[-1] == 0.0
    # layer amounts: independent evaluation in numpy (float64 throughout, REAL*4 constants as such)
    f32 = lambda x: float(np.float32(x))
    p = user[:, 1]
    mix = {1: 44.0, 3: 44.0, 4: 28.0, 5: 16.0, 6: 32.0, 2: 48.0}
    col = {1: 4, 2: 5, 3: 6, 4: 7, 5: 8, 6: 9}
    for k in range(1, 7):
        lev = user[:, col[k]] * f32(1e-6) * mix[k] / f32(28.97)
        want = (p[:-1] - p[1:]) * (lev[:-1] + lev[1:]) / 2.0 * ab.ATMOCM[k]
        assert np.allclose(ro0[k, :49], want, rtol=1e-14, atol=0) and ro0[k, 49] == lev[49]          # the top level keeps its level value
    assert np.allclose(ro0[7, :49], ((p[:-1] - p[1:]) * (2 * 2.3e-5 * f32(1e-6) * 46 / f32(28.9)) / 2.0 * ab.ATMOCM[7]), rtol=1e-14)
    # requested amounts are met: H2O column (g/cm2), O3 column (Dobson), CO2 / CH4 surface concentration (ppmv)
    u1, _, ro1 = ab.atmosphere(0, f, psurf=1000.0, h2o=2.5, o3=300.0, co2=420.0, ch4=1.9)
    assert abs(ro1[0].sum() / f32(6.022e23) * 18.0 - 2.5) < 1e-12
    assert abs(ro1[2].sum() / f32(6.022e23) * 48.0 * f32(466.23) * 1000.0 - 300.0) < 1e-9
    assert abs(u1[0, 4] / 420.0 - 1) < 1e-7 and abs(u1[0, 8] / 1.9 - 1) < 1e-7          # 44.0E-06 is not 1.0E-06 * 44.0 in REAL*4
    assert _same(u1[:, 1], user[:, 1])                            # a user profile keeps its own pressure column (DATATM :922-926)
    ro_p = ab.atmosphere(0, f, psurf=1000.0)[2]
    assert np.allclose(ro_p[6, :49] / ro0[6, :49], 1000.0 / user[0, 1], rtol=1e-13)
    with pytest.raises(ValueError, match="1021"):
        (tmp_path / "short.txt").write_text("1 0. 1013. 288.\n")
        ab.read_user_profile(str(tmp_path / "short.txt"))
    with pytest.raises(ValueError, match="UserFile"):
        ab.atmosphere(0, None)
    with pytest.raises(ValueError):
        ab.atmosphere(9, None)
    if os.path.isdir(REFROOT):
        u6, alt6, ro6 = ab.atmosphere(6, sos_abs_root=REFROOT, h2o=1.42)
        assert abs(ro6[0].sum() / f32(6.022e23) * 18.0 - 1.42) < 1e-12 and (u6[:, 11] > 0).all() and (u6[:, 12] > 0).all()
        assert 4.4e24 < ro6[6, :49].sum() < 4.6e24                 # O2 column of a standard atmosphere, molecules / cm2


def test_prepare_reads_ckd_tables(tmp_path):
    ab = _ab()
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    root = str(tmp_path)
    t = pc.ckd_tables(4)
    pc.write_ckd_files(root, t)
    f = str(tmp_path / "profile.txt")
    _write_profile(f, _user())
    wl = [1e4 / 13255.0, 1e4 / 13004.0, 1e4 / 13495.0]
    g = ab.prepare(lib, wl, 10.0, 0, f, sos_abs_root=root)
    assert g["lamb1"] == [25, 50, 1]                              # 1 + INT((NUMAX - NU) / NUSTEP), NUMAX = 13500
    assert g["tables"][0] is g["tables"][1] is g["tables"][2]     # one coefficient file, read once
    for k in ("nexp", "ai", "ki", "kh", "tab_temp", "tab_pres", "tab_conc"):
        assert np.array_equal(g["tables"][0][k], t[k]), k
    assert g["kdis_ai"][0] is g["tables"][0]["ai"] and g["userprofil"].shape == (50, 13) and g["ro"].shape == (8, 50)
    # a list that spans two coefficient files: each wavelength gets the tables of its own file
    t2 = pc.ckd_tables(5)
    pc.write_ckd_files(root, t2, numax=13000)
    g2 = ab.prepare(lib, [1e4 / 13255.0, 1e4 / 12800.0, 1e4 / 13100.0], 10.0, 0, f, sos_abs_root=root)
    assert g2["lamb1"] == [25, 21, 41] and g2["tables"][0] is g2["tables"][2] and g2["tables"][1] is not g2["tables"][0]
    assert np.array_equal(g2["tables"][1]["ki"], t2["ki"]) and np.array_equal(g2["tables"][0]["ki"], t["ki"])
    with pytest.raises(ValueError, match="READ_CKD_COEFF"):       # no file for 13500 .. 14000 cm-1
        ab.prepare(lib, [1e4 / 13500.0], 10.0, 0, f, sos_abs_root=root)
    with pytest.raises(ValueError, match="905"):
        ab.prepare(lib, [5.0], 10.0, 0, f, sos_abs_root=root)
    with pytest.raises(ValueError, match="READ_CKD_COEFF"):
        ab.prepare(lib, [1e4 / 13255.0], 1.0, 0, f, sos_abs_root=root)


def _ref_read_ckd(ref, root, k, jabs, nu, nustep):
    os.environ["SOS_ABS_ROOT"] = root
    nexp = np.zeros((pc.NBABS, pc.NWVL), dtype=np.int32, order="F")
    ai = np.zeros((pc.NAI, pc.NBABS, pc.NWVL), order="F")
    ki = np.zeros((pc.NTMAX, pc.NPMAX, pc.NAI, pc.NBABS, pc.NWVL), order="F")
    kh = np.zeros((pc.NTMAX, pc.NPMAX, pc.NCMAX, pc.NAI, pc.NWVL), order="F")
    tp, tt, tc = np.zeros(pc.NPMAX), np.zeros(pc.NTMAX), np.zeros(pc.NCMAX)
    numax, numin, nbp, nbt, nbc, ier = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(99)
    I = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    ref.read_ckd_coeff_(C.byref(C.c_int(k)), C.byref(C.c_int(jabs)), C.byref(C.c_double(nu)), C.byref(C.c_double(nustep)), I(nexp), _P(ai), _P(ki),
                        _P(kh), C.byref(numax), C.byref(numin), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc), C.byref(nbc), C.byref(ier))
    return ier.value, dict(nexp=nexp, ai=ai, ki=ki, kh=kh, tab_pres=tp, tab_temp=tt, tab_conc=tc, nb=(nbp.value, nbt.value, nbc.value),
                           rng=(numax.value, numin.value))


def _mine_read_ckd(lib, root, k, jabs, nu, nustep):
    nexp = np.zeros((pc.NBABS, pc.NWVL), dtype=np.int32, order="F")
    ai = np.zeros((pc.NAI, pc.NBABS, pc.NWVL), order="F")
    ki = np.zeros((pc.NTMAX, pc.NPMAX, pc.NAI, pc.NBABS, pc.NWVL), order="F")
    kh = np.zeros((pc.NTMAX, pc.NPMAX, pc.NCMAX, pc.NAI, pc.NWVL), order="F")
    tp, tt, tc = np.zeros(pc.NPMAX), np.zeros(pc.NTMAX), np.zeros(pc.NCMAX)
    numax, numin, nbp, nbt, nbc = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0), C.c_int(0)
    lib.sosgpu_read_ckd_coeff.restype = C.c_int
    rc = lib.sosgpu_read_ckd_coeff(root.encode(), C.c_int(k), C.c_int(jabs), C.c_double(nu), C.c_double(nustep), nexp.ctypes.data_as(C.POINTER(C.c_int)),
                                   _P(ai), _P(ki), _P(kh), C.byref(numax), C.byref(numin), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc),
                                   C.byref(nbc))
    return rc, dict(nexp=nexp, ai=ai, ki=ki, kh=kh, tab_pres=tp, tab_temp=tt, tab_conc=tc, nb=(nbp.value, nbt.value, nbc.value),
                    rng=(numax.value, numin.value))


def test_ckd_reader_vs_the_references_reader(tmp_path):
    """sosgpu_read_ckd_coeff (host code of libsosgpu.so) against READ_CKD_COEFF itself (SOS_SUB_TRS.F:481-905, in libsosref.so): every
    array identical -- on generated files for the eight gases (H2O with its concentration dimension) and, where the reference tree is
    present, on the reference's own coefficient files (O2 A band, CO2, CH4, O3 ... at 10, 5 and 1 cm-1)."""
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "read_ckd_coeff_"):
        pytest.skip("oracle/_ref/libsosref.so (with READ_CKD_COEFF) not available")
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    root = str(tmp_path)
    pc.write_ckd_files(root, pc.ckd_tables(4))
    cases = [(root, k, 1, 13255.0, 10.0) for k in range(1, 9)] + [(root, 7, 0, 13255.0, 10.0), (root, 1, 0, 13255.0, 10.0)]
    if os.path.isdir(os.path.join(REFROOT, "fic", "COEFF_CKD")):
        for step, nus in ((10, (13100.0, 4800.0, 20950.0)), (5, (13102.5,)), (1, (13120.0,))):
            d = os.path.join(REFROOT, "fic", "COEFF_CKD", "%dcmm1" % step)
            for k, g in enumerate(pc.GAS):
                for nu in nus:
                    lo = (int(27500 - nu) // (50 * step) + 1) * 50 * step
                    if os.path.exists(os.path.join(d, "coef_%s_%d_%d_%dcmm1" % (g, 27500 - lo + 50 * step, 27500 - lo, step))):
                        cases.append((REFROOT, k + 1, 1, nu, float(step)))
    nreal = 0
    for rt, k, jabs, nu, step in cases:
        ier, a = _ref_read_ckd(ref, rt, k, jabs, nu, step)
        rc, b = _mine_read_ckd(lib, rt, k, jabs, nu, step)
        assert ier == rc == 0, (rt, k, nu, step, ier, rc)
        nreal += rt == REFROOT
        for key in ("nexp", "ai", "ki", "kh", "tab_pres", "tab_temp", "tab_conc"):
            if key in ("tab_conc",) and k != 1:
                continue
            assert np.array_equal(a[key], b[key]), (rt, k, nu, step, key)
        if jabs:
            assert a["rng"] == b["rng"] and a["nb"][:2] == b["nb"][:2], (rt, k, nu)
    print("\n[CKD reader] %d cases identical to READ_CKD_COEFF, %d of them on the reference's own coefficient files" % (len(cases), nreal))
    for bad in ((root, 7, 1, 13255.0, 2.0), (root, 7, 1, 5000.0, 10.0)):       # unsupported resolution, missing file: IER = -1 on both sides
        assert _ref_read_ckd(ref, *bad)[0] == -1 and _mine_read_ckd(lib, *bad)[0] == -1


def test_prepare_vs_sos_prepa_absprofile(tmp_path):
    """absprofile.prepare against SOS_PREPA_ABSPROFILE itself (SOS_PREPA_ABSPROFILE.F:248-752, in libsosref.so; it calls DATATM and
    READ_CKD_COEFF of the same library): USERPROFIL, ALTABS, RO, the CKD tables, LAMB1 -- BIT-IDENTICAL -- for a user profile with
    and without the scalings of -AP.Psurf / H2O / O3 / CO2 / CH4 and, where the reference tree is present, for the predefined
    atmospheres with the reference's own SO2-NO2 and coefficient files."""
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_prepa_absprofile_"):
        pytest.skip("oracle/_ref/libsosref.so (with SOS_PREPA_ABSPROFILE) not available")
    ab = _ab()
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    root = str(tmp_path)
    pc.write_ckd_files(root, pc.ckd_tables(4))
    f = str(tmp_path / "profile.txt")
    _write_profile(f, _user())
    I = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))

    def reference(rt, wa, step, absprofil, fic, psurf, h2o, o3, co2, ch4):
        os.environ["SOS_ABS_ROOT"] = rt
        nexp = np.zeros((pc.NBABS, pc.NWVL), dtype=np.int32, order="F")
        ai = np.zeros((pc.NAI, pc.NBABS, pc.NWVL), order="F")
        ki = np.zeros((pc.NTMAX, pc.NPMAX, pc.NAI, pc.NBABS, pc.NWVL), order="F")
        kh = np.zeros((pc.NTMAX, pc.NPMAX, pc.NCMAX, pc.NAI, pc.NWVL), order="F")
        tp, tt, tc = np.zeros(pc.NPMAX), np.zeros(pc.NTMAX), np.zeros(pc.NCMAX)
        user, altabs, ro = np.zeros((50, 13), order="F"), np.zeros(50), np.zeros((8, 50), order="F")
        iabs = np.zeros(8, dtype=np.int32)
        nu, lamb1, nbp, nbt, nbc, ier = C.c_double(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(99)
        d = lambda v: C.byref(C.c_double(v))
        ref.sos_prepa_absprofile_(d(wa), d(step), d(psurf), d(h2o), d(o3), d(co2), d(ch4), C.byref(C.c_int(absprofil)), refdirect._fs(fic or "NONE"),
                                  C.byref(C.c_int(0)), C.byref(C.c_int(0)), C.byref(nu), C.byref(lamb1), I(iabs), _P(user), _P(altabs), _P(ro),
                                  I(nexp), _P(ai), _P(ki), _P(kh), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc), C.byref(nbc), C.byref(ier),
                                  refdirect._L)
        return ier.value, dict(user=user, altabs=altabs, ro=ro, lamb1=lamb1.value, nexp=nexp, ai=ai, ki=ki, kh=kh, tab_pres=tp, tab_temp=tt, tab_conc=tc)

    nd = -999.0
    cases = [(root, 1e4 / 13255.0, 10.0, 0, f, nd, nd, nd, nd, nd), (root, 1e4 / 13004.0, 10.0, 0, f, 1000.0, 2.5, 300.0, 420.0, 1.9),
             (root, 1e4 / 13495.0, 10.0, 0, f, nd, 0.8, nd, nd, 1.7)]
    if os.path.isdir(os.path.join(REFROOT, "fic", "COEFF_CKD")) and all(
            os.path.exists(os.path.join(REFROOT, "fic", "COEFF_CKD", "10cmm1", "coef_%s_13500_13000_10cmm1" % g)) for g in pc.GAS):
        cases += [(REFROOT, 0.7625, 10.0, 6, None, nd, nd, nd, nd, nd), (REFROOT, 0.7625, 10.0, 1, None, 1005.0, 4.1, 250.0, nd, nd),
                  (REFROOT, 0.765, 10.0, 3, None, nd, nd, nd, 400.0, nd)]
    for rt, wa, step, absprofil, fic, psurf, h2o, o3, co2, ch4 in cases:
        ier, r = reference(rt, wa, step, absprofil, fic, psurf, h2o, o3, co2, ch4)
        assert ier == 0, (rt, wa, absprofil)
        g = ab.prepare(lib, [wa], step, absprofil, fic, psurf, h2o, o3, co2, ch4, sos_abs_root=rt)
        assert g["lamb1"] == [r["lamb1"]]
        assert _same(g["userprofil"], r["user"]) and _same(g["altabs"], r["altabs"]) and _same(g["ro"], r["ro"]), (rt, wa, absprofil)
        for key in ("nexp", "ai", "ki", "kh", "tab_pres", "tab_temp", "tab_conc"):
            assert np.array_equal(g["tables"][0][key], r[key]), key
