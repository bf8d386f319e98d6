"""The gas atmosphere of a run from the -AP.* keywords (absprofile.py = SOS_PREPA_ABSPROFILE + DATATM on the host; SURVEY 8f N1 rest).

DATATM (SOS_SUB_TRS.F:908-1003) against the routine itself in oracle/_ref/libsosref.so -- BIT-IDENTICAL -- for a user profile and,
where the reference tree is present, for the six predefined atmospheres, whose tables absprofile.py reads from the DATA statements
of the reference's source file at run time (the translated routines TROPICA .. USTAD62 are the check of that reader).  The rest of
SOS_PREPA_ABSPROFILE (:473-541: about 25 arithmetic statements; the routine itself does not translate) is checked by what it has
to achieve -- column amounts and surface concentrations equal to the requested ones, layer amounts against an independent
vectorised evaluation -- and says so: PARITY UNPINNED for those statements.  CKD tables through the host reader of libsosgpu.so."""
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
