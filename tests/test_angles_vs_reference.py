"""The angle sets of a run against the reference's own SOS_ANGLES + SOS_ANGLES_GAUSS_USER (SOS_ANGLES.F:227-650, 713-975) in
oracle/_ref/libsosref.so: the radiance angles (Gauss + user angles + the solar angle), the phase-function angles, the expansion
orders and the two angle files it writes -- what synth.sos_angles, frontend.mie_angles, frontend.expansion_orders and
formats.write_angles restate on the host.  Every value is compared through the files' D21.14 fields, i.e. exactly as the
downstream routines of the reference read them (SOS_PREPA_OS, SOS_SURFACE, SOS_AEROSOLS)."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import refdirect

_fs, _L, _ip, _dp = refdirect._fs, refdirect._L, refdirect._ip, refdirect._dp


def _ref_angles(ref, tmp, nb_lum, tetas, user_lum, nb_mie, user_mie):
    flum, fmie = os.path.join(tmp, "SOS_UsedAngles.txt"), os.path.join(tmp, "Aer_UsedAngles.txt")
    for f in (flum, fmie):
        if os.path.exists(f):
            os.remove(f)
    ier = C.c_int(99)
    ref.sos_angles_(_ip(nb_lum), _dp(tetas), _fs(user_lum or "NO_USER_ANGLES"), _ip(nb_mie), _fs(user_mie or "NO_USER_ANGLES"),
                    _fs("NO_LOG_FILE"), _fs(flum), _fs(fmie), C.byref(ier), _L, _L, _L, _L, _L)
    return ier.value, flum, fmie


def _lines(path):
    return [ln.rstrip() for ln in open(path).read().split("\n") if ln.strip()]


@pytest.mark.parametrize("nb_lum,tetas,user,nb_mie,user_mie", [
    (12, 35.0, None, 20, None),
    (12, 35.0, [10.0, 40.0], 20, None),
    (24, 0.0, [0.0, 60.0, 89.5], 40, [5.0, 20.5, 60.0]),           # sun at the zenith, a user angle at the zenith
    (40, 72.5, [72.5], 40, None),                                  # a user angle equal to the solar angle
    (8, 89.0, None, 12, [33.0]),
])
def test_angles_and_angle_files(tmp_path, nb_lum, tetas, user, nb_mie, user_mie):
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_angles_"):
        pytest.skip("oracle/_ref/libsosref.so (with SOS_ANGLES) not available")
    pkg = importlib.import_module("radiativetransfer-sos_b200")
    fe = importlib.import_module("radiativetransfer-sos_b200.frontend")
    syn, fm = pkg.synth, pkg.formats
    tmp = str(tmp_path)
    files = []
    for k, lst in enumerate((user, user_mie)):
        if lst:
            f = os.path.join(tmp, "user%d.txt" % k)
            open(f, "w").write("".join("%g\n" % a for a in lst))
            files.append(f)
        else:
            files.append(None)
    ier, flum, fmie = _ref_angles(ref, tmp, nb_lum, tetas, files[0], nb_mie, files[1])
    assert ier == 0
    os_nb, os_ns, os_nm = fe.expansion_orders(nb_mie, nb_lum)
    # ---- radiance angles ----
    rmu, ga, n0, flags = syn.sos_angles(nb_lum, tetas, fe.read_user_angles(files[0]) if files[0] else [])
    N = (rmu.size - 1) // 2
    mine = os.path.join(tmp, "mine_lum.txt")
    fm.write_angles(mine, rmu[N + 1:], ga[N + 1:], flags, nb_lum, tetas, n0, os_nb, os_ns, os_nm, userfile=files[0] or "NO_USER_ANGLES")
    a, b = _lines(mine), _lines(flum)
    assert a == b, [(x, y) for x, y in zip(a, b) if x != y][:5]
    # ---- phase-function angles (ascending mu, in the file as in the solver's arrays) ----
    n, xmu, xhr = fe.mie_angles(nb_mie, fe.read_user_angles(files[1]) if files[1] else [])
    rows = [ln.split() for ln in _lines(fmie)[5:]]
    assert int(_lines(fmie)[0].split(":")[1]) == n == len(rows)
    fmu = np.array([float(r[1].replace("D", "E")) for r in rows])
    fw = np.array([float(r[2].replace("D", "E")) for r in rows])
    assert np.array_equal(fmu, xmu[n + 1:]) and np.array_equal(fw, xhr[n + 1:])
    assert int(_lines(fmie)[3].split(":")[1]) == os_nb


def test_solar_angle_insertion_sweep(tmp_path):
    """Many solar angles, including the nine tenth-degree values below 90 where TETAS*PI/180 and TETAS*(PI/180) give different
    cosines after the D21.14 field (SOS_ANGLES.F:296, 402 use the second), solar angles on a Gauss node (within
    CTE_SEUIL_ECART_MUS, :401-408: no insertion) and next to one."""
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_angles_"):
        pytest.skip("oracle/_ref/libsosref.so (with SOS_ANGLES) not available")
    pkg = importlib.import_module("radiativetransfer-sos_b200")
    syn = pkg.synth
    tmp = str(tmp_path)
    rng = np.random.default_rng(3)
    mu, _ = syn.sos_gauss(13)
    node = float(np.degrees(np.arccos(mu[4])))
    thetas = [67.7, 76.9, 84.6, 84.9, 86.4, 87.1, 89.0, 89.7, 89.8, 0.0, node, node + 1e-4, node - 1e-4, node + 2e-3] + list(rng.uniform(0.0, 89.9, 40))
    ninsert = 0
    for tetas in thetas:
        ier, flum, _ = _ref_angles(ref, tmp, 12, tetas, None, 20, None)
        assert ier == 0
        ln = _lines(flum)
        rows = [r.split() for r in ln[9:]]
        rmu, ga, n0, flags = syn.sos_angles(12, tetas)
        N = (rmu.size - 1) // 2
        assert int(ln[0].split(":")[1]) == N == len(rows) and int(ln[4].split(":")[1]) == n0, tetas
        assert np.array_equal(np.array([float(r[1].replace("D", "E")) for r in rows]), rmu[N + 1:]), tetas
        assert np.array_equal(np.array([float(r[2].replace("D", "E")) for r in rows]), ga[N + 1:]), tetas
        ninsert += N == 13
    assert 0 < ninsert < len(thetas)                                  # both branches of the insertion were taken
