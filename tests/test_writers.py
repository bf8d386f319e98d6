"""SOS_Up.txt / SOS_Down.txt writer of the drop-in surface: byte compatibility with the reference's
SOS_ABS_MAIN.F:2250-2519.  Headers: tests/golden/updown_headers.npz, derived from the reference's own WRITE statements
(SOS_TRPHI.F:1570-1796) by tests/make_golden_headers.py.  Records: formats 55 / 56 (SOS_ABS_MAIN.F:3095-3096) re-implemented
here in Python, independently of the library's C++ formatter.  Host-only: runs without a GPU."""
import importlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "updown_headers.npz"))


def _tables(rng, nphi, N):
    t = rng.standard_normal((7, nphi, N))
    t[0] = rng.uniform(0.0, 180.0, (nphi, N))                # scattering angle
    t[1] = np.abs(t[1]) * 10.0 ** rng.integers(-6, 1, (nphi, N))
    t[2] *= 10.0 ** rng.integers(-9, -1, (nphi, N)); t[3] *= 10.0 ** rng.integers(-9, -1, (nphi, N))
    t[4] = rng.uniform(-90.0, 90.0, (nphi, N)); t[5] = rng.uniform(0.0, 100.0, (nphi, N)); t[6] = np.abs(t[6]) * 1e-3
    t[2, 0, 0] = 0.0; t[4, 0, 0] = -999.0                    # the reference's "undefined" marker and an exact zero
    return t


@pytest.mark.parametrize("tag,phios,zout", [("toa", 0.0, -1.0), ("z3_phi20", 20.0, 3.0)])
def test_updown_view1_bytes(pkg, tmp_path, tag, phios, zout):
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    fm = pkg.formats
    N = 7
    rng = np.random.default_rng(1)
    theta = np.sort(rng.uniform(1.0, 89.0, N))
    up, down = _tables(rng, 2, N), _tables(rng, 2, N)
    fu, fd = tmp_path / "SOS_Up.txt", tmp_path / "SOS_Down.txt"
    api.write_updown(fu, fd, N, 1, phios, 0, zout, None, theta, up, down)
    g = _golden()
    E, F = (lambda x: fm.fortran_e(x, 13, 6)), (lambda x: fm.fortran_f(x, 7, 2))
    for path, tab, ud in ((fu, up, 1), (fd, down, 2)):
        want = g["view1_%s_%d" % (tag, ud)].tobytes().decode()
        for sign, ip, order in ((-1.0, 0, range(N - 1, -1, -1)), (1.0, 1, range(N))):   # FORMAT 55
            for jj in order:
                want += "  %s  %s  %s  %s  %s  %s  %s  %s\n" % (F(sign * theta[jj]), F(tab[0, ip, jj]), E(tab[1, ip, jj]), E(tab[2, ip, jj]),
                                                               E(tab[3, ip, jj]), F(tab[4, ip, jj]), F(tab[5, ip, jj]), E(tab[6, ip, jj]))
        assert open(path, "rb").read() == want.encode()


@pytest.mark.parametrize("tag,zout,pas", [("toa", -1.0, 30), ("z3_phi20", 3.0, 45)])
def test_updown_view2_bytes(pkg, tmp_path, tag, zout, pas):
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    fm = pkg.formats
    N, nphi = 5, 360 // pas + 1
    rng = np.random.default_rng(2)
    theta = np.sort(rng.uniform(1.0, 89.0, N))
    phi = np.arange(0, 361, pas, dtype=np.float64)
    up, down = _tables(rng, nphi, N), _tables(rng, nphi, N)
    fu, fd = tmp_path / "SOS_Up.txt", tmp_path / "SOS_Down.txt"
    api.write_updown(fu, fd, N, 2, 0.0, pas, zout, phi, theta, up, down)
    g = _golden()
    E, F = (lambda x: fm.fortran_e(x, 13, 6)), (lambda x: fm.fortran_f(x, 7, 2))
    for path, tab, ud in ((fu, up, 1), (fd, down, 2)):
        want = g["view2_%s_%d" % (tag, ud)].tobytes().decode()
        for ip in range(nphi):                               # FORMAT 56
            for jj in range(N):
                sca = tab[0, ip, jj]
                if ud == 1:                                  # SCA_UP_FIN(IPHI,JJ): indexed by degrees in the reference (:2467)
                    sca = tab[0, ip * pas, jj] if ip * pas < nphi else 0.0
                want += "  %s  %s  %s %s  %s  %s   %s %s%s\n" % (F(phi[ip]), F(theta[jj]), F(sca), E(tab[1, ip, jj]), E(tab[2, ip, jj]),
                                                               E(tab[3, ip, jj]), F(tab[4, ip, jj]), F(tab[5, ip, jj]), E(tab[6, ip, jj]))
        assert open(path, "rb").read() == want.encode()
    # the documented fix of the quirk
    api.write_updown(fu, fd, N, 2, 0.0, pas, zout, phi, theta, up, down, fix_sca_index=True)
    line = open(fu).read().splitlines()[-1]
    assert line.split()[2] == fm.fortran_f(up[0, nphi - 1, N - 1], 7, 2).strip()


def _F(x, w, d):
    s = "%*.*f" % (w, d, x)
    return s if len(s) <= w else "*" * w


def test_trans_and_flux_files(tmp_path):
    """-SOS.Trans / -SOS.Flux (SOS_PROC.F:3779-3874): the formatted records against an independent rendering of FORMAT
    1005 / 1006 / 1010 / 2010 / 2016-2018 / 2020 (:4944-4951); the list-directed records against gfortran's layout rules
    for a few known values (documented behaviour: `print *, 0.1d0` gives `  0.10000000000000001     `)."""
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    rng = np.random.default_rng(3)
    N = 24
    rmu = np.sort(rng.uniform(0.02, 1.0, N))
    tdg = rng.uniform(0.0, 0.3, N)
    tetas, tt, tv, tds = 32.5, 0.4123, 0.4671, 0.0934
    ft = str(tmp_path / "SOS_Trans.txt")
    api.write_trans(ft, tetas, tt, tv, tds, rmu, tdg)
    pi = np.arccos(-1.0)
    mus = np.cos(pi * tetas / 180.0)
    want = ["Solar Zenith Angle : " + _F(tetas, 7, 3), "Direct transmission TOA -> surface : " + _F(np.exp(-tv / mus), 8, 4), "  ",
            " Diffuse transmittance : TOA -> surface",
            "    thetas = " + _F(tetas, 6, 3) + "   td(thetas) = " + _F(tds + np.exp(-tt / mus) - np.exp(-tv / mus), 7, 4), "  ",
            " Diffuse transmittance : surface -> TOA"]
    for j in range(N):
        want.append("    thetav = " + _F(np.degrees(np.arccos(rmu[j])), 6, 3) + "   td(thetav) = "
                    + _F(tdg[j] + np.exp(-tt / rmu[j]) - np.exp(-tv / rmu[j]), 7, 4))
    assert open(ft).read().split("\n") == want + [""]
    z = np.concatenate([np.arange(0, 25.0), np.arange(25.0, 50.0, 2.5), np.arange(50.0, 121.0, 5.0)])
    tau = np.linspace(0.0, 0.83, 50)
    ff = str(tmp_path / "SOS_Flux.txt")
    for eplus, shown in ((0.1, "  0.10000000000000001     "), (1.5, "   1.5000000000000000     "), (0.0876, "   8.7599999999999997E-002"),
                         (0.25, "  0.25000000000000000     ")):
        tdv, fdd, fd = api.write_flux(ff, tetas, tt, tv, 0.2, eplus, 0.05, 8.0, 0.2, 2.0, z, tau)
        assert (tdv, fdd, fd) == (np.exp(-tv / mus), 0.2 + np.exp(-tt / mus) - np.exp(-tv / mus), 0.2 + np.exp(-tt / mus))
        got = open(ff).read().split("\n")
        assert got[:7] == ["Solar Zenith Angle : " + _F(tetas, 7, 3), "  ", " Downward fluxes at BOA (normalized by TOA solar flux)",
                           "   - Downward direct flux at BOA : " + _F(tdv, 9, 5), "   - Downward diffuse flux at BOA: " + _F(fdd, 9, 5),
                           "   ==> Downward total flux at BOA: " + _F(fd, 9, 5), "  "]
        assert got[7] == " Upward diffuse flux at TOA (normalized by TOA solar flux):" + shown
        assert got[8:12] == ["", "", " According to the following profile", " Z(km)    MOT     AOT     GOT     TOTAL"]
        for n, i in enumerate(range(50, 0, -1)):
            a, b, c = 0.05 * np.exp(-z[i - 1] / 8.0), 0.2 * np.exp(-z[i - 1] / 2.0), tau[50 - i]
            assert got[12 + n] == _F(z[i - 1], 7, 2) + "  " + " ".join(_F(v, 7, 4) for v in (a, b, c, a + b + c))
        assert got[62:] == [""]
    # 'NO_OUTPUT' writes nothing and still returns the three scalars SOS_PROC hands back
    assert api.write_flux("NO_OUTPUT", tetas, tt, tv, 0.2, 0.1, 0.05, 8.0, 0.2, 2.0, z, tau)[0] == np.exp(-tv / mus)
    assert not os.path.exists("NO_OUTPUT")
