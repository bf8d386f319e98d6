"""-AER.Model 4 (external phase functions, -AER.ExtData) and 5 (user mixture, -AER.DefMixture): host side (SURVEY 8f N3 rest).
The spline interpolation of SOS_AEROSOLS.F:4822-5105 restated in aerosols.py against SOS_INTERPO_SPLINT of
oracle/_ref/libsosref.so -- BIT-IDENTICAL; the readers of the two user files; the weights of a user mixture
(SOS_AEROSOLS.F:2383-2594) and the plan handed to the device.  The device work itself (SOS_DECOMPO_LEGENDRE, the Mie chain, the
mixture) is what tests/test_aerosol_chain.py checks; here a stand-in records what it is asked."""
import ctypes as C
import importlib

import numpy as np
import pytest

import refdirect

_P, _ip = refdirect._P, refdirect._ip


def _aer():
    return importlib.import_module("radiativetransfer-sos_b200.aerosols")


def test_spline_interpolation_bit_identical():
    aer = _aer()
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_interpo_splint_"):
        pytest.skip("oracle/_ref/libsosref.so not available")
    rng = np.random.default_rng(11)
    for n, nout in ((5, 7), (37, 41), (181, 83), (200, 201)):
        ang = np.sort(rng.uniform(0.0, 180.0, n))
        ang[0], ang[-1] = 0.0, 180.0
        xin = np.cos(np.radians(ang))                                  # descending: the routine sorts
        yin = np.exp(3.0 * xin) * (1.0 + 0.1 * rng.standard_normal(n))
        xout = np.sort(rng.uniform(-1.0, 1.0, nout))
        yout = np.zeros(nout)
        ier = C.c_int(99)
        ref.sos_interpo_splint_(_ip(n), _P(xin), _P(yin), _ip(nout), _P(xout), _P(yout), C.byref(ier))
        assert ier.value == 0
        mine = aer.interpo_splint(xin, yin, xout)
        assert np.array_equal(mine.view(np.uint64), yout.view(np.uint64)), (n, np.abs(mine - yout).max())
    with pytest.raises(ValueError):
        aer.interpo_splint([0.0, 0.5, 0.5, 1.0], [1.0, 2.0, 3.0, 4.0], [0.5])


def _write_ext(path, n=37):
    ang = np.linspace(0.0, 180.0, n)
    mu = np.cos(np.radians(ang))
    f11 = (1 - 0.7 ** 2) / (1 + 0.7 ** 2 - 2 * 0.7 * mu) ** 1.5
    with open(path, "w") as f:
        f.write("EXTINCTION CROSS SECTION (mic^2) : 2.5000\nSCATTERING CROSS SECTION (mic^2) : 2.2500\nNUMBER OF ANGLES : %d\n" % n)
        f.write("  ANGLE     F11         -F12/F11      F22/F11      F33/F11\n")
        for a, m, p in zip(ang, mu, f11):
            f.write("%8.3f %14.7E %14.7E %14.7E %14.7E\n" % (a, p, 0.3 * (1 - m * m) / (1 + m * m), 1.0, 2 * m / (1 + m * m)))
    return ang, f11


class _Dev:
    def __init__(self):
        self.asked = None

    def decompo_legendre(self, itronc, nbmu, xmu, xhr, os_nb, p11, p12, p22, p33):
        self.asked = dict(itronc=itronc, nbmu=nbmu, p11=np.array(p11), p12=np.array(p12), p22=np.array(p22), p33=np.array(p33))
        z = np.zeros(os_nb + 1)
        b = z.copy()
        b[:2] = 1.0, 2.1
        return dict(alp=z, beta11=b, beta22=b, gamma12=z, delta33=z, zeta=z, p11=p11, ttt=p11, coef_tronca=0.4, z1=1.0, itronc=1, ier=0)

    def aerosols(self, nbmu, xmu, xhr, components, models, os_nb, want_phase=True):
        self.components, self.models = list(components), list(models)
        nc, nm = len(components), len(models)
        ck = np.array([[1.0 + i + 10 * c[0], 0.9 * (1.0 + i), 1.0] for i, c in enumerate(components)])
        return dict(comp_k=ck, comp_ier=np.zeros(nc, np.int32), model_ier=np.zeros(nm, np.int32),
                    scal=np.tile([2.0, 1.8, 0.9, 0.88, 0.3, 0.7, 1.0, 1.0], (nm, 1)), coef=np.zeros((nm, 6, os_nb + 1)))


def test_external_data_model(tmp_path):
    aer = _aer()
    fe = importlib.import_module("radiativetransfer-sos_b200.frontend")
    path = str(tmp_path / "ext.txt")
    ang, f11 = _write_ext(path)
    k1, k2, mu, g11, g12, g22, g33 = aer.read_external_data(path)
    assert (k1, k2) == (2.5, 2.25) and mu.size == 37 and mu[0] == 1.0 and abs(mu[-1] + 1.0) < 1e-15
    assert np.allclose(g11, f11, rtol=1e-7) and (g12 <= 0).all() and np.array_equal(g22, g11)        # F12 = -(-F12/F11) * F11
    n, xmu, xhr = fe.mie_angles(20)
    dev = _Dev()
    o = aer.external_data(dev, path, n, xmu, xhr, 40, 1, 0.55, 0.3)
    assert dev.asked["nbmu"] == n and dev.asked["p11"].shape == (2 * n + 1,)
    want = (1 - 0.49) / (1 + 0.49 - 1.4 * xmu) ** 1.5
    assert np.abs(dev.asked["p11"] / want - 1).max() < 2e-3                       # the spline through 37 nodes of a HG function
    assert o.ta == 0.3 and o.kmat1 == 2.5 and o.piz == 0.9 and o.coef_tronca == 0.4
    assert o.piztr == 0.9 * (1. - 0.4 / 2.) / (1. - 0.9 * 0.4 / 2.) and o.asym == 0.4 / 2. + (1. - 0.4 / 2.) * 2.1 / 3.
    (tmp_path / "bad.txt").write_text("A : 1.\nB : 1.\nN : 300\n\n")
    with pytest.raises(ValueError, match="950"):
        aer.read_external_data(str(tmp_path / "bad.txt"))
    (tmp_path / "short.txt").write_text("A : 1.\nB : 1.\nN : 3\nheader\n0. 1. 0. 1. 1.\n")
    with pytest.raises(ValueError, match="942"):
        aer.read_external_data(str(tmp_path / "short.txt"))


MIX = """Number of modes : 2
Mode 1 size distribution : LND
  modal radius (microns) : 0.10
  standard deviation : 0.46
  refractive index at WA, real part : 1.40
  refractive index at WA, imaginary part : -0.001
  refractive index at WAREF, real part : 1.41
  refractive index at WAREF, imaginary part : -0.002
  AOT rate at WAREF : 0.7
Mode 2 size distribution : JUNGE
  slope : 4.0
  minimal radius (microns) : 0.05
  maximal radius (microns) : 5.0
  refractive index at WA, real part : 1.50
  refractive index at WA, imaginary part : -0.01
  refractive index at WAREF, real part : 1.51
  refractive index at WAREF, imaginary part : -0.02
  AOT rate at WAREF : 0.3
"""


def test_user_mixture_model(tmp_path):
    aer = _aer()
    path = tmp_path / "mix.txt"
    path.write_text(MIX)
    m = aer.read_mixture_file(str(path), 0.55)
    assert len(m.modes) == 2 and m.modes[0] == (1, 0.10, 0.46, 0.0, 1.40, -0.001, 1.41, -0.002, 0.7)
    assert m.modes[1] == (2, 0.05, 4.0, 5.0, 1.50, -0.01, 1.51, -0.02, 0.3)                       # SOS_GRANU's (rmin, slope, rmax)
    dev = _Dev()
    out = aer.run(dev, 20, np.zeros(41), np.zeros(41), 40, m, [0.865], waref=0.55, aot_ref=0.2, itronc=1)
    # the last device call: both wavelengths (0.865 and the reference one), two components each, same weights
    assert len(dev.components) == 4 and len(dev.models) == 2
    c = dev.components
    assert c[0][:2] == (1.40, -0.001) and c[2][:2] == (1.41, -0.002) and c[1][4:8] == (2, 0.05, 4.0, 5.0) and c[3][:2] == (1.51, -0.02)
    af_lnd = aer.alphaf_of(aer.lnd_rmax(0.10, 0.46), aer.WAMIN)
    assert c[0][3] == af_lnd == c[2][3] and c[1][3] == aer.alphaf_of(5.0, aer.WAMIN) and c[0][8] == 0.865 and c[2][8] == 0.55
    # weights (SOS_AEROSOLS.F:2583-2594) from the stand-in's cross sections at the reference wavelength: K1 = 1 + i + 10 rn
    k = [1.0 + 0 + 10 * 1.41, 1.0 + 1 + 10 * 1.51]
    ca = [0.2 * 0.7 / k[0], 0.2 * 0.3 / k[1]]
    w = [ca[0] / (ca[0] + ca[1]), ca[1] / (ca[0] + ca[1])]
    assert dev.models[0][0] == 2 and dev.models[0][1] == [0, 1] and dev.models[1][1] == [2, 3]
    assert dev.models[0][2] == w and dev.models[1][2] == w
    assert out[0].ta == (2.0 / 2.0) * 0.2                                          # K(WA) / K(WAREF) * AOT_REF with the stand-in's scalars
    path.write_text(MIX.replace("AOT rate at WAREF : 0.3", "AOT rate at WAREF : 0.31"))
    with pytest.raises(ValueError, match="963"):
        aer.read_mixture_file(str(path), 0.55)
    path.write_text(MIX.replace("JUNGE", "GAMMA"))
    with pytest.raises(ValueError, match="962"):
        aer.read_mixture_file(str(path), 0.55)
    path.write_text("Number of modes : 5\n" + "".join(MIX.split("\n", 1)[1].split("Mode 2")[0].replace("0.7", "0.2") for _ in range(5)))
    five = aer.read_mixture_file(str(path), 0.55)
    with pytest.raises(NotImplementedError):
        aer.run(dev, 20, np.zeros(41), np.zeros(41), 40, five, [0.865], waref=0.55, aot_ref=0.2)
