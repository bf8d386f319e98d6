"""Inputs and reference-side calls of the aerosol chain (SOS_MIE -> SOS_GRANU -> mixture -> SOS_DECOMPO_LEGENDRE).
TEST INFRASTRUCTURE.  Inputs are generated here (Gauss angles from numpy, refractive indices and size distributions from
literals), so the GPU box builds identical inputs without /root/reference; the reference routines are called in
oracle/_ref/libsosref.so (translated from the reference's own Fortran) through the fixed strides of inc/SOS.h."""
import ctypes as C
import os
import struct

import numpy as np

MM = 100          # CTE_MIE_NBMU_MAX
NBM = 200         # CTE_OS_NB_MAX
_ip = lambda v: C.byref(C.c_int(int(v)))
_dp = lambda v: C.byref(C.c_double(float(v)))
_P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
_L = C.c_size_t(500)


def _fs(s):
    return C.create_string_buffer(s.encode().ljust(500), 500)


def mie_angles(nb_gauss, user_deg=()):
    """Angles of the Mie / phase-function calculations as SOS_ANGLES lays them out: the nb_gauss positive Gauss abscissae of a
    2*nb_gauss-point rule on (-1, 1) plus user angles of weight 0, merged in ascending mu; returns (nbmu, xmu[2n+1], xhr[2n+1])
    with V(-n:n) at [j+n], mu(0) = 0 with weight 0."""
    x, w = np.polynomial.legendre.leggauss(2 * nb_gauss)
    mu = list(x[nb_gauss:])
    wt = list(w[nb_gauss:])
    for a in user_deg:
        mu.append(float(np.cos(np.radians(a))))
        wt.append(0.0)
    order = np.argsort(mu)
    mu, wt = np.array(mu)[order], np.array(wt)[order]
    n = mu.size
    xmu, xhr = np.zeros(2 * n + 1), np.zeros(2 * n + 1)
    xmu[n + 1:], xhr[n + 1:] = mu, wt
    xmu[:n], xhr[:n] = -mu[::-1], wt[::-1]
    return n, xmu, xhr


def _fixed(v, n):
    out = np.zeros(2 * MM + 1)
    out[MM - n:MM + n + 1] = v
    return out


def read_mie_file(path, nbmu):
    """The unformatted sequential file SOS_MIE writes (gfortran records): header RN, IN, ALPHAF, MIE_NBMU; one record per size
    parameter: ALPHA, QEXT, QSCA (REAL*4), G (REAL*8), IMIE, QMIE, UMIE (-nbmu:nbmu) (REAL*4)."""
    raw = open(path, "rb").read()
    nang = 2 * nbmu + 1
    (l0,) = struct.unpack_from("i", raw, 0)
    assert l0 == 28
    rn, in_, alphaf, nb = struct.unpack_from("=dddi", raw, 4)
    assert nb == nbmu
    pos, reclen = 4 + l0 + 4, 12 + 8 + 12 * nang
    nrec = (len(raw) - pos) // (reclen + 8)
    assert pos + nrec * (reclen + 8) == len(raw)
    dt = np.dtype([("l0", "<i4"), ("rec", "<f4", 3), ("g", "<f8"), ("i", "<f4", nang), ("q", "<f4", nang), ("u", "<f4", nang), ("l1", "<i4")])
    a = np.frombuffer(raw, dtype=dt, count=nrec, offset=pos)
    assert (a["l0"] == reclen).all() and (a["l1"] == reclen).all()
    return dict(rn=rn, in_=in_, alphaf=alphaf, rec=np.ascontiguousarray(a["rec"]), g=np.ascontiguousarray(a["g"]),
                imie=np.ascontiguousarray(a["i"]), qmie=np.ascontiguousarray(a["q"]), umie=np.ascontiguousarray(a["u"]))


def ref_mie(ref, tmp, nbmu, xmu, xhr, rn, in_, alpha0, alphaf, name="MIE.bin"):
    """SOS_MIE (SOS_MIE.F:205) through its file."""
    f = os.path.join(tmp, name)
    if os.path.exists(f):
        os.remove(f)
    ier = C.c_int(99)
    ref.sos_mie_(_ip(nbmu), _P(_fixed(xmu, nbmu)), _P(_fixed(xhr, nbmu)), _dp(rn), _dp(in_), _dp(alpha0), _dp(alphaf), _fs(f),
                 _fs("NO_LOG_FILE"), C.byref(ier), _L, _L)
    assert ier.value == 0, "reference SOS_MIE IER=%d" % ier.value
    return f, read_mie_file(f, nbmu)


def ref_granu(ref, ficmie, igranu, v1, v2, v3, wa, nbmu, xmu):
    """SOS_GRANU (SOS_AEROSOLS.F:4392) on a Mie file -> (ier, kmat1, kmat2, somme_nr, p11, p12, p33)."""
    p11, p12, p33 = (np.zeros(2 * MM + 1) for _ in range(3))
    k1, k2, snr, ier = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(99)
    ref.sos_granu_(_fs(ficmie), _ip(igranu), _dp(v1), _dp(v2), _dp(v3), _dp(wa), _ip(nbmu), _P(_fixed(xmu, nbmu)), _ip(0),
                   C.byref(k1), C.byref(k2), C.byref(snr), _P(p11), _P(p12), _P(p33), C.byref(ier), _L)
    cut = slice(MM - nbmu, MM + nbmu + 1)
    return ier.value, k1.value, k2.value, snr.value, p11[cut].copy(), p12[cut].copy(), p33[cut].copy()


def ref_decompo(ref, itronc, nbmu, xmu, xhr, os_nb, p11, p12, p22, p33):
    """SOS_DECOMPO_LEGENDRE (SOS_AEROSOLS.F:3924) -> dict with the (possibly truncated) P11, TTT, COEF_TRONCA, Z1, ITRONC on exit
    and the six coefficient arrays (0:os_nb)."""
    a = {k: _fixed(v, nbmu) for k, v in dict(p11=p11, p12=p12, p22=p22, p33=p33).items()}
    ttt = np.zeros(2 * MM + 1)
    co = {k: np.zeros(NBM + 1) for k in ("alp", "beta11", "beta22", "gamma12", "delta33", "zeta")}
    it, ct, z1, ier = C.c_int(itronc), C.c_double(0), C.c_double(0), C.c_int(0)
    ref.sos_decompo_legendre_(C.byref(it), _ip(0), _ip(nbmu), _P(_fixed(xmu, nbmu)), _P(_fixed(xhr, nbmu)), _ip(os_nb), _P(a["p11"]), _P(ttt),
                              _P(a["p12"]), _P(a["p22"]), _P(a["p33"]), C.byref(ct), C.byref(z1), _P(co["alp"]), _P(co["beta11"]),
                              _P(co["beta22"]), _P(co["gamma12"]), _P(co["delta33"]), _P(co["zeta"]), C.byref(ier))
    cut = slice(MM - nbmu, MM + nbmu + 1)
    out = {k: v[:os_nb + 1].copy() for k, v in co.items()}
    out.update(p11=a["p11"][cut].copy(), ttt=ttt[cut].copy(), coef_tronca=ct.value, z1=z1.value, itronc=it.value, ier=ier.value)
    return out


def lnd_rmax(rm, sig):
    """Upper radius of a log-normal mode: where n(r) falls to CTE_COEF_NRMAX = 0.0001 (REAL*4 literal) of its maximum
    (SOS_AEROSOLS.F:1175-1177)."""
    return rm * np.exp(sig * sig) * np.exp(sig * np.sqrt(-2.0 * np.log(np.float64(np.float32(0.0001)))))


def alphaf_for(rmax, wa):
    """ALPHAF = REAL(100 + 100*DINT(2 pi RMAX / (100 WA))) (SOS_AEROSOLS.F:1944-1945)."""
    return float(np.float32(100 + 100 * np.trunc(np.float64(np.float32(2.0)) * np.pi * rmax / (np.float64(np.float32(100.0)) * wa))))


# components: (rn, in, igranu, v1, v2, v3) -- a coarse and a fine log-normal mode, a Junge law
COARSE = (1.45, -0.004, 1, 0.40, 0.60, -999.0)
FINE = (1.42, -0.008, 1, 0.08, 0.45, -999.0)
JUNGE = (1.40, -0.002, 2, 0.05, 4.0, 3.0)


def write_wmo_file(path, seed=3):
    """A data file in the layout SOS_INIT_PARAMWMO reads (formats 333 / 444 / 555, SOS_AEROSOLS.F:3545-3547): modal radii,
    log10 sigmas, particle volumes, then wavelength + four complex indices per line.  Values generated here (not the reference's
    table), so the GPU box can build it."""
    rng = np.random.default_rng(seed)
    rm = [0.45, 0.006, 0.28, 0.012]
    s10 = [0.46, 0.47, 0.40, 0.30]
    vol = [95.0, 1.1e-4, 5.2, 6.0e-5]
    with open(path, "w") as f:
        f.write("".join(" %9.5f" % x for x in rm) + "\n")
        f.write("".join(" %9.5f" % x for x in s10) + "\n")
        f.write(" %9.5f %12.5E %9.5f %12.5E\n" % tuple(vol))
        for wa in (0.3, 0.4, 0.55, 0.7, 0.9, 1.1, 1.6, 2.2, 3.0):
            row = [wa]
            for i in range(4):
                row += [1.35 + 0.1 * i + 0.02 * rng.random(), -abs(rng.normal()) * 10.0 ** (-1 - (i % 3))]
            f.write("".join(" %9.5f" % x for x in row) + "\n")
    return path


def ref_wmo_params(ref, path, wa):
    """SOS_INIT_PARAMWMO (SOS_AEROSOLS.F:3334) of the reference library."""
    v1, v2, mr, mi = (np.zeros(5) for _ in range(4))
    vol = np.zeros(4)
    ier = C.c_int(99)
    ref.sos_init_paramwmo_(_fs(path), _dp(wa), _P(v1), _P(v2), _P(mr), _P(mi), _P(vol), C.byref(ier), _L)
    return ier.value, v1[:4].copy(), v2[:4].copy(), mr[:4].copy(), mi[:4].copy(), vol.copy()


def write_sf_files(dirfic, seed=5):
    """Data files in the layout SOS_INIT_PARAMSF reads (formats 222 / 333 / 555, SOS_AEROSOLS.F:3838-3840) under the reference's
    file names: log10 sigmas, modal radii per relative humidity, and per component a table wavelength x 8 humidities of complex
    indices.  Values generated here."""
    rng = np.random.default_rng(seed)
    names = ("Data_SF_cor_2015_12_16", "IRefrac_SR_cor_2015_12_16", "IRefrac_LR", "IRefrac_SU_cor_2015_12_16", "IRefrac_LU_cor_2015_12_16",
             "IRefrac_OM_cor_2015_12_16")
    os.makedirs(dirfic, exist_ok=True)
    rhs = (0.0, 50.0, 70.0, 80.0, 90.0, 95.0, 98.0, 99.0)
    with open(os.path.join(dirfic, names[0]), "w") as f:
        f.write("".join(" %9.5f" % x for x in (0.35, 0.40, 0.35, 0.40, 0.40)) + "\n")
        base = np.array([0.027, 0.43, 0.025, 0.40, 0.16])
        for k, rh in enumerate(rhs):
            f.write(" %05.2f" % rh + "".join(" %9.5f" % x for x in base * (1.0 + 0.12 * k * k / 7.0)) + "\n")
    for i in range(5):
        with open(os.path.join(dirfic, names[1 + i]), "w") as f:
            for wa in (0.2, 0.3, 0.4, 0.55, 0.7, 0.9, 1.1, 1.6, 2.2, 3.0):
                row = [wa]
                for h in range(8):
                    row += [1.53 - 0.02 * h - 0.01 * i + 0.005 * rng.random(), -abs(rng.normal()) * 10.0 ** (-2 - (i % 2)) / (1 + h)]
                f.write("".join(" %9.5f" % x for x in row) + "\n")
    return dirfic


def ref_sf_params(ref, dirfic, wa, rh):
    """SOS_INIT_PARAMSF (SOS_AEROSOLS.F:3557) of the reference library."""
    v1, v2, mr, mi = (np.zeros(5) for _ in range(4))
    ier = C.c_int(99)
    ref.sos_init_paramsf_(C.create_string_buffer(dirfic.encode().ljust(350), 350), _dp(wa), _dp(rh), _P(v1), _P(v2), _P(mr), _P(mi),
                          C.byref(ier), C.c_size_t(350))
    return ier.value, v1, v2, mr, mi
