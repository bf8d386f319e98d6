"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (radiativetransfer-sos_b200) never does.

PIN: bit-identical to oracle/_ref/libsosref.so, the reference's own Fortran sources translated mechanically to C
(oracle/build_ref.py, tests/test_oracle_vs_reference.py); no Fortran compiler exists in this image and the reference
ships no golden vectors, so a gfortran build is not part of the pin (see oracle/sos_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_fp = C.POINTER(C.c_float)


def build(force=False):
    """Compile liboracle.so with gcc (seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_sos_os.restype = C.c_int
        _LIB.orc_sos.restype = C.c_int
        _LIB.orc_aggregate.restype = C.c_int
        _LIB.orc_trphi.restype = C.c_int
        _LIB.orc_trphi_option.restype = C.c_int
    return _LIB


def _d(a):
    return a.ctypes.data_as(c_dp)


def _dn(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(np.array(a, dtype=np.float64, copy=True))


def noyaux(is_, rmu, os_nb, alpha, beta, gamma, zeta):
    """SOS_NOYAUX. rmu: [2N+1] with rmu[N]=mu_s. Returns dict of arrays; kernels as [2N+1(k), 2N+1(j)]
    numpy arrays such that K[k+N, j+N] == Fortran K(j,k)."""
    rmu = _f64(rmu)
    W = rmu.size
    N = (W - 1) // 2
    a, b, g, z = (_f64(x) for x in (alpha, beta, gamma, zeta))
    out = {n: np.zeros(W) for n in ("xpl", "xrl", "xtl")}
    ker = {n: np.zeros((W, W)) for n in ("bp", "gr", "gt", "arr", "art", "att")}
    lib().orc_noyaux(C.c_int(is_), C.c_int(N), _d(rmu), C.c_int(os_nb), _d(a), _d(b), _d(g), _d(z),
                     _d(out["xpl"]), _d(out["xrl"]), _d(out["xtl"]),
                     _d(ker["bp"]), _d(ker["gr"]), _d(ker["gt"]), _d(ker["arr"]), _d(ker["art"]), _d(ker["att"]))
    out.update(ker)
    return out


def fsource_ordreig(is_, nt, xdel, ydel, beta0, beta2, gamma2, alpha2, ker, i1, q1, u1, ga):
    """SOS_FSOURCE_ORDREIG. fields are [2N+1, NT+1] arrays (F[k+N, i])."""
    W = ga.size
    N = (W - 1) // 2
    i2, q2, u2 = (np.zeros((W, nt + 1)) for _ in range(3))
    lib().orc_fsource_ordreig(C.c_int(is_), C.c_int(N), C.c_int(nt), _d(_f64(xdel)), _d(_f64(ydel)),
                              C.c_double(beta0), C.c_double(beta2), C.c_double(gamma2), C.c_double(alpha2),
                              _d(ker["xpl"]), _d(ker["xrl"]), _d(ker["xtl"]),
                              _d(_f64(i1)), _d(_f64(q1)), _d(_f64(u1)),
                              _d(ker["bp"]), _d(ker["gr"]), _d(ker["gt"]), _d(ker["arr"]), _d(ker["art"]),
                              _d(ker["att"]), _d(_f64(ga)), _d(i2), _d(q2), _d(u2))
    return i2, q2, u2


def integr_epopt(rmu, nt, h, i2, q2, u2, i1, q1, u1):
    """SOS_INTEGR_EPOPT; i1,q1,u1 carry the ground boundary values (k>0, level NT) in; returns new fields."""
    rmu = _f64(rmu)
    N = (rmu.size - 1) // 2
    i1, q1, u1 = _f64(i1), _f64(q1), _f64(u1)
    lib().orc_integr_epopt(C.c_int(N), _d(rmu), C.c_int(nt), _d(_f64(h)), _d(_f64(i2)), _d(_f64(q2)), _d(_f64(u2)),
                           _d(i1), _d(q1), _d(u1))
    return i1, q1, u1


class OsResult:
    pass


def sos_os(nbmu, rmu, ga, os_nb, nt, n0, tetas, ro, imat_surf, ifresnel, ind_surf, h, xdel, ydel, zprof, ron,
           alpha, beta, gamma, zeta, zout, igmax, iborm, ipolar, surf=None):
    """SOS_OS in memory.  Arrays are copied (the reference mutates rmu(0) and alpha/gamma/zeta)."""
    N = nbmu
    W = 2 * N + 1
    rmu, ga = _f64(rmu), _f64(ga)
    a, b, g, z = (_f64(x) for x in (alpha, beta, gamma, zeta))
    rec = np.zeros((iborm + 1, 3, W))
    nf = C.c_int(0)
    nsc = np.zeros(iborm + 1, dtype=np.int32)
    rsn = np.full(iborm + 1, -1, dtype=np.int32)
    em, ep = C.c_double(0), C.c_double(0)
    sp = None
    if imat_surf == 1:
        surf = np.ascontiguousarray(surf, dtype=np.float32)
        assert surf.size >= (iborm + 1) * 9 * N * N
        sp = surf.ctypes.data_as(c_fp)
    ier = lib().orc_sos_os(C.c_int(N), _d(rmu), _d(ga), C.c_int(os_nb), C.c_int(nt), C.c_int(n0), C.c_double(tetas),
                           C.c_double(ro), C.c_int(imat_surf), C.c_int(ifresnel), C.c_double(ind_surf),
                           _d(_f64(h)), _d(_f64(xdel)), _d(_f64(ydel)), _d(_f64(zprof)), C.c_double(ron),
                           _d(a), _d(b), _d(g), _d(z), C.c_double(zout), C.c_int(igmax), C.c_int(iborm),
                           C.c_int(ipolar), sp, _d(rec), C.byref(nf), nsc.ctypes.data_as(c_ip),
                           rsn.ctypes.data_as(c_ip), C.byref(em), C.byref(ep))
    r = OsResult()
    r.ier, r.n_fourier = ier, nf.value
    r.rec = rec[:nf.value].copy()
    r.n_scatter, r.stop_reason = nsc[:nf.value].copy(), rsn[:nf.value].copy()
    r.emoins, r.eplus = em.value, ep.value
    r.rmu_out = rmu
    return r


def sos(nt, zout, igmax, ipolar, ron, ind_surf, rho, imat_surf, ifresnel, surf, n0, piz, piztr, a, rmu, ga, tetas,
        os_nb, nbmu, alpha, beta, gamma, zeta, zprof, h, pcaer, pcmol, want_trans=False):
    """SOS (SOS.F:340) in memory: truncation adaptation + SOS_OS (+ optional transmissions)."""
    N = nbmu
    W = 2 * N + 1
    rmu, ga = _f64(rmu), _f64(ga)
    al, be, gm, ze = (_f64(x) for x in (alpha, beta, gamma, zeta))
    rec = np.zeros((os_nb + 1, 3, W))
    nf = C.c_int(0)
    nsc = np.zeros(os_nb + 1, dtype=np.int32)
    rsn = np.full(os_nb + 1, -1, dtype=np.int32)
    sc = [C.c_double(0) for _ in range(6)]  # ttot_tronc, ttot_vrai, tauout, tdifmus, emoins, eplus
    tdifmug = np.zeros(W)
    L = nt + 1
    h_tr, xdel_tr, ydel_tr = np.zeros(L), np.zeros(L), np.zeros(L)
    sp = None
    if imat_surf == 1:
        surf = np.ascontiguousarray(surf, dtype=np.float32)
        sp = surf.ctypes.data_as(c_fp)
    ier = lib().orc_sos(C.c_int(nt), C.c_double(zout), C.c_int(igmax), C.c_int(ipolar), C.c_double(ron),
                        C.c_double(ind_surf), C.c_double(rho), C.c_int(imat_surf), C.c_int(ifresnel), sp,
                        C.c_int(n0), C.c_double(piz), C.c_double(piztr), C.c_double(a), _d(rmu), _d(ga),
                        C.c_double(tetas), C.c_int(os_nb), C.c_int(N), _d(al), _d(be), _d(gm), _d(ze),
                        _d(_f64(zprof)), _d(_f64(h)), _d(_f64(pcaer)), _d(_f64(pcmol)), C.c_int(int(want_trans)),
                        _d(rec), C.byref(nf), nsc.ctypes.data_as(c_ip), rsn.ctypes.data_as(c_ip),
                        C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]), C.byref(sc[3]), _d(tdifmug),
                        C.byref(sc[4]), C.byref(sc[5]), _d(h_tr), _d(xdel_tr), _d(ydel_tr))
    r = OsResult()
    r.ier, r.n_fourier = ier, nf.value
    r.rec = rec[:nf.value].copy()
    r.n_scatter, r.stop_reason = nsc[:nf.value].copy(), rsn[:nf.value].copy()
    r.ttot_tronc, r.ttot_vrai, r.tauout, r.tdifmus, r.emoins, r.eplus = (x.value for x in sc)
    r.tdifmug = tdifmug
    r.h, r.xdel, r.ydel = h_tr, xdel_tr, ydel_tr
    return r


class Aggregate:
    """Running CKD aggregate: mirrors the caller-owned accumulators of SOS_PROC.F:1292-1302 + SOS_AGGREGATE."""

    def __init__(self, nbmu, max_rec):
        self.N = nbmu
        self.W = 2 * nbmu + 1
        self.res = np.zeros((max_rec + 64, 3, self.W))
        self.nres = 0
        self.have = False
        self.sc = dict(ttot_tronc=0.0, ttot_vrai=0.0, tauout=0.0, tdifmus=0.0, emoins=0.0, eplus=0.0)
        self.tdifmug = np.zeros(self.W)

    def add(self, aik, r):
        v = {k: C.c_double(x) for k, x in self.sc.items()}
        tmp = np.ascontiguousarray(r.rec)
        tg = getattr(r, "tdifmug", np.zeros(self.W))
        n = lib().orc_aggregate(C.c_int(self.N), C.c_double(aik), _d(tmp), C.c_int(r.n_fourier), _d(self.res),
                                C.c_int(self.nres), C.c_int(int(self.have)),
                                C.c_double(r.ttot_tronc), C.c_double(r.ttot_vrai), C.c_double(r.tauout),
                                C.c_double(getattr(r, "tdifmus", 0.0)), _d(_f64(tg)), C.c_double(r.emoins),
                                C.c_double(r.eplus),
                                C.byref(v["ttot_tronc"]), C.byref(v["ttot_vrai"]), C.byref(v["tauout"]),
                                C.byref(v["tdifmus"]), _d(self.tdifmug), C.byref(v["emoins"]), C.byref(v["eplus"]))
        self.nres = n
        self.have = True
        self.sc = {k: x.value for k, x in v.items()}
        return n


def trphi(rec, nbmu, rmu, tau, tauout, phi, igli, n0, wind, ind_surf, ifresnel, ipolar):
    W = 2 * nbmu + 1
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    xi, xq, xu, ang = (np.zeros(W) for _ in range(4))
    ier = lib().orc_trphi(_d(rec), C.c_int(rec.shape[0]), C.c_int(nbmu), _d(_f64(rmu)), C.c_double(tau),
                          C.c_double(tauout), C.c_double(phi), C.c_int(igli), C.c_int(n0), C.c_double(wind),
                          C.c_double(ind_surf), C.c_int(ifresnel), C.c_int(ipolar), _d(xi), _d(xq), _d(xu), _d(ang))
    return ier, xi, xq, xu, ang


def trphi_option(rec, nbmu, rmu, tau, tauout, igli, n0, wind, ind_surf, ifresnel, itrphi, phios, pas_phi, ipolar):
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    cap = 2 if itrphi == 1 else 360 // max(pas_phi, 1) + 1
    phi_fin = np.zeros(cap)
    theta = np.zeros(nbmu)
    up = np.zeros((7, cap, nbmu))
    down = np.zeros((7, cap, nbmu))
    n = lib().orc_trphi_option(_d(rec), C.c_int(rec.shape[0]), C.c_int(nbmu), _d(_f64(rmu)), C.c_double(tau),
                               C.c_double(tauout), C.c_int(igli), C.c_int(n0), C.c_double(wind),
                               C.c_double(ind_surf), C.c_int(ifresnel), C.c_int(itrphi), C.c_double(phios),
                               C.c_int(pas_phi), C.c_int(ipolar), _d(phi_fin), _d(theta), _d(up), _d(down),
                               C.c_int(cap))
    return n, phi_fin[:max(n, 0)], theta, up[:, :max(n, 0)], down[:, :max(n, 0)]


def gsf_pair(c1, c2, sig, os_nm):
    """SOS_GSF for one pair: (IL, E[0:os_nm+1])."""
    e = np.zeros(os_nm + 1)
    lib().orc_gsf_pair.restype = C.c_int
    il = lib().orc_gsf_pair(C.c_double(c1), C.c_double(c2), C.c_double(sig), C.c_int(os_nm), _d(e))
    return il, e


def mat_fresnel(nbmu, rmu, chr_, ind, os_ns):
    out = [np.zeros(os_ns + 1) for _ in range(4)]
    lib().orc_mat_fresnel(C.c_int(nbmu), _d(_f64(rmu)), _d(_f64(chr_)), C.c_double(ind), C.c_int(os_ns),
                          _d(out[0]), _d(out[1]), _d(out[2]), _d(out[3]))
    return dict(alpha=out[0], beta=out[1], gamma=out[2], zeta=out[3])


def glitter(nbmu, rmu, chr_, wind, ind, os_nb, os_ns, os_nm):
    """SOS_GLITTER: surface-file records [os_nb+1, 9, N(J), N(I)] REAL*4 and the G-series lengths IL per pair."""
    surf = np.zeros((os_nb + 1, 9, nbmu, nbmu), dtype=np.float32)
    il = np.zeros(nbmu * (nbmu + 1) // 2, dtype=np.int32)
    lib().orc_glitter.restype = C.c_int
    ier = lib().orc_glitter(C.c_int(nbmu), _d(_f64(rmu)), _d(_f64(chr_)), C.c_double(wind), C.c_double(ind),
                            C.c_int(os_nb), C.c_int(os_ns), C.c_int(os_nm), surf.ctypes.data_as(c_fp),
                            il.ctypes.data_as(c_ip))
    assert ier == 0
    return surf, il
