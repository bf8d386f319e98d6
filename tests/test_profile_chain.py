"""The per-term profile chain (SOS_ABSPROFILE + COEFF_ABS_CKD -> SOS_PROFILE + SOS_DISC -> PROFIL_TMP text hop).

CPU part (this file, `-m "not gpu"`): the device functions of csrc/profile_chain.cuh, compiled for the host by
tests/profile_host.cpp, against the reference's own routines in oracle/_ref/libsosref.so on seeded synthetic atmospheres and
CKD tables; the decimal round trip against the C library's printf/strtod; the CKD file reader of libsosgpu.so (host code)
against the tables the files were written from.  The GPU part (same cases through the kernels) is in
tests/test_gpu_vs_reference.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import profile_cases as pc
import refdirect

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
_I = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pch") / "libpch.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "profile_host.cpp"), "-lm"], check=True)
    lib = C.CDLL(out)
    lib.pch_round_e8.restype = C.c_double
    lib.pch_round_e8.argtypes = [C.c_double]
    lib.pch_round_f5.restype = C.c_double
    lib.pch_round_f5.argtypes = [C.c_double]
    return lib


@pytest.fixture(scope="module")
def ref():
    lib = refdirect.lib()
    if lib is None or not hasattr(lib, "coeff_abs_ckd_"):
        pytest.skip("oracle/_ref/libsosref.so (with COEFF_ABS_CKD) not available")
    return lib


def host_absprofile(host, t, user, ro, term):
    tau = np.zeros(50)
    ik = np.array(term["ik"], dtype=np.int32)
    rc = host.pch_absprofile(t["nb_temp"], t["nb_pres"], t["nb_conc"], _P(t["tab_temp"]), _P(t["tab_pres"]), _P(t["tab_conc"]),
                             _I(t["nexp"]), _P(t["ki"]), _P(t["kh"]), _P(user), _P(ro), term["lamb1"], _I(ik), _P(tau))
    return rc, tau


def host_profile(host, altabs, tabs, term, text_hop, wide=0):
    z, h, pa, pm = (np.zeros(pc.NT_MAX + 1) for _ in range(4))
    nt = C.c_int(0)
    rc = host.pch_profile(term["iprofil"], C.c_double(term["tr"]), C.c_double(term["hr"]), C.c_double(term["ta"]), C.c_double(term["ha"]),
                          C.c_double(term["zmin"]), C.c_double(term["zmax"]), term["absprofil"], _P(altabs), _P(tabs), text_hop, wide,
                          C.byref(nt), _P(z), _P(h), _P(pa), _P(pm))
    return rc, nt.value, z, h, pa, pm


def test_round_trip_matches_printf_strtod(host):
    """pc_round_e8 / pc_round_f5 == float('%.7E' % x) / float('%.5f' % x) (what a Fortran E15.8 / F10.5 write + read does)."""
    rng = np.random.default_rng(5)
    xs = np.concatenate([10.0 ** rng.uniform(-30, 3, 20000) * rng.choice([-1.0, 1.0], 20000), rng.random(20000),
                         [0.0, 1.0, 0.1, 0.99999999, 0.999999995, 123456785e-8, 0.5e-7, 1e-22, 9.9999999e-23, 1.5, 999.0,
                          0.125, 2.5e-6, 0.000244140625]])
    bad = 0
    for x in xs:
        want = float("%.7E" % x)
        got = host.pch_round_e8(float(x))
        if got != want:
            bad += 1
            assert abs(got - want) <= 2e-16 * abs(want), (x, got, want)      # beyond 1e-15 the way back is not one rounding
            assert abs(x) < 1e-15, (x, got, want)
    assert bad <= 5, bad
    zs = np.concatenate([rng.uniform(0, 120, 20000), [0.0, 120.0, 119.95, 0.015625, 0.000005, 0.000015, 59.0000050, 1e-7]])
    for z in zs:
        assert host.pch_round_f5(float(z)) == float("%.5f" % z), z


@pytest.mark.parametrize("seed", [0, 1])
def test_absprofile_host_vs_reference(host, ref, seed):
    """SOS_ABSPROFILE: identical optical thickness profiles (bit for bit: same operations in the same order)."""
    user, altabs, ro = pc.gas_atmosphere(seed)
    t = pc.ckd_tables(seed)
    nstrong = nfallback = 0
    for term in pc.make_terms(t, 150, seed):
        ier, want = refdirect.absprofile(ref, t, user, altabs, ro, term)
        rc, got = host_absprofile(host, t, user, ro, term) if term["absprofil"] != 7 else (0, np.zeros(50))
        assert (rc != 0) == (ier != 0), (term, rc, ier)
        if ier == 0:
            assert np.array_equal(got, want), (term, np.abs(got - want).max())
            nstrong += want[-1] > 1.5
    assert nstrong >= 5                                                    # the saturated branch of SOS_PROFILE gets inputs


@pytest.mark.parametrize("wide", [0, 1])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_profile_host_vs_reference(host, ref, seed, wide, tmp_path):
    """SOS_PROFILE: same NT, and the PROFIL_TMP values SOS reads back are identical (the device functions with the text hop
    against the reference's file).  wide = 1: the kernels' search strategy (31 bisection candidates / 32 first-level steps at
    a time), emulated lane by lane, must take exactly the path of the reference's serial loops."""
    user, altabs, ro = pc.gas_atmosphere(seed)
    t = pc.ckd_tables(seed)
    cases = {"nogas": 0, "weak": 0, "strong": 0}
    for term in pc.make_terms(t, 150 if seed < 2 else 60, seed):
        _, tabs = refdirect.absprofile(ref, t, user, altabs, ro, term)
        ier, nt, text, z, h, pa, pm = refdirect.profile(ref, str(tmp_path), altabs, tabs, term)
        rc, nt2, z2, h2, pa2, pm2 = host_profile(host, altabs, tabs, term, 1, wide)
        assert (rc != 0) == (ier != 0), (term, rc, ier)
        if ier != 0:
            continue
        assert nt2 == nt, (term, nt2, nt)
        for a, b, what in ((z2, z, "zprof"), (h2, h, "h"), (pa2, pa, "pcaer"), (pm2, pm, "pcmol")):
            assert np.array_equal(a[:nt + 1], b), (term, what, np.abs(a[:nt + 1] - b).max())
        cases["nogas" if tabs[-1] == 0 else "strong" if tabs[-1] > 1.5 else "weak"] += 1
    assert min(cases.values()) >= 2, cases


def test_profile_too_many_levels_is_an_error(host):
    """2 (TR + TA) + tau_gas above about 3 needs more than CTE_OS_NT = 600 levels: the reference writes past its arrays there
    (observed: it never returns); the device function stops with 9600."""
    user, altabs, ro = pc.gas_atmosphere(0)
    tabs = np.linspace(0.0, 1.31, 50)
    term = dict(iprofil=1, absprofil=2, tr=0.345, hr=8.0, ta=1.17, ha=2.25, zmin=0.0, zmax=5.0)
    rc, nt, *_ = host_profile(host, altabs, tabs, term, 1)
    assert rc == 9600 and nt == 0
    term = dict(term, tr=3.5, ta=0.0)                                      # already the profile without gas: 700 levels
    assert host_profile(host, altabs, tabs, term, 1)[0] == 9600


def test_profile_iprofil2_host_vs_reference(host, ref, tmp_path):
    """IPROFIL = 2 (aerosols between two altitudes, SOS_PROFIL.F:800-905), including its error exits."""
    user, altabs, ro = pc.gas_atmosphere(3)
    t = pc.ckd_tables(3)
    tabs = np.zeros(50)
    n_ok = 0
    src = os.path.join(ROOT, "oracle", "_ref", "libsosref.so")
    for n, term in enumerate(pc.make_terms(t, 40, 3, iprofil=2)):
        term["absprofil"] = 7
        # Hmol(0) is read before it is set (SOS_PROFIL.F:857): the reference relies on fresh storage, which the translated
        # library (static locals) only has on the first call of a process image -> a fresh copy of the library per case
        fresh = str(tmp_path / ("libsosref_%d.so" % n))
        shutil.copy(src, fresh)
        ier, nt, text, z, h, pa, pm = refdirect.profile(C.CDLL(fresh), str(tmp_path), altabs, tabs, term)
        os.remove(fresh)
        rc, nt2, z2, h2, pa2, pm2 = host_profile(host, altabs, tabs, term, 1)
        if ier != 0:
            assert rc != 0, term
            continue
        if rc != 0:                                                        # NBSC_C2 <= 0: the reference runs on with garbage
            continue
        n_ok += 1
        assert nt2 == nt
        for a, b in ((z2, z), (h2, h), (pa2, pa), (pm2, pm)):
            assert np.array_equal(a[:nt + 1], b), term
    assert n_ok >= 10


def test_ckd_reader_round_trip(tmp_path):
    """sosgpu_read_ckd_coeff (host code of libsosgpu.so, READ_CKD_COEFF SOS_SUB_TRS.F:481-905) on files written from known
    tables: every array comes back, for all 8 gases; the not-selected default; a real reference file when it is present."""
    so = os.path.join(ROOT, "radiativetransfer-sos_b200", "libsosgpu.so")
    if not os.path.exists(so):
        pytest.skip("libsosgpu.so not built")
    lib = C.CDLL(so)
    t = pc.ckd_tables(4)
    pc.write_ckd_files(str(tmp_path), t)
    nexp = np.zeros((pc.NBABS, pc.NWVL), dtype=np.int32, order="F")
    ai = np.zeros((pc.NAI, pc.NBABS, pc.NWVL), order="F")
    ki = np.zeros_like(t["ki"]); kh = np.zeros_like(t["kh"])
    tp, tt, tc = np.zeros(pc.NPMAX), np.zeros(pc.NTMAX), np.zeros(pc.NCMAX)
    numax, numin = C.c_double(0), C.c_double(0)
    nbp, nbt, nbc = C.c_int(0), C.c_int(0), C.c_int(0)
    for k in range(1, 9):
        rc = lib.sosgpu_read_ckd_coeff(str(tmp_path).encode(), k, 1, C.c_double(13255.0), C.c_double(10.0), _I(nexp), _P(ai), _P(ki), _P(kh),
                                       C.byref(numax), C.byref(numin), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc), C.byref(nbc))
        assert rc == 0, k
        assert (numax.value, numin.value) == (13500.0, 13000.0)
    assert (nbp.value, nbt.value, nbc.value) == (pc.NPMAX, pc.NTMAX, pc.NCMAX)
    assert np.array_equal(nexp, t["nexp"])
    assert np.array_equal(tp, t["tab_pres"]) and np.array_equal(tt, t["tab_temp"]) and np.array_equal(tc, t["tab_conc"])
    assert np.array_equal(ai, t["ai"]) and np.array_equal(ki, t["ki"]) and np.array_equal(kh, t["kh"])
    # gas not selected: one exponential of weight 1 and k = 0 everywhere (SOS_SUB_TRS.F:574-597)
    rc = lib.sosgpu_read_ckd_coeff(str(tmp_path).encode(), 7, 0, C.c_double(13255.0), C.c_double(10.0), _I(nexp), _P(ai), _P(ki), _P(kh),
                                   C.byref(numax), C.byref(numin), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc), C.byref(nbc))
    assert rc == 0 and (nexp[6] == 1).all() and (ai[0, 6] == 1.0).all() and not ki[:, :, 0, 6, :].any()
    # wrong resolution / missing file -> -1 (the reference's IER)
    assert lib.sosgpu_read_ckd_coeff(str(tmp_path).encode(), 7, 1, C.c_double(13255.0), C.c_double(2.0), _I(nexp), _P(ai), _P(ki), _P(kh),
                                     C.byref(numax), C.byref(numin), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc), C.byref(nbc)) == -1
    assert lib.sosgpu_read_ckd_coeff(str(tmp_path).encode(), 7, 1, C.c_double(5000.0), C.c_double(10.0), _I(nexp), _P(ai), _P(ki), _P(kh),
                                     C.byref(numax), C.byref(numin), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc), C.byref(nbc)) == -1
    real = "/root/reference/fic/COEFF_CKD/10cmm1/coef_O2_13500_13000_10cmm1"
    if os.path.exists(real):                                               # the reference's own O2 A-band file, parsed independently
        rc = lib.sosgpu_read_ckd_coeff(b"/root/reference", 7, 1, C.c_double(13100.0), C.c_double(10.0), _I(nexp), _P(ai), _P(ki), _P(kh),
                                       C.byref(numax), C.byref(numin), _P(tp), C.byref(nbp), _P(tt), C.byref(nbt), _P(tc), C.byref(nbc))
        assert rc == 0
        toks = open(real).read().split("\n", 18)[18].split()
        pos = 3
        nt_ = int(toks[pos]); pos += 1
        assert np.array_equal(tt[:nt_], np.array(toks[pos:pos + nt_], dtype=float)); pos += nt_
        np_ = int(toks[pos]); pos += 1
        assert np.array_equal(tp[:np_], np.array(toks[pos:pos + np_], dtype=float)); pos += np_
        for l in range(50):
            n = int(toks[pos + 5]); pos += 6
            if n == 0:
                assert nexp[6, l] == 1 and ai[0, 6, l] == 1.0 and not ki[:, :, 0, 6, l].any()
                continue
            assert nexp[6, l] == n
            assert np.array_equal(ai[:n, 6, l], np.array(toks[pos:pos + n], dtype=float)); pos += n
            for i in range(n):
                for p in range(np_):
                    assert (int(toks[pos]), int(toks[pos + 1])) == (i + 1, p + 1); pos += 2
                    assert np.array_equal(ki[:nt_, p, i, 6, l], np.array(toks[pos:pos + nt_], dtype=float)); pos += nt_
        assert pos == len(toks)
