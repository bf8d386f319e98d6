"""Nadal's BPDF surface file (-SURF.Type 6): SOS_F21SF_NADAL + SOS_CALC_F21_NADAL_SUR_FRESNEL -> SOS_MAT_FRESNEL ->
SOS_MAT_REFLEXION -> SOS_MISE_FORMAT (SOS_SURFACE_BPDF.F:219-392, 686-1223).

CPU part (`-m "not gpu"`): the functions of csrc/nadal_series.cuh, compiled for the host by tests/surface_host.cpp and stepped the
way the kernel gives them to its threads, against SOS_F21SF_NADAL of oracle/_ref/libsosref.so -- series lengths identical and
coefficients BIT-IDENTICAL (same libm on both sides); and the pairing of the reference (below) against SOS_SURFACE_BPDF itself.

The pairing: SOS_F21SF_NADAL writes one series per (I1, I2) for ALL N^2 pairs (:773-782), SOS_MAT_REFLEXION reads the series file
sequentially for the N(N+1)/2 pairs (I, J <= I) without looking at the indices stored in the records (SOS_SURFACE.F:1832-1842).
The p-th pair (I, J) therefore gets the series of (I1, I2) = (p / N + 1, p mod N + 1).  The reference's files are what a drop-in
has to reproduce; the device path does (`pairing="reference"`), and offers the series of the pair itself on request.
GPU part (`-m gpu`): sosgpu_surface_nadal against SOS_SURFACE_BPDF of the reference library."""
import ctypes as C
import importlib
import os
import struct
import subprocess

import numpy as np
import pytest

import refdirect

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_P, _ip, _dp, _fs, _L = refdirect._P, refdirect._ip, refdirect._dp, refdirect._fs, refdirect._L
IND, ALPHA, BETA = 1.5, 0.0159, 44.8                       # alpha, beta of the order of Nadal & Breon's fits (F6.4 / F4.1 fields)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sfh") / "libsfh.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "surface_host.cpp"), "-lm"], check=True)
    lib = C.CDLL(out)
    lib.sfh_nadal_f.restype = C.c_double
    lib.sfh_nadal_f.argtypes = [C.c_double] * 6
    lib.sfh_nadal_series.argtypes = [C.c_double] * 5 + [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return lib


@pytest.fixture(scope="module")
def ref():
    lib = refdirect.lib()
    if lib is None or not hasattr(lib, "sos_f21sf_nadal_"):
        pytest.skip("oracle/_ref/libsosref.so (with SOS_SURFACE_BPDF) not available")
    return lib


def _pkg():
    pkg = importlib.import_module("radiativetransfer-sos_b200")
    return pkg.synth, pkg.formats


def _records(path):
    """gfortran unformatted sequential records (4-byte length before and after)."""
    out, b, k = [], open(path, "rb").read(), 0
    while k < len(b):
        n = struct.unpack_from("<i", b, k)[0]
        out.append(b[k + 4:k + 4 + n])
        assert struct.unpack_from("<i", b, k + 4 + n)[0] == n
        k += n + 8
    return out


def ref_series(ref, tmp, N, rmu, os_nb, ind=IND, alpha=ALPHA, beta=BETA):
    """SOS_F21SF_NADAL -> [(I1, I2, IL, E[0..IL])] in file order."""
    r, _ = refdirect._angles(rmu, np.zeros_like(rmu), N)
    f = os.path.join(tmp, "NADAL_GSF")
    if os.path.exists(f):
        os.remove(f)
    ier = C.c_int(99)
    ref.sos_f21sf_nadal_(_ip(N), _P(r), _dp(ind), _dp(alpha), _dp(beta), _ip(os_nb), _fs(f), _ip(0), C.byref(ier), _L)
    assert ier.value == 0
    out = []
    for rec in _records(f):
        i1, i2, il = struct.unpack_from("<3i", rec, 0)
        e = np.frombuffer(rec, dtype="<f8", offset=12)
        assert e.size == il + 1
        out.append((i1, i2, il, e.copy()))
    return out


def host_series(host, c1, c2, os_nb, ind=IND, alpha=ALPHA, beta=BETA):
    e, b1 = np.zeros(os_nb + 1), np.zeros(os_nb + 1)
    il = host.sfh_nadal_series(ind, alpha, beta, c1, c2, os_nb, float(np.arccos(-1.0)), _P(e), _P(b1))
    return il, e, b1


def ref_surface_bpdf_nadal(ref, fm, tmp, N, rmu, ga, ind, alpha, beta, os_nb, os_ns, os_nm):
    r, g = refdirect._angles(rmu, ga, N)
    f = os.path.join(tmp, "BPDF_NADAL.bin")
    if os.path.exists(f):
        os.remove(f)
    ier = C.c_int(99)
    ref.sos_surface_bpdf_(_ip(N), _P(r), _P(g), _dp(ind), _ip(6), _dp(alpha), _dp(beta), _dp(0.0), _ip(os_nb), _ip(os_ns), _ip(os_nm),
                          _fs(os.path.join(tmp, "N_GSF")), _fs(os.path.join(tmp, "N_FRESNEL")), _fs(os.path.join(tmp, "N_MAT_REFLEX")),
                          _fs(f), _ip(0), C.byref(ier), _L, _L, _L, _L)
    assert ier.value == 0, "reference SOS_SURFACE_BPDF IER=%d" % ier.value
    return fm.read_surface_bin(f, N)


def test_nadal_function_and_series_bit_identical(host, ref, tmp_path):
    syn, fm = _pkg()
    rmu, ga, n0, _ = syn.sos_angles(8, 35.0)
    N = (rmu.size - 1) // 2
    os_nb = 24
    # the function itself against SOS_CALC_F21_NADAL_SUR_FRESNEL
    rng = np.random.default_rng(5)
    for _ in range(200):
        c1, c2 = rmu[N + 1 + rng.integers(N)], rmu[N + 1 + rng.integers(N)]
        phi = float(rng.uniform(0.0, np.pi))
        f = C.c_double(0.0)
        ref.sos_calc_f21_nadal_sur_fresnel_(_dp(IND), _dp(ALPHA), _dp(BETA), _dp(c1), _dp(np.sqrt(1 - c1 * c1)), _dp(c2), _dp(np.sqrt(1 - c2 * c2)),
                                            _dp(phi), C.byref(f))
        assert host.sfh_nadal_f(IND, ALPHA, BETA, c1, c2, phi) == f.value
    recs = ref_series(ref, str(tmp_path), N, rmu, os_nb)
    assert len(recs) == N * N and [(a, b) for a, b, _, _ in recs] == [(i, j) for i in range(1, N + 1) for j in range(1, N + 1)]
    ils = []
    for i1, i2, il, e in recs:
        il_h, e_h, _ = host_series(host, rmu[N + i1], rmu[N + i2], os_nb)
        assert il_h == il, (i1, i2, il_h, il)
        assert np.array_equal(e_h[:il + 1].view(np.uint64), e.view(np.uint64)), (i1, i2)
        ils.append(il)
    assert min(ils) < os_nb and len(set(ils)) > 1                   # the cut is exercised
    # a model whose series does not converge within OS_NB and one that stops at order 0
    for alpha, beta, nb in ((0.02, 80.0, 6), (0.0005, 0.1, 12)):
        rr = ref_series(ref, str(tmp_path), N, rmu, nb, alpha=alpha, beta=beta)
        for i1, i2, il, e in rr[::7]:
            il_h, e_h, _ = host_series(host, rmu[N + i1], rmu[N + i2], nb, alpha=alpha, beta=beta)
            assert il_h == il and np.array_equal(e_h[:il + 1].view(np.uint64), e.view(np.uint64))


def test_nadal_pairing_of_the_reference(host, ref, tmp_path):
    """SOS_SURFACE_BPDF(ISURF = 6) as a whole == SOS_MAT_REFLEXION fed with, for the p-th pair (I, J <= I), the host-stepped series
    of (p / N + 1, p mod N + 1): the pairing the device path reproduces."""
    syn, fm = _pkg()
    tmp = str(tmp_path)
    rmu, ga, n0, _ = syn.sos_angles(6, 40.0)
    N = (rmu.size - 1) // 2
    os_nb, os_ns = 16, 12
    os_nm = os_nb + os_ns
    full = ref_surface_bpdf_nadal(ref, fm, tmp, N, rmu, ga, IND, ALPHA, BETA, os_nb, os_ns, os_nm)
    gsf = os.path.join(tmp, "MY_GSF")
    with open(gsf, "wb") as f:
        for p in range(N * (N + 1) // 2):
            a1, a2 = p // N + 1, p % N + 1
            il, e, _ = host_series(host, rmu[N + a1], rmu[N + a2], os_nb)
            fm._write_record(f, struct.pack("<3i", a1, a2, il) + e[:il + 1].astype("<f8").tobytes())
    refdirect.mat_fresnel(ref, tmp, N, rmu, ga, IND, os_ns)                       # writes tmp/RES_FRESNEL
    r, _ = refdirect._angles(rmu, ga, N)
    ier = C.c_int(99)
    mr, out = os.path.join(tmp, "MY_MAT_REFLEX"), os.path.join(tmp, "MY_BPDF.bin")
    ref.sos_mat_reflexion_(_dp(1.0), _ip(N), _P(r), _ip(os_nb), _ip(os_ns), _ip(os_nm), _fs(os.path.join(tmp, "RES_FRESNEL")), _fs(gsf), _fs(mr),
                           C.byref(ier), _L, _L, _L)
    assert ier.value == 0
    ref.sos_mise_format_(_fs(mr), _fs(out), _ip(N), _ip(os_nb), C.byref(ier), _L, _L)
    assert ier.value == 0
    mine = fm.read_surface_bin(out, N)
    assert np.array_equal(mine.view(np.uint32), full.view(np.uint32))
    assert np.abs(full[:, 0]).max() > 0 and np.abs(full[:, 1]).max() > 0           # P11 and P12 are populated


@pytest.mark.gpu
def test_gpu_surface_nadal_vs_reference(solver, ref, tmp_path):
    """sosgpu_surface_nadal (k_glitter, gmodel 4) against SOS_SURFACE_BPDF(ISURF = 6) of the reference library: series lengths
    identical (the decisions of the cut have margins >= 1e-5 for these models, far above the last-bit differences of the device's
    exp / cos), REAL*4 records identical except for single-precision rounding ties; then + Roujean as SOS_SURFACE does."""
    syn, fm = _pkg()
    tmp = str(tmp_path)
    rmu, ga, n0, _ = syn.sos_angles(12, 35.0)
    N = (rmu.size - 1) // 2
    os_nb, os_ns = 40, 24
    os_nm = os_nb + os_ns
    for alpha, beta in ((ALPHA, BETA), (0.02, 80.0)):
        full = ref_surface_bpdf_nadal(ref, fm, tmp, N, rmu, ga, IND, alpha, beta, os_nb, os_ns, os_nm)
        recs = ref_series(ref, tmp, N, rmu, os_nb, alpha=alpha, beta=beta)
        il_ref = np.array([recs[p][2] for p in range(N * (N + 1) // 2)])          # the series SOS_MAT_REFLEXION reads for pair p
        surf, il = solver.surface_nadal(N, rmu, ga, IND, alpha, beta, os_nb, os_ns, os_nm)
        nbad = int((il != il_ref).sum())
        eq = surf.view(np.uint32) == full.view(np.uint32)
        sig = np.abs(full) > 1e-6 * np.abs(full).max()
        print("\n[Nadal alpha=%g beta=%g N=%d] series-length mismatches %d / %d pairs; REAL*4 records bit-identical: %.4f %% of all entries, "
              "%.4f %% of those above 1e-6 of the largest; max |diff| / max %.1e"
              % (alpha, beta, N, nbad, il.size, 100 * eq.mean(), 100 * eq[sig].mean(), np.abs(surf - full).max() / np.abs(full).max()))
        assert nbad == 0
        assert eq[sig].mean() > 0.99 and np.abs(surf - full).max() <= 2e-7 * np.abs(full).max()
    own, il_own = solver.surface_nadal(N, rmu, ga, IND, alpha, beta, os_nb, os_ns, os_nm, pairing="own")
    assert own.shape == surf.shape and np.isfinite(own).all() and np.array_equal(own[:, :, 0, 0], surf[:, :, 0, 0])   # pair 0 is (1, 1) either way
    ier, rj_ref = refdirect.roujean(ref, fm, tmp, N, rmu, ga, os_nb, 0.25, 0.04, 0.30)
    assert ier == 0
    s_ref = refdirect.bpdf_ajout_brdf(ref, fm, tmp, full, rj_ref)
    s_gpu = solver.bpdf_ajout_brdf(surf, solver.roujean(N, rmu, os_nb, 0.25, 0.04, 0.30))
    assert np.abs(s_gpu - s_ref).max() <= 2e-7 * np.abs(s_ref).max()
