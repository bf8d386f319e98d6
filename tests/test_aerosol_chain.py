"""Aerosol optics per wavelength (SOS_MIE + SOS_FPHASE_MIE -> SOS_GRANU -> mixture -> SOS_DECOMPO_LEGENDRE).

CPU part (`-m "not gpu"`): the functions of csrc/aerosol_chain.cuh, compiled for the host by tests/aerosol_host.cpp (one thread,
no barrier), against the reference's own routines in oracle/_ref/libsosref.so -- required BIT-IDENTICAL (same libm on both sides);
the mixture against the statements of SOS_AEROSOLS.F:2085-2110 restated in numpy; the result-file writer of libsosgpu.so (host
code) against the formats of SOS_AEROSOLS.F:3043-3052 and the way SOS_PREPA_OS.F:669-693 reads the file back.
GPU part (`-m gpu`): the same cases through the kernels of sosgpu_aerosols.cu and through the gfortran-ABI symbols; there the
device's sin / cos / exp / log / pow differ from glibc's in the last bit, so REAL*4 records are compared as "identical except
for single-precision rounding ties" and doubles to 1e-11."""
import ctypes as C
import os
import re
import subprocess
import time

import numpy as np
import pytest

import aerosol_cases as ac
import refdirect

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_P = ac._P
_F = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
_I = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
NAMES = ["alp", "beta11", "gamma12", "zeta", "beta22", "delta33"]          # order of the coefficient blocks of the C ABI
TABLES = [(1.45, -0.004, 0.0001, 200.0), (1.33, 0.0, 0.0001, 60.0), (1.75, -0.44, 0.0001, 1200.0)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ach") / "libach.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "aerosol_host.cpp"), "-lm"], check=True)
    lib = C.CDLL(out)
    lib.ach_mie_count.argtypes = [C.c_double, C.c_double]
    return lib


@pytest.fixture(scope="module")
def ref():
    lib = refdirect.lib()
    if lib is None or not hasattr(lib, "sos_mie_") or not hasattr(lib, "sos_granu_"):
        pytest.skip("oracle/_ref/libsosref.so (with SOS_MIE, SOS_GRANU) not available")
    return lib


def host_mie(host, nbmu, xmu, rn, in_, a0, af):
    n = host.ach_mie_count(a0, af)
    nang = 2 * nbmu + 1
    rec, g = np.zeros((n, 3), np.float32), np.zeros(n)
    im, qm, um = (np.zeros((n, nang), np.float32) for _ in range(3))
    k = host.ach_mie(nbmu, _P(xmu), C.c_double(rn), C.c_double(in_), C.c_double(a0), C.c_double(af), n, _F(rec), _P(g), _F(im), _F(qm), _F(um))
    assert k == n
    return dict(rec=rec, g=g, imie=im, qmie=qm, umie=um, alphaf=af)


def host_granu(host, nbmu, t, ig, v1, v2, v3, wa):
    nang = 2 * nbmu + 1
    out, p11, p12, p33 = np.zeros(3), np.zeros(nang), np.zeros(nang), np.zeros(nang)
    e = host.ach_granu(t["rec"].shape[0], _F(t["rec"]), _F(t["imie"]), _F(t["qmie"]), _F(t["umie"]), nang, C.c_double(t["alphaf"]), ig,
                       C.c_double(v1), C.c_double(v2), C.c_double(v3), C.c_double(wa), _P(out), _P(p11), _P(p12), _P(p33))
    return e, out[0], out[1], out[2], p11, p12, p33


def host_model(host, nbmu, xmu, xhr, comp_k, p11c, p12c, p33c, ncomp, comp, w, itronc, os_nb, p22c=None):
    nang = 2 * nbmu + 1
    scal, coef, ph = np.zeros(8), np.zeros((6, os_nb + 1)), np.zeros((4, nang))
    ci, ww = np.zeros(4, np.int32), np.zeros(4)
    ci[:len(comp)], ww[:len(w)] = comp, w
    e = host.ach_model(nbmu, _P(xmu), _P(xhr), _P(np.ascontiguousarray(comp_k)), _P(np.ascontiguousarray(p11c)),
                       _P(np.ascontiguousarray(p12c)), _P(np.ascontiguousarray(p33c)), _P(p22c) if p22c is not None else None,
                       ncomp, _I(ci), _P(ww), itronc, os_nb, _P(scal), _P(coef), _P(ph))
    return e, scal, coef, ph


def components_at(wa, comps):
    """(rn, in, alpha0, alphaf, igranu, v1, v2, v3, wa) per component, ALPHAF as SOS_AEROSOLS derives it (:1944-1945, :1182)."""
    out = []
    for rn, in_, ig, v1, v2, v3 in comps:
        rmax = ac.lnd_rmax(v1, v2) if ig == 1 else v3
        out.append((rn, in_, 0.0001, ac.alphaf_for(rmax, wa), ig, v1, v2, v3, wa))
    return out


def mix_numpy(w, k1c, k2c, p11c, p12c, p33c):
    """SOS_AEROSOLS.F:2085-2110 (bimodal) = :1455-1490 (WMO): the mixture of the components' cross sections and phase functions."""
    k1 = k2 = 0.0
    p = [np.zeros_like(p11c[0]) for _ in range(3)]
    for i in range(len(w)):
        if w[i] == 0.0:
            continue
        k1 = k1 + w[i] * k1c[i]
        k2 = k2 + w[i] * k2c[i]
        for q, src in zip(p, (p11c, p12c, p33c)):
            q += w[i] * src[i] * k2c[i]
    return k1, k2, [q / k2 for q in p]


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("case", TABLES)
def test_host_mie_bit_identical(host, ref, tmp_path, case):
    """SOS_MIE records (REAL*4 ALPHA, QEXT, QSCA and phase functions, REAL*8 G): 0 differing values.  The third table reaches
    size parameter 1200, where C_n diverges before 2 alpha + 5 (series cut, SOS_MIE.F:463-467) and S_n is renormalised
    (:509-515)."""
    rn, in_, a0, af = case
    nbmu, xmu, xhr = ac.mie_angles(12, (0.0, 30.0))
    _, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, rn, in_, a0, af)
    h = host_mie(host, nbmu, xmu, rn, in_, a0, af)
    assert h["rec"].shape[0] == r["rec"].shape[0]
    for k in ("rec", "g", "imie", "qmie", "umie"):
        assert np.array_equal(h[k], r[k]), k
    print("[Mie m=%g%+gi alphaf=%g] %d records x %d angles identical to SOS_MIE" % (rn, in_, af, r["g"].size, 2 * nbmu + 1))


def test_host_mie_lane_layout_bit_identical(host, ref, tmp_path):
    """The layout the kernels run -- work list of all tables sorted by size parameter, 32 size parameters per warp with interleaved
    work arrays, a_n / b_n in place, chunks that fit an arena budget -- stepped lane by lane on the host with the kernels' own plan
    and phase functions: still 0 differing values against SOS_MIE, with one chunk and with many."""
    nbmu, xmu, xhr = ac.mie_angles(6, (0.0,))
    nang = 2 * nbmu + 1
    for tables in (TABLES, TABLES[:2] + [(1.5, -0.01, 0.5, 30.0)]):          # one shared grid (linear-time plan) / grids that differ (sort)
        _lane_layout_case(host, ref, tmp_path, nbmu, xmu, xhr, nang, tables)


def _lane_layout_case(host, ref, tmp_path, nbmu, xmu, xhr, nang, TABLES):
    refs = [ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, *t, name="L%d.bin" % i)[1] for i, t in enumerate(TABLES)]
    n = sum(r["g"].size for r in refs)
    tab = np.array(TABLES).ravel().copy()
    for budget in (1 << 40, 3_000_000):
        rec, g = np.zeros((n, 3), np.float32), np.zeros(n)
        im, qm, um = (np.zeros((n, nang), np.float32) for _ in range(3))
        k = host.ach_mie_lanes(nbmu, _P(xmu), len(TABLES), _P(tab), C.c_longlong(budget), n, _F(rec), _P(g), _F(im), _F(qm), _F(um))
        assert k == n
        o = 0
        for r in refs:
            m = r["g"].size
            for a, b in ((rec, "rec"), (g, "g"), (im, "imie"), (qm, "qmie"), (um, "umie")):
                assert np.array_equal(a[o:o + m], r[b]), (budget, b)
            o += m


def test_host_mie_largest_table_bit_identical(host, ref, tmp_path):
    """The largest table inc/SOS.h allows (2 alphaf + 20 <= CTE_MIE_DIM = 10000): 8 790 size parameters up to 4990, series of up to
    9 985 orders, through the lane layout of the kernels -- 0 differing values against SOS_MIE."""
    nbmu, xmu, xhr = ac.mie_angles(2)
    nang = 2 * nbmu + 1
    case = (1.53, -0.008, 0.0001, 4990.0)
    r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, *case)[1]
    n = r["g"].size
    assert n == 8790
    rec, g = np.zeros((n, 3), np.float32), np.zeros(n)
    im, qm, um = (np.zeros((n, nang), np.float32) for _ in range(3))
    tab = np.array(case)
    assert host.ach_mie_lanes(nbmu, _P(xmu), 1, _P(tab), C.c_longlong(1 << 40), n, _F(rec), _P(g), _F(im), _F(qm), _F(um)) == n
    for a, b in ((rec, "rec"), (g, "g"), (im, "imie"), (qm, "qmie"), (um, "umie")):
        assert np.array_equal(a, r[b]), b


def test_host_mie_count_is_grid(host):
    assert host.ach_mie_count(0.0001, 200.0) == 4000
    assert host.ach_mie_count(0.0001, 4990.0) == 3900 + 4890
    assert host.ach_mie_count(0.0001, 4991.0) == -1                  # 2 alphaf + 20 > CTE_MIE_DIM


def test_host_granu_decompo_bit_identical(host, ref, tmp_path):
    """SOS_GRANU (log-normal coarse / fine modes, Junge law) on the reference's own MIE files and SOS_DECOMPO_LEGENDRE with and
    without truncation (the fine mode and the Junge law run into the 'truncation too weak: cancelled' branch): 0 differing
    doubles."""
    nbmu, xmu, xhr = ac.mie_angles(24, (0.0, 30.0))
    os_nb, wa = 48, 0.55
    for ci, c in enumerate(components_at(wa, [ac.COARSE, ac.FINE, ac.JUNGE])):
        rn, in_, a0, af, ig, v1, v2, v3, _ = c
        f, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, rn, in_, a0, af, "MIE%d.bin" % ci)
        e0, k1, k2, snr, p11, p12, p33 = ac.ref_granu(ref, f, ig, v1, v2, v3, wa, nbmu, xmu)
        e1, h1, h2, hs, q11, q12, q33 = host_granu(host, nbmu, r, ig, v1, v2, v3, wa)
        assert e0 == 0 and e1 == 0
        assert (h1, h2, hs) == (k1, k2, snr)
        assert np.array_equal(q11, p11) and np.array_equal(q12, p12) and np.array_equal(q33, p33)
        cancelled = 0
        for itr in (0, 1):
            d = ac.ref_decompo(ref, itr, nbmu, xmu, xhr, os_nb, p11, p12, p11, p33)
            e, scal, coef, ph = host_model(host, nbmu, xmu, xhr, np.array([k1, k2, snr]), q11, q12, q33, 0, [0], [1.0], itr, os_nb)
            assert e == 0 and int(scal[7]) == d["itronc"] and scal[4] == d["coef_tronca"] and scal[6] == d["z1"]
            for i, n in enumerate(NAMES):
                assert np.array_equal(coef[i], d[n]), (ci, itr, n)
            assert np.array_equal(ph[0], d["p11"]) and np.array_equal(ph[3], d["ttt"])
            cancelled += itr == 1 and d["itronc"] == 0
        print("[GRANU/DECOMPO component %d] identical; truncation %s" % (ci, "cancelled" if cancelled else "applied"))


def test_host_granu_short_table(host, ref, tmp_path):
    """A table that ends before the size distribution does: the reference reads past the end of its file (IER = -1)."""
    nbmu, xmu, xhr = ac.mie_angles(6)
    f, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, 1.4, -0.01, 0.0001, 20.0)
    cut = {k: (v[:-40] if isinstance(v, np.ndarray) else v) for k, v in r.items()}
    assert host_granu(host, nbmu, cut, 1, 0.3, 0.5, 0.0, 0.55)[0] == -1
    assert host_granu(host, nbmu, r, 1, 0.3, 0.5, 0.0, 0.55)[0] == 0
    # Junge law leaving at r > rmax before the end of the table
    e0, k1, k2, snr, p11, _, _ = ac.ref_granu(ref, f, 2, 0.05, 4.0, 1.0, 0.55, nbmu, xmu)
    e1, h1, h2, hs, q11, _, _ = host_granu(host, nbmu, r, 2, 0.05, 4.0, 1.0, 0.55)
    assert e0 == 0 and e1 == 0 and (h1, h2, hs) == (k1, k2, snr) and np.array_equal(p11, q11)


def test_host_mixture_and_decompo_with_p22(host, ref, tmp_path):
    """Bimodal mixture (SOS_AEROSOLS.F:2085-2110) restated in numpy -> SOS_DECOMPO_LEGENDRE of the reference; and the
    expansion with a P22 that differs from P11 (non-spherical data path of the routine's interface)."""
    nbmu, xmu, xhr = ac.mie_angles(20)
    os_nb, wa = 40, 0.865
    comps = components_at(wa, [ac.COARSE, ac.FINE])
    kc, pc = [], []
    for ci, (rn, in_, a0, af, ig, v1, v2, v3, _) in enumerate(comps):
        f, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, rn, in_, a0, af, "M%d.bin" % ci)
        e, k1, k2, snr, p11, p12, p33 = ac.ref_granu(ref, f, ig, v1, v2, v3, wa, nbmu, xmu)
        kc.append([k1, k2, snr]); pc.append([p11, p12, p33])
    kc, pc = np.array(kc), np.array(pc)
    cv = np.array([0.3, 0.7])
    w = cv / (cv[0] + cv[1])
    k1, k2, (p11, p12, p33) = mix_numpy(w, kc[:, 0], kc[:, 1], pc[:, 0], pc[:, 1], pc[:, 2])
    for itr in (0, 1):
        d = ac.ref_decompo(ref, itr, nbmu, xmu, xhr, os_nb, p11, p12, p11, p33)
        e, scal, coef, ph = host_model(host, nbmu, xmu, xhr, kc, pc[:, 0], pc[:, 1], pc[:, 2], 2, [0, 1], w, itr, os_nb)
        assert e == 0 and scal[0] == k1 and scal[1] == k2 and int(scal[7]) == d["itronc"] and scal[4] == d["coef_tronca"]
        for i, n in enumerate(NAMES):
            assert np.array_equal(coef[i], d[n]), (itr, n)
        piz = k2 / k1
        ct = d["coef_tronca"]
        assert scal[2] == piz and scal[3] == piz * (1.0 - ct / 2.0) / (1.0 - piz * ct / 2.0)
    p22 = p11 * (1.0 - 0.05 * (1.0 - xmu))
    d = ac.ref_decompo(ref, 1, nbmu, xmu, xhr, os_nb, p11, p12, p22, p33)
    e, scal, coef, ph = host_model(host, nbmu, xmu, xhr, np.array([k1, k2, 1.0]), p11, p12, p33, 0, [0], [1.0], 1, os_nb, p22c=p22)
    for i, n in enumerate(NAMES):
        assert np.array_equal(coef[i], d[n]), n


def test_write_aerosols_file(tmp_path):
    """The result file: the fixed text of formats 39-49 (taken from the reference source when it is present), E13.5 / F9.5 / I4 /
    E15.8 fields, and the values SOS_PREPA_OS.F:669-693 reads back (list-directed after the ':' resp. four columns)."""
    import importlib
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    rng = np.random.default_rng(5)
    nb = 24
    co = [rng.normal(size=nb + 1) * 10.0 ** rng.integers(-12, 3, nb + 1) for _ in range(4)]
    co[0][0] = 0.0
    path = str(tmp_path / "Aerosols.txt")
    api.write_aerosols(path, nb, 2.7175527644569626, 2.519186564312819, 0.71234, 0.5343377513101681, 0.96123456, *co)
    lines = open(path).read().split("\n")
    assert lines[-1] == "" and len(lines) == 8 + nb + 1 + 1
    src = "/root/reference/src/SOS_AEROSOLS.F"
    if os.path.exists(src):
        holl = {}
        for m in re.finditer(r"^\s+(\d+) FORMAT\((\d+)h(.*)$", open(src, encoding="latin-1").read(), re.M):
            holl[int(m.group(1))] = m.group(3)[:int(m.group(2))]
        for i, lab in enumerate((40, 41, 42, 46, 47)):
            assert lines[i].startswith(holl[lab]) and len(holl[lab]) == 38
        assert lines[5] == holl[39] and lines[6] == holl[48] + "%4d" % nb and lines[7] == holl[49]
    assert lines[0].endswith("  0.27176E+01") and lines[1].endswith("  0.25192E+01") and lines[2].endswith("  0.71234E+00")
    assert lines[3].endswith(":  0.53434") and lines[4].endswith(":  0.96123")
    a = float(lines[3].split(":")[1]); piztr = float(lines[4].split(":")[1])
    assert abs(a - 0.53434) < 1e-12 and abs(piztr - 0.96123) < 1e-12
    for k in range(nb + 1):
        ln = lines[8 + k]
        assert len(ln) == 15 * 4 + 3
        vals = [float(x) for x in ln.split()]
        for v, c in zip(vals, co):
            assert v == float("%.7E" % c[k])                              # 8 significant digits, correctly rounded


# ------------------------------------------------------------------------------------------------ GPU
def _float_ties(a, b, floor=None):
    """REAL*4 arrays equal except where a last-bit difference of the FP64 value crossed a single-precision rounding boundary:
    returns (fraction identical, largest |a - b| in units of the last place of max(|b|, floor)).  floor: for the polarised phase
    functions, which are differences of nearly equal terms, 1e-9 of the intensity phase function (the FP64 noise level there)."""
    a, b = np.asarray(a, np.float32).ravel(), np.asarray(b, np.float32).ravel()
    ref = np.abs(b.astype(np.float64))
    if floor is not None:
        ref = np.maximum(ref, np.asarray(floor, np.float64).ravel())
    ulp = np.maximum(np.spacing(ref.astype(np.float32)).astype(np.float64), 1e-45)
    return (a == b).mean(), float((np.abs(a.astype(np.float64) - b.astype(np.float64)) / ulp).max())


def _check_mie_tables(g, r):
    worst, fracs = 0.0, []
    for k in ("rec", "imie", "qmie", "umie"):
        floor = None if k in ("rec", "imie") else 1e-2 * np.abs(r["imie"].astype(np.float64))
        frac, u = _float_ties(g[k], r[k], floor)
        worst = max(worst, u)
        fracs.append(frac)
        assert frac > 0.99 and u <= 4.0, (k, frac, u)
    return min(fracs), worst


@pytest.mark.gpu
@pytest.mark.parametrize("case", TABLES)
def test_gpu_mie_vs_reference(solver, ref, tmp_path, case):
    rn, in_, a0, af = case
    nbmu, xmu, xhr = ac.mie_angles(12, (0.0, 30.0))
    t0 = time.time()
    _, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, rn, in_, a0, af)
    t_ref = time.time() - t0
    g = solver.mie(nbmu, xmu, rn, in_, a0, af)
    ms = solver.last_kernel_ms
    assert g["rec"].shape == r["rec"].shape
    assert np.array_equal(g["rec"][:, 0], r["rec"][:, 0])                  # the size-parameter grid itself is exact
    frac, worst = _check_mie_tables(g, r)
    gerr = float(np.abs(g["g"] - r["g"]).max())
    assert gerr <= 1e-10, gerr
    print("[GPU Mie m=%g%+gi alphaf=%g] %d records x %d angles: REAL*4 values identical %.5f %%, rest within %.1f ulp(float); G within "
          "%.1e; k_mie %.2f ms, reference routine %.0f ms on one host core"
          % (rn, in_, af, r["g"].size, 2 * nbmu + 1, 100 * frac, worst, gerr, ms, 1e3 * t_ref))


@pytest.mark.gpu
def test_gpu_granu_decompo_vs_reference(solver, ref, tmp_path):
    """k_granu on the reference's own Mie tables and k_model on the reference's own phase functions (each stage alone)."""
    nbmu, xmu, xhr = ac.mie_angles(24, (0.0, 30.0))
    os_nb, wa = 48, 0.55
    for ci, c in enumerate(components_at(wa, [ac.COARSE, ac.FINE, ac.JUNGE])):
        rn, in_, a0, af, ig, v1, v2, v3, _ = c
        f, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, rn, in_, a0, af, "MIE%d.bin" % ci)
        e0, k1, k2, snr, p11, p12, p33 = ac.ref_granu(ref, f, ig, v1, v2, v3, wa, nbmu, xmu)
        e1, h1, h2, hs, q11, q12, q33 = solver.granu(nbmu, r, ig, v1, v2, v3, wa)
        assert e0 == 0 and e1 == 0
        assert np.allclose([h1, h2, hs], [k1, k2, snr], rtol=1e-12, atol=0)
        for a, b in ((q11, p11), (q12, p12), (q33, p33)):
            assert np.abs(a - b).max() <= 1e-12 * np.abs(p11).max()
        for itr in (0, 1):
            d = ac.ref_decompo(ref, itr, nbmu, xmu, xhr, os_nb, p11, p12, p11, p33)
            o = solver.decompo_legendre(itr, nbmu, xmu, xhr, os_nb, p11, p12, p11, p33)
            assert o["ier"] == 0 and o["itronc"] == d["itronc"]
            assert abs(o["coef_tronca"] - d["coef_tronca"]) <= 1e-12 and abs(o["z1"] - d["z1"]) <= 1e-12 * abs(d["z1"])
            scale = np.abs(d["beta11"]).max()
            for n in NAMES:
                assert np.abs(o[n] - d[n]).max() <= 1e-11 * scale, (ci, itr, n, np.abs(o[n] - d[n]).max())
            assert np.allclose(o["p11"], d["p11"], rtol=1e-11, atol=0) and np.array_equal(o["ttt"], d["ttt"])
    e, *_ = solver.granu(nbmu, {k: (v[:-40] if isinstance(v, np.ndarray) else v) for k, v in r.items()}, 1, 0.4, 0.6, 0.0, wa)
    assert e == -1                                                          # table shorter than the size distribution


def _reference_chain(ref, tmp, nbmu, xmu, xhr, comps, models, os_nb):
    """The reference's flow for a list of components and models: SOS_MIE file per distinct table -> SOS_GRANU -> mixture
    (numpy restatement of SOS_AEROSOLS.F:2085-2110) -> SOS_DECOMPO_LEGENDRE."""
    files, kc, pc = {}, [], []
    for c in comps:
        rn, in_, a0, af, ig, v1, v2, v3, wa = c
        key = (rn, in_, a0, af)
        if key not in files:
            files[key] = ac.ref_mie(ref, tmp, nbmu, xmu, xhr, rn, in_, a0, af, "T%d.bin" % len(files))[0]
        e, k1, k2, snr, p11, p12, p33 = ac.ref_granu(ref, files[key], ig, v1, v2, v3, wa, nbmu, xmu)
        assert e == 0
        kc.append([k1, k2, snr]); pc.append([p11, p12, p33])
    kc, pc = np.array(kc), np.array(pc)
    out = []
    for ncomp, ci, w, itr in models:
        if ncomp == 0:
            k1, k2, (p11, p12, p33) = kc[ci[0], 0], kc[ci[0], 1], pc[ci[0]]
        else:
            k1, k2, (p11, p12, p33) = mix_numpy(w, kc[ci, 0], kc[ci, 1], pc[ci, 0], pc[ci, 1], pc[ci, 2])
        d = ac.ref_decompo(ref, itr, nbmu, xmu, xhr, os_nb, p11, p12, p11, p33)
        d.update(kmat1=k1, kmat2=k2)
        out.append(d)
    return kc, pc, out, len(files)


@pytest.mark.gpu
def test_gpu_aerosols_chain_vs_reference(solver, ref, tmp_path):
    """sosgpu_aerosols (tables never leave the device) against the reference's flow for six wavelengths of a bimodal log-normal
    model (the two modes mixed by volume concentrations, SOS_AEROSOLS.F:1706-2123) plus a mono-modal Junge model; the result
    files compared line by line."""
    import importlib
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    nbmu, xmu, xhr = ac.mie_angles(24, (0.0,))
    os_nb = 48
    comps, models = [], []
    for wa in (0.443, 0.55, 0.67, 0.865, 1.24, 2.13):
        n0 = len(comps)
        comps += components_at(wa, [ac.COARSE, ac.FINE])
        cv = np.array([0.6, 0.4])
        models.append((2, [n0, n0 + 1], list(cv / (cv[0] + cv[1])), 1))
    comps += components_at(0.55, [ac.JUNGE])
    models.append((0, [len(comps) - 1], [1.0], 1))
    t0 = time.time()
    kc, pc, refs, ntab = _reference_chain(ref, str(tmp_path), nbmu, xmu, xhr, comps, models, os_nb)
    t_ref = time.time() - t0
    t0 = time.time()
    o = solver.aerosols(nbmu, xmu, xhr, comps, models, os_nb)
    t_gpu = time.time() - t0
    assert (o["comp_ier"] == 0).all() and (o["model_ier"] == 0).all()
    assert np.allclose(o["comp_k"], kc, rtol=2e-7, atol=0)
    for c in range(len(comps)):
        assert np.abs(o["comp_phase"][c] - pc[c]).max() <= 2e-7 * np.abs(pc[c, 0]).max(), c
    same_lines = total_lines = 0
    for m, d in enumerate(refs):
        s = o["scal"][m]
        assert int(s[7]) == d["itronc"], m
        assert np.isclose(s[0], d["kmat1"], rtol=2e-7) and np.isclose(s[1], d["kmat2"], rtol=2e-7) and abs(s[4] - d["coef_tronca"]) < 2e-7
        scale = np.abs(d["beta11"]).max()
        for i, n in enumerate(NAMES):
            assert np.abs(o["coef"][m, i] - d[n]).max() <= 2e-7 * scale, (m, n)
        piz = d["kmat2"] / d["kmat1"]
        ct = d["coef_tronca"]
        fa, fb = str(tmp_path / ("gpu%d.txt" % m)), str(tmp_path / ("ref%d.txt" % m))
        api.write_aerosols(fa, os_nb, s[0], s[1], s[5], s[4], s[3], o["coef"][m, 0], o["coef"][m, 1], o["coef"][m, 2], o["coef"][m, 3])
        api.write_aerosols(fb, os_nb, d["kmat1"], d["kmat2"], ct / 2.0 + (1.0 - ct / 2.0) * d["beta11"][1] / 3.0, ct,
                           piz * (1.0 - ct / 2.0) / (1.0 - piz * ct / 2.0), d["alp"], d["beta11"], d["gamma12"], d["zeta"])
        la, lb = open(fa).read().split("\n"), open(fb).read().split("\n")
        assert la[:8] == lb[:8], m                                          # cross sections, albedo, truncation: same decimals
        same_lines += sum(x == y for x, y in zip(la, lb)); total_lines += len(lb)
    assert same_lines >= 0.98 * total_lines
    print("[GPU aerosol chain] %d models / %d components / %d Mie tables: truncation decisions identical, coefficients within 2e-7 "
          "of beta11's scale (REAL*4 Mie records), result-file lines identical %d / %d; device time of the chain %.1f ms (call %.0f "
          "ms), reference flow %.1f s on one host core" % (len(models), len(comps), ntab, same_lines, total_lines,
                                                           solver.last_kernel_ms, 1e3 * t_gpu, t_ref))


@pytest.mark.gpu
def test_gpu_aerosol_shims(ref, tmp_path):
    """The gfortran-ABI symbols sos_mie_ / sos_granu_ / sos_decompo_legendre_ of libsosgpu.so with the reference's strides and
    files, next to the same calls into the reference library."""
    import importlib
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    nbmu, xmu, xhr = ac.mie_angles(10, (0.0,))
    rn, in_, af, wa = 1.5, -0.01, 100.0, 0.55
    fr, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, rn, in_, 0.0001, af, "REF.bin")
    fg, g = ac.ref_mie(lib, str(tmp_path), nbmu, xmu, xhr, rn, in_, 0.0001, af, "GPU.bin")
    assert os.path.getsize(fr) == os.path.getsize(fg)
    assert (g["rn"], g["in_"], g["alphaf"]) == (r["rn"], r["in_"], r["alphaf"])
    _check_mie_tables(g, r)
    a = ac.ref_granu(ref, fr, 1, 0.3, 0.5, 0.0, wa, nbmu, xmu)
    b = ac.ref_granu(lib, fr, 1, 0.3, 0.5, 0.0, wa, nbmu, xmu)          # the GPU symbol reading the reference's file
    assert a[0] == 0 and b[0] == 0 and np.allclose(a[1:4], b[1:4], rtol=1e-12)
    for x, y in zip(a[4:], b[4:]):
        assert np.abs(x - y).max() <= 1e-12 * np.abs(a[4]).max()
    da = ac.ref_decompo(ref, 1, nbmu, xmu, xhr, 20, a[4], a[5], a[4], a[6])
    db = ac.ref_decompo(lib, 1, nbmu, xmu, xhr, 20, a[4], a[5], a[4], a[6])
    assert db["ier"] == 0 and da["itronc"] == db["itronc"] and abs(da["coef_tronca"] - db["coef_tronca"]) < 1e-12
    for n in NAMES:
        assert np.abs(da[n] - db[n]).max() <= 1e-11 * np.abs(da["beta11"]).max(), n


@pytest.mark.gpu
def test_gpu_aerosols_sweep_timing(solver):
    """A hyperspectral-sweep-sized call (BASELINE configs[3]: bimodal log-normal aerosol, refractive index varying with the
    wavelength, so every wavelength needs its own two Mie tables): timing only, plus sanity of the outputs."""
    nbmu, xmu, xhr = ac.mie_angles(40)
    os_nb = 80
    comps, models = [], []
    for i, wa in enumerate(np.linspace(0.4, 2.5, 64)):
        n0 = len(comps)
        coarse = (1.45 - 0.01 * i / 64, -0.004, 1, 0.40, 0.60, -999.0)
        fine = (1.42 + 0.01 * i / 64, -0.008, 1, 0.08, 0.45, -999.0)
        comps += components_at(float(wa), [coarse, fine])
        models.append((2, [n0, n0 + 1], [0.5, 0.5], 1))
    t0 = time.time()
    o = solver.aerosols(nbmu, xmu, xhr, comps, models, os_nb, want_phase=False)
    dt = time.time() - t0
    assert (o["comp_ier"] == 0).all() and (o["model_ier"] == 0).all()
    assert np.allclose(o["coef"][:, 1, 0], 1.0, rtol=0, atol=1e-12)          # beta11(0) = 1 after the normalisation
    assert ((o["scal"][:, 2] > 0.8) & (o["scal"][:, 2] < 1.0)).all()          # single-scattering albedo
    nrec = sum(solver.mie_count(c[2], c[3]) for c in comps)
    print("[GPU aerosol sweep] 64 wavelengths x 2 modes = 128 Mie tables (%d records x 81 angles), expansions to order 80: device "
          "time %.1f ms, call %.0f ms" % (nrec, solver.last_kernel_ms, 1e3 * dt))


# ------------------------------------------------------------------------------------------------ host front end (aerosols.py)
def _aer():
    import importlib
    return importlib.import_module("radiativetransfer-sos_b200.aerosols")


def test_wmo_params_vs_reference(ref, tmp_path):
    """aerosols.wmo_params against SOS_INIT_PARAMWMO of the reference library: a generated file in the reference's fixed-column
    layout at several wavelengths, and the reference's own WMO table where the reference tree is present."""
    aer = _aer()
    files = [ac.write_wmo_file(str(tmp_path / "Data_WMO_test"))]
    real = "/root/reference/fic/Data_WMO_cor_2015_12_16"
    if os.path.exists(real):
        files.append(real)
    if not hasattr(ref, "sos_init_paramwmo_"):
        pytest.skip("SOS_INIT_PARAMWMO not in the reference library")
    for path in files:
        for wa in (0.4, 0.443, 0.55, 0.865, 0.91, 1.6, 2.13):
            e, v1, v2, mr, mi, vol = ac.ref_wmo_params(ref, path, wa)
            g1, g2, gr, gi, gv = aer.wmo_params(path, wa)
            assert e == 0
            assert np.array_equal(v1, g1) and np.array_equal(v2, g2) and np.array_equal(vol, gv), (path, wa)
            assert np.array_equal(mr, gr) and np.array_equal(mi, gi), (path, wa, mr, gr, mi, gi)


def test_sf_params_vs_reference(ref, tmp_path):
    """aerosols.sf_params against SOS_INIT_PARAMSF of the reference library: generated data files in the reference's layout, and
    the reference's own Shettle & Fenn tables where the reference tree is present; humidities on and between the table's rows."""
    aer = _aer()
    if not hasattr(ref, "sos_init_paramsf_"):
        pytest.skip("SOS_INIT_PARAMSF not in the reference library")
    dirs = [ac.write_sf_files(str(tmp_path / "fic"))]
    if os.path.exists("/root/reference/fic/Data_SF_cor_2015_12_16"):
        dirs.append("/root/reference/fic")
    for d in dirs:
        for wa in (0.4, 0.55, 0.865, 1.6, 2.13):
            for rh in (0.0, 30.0, 50.0, 70.0, 75.5, 90.0, 99.0):
                e, v1, v2, mr, mi = ac.ref_sf_params(ref, d, wa, rh)
                g1, g2, gr, gi = aer.sf_params(d, wa, rh)
                assert e == 0
                assert np.array_equal(v1, g1) and np.array_equal(v2, g2) and np.array_equal(mr, gr) and np.array_equal(mi, gi), (d, wa, rh)
    # the maritime model: small rural + oceanic, number densities 0.99 / 0.01 as REAL*4 literals, ALPHAF 70 resp. from the radius
    p = aer.plan(aer.ShettleFenn(dirs[0], 3, 70.0), [0.55])
    v1, v2, mr, mi = aer.sf_params(dirs[0], 0.55, 70.0)
    assert p.models == [(2, [0, 1], [float(np.float32(0.99)), float(np.float32(0.01))], 1)]
    assert p.components[0] == (mr[0], mi[0], 0.0001, 70.0, 1, v1[0], v2[0], -999.0, 0.55)
    assert p.components[1][3] == ac.alphaf_for(ac.lnd_rmax(v1[4], v2[4]), 0.55) and p.components[1][:2] == (mr[4], mi[4])
    assert [c[3] for c in aer.plan(aer.ShettleFenn(dirs[0], 2, 0.0), [0.865]).components][0] == 90.0      # urban: small urban first


def test_aerosol_plan_host_logic(tmp_path):
    """The component / model lists aerosols.plan builds: Mie table ranges (ALPHAF), index rounding, mixture weights."""
    aer = _aer()
    # mono-modal: the table is sized for CTE_WAMIN whatever the wavelength (SOS_AEROSOLS.F:1183)
    p = aer.plan(aer.MonoModal(1.4449, -0.0040049, 1, 0.40, 0.60), [0.55, 0.865])
    assert len(p.components) == 2 and p.models == [(0, [0], [1.0], 1), (0, [1], [1.0], 1)]
    assert p.components[0][:2] == (1.445, -0.004) and p.components[0][3] == p.components[1][3] == 200.0
    assert aer.round_index(1.4445, -0.004005) == (1.445, -0.00401)            # DNINT: halves away from zero
    # bimodal: one table per mode and wavelength, ALPHAF from the wavelength; CVI normalised
    b = aer.BimodalLnd(1.45, -0.004, 0.40, 0.60, 1.42, -0.008, 0.08, 0.45, cv_coarse=0.3, cv_fine=0.9)
    p = aer.plan(b, [0.443, 2.13], itronc=0)
    assert [c[3] for c in p.components] == [200.0, 100.0, 100.0, 100.0]
    assert p.models[1] == (2, [2, 3], [0.3 / 1.2, 0.9 / 1.2], 0)
    for c, (rm, sg) in zip(p.components[:2], ((0.40, 0.60), (0.08, 0.45))):
        assert c[3] == ac.alphaf_for(ac.lnd_rmax(rm, sg), 0.443)
    # WMO maritime: water-soluble + oceanic, weights N(I) / NTOT with N = C / V
    f = ac.write_wmo_file(str(tmp_path / "wmo"))
    p = aer.plan(aer.Wmo(f, 2), [0.91])
    v1, v2, mr, mi, vol = aer.wmo_params(f, 0.91)
    n = [float(np.float32(0.05)) / vol[1], float(np.float32(0.95)) / vol[2]]    # C(2) = 0.05, C(3) = 0.95 are REAL*4 literals (:1345-1346)
    assert p.models == [(2, [0, 1], [n[0] / (0.0 / vol[0] + n[0] + n[1] + 0.0 / vol[3]), n[1] / (n[0] + n[1])], 1)]
    assert [c[3] for c in p.components] == [50.0, 800.0] and p.components[1][:2] == (mr[2], mi[2])
    with pytest.raises(ValueError):
        aer.plan(aer.MonoModal(1.4, 0.01, 1, 0.4, 0.6), [0.55])               # positive imaginary part


@pytest.mark.gpu
def test_gpu_aerosols_front_end_wmo_demo(solver, ref, tmp_path):
    """aerosols.run for the demo's aerosol description (-AER.Model 1 -AER.WMO.Model 2 -AER.Waref 0.550 -AER.AOTref 0.3 at
    0.910 microns, exe/runSOS-ABS_demo.ksh) on a generated WMO table, against the reference's flow: SOS_INIT_PARAMWMO ->
    SOS_MIE files -> SOS_GRANU -> mixture -> SOS_DECOMPO_LEGENDRE, and TA = KMAT1(WA) / KMAT1(WAREF) * AOT_REF."""
    aer = _aer()
    nbmu, xmu, xhr = ac.mie_angles(24, (0.0,))
    os_nb, wa, waref, aot = 48, 0.910, 0.550, 0.3
    f = ac.write_wmo_file(str(tmp_path / "wmo"))
    got = aer.run(solver, nbmu, xmu, xhr, os_nb, aer.Wmo(f, 2), [wa], waref=waref, aot_ref=aot, itronc=1)[0]
    k1 = {}
    for w in (wa, waref):
        e, v1, v2, mr, mi, vol = ac.ref_wmo_params(ref, f, w)
        comps = [(mr[i], mi[i], 0.0001, (4000.0, 50.0, 800.0, 10.0)[i], 1, v1[i], v2[i], -999.0, w) for i in (1, 2)]
        n = np.array([0.0 / vol[0], np.float64(np.float32(0.05)) / vol[1], np.float64(np.float32(0.95)) / vol[2], 0.0 / vol[3]])   # C(2) = 0.05, C(3) = 0.95: REAL*4 literals
        ntot = 0.0
        for x in n:
            ntot = ntot + x
        kc, pc, refs, _ = _reference_chain(ref, str(tmp_path), nbmu, xmu, xhr, comps, [(2, [0, 1], [n[1] / ntot, n[2] / ntot], 1)], os_nb)
        k1[w] = refs[0]
    d = k1[wa]
    assert got.itronc == d["itronc"] and np.isclose(got.kmat1, d["kmat1"], rtol=2e-7) and abs(got.coef_tronca - d["coef_tronca"]) < 2e-7
    assert np.isclose(got.ta, d["kmat1"] / k1[waref]["kmat1"] * aot, rtol=4e-7)
    scale = np.abs(d["beta11"]).max()
    for a, n in ((got.alpha, "alp"), (got.beta, "beta11"), (got.gamma, "gamma12"), (got.zeta, "zeta")):
        assert np.abs(a - d[n]).max() <= 2e-7 * scale, n
    print("[GPU aerosols front end] WMO maritime at 0.910 um: TA = %.6f (reference flow %.6f), truncation coefficient %.6f (%.6f)"
          % (got.ta, d["kmat1"] / k1[waref]["kmat1"] * aot, got.coef_tronca, d["coef_tronca"]))


def test_mie_file_cache_interchange(host, ref, tmp_path):
    """MIE file names against SOS_NOM_FICMIE of the reference library; a table written by aerosols.write_mie_file is byte for byte the
    file SOS_MIE writes, and SOS_GRANU of the reference reads it."""
    aer = _aer()
    if hasattr(ref, "sos_nom_ficmie_"):
        for nb, rn, in_, a0, af in ((40, 1.45, -0.004, 0.0001, 200.0), (24, 1.333, 0.0, 0.0001, 100.0), (8, 1.75, -0.44, 0.0001, 4900.0),
                                    (83, 1.501, -0.00563, 0.0001, 70.0), (40, 1.386, -0.00001, 0.0001, 12300.0)):
            out = C.create_string_buffer(b" " * 150, 150)
            ref.sos_nom_ficmie_(ac._ip(nb), ac._fs("NO_USER_ANGLES"), ac._dp(rn), ac._dp(in_), ac._dp(a0), ac._dp(af), out, ac._L, C.c_size_t(150))
            assert out.raw.decode().strip() == aer.mie_file_name(nb, rn, in_, a0, af), (nb, rn, in_, af)
    nbmu, xmu, xhr = ac.mie_angles(8, (0.0,))
    rn, in_, a0, af = 1.45, -0.004, 0.0001, 60.0
    f_ref, r = ac.ref_mie(ref, str(tmp_path), nbmu, xmu, xhr, rn, in_, a0, af, "REF.bin")
    t = host_mie(host, nbmu, xmu, rn, in_, a0, af)
    f_mine = str(tmp_path / aer.mie_file_name(8, rn, in_, a0, af))
    aer.write_mie_file(f_mine, t, rn, in_, nbmu)
    assert open(f_mine, "rb").read() == open(f_ref, "rb").read()
    back = aer.read_mie_file(f_mine, nbmu)
    assert back["alphaf"] == af and np.array_equal(back["imie"], r["imie"]) and np.array_equal(back["g"], r["g"])
    a = ac.ref_granu(ref, f_mine, 1, 0.3, 0.5, 0.0, 0.55, nbmu, xmu)
    b = ac.ref_granu(ref, f_ref, 1, 0.3, 0.5, 0.0, 0.55, nbmu, xmu)
    assert a[0] == 0 and a[1:4] == b[1:4] and np.array_equal(a[4], b[4])
