"""The oracle against the REFERENCE ITSELF: oracle/build_ref.py translates the reference's Fortran sources of the hot
path mechanically to C (oracle/f77_to_c.py: Fortran typing rules, REAL*4 sub-expressions in single precision, libgcc's
integer powers, column-major arrays with the compile-time extents of SOS.h, gfortran's unformatted record I/O) and
compiles them into oracle/_ref/libsosref.so.  These tests run SOS_OS (driver + its 11 subroutines), SOS_TRPHI_OPTION /
SOS_TRPHI and SOS_GSF from that library -- i.e. the reference's own statements -- and require the hand-written
restatement in oracle/sos_oracle.c to reproduce them BIT FOR BIT: record counts, every Stokes value, fluxes, side
effects.  Skipped only when neither /root/reference nor a prebuilt oracle/_ref/libsosref.so is available."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MX, NTM, NBM = 80, 600, 200


@pytest.fixture(scope="module")
def ref():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sos_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    build_ref = importlib.util.module_from_spec(spec)             # by path: oracle/ must not shadow the oracle package
    spec.loader.exec_module(build_ref)
    lib = build_ref.build(verbose=False)
    if lib is None:
        pytest.skip("no /root/reference and no prebuilt oracle/_ref/libsosref.so")
    return C.CDLL(lib)


def _fs(s):
    return C.create_string_buffer(s.encode().ljust(500), 500)


_ip = lambda v: C.byref(C.c_int(v))
_dp = lambda v: C.byref(C.c_double(v))
_P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
_L = C.c_size_t(500)


def _pad(v, n):
    a = np.zeros(n)
    a[:len(v)] = v
    return a


def run_reference_sos_os(ref, fm, o, h, xd, yd, z, iborm, tmp, surf=None, ifresnel=0, zout=-1.0, ipolar=None):
    N, NT = o.nbmu, len(h) - 1
    rmu, ga = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    rmu[MX - N:MX + N + 1], ga[MX - N:MX + N + 1] = o.rmu, o.ga
    H, XD, YD, Z = (_pad(v, NTM + 1) for v in (h, xd, yd, z))
    al, be, gm, ze = (_pad(v, NBM + 1) for v in (o.alpha, o.beta, o.gamma, o.zeta))
    fos, fsurf = os.path.join(tmp, "REF_OS.bin"), os.path.join(tmp, "SURF.bin")
    if os.path.exists(fos):
        os.remove(fos)
    if surf is not None:
        fm.write_surface_bin(fsurf, surf)
    em, ep, ier = C.c_double(0), C.c_double(0), C.c_int(99)
    ref.sos_os_(_ip(N), _P(rmu), _P(ga), _ip(o.os_nb), _ip(NT), _fs(fsurf), _fs(fos), _ip(o.n0), _dp(o.tetas), _dp(o.rho),
                _ip(1 if surf is not None else 0), _ip(ifresnel), _dp(o.ind_surf), _P(H), _P(XD), _P(YD), _P(Z), _dp(o.ron),
                _P(al), _P(be), _P(gm), _P(ze), _dp(zout), _ip(o.igmax), _ip(iborm), _ip(o.ipolar if ipolar is None else ipolar),
                _ip(0), _ip(6), C.byref(em), C.byref(ep), C.byref(ier), _L, _L)
    rec = fm.read_result_bin(fos, N) if os.path.exists(fos) else np.zeros((0, 3, 2 * N + 1))
    return dict(ier=ier.value, rec=rec, emoins=em.value, eplus=ep.value, rmu0=rmu[MX], alpha=al, gamma=gm, zeta=ze)


CASES = {
    "lambert_black": dict(surface="lambert", rho=0.0),
    "lambert_0.3": dict(surface="lambert", rho=0.3),
    "brdf_matrix": dict(surface="brdf", rho=0.05),
    "flat_fresnel": dict(surface="lambert", rho=0.0, ifresnel=1),
    "unpolarized": dict(surface="lambert", rho=0.1, ipolar=0),
    "output_altitude": dict(surface="lambert", rho=0.1, zout=3.0),
    "rayleigh_only": dict(surface="lambert", rho=0.0, aerosol=0.0, iborm=2),
    "thick_aerosol": dict(surface="lambert", rho=0.2, aerosol=1.2, nb_gauss=16, os_nb=40),
    "n25": dict(surface="lambert", rho=0.1, nb_gauss=24, os_nb=32),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_sos_os_restatement_is_bit_identical_to_the_reference(pkg, orc, ref, case, tmp_path):
    syn, fm = pkg.synth, pkg.formats
    c = CASES[case]
    o = syn.make_optics(nb_gauss=c.get("nb_gauss", 12), tetas=35.0, os_nb=c.get("os_nb", 24), surface=c["surface"],
                        rho=c["rho"], zout=c.get("zout", -1.0), ipolar=c.get("ipolar", 1))
    z, h, xa, ym = syn.profile(0.08, 8.0, c.get("aerosol", 0.25), 2.0, 0.05)
    xd = np.array(xa) * o.piztr
    iborm = c.get("iborm", o.os_nb)
    ifr = c.get("ifresnel", 0)
    surf = o.surf if o.imat_surf == 1 else None
    al, gm, ze = o.alpha.copy(), o.gamma.copy(), o.zeta.copy()
    rmu = o.rmu.copy()
    r = orc.sos_os(o.nbmu, rmu, o.ga, o.os_nb, len(h) - 1, o.n0, o.tetas, o.rho, o.imat_surf, ifr, o.ind_surf, h, xd, ym, z,
                   o.ron, al, o.beta, gm, ze, o.zout, o.igmax, iborm, o.ipolar, surf)
    f = run_reference_sos_os(ref, fm, o, h, xd, ym, z, iborm, str(tmp_path), surf=surf, ifresnel=ifr, zout=o.zout)
    assert f["ier"] == r.ier == 0
    assert f["rec"].shape[0] == r.n_fourier                     # same number of Fourier orders written
    assert np.array_equal(f["rec"], r.rec)                      # every Q, U, I value: identical bits
    assert f["emoins"] == r.emoins and f["eplus"] == r.eplus
    assert f["rmu0"] == -o.rmu[o.nbmu + o.n0]                    # caller-visible side effect RMU(0) = mu_s
    if o.ipolar == 0:                                            # and the zeroing of alpha, gamma, zeta
        assert not f["alpha"].any() and not f["gamma"].any() and not f["zeta"].any()


def test_trphi_option_restatement_against_the_reference(pkg, orc, ref, tmp_path):
    """SOS_TRPHI_OPTION / SOS_TRPHI / SOS_GLITTE / SOS_ANGLE / SOS_REFLEX / SOS_MATRIC / SOS_POLAR from the reference."""
    syn, fm = pkg.synth, pkg.formats
    o = syn.make_optics(nb_gauss=12, tetas=35.0, os_nb=24, surface="glitter")
    z, h, xa, ym = syn.profile(0.05, 8.0, 0.2, 2.0, 0.0)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import oracle_term
    r = oracle_term(orc, o, syn.Term(0, 1.0, z, h, xa, ym))
    N = o.nbmu
    fos = str(tmp_path / "SOS_Result.bin")
    fm.write_result_bin(fos, r.rec)
    rmu, ga = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    rmu[MX - N:MX + N + 1], ga[MX - N:MX + N + 1] = o.rmu, o.ga
    for itrphi, phios, pas, igli, ifr in ((1, 0.0, 0, 1, 0), (1, 37.5, 0, 1, 0), (2, 0.0, 30, 1, 0), (2, 0.0, 45, 0, 1)):
        n0, pf0, th0, up0, dn0 = orc.trphi_option(r.rec, N, o.rmu, r.ttot_tronc, r.tauout, igli, o.n0, o.wind, o.ind_surf,
                                                  ifr, itrphi, phios, pas, 1)
        pf, th = np.zeros(361), np.zeros(MX + 1)
        tabs = [np.zeros((MX + 1, 361)) for _ in range(14)]
        ier = C.c_int(99)
        ref.sos_trphi_option_(_ip(N), _P(rmu), _P(ga), _fs(fos), _dp(r.ttot_tronc), _dp(r.tauout), _dp(-1.0), _ip(igli),
                              _ip(o.n0), _dp(o.wind), _dp(o.ind_surf), _ip(ifr), _ip(0), _dp(0.0), _dp(0.0), _dp(0.0), _ip(0),
                              _ip(0), _ip(0), _dp(0.0), _dp(0.0), _ip(0), _dp(0.0), _ip(itrphi), _dp(phios), _ip(pas), _ip(1),
                              _P(pf), _P(th), *[_P(t) for t in tabs], C.byref(ier), _L)
        assert ier.value == 0
        for tb in range(7):
            assert np.array_equal(tabs[tb][:N, :n0].T, up0[tb]), (itrphi, "up", tb)
            assert np.array_equal(tabs[7 + tb][:N, :n0].T, dn0[tb]), (itrphi, "down", tb)
        assert np.array_equal(th[:N], th0)


def test_gsf_restatement_against_the_reference(pkg, orc, ref, tmp_path):
    """SOS_GSF / SOS_CALCG: Fourier series of the Cox-Munk G function for every (theta1 >= theta2) pair: the data-
    dependent series lengths IL and every coefficient."""
    rmu, ga, n0, _ = pkg.synth.sos_angles(10, 35.0)
    N = (rmu.size - 1) // 2
    wind, os_nm = 5.0, 48
    sig = float(np.float32(0.003) + np.float32(0.00512) * np.float32(wind))
    r = np.zeros(2 * MX + 1)
    r[MX - N:MX + N + 1] = rmu
    fgsf = str(tmp_path / "RES_GSF")
    ier = C.c_int(99)
    ref.sos_gsf_(_ip(N), _P(r), _dp(sig), _ip(os_nm), _fs(fgsf), C.byref(ier), _L)
    assert ier.value == 0
    raw = open(fgsf, "rb").read()
    pos, npair = 0, 0
    while pos < len(raw):
        n = int(np.frombuffer(raw, dtype=np.int32, count=1, offset=pos)[0])
        i1, i2, il = np.frombuffer(raw, dtype=np.int32, count=3, offset=pos + 4)
        e = np.frombuffer(raw, dtype=np.float64, count=il + 1, offset=pos + 16)
        il0, e0 = orc.gsf_pair(rmu[N + i1], rmu[N + i2], sig, os_nm)
        assert il0 == il, (i1, i2)
        assert np.array_equal(e0[:il + 1], e), (i1, i2)
        pos += n + 8
        npair += 1
    assert npair == N * (N + 1) // 2


@pytest.mark.parametrize("nbg,wind", [(10, 2.0), (12, 7.5)])
def test_glitter_chain_restatement_is_bit_identical_to_the_reference(pkg, orc, ref, nbg, wind, tmp_path):
    """SOS_GLITTER from the reference: SOS_GSF -> SOS_MAT_FRESNEL (through its 4(E15.8) text file) -> SOS_MAT_REFLEXION +
    SOS_NOYAUX_FRESNEL -> SOS_MISE_FORMAT, intermediate files deleted at the end as the reference does.  The REAL*4
    records of the surface file must equal the oracle's, bit for bit."""
    syn, fm = pkg.synth, pkg.formats
    rmu, ga, n0, _ = syn.sos_angles(nbg, 35.0)
    N = (rmu.size - 1) // 2
    os_nb = os_ns = 2 * nbg
    os_nm = os_nb + os_ns
    r, g = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    r[MX - N:MX + N + 1], g[MX - N:MX + N + 1] = rmu, ga
    fgl = str(tmp_path / "GLITTER.bin")
    ier = C.c_int(99)
    ref.sos_glitter_(_ip(N), _P(r), _P(g), _dp(wind), _dp(1.34), _ip(os_nb), _ip(os_ns), _ip(os_nm), _fs(str(tmp_path / "GSF")),
                     _fs(str(tmp_path / "FRESNEL")), _fs(str(tmp_path / "MAT_REFLEX")), _fs(fgl), _ip(0), C.byref(ier),
                     _L, _L, _L, _L)
    assert ier.value == 0
    assert sorted(os.listdir(tmp_path)) == ["GLITTER.bin"]            # the three intermediate files are gone
    got = fm.read_surface_bin(fgl, N)
    want, _ = orc.glitter(N, rmu, ga, wind, 1.34, os_nb, os_ns, os_nm)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("a_trunc,trans", [(0.0, False), (0.35, True)])
def test_sos_restatement_is_bit_identical_to_the_reference(pkg, orc, ref, a_trunc, trans, tmp_path):
    """SOS (SOS.F:340-697) from the reference: formatted read of the profile file, truncation adaptation, SOS_OS, tau at
    the output level, and the 1+N black-surface IBORM=0 solves of -SOS.Trans."""
    syn, fm = pkg.synth, pkg.formats
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import oracle_term
    o = syn.make_optics(nb_gauss=10, tetas=40.0, os_nb=20, surface="lambert", rho=0.1, a_trunc=a_trunc, piztr=0.97)
    z, h, xa, ym = syn.profile(0.05, 8.0, 0.2, 2.0, 0.3)
    fprof = str(tmp_path / "PROFIL_TMP")
    fm.write_profile(fprof, z, h, xa, ym)
    z, h, xa, ym = fm.read_profile(fprof)
    NT, N = len(h) - 1, o.nbmu
    r = oracle_term(orc, o, syn.Term(0, 1.0, z, h, xa, ym), want_trans=trans)
    rmu, ga = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    rmu[MX - N:MX + N + 1], ga[MX - N:MX + N + 1] = o.rmu, o.ga
    al, be, gm, ze = (_pad(v, NBM + 1) for v in (o.alpha, o.beta, o.gamma, o.zeta))
    fos = str(tmp_path / "SOS_Result.bin")
    sc = [C.c_double(0) for _ in range(6)]               # ttot_tronc, ttot_vrai, tauout, tdifmus, emoins, eplus
    tdg = np.zeros(2 * MX + 1)
    ier = C.c_int(99)
    ref.sos_(_fs(fos), _fs(str(tmp_path / "Trans.txt") if trans else "NO_OUTPUT"), _fs(fprof), _ip(NT), _dp(-1.0), _ip(o.igmax),
             _ip(1), _dp(o.ron), _dp(o.ind_surf), _dp(o.rho), _ip(0), _ip(0), _fs("none"), _ip(o.n0), _dp(o.piz), _dp(o.piztr),
             _dp(o.a_trunc), _P(rmu), _P(ga), _dp(o.tetas), _ip(o.os_nb), _ip(N), _P(al), _P(be), _P(gm), _P(ze),
             C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]), C.byref(sc[3]), _P(tdg), C.byref(sc[4]), C.byref(sc[5]),
             _ip(0), _ip(6), C.byref(ier), _L, _L, _L, _L)
    assert ier.value == r.ier == 0
    rec = fm.read_result_bin(fos, N)
    assert rec.shape[0] == r.n_fourier and np.array_equal(rec, r.rec)
    assert sc[0].value == r.ttot_tronc and sc[1].value == r.ttot_vrai and sc[2].value == r.tauout
    assert sc[4].value == r.emoins and sc[5].value == r.eplus
    if trans:
        assert sc[3].value == r.tdifmus
        assert np.array_equal(tdg[MX + 1:MX + N + 1], r.tdifmug[N + 1:])


def test_aggregate_restatement_is_bit_identical_to_the_reference(pkg, orc, ref, tmp_path):
    """SOS_AGGREGATE from the reference (file read-modify-write through a temporary file and `mv`), three CKD terms whose
    Fourier series have different lengths, in an order that triggers the trailing all-zero record; records and the
    scalar accumulators (incl. the -log(sum a exp(-tau)) aggregation of the optical depths) must equal the oracle's."""
    syn, fm = pkg.synth, pkg.formats
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import oracle_term
    o = syn.make_optics(nb_gauss=8, tetas=35.0, os_nb=16, surface="lambert", rho=0.1)
    N = o.nbmu
    terms = [syn.Term(0, a, *syn.profile(0.05, 8.0, aer, 2.0, gas)) for a, aer, gas in ((0.2, 0.25, 0.02), (0.5, 0.0, 0.5), (0.3, 0.25, 3.0))]
    rs = [oracle_term(orc, o, t) for t in terms]
    assert len({r.n_fourier for r in rs}) > 1                      # ragged series lengths
    agg = orc.Aggregate(N, o.os_nb + 4)
    fres, ftmp, fagg = str(tmp_path / "SOS_Result.bin"), str(tmp_path / "OS_TMP.bin"), str(tmp_path / "AGG_TMP.bin")
    acc = [C.c_double(0) for _ in range(6)]                        # ttot_tronc, ttot_vrai, tauout, tdifmus, emoins, eplus
    tdg_tmp, tdg = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    for t, r in zip(terms, rs):
        fm.write_result_bin(ftmp, r.rec)
        ier = C.c_int(0)
        ref.sos_aggregate_(_ip(N), _dp(t.aik), _fs(ftmp), _dp(r.ttot_tronc), _dp(r.ttot_vrai), _dp(r.tauout), _dp(0.0), _P(tdg_tmp),
                           _dp(r.emoins), _dp(r.eplus), _fs(fagg), _fs(fres), C.byref(acc[0]), C.byref(acc[1]), C.byref(acc[2]),
                           C.byref(acc[3]), _P(tdg), C.byref(acc[4]), C.byref(acc[5]), C.byref(ier), _L, _L, _L)
        assert ier.value == 0
        agg.add(t.aik, r)
        res = fm.read_result_bin(fres, N)
        assert res.shape[0] == agg.nres                            # incl. the reference's trailing zero record
        assert np.array_equal(res, agg.res[:agg.nres])
    assert acc[0].value == agg.sc["ttot_tronc"] and acc[1].value == agg.sc["ttot_vrai"] and acc[2].value == agg.sc["tauout"]
    assert acc[4].value == agg.sc["emoins"] and acc[5].value == agg.sc["eplus"]


def test_committed_golden_vectors_are_what_the_reference_produces(pkg, ref, tmp_path):
    """tests/golden/oracle_small.npz (the fixtures the GPU parity tests use on a box that has no /root/reference) against
    the reference's own SOS run here from oracle/_ref: identical record counts and identical bits."""
    import importlib.util
    syn, fm = pkg.synth, pkg.formats
    spec = importlib.util.spec_from_file_location("sos_make_golden", os.path.join(ROOT, "tests", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_small.npz"))
    for name, (o, t) in mg.cases(pkg).items():
        N = o.nbmu
        fprof, fos, fsurf = (str(tmp_path / (name + s)) for s in ("_PROFIL_TMP", "_OS.bin", "_SURF.bin"))
        fm.write_profile(fprof, t.zprof, t.h, t.pcaer, t.pcmol)
        back = fm.read_profile(fprof)
        for a, b in zip(back, (t.zprof, t.h, t.pcaer, t.pcmol)):
            assert np.array_equal(a, np.asarray(b))               # the synthetic profiles are exact in the file format
        if o.imat_surf == 1:
            fm.write_surface_bin(fsurf, o.surf)
        rmu, ga = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
        rmu[MX - N:MX + N + 1], ga[MX - N:MX + N + 1] = o.rmu, o.ga
        al, be, gm, ze = (_pad(v, NBM + 1) for v in (o.alpha, o.beta, o.gamma, o.zeta))
        sc = [C.c_double(0) for _ in range(6)]
        tdg = np.zeros(2 * MX + 1)
        ier = C.c_int(99)
        ref.sos_(_fs(fos), _fs("NO_OUTPUT"), _fs(fprof), _ip(t.nt), _dp(o.zout), _ip(o.igmax), _ip(o.ipolar), _dp(o.ron),
                 _dp(o.ind_surf), _dp(o.rho), _ip(o.imat_surf), _ip(o.ifresnel), _fs(fsurf), _ip(o.n0), _dp(o.piz), _dp(o.piztr),
                 _dp(o.a_trunc), _P(rmu), _P(ga), _dp(o.tetas), _ip(o.os_nb), _ip(N), _P(al), _P(be), _P(gm), _P(ze),
                 C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]), C.byref(sc[3]), _P(tdg), C.byref(sc[4]), C.byref(sc[5]),
                 _ip(0), _ip(6), C.byref(ier), _L, _L, _L, _L)
        assert ier.value == 0, name
        rec = fm.read_result_bin(fos, N)
        assert rec.shape[0] == int(gold[name + "_nf"]), name
        assert np.array_equal(rec, gold[name + "_rec"]), name


def test_gauss_angles_of_the_input_generator_equal_the_reference(pkg, ref):
    """SOS_GAUSS (SOS_ANGLES.F:1022-1103) from the reference against synth.sos_gauss, which produces the Gauss angles and
    weights of every synthetic workload (tests and bench)."""
    for mm in (5, 13, 25, 41):
        mxa = 100                                                    # CTE_NBANGLES_MAX (SOS.h:555)
        amu, pmu = np.zeros(2 * mxa + 1), np.zeros(2 * mxa + 1)
        ref.sos_gauss_(_ip(mm), _P(amu), _P(pmu))
        mu, w = pkg.synth.sos_gauss(mm)
        got_mu, got_w = amu[mxa + 1:mxa + mm], pmu[mxa + 1:mxa + mm]
        assert np.array_equal(np.asarray(mu), got_mu) and np.array_equal(np.asarray(w), got_w), mm


def test_random_configurations_bit_identical(pkg, orc, ref, tmp_path):
    """Seeded random sweep (angles, expansion order, surface type, albedo, flat-sea flag, polarization switch, output
    altitude, solar angle, IGMAX, aerosol and gas loads): the restatement and the reference agree on every bit."""
    syn, fm = pkg.synth, pkg.formats
    rng = np.random.default_rng(12345)
    for it in range(12):
        nbg = int(rng.choice([4, 6, 8, 12, 16, 24]))
        os_nb = int(rng.choice([8, 12, 16, 24, 40]))
        surface = str(rng.choice(["lambert", "brdf", "lambert"]))
        rho = float(rng.choice([0.0, 0.05, 0.3, 0.8]))
        ifr = int(rng.random() < 0.25) if surface == "lambert" else 0
        ipolar = int(rng.random() < 0.8)
        zout = float(rng.choice([-1.0, -1.0, 0.0, 1.7, 10.0, 55.5]))
        tetas = float(rng.uniform(5, 80))
        igmax = int(rng.choice([100, 100, 3, 7]))
        o = syn.make_optics(nb_gauss=nbg, tetas=tetas, os_nb=os_nb, surface=surface, rho=rho, zout=zout, ipolar=ipolar,
                            igmax=igmax, seed=it)
        aer, gas = float(rng.choice([0.0, 0.05, 0.3, 1.0])), float(rng.choice([0.0, 0.1, 2.0, 10.0]))
        z, h, xa, ym = syn.profile(float(rng.uniform(0.01, 0.3)), 8.0, aer, 2.0, gas)
        xd = np.array(xa) * o.piztr
        iborm = 2 if aer == 0.0 else o.os_nb
        surf = o.surf if o.imat_surf == 1 else None
        r = orc.sos_os(o.nbmu, o.rmu.copy(), o.ga, o.os_nb, len(h) - 1, o.n0, o.tetas, o.rho, o.imat_surf, ifr, o.ind_surf, h, xd,
                       ym, z, o.ron, o.alpha.copy(), o.beta, o.gamma.copy(), o.zeta.copy(), o.zout, o.igmax, iborm, o.ipolar, surf)
        f = run_reference_sos_os(ref, fm, o, h, xd, ym, z, iborm, str(tmp_path), surf=surf, ifresnel=ifr, zout=o.zout)
        assert f["ier"] == r.ier, it
        assert f["rec"].shape[0] == r.n_fourier and np.array_equal(f["rec"], r.rec), it
        assert f["emoins"] == r.emoins and f["eplus"] == r.eplus, it
