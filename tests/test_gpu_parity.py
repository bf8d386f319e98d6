"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle on identical (format-rounded) inputs.

Bar (BASELINE.json north_star): numbers of Fourier orders and of scattering orders bit-identical;
Stokes I/Q/U within 1e-9 relative (1e-12 absolute floor) at every output angle.
"""
import numpy as np
import pytest

from util import assert_stokes_close, oracle_term

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("is_", [0, 1, 2, 3, 5, 16])
def test_noyaux(pkg, orc, solver, is_):
    """SOS_NOYAUX (SOS_OS.F:1857-2158): six phase-matrix kernels + l=2 rows."""
    o = pkg.synth.make_optics(nb_gauss=12, tetas=35.0, os_nb=24)
    N = o.nbmu
    rmu = o.rmu.copy()
    rmu[N] = -o.rmu[N + o.n0]
    ref = orc.noyaux(is_, rmu, o.os_nb, o.alpha, o.beta, o.gamma, o.zeta)
    got = solver.noyaux(is_, rmu, o.os_nb, o.alpha, o.beta, o.gamma, o.zeta)
    for k in ("xpl", "xrl", "xtl", "bp", "gr", "gt", "arr", "art", "att"):
        scale = np.abs(ref[k]).max() + 1e-300
        assert np.abs(got[k] - ref[k]).max() <= 1e-13 * scale, k


@pytest.mark.parametrize("is_,nbg,nt", [(0, 8, 9), (2, 8, 70), (3, 12, 130), (6, 40, 100)])
def test_order_step(pkg, orc, solver, is_, nbg, nt):
    """SOS_FSOURCE_ORDREIG + SOS_INTEGR_EPOPT, one fused DMMA step vs the scalar loops."""
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=nbg, tetas=35.0, os_nb=2 * nbg)
    N, W = o.nbmu, 2 * o.nbmu + 1
    rmu = o.rmu.copy()
    rmu[N] = -o.rmu[N + o.n0]
    rng = np.random.default_rng(7)
    h = np.concatenate([[0.0], np.cumsum(rng.uniform(0.001, 0.006, nt))])
    xdel, ydel = rng.uniform(0.1, 0.6, nt + 1), rng.uniform(0.1, 0.4, nt + 1)
    i1, q1, u1 = (rng.standard_normal((W, nt + 1)) for _ in range(3))
    ker = orc.noyaux(is_, rmu, o.os_nb, o.alpha, o.beta, o.gamma, o.zeta)
    aaa = o.ron / (2 - o.ron)
    aaa = (1 - aaa) / (1 + 2 * aaa)
    i2, q2, u2 = orc.fsource_ordreig(is_, nt, xdel, ydel, 1.0 if is_ == 0 else 0.0, 0.5 * aaa, -aaa * np.sqrt(1.5),
                                     3.0 * aaa, ker, i1, q1, u1, o.ga)
    zi, zq, zu = (np.zeros((W, nt + 1)) for _ in range(3))
    ri, rq, ru = orc.integr_epopt(rmu, nt, h, i2, q2, u2, zi, zq, zu)
    gi, gq, gu, ji, jq, ju = solver.order_step(is_, rmu, o.ga, o.os_nb, o.alpha, o.beta, o.gamma, o.zeta, o.ron, 1,
                                               nt, h, xdel, ydel, i1, q1, u1)
    keep = np.arange(W) != N
    for g, r, nm in ((ji, i2, "I2"), (jq, q2, "Q2"), (ju, u2, "U2"), (gi, ri, "I1"), (gq, rq, "Q1"), (gu, ru, "U1")):
        scale = np.abs(r[keep]).max()
        assert np.abs(g[keep] - r[keep]).max() <= 1e-11 * scale, nm      # random fields: heavy cancellation


def _check_terms(pkg, orc, solver, wl, ids=None):
    tr, gr = solver.solve(wl, ids)
    ids = range(len(wl.terms)) if ids is None else ids
    for n, i in enumerate(ids):
        t = wl.terms[i]
        o = wl.optics[t.optics]
        r = oracle_term(orc, o, t)
        W = 2 * o.nbmu + 1
        assert tr.ier[n] == r.ier
        assert tr.n_fourier[n] == r.n_fourier, ("n_fourier", i, tr.n_fourier[n], r.n_fourier)
        assert np.array_equal(tr.n_scatter[n, :r.n_fourier], r.n_scatter), ("n_scatter", i, tr.n_scatter[n, :r.n_fourier], r.n_scatter)
        assert np.array_equal(tr.stop_reason[n, :r.n_fourier], r.stop_reason)
        assert_stokes_close(tr.rec[n, :r.n_fourier, :, :W], r.rec, "term %d records" % i)
        assert_stokes_close(tr.emoins[n], r.emoins, "emoins")
        assert_stokes_close(tr.eplus[n], r.eplus, "eplus")
        for k in ("ttot_tronc", "ttot_vrai", "tauout"):
            assert getattr(tr, k)[n] == getattr(r, k), k
    return tr, gr


def test_solve_golden_cases(pkg, orc, solver):
    """The four small cases of tests/golden (Lambert, BRDF/BPDF matrix, flat sea, output altitude)."""
    import make_golden
    syn = pkg.synth
    for name, (o, t) in make_golden.cases(pkg).items():
        wl = syn.Workload(name, [o], [t])
        _check_terms(pkg, orc, solver, wl)


def test_solve_rayleigh_only(pkg, orc, solver):
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=12, tetas=35.0, os_nb=24, surface="lambert", rho=0.2, a_trunc=0.0, piztr=1.0)
    wl = syn.Workload("ray", [o], [syn.Term(0, 1.0, *syn.profile(0.3, 8.0, 0.0, 2.0, 0.0))])
    tr, _ = _check_terms(pkg, orc, solver, wl)
    assert tr.n_fourier[0] == 3


def test_solve_unpolarized(pkg, orc, solver):
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=8, tetas=35.0, os_nb=16, surface="brdf", rho=0.05, ipolar=0)
    wl = syn.Workload("unpol", [o], [syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.2, 2.0, 0.0))])
    tr, _ = _check_terms(pkg, orc, solver, wl)
    assert np.all(tr.rec[0, :, 0] == 0.0) and np.all(tr.rec[0, :, 1] == 0.0)


def test_solve_demo_config_and_aggregate(pkg, orc, solver):
    """configs[0]-like (N=41, OS_NB=80, 5 CKD terms, ragged NT) + the SOS_AGGREGATE chain."""
    syn = pkg.synth
    wl = syn.config_demo(nterm=5, nb_gauss=40, os_nb=80, surface="brdf")
    tr, gr = _check_terms(pkg, orc, solver, wl)
    o = wl.optics[0]
    agg = orc.Aggregate(o.nbmu, o.os_nb + 1)
    for i, t in enumerate(wl.terms):
        agg.add(t.aik, oracle_term(orc, o, t))
    nmax = int(tr.n_fourier.max())
    assert gr.n_rec[0] == nmax
    assert_stokes_close(gr.rec[0, :nmax], agg.res[:nmax], "aggregated records")
    assert np.all(agg.res[nmax:agg.nres] == 0.0)      # the reference's padding records carry no signal
    for k in ("emoins", "eplus", "ttot_tronc", "ttot_vrai", "tauout"):
        assert_stokes_close(getattr(gr, k)[0], agg.sc[k], k)


def test_solve_mixed_batch_multiwave(pkg, orc, solver):
    """Several wavelengths with different optics / N in one batch, forced through 3-order waves."""
    syn = pkg.synth
    wl = syn.Workload("mixed")
    wl.optics.append(syn.make_optics(nb_gauss=8, tetas=35.0, os_nb=16, surface="lambert", rho=0.1))
    wl.optics.append(syn.make_optics(nb_gauss=12, tetas=60.0, os_nb=24, surface="brdf", rho=0.0, seed=3))
    wl.optics.append(syn.make_optics(nb_gauss=6, tetas=10.0, os_nb=12, surface="fresnel"))
    rng = np.random.default_rng(11)
    for p in range(3):
        for k in range(3):
            tg = float(np.exp(rng.uniform(np.log(1e-3), np.log(5.0))))
            wl.terms.append(syn.Term(p, 1.0 / 3, *syn.profile(0.05, 8.0, 0.15 + 0.1 * p, 2.0, tg)))
    solver.set_options(0, 3)
    try:
        _check_terms(pkg, orc, solver, wl)
    finally:
        solver.set_options(0, 0)


def test_solve_long_profile(pkg, orc, solver):
    """NT at the cap (600 layers) and a thick absorbing term."""
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=8, tetas=35.0, os_nb=16, surface="lambert", rho=0.3)
    wl = syn.Workload("thick", [o], [syn.Term(0, 1.0, *syn.profile(0.1, 8.0, 0.5, 2.0, 20.0)),
                                     syn.Term(0, 1.0, *syn.profile(0.1, 8.0, 0.5, 2.0, 1.2))])
    assert wl.terms[0].nt == 600
    _check_terms(pkg, orc, solver, wl)


def test_trphi_option(pkg, orc, solver):
    """SOS_TRPHI_OPTION (SOS_TRPHI.F:285-636): azimuth synthesis + sun-glint direct term, both view modes."""
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=12, tetas=35.0, os_nb=24, surface="glitter")
    t = syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.2, 2.0, 0.0))
    r = oracle_term(orc, o, t)
    for itrphi, phios, pas in ((1, 0.0, 0), (1, 37.5, 0), (2, 0.0, 30)):
        n0, pf0, th0, up0, dn0 = orc.trphi_option(r.rec, o.nbmu, o.rmu, r.ttot_tronc, r.tauout, 1, o.n0, o.wind,
                                                  o.ind_surf, 0, itrphi, phios, pas, 1)
        n1, pf1, th1, up1, dn1 = solver.trphi_option(r.rec, o.nbmu, o.rmu, r.ttot_tronc, r.tauout, 1, o.n0, o.wind,
                                                     o.ind_surf, 0, itrphi, phios, pas, 1)
        assert n0 == n1
        assert np.allclose(up1[0], up0[0], rtol=0, atol=1e-5) and np.allclose(dn1[0], dn0[0], rtol=0, atol=1e-5)   # degrees; acos is ill-conditioned at 0/180
        for tb in (1, 2, 3):
            assert_stokes_close(up1[tb], up0[tb], "up table %d" % tb)
            assert_stokes_close(dn1[tb], dn0[tb], "down table %d" % tb)
        assert np.allclose(up1[4:], up0[4:], rtol=1e-7, atol=1e-9) and np.allclose(dn1[4:], dn0[4:], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("nbg,wind", [(12, 2.0), (24, 7.5)])
def test_glitter(pkg, orc, solver, nbg, wind):
    """SOS_GLITTER pipeline (SOS_GSF, SOS_MAT_FRESNEL, SOS_MAT_REFLEXION, SOS_NOYAUX_FRESNEL, SOS_MISE_FORMAT):
    REAL*4 surface-file records and the integer series lengths IL of every angle pair."""
    rmu, ga, n0, _ = pkg.synth.sos_angles(nbg, 35.0)
    N = (rmu.size - 1) // 2
    os_ns, os_nb = 2 * nbg, 2 * nbg
    os_nm = os_nb + os_ns
    ref, il0 = orc.glitter(N, rmu, ga, wind, 1.34, os_nb, os_ns, os_nm)
    got, il1 = solver.glitter(N, rmu, ga, wind, 1.34, os_nb, os_ns, os_nm)
    assert np.array_equal(il0, il1)                     # data-dependent series cut: bit-identical counts
    exact = np.mean(got.view(np.uint32) == ref.view(np.uint32))
    assert exact > 0.999, exact                         # REAL*4 storage: bit-identical except rare rounding ties
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-7 * scale


def test_solve_with_glitter_surface(pkg, orc, solver):
    """configs[1]-like: rough sea; the surface matrix comes from the glitter pipeline and feeds the solver."""
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=12, tetas=35.0, os_nb=24, surface="glitter", rho=0.0)
    surf, _ = solver.glitter(o.nbmu, o.rmu, o.ga, o.wind, o.ind_surf, o.os_nb, 24, 48)
    ref, _ = orc.glitter(o.nbmu, o.rmu, o.ga, o.wind, o.ind_surf, o.os_nb, 24, 48)
    o.surf = ref                                        # same REAL*4 file for both sides
    assert np.mean(surf.view(np.uint32) == ref.view(np.uint32)) > 0.999
    wl = syn.Workload("sea", [o], [syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.2, 2.0, 0.0))])
    tr, _ = _check_terms(pkg, orc, solver, wl)
    assert tr.n_fourier[0] > 3


def test_gfortran_abi_sos_os_and_aggregate(pkg, orc, solver, tmp_path):
    """Drop-in symbols sos_os_ / sos_aggregate_ (SOS_OS.F:303-308, SOS_AGGREGATE.F:172-178): F77 by-reference
    arguments, SOS.h fixed strides, hidden CHARACTER*500 lengths, surface file in, SOS_Result.bin records out."""
    import ctypes as C
    import importlib
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    fm, syn = pkg.formats, pkg.synth
    lib = api.load_library()
    o = syn.make_optics(nb_gauss=8, tetas=50.0, os_nb=16, surface="brdf", rho=0.05)
    t = syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.1, 2.0, 0.5))
    r = oracle_term(orc, o, t)                       # SOS level; hand SOS_OS its truncation-adapted profile
    N, NT = o.nbmu, t.nt
    MX, NTM, NBM = 80, 600, 200

    def strided(v, cap, off):
        a = np.zeros(cap)
        a[off:off + len(v)] = v
        return a
    rmu = np.zeros(2 * MX + 1); ga = np.zeros(2 * MX + 1)
    rmu[MX - N:MX + N + 1] = o.rmu; ga[MX - N:MX + N + 1] = o.ga
    h, xd, yd, zp = (strided(v, NTM + 1, 0) for v in (r.h, r.xdel, r.ydel, t.zprof))
    al, be, gm, ze = (strided(v, NBM + 1, 0) for v in (o.alpha, o.beta, o.gamma, o.zeta))
    fsurf, fos = str(tmp_path / "SURF.bin"), str(tmp_path / "FICOS_TMP")
    fm.write_surface_bin(fsurf, o.surf)

    def fstr(s):
        return C.create_string_buffer(s.encode().ljust(500), 500)
    ip = lambda v: C.byref(C.c_int(v))
    dp = lambda v: C.byref(C.c_double(v))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    em, ep, ier = C.c_double(0), C.c_double(0), C.c_int(99)
    iborm = o.os_nb
    lib.sos_os_(ip(N), P(rmu), P(ga), ip(o.os_nb), ip(NT), fstr(fsurf), fstr(fos), ip(o.n0), dp(o.tetas), dp(o.rho),
                ip(1), ip(0), dp(o.ind_surf), P(h), P(xd), P(yd), P(zp), dp(o.ron), P(al), P(be), P(gm), P(ze),
                dp(-1.0), ip(o.igmax), ip(iborm), ip(1), ip(0), ip(6), C.byref(em), C.byref(ep), C.byref(ier),
                C.c_size_t(500), C.c_size_t(500))
    assert ier.value == 0
    assert rmu[MX] == -o.rmu[N + o.n0]               # caller-visible side effect RMU(0) = mu_s (SOS_OS.F:715)
    rec = fm.read_result_bin(fos, N)
    assert rec.shape[0] == r.n_fourier
    assert_stokes_close(rec, r.rec, "sos_os_ records")
    assert_stokes_close(em.value, r.emoins, "EMOINS")
    # aggregate twice into a fresh result file (accumulators start at 0, SOS_PROC.F:1292-1302)
    fres, fagg = str(tmp_path / "SOS_Result.bin"), str(tmp_path / "AGG_TMP")
    tdg_tmp, tdg = np.zeros(2 * MX + 1), np.zeros(2 * MX + 1)
    acc = [C.c_double(0) for _ in range(6)]          # ttot_tronc, ttot_vrai, tauout, tdifmus, emoins, eplus
    agg = orc.Aggregate(N, o.os_nb + 1)
    for aik in (0.25, 0.75):
        ier2 = C.c_int(0)
        lib.sos_aggregate_(ip(N), dp(aik), fstr(fos), dp(r.ttot_tronc), dp(r.ttot_vrai), dp(r.tauout), dp(0.0),
                           P(tdg_tmp), dp(r.emoins), dp(r.eplus), fstr(fagg), fstr(fres),
                           C.byref(acc[0]), C.byref(acc[1]), C.byref(acc[2]), C.byref(acc[3]), P(tdg),
                           C.byref(acc[4]), C.byref(acc[5]), C.byref(ier2), C.c_size_t(500), C.c_size_t(500),
                           C.c_size_t(500))
        assert ier2.value == 0
        agg.add(aik, r)
    res = fm.read_result_bin(fres, N)
    assert res.shape[0] == agg.nres                  # incl. the reference's trailing zero record
    assert_stokes_close(res, agg.res[:agg.nres], "sos_aggregate_ records")
    assert_stokes_close(acc[0].value, agg.sc["ttot_tronc"], "TTOT_TRONC")
    assert_stokes_close(acc[5].value, agg.sc["eplus"], "EPLUS")


def test_gfortran_abi_sos_glitter_trphi(pkg, orc, solver, tmp_path):
    """Drop-in symbols sos_glitter_ -> sos_ (profile file in, truncation, -SOS.Trans) -> sos_trphi_option_ /
    sos_trphi_ (SOS_GLITTER.F:229, SOS.F:340, SOS_TRPHI.F:285,749) chained through real files, as SOS_PROC does."""
    import ctypes as C
    import importlib
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    fm, syn = pkg.formats, pkg.synth
    lib = api.load_library()
    o = syn.make_optics(nb_gauss=10, tetas=40.0, os_nb=20, surface="glitter", rho=0.0)
    N, MX, NBM = o.nbmu, 80, 200
    os_ns, os_nm = 20, 40

    def fstr(s):
        return C.create_string_buffer(s.encode().ljust(500), 500)
    ip = lambda v: C.byref(C.c_int(v))
    dp = lambda v: C.byref(C.c_double(v))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    L500 = C.c_size_t(500)

    def strided(v, cap):
        a = np.zeros(cap)
        a[:len(v)] = v
        return a
    rmu = np.zeros(2 * MX + 1); ga = np.zeros(2 * MX + 1)
    rmu[MX - N:MX + N + 1] = o.rmu; ga[MX - N:MX + N + 1] = o.ga

    # --- SOS_GLITTER: surface file
    fgl = str(tmp_path / "GLITTER.bin")
    ier = C.c_int(99)
    lib.sos_glitter_(ip(N), P(rmu), P(ga), dp(o.wind), dp(o.ind_surf), ip(o.os_nb), ip(os_ns), ip(os_nm),
                     fstr(str(tmp_path / "RES_GSF")), fstr(str(tmp_path / "RES_FRESNEL")),
                     fstr(str(tmp_path / "RES_MAT_REFLEX")), fstr(fgl), ip(0), C.byref(ier), L500, L500, L500, L500)
    assert ier.value == 0
    ref_surf, _ = orc.glitter(N, o.rmu, o.ga, o.wind, o.ind_surf, o.os_nb, os_ns, os_nm)
    got_surf = fm.read_surface_bin(fgl, N)
    assert got_surf.shape == ref_surf.shape
    assert np.mean(got_surf.view(np.uint32) == ref_surf.view(np.uint32)) > 0.999
    lib.sos_glitter_(ip(N), P(rmu), P(ga), dp(o.wind), dp(o.ind_surf), ip(o.os_nb), ip(os_ns), ip(os_nm),
                     fstr("a"), fstr("b"), fstr("c"), fstr(fgl), ip(0), C.byref(ier), L500, L500, L500, L500)
    assert ier.value == -1                               # STATUS='NEW' (SOS_SURFACE.F:2360): refuses an existing file
    fm.write_surface_bin(fgl + ".ref", ref_surf)         # both sides read the same REAL*4 file from here on

    # --- SOS: profile file, truncation adaptation, transmissions
    z, h, xa, ym = syn.profile(0.05, 8.0, 0.2, 2.0, 0.3)
    fprof = str(tmp_path / "PROFIL_TMP")
    fm.write_profile(fprof, z, h, xa, ym)
    z, h, xa, ym = fm.read_profile(fprof)
    NT = len(h) - 1
    a_tr, piztr = 0.35, 0.97
    piz = piztr / (1 + 0.5 * a_tr * (piztr - 1))         # SOS_PREPA_OS.F:700
    o2 = syn.make_optics(nb_gauss=10, tetas=40.0, os_nb=20, surface="glitter", rho=0.0)
    o2.surf, o2.imat_surf, o2.a_trunc, o2.piztr = ref_surf, 1, a_tr, piztr
    assert o2.piz == piz
    r = oracle_term(orc, o2, syn.Term(0, 1.0, z, h, xa, ym), want_trans=True)
    al, be, gm, ze = (strided(v, NBM + 1) for v in (o.alpha, o.beta, o.gamma, o.zeta))
    fos = str(tmp_path / "SOS_Result.bin")
    sc = [C.c_double(0) for _ in range(6)]               # ttot_tronc, ttot_vrai, tauout, tdifmus, emoins, eplus
    tdg = np.zeros(2 * MX + 1)
    ier = C.c_int(99)
    lib.sos_(fstr(fos), fstr(str(tmp_path / "Trans.txt")), fstr(fprof), ip(NT), dp(-1.0), ip(o.igmax), ip(1), dp(o.ron),
             dp(o.ind_surf), dp(0.0), ip(1), ip(0), fstr(fgl + ".ref"), ip(o.n0), dp(piz), dp(piztr), dp(a_tr),
             P(rmu), P(ga), dp(o.tetas), ip(o.os_nb), ip(N), P(al), P(be), P(gm), P(ze),
             C.byref(sc[0]), C.byref(sc[1]), C.byref(sc[2]), C.byref(sc[3]), P(tdg), C.byref(sc[4]), C.byref(sc[5]),
             ip(0), ip(6), C.byref(ier), L500, L500, L500, L500)
    assert ier.value == 0
    rec = fm.read_result_bin(fos, N)
    assert rec.shape[0] == r.n_fourier
    assert_stokes_close(rec, r.rec, "sos_ records")
    for got, ref, what in zip(sc, (r.ttot_tronc, r.ttot_vrai, r.tauout, r.tdifmus, r.emoins, r.eplus),
                              ("TTOT_TRONC", "TTOT_VRAI", "TAUOUT", "TDIFMUS", "EMOINS", "EPLUS")):
        assert_stokes_close(got.value, ref, what)
    assert_stokes_close(tdg[MX + 1:MX + N + 1], r.tdifmug[N + 1:], "TDIFMUG")
    assert rmu[MX] == -o.rmu[2 * N]                       # RMU(0) after the last transmission solve (N0 = N)
    rmu[MX] = 0.0

    # --- SOS_TRPHI_OPTION / SOS_TRPHI on the result file
    for itrphi, phios, pas in ((1, 20.0, 0), (2, 0.0, 45)):
        n0, pf0, th0, up0, dn0 = orc.trphi_option(r.rec, N, o.rmu, r.ttot_tronc, r.tauout, 1, o.n0, o.wind,
                                                  o.ind_surf, 0, itrphi, phios, pas, 1)
        pf = np.zeros(361); th = np.zeros(MX + 1)
        tabs = [np.zeros((MX + 1, 361)) for _ in range(14)]   # Fortran (0:360,0:80): element (IP,JJ) at [JJ, IP]
        ier = C.c_int(99)
        lib.sos_trphi_option_(ip(N), P(rmu), P(ga), fstr(fos), dp(r.ttot_tronc), dp(r.tauout), dp(-1.0), ip(1), ip(o.n0),
                              dp(o.wind), dp(o.ind_surf), ip(0), ip(0), dp(0.0), dp(0.0), dp(0.0), ip(0), ip(0), ip(0),
                              dp(0.0), dp(0.0), ip(0), dp(0.0), ip(itrphi), dp(phios), ip(pas), ip(1), P(pf), P(th),
                              *[P(t) for t in tabs], C.byref(ier), L500)
        assert ier.value == 0
        assert np.array_equal(pf[:1], pf0[:1]) and (itrphi == 1 or np.array_equal(pf[:n0], pf0))
        assert np.allclose(th[:N], th0, rtol=0, atol=1e-9)
        for tb in (1, 2, 3):
            assert_stokes_close(tabs[tb][:N, :n0].T, up0[tb], "FIN up table %d" % tb)
            assert_stokes_close(tabs[7 + tb][:N, :n0].T, dn0[tb], "FIN down table %d" % tb)
        assert np.allclose(tabs[0][:N, :n0].T, up0[0], rtol=0, atol=1e-5)
    phi = 0.6
    _, xi0, xq0, xu0, an0 = orc.trphi(r.rec, N, o.rmu, r.ttot_tronc, r.tauout, phi, 1, o.n0, o.wind, o.ind_surf, 0, 1)
    xi, xq, xu, an = (np.zeros(2 * MX + 1) for _ in range(4))
    ier = C.c_int(99)
    lib.sos_trphi_(fstr(fos), ip(N), P(rmu), dp(r.ttot_tronc), dp(r.tauout), dp(phi), ip(1), ip(o.n0), dp(o.wind),
                   dp(o.ind_surf), ip(0), ip(0), dp(0.0), dp(0.0), dp(0.0), ip(0), ip(0), ip(0), dp(0.0), dp(0.0), ip(0),
                   dp(0.0), ip(1), P(xi), P(xq), P(xu), P(an), C.byref(ier), L500)
    assert ier.value == 0
    sel = np.r_[MX - N:MX, MX + 1:MX + N + 1]
    sel0 = np.r_[0:N, N + 1:2 * N + 1]
    assert_stokes_close(xi[sel], xi0[sel0], "XIT"); assert_stokes_close(xq[sel], xq0[sel0], "XQT")
    assert_stokes_close(xu[sel], xu0[sel0], "XUT")
    assert np.allclose(an[sel], an0[sel0], rtol=0, atol=1e-5)
    lib.sos_trphi_(fstr(fos), ip(N), P(rmu), dp(r.ttot_tronc), dp(r.tauout), dp(phi), ip(1), ip(o.n0), dp(o.wind),
                   dp(o.ind_surf), ip(0), ip(1), dp(0.1), dp(0.1), dp(0.1), ip(0), ip(0), ip(0), dp(0.0), dp(0.0), ip(0),
                   dp(0.0), ip(1), P(xi), P(xq), P(xu), P(an), C.byref(ier), L500)
    assert ier.value == 0                                # Roujean direct term (SOS_TRPHI.F:1047-1072; values: test_gpu_vs_reference)
    assert np.abs(xi[MX + 1:MX + N + 1] - xi0[N + 1:]).max() > 1e-6   # the direct term is there


@pytest.mark.parametrize("nbg,os_nb", [(24, 48), (79, 24)])
def test_solve_other_angle_counts(pkg, orc, solver, nbg, os_nb):
    """N=25 (default -ANG.Rad.NbGauss 24: 5 row groups per direction) and N=80 (the CTE_OS_NBMU_MAX cap: KP=480,
    two row tiles per direction, attenuation table not staged in shared memory)."""
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=nbg, tetas=35.0, os_nb=os_nb, surface="lambert", rho=0.15)
    assert o.nbmu == nbg + 1
    wl = syn.Workload("n%d" % o.nbmu, [o], [syn.Term(0, 0.5, *syn.profile(0.08, 8.0, 0.25, 2.0, 0.02)),
                                            syn.Term(0, 0.5, *syn.profile(0.08, 8.0, 0.25, 2.0, 0.9))])
    _check_terms(pkg, orc, solver, wl)


def test_band_properties_at_bench_size(pkg, orc, solver):
    """BASELINE configs[2] at the bench size (96 spectral points, ~600 term-solves): size-independent properties
    (the aggregated record is the AIK-weighted sum of the term records; counts within bounds; finite) plus a
    random sample of terms against the oracle."""
    syn = pkg.synth
    wl = syn.config_ckd_band(npoints=96, seed=20261021, nb_gauss=40, os_nb=80, surface="lambert", rho=0.1)
    tr, gr = solver.solve(wl)
    assert np.all(tr.ier == 0) and np.isfinite(tr.rec).all() and np.isfinite(gr.rec).all()
    assert np.all(tr.n_fourier >= 3) and np.all(tr.n_fourier <= 81)
    for i in range(len(wl.terms)):
        nf = tr.n_fourier[i]
        assert np.all(tr.n_scatter[i, :nf] >= 2) and np.all(tr.n_scatter[i, :nf] <= 100)
        assert np.all(tr.rec[i, nf:] == 0.0)
    aik = np.array([t.aik for t in wl.terms])
    grp = np.array([t.optics for t in wl.terms])
    for g in range(0, 96, 7):
        idx = np.where(grp == g)[0]
        ref = np.zeros_like(gr.rec[g])
        for i in idx:                                    # SOS_AGGREGATE order (SOS_AGGREGATE.F:397-413)
            ref = ref + aik[i] * tr.rec[i]
        assert np.array_equal(gr.rec[g], ref)
        assert abs(aik[idx].sum() - 1.0) < 1e-12
    rng = np.random.default_rng(3)
    for i in rng.choice(len(wl.terms), 6, replace=False):
        t = wl.terms[i]
        r = oracle_term(orc, wl.optics[t.optics], t)
        assert tr.n_fourier[i] == r.n_fourier and np.array_equal(tr.n_scatter[i, :r.n_fourier], r.n_scatter)
        assert_stokes_close(tr.rec[i, :r.n_fourier], r.rec, "bench-size term %d" % i)


def test_batch_trphi_band_solve(pkg, orc, solver):
    """Band-solve = term-solves + CKD sum + azimuth synthesis, all on the device (sosgpu_batch_trphi)."""
    syn = pkg.synth
    wl = syn.config_ckd_band(npoints=3, seed=5, nb_gauss=8, os_nb=16, surface="lambert", rho=0.1, max_terms=5)
    b = solver.upload(wl)
    tr, gr = solver.run(b)
    n, up, down = solver.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1)
    b.free()
    assert n == 13
    o = wl.optics[0]
    for g in range(3):
        nr = int(gr.n_rec[g])
        n0, pf, th, up0, dn0 = orc.trphi_option(gr.rec[g, :nr], o.nbmu, o.rmu, gr.ttot_tronc[g], gr.tauout[g], 0, o.n0,
                                                2.0, 1.34, 0, 2, 0.0, 30, 1)
        assert n0 == 13
        for tb in (1, 2, 3):
            assert_stokes_close(up[g, tb], up0[tb], "band up %d" % tb)
            assert_stokes_close(down[g, tb], dn0[tb], "band down %d" % tb)


def test_transmissions(pkg, orc, solver):
    """-SOS.Trans (SOS.F:605-637): TDIFMUS and TDIFMUG(1:N) from 1+N black-surface IS=0 solves per term."""
    syn = pkg.synth
    o = syn.make_optics(nb_gauss=8, tetas=35.0, os_nb=16, surface="brdf", rho=0.05)
    wl = syn.Workload("trans", [o], [syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.2, 2.0, 0.3))])
    r = oracle_term(orc, o, wl.terms[0], want_trans=True)
    tdifmus, tdifmug = solver.transmissions(wl)
    N = o.nbmu
    assert_stokes_close(tdifmus[0], r.tdifmus, "TDIFMUS")
    assert_stokes_close(tdifmug[0], r.tdifmug[N + 1:], "TDIFMUG")


def test_comm_api_single_rank(pkg, orc, solver):
    """The library-owned NCCL communicator with ONE rank: sosgpu_comm_init, the in-place reduce of the band sums + group
    metadata (identity here), sosgpu_batch_groups, and the table gather must reproduce the plain single-GPU results."""
    syn = pkg.synth
    wl = syn.config_ckd_band(npoints=3, seed=5, nb_gauss=8, os_nb=16, surface="lambert", rho=0.1, max_terms=5)
    solver.comm_init(1, 0, solver.comm_unique_id())
    b = solver.upload(wl)
    try:
        tr, gr = solver.run(b)
        n, up, down = solver.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1)
        b2 = solver.upload(wl)
        solver.run(b2, want_terms=False, want_groups=False, part_only=True)
        solver.reduce_groups(b2, root=0)
        g2 = solver.groups(b2)
        assert np.array_equal(g2.n_rec, gr.n_rec)
        assert np.array_equal(g2.rec, gr.rec)
        for k in ("emoins", "eplus"):
            assert_stokes_close(getattr(g2, k), getattr(gr, k), k)
        for k in ("ttot_tronc", "ttot_vrai", "tauout"):       # closed form -log(sum a e^-tau) vs the reference's running form
            assert np.allclose(getattr(g2, k), getattr(gr, k), rtol=1e-12, atol=1e-15), k
        n2, up2, down2 = solver.batch_trphi(b2, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1)
        for tb in (1, 2, 3):
            assert_stokes_close(up2[:, tb], up[:, tb], "reduced-band up table %d" % tb)
        # wavelength-sharded layout: gather of the synthesised tables (one rank: a device-to-host copy in rank order)
        solver.batch_trphi(b, 0, 2.0, 1.34, 0, 2, 0.0, 30, 1, download=False)
        gu, gd = solver.gather_tables(b, [b.ngroup], n, root=0)
        assert np.array_equal(gu, up) and np.array_equal(gd, down)
        b2.free()
    finally:
        b.free()
        solver.lib.sosgpu_comm_destroy(solver.ctx)
