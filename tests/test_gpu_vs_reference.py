"""GPU parity tests against the REFERENCE ITSELF (oracle/_ref/libsosref.so = the reference's Fortran statements translated
to C), one per BASELINE.json config, without the hand restatement in between:

  cfg1  demo:      N=41, OS_NB=80, 5 CKD terms, Lambert rho=0 + Cox-Munk glitter surface, View 1 (phi = 0 / 180)
  cfg2  demoPolar: the same solve synthesised on View 2 (dphi = 30, 13 azimuths)
  cfg3  CKD band:  the bench band, ALL 620 term-solves of the 96 spectral points
  cfg4  hyperspectral sweep: 64 wavelengths at N=25 with a BRDF/BPDF surface matrix per wavelength
  cfg5  angular stress: N=80 (CTE_OS_NBMU_MAX), OS_NB=200 (CTE_OS_NB_MAX), output at an altitude

The whole chain runs on each side: for cfg1/2 the GPU-generated glitter file feeds the GPU solve and the
reference-generated file feeds the reference solve.  Bar: number of Fourier orders identical, Stokes within 1e-9 relative
(1e-12 floor); every test prints its count-mismatch tally (mismatches / terms)."""
import copy
import os

import numpy as np
import pytest

import refdirect
from util import assert_stokes_close

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 1


@pytest.fixture(scope="module")
def ref():
    lib = refdirect.lib()
    if lib is None:
        pytest.fail("oracle/_ref/libsosref.so is missing: run __graft_entry__.build() where /root/reference exists")
    return lib


def _solve_both(pkg, solver, wl, ids=None):
    ids = list(range(len(wl.terms))) if ids is None else list(ids)
    rr = refdirect.runner()
    res, wall, busy = rr.solve_terms(wl, ids, CORES)
    b = solver.upload(wl, ids)
    tr, gr = solver.run(b)
    return ids, res, b, tr, gr


def _tally(name, bad, n):
    print("\n[count parity] %s: %d mismatches / %d term-solves %s" % (name, len(bad), n, bad[:5] if bad else ""))


def test_cfg1_cfg2_demo_glitter_chain(pkg, solver, ref, tmp_path):
    syn, fm = pkg.synth, pkg.formats
    wl = syn.config_demo(nterm=5, nb_gauss=40, os_nb=80, surface="glitter")
    o = wl.optics[0]
    N, os_ns, os_nm = o.nbmu, 80, 160
    assert N == 41 and o.rho == 0.0
    # surface file: each side makes its own
    surf_gpu, il = solver.glitter(N, o.rmu, o.ga, o.wind, o.ind_surf, o.os_nb, os_ns, os_nm)
    surf_ref = refdirect.glitter(ref, fm, str(tmp_path), N, o.rmu, o.ga, o.wind, o.ind_surf, o.os_nb, os_ns, os_nm)
    same = np.mean(surf_gpu.view(np.uint32) == surf_ref.view(np.uint32))
    print("\n[glitter N=41] REAL*4 records bit-identical: %.4f %%, max |diff| / max %.2e"
          % (100 * same, np.abs(surf_gpu - surf_ref).max() / np.abs(surf_ref).max()))
    assert same > 0.999
    wl_ref, wl_gpu = copy.deepcopy(wl), copy.deepcopy(wl)
    wl_ref.optics[0].surf, wl_gpu.optics[0].surf = surf_ref, surf_gpu
    rr = refdirect.runner()
    ids = list(range(len(wl.terms)))
    res, _, _ = rr.solve_terms(wl_ref, ids, CORES)
    b = solver.upload(wl_gpu)
    try:
        tr, gr = solver.run(b)
        bad = refdirect.compare_terms(tr, res, ids, wl, assert_stokes_close, "cfg1")
        _tally("cfg1/cfg2 demo + glitter chain (N=41, OS_NB=80)", bad, len(ids))
        assert not bad
        # SOS_AGGREGATE of the reference over its own term files
        agg_rec, agg_sc = rr.aggregate_point(ref, fm, str(tmp_path), N, [(wl.terms[i].aik, res[i]) for i in ids])
        nr = int(gr.n_rec[0])
        assert_stokes_close(gr.rec[0, :nr, :, :2 * N + 1], agg_rec[:nr], "aggregated records")
        assert not agg_rec[nr:].any()
        for k in ("emoins", "eplus", "ttot_tronc", "ttot_vrai", "tauout"):
            assert_stokes_close(getattr(gr, k)[0], agg_sc[k], k)
        # View 1 (cfg1) and View 2 (cfg2) with the sun-glint direct term
        for itrphi, phios, pas in ((1, 0.0, 0), (2, 0.0, 30)):
            n1, up, dn = solver.batch_trphi(b, 1, o.wind, o.ind_surf, 0, itrphi, phios, pas, 1)
            n0, pf, th, up0, dn0 = refdirect.trphi_option(ref, fm, str(tmp_path), agg_rec[:nr], N, o.rmu, o.ga, agg_sc["ttot_tronc"],
                                                          agg_sc["tauout"], 1, o.n0, o.wind, o.ind_surf, 0, itrphi, phios, pas)
            assert n0 == n1
            for tb in (1, 2, 3):
                assert_stokes_close(up[0, tb, :n0, :N], up0[tb], "view %d up table %d" % (itrphi, tb))
                assert_stokes_close(dn[0, tb, :n0, :N], dn0[tb], "view %d down table %d" % (itrphi, tb))
    finally:
        b.free()


def test_cfg3_bench_band_all_terms(pkg, solver, ref):
    """Every term-solve of the bench workload (BASELINE configs[2]: 96 points, 620 terms) against the reference."""
    wl = pkg.synth.config_ckd_band(npoints=96, seed=20261021, nb_gauss=40, os_nb=80, surface="lambert", rho=0.1)
    ids, res, b, tr, gr = _solve_both(pkg, solver, wl)
    b.free()
    bad = refdirect.compare_terms(tr, res, ids, wl, assert_stokes_close, "cfg3")
    _tally("cfg3 O2-A-like CKD band, all terms (N=41, OS_NB=80)", bad, len(ids))
    assert not bad


def test_cfg4_hyperspectral_64_wavelengths(pkg, solver, ref):
    wl = pkg.synth.config_hyperspectral(nwave=64, nb_gauss=24, os_nb=80)
    assert wl.optics[0].nbmu == 25
    ids, res, b, tr, gr = _solve_both(pkg, solver, wl)
    b.free()
    bad = refdirect.compare_terms(tr, res, ids, wl, assert_stokes_close, "cfg4")
    _tally("cfg4 hyperspectral sweep, 64 wavelengths (N=25, OS_NB=80, surface matrix)", bad, len(ids))
    assert not bad


def test_cfg5_angular_stress_caps(pkg, solver, ref):
    """N = CTE_OS_NBMU_MAX = 80 and OS_NB = CTE_OS_NB_MAX = 200 (SOS.h:471,480): 201 Fourier orders x 480-row contractions,
    output level inside the atmosphere."""
    wl = pkg.synth.config_angular_stress(nterm=8)
    o = wl.optics[0]
    assert o.nbmu == 80 and o.os_nb == 200 and o.zout != -1.0
    ids, res, b, tr, gr = _solve_both(pkg, solver, wl)
    b.free()
    bad = refdirect.compare_terms(tr, res, ids, wl, assert_stokes_close, "cfg5")
    _tally("cfg5 angular stress (N=80, OS_NB=200, zout=3 km)", bad, len(ids))
    assert not bad
    assert int(tr.n_fourier.max()) > 81                      # beyond what any OS_NB=80 case reaches


@pytest.mark.parametrize("model", ["roujean", "rondeaux", "breon", "nadal", "maignan"])
def test_trphi_land_surface_direct_terms(pkg, solver, ref, model, tmp_path):
    """SOS_TRPHI's direct-beam terms of the land-surface models (SOS_TRPHI.F:1047-1200 with SOS_CALC_F_ROUJEAN,
    SOS_ROUJEAN.F:891, and SOS_CALCG_MAIGNAN, SOS_SURFACE_BPDF.F:1606) against the reference's SOS_TRPHI_OPTION."""
    syn, fm = pkg.synth, pkg.formats
    o = syn.make_optics(nb_gauss=24, tetas=35.0, os_nb=32, surface="brdf", rho=0.0)
    wl = syn.Workload("land", [o], [syn.Term(0, 1.0, *syn.profile(0.05, 8.0, 0.2, 2.0, 0.1))])
    tr, gr = solver.solve(wl)
    nr = int(gr.n_rec[0])
    N = o.nbmu
    rec = gr.rec[0, :nr, :, :2 * N + 1]
    kw = dict(roujean=dict(roujean=(0.25, 0.04, 0.30)), rondeaux=dict(irondeaux=1), breon=dict(ibreon=1),
              nadal=dict(nadal=(0.017, 75.0)), maignan=dict(maignan=6.0))[model]
    solver.set_direct_models(**kw)
    try:
        for itrphi, phios, pas in ((1, 15.0, 0), (2, 0.0, 45)):
            n1, pf1, th1, up1, dn1 = solver.trphi_option(rec, N, o.rmu, gr.ttot_tronc[0], gr.tauout[0], 0, o.n0, 2.0, 1.5, 0,
                                                         itrphi, phios, pas, 1)
            bp = {k: v for k, v in (("irondeaux", kw.get("irondeaux", 0)), ("ibreon", kw.get("ibreon", 0)))}
            if "nadal" in kw:
                bp.update(inadal=1, alpha=kw["nadal"][0], beta=kw["nadal"][1])
            if "maignan" in kw:
                bp.update(imaignan=1, coef=kw["maignan"])
            n0, pf0, th0, up0, dn0 = refdirect.trphi_option(ref, fm, str(tmp_path), rec, N, o.rmu, o.ga, gr.ttot_tronc[0], gr.tauout[0],
                                                            0, o.n0, 2.0, 1.5, 0, itrphi, phios, pas, roujean=kw.get("roujean"), bpdf=bp)
            assert n0 == n1
            for tb in (1, 2, 3):
                assert_stokes_close(up1[tb], up0[tb], "%s up table %d" % (model, tb))
                assert_stokes_close(dn1[tb], dn0[tb], "%s down table %d" % (model, tb))
            # the direct term is really there: the table differs from the one without it
            solver.set_direct_models()
            _, _, _, upn, _ = solver.trphi_option(rec, N, o.rmu, gr.ttot_tronc[0], gr.tauout[0], 0, o.n0, 2.0, 1.5, 0, itrphi, phios, pas, 1)
            assert np.abs(up1[1] - upn[1]).max() > 1e-6
            solver.set_direct_models(**kw)
    finally:
        solver.set_direct_models()


@pytest.mark.parametrize("nbg,ind", [(12, 1.34), (40, 1.34), (24, 1.5)])
def test_mat_fresnel_isolated(pkg, solver, ref, nbg, ind, tmp_path):
    """SOS_MAT_FRESNEL (SOS_SURFACE.F:1235-1603) on its own: the device expansion + the E15.8 channel must give the very
    numbers the reference writes to RES_FRESNEL (8 significant digits: identical decimals, i.e. identical doubles)."""
    rmu, ga, n0, _ = pkg.synth.sos_angles(nbg, 35.0)
    N = (rmu.size - 1) // 2
    os_ns = 2 * nbg
    got = solver.mat_fresnel(N, rmu, ga, ind, os_ns)
    want = refdirect.mat_fresnel(ref, str(tmp_path), N, rmu, ga, ind, os_ns)
    for name, g, w in zip(("alpha", "beta", "gamma", "zeta"), got, want):
        assert np.array_equal(g, w), (name, np.abs(g - w).max())


def test_order1_isolated(pkg, solver, ref):
    """SOS_FSOURCE_ORDRE1 + boundary values + SOS_INTEGR_EPOPT alone (k_order1): IGMAX = 1 stops the reference after the first
    scattering order, for a Lambertian ground, a BRDF matrix and a flat Fresnel sea (SOS_FSOURCE_DIFF_FRESNEL1)."""
    syn = pkg.synth
    wl = syn.Workload("order1")
    for surface, rho in (("lambert", 0.3), ("brdf", 0.05), ("fresnel", 0.0)):
        wl.optics.append(syn.make_optics(nb_gauss=16, tetas=40.0, os_nb=32, surface=surface, rho=rho, igmax=1, seed=len(wl.optics)))
        wl.terms.append(syn.Term(len(wl.optics) - 1, 1.0, *syn.profile(0.08, 8.0, 0.25, 2.0, 0.4)))
    ids, res, b, tr, gr = _solve_both(pkg, solver, wl)
    b.free()
    bad = refdirect.compare_terms(tr, res, ids, wl, assert_stokes_close, "order1")
    assert not bad
    assert np.all(tr.n_scatter[:, :3] <= 2)


def test_land_surface_files_roujean_breon(pkg, solver, ref, tmp_path):
    """SOS_ROUJEAN (BRDF Fourier series), SOS_SURFACE_BPDF for the Rondeaux and Breon models and
    SOS_BPDF_AJOUT_BRDF, each against the reference through its files; then cfg4's surface for real: the Roujean + Breon
    file made on each side feeds each side's solve (N=25, OS_NB=80)."""
    syn, fm = pkg.synth, pkg.formats
    rmu, ga, n0, _ = syn.sos_angles(24, 35.0)
    N, os_nb, os_ns = (rmu.size - 1) // 2, 80, 48
    k = (0.25, 0.04, 0.30)
    ier, rj_ref = refdirect.roujean(ref, fm, str(tmp_path), N, rmu, ga, os_nb, *k)
    assert ier == 0
    rj = solver.roujean(N, rmu, os_nb, *k)
    same = np.mean(rj.view(np.uint32) == rj_ref.view(np.uint32))
    nz_ref = (np.abs(rj_ref[:, 0]).reshape(os_nb + 1, -1).max(axis=1) > 0).sum()
    nz = (np.abs(rj[:, 0]).reshape(os_nb + 1, -1).max(axis=1) > 0).sum()
    print("\n[roujean N=%d] REAL*4 records bit-identical: %.4f %%; orders with a non-zero coefficient: %d (reference %d)"
          % (N, 100 * same, nz, nz_ref))
    assert same > 0.995 and np.abs(rj - rj_ref).max() <= 2e-7 * np.abs(rj_ref).max()
    for isurf in (7, 4, 5):                                  # Maignan (C = 6), Rondeaux, Breon (the last one is used below)
        b_ref = refdirect.surface_bpdf(ref, fm, str(tmp_path), isurf, N, rmu, ga, 1.5, os_nb, os_ns, os_nb + os_ns, coef_c=6.0)
        b = solver.surface_bpdf(isurf, N, rmu, ga, 1.5, os_nb, os_ns, os_nb + os_ns, coef_c=6.0)
        eq = b.view(np.uint32) == b_ref.view(np.uint32)
        sig = np.abs(b_ref) > 1e-6 * np.abs(b_ref).max()
        print("[BPDF isurf=%d] REAL*4 records bit-identical: %.4f %% of all entries, %.4f %% of those above 1e-6 of the largest; "
              "max |diff| / max %.1e" % (isurf, 100 * eq.mean(), 100 * eq[sig].mean(), np.abs(b - b_ref).max() / np.abs(b_ref).max()))
        # Maignan's cusped G keeps (nearly) all OS_NM+1 orders of its series; the high output orders are cancellation noise of
        # about 1e-15 (against 0.03), where the last bits are arbitrary on either side: 1 % of the entries, none above 1e-6 of max
        assert eq[sig].mean() > 0.999 and eq.mean() > (0.98 if isurf == 7 else 0.999), isurf
        assert np.abs(b - b_ref).max() <= 2e-7 * np.abs(b_ref).max()
    s_ref = refdirect.bpdf_ajout_brdf(ref, fm, str(tmp_path), b_ref, rj_ref)
    s_gpu = solver.bpdf_ajout_brdf(b, rj)
    assert np.array_equal(solver.bpdf_ajout_brdf(b_ref, rj_ref).view(np.uint32), s_ref.view(np.uint32))
    # cfg4 for real: Roujean BRDF + Breon BPDF surface, each side with its own file
    import copy
    wl = syn.config_hyperspectral(nwave=3, nb_gauss=24, os_nb=80)
    for o in wl.optics:
        o.rho = 0.0
    wl_ref, wl_gpu = copy.deepcopy(wl), copy.deepcopy(wl)
    for o in wl_ref.optics:
        o.surf = s_ref
    for o in wl_gpu.optics:
        o.surf = s_gpu
    ids = list(range(len(wl.terms)))
    res, _, _ = refdirect.runner().solve_terms(wl_ref, ids, CORES)
    tr, gr = solver.solve(wl_gpu)
    bad = refdirect.compare_terms(tr, res, ids, wl, assert_stokes_close, "cfg4-real")
    _tally("cfg4 with the Roujean + Breon surface file made on each side (N=25)", bad, len(ids))
    assert not bad


# ---- the per-term profile chain ------------------------------------------------------------------------
def _chain_reference(ref, tmp, t, user, altabs, ro, terms):
    out = []
    for term in terms:
        ier, tau = refdirect.absprofile(ref, t, user, altabs, ro, term)
        assert ier == 0
        out.append((tau,) + refdirect.profile(ref, tmp, altabs, tau, term))
    return out


@pytest.mark.parametrize("seed", [0, 1])
def test_profile_chain_vs_reference(solver, ref, seed, tmp_path):
    """SOS_ABSPROFILE -> SOS_PROFILE -> PROFIL_TMP for 150 (wavelength, CKD term) entries in one device call, against the
    reference's routines called term by term: every NT identical, the absorption profile within 1e-13 (device exp / log are
    not glibc's), and the values SOS reads back from the text file identical except where a 1-ulp difference crosses a
    rounding boundary of the 8-digit decimal (counted, and bounded)."""
    import profile_cases as pc
    user, altabs, ro = pc.gas_atmosphere(seed)
    t = pc.ckd_tables(seed)
    terms = pc.make_terms(t, 150, seed)
    want = _chain_reference(ref, str(tmp_path), t, user, altabs, ro, terms)
    nt, z, h, pa, pm, ier, tau = solver.profile_chain(t, user, altabs, ro, terms, text_hop=True, want_tauabs=True)
    nvals = ndiff = nbad_nt = 0
    branches = {"nogas": 0, "weak": 0, "strong": 0}
    for i, (tau_r, ier_r, nt_r, _, z_r, h_r, pa_r, pm_r) in enumerate(want):
        assert (ier[i] != 0) == (ier_r != 0), (i, ier[i], ier_r)
        if ier_r != 0:
            continue
        np.testing.assert_allclose(tau[i], tau_r, rtol=1e-13, atol=1e-300, err_msg="TAUABSTOT term %d" % i)
        branches["nogas" if tau_r[-1] == 0 else "strong" if tau_r[-1] > 1.5 else "weak"] += 1
        if nt[i] != nt_r:
            nbad_nt += 1
            continue
        for a, b in ((z[i], z_r), (h[i], h_r), (pa[i], pa_r), (pm[i], pm_r)):
            a = a[:nt_r + 1]
            nvals += a.size
            ndiff += int((a != b).sum())
            np.testing.assert_allclose(a, b, rtol=2e-7, atol=1e-5, err_msg="profile term %d" % i)
        assert not h[i, nt_r + 1:].any()
    print("\n[profile chain seed %d] NT mismatches %d / %d terms; text values differing %d / %d; branches %s"
          % (seed, nbad_nt, len(terms), ndiff, nvals, branches))
    assert nbad_nt == 0
    assert ndiff <= max(2, nvals // 100000)
    assert min(branches.values()) >= 2


def test_profile_stages_and_errors(solver, ref, tmp_path):
    """The two stages separately (sosgpu_absprofile, sosgpu_profile incl. IPROFIL = 2 and the unrounded output), the error
    codes per term, and the solver fed from the chain's arrays."""
    import profile_cases as pc
    user, altabs, ro = pc.gas_atmosphere(2)
    t = pc.ckd_tables(2)
    terms = pc.make_terms(t, 24, 2)
    tau, ier = solver.absprofile(t, user, ro, terms)
    assert not ier.any()
    for i, term in enumerate(terms):
        _, tau_r = refdirect.absprofile(ref, t, user, altabs, ro, term)
        np.testing.assert_allclose(tau[i], tau_r, rtol=1e-13, atol=1e-300)
    # stage 2 from the reference's absorption profiles: the text-hop values equal the reference's file
    taus = np.array([refdirect.absprofile(ref, t, user, altabs, ro, term)[1] for term in terms])
    nt, z, h, pa, pm, ier = solver.profile(altabs, taus, terms, text_hop=True)
    nt0, z0, h0, pa0, pm0, _ = solver.profile(altabs, taus, terms, text_hop=False)
    assert not ier.any() and np.array_equal(nt, nt0)
    for i, term in enumerate(terms):
        ier_r, nt_r, _, z_r, h_r, pa_r, pm_r = refdirect.profile(ref, str(tmp_path), altabs, taus[i], term)
        assert ier_r == 0 and nt[i] == nt_r
        np.testing.assert_allclose(h[i, :nt_r + 1], h_r, rtol=2e-7)
        assert np.mean(h[i, :nt_r + 1] == h_r) > 0.99
        np.testing.assert_allclose(h0[i, :nt_r + 1], h_r, rtol=1e-7, atol=1e-12)      # unrounded values sit within the text precision
        assert (np.diff(h0[i, :nt_r + 1]) > 0).all() and (np.diff(z0[i, :nt_r + 1]) < 0).all()
    # IPROFIL = 2 and the error exits (bad altitudes 1010, too little aerosol 1020, IPROFIL 940)
    t2 = [dict(terms[0], iprofil=2, absprofil=7, tr=0.2, ta=0.3, zmin=1.0, zmax=4.0),
          dict(terms[0], iprofil=2, absprofil=7, tr=0.2, ta=0.3, zmin=4.0, zmax=1.0),
          dict(terms[0], iprofil=2, absprofil=7, tr=0.2, ta=1e-9, zmin=0.0, zmax=3.0),
          dict(terms[0], iprofil=3),
          dict(terms[0], iprofil=1, absprofil=2, tr=0.345, ta=1.17, ha=2.25)]     # > 600 levels: the reference never returns
    tau2 = np.zeros((5, 50))
    tau2[4] = np.linspace(0.0, 1.31, 50)
    nt, z, h, pa, pm, ier = solver.profile(altabs, tau2, t2, text_hop=True)
    assert list(ier) == [0, 1010, 1020, 940, 9600] and not nt[1:].any()
    import ctypes as C
    import shutil
    fresh = str(tmp_path / "libsosref_fresh.so")                   # IPROFIL = 2 reads Hmol(0) before setting it: fresh image
    shutil.copy(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsosref.so"), fresh)
    ier_r, nt_r, _, z_r, h_r, pa_r, pm_r = refdirect.profile(C.CDLL(fresh), str(tmp_path), altabs, np.zeros(50), t2[0])
    assert ier_r == 0 and nt[0] == nt_r
    np.testing.assert_allclose(h[0, :nt_r + 1], h_r, rtol=2e-7)
    np.testing.assert_allclose(pa[0, :nt_r + 1], pa_r, rtol=2e-7, atol=1e-12)


def test_profile_chain_feeds_the_solver(pkg, solver, ref, tmp_path):
    """End of the N1 row: profiles made by the device chain go into the device solve, profiles made by the reference chain go
    into the reference solve; same Fourier-order counts and Stokes within 1e-9 on every term."""
    import profile_cases as pc
    syn = pkg.synth
    user, altabs, ro = pc.gas_atmosphere(4)
    t = pc.ckd_tables(4)
    terms = [dict(x, absprofil=2, ta=0.15, ha=2.0, tr=0.05, hr=8.0) for x in pc.make_terms(t, 12, 4)]
    nt, z, h, pa, pm, ier = solver.profile_chain(t, user, altabs, ro, terms, text_hop=True)
    assert not ier.any()
    wl_ref, wl_gpu = syn.Workload("chain-ref"), syn.Workload("chain-gpu")
    for w in (wl_ref, wl_gpu):
        w.optics.append(syn.make_optics(nb_gauss=24, tetas=35.0, os_nb=40, surface="lambert", rho=0.1))
    for i, term in enumerate(terms):
        ier_r, tau_r = refdirect.absprofile(ref, t, user, altabs, ro, term)
        _, nt_r, _, z_r, h_r, pa_r, pm_r = refdirect.profile(ref, str(tmp_path), altabs, tau_r, term)
        assert nt[i] == nt_r
        n = int(nt[i]) + 1
        wl_ref.terms.append(syn.Term(0, 1.0 / len(terms), z_r, h_r, pa_r, pm_r))
        wl_gpu.terms.append(syn.Term(0, 1.0 / len(terms), z[i, :n].copy(), h[i, :n].copy(), pa[i, :n].copy(), pm[i, :n].copy()))
    ids = list(range(len(terms)))
    res, _, _ = refdirect.runner().solve_terms(wl_ref, ids, CORES)
    b = solver.upload(wl_gpu)
    try:
        tr, gr = solver.run(b)
        bad = refdirect.compare_terms(tr, res, ids, wl_ref, assert_stokes_close, "profile chain -> solve")
        _tally("profile chain -> solve (12 terms, N=25)", bad, len(ids))
        assert not bad
    finally:
        b.free()


def test_profile_shims_write_the_reference_file(pkg, ref, tmp_path):
    """The gfortran-ABI symbols sos_absprofile_ / sos_profile_ of libsosgpu.so (INTEGER*2 arguments, hidden string length,
    PROFIL_TMP written with format 20): the file is the reference's file byte for byte."""
    import ctypes as C
    import profile_cases as pc
    import importlib
    lib = importlib.import_module("radiativetransfer-sos_b200.api").load_library()
    user, altabs, ro = pc.gas_atmosphere(5)
    t = pc.ckd_tables(5)
    ip, dp, sp = (lambda v: C.byref(C.c_int(v))), (lambda v: C.byref(C.c_double(v))), (lambda v: C.byref(C.c_short(v)))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    nsame = 0
    terms = pc.make_terms(t, 16, 5)
    for term in terms:
        ier_r, tau_r = refdirect.absprofile(ref, t, user, altabs, ro, term)
        tau, ier = np.zeros(50), C.c_int(99)
        iabs = np.ones(8, dtype=np.int16)
        lib.sos_absprofile_(sp(term["absprofil"]), dp(13000.0), ip(term["lamb1"]), iabs.ctypes.data_as(C.POINTER(C.c_short)), P(user),
                            P(altabs), P(ro), t["nexp"].ctypes.data_as(C.POINTER(C.c_int)), P(t["ki"]), P(t["kh"]),
                            *[ip(v) for v in term["ik"]], P(t["tab_pres"]), ip(t["nb_pres"]), P(t["tab_temp"]), ip(t["nb_temp"]),
                            P(t["tab_conc"]), ip(t["nb_conc"]), P(tau), ip(0), ip(0), C.byref(ier))
        assert ier.value == 0 and ier_r == 0
        np.testing.assert_allclose(tau, tau_r, rtol=1e-13, atol=1e-300)
        _, nt_r, text_r, *_ = refdirect.profile(ref, str(tmp_path), altabs, tau_r, term)
        f = str(tmp_path / "PROFIL_GPU.txt")
        nt, ier = C.c_int(0), C.c_int(99)
        lib.sos_profile_(sp(term["iprofil"]), dp(term["tr"]), dp(term["hr"]), dp(term["ta"]), dp(term["ha"]), dp(term["zmin"]),
                         dp(term["zmax"]), sp(term["absprofil"]), P(altabs), P(tau_r), ip(0), ip(0),
                         C.create_string_buffer(f.encode().ljust(500), 500), C.byref(nt), C.byref(ier), C.c_size_t(500))
        assert ier.value == 0 and nt.value == nt_r
        text = open(f, "rb").read()
        assert len(text) == len(text_r)
        nsame += text == text_r
    print("\n[PROFIL_TMP] files byte-identical to the reference's: %d / %d" % (nsame, len(terms)))
    assert nsame >= len(terms) - 1


def _parse_updown(path):
    rows = []
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        rows.append([float(x) for x in line.split()])
    return np.array(rows)


def test_band_front_end_files(pkg, solver, ref, tmp_path):
    """the multi-wavelength front end (band.run_band: CKD term list -> device profile chain -> device solves
    and CKD sums -> device synthesis -> SOS_Up/Down.txt, SOS_Result.bin, Trans / Flux files per wavelength) against the
    reference's flow run stage by stage for every wavelength: SOS_ABSPROFILE, SOS_PROFILE, SOS + SOS_OS, SOS_AGGREGATE,
    SOS_TRPHI_OPTION.  CKD weights against a direct rendering of the nested loops."""
    import importlib
    import itertools
    import profile_cases as pc
    band = importlib.import_module("radiativetransfer-sos_b200.band")
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    syn, fm = pkg.synth, pkg.formats
    user, altabs, ro = pc.gas_atmosphere(6)
    t = pc.ckd_tables(6)
    counts = [int(np.prod(t["nexp"][:, l])) for l in range(pc.NWVL)]
    lambs = [l + 1 for l in np.argsort(counts) if 2 <= counts[l] <= 12][:2] + [int(np.argmin(counts)) + 1]
    waves = []
    for n, l in enumerate(lambs):
        o = syn.make_optics(nb_gauss=12, tetas=30.0 + 5 * n, os_nb=24, surface="lambert", rho=0.05 + 0.1 * n)
        waves.append(band.Wavelength(optics=o, lamb1=l, tr=0.08 - 0.02 * n, ta=0.1 + 0.05 * n, name="w%d" % n))
    waves[-1].absprofil = 7                                         # one wavelength without gaseous absorption
    # CKD weights: the eight nested loops, IK8 innermost
    iks, aik = band.enumerate_ckd_terms(t["nexp"], t["ai"], lambs[0])
    ne = [int(t["nexp"][k, lambs[0] - 1]) for k in range(8)]
    want = list(itertools.product(*[range(1, n + 1) for n in ne]))
    assert iks == want
    raw = np.array([np.prod([t["ai"][ik[k] - 1, k, lambs[0] - 1] for k in range(8)]) for ik in want])
    np.testing.assert_allclose(aik, raw / raw.sum(), rtol=1e-14)
    out = str(tmp_path / "band")
    res = band.run_band(solver, t, t["ai"], user, altabs, ro, waves, itrphi=2, phios=0.0, pas_phi=60, outdir=out, trans=True, flux=True)
    assert res.nterm == [len(iks), int(np.prod(t["nexp"][:, lambs[1] - 1])), 1]
    rr = refdirect.runner()
    nbad = nlines = nsame = 0
    for w, wv in enumerate(waves):
        o, N = wv.optics, wv.optics.nbmu
        if wv.absprofil == 7:
            ik_w, aik_w = [(1,) * 8], [1.0]
        else:
            ik_w, aik_w = band.enumerate_ckd_terms(t["nexp"], t["ai"], wv.lamb1)
        wl_ref = syn.Workload("ref")
        wl_ref.optics.append(o)
        for ik, a in zip(ik_w, aik_w):
            term = dict(lamb1=wv.lamb1, ik=ik, absprofil=wv.absprofil, iprofil=1, tr=wv.tr, hr=wv.hr, ta=wv.ta, ha=wv.ha, zmin=0.0, zmax=0.0)
            _, tau_r = refdirect.absprofile(ref, t, user, altabs, ro, term)
            ier_r, nt_r, _, z_r, h_r, pa_r, pm_r = refdirect.profile(ref, str(tmp_path), altabs, tau_r, term)
            assert ier_r == 0
            wl_ref.terms.append(syn.Term(0, a, z_r, h_r, pa_r, pm_r))
        ids = list(range(len(ik_w)))
        r, _, _ = rr.solve_terms(wl_ref, ids, CORES)
        agg_rec, agg_sc = rr.aggregate_point(ref, fm, str(tmp_path), N, [(aik_w[i], r[i]) for i in ids])
        nr = int(res.groups.n_rec[w])
        if agg_rec[nr:].any() or nr > agg_rec.shape[0]:
            nbad += 1
            continue
        got = fm.read_result_bin(os.path.join(res.dirs[w], "SOS_Result.bin"), N)
        assert_stokes_close(got, agg_rec[:nr], "SOS_Result.bin of wavelength %d" % w)
        for k in ("emoins", "eplus", "ttot_tronc", "ttot_vrai", "tauout"):
            assert_stokes_close(getattr(res.groups, k)[w], agg_sc[k], k)
        n0, pf, th, up0, dn0 = refdirect.trphi_option(ref, fm, str(tmp_path), agg_rec[:nr], N, o.rmu, o.ga, agg_sc["ttot_tronc"],
                                                      agg_sc["tauout"], o.igli, o.n0, o.wind, o.ind_surf, o.ifresnel, 2, 0.0, 60)
        assert n0 == res.nphi == 7
        fu, fd = str(tmp_path / "REF_Up.txt"), str(tmp_path / "REF_Down.txt")
        api.write_updown(fu, fd, N, 2, 0.0, 60, o.zout, pf, th, up0, dn0)
        for mine, theirs in ((os.path.join(res.dirs[w], "SOS_Up.txt"), fu), (os.path.join(res.dirs[w], "SOS_Down.txt"), fd)):
            a, b = open(mine).read().split("\n"), open(theirs).read().split("\n")
            assert len(a) == len(b)
            nlines += len(a)
            nsame += sum(x == y for x, y in zip(a, b))
            pa_, pb_ = _parse_updown(mine), _parse_updown(theirs)
            assert pa_.shape == pb_.shape and pa_.shape[0] == 7 * N
            np.testing.assert_allclose(pa_[:, :3], pb_[:, :3], atol=0.011)            # angles, printed with 2 decimals
            np.testing.assert_allclose(pa_[:, 3:6], pb_[:, 3:6], rtol=3e-6, atol=1e-12)   # I, Q, U printed with 6 digits
        # flux file: the three scalars SOS_PROC derives, and the files exist
        for f in ("SOS_Trans.txt", "SOS_Flux.txt"):
            assert os.path.getsize(os.path.join(res.dirs[w], f)) > 200
    print("\n[band front end] count mismatches %d / %d wavelengths; SOS_Up/Down lines identical to the reference-side rendering: %d / %d"
          % (nbad, len(waves), nsame, nlines))
    assert nbad == 0 and nsame >= nlines - 4
