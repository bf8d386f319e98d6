"""The reference's own driver SOS_PROC (SOS_PROC.F:415-481, the entry binding/run_sos.py calls through f2py; in oracle/_ref with
SOS_PREPA_OS and the surface file names) run from its arguments to its result files.

CPU part: the flow the GPU tests of the keyword front end compare with (the reference's routines called one after the other by
tests/test_frontend.py::test_gpu_run_from_keywords_demo: SOS_AEROSOLS chain -> SOS_GLITTER -> SOS_PROFILE -> SOS -> SOS_AGGREGATE ->
SOS_TRPHI_OPTION) gives what the driver gives: aerosol file, SOS_Result.bin and the 14 output tables.  So "front end vs the
routines one after the other" is "front end vs SOS_PROC".

GPU part: sos.sos_proc (the f2py-shaped entry over the device front end) against SOS_PROC itself on the same arguments."""
import importlib
import os

import numpy as np
import pytest

import aerosol_cases as ac
import refdirect

ARGS = dict(wa_simu=0.910, tetas=35.0, nbmu_gauss_lum=12, nbmu_gauss_mie=20, waref_aot=0.55, aot_ref=0.3, itronc_aer=1, imod_aer=1,
            imodele_wmo=2, hr=8.0, ha=2.0, iprofil=1, psurf=1013.0, absprofil=7, isurf=1, surf_ind=1.34, wind=2.0, rho=0.0, itrphi=2,
            pas_phi=60, igmax=100, zout=-1.0, ier=0)


def _installation(tmp):
    """A stand-in for the user's installation of the reference ($SOS_ABS_ROOT/fic with the WMO data file)."""
    os.makedirs(os.path.join(tmp, "abs_root", "fic"), exist_ok=True)
    os.environ["SOS_ABS_ROOT"] = os.path.join(tmp, "abs_root")
    ac.write_sf_files(os.path.join(os.environ["SOS_ABS_ROOT"], "fic"))
    return ac.write_wmo_file(os.path.join(os.environ["SOS_ABS_ROOT"], "fic", "Data_WMO_cor_2015_12_16"))


def _driver(ref, tmp, **more):
    root = os.path.join(tmp, "drv")
    ier, out = refdirect.sos_proc(ref, resroot=root, dir_mie=os.path.join(root, "MIE"), dir_surf=os.path.join(root, "SURF"), trace=0,
                                  **{**ARGS, **more})
    assert ier == 0, "reference SOS_PROC IER=%d" % ier
    return root, out


def test_driver_refuses_what_the_front_end_refuses(tmp_path):
    """SOS_PROC's argument checks on the cases sos.to_keywords / the front end refuse too (SOS_PROC.F:1311, 1320, 2210-2226)."""
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_proc_"):
        pytest.skip("oracle/_ref without SOS_PROC")
    tmp = str(tmp_path)
    _installation(tmp)
    root = os.path.join(tmp, "drv")
    base = dict(resroot=root, dir_mie=root + "/MIE", dir_surf=root + "/SURF", trace=0)
    assert refdirect.sos_proc(ref, **{**ARGS, "trace": 0})[0] == 1                              # ERROR_1000: no working folder
    a = dict(ARGS)
    del a["aot_ref"]
    assert refdirect.sos_proc(ref, **base, **a)[0] == 1                                         # ERROR_2301: AOT not defined
    assert refdirect.sos_proc(ref, **base, **{**ARGS, "isurf": 6, "rho": 0.1})[0] == 1          # Nadal's BPDF not supported
    assert refdirect.sos_proc(ref, **base, **{**ARGS, "igmax": -999})[0] == 1                   # ERROR_2604


def _one_after_the_other(ref, pkg, tmp, wmo, wa, gas=None, mode=2):
    """The reference's routines called one after the other, as the GPU tests of the keyword front end run them, for the description
    ARGS at wavelength wa.  gas: what absprofile.prepare returns (pinned against SOS_PREPA_ABSPROFILE by tests/test_absprofile.py);
    mode: -SOS.AbsModeCKD.  -> dict(aer lines, rec, scalars, nphi, phi, theta, up, down, nterm, tauabs of the Flux file)."""
    fe = importlib.import_module("radiativetransfer-sos_b200.frontend")
    band = importlib.import_module("radiativetransfer-sos_b200.band")
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    import test_aerosol_chain as tac
    syn, fm = pkg.synth, pkg.formats
    waref, aot = ARGS["waref_aot"], ARGS["aot_ref"]
    os_nb, os_ns, os_nm = fe.expansion_orders(20, 12)
    nbm, xmu, xhr = fe.mie_angles(20)
    k1 = {}
    for w in (wa, waref):
        e, v1, v2, mr, mi, vol = ac.ref_wmo_params(ref, wmo, w)
        comps = [(mr[i], mi[i], 0.0001, (4000.0, 50.0, 800.0, 10.0)[i], 1, v1[i], v2[i], -999.0, w) for i in (1, 2)]
        n = np.array([0.0 / vol[0], np.float64(np.float32(0.05)) / vol[1], np.float64(np.float32(0.95)) / vol[2], 0.0 / vol[3]])
        ntot = 0.0
        for x in n:
            ntot = ntot + x
        k1[w] = tac._reference_chain(ref, tmp, nbm, xmu, xhr, comps, [(2, [0, 1], [n[1] / ntot, n[2] / ntot], 1)], os_nb)[2][0]
    dd = k1[wa]
    ta = dd["kmat1"] / k1[waref]["kmat1"] * aot
    piz = dd["kmat2"] / dd["kmat1"]
    ct = dd["coef_tronca"]
    fa = os.path.join(tmp, "Aer_chain.txt")
    api.write_aerosols(fa, os_nb, dd["kmat1"], dd["kmat2"], ct / 2.0 + (1.0 - ct / 2.0) * dd["beta11"][1] / 3.0, ct,
                       piz * (1.0 - ct / 2.0) / (1.0 - piz * ct / 2.0), dd["alp"], dd["beta11"], dd["gamma12"], dd["zeta"])
    r8 = lambda v: np.array([float("%.7E" % x) for x in v])
    rmu, ga, n0, _ = syn.sos_angles(12, 35.0)
    N = (rmu.size - 1) // 2
    surf = refdirect.glitter(ref, fm, tmp, N, rmu, ga, 2.0, 1.34, os_nb, os_ns, os_nm)
    o = syn.Optics(nbmu=N, rmu=rmu.copy(), ga=ga, n0=n0, tetas=35.0, os_nb=os_nb, alpha=r8(dd["alp"]), beta=r8(dd["beta11"]),
                   gamma=r8(dd["gamma12"]), zeta=r8(dd["zeta"]), a_trunc=float("%.5f" % ct),
                   piztr=float("%.5f" % (piz * (1.0 - ct / 2.0) / (1.0 - piz * ct / 2.0))), rho=0.0, imat_surf=1, igli=1, surf=surf,
                   ind_surf=1.34, wind=2.0, igmax=100, ipolar=1, zout=-1.0)
    tr = fe.rayleigh_thickness(1013.0, wa)
    term = dict(lamb1=1, ik=(1,) * 8, absprofil=7, iprofil=1, tr=tr, hr=8.0, ta=ta, ha=2.0, zmin=0.0, zmax=0.0)
    altabs = np.linspace(120.0, 0.0, 50)
    if gas is None:
        terms, aik, taus = [term], [1.0], [np.zeros(50)]
    else:
        t, altabs = gas["tables"][0], gas["altabs"]
        iks, aik = band.enumerate_ckd_terms(t["nexp"], gas["kdis_ai"][0], gas["lamb1"][0])
        terms = [dict(term, lamb1=gas["lamb1"][0], ik=ik, absprofil=0) for ik in iks]
        taus = []
        for tm in terms:
            ier, tau = refdirect.absprofile(ref, t, gas["userprofil"], altabs, gas["ro"], tm)
            assert ier == 0
            taus.append(tau)
        if mode == 2:                                             # SOS_PROC.F:3613-3662
            terms, aik, taus = terms[:1], [1.0], [band.estimated_absorption(aik, np.array(taus))]
    wl = syn.Workload("ref")
    wl.optics.append(o)
    for tm, a, tau in zip(terms, aik, taus):
        ier_r, nt_r, _, z_r, h_r, pa_r, pm_r = refdirect.profile(ref, tmp, altabs, tau, tm)
        assert ier_r == 0
        wl.terms.append(syn.Term(0, a, z_r, h_r, pa_r, pm_r))
    rr = refdirect.runner()
    r, _, _ = rr.solve_terms(wl, list(range(len(terms))), min(8, os.cpu_count() or 1))
    if len(terms) == 1 and (gas is None or mode == 2):            # one solve, no SOS_AGGREGATE (SOS_PROC.F:3700-3708)
        rec, sc = r[0]["rec"], {k: r[0][k] for k in ("ttot_tronc", "ttot_vrai", "tauout", "emoins", "eplus")}
    else:
        rec, sc = rr.aggregate_point(ref, fm, tmp, N, [(a, r[i]) for i, a in enumerate(aik)])
    nphi, pf, th, up, down = refdirect.trphi_option(ref, fm, tmp, rec, N, o.rmu, o.ga, sc["ttot_tronc"], sc["tauout"], 1, o.n0, 2.0, 1.34,
                                                    0, 2, 0.0, 60)
    return dict(aer=open(fa).read().split("\n"), ct=ct, rec=rec, sc=sc, nphi=nphi, phi=pf, theta=th, up=up, down=down, N=N, rmu=rmu,
                nterm=len(terms), tauabs=taus[-1], ta=ta, tr=tr)


def _check_against_driver(root, out, m, what):
    n_drv, ind, phi_d, vza_d, tabs_d, (tdir, fdd, fd, eplus, ct_d) = out[0], out[1], out[2], out[3], out[4:18], out[18:]
    N, nphi = m["N"], m["nphi"]
    la = open(os.path.join(root, "SOS", "Aerosols.txt")).read().split("\n")
    assert la == m["aer"], [(x, y) for x, y in zip(la, m["aer"]) if x != y][:3]
    assert m["ct"] == ct_d and N == n_drv and not ind.any()
    fm = importlib.import_module("radiativetransfer-sos_b200.formats")
    got = fm.read_result_bin(os.path.join(root, "SOS", "SOS_Result.bin"), N)
    nr = got.shape[0]
    rec = m["rec"]
    assert rec.shape[0] >= nr and not rec[nr:].any(), "number of Fourier orders differs"
    assert np.array_equal(got, rec[:nr]), (what, np.abs(got - rec[:nr]).max())
    assert eplus == m["sc"]["eplus"]
    assert nphi == 7 and np.array_equal(m["phi"], phi_d[:7]) and not phi_d[7:].any()
    assert np.array_equal(m["theta"], vza_d[:N]) and not vza_d[N:].any()
    for t in range(7):
        assert np.array_equal(m["up"][t], tabs_d[t][:7, :N]) and np.array_equal(m["down"][t], tabs_d[7 + t][:7, :N]), (what, t)
        assert not tabs_d[t][7:].any() and not tabs_d[t][:, N:].any()
    return nr, tdir, fdd, fd, eplus


def test_routines_one_after_the_other_are_the_driver(pkg, tmp_path):
    """No gaseous absorption (the demo's description at 12 / 20 Gauss angles): SOS_PROC solves once and does not call SOS_AGGREGATE
    (IMODE_CKD_CALCUL = 2 is forced, SOS_PROC.F:2366), so the optical thicknesses SOS_TRPHI_OPTION gets are those of SOS itself."""
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_proc_"):
        pytest.skip("oracle/_ref without SOS_PROC")
    tmp = str(tmp_path)
    wmo = _installation(tmp)
    root, out = _driver(ref, tmp)
    assert sorted(os.listdir(os.path.join(root, "SOS"))) == ["Aer_UsedAngles.txt", "Aerosols.txt", "SOS_Result.bin", "SOS_UsedAngles.txt"]
    m = _one_after_the_other(ref, pkg, tmp, wmo, ARGS["wa_simu"])
    nr, tdir, fdd, fd, eplus = _check_against_driver(root, out, m, "no gas")
    # what the host side of the product derives for the same description
    N, rmu, sc = m["N"], m["rmu"], m["sc"]
    assert np.allclose(np.degrees(np.arccos(np.asarray(rmu)[N + 1:2 * N + 1])), out[3][:N], rtol=1e-14, atol=0)
    t_d, fdd_d, fd_d = api.write_flux("NO_OUTPUT", 35.0, sc["ttot_tronc"], sc["ttot_vrai"], sc["emoins"], sc["eplus"], 0.0, 8.0, 0.0, 2.0,
                                      np.zeros(50), np.zeros(50))
    assert abs(t_d - tdir) < 1e-15 and abs(fdd_d - fdd) < 1e-15 and abs(fd_d - fd) < 1e-15
    print("\n[reference driver] SOS_PROC (12 / 20 Gauss angles, WMO maritime, rough sea, no gas) = its routines one after the other: "
          "Aerosols.txt %d lines, SOS_Result.bin %d records bit for bit, 14 tables x %d azimuths x %d angles bit for bit, "
          "Tdir %.6f Fdd %.6f Fd %.6f E+ %.6f" % (len(m["aer"]), nr, m["nphi"], N, tdir, fdd, fd, eplus))


@pytest.mark.parametrize("mode", [1, 2])
def test_driver_with_gas_both_ckd_modes(pkg, tmp_path, mode):
    """Gaseous absorption from a user profile file and CKD coefficient files (a stand-in installation): -SOS.AbsModeCKD 1 (one solve
    per CKD term, SOS_AGGREGATE) and 2 (one solve on the absorption estimated from the CKD terms: band.estimated_absorption) --
    the term list, the weights, the estimated profile and the order of the calls against SOS_PROC itself, bit for bit."""
    import profile_cases as pc
    import test_absprofile as tab
    ab = importlib.import_module("radiativetransfer-sos_b200.absprofile")
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_proc_"):
        pytest.skip("oracle/_ref without SOS_PROC")
    tmp = str(tmp_path)
    wmo = _installation(tmp)
    inst = os.environ["SOS_ABS_ROOT"]
    pc.write_ckd_files(inst, pc.ckd_tables(4))
    fprof = os.path.join(tmp, "profile.txt")
    tab._write_profile(fprof, tab._user())
    wa = 1e4 / 13255.0
    more = dict(wa_simu=wa, absprofil=0, ficabsprofil=fprof, nustep=10.0, imode_ckd_calcul=mode, ficflux="Flux.txt")
    root, out = _driver(ref, tmp, **more)
    nd = -999.0
    gas = ab.prepare(api.load_library(), [wa], 10.0, 0, fprof, 1013.0, nd, nd, nd, nd, sos_abs_root=inst)
    m = _one_after_the_other(ref, pkg, tmp, wmo, wa, gas=gas, mode=mode)
    nr, tdir, fdd, fd, eplus = _check_against_driver(root, out, m, "gas, mode %d" % mode)
    # the Flux file of the driver against the product's writer on the same values (TAUABS as SOS_PROC leaves it, :3860)
    mine = os.path.join(tmp, "Flux_mine.txt")
    sc = m["sc"]
    api.write_flux(mine, 35.0, sc["ttot_tronc"], sc["ttot_vrai"], sc["emoins"], sc["eplus"], m["tr"], 8.0, m["ta"], 2.0,
                   np.asarray(gas["userprofil"])[:, 0], m["tauabs"])
    a, b = open(os.path.join(root, "SOS", "Flux.txt")).read().split("\n"), open(mine).read().split("\n")
    it = iter(b)                                                  # the translated library writes the FORMAT-ted lines only (list-directed
    same = sum(any(x == y for y in it) for x in a)                # output has no pinned layout): they are mine, in order
    print("\n[reference driver] gas (user profile, CKD files), -SOS.AbsModeCKD %d: %d solve(s), SOS_Result.bin %d records and the 14 "
          "tables bit for bit; formatted lines of the Flux file identical %d / %d; Tdir %.6f Fd %.6f" % (mode, m["nterm"], nr, same, len(a), tdir, fd))
    assert same == len(a) == 55


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["no gas", "gas, -SOS.AbsModeCKD 2", "gas, -SOS.AbsModeCKD 1"], ids=["nogas", "gas_mode2", "gas_mode1"])
def test_gpu_sos_proc_against_the_reference_driver(solver, tmp_path, case):
    """sos.sos_proc on the device vs SOS_PROC of the reference library, same arguments: the 23 outputs of the f2py wrapper and
    SOS_Result.bin.  Gas: user profile file + CKD coefficient files of a stand-in installation, both CKD modes."""
    import profile_cases as pc
    import test_absprofile as tab
    sos = importlib.import_module("radiativetransfer-sos_b200.sos")
    fm = importlib.import_module("radiativetransfer-sos_b200.formats")
    ref = refdirect.lib()
    assert ref is not None and hasattr(ref, "sos_proc_")
    tmp = str(tmp_path)
    _installation(tmp)
    more = {}
    if case != "no gas":
        pc.write_ckd_files(os.environ["SOS_ABS_ROOT"], pc.ckd_tables(4))
        fprof = os.path.join(tmp, "profile.txt")
        tab._write_profile(fprof, tab._user())
        more = dict(wa_simu=1e4 / 13255.0, absprofil=0, ficabsprofil=fprof, nustep=10.0, imode_ckd_calcul=int(case[-1]))
    root, want = _driver(ref, tmp, **more)
    got = sos.sos_proc(solver=solver, resroot=os.path.join(tmp, "gpu"), trace=False, **{**ARGS, **more})
    assert len(got) == len(want) == 23
    n = want[0]
    assert got[0] == n and np.array_equal(got[1], want[1])
    assert np.array_equal(got[2], want[2]), "PHI_FIN"
    np.testing.assert_allclose(got[3], want[3], rtol=1e-13, atol=0, err_msg="THETA_FIN")
    a = fm.read_result_bin(os.path.join(tmp, "gpu", "SOS", "%.6f" % {**ARGS, **more}["wa_simu"], "SOS_Result.bin"), n)
    b = fm.read_result_bin(os.path.join(root, "SOS", "SOS_Result.bin"), n)
    # every SOS_AGGREGATE call after the first ends its file with one more record of zeros (both inputs at their end, :372-393):
    # the driver's file of K aggregated terms carries K - 1 of them after the longest series; the front end's file ends there
    nb = int(np.flatnonzero(b.reshape(b.shape[0], -1).any(axis=1))[-1]) + 1
    assert a.shape[0] >= nb and not a[nb:].any(), "number of Fourier orders: %d, reference %d" % (a.shape[0], nb)
    err = float(np.abs(a[:nb] - b[:nb]).max() / np.abs(b).max())
    assert err < 5e-6, err                                        # the aerosol thickness goes through REAL*4 Mie records (2e-7)
    worst = 0.0
    for t in range(14):
        g, w = np.asarray(got[4 + t]), np.asarray(want[4 + t])
        assert g.shape == w.shape == (361, 81) and not g[7:].any() and not g[:, n:].any()
        if t % 7 == 0:
            # acos at +-1: one ulp of the cosine is 8.5e-7 degrees (seen on the B200: the solar direction of the downward table)
            np.testing.assert_allclose(g, w, rtol=1e-12, atol=3e-6, err_msg="scattering angle")
        elif t % 7 in (1, 2, 3):                                  # I, Q, U against the largest radiance of the direction
            scale = float(np.abs(np.asarray(want[4 + 7 * (t // 7) + 1])).max())
            e = float(np.abs(g - w).max() / scale)
            assert e < 1e-5, ("Stokes table %d" % t, e)
            worst = max(worst, e)
    for k, name in zip(range(18, 23), ("Tdir", "Fdd", "Fd", "E+", "COEF_TRONCA")):
        assert abs(got[k] - want[k]) <= 1e-5 * max(abs(want[k]), 1e-3), (name, got[k], want[k])
    print("\n[sos_proc vs SOS_PROC] %s, 12 / 20 Gauss angles, 7 azimuths: SOS_Result.bin %d records within %.1e of scale; I Q U up and "
          "down within %.1e of scale; Tdir %.6f (%.6f) E+ %.6f (%.6f) COEF_TRONCA %.8f (%.8f)"
          % (case, nb, err, worst, got[18], want[18], got[21], want[21], got[22], want[22]))


def test_run_band_ckd_mode_2_host_flow(pkg, tmp_path):
    """band.run_band with ckd_mode = 2 on a stand-in for the device: per-term absorption profiles -> one estimated profile per
    wavelength -> SOS_PROFILE once per wavelength -> one term of weight 1 per wavelength; mode 1 keeps the CKD terms."""
    import profile_cases as pc
    from test_frontend import _StubSolver
    band = importlib.import_module("radiativetransfer-sos_b200.band")
    syn = pkg.synth

    class Stub(_StubSolver):
        def profile_chain(self, tables, userprofil, altabs, ro, terms, text_hop=True, want_tauabs=False):
            out = self.profile(altabs, None, terms)
            self.calls.pop()
            tau = np.array([[0.0] * 50 if t["absprofil"] == 7 else
                            [0.01 * (1 + sum(t["ik"]) % 5) * (i + 1) / (1.0 + 0.1 * t["lamb1"]) for i in range(50)] for t in terms])
            self.chain_tau = tau
            return out + (tau,)

        def profile(self, altabs, tau, terms, text_hop=True):
            self.calls.append(("profile", None if tau is None else np.array(tau), list(terms)))
            return super().profile(altabs, tau, terms, text_hop)

        def upload(self, wl, groups=None, ngroup=None):
            self.uploaded = (wl, list(groups), ngroup)
            return super().upload(wl, groups=groups, ngroup=ngroup)

    user, altabs, ro = pc.gas_atmosphere(6)
    t = pc.ckd_tables(6)
    counts = [int(np.prod(t["nexp"][:, l])) for l in range(pc.NWVL)]
    lambs = [l + 1 for l in np.argsort(counts) if 2 <= counts[l] <= 12][:2] + [int(np.argmin(counts)) + 1]
    waves = []
    for n, l in enumerate(lambs):
        o = syn.make_optics(nb_gauss=12, tetas=30.0, os_nb=24, surface="lambert", rho=0.1)
        waves.append(band.Wavelength(optics=o, lamb1=l, tr=0.08, ta=0.1, name="w%d" % n))
    waves[-1].absprofil = 7
    nterm = [counts[lambs[0] - 1], counts[lambs[1] - 1], 1]
    s1 = Stub()
    r1 = band.run_band(s1, t, t["ai"], user, altabs, ro, waves, itrphi=2, pas_phi=60, outdir=str(tmp_path / "m1"), flux=True, ckd_mode=1)
    assert r1.nterm == nterm and len(s1.uploaded[0].terms) == sum(nterm) and not [c for c in s1.calls if c[0] == "profile"]
    s2 = Stub()
    r2 = band.run_band(s2, t, t["ai"], user, altabs, ro, waves, itrphi=2, pas_phi=60, outdir=str(tmp_path / "m2"), flux=True, ckd_mode=2)
    assert r2.nterm == [1, 1, 1]
    wl, groups, ngroup = s2.uploaded
    assert groups == [0, 1, 2] and ngroup == 3 and [tm.aik for tm in wl.terms] == [1.0, 1.0, 1.0] and [tm.optics for tm in wl.terms] == [0, 1, 2]
    (call,) = [c for c in s2.calls if c[0] == "profile"]
    tau2, terms2 = call[1], call[2]
    assert tau2.shape == (3, 50) and [tm["lamb1"] for tm in terms2] == lambs and [tm["absprofil"] for tm in terms2] == [2, 2, 7]
    first = np.cumsum([0] + nterm)
    for w in range(2):
        _, aik = band.enumerate_ckd_terms(t["nexp"], t["ai"], lambs[w])
        trs = np.zeros(50)
        for a, row in zip(aik, s2.chain_tau[first[w]:first[w + 1]]):
            trs += a * np.exp(-row)
        assert np.array_equal(tau2[w], np.maximum(-np.log(trs), 0.0))
        assert tau2[w].min() >= s2.chain_tau[first[w]:first[w + 1]].min(axis=0).min() - 1e-15
    assert not tau2[2].any()
    for w in range(3):
        assert sorted(os.listdir(os.path.join(str(tmp_path / "m2"), "w%d" % w))) == ["SOS_Down.txt", "SOS_Flux.txt", "SOS_Result.bin", "SOS_Up.txt"]
    flux = open(os.path.join(str(tmp_path / "m2"), "w0", "SOS_Flux.txt")).read().split("\n")
    got = float(flux[-2].split()[3])                               # GOT at the surface = the estimated profile's last level (:3860)
    assert abs(got - tau2[0][-1]) < 5e-5
    with pytest.raises(ValueError, match="2515"):
        band.run_band(s2, t, t["ai"], user, altabs, ro, waves, ckd_mode=3)


from test_aerosol_chain import host  # noqa: E402,F401  (fixture: the aerosol device functions compiled for the host)

_LAND = dict(isurf=5, rho=0.0, k0_roujean=0.25, k1_roujean=0.04, k2_roujean=0.3, surf_ind=1.5)
_FRONT_END_CASES = {
    "ocean_wmo_view2": {},
    "ocean_view1_phi30_unpolarized": dict(itrphi=1, phios=30.0, ipolar=0, igmax=6),
    "land_roujean_breon_lnd": dict(_LAND, imod_aer=0, rn_wa=1.45, in_wa=-0.005, rn_waref=1.45, in_waref=-0.005, igranu=1,
                                   lnd_radius_mmd_aer=0.12, lnd_lnvar_mmd_aer=0.5, wa_simu=0.670),
    "lambert_flat_sea_no_aerosols": dict(isurf=2, rho=0.03, aot_ref=0.0, zout=3.0),
    "roujean_only_junge_no_truncation": dict(isurf=3, rho=0.0, k0_roujean=0.3, k1_roujean=0.05, k2_roujean=0.4, imod_aer=0, rn_wa=1.40, in_wa=-0.002,
                                             rn_waref=1.42, in_waref=-0.001, igranu=2, jd_slope_mmd_aer=4.0, jd_rmin_mmd_aer=0.05,
                                             jd_rmax_mmd_aer=10.0, itronc_aer=0, tetas=50.0),
    "rondeaux_shettle_fenn": dict(_LAND, isurf=4, imod_aer=2, imodele_sf=2, rh=70.0, wa_simu=0.865, psurf=950.0, hr=7.5, ha=1.5),
    "maignan_bimodal_volumes": dict(_LAND, isurf=7, coef_c_maignan=6.5, imod_aer=3, mode_param_bilnd=1, user_cv_coarse=0.02, user_cv_fine=0.01,
                                    bmd_cm_mrwa=1.50, bmd_cm_miwa=-0.003, bmd_cm_mrwaref=1.50, bmd_cm_miwaref=-0.003, bmd_cm_rmodal=0.8,
                                    bmd_cm_var=0.6, bmd_fm_mrwa=1.43, bmd_fm_miwa=-0.005, bmd_fm_mrwaref=1.44, bmd_fm_miwaref=-0.005,
                                    bmd_fm_rmodal=0.08, bmd_fm_var=0.45),
    "bimodal_share_aerosol_layer_given_rayleigh": dict(imod_aer=3, mode_param_bilnd=2, rtauct_waref=0.4, bmd_cm_mrwa=1.50, bmd_cm_miwa=-0.003,
                                                       bmd_cm_mrwaref=1.50, bmd_cm_miwaref=-0.003, bmd_cm_rmodal=0.8, bmd_cm_var=0.6,
                                                       bmd_fm_mrwa=1.43, bmd_fm_miwa=-0.005, bmd_fm_mrwaref=1.44, bmd_fm_miwaref=-0.005,
                                                       bmd_fm_rmodal=0.08, bmd_fm_var=0.45, iprofil=2, zmin=1.0, zmax=3.0, tr=0.05),
    "wmo_user_volumes": dict(imodele_wmo=4, c_wmo_dl=0.0, c_wmo_ws=0.5, c_wmo_oc=0.4, c_wmo_so=0.1, isurf=0, rho=0.1),
    "user_angle_files": dict(user_angles=[5.0, 20.5, 60.0], user_mie_angles=[10.0, 75.0]),
    "aerosol_and_surface_files_of_the_user": dict(from_files=True, wa_simu=0.55),
    "gas_mode1_lambert": dict(isurf=0, rho=0.2, gas=1),
    "gas_mode2_lambert": dict(isurf=0, rho=0.2, gas=2),
}


@pytest.mark.parametrize("name", list(_FRONT_END_CASES))
def test_front_end_host_logic_is_the_drivers(pkg, host, tmp_path, name):
    """sos.sos_proc -> frontend.run -> band.run_band with every device stage replaced by the reference's own routine
    (tests/reference_flow_solver.py) against SOS_PROC itself on the same arguments: the 23 outputs BIT FOR BIT, the aerosol file line
    by line, SOS_Result.bin record by record.  What is compared is the host side of the product: keyword mapping, reference-wavelength
    scaling of the optical thickness, the result-file hop of the coefficients, Rayleigh thickness, CKD term lists / weights / modes,
    aggregated and direct groups, direct-term models, output layout."""
    import profile_cases as pc
    import test_absprofile as tab
    from reference_flow_solver import ReferenceFlowSolver
    sos = importlib.import_module("radiativetransfer-sos_b200.sos")
    fm = importlib.import_module("radiativetransfer-sos_b200.formats")
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_proc_"):
        pytest.skip("oracle/_ref without SOS_PROC")
    tmp = str(tmp_path)
    _installation(tmp)
    more = dict(_FRONT_END_CASES[name])
    gas = more.pop("gas", None)
    user = more.pop("user_angles", [])
    for key, arg, fname in (("user_angles", "ficangles_user_lum", "rad_angles.txt"), ("user_mie_angles", "ficangles_user_mie", "mie_angles.txt")):
        angles = user if key == "user_angles" else more.pop(key, [])
        if angles:
            with open(os.path.join(tmp, fname), "w") as f:
                f.write("".join("%.2f\n" % a for a in angles))
            more[arg] = os.path.join(tmp, fname)
    if gas:
        pc.write_ckd_files(os.environ["SOS_ABS_ROOT"], pc.ckd_tables(4))
        fprof = os.path.join(tmp, "profile.txt")
        tab._write_profile(fprof, tab._user())
        more.update(wa_simu=1e4 / 13255.0, absprofil=0, ficabsprofil=fprof, nustep=10.0, imode_ckd_calcul=gas)
    if more.pop("from_files", False):
        # -AER.UserFile and -SURF.File: the aerosol file and the surface file a first run of the driver leaves behind, given back
        # to both sides as the user's files (SOS_PROC.F:1867-1876, 3186-3189)
        first, _ = _driver(ref, tmp, **more)
        import glob
        import shutil
        (fsurf,) = glob.glob(os.path.join(first, "SURF", "GLITTER", "*")) or glob.glob(os.path.join(first, "SURF", "*", "*"))
        shutil.copy(os.path.join(first, "SOS", "Aerosols.txt"), os.path.join(tmp, "user_aerosols.txt"))
        shutil.copy(fsurf, os.path.join(tmp, "user_surface.bin"))
        shutil.rmtree(first)
        more.update(ficuser_aer=os.path.join(tmp, "user_aerosols.txt"), ficsurf=os.path.join(tmp, "user_surface.bin"), ficgranu=None)
        for k in ("imod_aer", "imodele_wmo"):
            more[k] = None
    args = {k: v for k, v in {**ARGS, **more}.items() if v is not None}
    root, want = _driver(ref, tmp, **{k: v for k, v in more.items() if v is not None or k in ARGS})
    _, ga, _, _ = pkg.synth.sos_angles(args["nbmu_gauss_lum"], args["tetas"], user)
    s = ReferenceFlowSolver(host, ref, tmp_path, ga)
    got = sos.sos_proc(solver=s, resroot=os.path.join(tmp, "mine"), trace=False, **args)
    assert len(got) == len(want) == 23
    names = ("NBMU", "IND_ANGOUT", "PHI", "THETA") + tuple("%s_%s" % (t, d) for d in ("UP", "DOWN") for t in
                                                           ("SCA", "I", "Q", "U", "POL_ANG", "POL_RATE", "L_POL")) + ("TDIR", "FDD", "FD", "EPLUS", "COEF_TRONCA")
    for k, (g, w) in enumerate(zip(got, want)):
        if names[k] == "THETA":                                   # acos of the Gauss cosines: numpy's against the reference's DACOS
            np.testing.assert_allclose(g, w, rtol=1e-14, atol=0)
        elif names[k] in ("TDIR", "FDD", "FD"):                   # exp / division of identical inputs by the product's own writer
            assert abs(g - w) <= 2e-16 * max(1.0, abs(w)), (names[k], g, w)
        else:
            assert np.array_equal(np.asarray(g), np.asarray(w)), (name, names[k], np.abs(np.asarray(g) - np.asarray(w)).max())
    a = fm.read_result_bin(os.path.join(tmp, "mine", "SOS", "%.6f" % args["wa_simu"], "SOS_Result.bin"), want[0])
    b = fm.read_result_bin(os.path.join(root, "SOS", "SOS_Result.bin"), want[0])
    nb = int(np.flatnonzero(b.reshape(b.shape[0], -1).any(axis=1))[-1]) + 1
    assert a.shape[0] == nb and np.array_equal(a, b[:nb])
    if "ficuser_aer" in args:                                     # no aerosol computation, no aerosol file on either side
        assert not os.path.exists(os.path.join(root, "SOS", "Aerosols.txt")) and not os.path.isdir(os.path.join(tmp, "mine", "AER"))
    elif args["aot_ref"] > 0.0:
        la = open(os.path.join(root, "SOS", "Aerosols.txt")).read().split("\n")
        lb = open(os.path.join(tmp, "mine", "AER", "Aerosols.txt")).read().split("\n")
        assert la == lb, [(x, y) for x, y in zip(la, lb) if x != y][:3]
    print("\n[front end = driver] %s: 23 outputs bit for bit, SOS_Result.bin %d records (+ %d of zeros in the driver's file), E+ %.8f"
          % (name, nb, b.shape[0] - nb, want[21]))
