"""Keyword-driven front end: the keyword set of SOS_ABS_MAIN and what frontend.py derives from it on the host
(CPU tests), and the whole run from keywords to SOS_Up / SOS_Down on the device against the reference's flow stage by stage
(GPU test: SOS_INIT_PARAMWMO -> SOS_MIE -> SOS_GRANU -> mixture -> SOS_DECOMPO_LEGENDRE -> aerosol result file -> SOS_GLITTER ->
SOS_PROFILE -> SOS + SOS_OS -> SOS_AGGREGATE -> SOS_TRPHI_OPTION of oracle/_ref/libsosref.so)."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

import aerosol_cases as ac
import refdirect

DEMO = ("-SOS_Main.Wa 0.910 -SOS_Main.ResRoot {root} -SOS_Main.Log SOS_Main_Demo.Log -ANG.Rad.NbGauss {nrad} -ANG.Aer.NbGauss {naer} "
        "-ANG.Thetas 35. -SOS.View 1 -SOS.View.Phi 0. -SOS.ResFileUp SOS_Up_Demo.txt -SOS.ResFileDown SOS_Down_Demo.txt -AP.Log Profile_Demo.Log "
        "-AP.Psurf 1013 -AP.AerProfile.Type 1 -AP.HR 8.0 -AP.AerHS.HA 2.0 -AP.AbsProfile.Type {abs} -AP.SpectralResol 10. -SOS.AbsModeCKD 1 "
        "-AER.DirMie {root}/MIE -AER.Model 1 -AER.WMO.Model 2 -AER.Waref 0.550 -AER.AOTref 0.3 -AER.ResFile Aerosols_Demo.txt "
        "-AER.Log Aerosols_Demo.Log -SURF.Dir {root}/SURF -SURF.Type 1 -SURF.Alb 0.00 -SURF.Ind 1.34 -SURF.Glitter.Wind 2.0")


def _mods():
    m = lambda n: importlib.import_module("radiativetransfer-sos_b200." + n)
    return m("keywords"), m("frontend"), m("aerosols")


def test_keyword_set_is_the_references():
    """Every keyword SOS_ABS_MAIN documents (SOS_ABS_MAIN.F:213-912) is known with the documented value type; the demo script's
    command line parses; unknown keywords and missing required ones are refused."""
    kw, fe, _ = _mods()
    src = "/root/reference/src/SOS_ABS_MAIN.F"
    if os.path.exists(src):
        lines = open(src, encoding="latin-1").read().split("\n")[212:915]
        cur, seen = None, {}
        for ln in lines:
            m = re.match(r"^C\s+Keyword\s*:\s*(-[\w.]+)", ln)
            if m:
                cur = m.group(1)
            elif cur and "Value format" in ln and cur not in seen:
                seen[cur] = "f" if "Float" in ln else "i" if "Integer" in ln else "s"
        assert len(seen) >= 90
        seen["-AER.SF.RH"] = seen.pop("-AER.SF.HR")                 # the documentation block misspells the keyword the code parses
        for k, t in seen.items():
            assert k in kw.KEYWORDS, k
            if k != "-AER.WMO.WS":                                  # documented as a string by a copy-paste slip, read as a float
                assert kw.KEYWORDS[k] == t, (k, t)
        parsed = set(re.findall(r'KEYWORD\.EQ\."(-[\w.]+)"', open(src, encoding="latin-1").read()))
        assert parsed == set(kw.KEYWORDS), (parsed ^ set(kw.KEYWORDS))   # exactly the names the argument loop compares with
        demo = open("/root/reference/exe/runSOS-ABS_demo.ksh", encoding="latin-1").read()
        for k in re.findall(r"(-[A-Z_a-z]+\.[\w.]+)\s", demo[demo.index("SOS_ABS_MAIN.exe"):]):
            assert k in kw.KEYWORDS, k
    d = kw.parse(DEMO.format(root="/tmp/x", nrad=40, naer=40, abs=1).split())
    assert d["-SOS_Main.Wa"] == 0.910 and d["-ANG.Rad.NbGauss"] == 40 and d["-SURF.Type"] == 1 and d["-SOS.IGmax"] == 100
    assert d["-SOS.ResBin"] == "SOS_Result.bin" and d["-SOS.Ipolar"] == 1 and d["-SOS.OutputAlt"] == -1.0
    assert kw.parse(["-AP.MOT", "1.D-2"], require=False)["-AP.MOT"] == 0.01
    with pytest.raises(ValueError):
        kw.parse(["-SOS.Wavelength", "0.5"], require=False)
    with pytest.raises(ValueError):
        kw.parse(["-SOS_Main.Wa", "0.5"])
    with pytest.raises(ValueError):
        kw.parse(["-SOS.IGmax", "many"], require=False)


def test_frontend_host_logic():
    kw, fe, aer = _mods()
    assert fe.expansion_orders(None, None) == (80, 48, 128)           # SOS_ANGLES.F:303-329
    assert fe.expansion_orders(40, 40) == (80, 80, 160) and fe.expansion_orders(20, None) == (40, 48, 128)
    n, xmu, xhr = fe.mie_angles(20)
    assert n == 20 and xmu[20] == 0.0 and (np.diff(xmu) > 0).all() and abs(xhr[21:].sum() - 1.0) < 1e-13
    syn = importlib.import_module("radiativetransfer-sos_b200.synth")
    mu, w = syn.sos_gauss(21)                                       # equal to the reference's SOS_GAUSS (test_oracle_vs_reference.py)
    assert np.array_equal(xmu[21:], [float("%.13E" % v) for v in np.sort(mu)])
    tr = fe.rayleigh_thickness(1013.0, 0.910)
    assert abs(tr - 1e-4 * (84.35 / 0.91 ** 4 - 1.225 / 0.91 ** 5 + 1.4 / 0.91 ** 6)) < 1e-9 and 0.012 < tr < 0.0125
    d = kw.parse(DEMO.format(root="/tmp/x", nrad=12, naer=20, abs=7).split())
    os.environ["SOS_ABS_ROOT"] = "/somewhere"
    m = fe.aerosol_model(d)
    assert isinstance(m, aer.Wmo) and m.imodele == 2 and m.datafile == "/somewhere/fic/Data_WMO_cor_2015_12_16"
    d.update({"-AER.Model": 3, "-AER.BMD.VCdef": 2, "-AER.BMD.RAOT": 0.4, "-AER.BMD.CM.MRwa": 1.45, "-AER.BMD.CM.MIwa": -0.004,
              "-AER.BMD.CM.MRwaref": 1.46, "-AER.BMD.CM.MIwaref": -0.005, "-AER.BMD.CM.SDradius": 0.4, "-AER.BMD.CM.SDvar": 0.6,
              "-AER.BMD.FM.MRwa": 1.42, "-AER.BMD.FM.MIwa": -0.008, "-AER.BMD.FM.MRwaref": 1.43, "-AER.BMD.FM.MIwaref": -0.009,
              "-AER.BMD.FM.SDradius": 0.08, "-AER.BMD.FM.SDvar": 0.45})
    b = fe.aerosol_model(d)
    assert b.rtauct == 0.4 and b.coarse_rn(0.55) == 1.46 and b.coarse_rn(0.91) == 1.45 and b.fine_in(0.55) == -0.009
    # the indices of the reference wavelength: used at that wavelength whatever the definition of the mixture (SOS_PROC.F:2896-2907),
    # ignored when the simulation wavelength IS the reference one (:1815-1822), required otherwise for VCdef 2 (error 2329)
    b1 = fe.aerosol_model(dict(d, **{"-AER.BMD.VCdef": 1, "-AER.BMD.CoarseVC": 0.3, "-AER.BMD.FineVC": 0.7}), [0.91])
    assert b1.rtauct is None and b1.cv_coarse == 0.3 and b1.coarse_rn(0.55) == 1.46 and b1.coarse_rn(0.91) == 1.45
    b2 = fe.aerosol_model(d, [0.55])
    assert b2.coarse_rn == 1.45 and b2.fine_in == -0.008
    with pytest.raises(ValueError, match="2329"):
        fe.aerosol_model({k: v for k, v in d.items() if k != "-AER.BMD.FM.MIwaref"}, [0.91])
    mono = {"-AER.Model": 0, "-AER.Waref": 0.55, "-AER.MMD.MRwa": 1.45, "-AER.MMD.MIwa": -0.004, "-AER.MMD.SDtype": 1, "-AER.MMD.LNDradius": 0.1,
            "-AER.MMD.LNDvar": 0.46}
    assert fe.aerosol_model(mono, [0.55]).rn == 1.45                  # simulation at the reference wavelength: one index
    with pytest.raises(ValueError, match="2314"):
        fe.aerosol_model(mono, [0.91])
    mm = fe.aerosol_model(dict(mono, **{"-AER.MMD.MRwaref": 1.47, "-AER.MMD.MIwaref": -0.006}), [0.91])
    assert mm.rn(0.55) == 1.47 and mm.rn(0.91) == 1.45 and mm.in_(0.55) == -0.006 and mm.in_(0.91) == -0.004
    d.update({"-AER.Model": 2, "-AER.SF.Model": 3, "-AER.SF.RH": 70.0})
    sf = fe.aerosol_model(d)
    assert isinstance(sf, aer.ShettleFenn) and sf.imodele == 3 and sf.rh == 70.0 and sf.dirfic == "/somewhere/fic"
    d["-AER.Model"] = 6
    with pytest.raises(ValueError):
        fe.aerosol_model(d)
    os.environ["SOS_ABS_ROOT"] = "/nonexistent"                      # a predefined atmosphere without the reference's installation
    with pytest.raises(ValueError, match="SOS_SUB_TRS.F"):
        fe.run_keywords(None, DEMO.format(root="/tmp/x", nrad=12, naer=20, abs=1).split())


def _parse_updown(path):
    rows = []
    for ln in open(path):
        p = ln.split()
        try:
            rows.append([float(x) for x in p])
        except ValueError:
            continue
    n = max(len(r) for r in rows)
    return np.array([r for r in rows if len(r) == n])


@pytest.mark.gpu
def test_gpu_run_from_keywords_demo(pkg, solver, tmp_path):
    """The demo's command line (exe/runSOS-ABS_demo.ksh: WMO maritime aerosols scaled from 0.550 microns, rough-sea surface, solar
    plane) with 12 / 20 Gauss angles and without gaseous absorption, from keywords to SOS_Up_Demo.txt / SOS_Down_Demo.txt on the
    device -- against the reference's routines run one after the other on the host."""
    kwm, fe, aer = _mods()
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    import test_aerosol_chain as tac
    ref = refdirect.lib()
    assert ref is not None and hasattr(ref, "sos_mie_")
    syn, fm = pkg.synth, pkg.formats
    root = str(tmp_path / "res")
    os.makedirs(os.path.join(str(tmp_path), "abs_root", "fic"))
    os.environ["SOS_ABS_ROOT"] = os.path.join(str(tmp_path), "abs_root")
    wmo = ac.write_wmo_file(os.path.join(os.environ["SOS_ABS_ROOT"], "fic", "Data_WMO_cor_2015_12_16"))
    argv = DEMO.format(root=root, nrad=12, naer=20, abs=7).split()
    res, aopt = fe.run_keywords(solver, argv)
    d = res.dirs[0]
    assert sorted(os.listdir(d)) == ["SOS_Down_Demo.txt", "SOS_Result.bin", "SOS_Up_Demo.txt"]
    assert os.path.exists(os.path.join(root, "AER", "Aerosols_Demo.txt"))
    # ---- the reference's flow ----
    wa, waref, aot, os_nb, os_ns, os_nm = 0.910, 0.550, 0.3, 40, 24, 64
    assert fe.expansion_orders(20, 12) == (os_nb, os_ns, os_nm)
    nbm, xmu, xhr = fe.mie_angles(20)
    k1 = {}
    for w in (wa, waref):
        e, v1, v2, mr, mi, vol = ac.ref_wmo_params(ref, wmo, w)
        comps = [(mr[i], mi[i], 0.0001, (4000.0, 50.0, 800.0, 10.0)[i], 1, v1[i], v2[i], -999.0, w) for i in (1, 2)]
        n = np.array([0.0 / vol[0], np.float64(np.float32(0.05)) / vol[1], np.float64(np.float32(0.95)) / vol[2], 0.0 / vol[3]])   # REAL*4 literals
        ntot = 0.0
        for x in n:
            ntot = ntot + x
        k1[w] = tac._reference_chain(ref, str(tmp_path), nbm, xmu, xhr, comps, [(2, [0, 1], [n[1] / ntot, n[2] / ntot], 1)], os_nb)[2][0]
    dd = k1[wa]
    ta = dd["kmat1"] / k1[waref]["kmat1"] * aot
    piz = dd["kmat2"] / dd["kmat1"]
    ct = dd["coef_tronca"]
    fa = str(tmp_path / "Aer_ref.txt")
    api.write_aerosols(fa, os_nb, dd["kmat1"], dd["kmat2"], ct / 2.0 + (1.0 - ct / 2.0) * dd["beta11"][1] / 3.0, ct,
                       piz * (1.0 - ct / 2.0) / (1.0 - piz * ct / 2.0), dd["alp"], dd["beta11"], dd["gamma12"], dd["zeta"])
    la, lb = open(os.path.join(root, "AER", "Aerosols_Demo.txt")).read().split("\n"), open(fa).read().split("\n")
    same_aer = sum(x == y for x, y in zip(la, lb))
    assert la[:8] == lb[:8] and same_aer >= len(lb) - 2
    r8 = lambda v: np.array([float("%.7E" % x) for x in v])
    rmu, ga, n0, _ = syn.sos_angles(12, 35.0)
    N = (rmu.size - 1) // 2
    surf_ref = refdirect.glitter(ref, fm, str(tmp_path), N, rmu, ga, 2.0, 1.34, os_nb, os_ns, os_nm)
    o = syn.Optics(nbmu=N, rmu=rmu.copy(), ga=ga, n0=n0, tetas=35.0, os_nb=os_nb, alpha=r8(dd["alp"]), beta=r8(dd["beta11"]),
                   gamma=r8(dd["gamma12"]), zeta=r8(dd["zeta"]), a_trunc=float("%.5f" % ct),
                   piztr=float("%.5f" % (piz * (1.0 - ct / 2.0) / (1.0 - piz * ct / 2.0))), rho=0.0, imat_surf=1, igli=1, surf=surf_ref,
                   ind_surf=1.34, wind=2.0, igmax=100, ipolar=1, zout=-1.0)
    tr = fe.rayleigh_thickness(1013.0, wa)
    term = dict(lamb1=1, ik=(1,) * 8, absprofil=7, iprofil=1, tr=tr, hr=8.0, ta=ta, ha=2.0, zmin=0.0, zmax=0.0)
    ier_r, nt_r, _, z_r, h_r, pa_r, pm_r = refdirect.profile(ref, str(tmp_path), np.linspace(120.0, 0.0, 50), np.zeros(50), term)
    assert ier_r == 0 and nt_r == int(res.nt[0])
    wl_ref = syn.Workload("ref")
    wl_ref.optics.append(o)
    wl_ref.terms.append(syn.Term(0, 1.0, z_r, h_r, pa_r, pm_r))
    rr = refdirect.runner()
    r, _, _ = rr.solve_terms(wl_ref, [0], 1)
    agg_rec, agg_sc = rr.aggregate_point(ref, fm, str(tmp_path), N, [(1.0, r[0])])
    nr = int(res.groups.n_rec[0])
    assert not agg_rec[nr:].any(), "number of Fourier orders differs"
    got = fm.read_result_bin(os.path.join(d, "SOS_Result.bin"), N)
    scale = np.abs(agg_rec[:nr]).max()
    err = np.abs(got - agg_rec[:nr]).max() / scale
    assert err < 5e-6, err                                        # aerosol optical thickness through REAL*4 Mie records (2e-7 on TA)
    nphi, pf, th, up0, dn0 = refdirect.trphi_option(ref, fm, str(tmp_path), agg_rec[:nr], N, o.rmu, o.ga, agg_sc["ttot_tronc"],
                                                    agg_sc["tauout"], 1, o.n0, 2.0, 1.34, 0, 1, 0.0, 30)
    fu, fd = str(tmp_path / "REF_Up.txt"), str(tmp_path / "REF_Down.txt")
    api.write_updown(fu, fd, N, 1, 0.0, 30, -1.0, pf, th, up0, dn0)
    nlines = nsame = 0
    for mine, theirs in ((os.path.join(d, "SOS_Up_Demo.txt"), fu), (os.path.join(d, "SOS_Down_Demo.txt"), fd)):
        a, b = open(mine).read().split("\n"), open(theirs).read().split("\n")
        assert len(a) == len(b)
        nlines += len(a)
        nsame += sum(x == y for x, y in zip(a, b))
        pa_, pb_ = _parse_updown(mine), _parse_updown(theirs)
        assert pa_.shape == pb_.shape
        np.testing.assert_allclose(pa_[:, 3:6], pb_[:, 3:6], rtol=2e-5, atol=1e-9)
    print("\n[keywords -> files] demo command line (12 / 20 Gauss angles, no gas): NT %d = reference's, Fourier orders %d = reference's, "
          "SOS_Result.bin within %.1e of scale, aerosol file lines identical %d / %d, SOS_Up/Down lines identical %d / %d; TA = %.6f"
          % (nt_r, nr, err, same_aer, len(lb), nsame, nlines, ta))


def test_sos_proc_arguments():
    """sos.sos_proc takes the arguments of the f2py wrapper binding/run_sos.py calls (names, order, "not defined" values)."""
    sos = importlib.import_module("radiativetransfer-sos_b200.sos")
    names = [a for a, _ in sos.ARGS]
    src = "/root/reference/src/SOS_PROC.F"
    if os.path.exists(src):
        txt = open(src, encoding="latin-1").read()
        ins = re.findall(r"^Cf2py intent\(in\)\s+(.*)$", txt, re.M)
        want = [w.strip().lower() for ln in ins for w in ln.split(",") if w.strip()]
        want = [sos._ALIASES.get(w, w) for w in want]
        assert want == names, [(a, b) for a, b in zip(want, names) if a != b]
        call = open("/root/reference/binding/run_sos.py", encoding="latin-1").read()
        used = re.findall(r"\b(\w+)=[A-Za-z_]+[,)]", call[call.index("sos.sos_proc("):call.index("print(i_up.shape)")])
        for u in used:
            assert sos._ALIASES.get(u, u) in names, u
    kw = sos.to_keywords(resroot="/tmp/r", wa_simu=0.44, tetas=40.0, nbmu_gauss_lum=24, nbmu_gauss_mie=-999, ficangles_user_lum="NO_USER_ANGLES",
                         ficanglog="Angles.Log", waref_aot=0.55, aot_ref=0.3, itronc_aer=1, imod_aer=1, imodele_wmo=1, c_wmo_dl=-999.0,
                         tr=0.23, hr=8, ha=2, iprofil=1, absprofil=7, isurf=1, surf_ind=1.33, wind=2.0, rho=0.0, igmax=30, itrphi=1, phios=35,
                         pas_phi=-999, zout=-999.0, ier=0, trace=True)
    assert kw["-SOS_Main.Wa"] == 0.44 and kw["-ANG.Rad.NbGauss"] == 24 and "-ANG.Aer.NbGauss" not in kw and "-ANG.Rad.UserAngFile" not in kw
    assert "-ANG.Log" not in kw and "-AER.WMO.DL" not in kw and kw["-SOS.IGmax"] == 30 and kw["-SOS.OutputAlt"] == -1.0 and kw["-AP.HR"] == 8.0
    with pytest.raises(TypeError):
        sos.to_keywords(wavelength=0.5)
    with pytest.raises(ValueError):
        sos.to_keywords(resroot="/tmp/r", wa_simu=0.44)             # required arguments not defined


@pytest.mark.gpu
def test_gpu_sos_proc_entry(solver, tmp_path):
    """The f2py-shaped entry end to end: same numbers as frontend.run_keywords for the same description, in the wrapper's shapes."""
    sos = importlib.import_module("radiativetransfer-sos_b200.sos")
    _, fe, _ = _mods()
    os.makedirs(os.path.join(str(tmp_path), "abs_root", "fic"))
    os.environ["SOS_ABS_ROOT"] = os.path.join(str(tmp_path), "abs_root")
    ac.write_wmo_file(os.path.join(os.environ["SOS_ABS_ROOT"], "fic", "Data_WMO_cor_2015_12_16"))
    out = sos.sos_proc(solver=solver, resroot=str(tmp_path / "a"), wa_simu=0.910, tetas=35.0, nbmu_gauss_lum=12, nbmu_gauss_mie=20, waref_aot=0.55,
                       aot_ref=0.3, itronc_aer=1, imod_aer=1, imodele_wmo=2, hr=8.0, ha=2.0, iprofil=1, psurf=1013.0, absprofil=7, isurf=1,
                       surf_ind=1.34, wind=2.0, rho=0.0, itrphi=2, pas_phi=60, igmax=-999, zout=-999.0, ier=0, trace=False)
    assert len(out) == 23
    n, ind, phi, vza, tabs, (tdir, fdd, fd, eplus, ct) = out[0], out[1], out[2], out[3], out[4:18], out[18:]
    res, aer = fe.run_keywords(solver, DEMO.format(root=str(tmp_path / "b"), nrad=12, naer=20, abs=7).replace("-SOS.View 1 -SOS.View.Phi 0.",
                                                                                                                "-SOS.View 2 -SOS.View.Dphi 60").split())
    assert n == res.up.shape[3] == 13 and ind.shape == (81,) and not ind.any() and phi.shape == (361,) and vza.shape == (81,) and tabs[0].shape == (361, 81)
    assert list(phi[:7]) == [0.0, 60.0, 120.0, 180.0, 240.0, 300.0, 360.0] and res.nphi == 7
    for t in range(7):
        assert np.allclose(tabs[t][:7, :n], res.up[0, t, :7, :n], rtol=1e-12, atol=0) and np.allclose(tabs[7 + t][:7, :n], res.down[0, t, :7, :n], rtol=1e-12, atol=0)
        assert not tabs[t][7:].any() and not tabs[t][:, n:].any()
    assert abs(ct - aer[0].coef_tronca) < 1e-14 and abs(eplus - float(res.groups.eplus[0])) < 1e-14 and 0.0 < tdir < 1.0 and abs(fd - (fdd + tdir)) < 1e-12


def test_user_angle_files(tmp_path):
    """-ANG.Rad.UserAngFile / -ANG.Aer.UserAngFile and the UserAng output files, host side: the file reader, the merged angle sets and
    the selection of the records of the user angles in the order SOS_ABS_MAIN writes them (both view modes)."""
    kw, fe, aer = _mods()
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    syn = importlib.import_module("radiativetransfer-sos_b200.synth")
    f = tmp_path / "ang.txt"
    f.write_text("5.\n 20.5\n\n60.D0\n")
    assert fe.read_user_angles(str(f)) == [5.0, 20.5, 60.0]
    (tmp_path / "bad.txt").write_text("95.\n")
    with pytest.raises(ValueError):
        fe.read_user_angles(str(tmp_path / "bad.txt"))
    n, xmu, xhr = fe.mie_angles(10, [5.0, 20.5, 60.0])
    assert n == 13 and (np.diff(xmu) > 0).all() and (xhr[n + 1:] == 0).sum() == 3 and abs(xhr[n + 1:].sum() - 1.0) < 1e-13
    assert np.isin([float("%.13E" % np.cos(np.radians(a))) for a in (5.0, 20.5, 60.0)], xmu).all()
    rmu, ga, n0, flags = syn.sos_angles(6, 35.0, [5.0, 20.5, 60.0])
    N = (rmu.size - 1) // 2
    assert N == 10 and flags.sum() == 3 and flags[n0 - 1] == 0
    theta = np.degrees(np.arccos(rmu[N + 1:]))
    for itrphi, nphi, phis in ((1, 2, np.array([0.0, 180.0])), (2, 4, np.arange(0.0, 361.0, 120.0))):
        up = np.arange(7 * nphi * N, dtype=float).reshape(7, nphi, N) * 1e-3 + 0.1
        fu, fd = str(tmp_path / ("u%d.txt" % itrphi)), str(tmp_path / ("d%d.txt" % itrphi))
        api.write_updown(fu, fd, N, itrphi, 0.0, 120, -1.0, phis, theta, up, up + 1.0)
        out = str(tmp_path / ("user%d.txt" % itrphi))
        fe.user_angle_file(fu, out, itrphi, N, flags)
        full = [ln for ln in open(fu).read().split("\n") if ln]
        got = [ln for ln in open(out).read().split("\n") if ln]
        head = [ln for ln in full if ln.lstrip().startswith("#")]
        rows = [ln for ln in got if not ln.lstrip().startswith("#")]
        assert got[:len(head)] == head and len(rows) == (2 if itrphi == 1 else nphi) * 3
        col = 0 if itrphi == 1 else 1
        ang = sorted({round(abs(float(r.split()[col])), 2) for r in rows})
        assert ang == [5.0, 20.5, 60.0]
        assert all(r in full for r in rows)
    with pytest.raises(ValueError):                                  # UserAng output without a user angle file
        fe.run(None, kw.parse((DEMO.format(root="/tmp/x", nrad=12, naer=20, abs=7) + " -SOS.ResFileUp.UserAng U.txt").split()))


@pytest.mark.gpu
def test_gpu_user_angles_do_not_change_the_field(solver, tmp_path):
    """User angles have quadrature weight 0: a run with user angle files gives, at the Gauss angles, the radiances of the run without
    them -- to the level of the stop tests (1e-5, inc/SOS.h:389-400), not to rounding, because the stop tests of SOS_OS take their
    maxima over ALL angles, user angles included, so the two runs may stop at different scattering / Fourier orders -- and the
    UserAng files hold exactly the user-angle records of the full files."""
    _, fe, _ = _mods()
    os.makedirs(os.path.join(str(tmp_path), "abs_root", "fic"))
    os.environ["SOS_ABS_ROOT"] = os.path.join(str(tmp_path), "abs_root")
    ac.write_wmo_file(os.path.join(os.environ["SOS_ABS_ROOT"], "fic", "Data_WMO_cor_2015_12_16"))
    ua = tmp_path / "user.txt"
    ua.write_text("10.\n40.\n")
    base = DEMO.format(root=str(tmp_path / "a"), nrad=12, naer=20, abs=7)
    res0, _ = fe.run_keywords(solver, base.split())
    # (radiance angles only: a user angle among the phase-function angles may become the anchor of the truncation line,
    # SOS_AEROSOLS.F:4036-4052, and then changes the aerosol coefficients -- in the reference too)
    extra = " -ANG.Rad.UserAngFile %s -SOS.ResFileUp.UserAng Up_user.txt -SOS.ResFileDown.UserAng Down_user.txt" % ua
    res1, _ = fe.run_keywords(solver, (base.replace(str(tmp_path / "a"), str(tmp_path / "b")) + extra).split())
    keep = np.flatnonzero(res1.ind_angout == 0)
    assert res1.ind_angout.sum() == 2 and keep.size == res0.up.shape[3] == 13 and res1.up.shape[3] == 15
    for a, b in ((res0.up, res1.up), (res0.down, res1.down)):
        for t in (1, 2, 3):                                          # I, Q, U
            x, y = a[0, t, :2, :13], b[0, t, :2][:, keep]
            assert np.abs(x - y).max() <= 1e-4 * np.abs(a[0, 1]).max(), t
    d = res1.dirs[0]
    for full, user in (("SOS_Up_Demo.txt", "Up_user.txt"), ("SOS_Down_Demo.txt", "Down_user.txt")):
        lf = open(os.path.join(d, full)).read().split("\n")
        rows = [ln for ln in open(os.path.join(d, user)).read().split("\n") if ln and not ln.lstrip().startswith("#")]
        assert len(rows) == 4 and all(r in lf for r in rows)
        assert sorted({round(abs(float(r.split()[0])), 2) for r in rows}) == [10.0, 40.0]


def test_user_angles_do_not_change_the_reference_field(tmp_path):
    """The property the GPU test above relies on, on the reference itself (oracle/_ref/libsosref.so, CPU): SOS_GLITTER ->
    SOS_PROFILE -> SOS + SOS_OS -> SOS_AGGREGATE -> SOS_TRPHI_OPTION with and without two user angles (weight 0) give the same
    radiances at the Gauss angles and at the solar angle, with the same number of Fourier orders."""
    _, fe, _ = _mods()
    pkg = importlib.import_module("radiativetransfer-sos_b200")
    syn, fm = pkg.synth, pkg.formats
    ref = refdirect.lib()
    if ref is None or not hasattr(ref, "sos_glitter_"):
        pytest.skip("oracle/_ref/libsosref.so not built")
    tmp = str(tmp_path)
    os_nb, os_ns, os_nm = fe.expansion_orders(20, 12)
    k = np.arange(os_nb + 1)
    beta = (2 * k + 1) * 0.75 ** k                                   # Henyey-Greenstein-like expansion, mildly polarizing
    gamma, alpha, zeta = np.where(k >= 2, -0.05 * beta, 0.0), np.where(k >= 2, 0.6 * beta, 0.0), np.where(k >= 2, 0.55 * beta, 0.0)
    term = dict(lamb1=1, ik=(1,) * 8, absprofil=7, iprofil=1, tr=fe.rayleigh_thickness(1013.0, 0.910), hr=8.0, ta=0.25, ha=2.0, zmin=0.0, zmax=0.0)
    ier, nt, _, z, h, pa, pm = refdirect.profile(ref, tmp, np.linspace(120.0, 0.0, 50), np.zeros(50), term)
    assert ier == 0
    rr = refdirect.runner()
    out = {}
    for name, user in (("plain", []), ("user", [10.0, 40.0])):
        rmu, ga, n0, flags = syn.sos_angles(12, 35.0, user)
        N = (rmu.size - 1) // 2
        surf = refdirect.glitter(ref, fm, tmp, N, rmu, ga, 2.0, 1.34, os_nb, os_ns, os_nm)
        o = syn.Optics(nbmu=N, rmu=rmu.copy(), ga=ga, n0=n0, tetas=35.0, os_nb=os_nb, alpha=alpha, beta=beta, gamma=gamma, zeta=zeta,
                       a_trunc=0.4, piztr=0.95, rho=0.0, imat_surf=1, igli=1, surf=surf, ind_surf=1.34, wind=2.0, igmax=100, ipolar=1, zout=-1.0)
        wl = syn.Workload("ref")
        wl.optics.append(o)
        wl.terms.append(syn.Term(0, 1.0, z, h, pa, pm))
        r, _, _ = rr.solve_terms(wl, [0], 1)
        rec, sc = rr.aggregate_point(ref, fm, tmp, N, [(1.0, r[0])])
        nr = int(np.flatnonzero(np.abs(rec).reshape(rec.shape[0], -1).max(axis=1))[-1]) + 1
        _, _, _, up, down = refdirect.trphi_option(ref, fm, tmp, rec[:nr], N, o.rmu, o.ga, sc["ttot_tronc"], sc["tauout"], 1, o.n0, 2.0, 1.34,
                                                   0, 1, 0.0, 30)
        out[name] = (N, nr, flags, up, down)
    keep = np.flatnonzero(out["user"][2] == 0)
    assert out["plain"][0] == 13 and out["user"][0] == 15 and keep.size == 13 and out["plain"][1] == out["user"][1]
    for a, b in ((out["plain"][3], out["user"][3]), (out["plain"][4], out["user"][4])):
        for t in (1, 2, 3):
            assert np.abs(a[t, :2, :13] - b[t, :2][:, keep]).max() <= 1e-9 * np.abs(a[1]).max(), t


class _StubSolver:
    """Stands in for api.Solver in the CPU test of the host-side flow below: same method names, argument lists and array shapes,
    values that depend only on the inputs (radiances are functions of the cosine of the angle, so zero-weight angles cannot change
    the values at the others).  TEST INFRASTRUCTURE; nothing here computes radiative transfer."""
    def __init__(self):
        self.calls = []

    def aerosols(self, nbmu, xmu, xhr, components, models, os_nb, want_phase=True):
        self.calls.append(("aerosols", len(components), len(models)))
        nc, nm = len(components), len(models)
        ck = np.array([[2.0 / c[8] + 0.1 * c[0], 1.8 / c[8], 1.0] for c in components]).reshape(nc, 3)
        scal, coef = np.zeros((nm, 8)), np.zeros((nm, 6, os_nb + 1))
        for m, (ncomp, idx, w, itronc) in enumerate(models):
            k1 = ck[idx[0], 0] if ncomp == 0 else sum(wi * ck[i, 0] for wi, i in zip(w, idx))
            scal[m] = [k1, 0.9 * k1, 0.9, 0.88, 0.3 * itronc, 0.7, 1.0, itronc]
            coef[m, 1] = 0.8 ** np.arange(os_nb + 1) * (2 * np.arange(os_nb + 1) + 1)
        return dict(comp_k=ck, comp_phase=None, comp_ier=np.zeros(nc, np.int32), scal=scal, coef=coef, phase=None, model_ier=np.zeros(nm, np.int32))

    def decompo_legendre(self, itronc, nbmu, xmu, xhr, os_nb, p11, p12, p22, p33):
        self.calls.append(("decompo", nbmu, os_nb, len(p11)))
        z, k = np.zeros(os_nb + 1), np.arange(os_nb + 1)
        b = 0.8 ** k * (2 * k + 1)
        return dict(alp=z, beta11=b, beta22=b, gamma12=z, delta33=z, zeta=z, p11=p11, ttt=p11, coef_tronca=0.3, z1=1.0, itronc=itronc, ier=0)

    def glitter(self, nbmu, rmu, ga, wind, ind, os_nb, os_ns, os_nm):
        self.calls.append(("glitter", nbmu, os_nb, os_ns, os_nm))
        return np.zeros((os_nb + 1, 9, nbmu, nbmu), np.float32), np.zeros(nbmu * (nbmu + 1) // 2, np.int32)

    def roujean(self, nbmu, rmu, os_nb, k0, k1, k2):
        return np.zeros((os_nb + 1, 9, nbmu, nbmu), np.float32)

    def surface_bpdf(self, isurf, nbmu, rmu, ga, ind, os_nb, os_ns, os_nm, coef_c=0.0):
        return np.zeros((os_nb + 1, 9, nbmu, nbmu), np.float32)

    def bpdf_ajout_brdf(self, a, b):
        return a + b

    def set_direct_models(self, **kw):
        self.calls.append(("direct", sorted(kw)))

    def profile(self, altabs, tau, terms, text_hop=True):
        n = len(terms)
        nt = np.full(n, 12, np.int32)
        z = np.tile(np.linspace(120.0, 0.0, 601), (n, 1))
        h = np.tile(np.linspace(0.0, 0.3, 601), (n, 1))
        return nt, z, h, np.full((n, 601), 0.5), np.full((n, 601), 0.5), np.zeros(n, np.int32)

    def profile_chain(self, tables, userprofil, altabs, ro, terms, text_hop=True, want_tauabs=False):
        self.calls.append(("profile_chain", len(terms), int(tables["nb_pres"]), float(np.asarray(ro)[6, :49].sum())))
        return self.profile(altabs, None, terms) + (np.zeros((len(terms), 50)),)

    def upload(self, wl, groups=None, ngroup=None):
        outer = self

        class B:
            def __init__(s):
                s.ngroup, s.wl = ngroup, wl
                s.wmax = max(2 * o.nbmu + 1 for o in wl.optics)

            def free(s):
                outer.calls.append(("free",))
        return B()

    def set_group_direct(self, batch, direct):
        self.calls.append(("direct_groups", list(direct)))

    def run(self, batch, want_terms=True, want_groups=True):
        g, w = batch.ngroup, batch.wmax

        class G:
            n_rec = np.full(g, 3, np.int32)
            rec = np.ones((g, 5, 3, w)) * 0.01
            ttot_tronc, ttot_vrai, tauout = np.full(g, 0.2), np.full(g, 0.25), np.zeros(g)
            emoins, eplus = np.full(g, 0.15), np.full(g, 0.05)
        return None, G

    def batch_trphi(self, batch, igli, wind, ind, ifresnel, itrphi, phios, pas_phi, ipolar, download=True):
        nphi = 2 if itrphi == 1 else 360 // pas_phi + 1
        nmax = (batch.wmax - 1) // 2
        up = np.zeros((batch.ngroup, 7, nphi, nmax))
        for g, o in enumerate(batch.wl.optics):
            mu = np.asarray(o.rmu)[o.nbmu + 1:]
            for t in range(7):
                for ip in range(nphi):
                    up[g, t, ip, :o.nbmu] = (t + 1) * 0.01 * mu + 0.001 * ip
        return nphi, up, up + 1.0


def test_host_flow_with_stub_solver(tmp_path):
    """frontend.run_keywords, the user-angle outputs and sos.sos_proc end to end on the host with a stand-in for the device (shapes,
    file names, argument plumbing, optical-thickness scaling); the numbers come from the stub, the GPU tests check the real ones."""
    kwm, fe, aer = _mods()
    sos = importlib.import_module("radiativetransfer-sos_b200.sos")
    os.makedirs(os.path.join(str(tmp_path), "abs_root", "fic"))
    os.environ["SOS_ABS_ROOT"] = os.path.join(str(tmp_path), "abs_root")
    ac.write_wmo_file(os.path.join(os.environ["SOS_ABS_ROOT"], "fic", "Data_WMO_cor_2015_12_16"))
    ac.write_sf_files(os.path.join(os.environ["SOS_ABS_ROOT"], "fic"))
    s = _StubSolver()
    base = DEMO.format(root=str(tmp_path / "a"), nrad=12, naer=20, abs=7)
    res, aopt = fe.run_keywords(s, base.split(), wavelengths=[0.865, 0.910])
    assert len(res.dirs) == 2 and res.nphi == 2 and res.up.shape == (2, 7, 2, 13)
    for d in res.dirs:
        assert sorted(os.listdir(d)) == ["SOS_Down_Demo.txt", "SOS_Result.bin", "SOS_Up_Demo.txt"]
    assert sorted(os.listdir(str(tmp_path / "a" / "AER"))) == ["Aerosols_Demo.txt_0.865000", "Aerosols_Demo.txt_0.910000"]
    assert ("glitter", 13, 40, 24, 64) in s.calls and ("aerosols", 6, 3) in s.calls        # 2 wavelengths + the reference one, 2 components each
    assert abs(aopt[0].ta / aopt[1].ta - aopt[0].kmat1 / aopt[1].kmat1) < 1e-14 and 0.0 < aopt[1].ta < 0.3      # TA = K(WA) / K(WAREF) * AOT_REF
    # user angles + UserAng files, view 2, Shettle & Fenn aerosols, Roujean + Breon surface, transmissions
    ua = tmp_path / "user.txt"
    ua.write_text("10.\n40.\n")
    argv = (base.replace(str(tmp_path / "a"), str(tmp_path / "b")).replace("-SOS.View 1 -SOS.View.Phi 0.", "-SOS.View 2 -SOS.View.Dphi 90")
            .replace("-AER.Model 1 -AER.WMO.Model 2", "-AER.Model 2 -AER.SF.Model 3 -AER.SF.RH 70.")
            .replace("-SURF.Type 1", "-SURF.Type 5 -SURF.Roujean.K0 0.1 -SURF.Roujean.K1 0.05 -SURF.Roujean.K2 0.3")
            + " -ANG.Rad.UserAngFile %s -ANG.Aer.UserAngFile %s -SOS.ResFileUp.UserAng U.txt -SOS.ResFileDown.UserAng D.txt" % (ua, ua))
    res2, _ = fe.run_keywords(s, argv.split())
    assert res2.ind_angout.sum() == 2 and res2.up.shape == (1, 7, 5, 15) and ("direct", ["ibreon", "roujean"]) in s.calls
    rows = [ln for ln in open(os.path.join(res2.dirs[0], "U.txt")).read().split("\n") if ln and not ln.lstrip().startswith("#")]
    assert len(rows) == 5 * 2
    keep = np.flatnonzero(res2.ind_angout == 0)
    res1, _ = fe.run_keywords(s, base.replace(str(tmp_path / "a"), str(tmp_path / "c")).replace("-SOS.View 1 -SOS.View.Phi 0.", "-SOS.View 2 -SOS.View.Dphi 90").split())
    assert np.allclose(res1.up[0, 1, :5, :13], res2.up[0, 1, :5][:, keep], rtol=0, atol=1e-15)
    # Maignan's BPDF + Roujean (-SURF.Type 7); Nadal (6) is refused, as SOS_PROC refuses it (SOS_PROC.F:2210-2226)
    argv7 = (base.replace(str(tmp_path / "a"), str(tmp_path / "e"))
             .replace("-SURF.Type 1", "-SURF.Type 7 -SURF.Roujean.K0 0.1 -SURF.Roujean.K1 0.05 -SURF.Roujean.K2 0.3 -SURF.Maignan.C 6.0"))
    res7, _ = fe.run_keywords(s, argv7.split())
    assert res7.up.shape == (1, 7, 2, 13) and ("direct", ["maignan", "roujean"]) in s.calls
    with pytest.raises(ValueError, match="-SURF.Maignan.C"):
        fe.run_keywords(s, argv7.replace(" -SURF.Maignan.C 6.0", "").split())
    with pytest.raises(ValueError, match="Nadal"):
        fe.run_keywords(s, argv7.replace("-SURF.Type 7", "-SURF.Type 6 -SURF.Nadal.Alpha 0.0159 -SURF.Nadal.Beta 44.8").split())
    # user files: the aerosol result file of the first run as -AER.UserFile (simulation at the reference wavelength), a surface matrix
    # file as -SURF.File -- no aerosol / surface computation is asked from the device
    fm = importlib.import_module("radiativetransfer-sos_b200.formats")
    fsurf = str(tmp_path / "surf.bin")
    fm.write_surface_bin(fsurf, np.full((41, 9, 13, 13), 0.25, np.float32))
    faer = os.path.join(str(tmp_path / "a"), "AER", "Aerosols_Demo.txt_0.910000")
    n_before = len(s.calls)
    argv_u = (base.replace(str(tmp_path / "a"), str(tmp_path / "f")).replace("-AER.Waref 0.550", "-AER.Waref 0.910")
              .replace(" -AER.ResFile Aerosols_Demo.txt", "") + " -AER.UserFile %s -SURF.File %s" % (faer, fsurf))
    res_u, aer_u = fe.run_keywords(s, argv_u.split())
    new_calls = [c[0] for c in s.calls[n_before:]]
    assert "aerosols" not in new_calls and "glitter" not in new_calls and res_u.up.shape == (1, 7, 2, 13)
    assert aer_u[0].ta == 0.3 and aer_u[0].coef_tronca == float("%.5f" % aopt[1].coef_tronca) and not os.path.exists(str(tmp_path / "f" / "AER"))
    assert np.array_equal(aer_u[0].beta, np.array([float("%.7E" % x) for x in aopt[1].beta]))
    with pytest.raises(ValueError, match="2350"):
        fe.run_keywords(s, argv_u.replace("-AER.Waref 0.910", "-AER.Waref 0.550").split())
    with pytest.raises(ValueError, match="2351"):
        fe.run_keywords(s, (argv_u + " -AER.ResFile A.txt").split())
    fm.write_surface_bin(fsurf, np.zeros((41, 9, 12, 12), np.float32))
    with pytest.raises(ValueError, match="expected 41 records"):
        fe.run_keywords(s, argv_u.split())
    # -AER.Model 4 (external phase functions, at the reference wavelength only) and 5 (user mixture)
    import test_aerosol_models as tam
    fext, fmix = str(tmp_path / "ext.txt"), tmp_path / "mix.txt"
    tam._write_ext(fext)
    fmix.write_text(tam.MIX)
    argv4 = (base.replace(str(tmp_path / "a"), str(tmp_path / "g")).replace("-AER.Waref 0.550", "-AER.Waref 0.910")
             .replace("-AER.Model 1 -AER.WMO.Model 2", "-AER.Model 4 -AER.ExtData " + fext))
    res4, aer4 = fe.run_keywords(s, argv4.split())
    assert ("decompo", 20, 40, 41) in s.calls and aer4[0].ta == 0.3 and aer4[0].kmat1 == 2.5 and aer4[0].coef_tronca == 0.3
    assert os.path.exists(str(tmp_path / "g" / "AER" / "Aerosols_Demo.txt")) and res4.up.shape == (1, 7, 2, 13)
    with pytest.raises(ValueError, match="2331"):
        fe.run_keywords(s, argv4.replace("-AER.Waref 0.910", "-AER.Waref 0.550").split())
    with pytest.raises(ValueError, match="2330"):
        fe.run_keywords(s, argv4.replace(" -AER.ExtData " + fext, "").split())
    argv5 = (base.replace(str(tmp_path / "a"), str(tmp_path / "h")).replace("-AER.Model 1 -AER.WMO.Model 2", "-AER.Model 5 -AER.DefMixture %s" % fmix))
    res5, aer5 = fe.run_keywords(s, argv5.split())
    assert ("aerosols", 4, 2) in s.calls and 0.0 < aer5[0].ta and res5.up.shape == (1, 7, 2, 13)
    with pytest.raises(ValueError, match="2340"):
        fe.run_keywords(s, argv5.replace(" -AER.DefMixture %s" % fmix, "").split())
    # gaseous absorption from the keywords: a user profile file, CKD coefficient files under $SOS_ABS_ROOT
    import profile_cases as pc
    import test_absprofile as tab
    band = importlib.import_module("radiativetransfer-sos_b200.band")
    t = pc.ckd_tables(4)
    pc.write_ckd_files(os.environ["SOS_ABS_ROOT"], t)
    fprof = str(tmp_path / "gas_profile.txt")
    tab._write_profile(fprof, tab._user())
    wa = 1e4 / 13255.0
    argv_g = (base.replace(str(tmp_path / "a"), str(tmp_path / "i")).replace("-SOS_Main.Wa 0.910", "-SOS_Main.Wa %.12f" % wa)
              .replace("-AP.AbsProfile.Type 7", "-AP.AbsProfile.Type 0 -AP.AbsProfile.UserFile %s -AP.H2O 2.0" % fprof))
    res_g, _ = fe.run_keywords(s, argv_g.split())
    nterm = len(band.enumerate_ckd_terms(t["nexp"], t["ai"], 25)[0])
    pcall = [c for c in s.calls if c[0] == "profile_chain"][-1]
    assert res_g.nterm == [nterm] and pcall[1] == nterm and pcall[2] == pc.NPMAX and 4.0e24 < pcall[3] < 5.0e24     # O2 column, molecules / cm2
    # two wavelengths in two coefficient files: one profile-chain call per file, the terms of each wavelength from its own tables
    t2 = pc.ckd_tables(5)
    pc.write_ckd_files(os.environ["SOS_ABS_ROOT"], t2, numax=13000)
    n_before = len(s.calls)
    res_g2, _ = fe.run_keywords(s, argv_g.replace(str(tmp_path / "i"), str(tmp_path / "j")).split(), wavelengths=[wa, 1e4 / 12800.0])
    nterm2 = len(band.enumerate_ckd_terms(t2["nexp"], t2["ai"], 21)[0])
    pcalls = [c for c in s.calls[n_before:] if c[0] == "profile_chain"]
    assert res_g2.nterm == [nterm, nterm2] and [c[1] for c in pcalls] == [nterm, nterm2] and len(res_g2.dirs) == 2
    with pytest.raises(ValueError, match="2513"):
        fe.run_keywords(s, argv_g.replace("-AP.AerProfile.Type 1", "-AP.AerProfile.Type 2 -AP.AerLayer.Zmin 1. -AP.AerLayer.Zmax 3.").split())
    # the f2py-shaped entry
    out = sos.sos_proc(solver=s, resroot=str(tmp_path / "d"), wa_simu=0.910, tetas=35.0, nbmu_gauss_lum=12, nbmu_gauss_mie=20, waref_aot=0.55,
                       aot_ref=0.3, itronc_aer=1, imod_aer=1, imodele_wmo=2, hr=8.0, ha=2.0, iprofil=1, psurf=1013.0, absprofil=7, isurf=1,
                       surf_ind=1.34, wind=2.0, rho=0.0, itrphi=2, pas_phi=60, igmax=-999, zout=-999.0, ficangles_user_lum=str(ua), ier=0, trace=False)
    assert len(out) == 23 and out[0] == 15 and out[1].shape == (81,) and out[1].sum() == 2 and out[2].shape == (361,) and out[4].shape == (361, 81)
    assert list(out[2][:7]) == [0.0, 60.0, 120.0, 180.0, 240.0, 300.0, 360.0] and not out[4][7:].any() and out[4][:7, :15].all()
    assert 0.0 < out[18] < 1.0 and abs(out[20] - (out[19] + out[18])) < 1e-12 and out[21] == 0.05 and out[22] == 0.3


def test_command_line_entry(capsys):
    """python -m radiativetransfer-sos_b200 <keywords>: exit status 1 with a message for a bad command line, and -- on a machine
    without a GPU -- for a good one (no CPU fallback)."""
    import ctypes
    main = importlib.import_module("radiativetransfer-sos_b200.__main__").main
    assert main(["-SOS_Main.Wa", "0.5"]) == 1 and "missing required keyword" in capsys.readouterr().err
    assert main(["-SOS.Wavelength", "0.5"]) == 1 and "unknown keyword" in capsys.readouterr().err
    api = importlib.import_module("radiativetransfer-sos_b200.api")
    lib = api.load_library()
    lib.sosgpu_device_count.restype = ctypes.c_int
    if lib.sosgpu_device_count() == 0:
        argv = "-SOS_Main.Wa 0.91 -SOS_Main.ResRoot /tmp/x -ANG.Thetas 35. -SOS.View 1 -SURF.Type 0 -SURF.Alb 0.1 -AP.AerProfile.Type 1 -AP.AbsProfile.Type 7"
        assert main(argv.split()) == 1 and "no CPU fallback" in capsys.readouterr().err
