"""Gas atmosphere and CKD tables of a run from the -AP.* keywords: what SOS_PREPA_ABSPROFILE (SOS_PREPA_ABSPROFILE.F:248-752) and
DATATM (SOS_SUB_TRS.F:908-1003) prepare on the host, once per run, before the per-term chain SOS_ABSPROFILE -> SOS_PROFILE runs
on the device (band.run_band).  Host work on 50 levels x 8 gases; the statements of the reference in their order, REAL*4 literals
as (double)(float) values.

    -AP.AbsProfile.Type 0   user profile file (-AP.AbsProfile.UserFile): 50 lines "level, altitude (km), pressure (mbar),
                            temperature (K), H2O, CO2, O3, N2O, CO, CH4, O2 (ppmv), air density, NO2, SO2"
    -AP.AbsProfile.Type 1-6 tropical, mid-latitude summer / winter, sub-arctic summer / winter, US standard 1962.  The tables of
                            these atmospheres are DATA statements of the reference's own source file: they are READ from the
                            user's installation of the reference ($SOS_ABS_ROOT/src/SOS_SUB_TRS.F), like the CKD coefficient
                            files and fic/SO2-NO2 -- this package holds no copy of the reference's data.
    -AP.Psurf, -AP.H2O, -AP.O3, -AP.CO2, -AP.CH4   scalings of the profile (pressure, column amounts, surface concentrations)
    -AP.SpectralResol       resolution of the CKD coefficient files read by sosgpu_read_ckd_coeff (READ_CKD_COEFF)

`prepare` returns the `gas=` dict of frontend.run / band.run_band."""
import ctypes as C
import os
import re

import numpy as np

NLEVEL, NBCOL, NBABS = 50, 13, 8                       # CTE_ABS_NBLEV, CTE_ABS_NBCOL, CTE_NBABS (SOS.h:246-254)
NWVL, NAI, NTMAX, NPMAX, NCMAX = 50, 5, 9, 31, 12      # CTE_CKD_* (SOS.h:278-282)
CKD_NUMAX, CKD_NUMIN = 27500.0, 2500.0                 # SOS.h:290-291
NOT_DEFINED = -999.0
_f = lambda x: float(np.float32(x))                    # a REAL*4 literal promoted to DOUBLE PRECISION
ATMOCM = [_f(v) for v in (3.410E+22, 1.395E+22, 1.279E+22, 1.395E+22, 2.192E+22, 3.837E+22, 1.918E+22, 1.3340E+22)]   # :391-392
PDSMOL = [18., 44., 48., 44., 28., 16., 32., 46.]
ROUTINES = {1: "TROPICA", 2: "MIDLASU", 3: "MIDLAWI", 4: "SUBSUMM", 5: "SUBWINT", 6: "USTAD62"}        # DATATM, SOS_SUB_TRS.F:945-956
# DONUSER column (1-based) of each DATA array of the atmosphere routines (SOS_SUB_TRS.F:958-971)
_COLUMN = {"ALT1": 1, "P1": 2, "T1": 3, "ROH2O1": 4, "ROCO21": 5, "ROO31": 6, "RON2O1": 7, "ROCO1": 8, "ROCH41": 9, "ROO21": 10, "DENS1": 11}


def _num(t):
    return float(t.replace("D", "E").replace("d", "e"))


def read_user_profile(path):
    """The user's gas profile (ABSPROFIL = 0, :443-449): 50 list-directed records II, USERPROFIL(I, 1:13) -> [50, 13]."""
    user = np.zeros((NLEVEL, NBCOL), order="F")
    try:
        with open(path) as f:
            rows = [ln.replace(",", " ").split() for ln in f if ln.strip()]
    except OSError:
        raise ValueError("-AP.AbsProfile.UserFile %s cannot be opened (SOS_PREPA_ABSPROFILE error 1012)" % path)
    try:
        for i in range(NLEVEL):
            user[i, :] = [_num(v) for v in rows[i][1:1 + NBCOL]]
    except (IndexError, ValueError):
        raise ValueError("-AP.AbsProfile.UserFile %s: 50 records of a level number and 13 values are expected (error 1021)" % path)
    return user


def standard_atmosphere(iatm, sos_abs_root=None):
    """Columns 1..11 of USERPROFIL for a predefined atmosphere, from the DATA statements of its routine in the user's copy of
    the reference's SOS_SUB_TRS.F (REAL*4 constants stored in DOUBLE PRECISION arrays) -> [50, 13] with columns 12, 13 zero."""
    root = sos_abs_root or os.environ.get("SOS_ABS_ROOT", "")
    src = os.path.join(root, "src", "SOS_SUB_TRS.F")
    if not os.path.exists(src):
        raise ValueError("-AP.AbsProfile.Type %d: the tables of the predefined atmospheres are read from $SOS_ABS_ROOT/src/SOS_SUB_TRS.F "
                         "(the reference's installation); %r does not exist" % (iatm, src))
    text = open(src, encoding="latin-1").read().split("\n")
    name = ROUTINES[iatm]
    start = next(i for i, ln in enumerate(text) if re.match(r"\s+SUBROUTINE\s+%s\b" % name, ln, re.I))
    end = next(i for i in range(start + 1, len(text)) if re.match(r"\s+END\s*$", text[i], re.I))
    user = np.zeros((NLEVEL, NBCOL), order="F")
    seen = set()
    i = start
    while i < end:
        m = re.match(r"\s{6}\s*DATA\s+(\w+)\s*/(.*)$", text[i], re.I) if text[i][:1] not in "Cc*!" else None
        if not m:
            i += 1
            continue
        arr, body = m.group(1).upper(), m.group(2)
        i += 1
        while "/" not in body:                             # continuation lines: any character in column 6
            if len(text[i]) > 5 and text[i][:5].strip() == "" and text[i][5] not in " 0":
                body += text[i][6:]
            i += 1
        vals = [float(np.float32(_num(v))) for v in body.split("/")[0].replace(",", " ").split()]
        if arr in _COLUMN:
            if len(vals) != NLEVEL:
                raise ValueError("%s: DATA %s of %s holds %d values, expected %d" % (src, arr, name, len(vals), NLEVEL))
            user[:, _COLUMN[arr] - 1] = vals
            seen.add(arr)
    if seen != set(_COLUMN):
        raise ValueError("%s: routine %s lacks the DATA arrays %s" % (src, name, sorted(set(_COLUMN) - seen)))
    return user


def datatm(user, iatm, psurf):
    """DATATM (SOS_SUB_TRS.F:908-1003).  user: [50, 13] -- the user's profile (iatm 0) or the predefined atmosphere's columns;
    returns (ro [8, 50] mass mixing ratios per level, p [50], t [50], alt [50], user with the pressure column scaled)."""
    user = np.array(user, dtype=np.float64, order="F")
    coef = 1.0
    if psurf > 0.0:
        coef = psurf / user[0, 1]
    if iatm > 0:                                          # predefined atmosphere: the profile handed on (DONUSER) gets the scaled
        p = user[:, 1].copy()                             # pressures, the pressures returned for the layer amounts do not (:957-962)
        user[:, 1] = p * coef
    else:                                                 # user profile: the other way round (:922-926)
        p = user[:, 1] * coef
    alt, t = user[:, 0].copy(), user[:, 2].copy()
    ro = np.zeros((NBABS, NLEVEL), order="F")
    e6 = _f(1.0E-06)
    m = _f(28.97)
    for k, col, mol in ((2, 5, 44.0), (4, 7, 44.0), (5, 8, 28.0), (6, 9, 16.0), (7, 10, 32.0), (3, 6, 48.0)):     # :977-989
        ro[k - 1] = user[:, col - 1] * e6 * mol / m
    h = user[:, 3] * e6 * 18.0 / m
    ro[0] = h / (1 + h)
    return ro, p, t, alt, user


def read_so2_no2(sos_abs_root=None):
    """fic/SO2-NO2 (:462-468): 50 records SO2, NO2 -> (no2 [50], so2 [50])."""
    root = sos_abs_root or os.environ.get("SOS_ABS_ROOT", "")
    if not root:
        raise ValueError("SOS_ABS_ROOT is not set (SOS_PREPA_ABSPROFILE error 925)")
    path = os.path.join(root, "fic", "SO2-NO2")
    try:
        rows = [[_num(v) for v in ln.replace(",", " ").split()[:2]] for ln in open(path) if ln.strip()][:NLEVEL]
        a = np.array(rows)
        assert a.shape == (NLEVEL, 2)
    except Exception:
        raise ValueError("error while reading %s (SOS_PREPA_ABSPROFILE error 927)" % path)
    return a[:, 1], a[:, 0]


def atmosphere(absprofil, ficabsprofil=None, psurf=NOT_DEFINED, h2o=NOT_DEFINED, o3=NOT_DEFINED, co2=NOT_DEFINED, ch4=NOT_DEFINED,
               sos_abs_root=None):
    """SOS_PREPA_ABSPROFILE.F:441-571 -> (userprofil [50, 13], altabs [50] descending, ro [8, 50] particles / cm2 per layer)."""
    if absprofil == 0:
        if not ficabsprofil:
            raise ValueError("-AP.AbsProfile.Type 0 requires -AP.AbsProfile.UserFile")
        user = read_user_profile(ficabsprofil)
    elif absprofil in ROUTINES:
        user = standard_atmosphere(absprofil, sos_abs_root)
        user[:, 11], user[:, 12] = read_so2_no2(sos_abs_root)
    else:
        raise ValueError("-AP.AbsProfile.Type %r: 0 (user file), 1 .. 6 (predefined atmospheres) or 7 (no gaseous absorption)" % (absprofil,))
    ro, p, t, altc, user = datatm(user, absprofil, psurf)
    ro[7] = user[:, 11] * _f(1.0E-06) * 46 / _f(28.9)                                   # NO2 (:473)
    co2_default = ro[1, 0] * _f(28.97) / _f(44.0E-06) if co2 >= 0.0 else None           # back to ppmv (:476-477)
    ch4_default = ro[5, 0] * _f(28.97) / _f(16.0E-06) if ch4 >= 0.0 else None
    altabs = np.ascontiguousarray(altc[::-1])                                           # :480-482
    for j in range(NLEVEL - 1):                                                         # :484-491; the top level keeps its level value
        dp = p[j] - p[j + 1]
        for k in range(NBABS):
            ro[k, j] = dp * (ro[k, j] + ro[k, j + 1]) / 2.0 * ATMOCM[k]
    avo = _f(6.022E+23)
    if h2o >= 0.0:                                                                      # :493-506
        q = 0.0
        for j in range(NLEVEL):
            q = q + ro[0, j]
        q = q / avo * PDSMOL[0]
        ro[0] = ro[0] * h2o / q
        user[:, 3] = user[:, 3] * h2o / q
    if o3 >= 0.0:                                                                       # :508-525
        o3 = o3 / _f(1000.)
        q = 0.0
        for j in range(NLEVEL):
            q = q + ro[2, j]
        q = q / avo * PDSMOL[2]
        q = q * _f(466.23)
        ro[2] = ro[2] * o3 / q
        user[:, 5] = user[:, 5] * o3 / q
    if co2 >= 0.0:                                                                      # :527-533
        ro[1] = ro[1] * co2 / co2_default
        user[:, 4] = user[:, 4] * co2 / co2_default
    if ch4 >= 0.0:                                                                      # :535-541
        ro[5] = ro[5] * ch4 / ch4_default
        user[:, 8] = user[:, 8] * ch4 / ch4_default
    return user, altabs, ro


def read_ckd(lib, nu, nustep, sos_abs_root=None):
    """READ_CKD_COEFF for the eight gases at wavenumber nu (SOS_PREPA_ABSPROFILE.F:553-563) through the host reader of libsosgpu.so
    -> (tables dict as band.run_band takes it, numax, numin)."""
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    nexp = np.zeros((NBABS, NWVL), dtype=np.int32, order="F")
    ai = np.zeros((NAI, NBABS, NWVL), order="F")
    ki = np.zeros((NTMAX, NPMAX, NAI, NBABS, NWVL), order="F")
    kh = np.zeros((NTMAX, NPMAX, NCMAX, NAI, NWVL), order="F")
    tp, tt, tc = np.zeros(NPMAX), np.zeros(NTMAX), np.zeros(NCMAX)
    numax, numin = C.c_double(0), C.c_double(0)
    nbp, nbt, nbc = C.c_int(0), C.c_int(0), C.c_int(0)
    root = (sos_abs_root or os.environ.get("SOS_ABS_ROOT", "")).encode()
    lib.sosgpu_read_ckd_coeff.restype = C.c_int
    for k in range(1, NBABS + 1):
        rc = lib.sosgpu_read_ckd_coeff(root, C.c_int(k), C.c_int(1), C.c_double(nu), C.c_double(nustep), nexp.ctypes.data_as(C.POINTER(C.c_int)),
                                       P(ai), P(ki), P(kh), C.byref(numax), C.byref(numin), P(tp), C.byref(nbp), P(tt), C.byref(nbt),
                                       P(tc), C.byref(nbc))
        if rc != 0:
            raise ValueError("READ_CKD_COEFF failed for gas %d at %.3f cm-1, resolution %g cm-1 under %r" % (k, nu, nustep, root.decode()))
    tables = dict(nb_temp=nbt.value, nb_pres=nbp.value, nb_conc=nbc.value, tab_temp=tt, tab_pres=tp, tab_conc=tc, nexp=nexp, ai=ai,
                  ki=ki, kh=kh)
    return tables, numax.value, numin.value


def prepare(lib, wavelengths, nustep, absprofil, ficabsprofil=None, psurf=NOT_DEFINED, h2o=NOT_DEFINED, o3=NOT_DEFINED, co2=NOT_DEFINED,
            ch4=NOT_DEFINED, sos_abs_root=None):
    """SOS_PREPA_ABSPROFILE for a list of wavelengths (microns): the `gas=` dict of frontend.run -- userprofil, altabs, ro of the
    run, and per wavelength the CKD tables of its coefficient file (50 spectral intervals per file; wavelengths of the same file
    share one set, read once), kdis_ai and lamb1 = 1 + INT((NUMAX - NU) / NUSTEP) (:565)."""
    user, altabs, ro = atmosphere(absprofil, ficabsprofil, psurf, h2o, o3, co2, ch4, sos_abs_root)
    files, tables, lamb1 = [], [], []                         # files: (numax, numin, tables) already read
    for wa in wavelengths:
        nu = _f(1.0E+4) / float(wa)
        if nu > CKD_NUMAX or nu < CKD_NUMIN:
            raise ValueError("the wavelength %g is not included in the spectral range of the CKD data, %g .. %g cm-1 (error 905)"
                             % (wa, CKD_NUMIN, CKD_NUMAX))
        hit = next((f for f in files if f[1] <= nu < f[0]), None)       # the file READ_CKD_COEFF picks: NUMIN <= NU < NUMAX
        if hit is None:
            t, numax, numin = read_ckd(lib, nu, nustep, sos_abs_root)
            hit = (numax, numin, t)
            files.append(hit)
        tables.append(hit[2])
        lamb1.append(1 + int((hit[0] - nu) / nustep))
    return dict(tables=tables, kdis_ai=[t["ai"] for t in tables], userprofil=user, altabs=altabs, ro=ro, lamb1=lamb1)
