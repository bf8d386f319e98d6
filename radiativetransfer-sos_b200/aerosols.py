"""Host side of the aerosol chain: what SOS_AEROSOLS decides on the host before and after its Mie / size-
distribution / Legendre work -- which components a model has, their refractive indices rounded to the MIE file-name precision,
the size-parameter range of each Mie table, the mixture weights, the aerosol optical thickness at the simulation wavelength --
for a LIST of wavelengths, with the work itself done on the device by Solver.aerosols (one call for all wavelengths).

    mono-modal (IMOD = 0)        SOS_AEROSOLS.F:1158-1304   log-normal or Junge size distribution
    WMO models (IMOD = 1)        SOS_AEROSOLS.F:1309-1510   dust-like / water-soluble / oceanic / soot, SOS_INIT_PARAMWMO :3334-3556
    Shettle & Fenn (IMOD = 2)    SOS_AEROSOLS.F:1514-1702   rural / urban / oceanic components at a relative humidity, SOS_INIT_PARAMSF
                                                            :3557-3843
    bimodal log-normal (IMOD=3)  SOS_AEROSOLS.F:1706-2123   volume concentrations given, or the coarse share of the optical
                                                            thickness at the reference wavelength (MODE_PARAM_BILND = 1 / 2)
    optical thickness at WA      SOS_PROC.F:2941-3063       TA = KMAT1(WA) / KMAT1(WAREF) * AOT_REF

External phase functions (4) and user mixtures (5) are not built.  Keywords: frontend.aerosol_model maps -AER.* to these models.
No CPU fallback: the numbers come from Solver.aerosols, which needs the GPU."""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

ALPHA0 = 0.0001                                    # CTE_MIE_ALPHAMIN (SOS.h:116)
COEF_NRMAX = float(np.float32(0.0001))             # CTE_COEF_NRMAX (SOS.h:134), REAL*4 literal
WAMIN = float(np.float32(0.364))                   # CTE_WAMIN (SOS.h:70)
NOT_DEFINED = -999.0                               # CTE_NOT_DEFINED_VALUE_DBLE (SOS.h:78)
ALPHAMAX_WMO = (4000.0, 50.0, 800.0, 10.0)         # CTE_ALPHAMAX_WMO_DL / WS / OC / SO (SOS.h:122-125)
ALPHAMAX_SF = {0: 70.0, 2: 90.0}                   # CTE_ALPHAMAX_SF_SR / SU (SOS.h:126-127): small rural, small urban
_f32 = lambda *v: tuple(float(np.float32(x)) for x in v)
# volume proportions of DL, WS, OC, SO (SOS_AEROSOLS.F:1340-1353) and number densities of the five Shettle & Fenn components
# (:1536-1551): REAL*4 literals assigned to DOUBLE PRECISION variables -- 0.70 is (double)0.70f there
WMO_VOLUMES = {1: _f32(0.70, 0.29, 0.0, 0.01), 2: _f32(0.0, 0.05, 0.95, 0.0), 3: _f32(0.17, 0.61, 0.0, 0.22)}
SF_DENSITIES = {1: _f32(1.0, 0.0, 0.0, 0.0, 0.0), 2: _f32(0.0, 0.0, 0.999875, 0.000125, 0.0), 3: _f32(0.99, 0.0, 0.0, 0.0, 0.01),
                4: _f32(0.995, 0.0, 0.0, 0.0, 0.005)}
SF_FILES = ("Data_SF_cor_2015_12_16", "IRefrac_SR_cor_2015_12_16", "IRefrac_LR", "IRefrac_SU_cor_2015_12_16", "IRefrac_LU_cor_2015_12_16",
            "IRefrac_OM_cor_2015_12_16")          # CTE_AER_DATASF, CTE_AER_SR_SF .. CTE_AER_OM_SF (SOS.h:149-160)
_PI = float(np.arccos(-1.0))


def dnint(x):
    """Fortran DNINT: nearest whole number, halves away from zero."""
    return float(np.trunc(x + 0.5)) if x >= 0 else float(-np.trunc(-x + 0.5))


def round_index(rn, in_):
    """Refractive index forced to the F5.3 / F8.5 precision of the MIE file name (SOS_AEROSOLS.F:1166-1167, 1766-1767)."""
    return dnint(rn * 1000.0) / 1000.0, -dnint(-in_ * 100000.0) / 100000.0


def lnd_rmax(rmodal, sigma):
    """Radius where a log-normal mode has fallen to CTE_COEF_NRMAX of its maximum (SOS_AEROSOLS.F:1175-1177, 1938-1940)."""
    return rmodal * np.exp(sigma * sigma) * np.exp(sigma * np.sqrt(-2.0 * np.log(COEF_NRMAX)))


def alphaf_of(rmax, wa):
    """ALPHAF = REAL(100 + 100 * DINT(2. * PI * RMAX / (100. * WA))) (SOS_AEROSOLS.F:1183-1184, 1942-1943)."""
    return float(np.float32(100 + 100 * np.trunc(2.0 * _PI * rmax / (100.0 * wa))))


def interpol(y1, y2, x1, x2, x):
    """SOS_INTERPOL (SOS_AEROSOLS.F:3844-3862)."""
    return ((y2 - y1) / (x2 - x1)) * (x - x2) + y2


def wmo_params(path, wa):
    """SOS_INIT_PARAMWMO (SOS_AEROSOLS.F:3334-3556): the WMO data file (fixed-column formats 333 / 444 / 555) -> modal radii,
    log-normal sigmas (the file's log10 values times ln 10), volumes of one particle, and the refractive indices interpolated to
    `wa` and rounded to the MIE file-name precision; four components DL, WS, OC, SO.  Indices stay 0 when `wa` is outside the
    table, as in the reference."""
    cols = _cols
    with open(path) as f:
        lines = f.read().split("\n")
    f9 = [(1, 9)] * 4
    v1 = cols(lines[0], f9)
    v2 = [x * np.log(10.0) for x in cols(lines[1], f9)]
    vol = cols(lines[2], [(1, 9), (1, 12), (1, 9), (1, 12)])
    mr, mi = [0.0] * 4, [0.0] * 4
    rows = [cols(ln, [(1, 9)] * 9) for ln in lines[3:] if ln.strip()]
    for a, b in zip(rows[:-1], rows[1:]):
        if a[0] <= wa <= b[0]:
            for i in range(4):
                r = interpol(a[1 + 2 * i], b[1 + 2 * i], a[0], b[0], wa)
                im = interpol(a[2 + 2 * i], b[2 + 2 * i], a[0], b[0], wa)
                mr[i], mi[i] = round_index(r, im)
            break
    return v1, v2, mr, mi, vol


def _cols(line, widths):
    out, pos = [], 0
    for skip, w in widths:
        pos += skip
        out.append(float(line[pos:pos + w].strip() or 0.0))
        pos += w
    return out


def sf_params(dirfic, wa, rh):
    """SOS_INIT_PARAMSF (SOS_AEROSOLS.F:3557-3843): the Shettle & Fenn data files (formats 222 / 333 / 555) -> modal radii at the
    relative humidity `rh` (percent), log-normal sigmas, refractive indices interpolated in wavelength and humidity and rounded to
    the MIE file-name precision; five components small rural, large rural, small urban, large urban, oceanic."""
    import os
    with open(os.path.join(dirfic, SF_FILES[0])) as f:
        lines = [ln for ln in f.read().split("\n") if ln.strip()]
    v2 = [x * np.log(10.0) for x in _cols(lines[0], [(1, 9)] * 5)]
    rows = [_cols(ln, [(1, 5)] + [(1, 9)] * 5) for ln in lines[1:]]
    rh1, rm1, cpt = rows[0][0], rows[0][1:], 1
    v1, rh2 = None, None
    if rh1 == rh:
        v1 = list(rm1)
    else:
        for row in rows[1:]:
            rh2, rm2 = row[0], row[1:]
            cpt += 1
            if rh1 < rh <= rh2:
                v1 = [interpol(a, b, rh1, rh2, rh) for a, b in zip(rm1, rm2)]
                break
            rh1, rm1 = rh2, rm2
    if v1 is None:
        raise ValueError("relative humidity %g outside the Shettle & Fenn table" % rh)
    mr, mi = [0.0] * 5, [0.0] * 5
    for i in range(5):
        with open(os.path.join(dirfic, SF_FILES[1 + i])) as f:
            tab = [_cols(ln, [(1, 9)] * 17) for ln in f.read().split("\n") if ln.strip()]
        for a, b in zip(tab[:-1], tab[1:]):
            if a[0] <= wa <= b[0]:
                col = lambda row, h, im: row[1 + 2 * (h - 1) + im]          # MR(h) / MI(h), h = 1..8 humidities
                if cpt == 1:
                    r = interpol(col(a, 1, 0), col(b, 1, 0), a[0], b[0], wa)
                    m = interpol(col(a, 1, 1), col(b, 1, 1), a[0], b[0], wa)
                else:
                    r = interpol(interpol(col(a, cpt - 1, 0), col(b, cpt - 1, 0), a[0], b[0], wa),
                                 interpol(col(a, cpt, 0), col(b, cpt, 0), a[0], b[0], wa), rh1, rh2, rh)
                    m = interpol(interpol(col(a, cpt - 1, 1), col(b, cpt - 1, 1), a[0], b[0], wa),
                                 interpol(col(a, cpt, 1), col(b, cpt, 1), a[0], b[0], wa), rh1, rh2, rh)
                mr[i], mi[i] = r, m
                break
        mr[i], mi[i] = round_index(mr[i], mi[i])
    return v1, v2, mr, mi


@dataclass
class ShettleFenn:
    """-AER.Model 2: imodele 1 tropospheric, 2 urban, 3 maritime, 4 coastal; rh = relative humidity in percent; dirfic = the
    directory of the Shettle & Fenn data files ($SOS_ABS_ROOT/fic)."""
    dirfic: str
    imodele: int
    rh: float


@dataclass
class MonoModal:
    """-AER.Model 0: one log-normal (igranu 1: v1 = modal radius, v2 = sigma) or Junge (igranu 2: v1 = r0, v2 = slope, v3 = rmax)
    mode; the refractive index may be a function of the wavelength."""
    rn: object
    in_: object
    igranu: int
    v1: float
    v2: float
    v3: float = NOT_DEFINED


@dataclass
class Wmo:
    """-AER.Model 1: imodele 1 continental, 2 maritime, 3 urban, 4 user volumes (percent / 100) of DL, WS, OC, SO."""
    datafile: str
    imodele: int
    user_volumes: Sequence[float] = (0.0, 0.0, 0.0, 0.0)


@dataclass
class BimodalLnd:
    """-AER.Model 3: coarse and fine log-normal modes.  Either the volume concentrations (cv_coarse, cv_fine), or rtauct = the
    coarse mode's share of the optical thickness at the reference wavelength (then the indices at that wavelength are needed:
    give the indices as functions of the wavelength, or constants)."""
    coarse_rn: object
    coarse_in: object
    coarse_rmodal: float
    coarse_sigma: float
    fine_rn: object
    fine_in: object
    fine_rmodal: float
    fine_sigma: float
    cv_coarse: Optional[float] = None
    cv_fine: Optional[float] = None
    rtauct: Optional[float] = None


@dataclass
class AerosolOptics:
    """What SOS_AEROSOLS leaves for one wavelength: the result file's contents + the optical thickness SOS_PROC derives."""
    wa: float
    kmat1: float
    kmat2: float
    piz: float
    piztr: float
    coef_tronca: float
    asym: float
    itronc: int
    alpha: np.ndarray
    beta: np.ndarray
    gamma: np.ndarray
    zeta: np.ndarray
    ta: float = 0.0


@dataclass
class Plan:
    components: List[tuple] = field(default_factory=list)      # (rn, in, alpha0, alphaf, igranu, v1, v2, v3, wa)
    models: List[tuple] = field(default_factory=list)          # (ncomp, component indices, weights, itronc)
    wavelengths: List[float] = field(default_factory=list)


def _at(x, wa):
    return float(x(wa)) if callable(x) else float(x)


def _lnd_component(rn, in_, rmodal, sigma, wa, wa_for_alphaf):
    rn, in_ = round_index(rn, in_)
    af = alphaf_of(lnd_rmax(rmodal, sigma), wa_for_alphaf)
    if ALPHA0 > af or af >= 1e5:
        raise ValueError("size-parameter range of the Mie table out of bounds (SOS_AEROSOLS error 1009)")
    return (rn, in_, ALPHA0, af, 1, rmodal, sigma, NOT_DEFINED, wa)


def plan(model, wavelengths, itronc=1, bilnd_weights=None):
    """Components and models of Solver.aerosols for `wavelengths` (microns).  For a BimodalLnd given by rtauct, bilnd_weights are
    the normalised CVI of the two modes (from reference_weights)."""
    p = Plan(wavelengths=[float(w) for w in wavelengths])
    for wa in p.wavelengths:
        n0 = len(p.components)
        if isinstance(model, MonoModal):
            rn, in_ = round_index(_at(model.rn, wa), _at(model.in_, wa))
            if in_ > 0.0:
                raise ValueError("imaginary parts of refractive indexes have to be negative")
            # the mono-modal table is sized for the shortest wavelength of the code, CTE_WAMIN, whatever WA is (:1183)
            rmax = lnd_rmax(model.v1, model.v2) if model.igranu == 1 else model.v3
            af = alphaf_of(rmax, WAMIN)
            if ALPHA0 > af or af >= 1e5:
                raise ValueError("size-parameter range of the Mie table out of bounds (SOS_AEROSOLS error 1009)")
            p.components.append((rn, in_, ALPHA0, af, model.igranu, model.v1, model.v2, model.v3, wa))
            p.models.append((0, [n0], [1.0], itronc))
        elif isinstance(model, Wmo):
            v1, v2, mr, mi, vol = wmo_params(model.datafile, wa)
            c = list(WMO_VOLUMES[model.imodele]) if model.imodele in WMO_VOLUMES else [float(x) for x in model.user_volumes]
            n = [ci / vi for ci, vi in zip(c, vol)]                     # N(I) = C(I) / V(I) (:1377-1380)
            ntot = 0.0
            for x in n:
                ntot = ntot + x
            idx, wts = [], []
            for i in range(4):
                if c[i] == 0.0:
                    continue
                idx.append(len(p.components))
                wts.append(n[i] / ntot)
                p.components.append((mr[i], mi[i], ALPHA0, ALPHAMAX_WMO[i], 1, v1[i], v2[i], NOT_DEFINED, wa))
            p.models.append((len(idx), idx, wts, itronc))
        elif isinstance(model, ShettleFenn):
            v1, v2, mr, mi = sf_params(model.dirfic, wa, model.rh)
            ni = SF_DENSITIES[model.imodele]
            idx, wts = [], []
            for i in range(5):
                if ni[i] == 0.0:
                    continue
                af = ALPHAMAX_SF[i] if i in ALPHAMAX_SF else alphaf_of(lnd_rmax(v1[i], v2[i]), wa)       # :1582-1591
                if ALPHA0 > af or af >= 1e5:
                    raise ValueError("size-parameter range of the Mie table out of bounds (SOS_AEROSOLS error 1009)")
                idx.append(len(p.components))
                wts.append(ni[i])                                      # number densities as they are (:1657-1664): they sum to 1
                p.components.append((mr[i], mi[i], ALPHA0, af, 1, v1[i], v2[i], NOT_DEFINED, wa))
            p.models.append((len(idx), idx, wts, itronc))
        elif isinstance(model, BimodalLnd):
            if in_pos(model, wa):
                raise ValueError("imaginary parts of refractive indexes have to be negative")
            p.components.append(_lnd_component(_at(model.coarse_rn, wa), _at(model.coarse_in, wa), model.coarse_rmodal, model.coarse_sigma, wa, wa))
            p.components.append(_lnd_component(_at(model.fine_rn, wa), _at(model.fine_in, wa), model.fine_rmodal, model.fine_sigma, wa, wa))
            if model.rtauct is None:
                cv = [float(model.cv_coarse), float(model.cv_fine)]
                ntot = cv[0] + cv[1]
                w = [cv[0] / ntot, cv[1] / ntot]                        # :2060-2062
            else:
                if bilnd_weights is None:
                    raise ValueError("BimodalLnd with rtauct: pass bilnd_weights=reference_weights(...)")
                w = list(bilnd_weights)
            p.models.append((2, [n0, n0 + 1], w, itronc))
        else:
            raise TypeError("unsupported aerosol model %r" % (model,))
    return p


def in_pos(model, wa):
    return _at(model.coarse_in, wa) > 0.0 or _at(model.fine_in, wa) > 0.0


def reference_weights(solver, nbmu, xmu, xhr, model, waref, aot_ref):
    """MODE_PARAM_BILND = 2 (SOS_AEROSOLS.F:1806-2058): the extinction cross sections of the two modes at the reference
    wavelength give CVI(coarse) = rtauct * AOT_REF / KMAT1c, CVI(fine) = (1 - rtauct) * AOT_REF / KMAT1f, then normalised."""
    comps = [_lnd_component(_at(model.coarse_rn, waref), _at(model.coarse_in, waref), model.coarse_rmodal, model.coarse_sigma, waref, waref),
             _lnd_component(_at(model.fine_rn, waref), _at(model.fine_in, waref), model.fine_rmodal, model.fine_sigma, waref, waref)]
    o = solver.aerosols(nbmu, xmu, xhr, comps, [], 2, want_phase=False)
    if o["comp_ier"].any():
        raise RuntimeError("SOS_GRANU failed at the reference wavelength")
    cv = [(model.rtauct * aot_ref) / o["comp_k"][0, 0], ((1.0 - model.rtauct) * aot_ref) / o["comp_k"][1, 0]]
    ntot = cv[0] + cv[1]
    return [cv[0] / ntot, cv[1] / ntot]


def run(solver, nbmu, xmu, xhr, os_nb, model, wavelengths, waref=None, aot_ref=None, itronc=1):
    """Aerosol optics of every wavelength in one device call (+ one for the reference wavelength when an optical thickness is to
    be scaled): list of AerosolOptics.  ta = KMAT1(wa) / KMAT1(waref) * aot_ref (SOS_PROC.F:3063), aot_ref itself at waref."""
    wl = [float(w) for w in wavelengths]
    with_ref = waref is not None and aot_ref is not None and aot_ref != 0.0
    weights = None
    if isinstance(model, BimodalLnd) and model.rtauct is not None:
        if not with_ref:
            raise ValueError("BimodalLnd with rtauct needs waref and aot_ref")
        weights = reference_weights(solver, nbmu, xmu, xhr, model, float(waref), float(aot_ref))
    allw = wl + ([float(waref)] if with_ref and float(waref) not in wl else [])
    p = plan(model, allw, itronc, weights)
    o = solver.aerosols(nbmu, xmu, xhr, p.components, p.models, os_nb, want_phase=False)
    if o["comp_ier"].any() or o["model_ier"].any():
        raise RuntimeError("aerosol chain failed: component codes %s, model codes %s" % (o["comp_ier"].tolist(), o["model_ier"].tolist()))
    k_ref = o["scal"][allw.index(float(waref)), 0] if with_ref else None
    out = []
    for m, wa in enumerate(wl):
        s, c = o["scal"][m], o["coef"][m]
        ta = 0.0
        if with_ref:
            ta = float(aot_ref) if wa == float(waref) else (s[0] / k_ref) * float(aot_ref)
        out.append(AerosolOptics(wa, s[0], s[1], s[2], s[3], s[4], s[5], int(s[7]), c[0].copy(), c[1].copy(), c[2].copy(), c[3].copy(), ta))
    return out


def through_result_file(a: AerosolOptics):
    """The values SOS_PREPA_OS reads back from the aerosol result file (E15.8 coefficients, F9.5 truncation coefficient and
    albedo, SOS_PREPA_OS.F:669-693): what reaches the solver in the reference's flow."""
    r8 = lambda v: np.array([float("%.7E" % x) for x in v])
    return dict(alpha=r8(a.alpha), beta=r8(a.beta), gamma=r8(a.gamma), zeta=r8(a.zeta), a_trunc=float("%.5f" % a.coef_tronca),
                piztr=float("%.5f" % a.piztr))


# ---- the MIE file cache of -AER.DirMie: names and format are the reference's, so either side can reuse the other's files ----
def mie_file_name(nbmu_gauss, rn, in_, alpha0, alphaf):
    """SOS_NOM_FICMIE (SOS_AEROSOLS.F:3128-3254) without a user angle file: MIEr.rrr-i.iiiii-a.aaaa-AAAAA.AA-MUnn, the digits
    being INT(RN*1000), INT(-IN*100000), INT(ALPHAO*10000), INT(ALPHAF*100) zero-padded (truncation, as the reference)."""
    crn, cin = "%04d" % int(rn * 1000), "%06d" % int(-in_ * 100000)
    ca0, caf = "%05d" % int(alpha0 * 10000), "%07d" % int(alphaf * 100)
    return "MIE%s.%s-%s.%s-%s.%s-%s.%s-MU%d" % (crn[0], crn[1:], cin[0], cin[1:], ca0[0], ca0[1:], caf[:5], caf[5:], nbmu_gauss)


def write_mie_file(path, table, rn, in_, nbmu):
    """A Mie table (dict of Solver.mie) as the unformatted sequential file SOS_MIE writes (SOS_MIE.F:395, 919-922): header record RN,
    IN, ALPHAF (REAL*8), MIE_NBMU (INTEGER*4); per size parameter ALPHA, QEXT, QSCA (REAL*4), G (REAL*8), IMIE, QMIE, UMIE
    (-nbmu:nbmu) (REAL*4); 4-byte record markers of gfortran."""
    import struct
    nang = 2 * nbmu + 1
    n = table["g"].size
    rec = np.dtype([("l0", "<i4"), ("rec", "<f4", 3), ("g", "<f8"), ("i", "<f4", nang), ("q", "<f4", nang), ("u", "<f4", nang), ("l1", "<i4")])
    a = np.zeros(n, dtype=rec)
    a["l0"] = a["l1"] = 12 + 8 + 12 * nang
    a["rec"], a["g"], a["i"], a["q"], a["u"] = table["rec"], table["g"], table["imie"], table["qmie"], table["umie"]
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 28) + struct.pack("<dddi", rn, in_, table["alphaf"], nbmu) + struct.pack("<i", 28))
        f.write(a.tobytes())


def read_mie_file(path, nbmu):
    """The inverse of write_mie_file (what SOS_GRANU reads, SOS_AEROSOLS.F:4505-4532): dict for Solver.granu."""
    import struct
    raw = open(path, "rb").read()
    nang = 2 * nbmu + 1
    l0, = struct.unpack_from("<i", raw, 0)
    rn, in_, alphaf, nb = struct.unpack_from("<dddi", raw, 4)
    if l0 != 28 or nb != nbmu:
        raise ValueError("MIE file %s: header does not match %d angles" % (path, nbmu))
    rec = np.dtype([("l0", "<i4"), ("rec", "<f4", 3), ("g", "<f8"), ("i", "<f4", nang), ("q", "<f4", nang), ("u", "<f4", nang), ("l1", "<i4")])
    a = np.frombuffer(raw, dtype=rec, offset=36)
    return dict(rn=rn, in_=in_, alphaf=alphaf, rec=np.ascontiguousarray(a["rec"]), g=np.ascontiguousarray(a["g"]),
                imie=np.ascontiguousarray(a["i"]), qmie=np.ascontiguousarray(a["q"]), umie=np.ascontiguousarray(a["u"]))


def export_mie_cache(solver, dir_mie, nbmu_gauss, nbmu, xmu, components):
    """Computes the Mie table of every distinct (rn, in, alpha0, alphaf) of `components` on the device and stores it in dir_mie under
    the reference's file name, unless the file exists (SOS_AEROSOLS.F:1208-1237): the reference then finds its MIE files already
    calculated.  Returns the file names."""
    import os
    os.makedirs(dir_mie, exist_ok=True)
    names = []
    for key in dict.fromkeys((c[0], c[1], c[2], c[3]) for c in components):
        name = mie_file_name(nbmu_gauss, *key)
        names.append(name)
        path = os.path.join(dir_mie, name)
        if not os.path.exists(path):
            write_mie_file(path, solver.mie(nbmu, xmu, *key), key[0], key[1], nbmu)
    return names
