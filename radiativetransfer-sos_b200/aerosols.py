"""Host side of the aerosol chain: what SOS_AEROSOLS decides on the host before and after its Mie / size-
distribution / Legendre work -- which components a model has, their refractive indices rounded to the MIE file-name precision,
the size-parameter range of each Mie table, the mixture weights, the aerosol optical thickness at the simulation wavelength --
for a LIST of wavelengths, with the work itself done on the device by Solver.aerosols (one call for all wavelengths).

    mono-modal (IMOD = 0)        SOS_AEROSOLS.F:1158-1304   log-normal or Junge size distribution
    WMO models (IMOD = 1)        SOS_AEROSOLS.F:1309-1510   dust-like / water-soluble / oceanic / soot, SOS_INIT_PARAMWMO :3334-3556
    bimodal log-normal (IMOD=3)  SOS_AEROSOLS.F:1706-2123   volume concentrations given, or the coarse share of the optical
                                                            thickness at the reference wavelength (MODE_PARAM_BILND = 1 / 2)
    optical thickness at WA      SOS_PROC.F:2941-3063       TA = KMAT1(WA) / KMAT1(WAREF) * AOT_REF

Shettle & Fenn (IMOD = 2), external phase functions (4) and user mixtures (5) are not built.  Keyword parsing stays with the
caller.  No CPU fallback: the numbers come from Solver.aerosols, which needs the GPU."""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

ALPHA0 = 0.0001                                    # CTE_MIE_ALPHAMIN (SOS.h:116)
COEF_NRMAX = float(np.float32(0.0001))             # CTE_COEF_NRMAX (SOS.h:134), REAL*4 literal
WAMIN = float(np.float32(0.364))                   # CTE_WAMIN (SOS.h:70)
NOT_DEFINED = -999.0                               # CTE_NOT_DEFINED_VALUE_DBLE (SOS.h:78)
ALPHAMAX_WMO = (4000.0, 50.0, 800.0, 10.0)         # CTE_ALPHAMAX_WMO_DL / WS / OC / SO (SOS.h:122-125)
WMO_VOLUMES = {1: (0.70, 0.29, 0.0, 0.01), 2: (0.0, 0.05, 0.95, 0.0), 3: (0.17, 0.61, 0.0, 0.22)}   # SOS_AEROSOLS.F:1340-1353
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
    def cols(line, widths):
        out, pos = [], 0
        for skip, w in widths:
            pos += skip
            out.append(float(line[pos:pos + w].strip() or 0.0))
            pos += w
        return out
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
