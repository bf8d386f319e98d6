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

    external phase functions (4) SOS_AEROSOLS.F:2143-2279   -AER.ExtData: spline interpolation to the phase-function angles (host,
                                                            <= 200 nodes), SOS_DECOMPO_LEGENDRE on the device
    user mixture (IMOD = 5)      SOS_AEROSOLS.F:2283-2760   -AER.DefMixture: log-normal / Junge modes with their shares of the
                                                            optical thickness at the reference wavelength (up to 4 modes)

Keywords: frontend.aerosol_model maps -AER.* to these models.
No CPU fallback: the numbers come from Solver.aerosols, which needs the GPU."""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import math

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


@dataclass
class UserMixture:
    """-AER.Model 5 (-AER.DefMixture, SOS_AEROSOLS.F:2283-2760): up to four log-normal / Junge modes, each with its refractive
    index at the simulation wavelength and at the reference wavelength and its share of the optical thickness at the reference
    wavelength.  modes: list of (igranu, v1, v2, v3, rn_wa, in_wa, rn_waref, in_waref, aot_rate); igranu 1: v1 modal radius,
    v2 sigma; igranu 2: v1 minimal radius, v2 slope, v3 maximal radius (the arguments of SOS_GRANU)."""
    modes: List[tuple]
    waref: float


GAP_TOLER_SUM_RATES = float(np.float32(0.000001))   # CTE_GAP_TOLER_SUM_RATES (SOS.h:184)
MAX_MODES_DEVICE = 4                                # components per model of sosgpu_aer_model (the reference allows 20, SOS.h:178)


def read_mixture_file(path, waref):
    """The mixture definition file of -AER.DefMixture as SOS_AEROSOLS reads it (SOS_AEROSOLS.F:2297-2385): every value after the
    ':' of its line -- number of modes; per mode: LND | JUNGE, (modal radius, standard deviation) or (slope, minimal radius,
    maximal radius), real and imaginary index at the simulation wavelength, then at the reference wavelength, share of the optical
    thickness at the reference wavelength.  The shares must sum to 1 within CTE_GAP_TOLER_SUM_RATES and are then normalised."""
    with open(path) as f:
        vals = [ln.split(":", 1)[1].split()[0] for ln in f.read().splitlines() if ":" in ln]
    num = lambda t: float(t.replace("D", "E").replace("d", "e"))
    it = iter(vals)
    try:
        n = int(num(next(it)))
        modes = []
        for _ in range(n):
            kind = next(it).strip("'\"")
            if kind == "LND":
                v1, v2, v3, ig = num(next(it)), num(next(it)), 0.0, 1
            elif kind == "JUNGE":
                v2, v1, v3, ig = num(next(it)), num(next(it)), num(next(it)), 2
            else:
                raise ValueError("%s: size distribution %r is neither LND nor JUNGE (SOS_AEROSOLS error 962)" % (path, kind))
            rn_wa, in_wa, rn_ref, in_ref, rate = (num(next(it)) for _ in range(5))
            modes.append((ig, v1, v2, v3, rn_wa, in_wa, rn_ref, in_ref, rate))
    except StopIteration:
        raise ValueError("%s: the mixture file ends before its %d modes are described (SOS_AEROSOLS error 961)" % (path, n))
    tot = 0.0
    for m in modes:
        tot = tot + m[8]
    if abs(tot - 1.0) > GAP_TOLER_SUM_RATES:
        raise ValueError("%s: the shares of the optical thickness sum to %r, not 1 (SOS_AEROSOLS error 963)" % (path, tot))
    return UserMixture(modes, float(waref))


def _mixture_component(mode, wa, waref):
    """Component of one mode of a UserMixture at wavelength wa: indices of the reference wavelength when wa == waref, the table
    sized for CTE_WAMIN (SOS_AEROSOLS.F:2626-2650); no rounding of the indices (unlike the other models)."""
    ig, v1, v2, v3, rn_wa, in_wa, rn_ref, in_ref, _ = mode
    rn, in_ = (rn_ref, in_ref) if wa == waref else (rn_wa, in_wa)
    af = alphaf_of(lnd_rmax(v1, v2) if ig == 1 else v3, WAMIN)
    if ALPHA0 > af or af >= 1e5:
        raise ValueError("size-parameter range of the Mie table out of bounds (SOS_AEROSOLS error 1009)")
    return (rn, in_, ALPHA0, af, ig, v1, v2, v3, wa)


# ---- -AER.Model 4: phase functions from an external file (SOS_AEROSOLS.F:2143-2279) ----
MAXNB_ANG_EXT = 200                                  # CTE_MAXNB_ANG_EXT (SOS.h:101)
_BIG = float(np.float32(0.99e30))                    # .99E30 (REAL*4 literal) of SOS_SPLINE


def spline(x, y, dy1, dyn):
    """SOS_SPLINE (SOS_AEROSOLS.F:4952-5002): second derivatives of the cubic spline through (x, y), x ascending, with the first
    derivatives dy1, dyn at the ends; the reference's statements in their order (host work on <= 200 nodes)."""
    n = len(x)
    d2, u = [0.0] * n, [0.0] * n
    if dy1 > _BIG:
        d2[0], u[0] = 0.0, 0.0
    else:
        d2[0] = -0.5
        u[0] = (3. / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - dy1)
    for k in range(1, n - 1):
        sig = (x[k] - x[k - 1]) / (x[k + 1] - x[k - 1])
        p = sig * d2[k - 1] + 2.
        d2[k] = (sig - 1.) / p
        u[k] = (6. * ((y[k + 1] - y[k]) / (x[k + 1] - x[k]) - (y[k] - y[k - 1]) / (x[k] - x[k - 1])) / (x[k + 1] - x[k - 1]) - sig * u[k - 1]) / p
    if dyn > _BIG:
        qn, un = 0.0, 0.0
    else:
        qn = 0.5
        un = (3. / (x[n - 1] - x[n - 2])) * (dyn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]))
    d2[n - 1] = (un - qn * u[n - 2]) / (qn * d2[n - 2] + 1.)
    for k in range(n - 2, -1, -1):
        d2[k] = d2[k] * d2[k + 1] + u[k]
    return d2


def splint(x, y, d2, xval):
    """SOS_SPLINT (SOS_AEROSOLS.F:5042-5105): the spline at xval (bisection for the interval, then the cubic)."""
    klo, khi = 1, len(x)
    while khi - klo > 1:
        k = (khi + klo) // 2
        if x[k - 1] > xval:
            khi = k
        else:
            klo = k
    h = x[khi - 1] - x[klo - 1]
    if h == 0.0:
        raise ValueError("SPLINT interpolation: two nodes share an abscissa (bad X table input)")
    a = (x[khi - 1] - xval) / h
    b = (xval - x[klo - 1]) / h
    return a * y[klo - 1] + b * y[khi - 1] + ((a * (a * a) - a) * d2[klo - 1] + (b * (b * b) - b) * d2[khi - 1]) * (h * h) / 6.


def interpo_splint(xin, yin, xout):
    """SOS_INTERPO_SPLINT (SOS_AEROSOLS.F:4822-4925): nodes sorted by ascending abscissa, end slopes from the end intervals."""
    order = sorted(range(len(xin)), key=lambda i: xin[i])
    x, y = [float(xin[i]) for i in order], [float(yin[i]) for i in order]
    if any(b == a for a, b in zip(x, x[1:])):                     # the reference divides by zero, then SOS_SPLINT returns IER = -1
        raise ValueError("spline interpolation: two nodes share an abscissa (SOS_SPLINT: bad X table input)")
    dy1 = (y[1] - y[0]) / (x[1] - x[0])
    dyn = (y[-1] - y[-2]) / (x[-1] - x[-2])
    d2 = spline(x, y, dy1, dyn)
    return np.array([splint(x, y, d2, float(v)) for v in xout])


def read_external_data(path):
    """The file of -AER.ExtData as SOS_AEROSOLS reads it (:2148-2172): extinction and scattering cross sections and the number of
    angles after the ':' of the first three lines, one header line, then rows ANGLE, F11, -F12/F11, F22/F11, F33/F11.
    Returns (kmat1, kmat2, mu[n], f11[n], f12[n], f22[n], f33[n])."""
    num = lambda t: float(t.replace("D", "E").replace("d", "e"))
    with open(path) as f:
        lines = f.read().splitlines()
    try:
        kmat1, kmat2 = (num(lines[i].split(":", 1)[1].split()[0]) for i in (0, 1))
        n = int(num(lines[2].split(":", 1)[1].split()[0]))
        if n > MAXNB_ANG_EXT:
            raise ValueError("%s: %d angles, more than CTE_MAXNB_ANG_EXT = %d (SOS_AEROSOLS error 950)" % (path, n, MAXNB_ANG_EXT))
        rows = np.array([[num(v) for v in ln.replace(",", " ").split()[:5]] for ln in lines[4:4 + n]])
    except (IndexError, ValueError) as e:
        if "CTE_MAXNB_ANG_EXT" in str(e):
            raise
        raise ValueError("%s: not in the layout of the external phase-function file (SOS_AEROSOLS error 941)" % path)
    if rows.shape != (n, 5):
        raise ValueError("%s: the file ends before its %d angles (SOS_AEROSOLS error 942)" % (path, n))
    ang, f11 = rows[:, 0], rows[:, 1]
    mu = np.array([math.cos(a * _PI / 180.0) for a in ang])          # libm's cos, as the reference's DCOS
    return kmat1, kmat2, mu, f11, -rows[:, 2] * f11, rows[:, 3] * f11, rows[:, 4] * f11


def external_data(solver, path, nbmu, xmu, xhr, os_nb, itronc, wa, aot):
    """-AER.Model 4: external phase functions -> spline interpolation to the phase-function angles (host, <= 200 nodes) ->
    SOS_DECOMPO_LEGENDRE on the device -> AerosolOptics (the common tail of SOS_AEROSOLS, :2775-2781).  The optical thickness is
    aot itself: SOS_PROC requires the simulation wavelength to be the reference one (error 2331)."""
    kmat1, kmat2, mu, f11, f12, f22, f33 = read_external_data(path)
    p11, p12, p22, p33 = (interpo_splint(mu, f, xmu) for f in (f11, f12, f22, f33))
    d = solver.decompo_legendre(itronc, nbmu, xmu, xhr, os_nb, p11, p12, p22, p33)
    if d["ier"] != 0:
        raise RuntimeError("SOS_DECOMPO_LEGENDRE failed on the external phase functions (IER = %d)" % d["ier"])
    ct = d["coef_tronca"]
    piz = kmat2 / kmat1
    piztr = piz * (1. - ct / 2.) / (1. - piz * ct / 2.)
    asym = ct / 2. + (1. - ct / 2.) * d["beta11"][1] / 3.
    return AerosolOptics(float(wa), kmat1, kmat2, piz, piztr, ct, asym, int(d["itronc"]), d["alp"], d["beta11"], d["gamma12"], d["zeta"], float(aot))


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
        elif isinstance(model, UserMixture):
            if len(model.modes) > MAX_MODES_DEVICE:
                raise NotImplementedError("a mixture of %d modes: the device mixes at most %d (the reference 20)"
                                          % (len(model.modes), MAX_MODES_DEVICE))
            if bilnd_weights is None:
                raise ValueError("UserMixture: pass bilnd_weights=reference_weights(...)")
            for m in model.modes:
                p.components.append(_mixture_component(m, wa, model.waref))
            p.models.append((len(model.modes), list(range(n0, n0 + len(model.modes))), list(bilnd_weights), itronc))
        else:
            raise TypeError("unsupported aerosol model %r" % (model,))
    return p


def in_pos(model, wa):
    return _at(model.coarse_in, wa) > 0.0 or _at(model.fine_in, wa) > 0.0


def reference_weights(solver, nbmu, xmu, xhr, model, waref, aot_ref):
    """MODE_PARAM_BILND = 2 (SOS_AEROSOLS.F:1806-2058): the extinction cross sections of the two modes at the reference
    wavelength give CVI(coarse) = rtauct * AOT_REF / KMAT1c, CVI(fine) = (1 - rtauct) * AOT_REF / KMAT1f, then normalised."""
    if isinstance(model, UserMixture):
        # SOS_AEROSOLS.F:2383, 2452-2465, 2486-2594: AOT of each mode = AOT_REF * share (shares normalised when their sum is not
        # exactly 1), COEF_ALPHA = AOT / KMAT1 of the mode at the reference wavelength, then normalised
        comps = [_mixture_component(m, waref, waref) for m in model.modes]
        o = solver.aerosols(nbmu, xmu, xhr, comps, [], 2, want_phase=False)
        if o["comp_ier"].any():
            raise RuntimeError("SOS_GRANU failed at the reference wavelength")
        aot = [aot_ref * m[8] for m in model.modes]
        tot = 0.0
        for m in model.modes:
            tot = tot + m[8]
        if tot != 1.0:
            aot = [a / tot for a in aot]
        ca = [a / o["comp_k"][i, 0] for i, a in enumerate(aot)]
        som = 0.0
        for c in ca:
            som = som + c
        return [c / som for c in ca]
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
    if (isinstance(model, BimodalLnd) and model.rtauct is not None) or isinstance(model, UserMixture):
        if not with_ref:
            raise ValueError("a mixture defined by shares of the optical thickness needs waref and aot_ref")
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
