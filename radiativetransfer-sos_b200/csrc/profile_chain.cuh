// profile_chain.cuh -- the per-term profile chain that precedes every term-solve, as functions one GPU thread
// runs for one (wavelength, CKD term):
//   absorption optical thickness of the 49 gas layers   SOS_ABSPROFILE        (SOS_ABSPROFILE.F:184-425)
//   CKD coefficient k_i(P, T, [H2O]) of one layer        COEFF_ABS_CKD         (SOS_SUB_TRS.F:171-393)
//   linear / cubic-spline interpolation                  SOS_INTERPOL, SOS_INTERPO_SPLINT, SOS_SPLINE, SOS_SPLINT
//                                                        (SOS_AEROSOLS.F:3844, 4822-5130)
//   discretisation of the atmosphere in optical depth    SOS_PROFILE, SOS_DISC (SOS_PROFIL.F:224-1156, 1210-1332)
//   the text hop PROFIL_TMP between SOS_PROFILE and SOS  WRITE format 20 (2X,I5,F10.5,3(E15.8), SOS_PROFIL.F:1095-1097,
//                                                        1152) read back by SOS.F:511-516: values reach the solver rounded to
//                                                        5 decimals (altitude) / 8 significant digits (H, PCAER, PCMOL)
// The statement order, the REAL*4 literals of SOS.h (CTE_TCOUCHE 0.005, CTE_DELTA_Z 0.05 ... are single-precision constants
// promoted to double) and the comparison operators follow the reference, because the number of levels NT is an integer result.
// The functions are __host__ __device__ so that tests/ can also compile them for the host and step through them against the
// reference library without a GPU; the library itself only ever runs them inside the kernels of sosgpu_profile.cu.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PC_HD __host__ __device__ inline
#define PC_HDM __host__ __device__
#else
#define PC_HD static inline
#define PC_HDM
#endif

// dimensions of SOS.h:246-282
#define PC_NBABS 8
#define PC_NLEV 50
#define PC_NCOL 13
#define PC_NWVL 50
#define PC_NAI 5
#define PC_NTMAX 9
#define PC_NPMAX 31
#define PC_NCMAX 12
#define PC_OS_NT 600
#define PC_OS_NT_MIN 100
#define PC_LEVELS (PC_OS_NT + 1)
// SOS.h:197-301 -- literals without a D exponent are REAL*4
#define PC_TOA_ALT 120.0
#define PC_TCOUCHE ((double)0.005f)
#define PC_FIRST_LAYER ((double)0.0002f)
#define PC_DELTA_Z ((double)0.05f)
#define PC_THRESHOLD_DZ ((double)0.001f)
#define PC_DZTRANSI ((double)0.010f)
#define PC_PROFIL_MIN_NBC 3
#define PC_TAUABS_MAX ((double)999.f)
#define PC_THRESHOLD_TAUABS ((double)1.5f)

// Fortran column-major indices (1-based) of the CKD tables as READ_CKD_COEFF fills them
#define PC_KI(it, ip, ik, nabs, lamb) \
  ((size_t)((it) - 1) + PC_NTMAX * ((size_t)((ip) - 1) + PC_NPMAX * ((size_t)((ik) - 1) + PC_NAI * ((size_t)((nabs) - 1) + PC_NBABS * (size_t)((lamb) - 1)))))
#define PC_KI_H2O(it, ip, ic, ik, lamb) \
  ((size_t)((it) - 1) + PC_NTMAX * ((size_t)((ip) - 1) + PC_NPMAX * ((size_t)((ic) - 1) + PC_NCMAX * ((size_t)((ik) - 1) + PC_NAI * (size_t)((lamb) - 1)))))
#define PC_NEXP(k, lamb) ((k) - 1 + PC_NBABS * ((lamb) - 1))
#define PC_USER(j, col) ((j) - 1 + PC_NLEV * ((col) - 1))        // USERPROFIL(NLEVEL, NBCOL)
#define PC_RO(k, j) ((k) - 1 + PC_NBABS * ((j) - 1))             // RO(NBABS, NLEVEL)

struct PcCkd {                 // device / host view of the tables of one CKD file set
  int nb_temp, nb_pres, nb_conc;
  const double *tab_temp, *tab_pres, *tab_conc;   // [NTMAX], [NPMAX], [NCMAX]
  const int *nexp;                                // NEXP(NBABS, NWVL)
  const double *ki, *ki_h2o;                      // KDIS_KI, KDIS_KI_H2O
};

// SOS_INTERPOL (SOS_AEROSOLS.F:3844)
PC_HD double pc_interpol(double y1, double y2, double x1, double x2, double x) { return ((y2 - y1) / (x2 - x1)) * (x - x2) + y2; }

// SOS_INTERPO_SPLINT for one abscissa (SOS_AEROSOLS.F:4822-4950): sort, natural end slopes from the end chords, SOS_SPLINE
// (:4952-5040), SOS_SPLINT (:5042-5130).  n <= PC_NTMAX.  Returns 0, or -1 for a repeated abscissa.
PC_HD int pc_interpo_splint(int n, const double *xin, const double *yin, double xval, double *yval)
{
  double x[PC_NTMAX], y[PC_NTMAX], d2[PC_NTMAX], u[PC_NTMAX];
  for (int j = 0; j < n; ++j) { x[j] = xin[j]; y[j] = yin[j]; }
  for (int j = 0; j < n; ++j)
    for (int k = j + 1; k < n; ++k)
      if (x[j] > x[k]) { double v = x[j]; x[j] = x[k]; x[k] = v; v = y[j]; y[j] = y[k]; y[k] = v; }
  for (int j = 0; j < n; ++j) d2[j] = 0.0;
  const double dy1 = (y[1] - y[0]) / (x[1] - x[0]);
  const double dyn = (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]);
  const double big = (double).99e30f;
  if (dy1 > big) { d2[0] = 0.0; u[0] = 0.0; }
  else { d2[0] = -0.5; u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - dy1); }
  for (int k = 1; k <= n - 2; ++k) {
    const double sig = (x[k] - x[k - 1]) / (x[k + 1] - x[k - 1]);
    const double p = sig * d2[k - 1] + 2.0;
    d2[k] = (sig - 1.0) / p;
    u[k] = (6.0 * ((y[k + 1] - y[k]) / (x[k + 1] - x[k]) - (y[k] - y[k - 1]) / (x[k] - x[k - 1])) / (x[k + 1] - x[k - 1]) - sig * u[k - 1]) / p;
  }
  double qn, un;
  if (dyn > big) { qn = 0.0; un = 0.0; }
  else { qn = 0.5; un = (3.0 / (x[n - 1] - x[n - 2])) * (dyn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2])); }
  d2[n - 1] = (un - qn * u[n - 2]) / (qn * d2[n - 2] + 1.0);
  for (int k = n - 2; k >= 0; --k) d2[k] = d2[k] * d2[k + 1] + u[k];
  int klo = 1, khi = n;                                          // SOS_SPLINT, 1-based as there
  while (khi - klo > 1) {
    const int k = (khi + klo) / 2;
    if (x[k - 1] > xval) khi = k; else klo = k;
  }
  const double h = x[khi - 1] - x[klo - 1];
  if (h == 0.0) return -1;
  const double a = (x[khi - 1] - xval) / h, b = (xval - x[klo - 1]) / h;
  *yval = a * y[klo - 1] + b * y[khi - 1] + ((a * a * a - a) * d2[klo - 1] + (b * b * b - b) * d2[khi - 1]) * (h * h) / 6.0;
  return 0;
}

// COEFF_ABS_CKD (SOS_SUB_TRS.F:171-393).  prs, tmp, conc are clamped in place as there (the caller's copies change).
// lamb and ik are 1-based.  Returns 0 or the reference's error label (900, 910, 915, 916, 920, 921, 922, 923).
PC_HD int pc_coeff_abs_ckd(const PcCkd &c, int nabs, int lamb, int ik, double &prs, double &tmp, double &conc, double *xk_out)
{
  const int nt = c.nb_temp, np = c.nb_pres, nc = c.nb_conc;
  if (c.tab_temp[nt - 1] < c.tab_temp[0]) return 900;
  if (tmp > c.tab_temp[nt - 1]) tmp = c.tab_temp[nt - 1];
  if (tmp < c.tab_temp[0]) tmp = c.tab_temp[0];
  if (c.tab_pres[np - 1] < c.tab_pres[0]) return 910;
  if (prs <= c.tab_pres[0]) { *xk_out = 0.0; return 0; }
  if (prs > c.tab_pres[np - 1]) prs = c.tab_pres[np - 1];
  if (c.tab_conc[nc - 1] < c.tab_conc[0]) return 915;
  if (conc > c.tab_conc[nc - 1]) conc = c.tab_conc[nc - 1];
  if (conc < c.tab_conc[0]) conc = c.tab_conc[0];
  int ip = 1;
  while (c.tab_pres[ip - 1] <= prs && ip < np) ++ip;
  --ip;
  if (ip == np) return 920;
  double xki[PC_NTMAX];
  const double px1 = c.tab_pres[ip - 1], px2 = c.tab_pres[ip];
  if (nabs == 1) {
    int ic = 1;
    while (c.tab_conc[ic - 1] <= conc && ic < nc) ++ic;
    --ic;
    if (ic == nc) return 916;
    const double cx1 = c.tab_conc[ic - 1], cx2 = c.tab_conc[ic];
    for (int it = 1; it <= nt; ++it) {                            // only the two pressure rows the next step reads
      const double lo = pc_interpol(c.ki_h2o[PC_KI_H2O(it, ip, ic, ik, lamb)], c.ki_h2o[PC_KI_H2O(it, ip, ic + 1, ik, lamb)], cx1, cx2, conc);
      const double hi = pc_interpol(c.ki_h2o[PC_KI_H2O(it, ip + 1, ic, ik, lamb)], c.ki_h2o[PC_KI_H2O(it, ip + 1, ic + 1, ik, lamb)], cx1, cx2, conc);
      xki[it - 1] = pc_interpol(lo, hi, px1, px2, prs);
    }
  } else {
    for (int it = 1; it <= nt; ++it)
      xki[it - 1] = pc_interpol(c.ki[PC_KI(it, ip, ik, nabs, lamb)], c.ki[PC_KI(it, ip + 1, ik, nabs, lamb)], px1, px2, prs);
  }
  double xk = 0.0;
  if (pc_interpo_splint(nt, c.tab_temp, xki, tmp, &xk) != 0) return 922;
  if (xk < 0.0) {                                                 // spline overshoot: linear in T instead
    int it = 1;
    while (c.tab_temp[it - 1] <= tmp && it < nt) ++it;
    --it;
    if (it == nt) return 921;
    xk = pc_interpol(xki[it - 1], xki[it], c.tab_temp[it - 1], c.tab_temp[it], tmp);
    if (xk < 0.0) return 923;
  }
  *xk_out = xk;
  return 0;
}

// One layer of SOS_ABSPROFILE (SOS_ABSPROFILE.F:330-364): optical thickness of layer j (1 = top) summed over the 8 gases in
// the reference's order.  ik[8] are the exponential indices IK1..IK8.
PC_HD int pc_absprofile_layer(const PcCkd &c, const double *userprofil, const double *ro, int lamb, const int *ik, int j, double *tau)
{
  double prs = (userprofil[PC_USER(PC_NLEV - j, 2)] + userprofil[PC_USER(PC_NLEV - j + 1, 2)]) / 2.0;
  double tmp = (userprofil[PC_USER(PC_NLEV - j, 3)] + userprofil[PC_USER(PC_NLEV - j + 1, 3)]) / 2.0;
  double conc = (userprofil[PC_USER(PC_NLEV - j, 4)] + userprofil[PC_USER(PC_NLEV - j + 1, 4)]) / 2.0;
  conc = conc * 1.0e-06;
  double t = 0.0;
  for (int k = 1; k <= PC_NBABS; ++k) {
    double xk = 0.0;
    if (c.nexp[PC_NEXP(k, lamb)] >= 1) {
      const int rc = pc_coeff_abs_ckd(c, k, lamb, ik[k - 1], prs, tmp, conc, &xk);
      if (rc != 0) return rc;
    }
    t = t + xk * ro[PC_RO(k, PC_NLEV - j)];
  }
  *tau = t;
  return 0;
}

// The running transmission of SOS_ABSPROFILE (:366-379): tauabs[0..49] from the 49 layer thicknesses.
PC_HD void pc_absprofile_scan(const double *tau_layer, double *tauabs)
{
  double trs = 1.0;
  tauabs[0] = 0.0;
  for (int j = 1; j <= PC_NLEV - 1; ++j) {
    trs = trs * exp(-tau_layer[j - 1]);
    tauabs[j] = (trs > 0.0) ? -log(trs) : PC_TAUABS_MAX;
  }
}

// Index j (2..50, 1-based) of the first gas level at or below altitude z: the linear search `J=2; DO WHILE(Z.LT.ALTABS(J))` of
// SOS_PROFIL.F as a binary search.  ALTABS is strictly descending (the host checks it), so the first j with z >= ALTABS(j) is
// where the predicate flips; z below ALTABS(50) (never: altitudes are >= 0 = ALTABS(50)) stops at 50 instead of running off.
PC_HD int pc_abs_bracket(const double *altabs, double z)
{
  if (!(z < altabs[1])) return 2;
  int lo = 2, hi = PC_NLEV;                                       // z < altabs(lo) holds; answer in (lo, hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (z < altabs[mid - 1]) lo = mid; else hi = mid;
  }
  return hi;
}

struct PcColumn {              // what the optical depth at an altitude depends on
  double ta, ha, tr, hr;
  const double *tabs, *altabs;
  double tg_zlim;              // SOS_DISC: gas is interpolated only when TG_ZLIM > 0
};

// total optical depth at altitude z as SOS_DISC forms it (SOS_PROFIL.F:1285-1306)
PC_HD double pc_disc_tau(const PcColumn &c, double zmoy)
{
  double tg;
  if (c.tg_zlim > 0.0) {
    const int j = pc_abs_bracket(c.altabs, zmoy);
    double zz;
    if (zmoy > c.altabs[0]) zz = 0.0;
    else zz = (zmoy - c.altabs[j - 2]) / (c.altabs[j - 1] - c.altabs[j - 2]);
    tg = (1 - zz) * c.tabs[j - 2] + zz * c.tabs[j - 1];
  } else tg = 0.0;
  return c.ta * exp(-zmoy / c.ha) + c.tr * exp(-zmoy / c.hr) + tg;
}
// ... and as the search of the first level forms it: without gas (SOS_PROFIL.F:428) / with gas (:661-678)
PC_HD double pc_first_tau_ng(const PcColumn &c, double z) { return c.tr * exp(-z / c.hr) + c.ta * exp(-z / c.ha); }
PC_HD double pc_first_tau_gas(const PcColumn &c, double z)
{
  const int j = pc_abs_bracket(c.altabs, z);
  double vg;
  if (z <= c.altabs[0]) {
    const double zz = (z - c.altabs[j - 2]) / (c.altabs[j - 1] - c.altabs[j - 2]);
    vg = (1 - zz) * c.tabs[j - 2] + zz * c.tabs[j - 1];
  } else vg = 0.0;
  const double vr = c.tr * exp(-z / c.hr), va = c.ta * exp(-z / c.ha);
  return vr + va + vg;
}

#define PC_DISC_TOL ((double).000001f)
// One step of SOS_DISC's bisection at the candidate zmoy: bit 0 = stop here (:1311-1312), bit 1 = continue with ZMIN = ZMOY
// (else ZMAX = ZMOY, :1316-1320).
PC_HD int pc_disc_step(const PcColumn &c, double ti, double zmoy)
{
  const double tzmoy = pc_disc_tau(c, zmoy);
  const double xd = fabs(ti - tzmoy);
  int r = 0;
  if (xd < PC_DISC_TOL) r |= 1;
  if (zmoy == 0.0) r |= 1;
  if ((ti - tzmoy) < 0.0) r |= 2;
  return r;
}

// SOS_DISC (SOS_PROFIL.F:1210-1332): altitude whose total optical depth is tim1 + dt, by bisection between zlim and zmax_init.
PC_HD double pc_disc(double dt, const PcColumn &c, double tim1, double zmax_init, double zlim)
{
  const double ti = tim1 + dt;
  double zmax = zmax_init, zmin = zlim, zmoy;
  for (;;) {
    zmoy = (zmax + zmin) / 2.0;
    const int r = pc_disc_step(c, ti, zmoy);
    if (r & 1) break;
    if (r & 2) zmin = zmoy; else zmax = zmoy;
  }
  return zmoy;
}

// The same bisection, five levels at a time: the 31 candidates of a depth-5 subtree are independent once the interval at its
// root is known, so 31 lanes evaluate them at once and the path the serial loop would have taken is read off their results.
// Node n (heap order, 1 = root, children 2n / 2n+1) sits at the end of the path spelled by the bits of n below its leading
// one, most significant first: bit 1 = "ZMIN = ZMOY" at that ancestor.  The midpoints are formed with the reference's own
// expression at every ancestor, so a node's candidate is the very double the serial loop would test there.
PC_HD void pc_tree_interval(int n, int depth, double zmin, double zmax, double *lo, double *hi)
{
  double l = zmin, h = zmax;
  for (int b = depth - 1; b >= 0; --b) {
    const double mid = (h + l) / 2.0;
    if ((n >> b) & 1) l = mid; else h = mid;
  }
  *lo = l; *hi = h;
}
PC_HD int pc_tree_depth(int n) { int d = 0; while ((n >> (d + 1)) != 0) ++d; return d; }
// stop / direction masks (bit n = node n) -> node where the serial loop stops (returns 1) or the depth-5 virtual node (32..63)
// whose interval is the next root (returns 0)
PC_HD int pc_tree_walk(unsigned stop, unsigned dir, int *node)
{
  int m = 1;
  for (int lvl = 0; lvl < 5; ++lvl) {
    if ((stop >> m) & 1u) { *node = m; return 1; }
    m = 2 * m + (int)((dir >> m) & 1u);
  }
  *node = m;
  return 0;
}

// Search strategies of pc_profile.  PcSerial is the reference's control flow statement by statement; the kernels use a
// warp-wide strategy built on pc_tree_* (sosgpu_profile.cu) and tests/profile_host.cpp holds a host emulation of that one.
struct PcSerial {
  PC_HDM double disc(double dt, const PcColumn &c, double tim1, double zmax_init, double zlim) const { return pc_disc(dt, c, tim1, zmax_init, zlim); }
  // first level: Z steps down by CTE_DELTA_Z until the optical depth reaches t_first (SOS_PROFIL.F:424-429, 657-679)
  PC_HDM void first(bool gas, const PcColumn &c, double t_first, double *z, double *dtau) const
  {
    double d = 0.0, zz = *z;
    while (d < t_first) {
      zz = zz - PC_DELTA_Z;
      d = gas ? pc_first_tau_gas(c, zz) : pc_first_tau_ng(c, zz);
    }
    *z = zz; *dtau = d;
  }
};

// ------------------------------------------------------------------------------------------------
// The text hop.  A value written with Ew.8 and read back is the double nearest to its 8-significant-digit decimal, the
// decimal itself being the correctly rounded one (gfortran formats through the C library; ties to even on the exact
// binary value).  x*10^k is formed exactly as a double-double with one FMA, so the rounding decision is exact; the
// way back is one correctly rounded division or multiplication by an exactly representable power of ten (|k| <= 22),
// which is what strtod's fast path returns, and a double-double division beyond that (|x| < 1e-15: never an optical depth).
PC_HD double pc_pow10(int k)                                       // exact for 0 <= k <= 22: every partial product is representable
{
  double p = 1.0;
  for (int i = 0; i < k; ++i) p *= 10.0;
  return p;
}
// round (p + e) to the nearest integer, ties to even: p a double with |p| < 2^52 and e the exact remainder of the product that
// made it (|e| <= ulp(p)/2).  d = (p - floor p) - 1/2 is exact and a multiple of ulp(p), so e only decides when d == 0.
PC_HD double pc_rint_dd(double p, double e)
{
  const double r = floor(p);
  const double d = (p - r) - 0.5;
  if (d > 0.0) return r + 1.0;
  if (d < 0.0) return r;
  if (e > 0.0) return r + 1.0;
  if (e < 0.0) return r;
  return (fmod(r, 2.0) != 0.0) ? r + 1.0 : r;
}
PC_HD double pc_round_e8(double x)                                 // WRITE E15.8 -> READ
{
  if (x == 0.0 || !(fabs(x) < 1e300)) return x;
  const double ax = fabs(x);
  int e10 = (int)floor(log10(ax));                                // 10^e10 <= ax < 10^(e10+1), corrected below
  double r = 0.0;
  for (int pass = 0; pass < 3; ++pass) {
    const int k = 7 - e10;                                        // scale so that the integer part has 8 digits
    double p, e;
    if (k >= 0 && k <= 22) { const double s = pc_pow10(k); p = ax * s; e = fma(ax, s, -p); }
    else if (k > 22 && k <= 44) {
      const double s2 = pc_pow10(k - 22);
      const double p1 = ax * 1e22, e1 = fma(ax, 1e22, -p1);
      p = p1 * s2; e = fma(p1, s2, -p) + e1 * s2;
    } else if (k < 0 && k >= -22) { const double s = pc_pow10(-k); p = ax / s; e = -fma(p, s, -ax) / s; }
    else return x;                                                // outside any profile quantity: left unrounded
    if (p + e < 1e7) { --e10; continue; }
    if (p + e >= 1e8) { ++e10; continue; }
    r = pc_rint_dd(p, e);
    break;
  }
  if (r >= 1e8) { r = 1e7; ++e10; }
  const int k = 7 - e10;
  double v;
  if (k >= 0 && k <= 22) v = r / pc_pow10(k);
  else if (k < 0) v = r * pc_pow10(-k);
  else {                                                          // r / (1e22 * 10^(k-22)), the divisor carried exactly enough
    const double s2 = pc_pow10(k - 22);
    const double d = 1e22 * s2, de = fma(1e22, s2, -d);            // divisor = d + de
    const double q0 = r / d;
    const double rem = fma(-q0, d, r) - q0 * de;
    v = q0 + rem / d;
  }
  return x < 0.0 ? -v : v;
}
PC_HD double pc_round_f5(double x)                                 // WRITE F10.5 -> READ
{
  const double ax = fabs(x);
  if (!(ax < 1e9)) return x;
  const double p = ax * 1e5, e = fma(ax, 1e5, -p);
  const double v = pc_rint_dd(p, e) / 1e5;
  return x < 0.0 ? -v : v;
}

// ------------------------------------------------------------------------------------------------
// SOS_PROFILE for one term.  Outputs zprof, h, pcaer, pcmol [0..nt] (unrounded; the caller applies the text hop) and NT.
// scratch: PC_LEVELS doubles (the altitudes of the profile without gas, kept for the profile with gas).
// Returns 0, or the reference's error label (940, 1010, 1020), or 9600 when the level count would leave the reference's arrays
// (CTE_OS_NT = 600: the reference writes out of bounds there; this implementation refuses).
template <class Search>
PC_HD int pc_profile(const Search &srch, int iprofil, double tr, double hr, double ta, double ha, double zmin_in, double zmax_in,
                     int absprofil, const double *altabs, const double *tabs, double *zng, double *zprof, double *h, double *pcaer,
                     double *pcmol, int *nt_out)
{
  if (iprofil != 1 && iprofil != 2) return 940;
  int nt = 0;
  if (iprofil == 1) {
    // ---- step 1: profile without gaseous absorption (:336-489) ----
    double ttot = tr + ta;
    int nt_ng;
    double t_layer_ng, t_first_ng;
    if ((ttot / PC_OS_NT_MIN) <= PC_FIRST_LAYER) {
      nt_ng = PC_OS_NT_MIN;
      t_layer_ng = ttot / nt_ng;
      t_first_ng = t_layer_ng;
    } else if ((ttot / PC_OS_NT_MIN) < PC_TCOUCHE) {
      nt_ng = PC_OS_NT_MIN + 1;
      t_first_ng = PC_FIRST_LAYER;
      t_layer_ng = (ttot - t_first_ng) / PC_OS_NT_MIN;
    } else {
      t_first_ng = PC_FIRST_LAYER;
      nt_ng = (int)((ttot - t_first_ng) / PC_TCOUCHE);
      t_layer_ng = (ttot - t_first_ng) / nt_ng;
      nt_ng = nt_ng + 1;
    }
    if (nt_ng > PC_OS_NT) return 9600;
    const bool no_gas = (absprofil == 7) || (tabs[PC_NLEV - 1] == 0.0);
    // Without gas the NG profile IS the result: it is written straight into the outputs; with gas only its altitudes are kept.
    double *hng = h, *pa = pcaer, *pm = pcmol;
    if (ta == 0.0) {
      for (int i = 0; i <= nt_ng; ++i) {
        double hm;
        if (i == 0) hm = 0.0;
        else if (i == 1) hm = t_first_ng;
        else hm = (i - 1) * t_layer_ng + t_first_ng;
        hng[i] = hm; pm[i] = (double)1.000f; pa[i] = (double)0.000f;
        zng[i] = (i == 0) ? PC_TOA_ALT : hr * log(tr / hm);
      }
    } else {
      zng[0] = PC_TOA_ALT; hng[0] = 0.0;
      const PcColumn cng = {ta, ha, tr, hr, tabs, altabs, 0.0};
      double dtau = 0.0, z = PC_TOA_ALT;
      srch.first(false, cng, t_first_ng, &z, &dtau);
      zng[1] = z;
      double vr = tr * exp(-z / hr), va = ta * exp(-z / ha);
      double hmol_prev = vr, haer_prev = va;
      hng[1] = dtau; pm[1] = vr / dtau; pa[1] = va / dtau;
      pm[0] = pm[1]; pa[0] = pa[1];
      for (int i = 2; i <= nt_ng - 1; ++i) {
        z = srch.disc(t_layer_ng, cng, hng[i - 1], zng[1], 0.0);
        zng[i] = z;
        vr = tr * exp(-z / hr); va = ta * exp(-z / ha);
        const double hm = vr, hae = va;
        hng[i] = vr + va;
        vr = vr - hmol_prev; va = va - haer_prev;
        pm[i] = vr / (vr + va); pa[i] = va / (vr + va);
        hmol_prev = hm; haer_prev = hae;
      }
      zng[nt_ng] = 0.0;
      hng[nt_ng] = tr + ta;
      vr = tr - hmol_prev; va = ta - haer_prev;
      pm[nt_ng] = vr / (vr + va); pa[nt_ng] = va / (vr + va);
    }
    if (no_gas) {
      nt = nt_ng;
      for (int i = 0; i <= nt; ++i) zprof[i] = zng[i];
      // H(I) = Hmol(I) + Haer(I) (:497) is H_NG(I) at every level: DTAU of level 1 and TR + TA of the ground are the same sums
    } else {
      // ---- step 2: profile with gaseous absorption (:500-790) ----
      for (int i = nt_ng + 1; i < PC_LEVELS; ++i) zng[i] = 0.0;    // the reference reads past NT_NG into zeroed storage
      ttot = tr + ta + tabs[PC_NLEV - 1];
      double zlim, tg_zlim, t_first, t_layer;
      const bool strong = tabs[PC_NLEV - 1] > PC_THRESHOLD_TAUABS;
      if (strong) {
        int i = 1;
        while (tabs[i - 1] < PC_THRESHOLD_TAUABS) ++i;
        const double alin = (tabs[i - 1] - tabs[i - 2]) / (altabs[i - 1] - altabs[i - 2]);
        const double blin = tabs[i - 1] - alin * altabs[i - 1];
        tg_zlim = PC_THRESHOLD_TAUABS;
        zlim = (tg_zlim - blin) / alin;
        t_first = PC_FIRST_LAYER;
        const double ttot_zlim0 = ta * exp(-zlim / ha) + tr * exp(-zlim / hr) + tg_zlim;
        if (PC_OS_NT - nt_ng - 2 <= 0) return 9600;
        t_layer = (ttot_zlim0 - t_first) / (PC_OS_NT - nt_ng - 2);
        t_layer = fmax(t_layer, PC_TCOUCHE);
      } else {
        zlim = (double)0.f;
        tg_zlim = tabs[PC_NLEV - 1];
        if ((ttot / PC_OS_NT_MIN) <= PC_FIRST_LAYER) {
          nt = PC_OS_NT_MIN;
          t_layer = ttot / nt;
          t_first = t_layer;
        } else if ((ttot / PC_OS_NT_MIN) < PC_TCOUCHE) {
          t_first = PC_FIRST_LAYER;
          t_layer = (ttot - t_first) / PC_OS_NT_MIN;
        } else {
          t_first = PC_FIRST_LAYER;
          nt = (int)((ttot - t_first) / PC_TCOUCHE);
          t_layer = (ttot - t_first) / nt;
        }
      }
      nt = 1;
      double z = PC_TOA_ALT;
      int ing = 1;
      double zing = zng[1];
      double hmol_prev = 0.0, haer_prev = 0.0, habs_prev = 0.0;    // Hmol, Haer, Habs of level NT-1
      double hmol_pp = 0.0, haer_pp = 0.0, habs_pp = 0.0;          // ... of level NT-2 (needed if the last level is dropped)
      h[0] = 0.0;
      const double ttot_zlim = ta * exp(-zlim / ha) + tr * exp(-zlim / hr) + tg_zlim;
      const PcColumn cg = {ta, ha, tr, hr, tabs, altabs, tg_zlim};
      while ((ttot_zlim - h[nt - 1]) > t_layer) {
        const int i = nt;
        if (i >= PC_OS_NT) return 9600;
        if (i == 1) {
          double dtau = 0.0;
          srch.first(true, cg, t_first, &z, &dtau);
          zprof[1] = z;
          h[1] = dtau;
          ing = 1;
        } else {
          z = srch.disc(t_layer, cg, h[i - 1], zprof[1], zlim);
        }
        if (z <= zing) {
          z = zing;
          ing = ing + 1;
          zing = (ing < PC_LEVELS) ? zng[ing] : 0.0;
        } else if ((z - zing) <= PC_THRESHOLD_DZ) {
          ing = ing + 1;
          zing = (ing < PC_LEVELS) ? zng[ing] : 0.0;
        }
        zprof[i] = z;
        const int j = pc_abs_bracket(altabs, z);
        double vg;
        if (z > altabs[0]) vg = tabs[j - 2];
        else {
          const double zz = (z - altabs[j - 2]) / (altabs[j - 1] - altabs[j - 2]);
          vg = (1 - zz) * tabs[j - 2] + zz * tabs[j - 1];
        }
        double vr = tr * exp(-z / hr), va = ta * exp(-z / ha);
        const double hm = vr, hae = va, hab = vg;
        h[i] = va + vr + vg;
        va = va - haer_prev; vr = vr - hmol_prev; vg = vg - habs_prev;
        pcaer[i] = va / (va + vr + vg);
        pcmol[i] = vr / (va + vr + vg);
        hmol_pp = hmol_prev; haer_pp = haer_prev; habs_pp = habs_prev;
        hmol_prev = hm; haer_prev = hae; habs_prev = hab;
        nt = nt + 1;
      }
      if ((zprof[nt - 1] - zlim) <= PC_THRESHOLD_DZ) {            // last computed level too close to the limit level: dropped
        nt = nt - 1;
        hmol_prev = hmol_pp; haer_prev = haer_pp; habs_prev = habs_pp;
      }
      if (nt < 1 || nt > PC_OS_NT) return 9600;
      zprof[nt] = zlim;
      {
        double vr = tr * exp(-zlim / hr), va = ta * exp(-zlim / ha), vg = tg_zlim;
        const double hm = vr, hae = va, hab = vg;
        h[nt] = vr + va + tg_zlim;
        va = va - haer_prev; vr = vr - hmol_prev; vg = vg - habs_prev;
        pcaer[nt] = va / (va + vr + vg);
        pcmol[nt] = vr / (va + vr + vg);
        hmol_prev = hm; haer_prev = hae; habs_prev = hab;
      }
      zprof[0] = PC_TOA_ALT;
      pcaer[0] = pcaer[1]; pcmol[0] = pcmol[1];
      h[0] = 0.0;
      if (strong) {                                               // one more layer from the limit level to the ground
        nt = nt + 1;
        if (nt > PC_OS_NT) return 9600;
        const double hm = tr, hae = ta, hab = tabs[PC_NLEV - 1];
        h[nt] = hm + hae + hab;
        const double vr = hm - hmol_prev, va = hae - haer_prev, vg = hab - habs_prev;
        pcaer[nt] = va / (va + vr + vg);
        pcmol[nt] = vr / (va + vr + vg);
        zprof[nt] = 0.0;
      }
    }
  } else {
    // ---- IPROFIL = 2: aerosols between two altitudes, no gas (:800-905) ----
    const double zmin = zmin_in, zmax = zmax_in;
    if (zmin < 0.0 || zmax <= zmin) return 1010;
    const double ttot = tr + ta;
    nt = (int)(ttot / PC_TCOUCHE);
    if (nt > PC_OS_NT) nt = PC_OS_NT;
    const double vr_c1 = tr * exp(-(zmax + PC_DZTRANSI) / hr);
    const double vr_c2 = tr * (exp(-zmin / hr) - exp(-(zmax + PC_DZTRANSI) / hr));
    double vr_c3;
    int nb_tr;
    if (zmin == 0.0) { vr_c3 = 0.0; nb_tr = 1; }
    else { vr_c3 = tr * (1.0 - exp(-(zmin - PC_DZTRANSI) / hr)); nb_tr = 2; }
    int nbsc1 = (int)((nt - nb_tr) * vr_c1 / (tr + ta));
    if (nbsc1 < PC_PROFIL_MIN_NBC) nbsc1 = PC_PROFIL_MIN_NBC;
    int nbsc3;
    if (zmin == 0.0) nbsc3 = 0;
    else { nbsc3 = (int)((nt - nb_tr) * vr_c3 / (tr + ta)); if (nbsc3 < PC_PROFIL_MIN_NBC) nbsc3 = PC_PROFIL_MIN_NBC; }
    const int nbsc2 = (nt - nb_tr) - nbsc1 - nbsc3;
    if (nbsc2 <= 0 || nt < 1) return 1020;                         // the reference divides by NBSC_C2 here
    if ((ta / nbsc2) < (double)0.00001f) return 1020;
    // Hmol in zprof[], Haer in zng[] until the final pass
    double *hmol = zprof, *haer = zng;
    hmol[0] = 0.0; haer[0] = 0.0;                                   // zeroed storage of the reference on entry
    double vr_sc = vr_c1 / nbsc1;
    for (int i = 1; i <= nbsc1; ++i) { hmol[i] = hmol[i - 1] + vr_sc; haer[i] = 0.0; }
    int i = nbsc1 + 1;
    hmol[i] = tr * exp(-zmax / hr);
    vr_sc = hmol[i] - hmol[i - 1];
    haer[i] = haer[i - 1] + (ta * vr_sc / vr_c2);
    const double delta_z = (zmax - zmin) / nbsc2;
    double z = zmax;
    for (i = nbsc1 + 2; i <= nbsc1 + nbsc2 + 1; ++i) {
      z = z - delta_z;
      hmol[i] = tr * exp(-z / hr);
      vr_sc = hmol[i] - hmol[i - 1];
      haer[i] = haer[i - 1] + (ta * vr_sc / vr_c2);
    }
    if (zmin != 0.0) {
      i = nbsc1 + nbsc2 + 2;
      hmol[i] = tr * exp(-(zmin - PC_DZTRANSI) / hr);
      haer[i] = haer[i - 1];
      vr_sc = vr_c3 / nbsc3;
      for (i = nbsc1 + nbsc2 + 3; i <= nt; ++i) { hmol[i] = vr_sc + hmol[i - 1]; haer[i] = haer[i - 1]; }
    }
    hmol[0] = tr * exp(-PC_TOA_ALT / hr);
    haer[0] = 0.0;
    h[0] = hmol[0] + haer[0];
    pcaer[0] = 0.0; pcmol[0] = 1.0;
    for (i = 1; i <= nt; ++i) {
      h[i] = hmol[i] + haer[i];
      if (haer[i] == haer[i - 1]) { pcaer[i] = 0.0; pcmol[i] = 1.0; }
      else { pcmol[i] = 1 / (1 + (ta / vr_c2)); pcaer[i] = 1 - pcmol[i]; }
    }
    for (i = nt; i >= 1; --i) zprof[i] = hr * log(tr / hmol[i]);    // in place: hmol[i] is read before zprof[i] is written
    zprof[0] = PC_TOA_ALT;
  }
  *nt_out = nt;
  return 0;
}
