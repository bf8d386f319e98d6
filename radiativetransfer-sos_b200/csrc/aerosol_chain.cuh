// aerosol_chain.cuh -- aerosol optics per wavelength, as the pieces the kernels of sosgpu_aerosols.cu give to
// single threads:
//   Mie theory for one size parameter               SOS_MIE, SOS_FPHASE_MIE       (SOS_MIE.F:205-690, 801-944)
//   integration over a size distribution            SOS_GRANU                     (SOS_AEROSOLS.F:4392-4767)
//   mixture of modes                                SOS_AEROSOLS                  (SOS_AEROSOLS.F:1455-1490, 2085-2110)
//   Legendre expansion alpha, beta, gamma, zeta     SOS_DECOMPO_LEGENDRE          (SOS_AEROSOLS.F:3924-4210)
//   with the optional truncation of the forward peak
// The serial recurrences of the reference (upward C_n / G_n, downward D_n / S_n, the sums over orders and over size
// parameters) are kept as serial chains in the reference's statement order -- each is one function here, run by one thread --
// and the kernels get their parallelism from running the independent chains of one size parameter on different warps, one
// scattering angle or one expansion order per thread, and one CTA per size parameter / component / model.  REAL*4 literals,
// REAL*4 sub-expressions (CO1, CO2 of the alpha / zeta sums) and the REAL*4 storage of the Mie records are the reference's.
// __host__ __device__ so that tests/aerosol_host.cpp can step the same functions against the reference library on a machine
// without a GPU; the library only ever runs them inside its kernels.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define AC_HD __host__ __device__ inline
#define AC_HDM __host__ __device__
#else
#define AC_HD static inline
#define AC_HDM
#endif

#define AC_MIE_DIM 10000            // CTE_MIE_DIM (SOS.h:96)
#define AC_MIE_NBMU_MAX 100         // CTE_MIE_NBMU_MAX (SOS.h:457)
#define AC_NB_MAX 200               // CTE_OS_NB_MAX (SOS.h:480)
#define AC_PI 3.141592653589793     // INCTE_PI = DACOS(-1.D+00): glibc returns this double
#define AC_MU1_TRONCA ((double)0.8f)    // CTE_AER_MU1_TRONCA (SOS.h:166), REAL*4 literal assigned to a double
#define AC_MU2_TRONCA ((double)0.94f)   // CTE_AER_MU2_TRONCA (SOS.h:167)
#define AC_SEUIL_TRONCA ((double)0.1f)  // CTE_PH_SEUIL_TRONCA (SOS.h:172)

// Work arrays of one size parameter: seven arrays of `stride` doubles, Fortran index i at [i + 1] (CNA, SNA, RGNA, IGNA start at
// -1).  The a_n / b_n of order i overwrite the G_n / D_n(m alpha) entries of the same order (each thread reads its seven inputs
// before it writes), the terms of the three sums over n go where C_n, S_n and D_n(alpha) were.
struct AcMieWork {
  double *cna, *sna, *rgna, *igna, *rdna, *rdnb, *idnb, *ra, *ia, *rb, *ib, *tq1, *tq2, *tg;
  size_t es;               // distance between consecutive elements of an array: 1, or 32 when the arrays of the 32 size parameters
                           // of a warp are interleaved (element i of lane l at [i * 32 + l]: coalesced)
};
#define AC_WORK_ARRAYS 7
AC_HD AcMieWork ac_work(double *base, size_t stride, size_t es = 1)
{
  AcMieWork w;
  const size_t a = stride * es;
  w.es = es;
  w.cna = base; w.sna = base + a; w.rgna = base + 2 * a; w.igna = base + 3 * a; w.rdna = base + 4 * a;
  w.rdnb = base + 5 * a; w.idnb = base + 6 * a;
  w.ra = w.rgna + es; w.ia = w.igna + es; w.rb = w.rdnb + es; w.ib = w.idnb + es;      // RA(i) in the slot of RGNA(i) ...
  w.tq1 = w.cna; w.tq2 = w.sna; w.tg = w.rdna;
  return w;
}

// Step of the size-parameter grid (SOS_MIE.F:406-411): REAL*4 literals assigned to / compared with doubles.
AC_HD double ac_mie_step(double alpha)
{
  double pas = (double)0.0001f;
  if (alpha > (double)0.1f) pas = (double)0.001f;
  if (alpha > (double)1.00f) pas = (double)0.01f;
  if (alpha > (double)10.f) pas = (double)0.05f;
  if (alpha > (double)30.f) pas = (double)0.10f;
  if (alpha > (double)100.f) pas = (double)1.00f;
  return pas;
}

// ---- SOS_MIE.F:424-470: upward recurrences.  C_n alone decides where the series are cut (N2, N1). ----
AC_HD void ac_mie_orders(double alpha, int *n1, int *n2)
{
  *n1 = (int)trunc(alpha + alpha + 20);
  *n2 = (int)trunc(alpha + alpha + 5);
}
AC_HD void ac_mie_chain_c(double alpha, const AcMieWork &w, int *n1, int *n2)
{
  double cm2 = -sin(alpha), cm1 = cos(alpha);
  w.cna[(size_t)(0) * w.es] = cm2; w.cna[(size_t)(1) * w.es] = cm1;
  const int n2_0 = *n2;
  for (int i = 1; i <= n2_0; ++i) {
    const double c = (2 * i - 1.0) * cm1 / alpha - cm2;
    w.cna[(size_t)(i + 1) * w.es] = c;
    cm2 = cm1; cm1 = c;
    if (c < 1.e+304) continue;
    *n2 = i; *n1 = i + 15;                              // C_n diverges: cut here (SOS_MIE.F:463-467)
    return;
  }
}
AC_HD void ac_mie_chain_g(double alpha, const AcMieWork &w, int n2)
{
  double x = 0.0, y = -1.0;
  w.rgna[(size_t)(0) * w.es] = 0.0; w.rgna[(size_t)(1) * w.es] = 0.0; w.igna[(size_t)(0) * w.es] = 0.0; w.igna[(size_t)(1) * w.es] = -1.0;
  for (int i = 1; i <= n2; ++i) {
    const double z = i / alpha;
    const double ww = ((z - x) * (z - x) + (y * y));
    const double xr = (z - x) / ww - z;
    const double yi = y / ww;
    w.rgna[(size_t)(i + 1) * w.es] = xr; w.igna[(size_t)(i + 1) * w.es] = yi;
    x = xr; y = yi;
  }
}

// ---- SOS_MIE.F:479-523: downward recurrences from N1 ----
AC_HD void ac_mie_chain_db(double alpha, double rn, double in, const AcMieWork &w, int n1)
{
  const double rbeta = rn * alpha, ibeta = in * alpha;
  const double x1 = rbeta * rbeta + ibeta * ibeta;
  const double x2 = rbeta / x1, x3 = ibeta / x1;
  double x = 0.0, y = 0.0;
  w.rdnb[(size_t)(n1 + 1) * w.es] = 0.0; w.idnb[(size_t)(n1 + 1) * w.es] = 0.0;
  for (int i = n1 - 1; i >= 0; --i) {
    const double z = x + (i + 1.0) * x2;
    const double ww = y - (i + 1.0) * x3;
    const double x4 = z * z + ww * ww;
    x = (i + 1.0) * x2 - z / x4;
    y = -((i + 1.0) * x3) + ww / x4;
    w.rdnb[(size_t)(i + 1) * w.es] = x; w.idnb[(size_t)(i + 1) * w.es] = y;
  }
}
AC_HD void ac_mie_chain_da(double alpha, const AcMieWork &w, int n1)
{
  double x = 0.0;
  w.rdna[(size_t)(n1 + 1) * w.es] = 0.0;
  for (int i = n1 - 1; i >= 0; --i) {
    const double z = (i + 1.0) / alpha;
    x = z - 1.0 / (x + z);
    w.rdna[(size_t)(i + 1) * w.es] = x;
  }
}
AC_HD void ac_mie_chain_s(double alpha, const AcMieWork &w, int n1, int n2)
{
  double sp1 = 0.0, s0 = 1.0;                           // SNA(I+1), SNA(I)
  w.sna[(size_t)(n1 + 1) * w.es] = 0.0; w.sna[(size_t)(n1) * w.es] = 1.0;
  for (int i = n1 - 1; i >= 0; --i) {
    const double sm1 = (2.0 * i + 1.0) * s0 / alpha - sp1;
    w.sna[(size_t)(i) * w.es] = sm1;                                     // SNA(I-1)
    if (sm1 > 1.e+304) {                                // renormalise what has been computed (SOS_MIE.F:509-515)
      const int test = i - 1;
      const double xx = sm1;
      for (int j = test; j <= n2; ++j) w.sna[(size_t)(j + 1) * w.es] = w.sna[(size_t)(j + 1) * w.es] / xx;
      sp1 = w.sna[(size_t)(i + 1) * w.es]; s0 = w.sna[(size_t)(i) * w.es];
    } else {
      sp1 = s0; s0 = sm1;
    }
  }
}

// SOS_MIE.F:527-531: S_n normalised so that S_0 = sin(alpha); q = SNA(0) / DSIN(ALPHA)
AC_HD double ac_mie_snorm(double alpha, const AcMieWork &w) { return w.sna[(size_t)(1) * w.es] / sin(alpha); }

// ---- SOS_MIE.F:541-590: a_n, b_n of order i (1 <= i <= n2) ----
AC_HD void ac_mie_ab(int i, double rn, double in, const AcMieWork &w)
{
  const double x1 = w.sna[(size_t)(i + 1) * w.es], x2 = w.cna[(size_t)(i + 1) * w.es], x3 = w.rdnb[(size_t)(i + 1) * w.es], x4 = w.idnb[(size_t)(i + 1) * w.es], x5 = w.rdna[(size_t)(i + 1) * w.es];
  const double x6 = w.rgna[(size_t)(i + 1) * w.es], x7 = w.igna[(size_t)(i + 1) * w.es];
  double y1 = x3 - rn * x5;
  double y2 = x4 - in * x5;
  double y3 = x3 - rn * x6 + in * x7;
  double y4 = x4 - rn * x7 - in * x6;
  const double y5 = rn * x3 - in * x4 - x5;
  const double y6 = in * x3 + rn * x4;
  const double y7 = rn * x3 - in * x4 - x6;
  const double y8 = in * x3 + rn * x4 - x7;
  const double z4 = y2 * y3 - y1 * y4;
  const double z3 = y1 * y3 + y2 * y4;
  const double z5 = x1 * x1 + x2 * x2;
  const double z6 = y3 * y3 + y4 * y4;
  const double z7 = y5 * y7 + y6 * y8;
  const double z8 = y6 * y7 - y5 * y8;
  const double z9 = y7 * y7 + y8 * y8;
  const int un = (i & 1) ? 1 : -1;
  double q = (i + i + 1.0) / i / (i + 1.0) * un;
  if (x2 > 1.e+300) {
    y1 = 0.0; y2 = 0.0; y3 = 0.0; y4 = 0.0;
  } else {
    y1 = x1 * (x1 * z3 + x2 * z4) / z5 / z6;
    y2 = x1 * (x1 * z4 - x2 * z3) / z5 / z6;
    y3 = x1 * (x1 * z7 + x2 * z8) / z5 / z9;
    y4 = x1 * (x1 * z8 - x2 * z7) / z5 / z9;
  }
  w.ra[(size_t)(i) * w.es] = y2 * q;
  w.ib[(size_t)(i) * w.es] = y3 * q;
  q = -q;
  w.rb[(size_t)(i) * w.es] = y4 * q;
  w.ia[(size_t)(i) * w.es] = y1 * q;
}

// ---- SOS_MIE.F:606-630: the terms of the three sums over n (each independent of the running sums) ----
AC_HD void ac_mie_qterms(int n, const AcMieWork &w)
{
  const double x = w.ra[(size_t)(n) * w.es], y = w.ia[(size_t)(n) * w.es], z = w.rb[(size_t)(n) * w.es], t = w.ib[(size_t)(n) * w.es];
  const double xx = w.ra[(size_t)(n + 1) * w.es], yy = w.ia[(size_t)(n + 1) * w.es], zz = w.rb[(size_t)(n + 1) * w.es], tt = w.ib[(size_t)(n + 1) * w.es];
  const double a2 = (n + 1.0);
  const int j = (n & 1) ? -1 : 1;
  w.tq1[(size_t)(n) * w.es] = n * a2 * j * (y - t);
  w.tq2[(size_t)(n) * w.es] = (n * n) * a2 * a2 / (n + a2) * (x * x + y * y + z * z + t * t);
  w.tg[(size_t)(n) * w.es] = a2 * n / (a2 + n) * (n * (a2 + (double)1.f) * (a2 + (double)1.f) / (double)(2.f * n + 3.f) *
                                 (y * yy + x * xx + t * tt + z * zz) + y * t + x * z);
}
// which = 0: QEXT, 1: QSCA, 2: G (before the final scalings)
AC_HD double ac_mie_qsum(int which, int n2, const AcMieWork &w)
{
  const double *t = which == 0 ? w.tq1 : which == 1 ? w.tq2 : w.tg;
  double s = 0.0;
  if (which == 2) for (int n = 1; n <= n2; ++n) s = s - t[(size_t)n * w.es];
  else for (int n = 1; n <= n2; ++n) s = s + t[(size_t)n * w.es];
  return s;
}
AC_HD void ac_mie_qfinal(double alpha, double *qext, double *qsca, double *g)
{
  const double w6 = 2.0 / alpha / alpha;
  *qext = w6 * *qext;
  *qsca = w6 * *qsca;
  *g = 4.0 * *g / *qsca / alpha / alpha;
}

// ---- SOS_FPHASE_MIE (SOS_MIE.F:880-916): one scattering angle, coefficients read through `coef(n, &ar, &ai, &br, &bi)` ----
template <class Coef> AC_HD void ac_mie_phase(double rmu_j, double alpha, double kma2, int n2, const Coef &coef, float *imie, float *qmie,
                                              float *umie)
{
  const double cf = 2.0 / kma2 / (alpha * alpha);
  const double x = -rmu_j;
  double pim = 0.0, piv = 1.0, tau = x, res1 = 0.0, res2 = 0.0, ims1 = 0.0, ims2 = 0.0;
  for (int n = 1; n <= n2; ++n) {
    double ar, ai, br, bi;
    coef(n, &ar, &ai, &br, &bi);
    res1 = res1 - ai * piv - bi * tau;
    res2 = res2 + ai * tau + bi * piv;
    ims1 = ims1 + ar * piv + br * tau;
    ims2 = ims2 - ar * tau - br * piv;
    const double pip = ((2.0 * n + 1.0) * x * piv - (n + 1.0) * pim) / n;
    pim = piv;
    piv = pip;
    tau = (n + 1.0) * x * piv - (n + 2.0) * pim;
  }
  const double y1 = res1 * res1 + ims1 * ims1;
  const double y2 = res2 * res2 + ims2 * ims2;
  const double y3 = 2.0 * res2 * res1;
  const double y4 = 2.0 * ims2 * ims1;
  *imie = (float)(cf * (y1 + y2));
  *qmie = (float)(cf * (y2 - y1));
  *umie = (float)(cf * (y3 + y4));
}

// ================= SOS_GRANU =================
// step PAS (REAL*4) in force after the record of size parameter alpha has been read (SOS_AEROSOLS.F:4549-4553); the value
// before the first record is 0.0001
AC_HD float ac_granu_step_after(float alpha, float pas_before)
{
  float pas = pas_before;
  if (alpha > 0.10f) pas = 0.001f;
  if (alpha > 1.00f) pas = 0.01f;
  if (alpha > 10.f) pas = 0.05f;
  if (alpha > 30.f) pas = 0.10f;
  if (alpha > 100.f) pas = 1.00f;
  return pas;
}
// One record of the Mie table: returns 1 when the reference leaves its loop at this record (before using it), else 0 and
// x1e = X1*QEXT, x1s = QSCA*X1 (also the weight of the phase functions), nrpr = NR*PR  (SOS_AEROSOLS.F:4537-4600)
AC_HD int ac_granu_record(float alpha, float alpha_prev, int first, float qext, float qsca, double alphaf, int igranu, double v1,
                          double v2, double v3, double wa, double *x1e, double *x1s, double *nrpr)
{
  const float pas_before = first ? 0.0001f : ac_granu_step_after(alpha_prev, 0.0001f);
  const double r = alpha * wa / 2.0 / AC_PI;
  if ((double)alpha >= (alphaf - (double)pas_before)) return 1;
  const float pas = ac_granu_step_after(alpha, pas_before);
  double nr = 0.0;
  if (igranu == 1) {
    const double b = log(r / v1) / v2;
    nr = exp(-(b * b) / (double)2.f) / (r * v2 * sqrt(2 * AC_PI));
  }
  if (igranu == 2) {
    const double nr0 = pow(v1, -v2);
    if (r > v3) return 1;
    nr = (r <= v1) ? nr0 : pow(r, -v2);
  }
  const double pr = wa * (double)pas / (double)2.f / AC_PI;
  const double x1 = nr * pr * AC_PI * (r * r);
  *x1e = x1 * (double)qext;
  *x1s = (double)qsca * x1;
  *nrpr = nr * pr;
  return 0;
}

// ================= SOS_DECOMPO_LEGENDRE =================
// Legendre polynomials P_k(mu) and the functions P^2_k(mu) of one angle, k = 0..nb, written with stride `st`
// (SOS_AEROSOLS.F:4100-4104, 4166-4178, 4185-4189)
AC_HD void ac_legendre_column(double xrmu, int nb, double *pl, double *pol, size_t st)
{
  double pm = 0.0, p = 1.0;
  for (int k = 0; k <= nb; ++k) {
    pl[k * st] = p;
    const double pn = ((double)(2 * k + 1.f) * xrmu * p - k * pm) / (double)(k + 1.f);
    pm = p; p = pn;
  }
  double qm = 0.0, q = 3.0 * (1.0 - xrmu * xrmu) / 2.0 / sqrt(6.0);
  pol[0] = 0.0;
  if (nb >= 1) pol[st] = 0.0;
  for (int k = 2; k <= nb; ++k) {
    pol[k * st] = q;
    const double d = (double)(2.f * k + 1.f) / sqrt(1.0 * (double)(k + 3.f) * (double)(k - 1.f));
    const double e = sqrt(1.0 * (double)(k + 2.f) * (double)(k - 2.f)) / (double)(2.f * k + 1.f);
    const double qn = d * (xrmu * q - e * qm);
    qm = q; q = qn;
  }
}
// index (0-based position in V(-n:n)) of the Gauss angle just below the truncation bound (SOS_AEROSOLS.F:4036-4052); -1 if none
AC_HD int ac_tronca_index(const double *xmu, const double *xhr, int nbmu, double bound)
{
  for (int j = 1; j <= nbmu; ++j)
    if (xmu[nbmu + j] > bound && xhr[nbmu + j] != 0.0) return j - 1;
  return -1;
}
// truncated phase function at angle j > kk (SOS_AEROSOLS.F:4075-4084)
AC_HD double ac_tronca_value(double p11_k, double p11_kk, double mu_k, double mu_kk, double mu_j)
{
  const double aa = (log10(p11_kk) - log10(p11_k)) / (acos(mu_kk) - acos(mu_k));
  const double x1 = log10(p11_kk), x2 = acos(mu_kk);
  const double c = x1 + aa * (acos(mu_j) - x2);
  return pow(10.0, c);
}
// alpha(i), zeta(i) from beta22, delta33 (after their (2k+1)/2 scaling), i >= 2 (SOS_AEROSOLS.F:4200-4222).  CO1, CO2 and the
// integer-valued factors are REAL*4 expressions in the reference.
AC_HD void ac_alpha_zeta(int i, const double *beta22, const double *delta33, double *alp, double *zeta)
{
  const double co1 = (double)(4 * (2 * i + 1.f) / i / (i - 1.f) / (i + 1.f) / (i + 2.f));
  double co2 = (double)(i * (i - 1.f) / ((i + 1.f) * (i + 2.f)));
  const double co3 = co2 * delta33[i];
  co2 = co2 * beta22[i];
  const int nn = (int)(i * .5f), mm = (int)((i - 1) * .5f);
  double som1 = 0.0, som2 = 0.0, som3 = 0.0, som4 = 0.0;
  for (int j = 1; j <= nn; ++j) {
    const double x2 = (double)((i - 1.f) * (i - 1.f) - 3.f * (2 * j - 1.f) * (i - j));
    som1 = som1 + x2 * beta22[i - 2 * j];
    som2 = som2 + x2 * delta33[i - 2 * j];
  }
  for (int j = 0; j <= mm; ++j) {
    const double x2 = (double)((i - 1.f) * (i - 1.f) - 3.f * j * (2 * i - 2 * j - 1.f));
    som3 = som3 + x2 * beta22[i - 2 * j - 1];
    som4 = som4 + x2 * delta33[i - 2 * j - 1];
  }
  *zeta = co3 - co1 * (som2 - som3);
  *alp = co2 - co1 * (som1 - som4);
}
// single-scattering albedo after truncation and the asymmetry factor of the result file (SOS_AEROSOLS.F:2763-2768, 2822-2823)
AC_HD double ac_piztr(double piz, double ct) { return piz * ((double)1.f - ct / (double)2.f) / ((double)1.f - piz * ct / (double)2.f); }
AC_HD double ac_asym(double ct, double beta11_1) { return ct / (double)2.f + ((double)1.f - ct / (double)2.f) * beta11_1 / (double)3.f; }

// =====================================================================================================================
// Compositions.  Every function below is entered by all `nthr` threads of one CTA (thread `tid`), `sync()` being the CTA
// barrier; loops over angles / orders / records are strided by nthr, the serial chains go to the first lanes of different
// warps (ac_role).  With tid = 0, nthr = 1 and an empty sync() the same code is the serial routine the CPU tests step.
// =====================================================================================================================
struct AcNoSync { AC_HDM void operator()() const {} };
AC_HD int ac_role(int k, int nthr) { return (k * 32) % nthr; }
AC_HD void ac_min_int(int *p, int v)
{
#ifdef __CUDA_ARCH__
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}

struct AcCoef {
  const double *ra, *ia, *rb, *ib;
  size_t es;
  AC_HDM void operator()(int n, double *ar, double *ai, double *br, double *bi) const
  {
    const size_t o = (size_t)n * es;
    *ar = ra[o]; *ai = ia[o]; *br = rb[o]; *bi = ib[o];
  }
};

// One record of the Mie table (SOS_MIE.F:399-646 for one ALPHA).  sh_n[2], sh_q[4]: CTA-shared.  Outputs: rec[0..2] = ALPHA, QEXT,
// QSCA as REAL*4, *g, and the three phase functions at the 2*nbmu+1 angles (REAL*4).
template <class Sync> AC_HD void ac_mie_record(int tid, int nthr, Sync sync, double alpha, double rn, double in, int nbmu, const double *rmu,
                                               const AcMieWork &w, int *sh_n, double *sh_q, float *rec, double *g, float *imie,
                                               float *qmie, float *umie)
{
  int n1, n2;
  ac_mie_orders(alpha, &n1, &n2);
  const int n2_0 = n2;
  if (tid == ac_role(0, nthr)) { ac_mie_chain_c(alpha, w, &n1, &n2); sh_n[0] = n1; sh_n[1] = n2; }
  if (tid == ac_role(1, nthr)) ac_mie_chain_g(alpha, w, n2_0);
  sync();
  n1 = sh_n[0]; n2 = sh_n[1];
  if (tid == ac_role(0, nthr)) ac_mie_chain_db(alpha, rn, in, w, n1);
  if (tid == ac_role(1, nthr)) ac_mie_chain_da(alpha, w, n1);
  if (tid == ac_role(2, nthr)) { ac_mie_chain_s(alpha, w, n1, n2); sh_q[3] = ac_mie_snorm(alpha, w); }
  sync();
  const double q = sh_q[3];
  for (int i = tid; i <= n2; i += nthr) w.sna[(size_t)(i + 1) * w.es] = w.sna[(size_t)(i + 1) * w.es] / q;
  sync();
  for (int i = 1 + tid; i <= n2; i += nthr) ac_mie_ab(i, rn, in, w);
  if (tid == 0) {
    w.ra[(size_t)(0) * w.es] = 0.0; w.ia[(size_t)(0) * w.es] = 0.0; w.rb[(size_t)(0) * w.es] = 0.0; w.ib[(size_t)(0) * w.es] = 0.0;
    w.ra[(size_t)(n2 + 1) * w.es] = 0.0; w.ia[(size_t)(n2 + 1) * w.es] = 0.0; w.rb[(size_t)(n2 + 1) * w.es] = 0.0; w.ib[(size_t)(n2 + 1) * w.es] = 0.0;
  }
  sync();
  for (int n = 1 + tid; n <= n2; n += nthr) ac_mie_qterms(n, w);
  sync();
  for (int k = 0; k < 3; ++k)
    if (tid == ac_role(k, nthr)) sh_q[k] = ac_mie_qsum(k, n2, w);
  sync();
  if (tid == 0) {
    double qe = sh_q[0], qs = sh_q[1], gg = sh_q[2];
    ac_mie_qfinal(alpha, &qe, &qs, &gg);
    sh_q[0] = qe; sh_q[1] = qs; sh_q[2] = gg;
    rec[0] = (float)alpha; rec[1] = (float)qe; rec[2] = (float)qs; *g = gg;
  }
  sync();
  AcCoef cf{w.ra, w.ia, w.rb, w.ib, w.es};
  const double kma2 = sh_q[1];
  for (int j = tid; j <= 2 * nbmu; j += nthr) ac_mie_phase(rmu[j], alpha, kma2, n2, cf, &imie[j], &qmie[j], &umie[j]);
  sync();
}

// SOS_GRANU for one component over a Mie table of nrec records (phase functions [rec][nang], nang = 2*nbmu+1).
// scratch: 3*nrec doubles; sh_k: CTA-shared int.  out[0..2] = KMAT1, KMAT2, SOMME_NR.  *ier = -1 when the table ends before
// the reference's loop would (its read error 9992).
template <class Sync> AC_HD void ac_granu(int tid, int nthr, Sync sync, int nrec, const float *rec, const float *imie, const float *qmie,
                                          const float *umie, int nang, double alphaf, int igranu, double v1, double v2, double v3,
                                          double wa, double *scratch, int *sh_k, double *out, double *p11, double *p12, double *p33,
                                          int *ier)
{
  double *x1e = scratch, *x1s = scratch + nrec, *nrpr = scratch + 2 * (size_t)nrec;
  if (tid == 0) *sh_k = nrec;
  sync();
  for (int k = tid; k < nrec; k += nthr) {
    const float a = rec[3 * (size_t)k], ap = k ? rec[3 * (size_t)(k - 1)] : 0.f;
    if (ac_granu_record(a, ap, k == 0, rec[3 * (size_t)k + 1], rec[3 * (size_t)k + 2], alphaf, igranu, v1, v2, v3, wa, &x1e[k], &x1s[k], &nrpr[k]))
      ac_min_int(sh_k, k);
  }
  sync();
  const int ks = *sh_k;
  if (ks >= nrec) { if (tid == 0) *ier = -1; return; }
  for (int j = tid; j < nang + 3; j += nthr) {
    if (j < nang) {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int k = 0; k < ks; ++k) {
        const double x = x1s[k];
        a = a + (double)imie[(size_t)k * nang + j] * x;
        b = b + (double)qmie[(size_t)k * nang + j] * x;
        c = c + (double)umie[(size_t)k * nang + j] * x;
      }
      p11[j] = a; p12[j] = b; p33[j] = c;
    } else {
      const double *t = j == nang ? x1e : j == nang + 1 ? x1s : nrpr;
      double s = 0.0;
      for (int k = 0; k < ks; ++k) s = s + t[k];
      out[j - nang] = s;
    }
  }
  sync();
  const double k2 = out[1];
  for (int j = tid; j < nang; j += nthr) { p11[j] = p11[j] / k2; p12[j] = p12[j] / k2; p33[j] = p33[j] / k2; }
  sync();
  if (tid == 0) { const double s = out[2]; out[0] = out[0] / s; out[1] = out[1] / s; *ier = 0; }
}

// Mixture of components + SOS_DECOMPO_LEGENDRE for one aerosol model.
struct AcModel {
  int ncomp;             // 0: one component taken as it is (mono-modal, SOS_AEROSOLS.F:1279-1301); 1..4: mixture (:1455-1490, :2085-2110)
  int comp[4];           // component indices
  double w[4];           // number fractions N(I)/NTOT resp. normalised CVI(I); a zero entry is skipped as in the reference
  int itronc, os_nb;
};
struct AcModelShared {   // CTA-shared (or host-local) work space
  double p11[2 * AC_MIE_NBMU_MAX + 1], p12[2 * AC_MIE_NBMU_MAX + 1], p22[2 * AC_MIE_NBMU_MAX + 1], p33[2 * AC_MIE_NBMU_MAX + 1];
  double ttt[2 * AC_MIE_NBMU_MAX + 1], xa[2 * AC_MIE_NBMU_MAX + 1], xb[2 * AC_MIE_NBMU_MAX + 1], xc[2 * AC_MIE_NBMU_MAX + 1];
  double beta11[AC_NB_MAX + 1], beta22[AC_NB_MAX + 1], gamma12[AC_NB_MAX + 1], delta33[AC_NB_MAX + 1], alp[AC_NB_MAX + 1], zeta[AC_NB_MAX + 1];
  double kmat1, kmat2;
  int k, kk, itronc;
};
// scal[8]: KMAT1, KMAT2, PIZ, PIZTR, COEF_TRONCA, asymmetry factor (no truncation), Z1, ITRONC on exit.
// coef: [6][os_nb+1] ALP, BETA11, GAMMA12, ZETA, BETA22, DELTA33.  phase (may be null): [4][nang] P11 (truncated), P12, P33, TTT.
// comp_p22 may be null (spherical particles: P22 = P11 before truncation).  pl, pol: tables [k][nang] of ac_legendre_column.
template <class Sync> AC_HD void ac_model(int tid, int nthr, Sync sync, int nbmu, const double *xmu, const double *xhr, const double *pl,
                                          const double *pol, const double *comp_k, const double *comp_p11, const double *comp_p12,
                                          const double *comp_p33, const double *comp_p22, const AcModel &m, AcModelShared &s,
                                          double *scal, double *coef, double *phase, int *ier)
{
  const int nang = 2 * nbmu + 1, nb = m.os_nb;
  // ---- mixture ----
  double k1 = 0.0, k2 = 0.0;
  if (m.ncomp == 0) { k1 = comp_k[3 * (size_t)m.comp[0]]; k2 = comp_k[3 * (size_t)m.comp[0] + 1]; }
  else
    for (int i = 0; i < m.ncomp; ++i) {
      if (m.w[i] == 0.0) continue;
      k1 = k1 + m.w[i] * comp_k[3 * (size_t)m.comp[i]];
      k2 = k2 + m.w[i] * comp_k[3 * (size_t)m.comp[i] + 1];
    }
  for (int j = tid; j < nang; j += nthr) {
    double a = 0.0, b = 0.0, c = 0.0;
    if (m.ncomp == 0) {
      const size_t o = (size_t)m.comp[0] * nang + j;
      a = comp_p11[o]; b = comp_p12[o]; c = comp_p33[o];
    } else {
      for (int i = 0; i < m.ncomp; ++i) {
        if (m.w[i] == 0.0) continue;
        const size_t o = (size_t)m.comp[i] * nang + j;
        const double kk2 = comp_k[3 * (size_t)m.comp[i] + 1];
        a = a + m.w[i] * comp_p11[o] * kk2;
        b = b + m.w[i] * comp_p12[o] * kk2;
        c = c + m.w[i] * comp_p33[o] * kk2;
      }
      a = a / k2; b = b / k2; c = c / k2;
    }
    s.p11[j] = a; s.p12[j] = b; s.p33[j] = c;
    s.p22[j] = (comp_p22 && m.ncomp == 0) ? comp_p22[(size_t)m.comp[0] * nang + j] : a;
    s.ttt[j] = a;
  }
  if (tid == 0) {
    s.kmat1 = k1; s.kmat2 = k2; s.itronc = m.itronc;
    s.k = ac_tronca_index(xmu, xhr, nbmu, AC_MU1_TRONCA);
    s.kk = ac_tronca_index(xmu, xhr, nbmu, AC_MU2_TRONCA);
  }
  sync();
  // ---- truncation of the forward peak (SOS_AEROSOLS.F:4022-4086) ----
  if (s.itronc != 0) {
    if (s.k < 0 || s.kk < 0) { if (tid == 0) *ier = -1; return; }      // the reference would use undefined indices
    const double pk = s.p11[nbmu + s.k], pkk = s.p11[nbmu + s.kk], mk = xmu[nbmu + s.k], mkk = xmu[nbmu + s.kk];
    sync();
    for (int j = nbmu + s.kk + 1 + tid; j < nang; j += nthr) s.p11[j] = ac_tronca_value(pk, pkk, mk, mkk, xmu[j]);
    sync();
  }
  double ct = 0.0;
  for (;;) {
    // ---- beta11 (SOS_AEROSOLS.F:4092-4112) ----
    for (int j = tid; j < nang; j += nthr) s.xa[j] = s.p11[j] * xhr[j];
    sync();
    for (int k = tid; k <= nb; k += nthr) {
      double b = 0.0;
      for (int j = 0; j < nang; ++j) {
        if (j == nbmu) continue;
        b = b + s.xa[j] * pl[(size_t)k * nang + j];
      }
      s.beta11[k] = (2 * k + 1) * b * (double).5f;
    }
    sync();
    ct = (s.itronc == 1) ? 2 * (1 - s.beta11[0]) : 0.0;
    if (s.itronc == 1 && ct < AC_SEUIL_TRONCA) {                       // truncation too weak: cancelled (:4126-4152)
      sync();
      for (int j = tid; j < nang; j += nthr) s.p11[j] = s.ttt[j];
      if (tid == 0) s.itronc = 0;
      sync();
      continue;
    }
    break;
  }
  // ---- gamma12, beta22, delta33 (SOS_AEROSOLS.F:4160-4198) ----
  for (int j = tid; j < nang; j += nthr) {
    s.xa[j] = xhr[j] * s.p12[j] * s.p11[j] / s.ttt[j];
    s.xb[j] = xhr[j] * s.p22[j] * (s.p11[j] / s.ttt[j]);
    s.xc[j] = xhr[j] * s.p33[j] * s.p11[j] / s.ttt[j];
  }
  sync();
  for (int k = tid; k <= nb; k += nthr) {
    double g = 0.0, b = 0.0, d = 0.0;
    for (int j = 0; j < nang; ++j) {
      if (j == nbmu) continue;
      const double p = pl[(size_t)k * nang + j];
      if (k >= 2) g = g + s.xa[j] * pol[(size_t)k * nang + j];
      b = b + s.xb[j] * p;
      d = d + s.xc[j] * p;
    }
    const double f = (double)(2.f * k + 1.f);
    s.beta22[k] = b * f * (double).5f;
    s.delta33[k] = d * f * (double).5f;
    s.gamma12[k] = g * f * (double).5f;
    if (k < 2) { s.alp[k] = 0.0; s.zeta[k] = 0.0; }
  }
  sync();
  for (int i = 2 + tid; i <= nb; i += nthr) ac_alpha_zeta(i, s.beta22, s.delta33, &s.alp[i], &s.zeta[i]);
  sync();
  const double z1 = s.beta11[0];
  sync();
  for (int k = tid; k <= nb; k += nthr) {
    const size_t st = (size_t)nb + 1;
    coef[k] = s.alp[k] / z1;
    coef[st + k] = s.beta11[k] / z1;
    coef[2 * st + k] = s.gamma12[k] / z1;
    coef[3 * st + k] = s.zeta[k] / z1;
    coef[4 * st + k] = s.beta22[k] / z1;
    coef[5 * st + k] = s.delta33[k] / z1;
  }
  if (phase)
    for (int j = tid; j < nang; j += nthr) {
      phase[j] = s.p11[j]; phase[nang + j] = s.p12[j]; phase[2 * nang + j] = s.p33[j]; phase[3 * nang + j] = s.ttt[j];
    }
  if (tid == 0) {
    const double piz = s.kmat2 / s.kmat1;
    scal[0] = s.kmat1; scal[1] = s.kmat2; scal[2] = piz; scal[3] = ac_piztr(piz, ct); scal[4] = ct;
    scal[5] = ac_asym(ct, s.beta11[1] / z1); scal[6] = z1; scal[7] = (double)s.itronc;
    *ier = 0;
  }
}

// =====================================================================================================================
// Mie tables with one size parameter per LANE (what the kernels run; ac_mie_record above is the same arithmetic with one size
// parameter per CTA and serves the CPU tests as the plain serial routine).  ncu on the CTA-per-size-parameter kernel showed 56 %
// of its FP64 instructions issued with one active lane (the five serial recurrences), each costing the FP64 pipe as much as a
// full warp.  Here the 32 lanes of a warp run the same recurrence for 32 size parameters of similar magnitude (the work list is
// sorted), their work arrays interleaved element by element (coalesced), and the phase functions are a second stage with one
// (angle, 32 size parameters) task per warp.
//   stage 1 (ac_coef_phase, phases 0..6 separated by CTA barriers; warp roles wr = 0..nw-1, nw >= 3):
//     0: C_n (decides N1, N2) | G_n         1: D_n(m alpha) | D_n(alpha) | S_n + its normalisation factor
//     2: S_n normalised                     3: a_n, b_n (orders strided over the warps)
//     4: terms of the sums over n           5: QEXT | QSCA | G sums               6: final scalings, record header
//   stage 2 (ac_mie_phase): Imie, Qmie, Umie of one angle, coefficients read from the stage-1 arrays
// =====================================================================================================================
struct AcLaneState { int n1, n2; double snorm, sum[3]; };
#define AC_COEF_PHASES 7

AC_HD void ac_coef_phase(int p, int wr, int nw, double alpha, double rn, double in, const AcMieWork &w, AcLaneState &st, float *rec, double *g,
                         int *out_n2, double *out_qsca)
{
  switch (p) {
  case 0: {
    int n1, n2;
    ac_mie_orders(alpha, &n1, &n2);
    if (wr == 0) { ac_mie_chain_c(alpha, w, &n1, &n2); st.n1 = n1; st.n2 = n2; }
    else if (wr == 1) ac_mie_chain_g(alpha, w, n2);
    break;
  }
  case 1:
    if (wr == 0) ac_mie_chain_db(alpha, rn, in, w, st.n1);
    else if (wr == 1) ac_mie_chain_da(alpha, w, st.n1);
    else if (wr == 2) { ac_mie_chain_s(alpha, w, st.n1, st.n2); st.snorm = ac_mie_snorm(alpha, w); }
    break;
  case 2:
    for (int i = wr; i <= st.n2; i += nw) w.sna[(size_t)(i + 1) * w.es] = w.sna[(size_t)(i + 1) * w.es] / st.snorm;
    break;
  case 3:
    for (int i = 1 + wr; i <= st.n2; i += nw) ac_mie_ab(i, rn, in, w);
    if (wr == 0) {
      const size_t e = (size_t)(st.n2 + 1) * w.es;
      w.ra[0] = 0.0; w.ia[0] = 0.0; w.rb[0] = 0.0; w.ib[0] = 0.0;
      w.ra[e] = 0.0; w.ia[e] = 0.0; w.rb[e] = 0.0; w.ib[e] = 0.0;
    }
    break;
  case 4:
    for (int n = 1 + wr; n <= st.n2; n += nw) ac_mie_qterms(n, w);
    break;
  case 5:
    if (wr < 3) st.sum[wr] = ac_mie_qsum(wr, st.n2, w);
    break;
  default:
    if (wr == 0) {
      double qe = st.sum[0], qs = st.sum[1], gg = st.sum[2];
      ac_mie_qfinal(alpha, &qe, &qs, &gg);
      rec[0] = (float)alpha; rec[1] = (float)qe; rec[2] = (float)qs; *g = gg;
      *out_n2 = st.n2; *out_qsca = qs;
    }
    break;
  }
}

// ---- host only ----
#include <algorithm>
#include <vector>
// Work list of the two stages: the records of all tables sorted by decreasing size parameter (work grows with it), cut into
// groups of 32 (one warp), padded with table = -1; per group the length of its work arrays (the largest N1 + 4 of its lanes) and
// its offset in the arena; groups packed into chunks that fit an arena of `budget` doubles (a chunk is one launch of each stage).
struct AcMiePlan {
  std::vector<int> item_table, item_rec;         // [ngroup * 32]
  std::vector<long long> group_off;              // doubles, relative to the arena start of the group's chunk
  std::vector<int> group_stride;
  std::vector<int> chunk_first;                  // first group of every chunk, then ngroup
  size_t arena = 0;                              // doubles
};
static inline AcMiePlan ac_mie_plan(const std::vector<long long> &rec0, const std::vector<int> &nrec, const std::vector<double> &alpha,
                                    size_t budget)
{
  AcMiePlan p;
  std::vector<std::pair<int, int>> items;
  items.reserve(alpha.size());
  bool same_grid = true;                          // tables that start at the same size parameter share one grid: record r of every
  int longest = 0;                                // table has the same alpha, and the sorted list is "r descending, table ascending"
  for (size_t t = 0; t < rec0.size(); ++t) {
    same_grid = same_grid && nrec[t] > 0 && alpha[(size_t)rec0[t]] == alpha[(size_t)rec0[0]];
    longest = std::max(longest, nrec[t]);
  }
  if (same_grid) {
    std::vector<int> by_len(rec0.size());
    for (size_t t = 0; t < rec0.size(); ++t) by_len[t] = (int)t;
    std::stable_sort(by_len.begin(), by_len.end(), [&](int a, int b) { return nrec[a] > nrec[b]; });
    std::vector<int> live;                        // tables with nrec > r, in table order
    size_t next = 0;
    for (int r = longest - 1; r >= 0; --r) {
      bool grew = false;
      while (next < by_len.size() && nrec[by_len[next]] > r) { live.push_back(by_len[next++]); grew = true; }
      if (grew) std::sort(live.begin(), live.end());
      for (int t : live) items.push_back({t, r});
    }
  } else {
    for (size_t t = 0; t < rec0.size(); ++t)
      for (int r = 0; r < nrec[t]; ++r) items.push_back({(int)t, r});
    std::stable_sort(items.begin(), items.end(), [&](const std::pair<int, int> &a, const std::pair<int, int> &b) {
      return alpha[(size_t)rec0[a.first] + a.second] > alpha[(size_t)rec0[b.first] + b.second];
    });
  }
  const size_t ngroup = (items.size() + 31) / 32;
  p.item_table.assign(ngroup * 32, -1);
  p.item_rec.assign(ngroup * 32, 0);
  p.group_off.resize(ngroup);
  p.group_stride.resize(ngroup);
  size_t used = 0;
  p.chunk_first.push_back(0);
  for (size_t g = 0; g < ngroup; ++g) {
    for (size_t l = 0; l < 32 && g * 32 + l < items.size(); ++l) {
      const size_t it = g * 32 + l;
      p.item_table[it] = items[it].first; p.item_rec[it] = items[it].second;
    }
    const double a = alpha[(size_t)rec0[items[g * 32].first] + items[g * 32].second];      // the list is sorted: lane 0 is the largest
    const int stride = (int)trunc(a + a + 20) + 4;
    const size_t need = (size_t)AC_WORK_ARRAYS * stride * 32;
    if (used > 0 && used + need > budget) { p.chunk_first.push_back((int)g); used = 0; }
    p.group_off[g] = (long long)used;
    p.group_stride[g] = stride;
    used += need;
    p.arena = std::max(p.arena, used);
  }
  p.chunk_first.push_back((int)ngroup);
  return p;
}
