/*
 * sos_surface_oracle.c -- TEST INFRASTRUCTURE ONLY (see sos_oracle.h).
 *
 * Scalar C restatement of the rough-sea (Cox & Munk) reflection-matrix pipeline of SOS-ABS V5.1:
 *   SOS_GLITTER (SOS_GLITTER.F:229-371) -> SOS_GSF (:451-711) + SOS_CALCG (:755-784)
 *   SOS_MAT_FRESNEL (SOS_SURFACE.F:1235-1603), SOS_MAT_REFLEXION (:1708-1973),
 *   SOS_NOYAUX_FRESNEL (:2029-2227), SOS_MISE_FORMAT (:2307-2443).
 * The temporary files of the reference are replaced by memory, but their lossy channels are kept:
 * RES_FRESNEL is written with 4(E15.8) (8 significant digits) and the M_ij are REAL*4.
 * PIN: bit-identical to the translated reference chain in oracle/_ref (see sos_oracle.h); no gfortran build available.
 */
#include "sos_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PH_NU 1024      /* CTE_PH_NU  SOS.h:319 */
#define PH_NQ 10        /* CTE_PH_NQ  SOS.h:325 */
#define PH_TEST 10000   /* CTE_PH_TEST SOS.h:312 */

static double orc_pi2(void) { return acos(-1.0); }

/* value of x after WRITE with E15.8 and READ back (SOS_SURFACE.F:1552 fmt 207 -> :1822 fmt 1300) */
static double round_e15_8(double x)
{
  char buf[64];
  snprintf(buf, sizeof buf, "%.7E", x);
  return strtod(buf, NULL);
}

/* SOS_CALCG, SOS_GLITTER.F:755-784 */
static double calcg(double cs12, double c12, double s12, double sig, double phi)
{
  double costetad = -c12 + s12 * cos(phi);
  double x = (1 - costetad) / cs12;
  return x * x * exp(-(x - 1) / sig);
}

/* SOS_GSF for one pair (SOS_GLITTER.F:523-683): Fourier series E(0:IL) of G(theta1,theta2,phi). Returns IL. */
int orc_gsf_pair(double c1, double c2, double sig, int os_nm, double *e)
{
  const double pi = orc_pi2();
  double u[PH_NU + 1];
  double s1 = sqrt(1 - c1 * c1), s2 = sqrt(1 - c2 * c2);
  double c12 = c1 * c2, s12 = s1 * s2;
  double cs12 = (c1 + c2);
  cs12 = .5 * cs12 * cs12;
  double g = calcg(cs12, c12, s12, sig, 0.0);
  u[0] = g;
  double gmax = g;
  g = calcg(cs12, c12, s12, sig, pi);
  u[PH_NU] = g;
  double gmin = g;
  double phib, q, t1 = 0.0, t2 = 0.0;
  double x = PH_TEST * gmin;
  if (x >= gmax) {                                           /* :568-578 */
    phib = pi;
    q = pi / PH_NU;
    for (int i = 1; i <= PH_NU; ++i) u[i] = calcg(cs12, c12, s12, sig, q * i);
  } else {                                                   /* bisection :586-620 */
    double phi1 = 0, phi2 = pi;
    for (;;) {
      phib = .5 * (phi1 + phi2);
      g = calcg(cs12, c12, s12, sig, phib);
      x = PH_TEST * g;
      if (fabs(x - gmax) < (double)0.01f * gmax) break;
      if (x <= gmax) phi2 = phib; else phi1 = phib;
    }
    q = phib / PH_NU;
    for (int i = 1; i <= PH_NU; ++i) u[i] = calcg(cs12, c12, s12, sig, q * i);
    gmin = u[PH_NU];
  }
  int il = os_nm;
  for (int is = 0; is <= os_nm; ++is) {                      /* :644-678 */
    double z = .5 * (gmax + gmin * cos(is * phib));
    int ia = 1;
    for (int i = 1; i <= PH_NQ; ++i) {
      ia = 2 * ia;
      int ip = PH_NU / ia;
      double y = 0;
      for (int j = 1; j <= ia; j += 2) {
        int k = ip * j;
        y = y + u[k] * cos((is * k) * q);
      }
      y = 2 * y / ia;
      double xt = fabs(z - y) / z;
      if (xt < (double)0.0001f) break;
      z = .5 * (y + z);
    }
    e[is] = phib * z / pi;
    if (is == 0) { t1 = e[0]; t2 = e[0]; continue; }
    t1 = t1 + 2 * e[is];
    t2 = t2 + 2 * cos(is * phib * .5) * e[is];
    double b1 = fabs(t1 - gmax) / gmax;
    if (b1 > (double)0.001f) continue;
    il = is;
    break;
  }
  (void)t2;
  return il;
}

/* SOS_MAT_FRESNEL (SOS_SURFACE.F:1327-1553): Legendre expansion coefficients of the Fresnel matrix,
 * returned after the 4(E15.8) text round trip.  rmu/chr: [2N+1]. */
void orc_mat_fresnel(int nbmu, const double *rmu, const double *chr, double ind, int os_ns,
                     double *alpha, double *beta, double *gamma, double *zeta)
{
  const int N = nbmu;
#define VV(a, j) ((a)[(j) + N])
  double *delta = (double *)calloc(os_ns + 1, sizeof(double));
  double *r11 = (double *)calloc(2 * N + 1, sizeof(double));
  double *r12 = (double *)calloc(2 * N + 1, sizeof(double));
  double *r33 = (double *)calloc(2 * N + 1, sizeof(double));
  double *pl = (double *)calloc(os_ns + 3, sizeof(double));   /* PL(-1:NS+1) -> pl[k+1] */
  double *pol = (double *)calloc(os_ns + 2, sizeof(double));
  for (int k = 0; k <= os_ns; ++k) { beta[k] = 0; gamma[k] = 0; alpha[k] = 0; zeta[k] = 0; }
  for (int j = -N; j <= N; ++j) {                            /* :1346-1381 */
    if (j == 0) continue;
    double c = VV(rmu, j);
    c = sqrt(.5 * (1 + c));
    double a = sqrt(ind * ind - 1.0 + c * c);
    double b = ind * ind * c;
    double rl = -(b - a) / (b + a);
    double rr = (c - a) / (c + a);
    VV(r11, j) = .5 * (rl * rl + rr * rr);
    VV(r12, j) = .5 * (rl * rl - rr * rr);
    VV(r33, j) = rl * rr;
  }
  for (int j = -N; j <= N; ++j) {                            /* :1387-1400 */
    if (j == 0) continue;
    double x = VV(r11, j) * VV(chr, j);
    double xrmu = VV(rmu, j);
    pl[0] = 0.0; pl[1] = 1.0;
    for (int k = 0; k <= os_ns; ++k) {
      pl[k + 2] = ((2 * k + 1.) * xrmu * pl[k + 1] - k * pl[k]) / (k + 1.);
      beta[k] = beta[k] + x * pl[k + 1];
    }
  }
  for (int k = 0; k <= os_ns; ++k) beta[k] = (2 * k + 1) * beta[k] * .5;
  for (int j = -N; j <= N; ++j) {                            /* :1433-1456 */
    if (j == 0) continue;
    double xxx = VV(chr, j) * VV(r12, j);
    double xx = VV(chr, j) * VV(r33, j);
    pol[0] = 0.0; pol[1] = 0.0;
    double xrmu = VV(rmu, j);
    pl[0] = 0.0; pl[1] = 1.0;
    pol[2] = 3. * (1. - xrmu * xrmu) / 2. / sqrt(6.0);
    for (int k = 2; k <= os_ns; ++k) {
      double d = (2. * k + 1.) / sqrt(1.0 * (k + 3.) * (k - 1.));
      double e = sqrt(1.0 * (k + 2.) * (k - 2.)) / (2. * k + 1.);
      if (k + 1 <= os_ns + 1) pol[k + 1] = d * (xrmu * pol[k] - e * pol[k - 1]);
      gamma[k] = gamma[k] + xxx * pol[k];
    }
    for (int k = 0; k <= os_ns; ++k) {
      pl[k + 2] = ((2. * k + 1.) * xrmu * pl[k + 1] - k * pl[k]) / (k + 1.);
      delta[k] = delta[k] + xx * pl[k + 1];
    }
  }
  for (int k = 0; k <= os_ns; ++k) {                         /* :1458-1461 */
    delta[k] = delta[k] * (2. * k + 1.) * .5;
    gamma[k] = gamma[k] * (2. * k + 1.) * .5;
  }
  for (int i = 2; i <= os_ns; ++i) {                         /* :1521-1546 */
    /* CO1 and the first CO2 are all-REAL*4 expressions: single-precision arithmetic */
    float co1f = 4 * (2 * i + 1.f) / (float)i / (i - 1.f) / (i + 1.f) / (i + 2.f);
    float co2f = i * (i - 1.f) / ((i + 1.f) * (i + 2.f));
    double co1 = co1f, co2 = co2f;
    double co3 = co2 * delta[i];
    co2 = co2 * beta[i];
    int nn = (int)(i * .5f), mm = (int)((i - 1) * .5f);
    double som1 = 0, som2 = 0, som3 = 0, som4 = 0;
    for (int j = 1; j <= nn; ++j) {
      double x2 = (double)((i - 1.f) * (i - 1.f) - 3.f * (2 * j - 1.f) * (i - j));
      som1 = som1 + x2 * beta[i - 2 * j];
      som2 = som2 + x2 * delta[i - 2 * j];
    }
    for (int j = 0; j <= mm; ++j) {
      double x2 = (double)((i - 1.f) * (i - 1.f) - 3.f * j * (2 * i - 2 * j - 1.f));
      som3 = som3 + x2 * beta[i - 2 * j - 1];
      som4 = som4 + x2 * delta[i - 2 * j - 1];
    }
    zeta[i] = co3 - co1 * (som2 - som3);
    alpha[i] = co2 - co1 * (som1 - som4);
  }
  for (int k = 0; k <= os_ns; ++k) {                         /* text file RES_FRESNEL, 4(E15.8) */
    alpha[k] = round_e15_8(alpha[k]); beta[k] = round_e15_8(beta[k]);
    gamma[k] = round_e15_8(gamma[k]); zeta[k] = round_e15_8(zeta[k]);
  }
  free(delta); free(r11); free(r12); free(r33); free(pl); free(pol);
#undef VV
}

/* SOS_NOYAUX_FRESNEL (SOS_SURFACE.F:2029-2227).  Kernels K[is*2 + (k-1)], k = 1,2. */
static void noyaux_fresnel(double rmu1, double rmu2, int os_ns, const double *alpha, const double *beta,
                           const double *gamma, const double *zeta,
                           double *bp, double *gr, double *gt, double *arr, double *art, double *att)
{
  const int LD = os_ns + 2;
  double *psl = (double *)calloc((size_t)2 * LD, sizeof(double));   /* (-1:NS, 2), persistent across IS like the reference */
  double *rsl = (double *)calloc((size_t)2 * LD, sizeof(double));
  double *tsl = (double *)calloc((size_t)2 * LD, sizeof(double));
#define PS(l, j) psl[((j) - 1) * LD + (l) + 1]
#define RS(l, j) rsl[((j) - 1) * LD + (l) + 1]
#define TS(l, j) tsl[((j) - 1) * LD + (l) + 1]
  const double rac3 = sqrt(3.0), x26 = 2. * sqrt(6.0);
  const double r[3] = {0.0, rmu1, rmu2};
  for (int is = 0; is <= os_ns; ++is) {
    if (is == 0) {
      for (int j = 1; j <= 2; ++j) {
        double c = r[j];
        PS(0, j) = 1; PS(1, j) = c;
        PS(2, j) = (3 * c * c - 1) * 0.5;
        RS(1, j) = 0;
        RS(2, j) = 3 * (1 - c * c) / x26;
        TS(1, j) = 0.; TS(2, j) = 0.;
      }
    } else if (is == 1) {
      for (int j = 1; j <= 2; ++j) {
        double c = r[j];
        double x = 1 - c * c;
        PS(0, j) = 0;
        PS(1, j) = sqrt(x * 0.5);
        PS(2, j) = c * PS(1, j) * rac3;
        TS(1, j) = 0.; RS(1, j) = 0;
        RS(2, j) = -c * sqrt(x) * 0.5;
        TS(2, j) = -sqrt(x) * 0.5;
      }
    } else {
      double a = 1;
      for (int i = 1; i <= is; ++i) { double x = i; a = a * sqrt((i + is) / x) * 0.5; }
      double b = a * sqrt(is / (is + 1.0)) * sqrt((is - 1.0) / (is + 2.));
      for (int j = 1; j <= 2; ++j) {
        double c = r[j];
        double xx = 1 - c * c;
        double yy = (double)(is * 0.5f);
        PS(is - 1, j) = 0.; RS(is - 1, j) = 0.; TS(is - 1, j) = 0.;
        double x = pow(xx, yy);
        PS(is, j) = a * x;
        yy = yy - 1;
        x = pow(xx, yy);
        RS(is, j) = b * (1 + c * c) * x;
        TS(is, j) = 2 * b * c * x;
      }
    }
    int k0 = 2;
    if (is > 2) k0 = is;
    for (int l = k0; l <= os_ns - 1; ++l) {
      double a = (2 * l + 1.) / sqrt((l + is + 1.0) * (l - is + 1.));
      double b = sqrt((double)((l + is) * (l - is))) / (2. * l + 1.);
      double d = (l + 1.) * (2 * l + 1.) / sqrt((l + 3.0) * (l - 1.) * (l + is + 1.) * (l - is + 1.));
      double e = sqrt((l + 2.0) * (l - 2.) * (l + is) * (l - is)) / (l * (2. * l + 1.));
      float ff = (2.f * is) / (l * (l + 1.f));                 /* all-REAL*4 (:2176) */
      double f = ff;
      for (int j = 1; j <= 2; ++j) {
        double c = r[j];
        PS(l + 1, j) = a * (c * PS(l, j) - b * PS(l - 1, j));
        double xr = d * (c * RS(l, j) - f * TS(l, j) - e * RS(l - 1, j));
        double xt = d * (c * TS(l, j) - f * RS(l, j) - e * TS(l - 1, j));
        RS(l + 1, j) = xr;
        TS(l + 1, j) = xt;
      }
    }
    for (int k = 1; k <= 2; ++k) {
      int j = 3 - k;
      double sbp = 0, sarr = 0, satt = 0, sgr = 0, sgt = 0, sart = 0;
      for (int l = is; l <= os_ns; ++l) {
        sbp = sbp + beta[l] * PS(l, j) * PS(l, k);
        sgr = sgr + gamma[l] * PS(l, j) * RS(l, k);
        sgt = sgt + gamma[l] * PS(l, j) * TS(l, k);
        satt = satt + alpha[l] * TS(l, j) * TS(l, k) + zeta[l] * RS(l, j) * RS(l, k);
        sarr = sarr + zeta[l] * TS(l, j) * TS(l, k) + alpha[l] * RS(l, j) * RS(l, k);
        sart = sart + alpha[l] * RS(l, k) * TS(l, j) + zeta[l] * RS(l, j) * TS(l, k);
      }
      bp[is * 2 + k - 1] = sbp; gr[is * 2 + k - 1] = sgr; gt[is * 2 + k - 1] = sgt;
      arr[is * 2 + k - 1] = sarr; art[is * 2 + k - 1] = sart; att[is * 2 + k - 1] = satt;
    }
  }
  free(psl); free(rsl); free(tsl);
#undef PS
#undef RS
#undef TS
}

/*
 * SOS_GLITTER (SOS_GLITTER.F:229-371) in memory.
 *   rmu, chr : [2N+1] cosines and weights (index j at [j+N])
 *   surf     : out, [os_nb+1][9][N][N] REAL*4, surf[s][m][(J-1)*N + (I-1)] = P_m(I,J) (SOS_SURFACE.F:2404-2412)
 *   il_out   : out (may be NULL), [N*(N+1)/2] series lengths IL of SOS_GSF in pair order (I1, I2<=I1)
 */
int orc_glitter(int nbmu, const double *rmu, const double *chr, double wind, double ind,
                int os_nb, int os_ns, int os_nm, float *surf, int *il_out)
{
  const int N = nbmu;
  const double sig = (double)0.003f + (double)0.00512f * wind;   /* SIG = .003 + .00512*WIND (:300) */
  const double coef = 1.0 / sig;                                 /* (1./SIG) :315 */
  double *alpha = (double *)calloc(os_ns + 1, sizeof(double)), *beta = (double *)calloc(os_ns + 1, sizeof(double));
  double *gamma = (double *)calloc(os_ns + 1, sizeof(double)), *zeta = (double *)calloc(os_ns + 1, sizeof(double));
  orc_mat_fresnel(N, rmu, chr, ind, os_ns, alpha, beta, gamma, zeta);
  double *g = (double *)calloc(os_nm + os_ns + os_nb + 2, sizeof(double));
  double *ker = (double *)calloc((size_t)6 * 2 * (os_ns + 1), sizeof(double));
  double *bp = ker, *gr = bp + 2 * (os_ns + 1), *gt = gr + 2 * (os_ns + 1), *arr = gt + 2 * (os_ns + 1),
         *art = arr + 2 * (os_ns + 1), *att = art + 2 * (os_ns + 1);
  const size_t NN = (size_t)N * N;
  int pair = 0;
  for (int i = 1; i <= N; ++i) {
    for (int j = 1; j <= i; ++j, ++pair) {
      memset(g, 0, (os_nm + os_ns + os_nb + 2) * sizeof(double));
      int lim = orc_gsf_pair(rmu[i + N], rmu[j + N], sig, os_nm, g);   /* GSF loops I1 (outer), I2<=I1 */
      for (int x = lim + 1; x <= os_nm; ++x) g[x] = 0.;                 /* SOS_SURFACE.F:1846-1848 */
      if (il_out) il_out[pair] = lim;
      noyaux_fresnel(rmu[i + N], rmu[j + N], os_ns, alpha, beta, gamma, zeta, bp, gr, gt, arr, art, att);
      for (int is = 0; is <= os_nb; ++is) {                             /* :1864-1933 */
        double x = coef * g[is] / 4.;
        double r111 = x * bp[0], r121 = x * gr[0], r122 = x * gr[1], r131 = 0, r132 = 0, r231 = 0, r232 = 0;
        double r211 = x * gr[1], r212 = x * gr[0], r221 = x * arr[1], r222 = x * arr[0];
        double r311 = 0, r312 = 0, r321 = 0, r322 = 0, r331 = x * att[1], r332 = x * att[0];
        int im = 1;
        for (int k = 1; k <= os_ns; ++k) {
          im = -im;
          int i1 = k + is, i2 = abs(k - is);
          if (i1 > lim && i2 > lim) continue;
          double xx = coef * im * (g[i1] + g[i2]) / 4.;
          double yy = coef * im * (g[i2] - g[i1]) / 4.;
          const double b1 = bp[k * 2], g1 = gr[k * 2], g2 = gr[k * 2 + 1], t1 = gt[k * 2], t2 = gt[k * 2 + 1];
          const double a1 = arr[k * 2], a2 = arr[k * 2 + 1], rt1 = art[k * 2], rt2 = art[k * 2 + 1];
          const double tt1 = att[k * 2], tt2 = att[k * 2 + 1];
          r111 = r111 + b1 * xx;
          r121 = r121 + g1 * xx;  r122 = r122 + g2 * xx;
          r131 = r131 + t1 * yy;  r132 = r132 + t2 * yy;
          r211 = r211 + g2 * xx;  r212 = r212 + g1 * xx;
          r221 = r221 + a2 * xx;  r222 = r222 + a1 * xx;
          r231 = r231 + rt2 * yy; r232 = r232 + rt1 * yy;
          r311 = r311 + t2 * yy;  r312 = r312 + t1 * yy;
          r321 = r321 + rt1 * yy; r322 = r322 + rt2 * yy;
          r331 = r331 + tt2 * xx; r332 = r332 + tt1 * xx;
        }
        /* M(IS,1) -> P(I,J), M(IS,2) -> P(J,I)  (SOS_MISE_FORMAT :2378-2395); REAL*4 storage */
        float *rec = surf + (size_t)is * 9 * NN;
        const float m1[9] = {(float)r111, (float)r121, (float)r131, (float)r211, (float)r221, (float)r231,
                             (float)-r311, (float)-r321, (float)-r331};
        const float m2[9] = {(float)r111, (float)r122, (float)r132, (float)r212, (float)r222, (float)r232,
                             (float)-r312, (float)-r322, (float)-r332};
        for (int m = 0; m < 9; ++m) {
          rec[m * NN + (size_t)(j - 1) * N + (i - 1)] = m1[m];   /* P(I,J) */
          rec[m * NN + (size_t)(i - 1) * N + (j - 1)] = m2[m];   /* P(J,I) */
        }
      }
    }
  }
  free(alpha); free(beta); free(gamma); free(zeta); free(g); free(ker);
  return 0;
}
