/*
 * sos_oracle.c -- TEST INFRASTRUCTURE ONLY (see sos_oracle.h).
 *
 * Scalar C restatement of the successive-orders solver of SOS-ABS V5.1:
 *   SOS_OS.F (all routines), SOS.F:496-637, SOS_AGGREGATE.F:289-488,
 *   SOS_TRPHI.F:285-636, 876-1218, 1278-1541, 1843-1907.
 * Same loop nests, same summation order, same REAL*4 literals / REAL*4 sub-expressions
 * as the Fortran source (gfortran evaluates un-suffixed literals and all-REAL*4
 * sub-expressions in single precision before promotion).
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction, no re-association).
 *
 * PIN: bit-identical to oracle/_ref (the reference's own source, mechanically translated; see sos_oracle.h);
 * not checked against a gfortran build (no Fortran compiler, no golden vectors).
 */
#include "sos_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* inc/SOS.h constants with their literal types */
#define SEUIL_CV_SG  ((double)0.00001f)   /* CTE_PH_SEUIL_CV_SG  SOS.h:389 (REAL*4) */
#define SEUIL_SUMDIF ((double)0.00001f)   /* CTE_PH_SEUIL_SUMDIF SOS.h:394 (REAL*4) */
#define SEUIL_VALDIF (1.0e-50)            /* CTE_PH_SEUIL_VALDIF SOS.h:395 (double) */
#define SEUIL_SF     ((double)0.00001f)   /* CTE_PH_SEUIL_SF     SOS.h:400 (REAL*4) */
#define SEUIL_Z      ((double)0.0001f)    /* CTE_SEUIL_Z         SOS.h:407 (REAL*4) */
#define SEUIL_X      ((double)0.00001f)   /* CTE_SEUIL_X         SOS.h:413 (REAL*4) */
#define THRESHOLD_QU (1.0e-15)            /* CTE_THRESHOLD_Q_U_NULL SOS.h:418      */
#define TOA_ALT      (120.0)              /* CTE_TOA_ALT SOS.h:197                  */
#define SOLAR_DISC   (6.8e-05)            /* CTE_SOLAR_DISC_SOLID_ANGLE SOS.h:426   */
#define VALEUR_INDEF (-999.0)             /* INCTE_VALEUR_INDEF SOS_TRPHI.F:134     */

static double orc_pi(void) { return acos(-1.0); } /* INCTE_PI = DACOS(-1.D+00) */

/* index helpers: N and L=NT+1 and W=2N+1 must be in scope */
#define V(a, j)     ((a)[(j) + N])
#define F(a, i, k)  ((a)[(size_t)((k) + N) * L + (i)])
#define P2(a, j, k) ((a)[(size_t)((k) + N) * W + ((j) + N)])

/* ------------------------------------------------------------------------- */
/* SOS_MAT_FRESNEL_PLAN_REFL, SOS_OS.F:1719-1782                              */
void orc_mat_fresnel_plan_refl(int nbmu, const double *rmu, double ind_surf, int ipolar,
                               double *f11, double *f12, double *f33)
{
  const int N = nbmu;
  for (int j = 0; j <= N; ++j) {
    double mu = (j == 0) ? -V(rmu, 0) : V(rmu, j);          /* :1757-1761 */
    double ind2 = ind_surf * ind_surf;
    double mu2 = mu * mu;
    double x = sqrt(ind2 - 1.0 + mu2);                       /* :1765 */
    double rl = (ind2 * mu - x) / (ind2 * mu + x);
    double rr = (mu - x) / (mu + x);
    f11[j] = (rl * rl + rr * rr) / 2.0;
    if (ipolar == 1) {
      f12[j] = (rl * rl - rr * rr) / 2.0;
      f33[j] = rl * rr;
    } else {
      f12[j] = 0.0;
      f33[j] = 0.0;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SOS_NOYAUX, SOS_OS.F:1857-2158                                             */
void orc_noyaux(int is, int nbmu, const double *rmu, int os_nb,
                const double *alpha, const double *beta, const double *gamma, const double *zeta,
                double *xpl, double *xrl, double *xtl,
                double *bp, double *gr, double *gt, double *arr, double *art, double *att)
{
  const int N = nbmu, W = 2 * nbmu + 1;
  /* PSL(-1:NB, -N:N): row l+1.  Zero-initialised like the reference's static storage. */
  const int LD = os_nb + 2;
  double *psl = (double *)calloc((size_t)LD * W, sizeof(double));
  double *rsl = (double *)calloc((size_t)LD * W, sizeof(double));
  double *tsl = (double *)calloc((size_t)LD * W, sizeof(double));
#define PS(l, j) psl[(size_t)((j) + N) * LD + ((l) + 1)]
#define RS(l, j) rsl[(size_t)((j) + N) * LD + ((l) + 1)]
#define TS(l, j) tsl[(size_t)((j) + N) * LD + ((l) + 1)]
  const double rac3 = sqrt(3.0);
  const double x26 = 2.0 * sqrt(6.0);

  if (is == 0) {                                             /* :1968-1993 */
    for (int j = 0; j <= N; ++j) {
      double c = V(rmu, j);
      PS(0, -j) = 1.0; PS(0, j) = 1.0;
      PS(1, j) = c;    PS(1, -j) = -c;
      double x = (3.0 * c * c - 1.0) * 0.5;
      PS(2, -j) = x;   PS(2, j) = x;
      RS(1, j) = 0.0;  RS(1, -j) = 0.0;
      x = 3.0 * (1.0 - c * c) / x26;
      RS(2, -j) = x;   RS(2, j) = x;
      TS(1, j) = 0.0;  TS(1, -j) = 0.0;
      TS(2, j) = 0.0;  TS(2, -j) = 0.0;
    }
    PS(1, 0) = V(rmu, 0);
    RS(1, 0) = 0.0;
  } else if (is == 1) {                                      /* :1997-2023 */
    for (int j = 0; j <= N; ++j) {
      double c = V(rmu, j);
      double x = 1.0 - c * c;
      PS(0, j) = 0.0;  PS(0, -j) = 0.0;
      PS(1, -j) = sqrt(x * 0.5);
      PS(1, j) = sqrt(x * 0.5);
      PS(2, j) = c * PS(1, j) * rac3;
      PS(2, -j) = -PS(2, j);
      RS(1, -j) = 0.0; RS(1, j) = 0.0;
      RS(2, j) = -c * sqrt(x) * 0.5;
      RS(2, -j) = -RS(2, j);
      TS(1, -j) = 0.0; TS(1, j) = 0.0;
      TS(2, j) = -sqrt(x) * 0.5;
      TS(2, -j) = -sqrt(x) * 0.5;
    }
    PS(2, 0) = -PS(2, 0);
    RS(2, 0) = -RS(2, 0);
    RS(1, 0) = 0.0;
    TS(1, 0) = 0.0;
  } else {                                                   /* :2027-2052 */
    double a = 1.0;
    for (int i = 1; i <= is; ++i) {
      double x = (double)i;
      a = a * sqrt((double)(i + is) / x) * 0.5;
    }
    double b = a * sqrt((double)is / ((double)is + 1.0)) * sqrt(((double)is - 1.0) / ((double)is + 2.0));
    for (int j = 0; j <= N; ++j) {
      double c = V(rmu, j);
      double xx = 1.0 - c * c;
      double yy = (double)((float)is * 0.5f - 1.0f);          /* YY=IS*0.5-1. : REAL*4 expr, exact */
      PS(is - 1, j) = 0.0; RS(is - 1, j) = 0.0; TS(is - 1, j) = 0.0;
      double x = a * pow(xx, (double)((float)is * 0.5f));     /* XX**(IS*0.5) */
      PS(is, -j) = x;  PS(is, j) = x;
      x = b * (1.0 + c * c) * pow(xx, yy);
      RS(is, -j) = x;  RS(is, j) = x;
      x = 2.0 * b * c * pow(xx, yy);
      TS(is, -j) = -x; TS(is, j) = x;
    }
  }

  /* recurrences, :2058-2100 */
  int k0 = 2;
  if (is > 2) k0 = is;
  if (k0 != os_nb) {
    int ig = -1;
    if (is == 1) ig = 1;
    for (int l = k0; l <= os_nb - 1; ++l) {
      int lp = l + 1, lm = l - 1;
      double a = (2.0 * l + 1.0) / sqrt(((double)(l + is) + 1.0) * ((double)(l - is) + 1.0));
      double b = sqrt((double)((l + is) * (l - is))) / (2.0 * l + 1.0);
      double d = ((double)l + 1.0) * (2.0 * l + 1.0) /
                 sqrt(((double)l + 3.0) * ((double)l - 1.0) * ((double)(l + is) + 1.0) * ((double)(l - is) + 1.0));
      double e = sqrt(((double)l + 2.0) * ((double)l - 2.0) * (double)(l + is) * (double)(l - is)) /
                 ((double)l * (2.0 * l + 1.0));
      /* F=2.*IS/(L*(L+1.)) is an all-REAL*4 expression: single-precision divide (:2079) */
      float ff = (2.0f * (float)is) / ((float)l * ((float)l + 1.0f));
      double f = (double)ff;
      for (int j = 0; j <= N; ++j) {
        double c = V(rmu, j);
        double x = a * (c * PS(l, j) - b * PS(lm, j));
        PS(lp, j) = x;
        x = d * (c * RS(l, j) - f * TS(l, j) - e * RS(lm, j));
        RS(lp, j) = x;
        x = d * (c * TS(l, j) - f * RS(l, j) - e * TS(lm, j));
        TS(lp, j) = x;
        if (j == 0) continue;
        PS(lp, -j) = ig * PS(lp, j);
        RS(lp, -j) = ig * RS(lp, j);
        TS(lp, -j) = -ig * TS(lp, j);
      }
      ig = -ig;
    }
  }

  for (int j = -N; j <= N; ++j) {                            /* :2107-2111 */
    V(xpl, j) = PS(2, j);
    V(xrl, j) = RS(2, j);
    V(xtl, j) = TS(2, j);
  }

  for (int j = -N; j <= N; ++j) {                            /* :2121-2155 */
    for (int k = -N; k <= N; ++k) {
      double sbp = 0, satt = 0, sarr = 0, sgr = 0, sgt = 0, sart = 0;
      if (!(is > os_nb)) {
        for (int l = is; l <= os_nb; ++l) {
          double r1 = TS(l, j) * TS(l, k);
          double r2 = RS(l, j) * RS(l, k);
          sbp = sbp + beta[l] * PS(l, j) * PS(l, k);
          satt = satt + alpha[l] * r1 + zeta[l] * r2;
          sarr = sarr + zeta[l] * r1 + alpha[l] * r2;
          sgr = sgr + gamma[l] * PS(l, j) * RS(l, k);
          sgt = sgt + gamma[l] * PS(l, j) * TS(l, k);
          sart = sart + alpha[l] * RS(l, k) * TS(l, j) + zeta[l] * RS(l, j) * TS(l, k);
        }
      }
      P2(bp, j, k) = sbp;
      P2(att, j, k) = satt;
      P2(arr, j, k) = sarr;
      P2(gr, j, k) = sgr;
      P2(gt, j, k) = sgt;
      P2(art, j, k) = sart;
    }
  }
  free(psl); free(rsl); free(tsl);
#undef PS
#undef RS
#undef TS
}

/* ------------------------------------------------------------------------- */
/* SOS_FSOURCE_ORDRE1, SOS_OS.F:2431-2565                                     */
void orc_fsource_ordre1(int is, int nbmu, int nt, int jk, const double *xdel, const double *ydel,
                        double beta0, double beta2, double gamma2,
                        const double *xpl, const double *xrl, const double *xtl,
                        const double *bp, const double *gr, const double *gt, const double *ch,
                        double *i2, double *q2, double *u2)
{
  const int N = nbmu, W = 2 * nbmu + 1, L = nt + 1;
  for (int j = -N; j <= N; ++j) {
    double sa1, sa2, sb1, sb2, sc1, sc2;
    if ((is - 2) > 0) {                                      /* :2539-2544 */
      sa2 = P2(bp, jk, j); sa1 = 0.0;
      sb2 = P2(gr, jk, j); sb1 = 0.0;
      sc2 = P2(gt, jk, j); sc1 = 0.0;
    } else {                                                 /* :2526-2532 */
      double spl = V(xpl, jk);
      sa1 = beta0 + beta2 * V(xpl, j) * spl;
      sa2 = P2(bp, jk, j);
      sb1 = gamma2 * V(xrl, j) * spl;
      sb2 = P2(gr, jk, j);
      sc1 = gamma2 * V(xtl, j) * spl;
      sc2 = P2(gt, jk, j);
    }
    for (int k = 0; k <= nt; ++k) {                          /* :2553-2560 */
      double attdir = ch[k], pcray = ydel[k], pcaer = xdel[k];
      F(i2, k, j) = attdir * (sa2 * pcaer + sa1 * pcray);
      F(q2, k, j) = attdir * (sb2 * pcaer + sb1 * pcray);
      F(u2, k, j) = -attdir * (sc2 * pcaer + sc1 * pcray);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SOS_FSOURCE_ORDREIG, SOS_OS.F:2663-3017                                    */
void orc_fsource_ordreig(int is, int nbmu, int nt, const double *xdel, const double *ydel,
                         double beta0, double beta2, double gamma2, double alpha2,
                         const double *xpl, const double *xrl, const double *xtl,
                         const double *i1, const double *q1, const double *u1,
                         const double *bp, const double *gr, const double *gt,
                         const double *arr, const double *art, const double *att, const double *ga,
                         double *i2, double *q2, double *u2)
{
  const int N = nbmu, W = 2 * nbmu + 1, L = nt + 1;
  const int ray = !((is - 2) > 0);                           /* :2810 */
  for (int k = 1; k <= N; ++k) {
    double xpk = 0, xrk = 0, xtk = 0, ypk = 0, yrk = 0, ytk = 0;
    if (ray) {
      xpk = V(xpl, k);  xrk = V(xrl, k);  xtk = V(xtl, k);
      ypk = V(xpl, -k); yrk = V(xrl, -k); ytk = V(xtl, -k);
    }
    for (int i = 0; i <= nt; ++i) {
      double ii1 = 0, ii2 = 0, qq1 = 0, qq2 = 0, uu1 = 0, uu2 = 0;
      double pcaer = xdel[i], pcray = ydel[i];
      for (int j = 1; j <= N; ++j) {
        double bpjk, bpjmk, gtjmk, gtjk, gtkmj, gtkj, grjk, grjmk, grkj, grkmj;
        double arrjk, arrjmk, artjk, artjmk, artkj, artkmj, attjmk, attjk;
        if (ray) {                                           /* :2852-2876 */
          double xpj = V(xpl, j), xrj = V(xrl, j), xtj = V(xtl, j);
          double yrj = V(xrl, -j), ytj = V(xtl, -j);
          bpjk = P2(bp, j, k) * pcaer + pcray * (beta0 + beta2 * xpj * xpk);
          bpjmk = P2(bp, j, -k) * pcaer + pcray * (beta0 + beta2 * xpj * ypk);
          gtjmk = P2(gt, j, -k) * pcaer + pcray * (gamma2 * xpj * ytk);
          gtjk = P2(gt, j, k) * pcaer + pcray * (gamma2 * xpj * xtk);
          gtkmj = P2(gt, k, -j) * pcaer + pcray * (gamma2 * xpk * ytj);
          gtkj = P2(gt, k, j) * pcaer + pcray * (gamma2 * xpk * xtj);
          grjk = P2(gr, j, k) * pcaer + pcray * (gamma2 * xpj * xrk);
          grjmk = P2(gr, j, -k) * pcaer + pcray * (gamma2 * xpj * yrk);
          grkj = P2(gr, k, j) * pcaer + pcray * (gamma2 * xpk * xrj);
          grkmj = P2(gr, k, -j) * pcaer + pcray * (gamma2 * xpk * yrj);
          arrjk = P2(arr, j, k) * pcaer + pcray * (alpha2 * xrj * xrk);
          arrjmk = P2(arr, j, -k) * pcaer + pcray * (alpha2 * xrj * yrk);
          artjk = P2(art, j, k) * pcaer + pcray * (alpha2 * xtj * xrk);
          artjmk = P2(art, j, -k) * pcaer + pcray * (alpha2 * xtj * yrk);
          artkj = P2(art, k, j) * pcaer + pcray * (alpha2 * xtk * xrj);
          artkmj = P2(art, k, -j) * pcaer + pcray * (alpha2 * xtk * yrj);
          attjmk = P2(att, j, -k) * pcaer + pcray * (alpha2 * xtj * ytk);
          attjk = P2(att, j, k) * pcaer + pcray * (alpha2 * xtj * xtk);
        } else {                                             /* :2951-2968 */
          bpjk = P2(bp, j, k) * pcaer;     bpjmk = P2(bp, j, -k) * pcaer;
          gtjmk = P2(gt, j, -k) * pcaer;   gtjk = P2(gt, j, k) * pcaer;
          gtkmj = P2(gt, k, -j) * pcaer;   gtkj = P2(gt, k, j) * pcaer;
          grjk = P2(gr, j, k) * pcaer;     grjmk = P2(gr, j, -k) * pcaer;
          grkj = P2(gr, k, j) * pcaer;     grkmj = P2(gr, k, -j) * pcaer;
          arrjk = P2(arr, j, k) * pcaer;   arrjmk = P2(arr, j, -k) * pcaer;
          artjk = P2(art, j, k) * pcaer;   artjmk = P2(art, j, -k) * pcaer;
          artkj = P2(art, k, j) * pcaer;   artkmj = P2(art, k, -j) * pcaer;
          attjmk = P2(att, j, -k) * pcaer; attjk = P2(att, j, k) * pcaer;
        }
        double z = V(ga, j);
        double xi1 = F(i1, i, j), xi2 = F(i1, i, -j);
        double xq1 = F(q1, i, j), xq2 = F(q1, i, -j);
        double xu1 = F(u1, i, j), xu2 = F(u1, i, -j);
        /* :2894-2905 */
        ii2 = ii2 + z * (xi1 * bpjk + xi2 * bpjmk + xq1 * grkj + xq2 * grkmj - xu1 * gtkj - xu2 * gtkmj);
        ii1 = ii1 + z * (xi1 * bpjmk + xi2 * bpjk + xq1 * grkmj + xq2 * grkj + xu1 * gtkmj + xu2 * gtkj);
        qq2 = qq2 + z * (xi1 * grjk + xi2 * grjmk + xq1 * arrjk + xq2 * arrjmk + xu2 * artjmk - xu1 * artjk);
        qq1 = qq1 + z * (xi1 * grjmk + xi2 * grjk + xq1 * arrjmk + xq2 * arrjk - xu1 * artjmk + xu2 * artjk);
        uu2 = uu2 - z * (xi1 * gtjk - xi2 * gtjmk + xq1 * artkj + xq2 * artkmj - xu1 * attjk - xu2 * attjmk);
        uu1 = uu1 - z * (xi1 * gtjmk - xi2 * gtjk - xq1 * artkmj - xq2 * artkj - xu1 * attjmk - xu2 * attjk);
      }
      F(i2, i, k) = ii2 * 0.5;  F(i2, i, -k) = ii1 * 0.5;
      F(q2, i, k) = qq2 * 0.5;  F(q2, i, -k) = qq1 * 0.5;
      F(u2, i, k) = uu2 * 0.5;  F(u2, i, -k) = uu1 * 0.5;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SOS_INTEGR_EPOPT, SOS_OS.F:2222-2357                                       */
void orc_integr_epopt(int nbmu, const double *rmu, int nt, const double *h,
                      const double *i2, const double *q2, const double *u2,
                      double *i1, double *q1, double *u1)
{
  const int N = nbmu, L = nt + 1;
  for (int k = 1; k <= N; ++k) {                             /* upward, :2279-2310 */
    double rmuk = V(rmu, k);
    double zi1 = F(i1, nt, k), zq1 = F(q1, nt, k), zu1 = F(u1, nt, k);
    for (int i = nt - 1; i >= 0; --i) {
      int jj = i + 1;
      double dtau = h[jj] - h[i];
      double att = exp(-dtau / rmuk);
      double matt = 1.0 - att;
      double attdtau = att * dtau;
      double b = F(i2, i, k);
      double a = (F(i2, jj, k) - b) / dtau;
      zi1 = zi1 * att + matt * (a * rmuk + b) - a * attdtau;
      F(i1, i, k) = zi1;
      b = F(q2, i, k);
      a = (F(q2, jj, k) - b) / dtau;
      zq1 = zq1 * att + matt * (a * rmuk + b) - a * attdtau;
      F(q1, i, k) = zq1;
      b = F(u2, i, k);
      a = (F(u2, jj, k) - b) / dtau;
      zu1 = zu1 * att + matt * (a * rmuk + b) - a * attdtau;
      F(u1, i, k) = zu1;
    }
  }
  for (int k = -N; k <= -1; ++k) {                           /* downward, :2320-2354 */
    double rmuk = V(rmu, k);
    F(i1, 0, k) = 0.0; F(q1, 0, k) = 0.0; F(u1, 0, k) = 0.0;
    double zi1 = 0.0, zq1 = 0.0, zu1 = 0.0;
    for (int i = 1; i <= nt; ++i) {
      int jj = i - 1;
      double dtau = h[i] - h[jj];
      double att = exp(dtau / rmuk);
      double matt = 1.0 - att;
      double attdtau = att * dtau;
      double b = F(i2, i, k);
      double a = (b - F(i2, jj, k)) / dtau;
      zi1 = zi1 * att + matt * (a * rmuk + b) + a * attdtau;
      F(i1, i, k) = zi1;
      b = F(q2, i, k);
      a = (b - F(q2, jj, k)) / dtau;
      zq1 = zq1 * att + matt * (a * rmuk + b) + a * attdtau;
      F(q1, i, k) = zq1;
      b = F(u2, i, k);
      a = (b - F(u2, jj, k)) / dtau;
      zu1 = zu1 * att + matt * (a * rmuk + b) + a * attdtau;
      F(u1, i, k) = zu1;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SOS_FSOURCE_DIFF_FRESNEL1, SOS_OS.F:3106-3295                              */
static void orc_fsource_diff_fresnel1(double f11sun, double f12sun, const double *xdel, const double *ydel, int nt,
                                      double beta0, double beta2, double gamma2, double alpha2,
                                      const double *bp, const double *gr, const double *gt,
                                      const double *arr, const double *art,
                                      const double *xpl, const double *xrl, const double *xtl,
                                      int is, int nbmu, double mus, const double *h,
                                      double *i2, double *q2, double *u2)
{
  const int N = nbmu, W = 2 * nbmu + 1, L = nt + 1;
  for (int j = -N; j <= N; ++j)
    for (int k = 0; k <= nt; ++k) { F(i2, k, j) = 0.0; F(q2, k, j) = 0.0; F(u2, k, j) = 0.0; }
  double coefnt = exp(2.0 * h[nt] / mus) / 4.0;              /* :3219 */
  double spl = V(xpl, 0);
  for (int k = 0; k <= nt - 1; ++k) {
    double yr = ydel[k], xp = xdel[k], yyr = ydel[k + 1], xxp = xdel[k + 1];
    for (int j = 1; j <= N; ++j) {
      double bp0mj, bp0j, grj0, gr0j, gr0mj, grmj0, gt0mj, gt0j, arr0mj, arr0j, artj0, artmj0;
      if (is <= 2) {                                         /* :3237-3252 */
        bp0mj = P2(bp, 0, -j) * xp + (beta0 + beta2 * V(xpl, -j) * spl) * yr;
        bp0j = P2(bp, 0, j) * xxp + (beta0 + beta2 * V(xpl, j) * spl) * yyr;
        grj0 = P2(gr, j, 0) * xxp + yyr * V(xrl, 0) * V(xpl, j) * gamma2;
        gr0j = P2(gr, 0, j) * xxp + yyr * V(xrl, j) * V(xpl, 0) * gamma2;
        gr0mj = P2(gr, 0, -j) * xp + yr * V(xrl, -j) * spl * gamma2;
        grmj0 = P2(gr, -j, 0) * xp + yr * gamma2 * V(xrl, 0) * V(xpl, -j);
        gt0mj = P2(gt, 0, -j) * xp + yr * gamma2 * spl * V(xtl, -j);
        gt0j = P2(gt, 0, j) * xxp + yyr * gamma2 * spl * V(xtl, j);
        arr0mj = P2(arr, 0, -j) * xp + alpha2 * yr * V(xrl, 0) * V(xrl, -j);
        arr0j = P2(arr, 0, j) * xxp + alpha2 * yyr * V(xrl, 0) * V(xrl, j);
        artj0 = P2(art, j, 0) * xxp + yyr * alpha2 * V(xtl, j) * V(xrl, 0);
        artmj0 = P2(art, -j, 0) * xp + yr * alpha2 * V(xtl, -j) * V(xrl, 0);
      } else {                                               /* :3256-3271 */
        bp0mj = P2(bp, 0, -j) * xp;   bp0j = P2(bp, 0, j) * xxp;
        grj0 = P2(gr, j, 0) * xxp;    gr0j = P2(gr, 0, j) * xxp;
        gr0mj = P2(gr, 0, -j) * xp;   grmj0 = P2(gr, -j, 0) * xp;
        gt0mj = P2(gt, 0, -j) * xp;   gt0j = P2(gt, 0, j) * xxp;
        arr0mj = P2(arr, 0, -j) * xp; arr0j = P2(arr, 0, j) * xxp;
        artj0 = P2(art, j, 0) * xxp;  artmj0 = P2(art, -j, 0) * xp;
      }
      double coefk = coefnt * exp(-h[k] / mus);              /* :3278 */
      F(i2, k, j) = coefk * (f11sun * bp0mj + f12sun * grmj0);
      F(q2, k, j) = coefk * (f11sun * gr0mj + f12sun * arr0mj);
      F(u2, k, j) = coefk * (f11sun * gt0mj + f12sun * artmj0);
      double coefkp1 = coefnt * exp(-h[k + 1] / mus);        /* :3285 */
      F(i2, k + 1, -j) = coefkp1 * (f11sun * bp0j + f12sun * grj0);
      F(q2, k + 1, -j) = coefkp1 * (f11sun * gr0j + f12sun * arr0j);
      F(u2, k + 1, -j) = coefkp1 * (f11sun * gt0j + f12sun * artj0);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SOS_PARAM_CONV, SOS_OS.F:3377-3461                                         */
static double orc_param_conv(int N, const double *a1, const double *b1, const double *c1,
                             const double *d1, const double *e1, const double *f1,
                             const double *g1, const double *h1, const double *p1,
                             const double *i3, const double *q3, const double *u3)
{
  double z1 = 0.0;
  for (int k = -N; k <= N; ++k) {
    if (k == 0) continue;
    if (V(a1, k) != 0.0 && V(d1, k) != 0.0 && V(i3, k) != 0.0) {
      double r = 1 - V(g1, k) / V(d1, k);
      double y = (V(g1, k) / V(d1, k) - V(d1, k) / V(a1, k)) / (r * r) * (V(g1, k) / V(i3, k));
      z1 = fmax(z1, fabs(y));
    }
    if (V(b1, k) != 0.0 && V(e1, k) != 0.0 && V(q3, k) != 0.0) {
      double r = 1 - V(h1, k) / V(e1, k);
      double y = (V(h1, k) / V(e1, k) - V(e1, k) / V(b1, k)) / (r * r) * (V(h1, k) / V(q3, k));
      z1 = fmax(z1, fabs(y));
    }
    if (V(c1, k) != 0.0 && V(f1, k) != 0.0 && V(u3, k) != 0.0) {
      double r = 1 - V(p1, k) / V(f1, k);
      double y = (V(p1, k) / V(f1, k) - V(f1, k) / V(c1, k)) / (r * r) * (V(p1, k) / V(u3, k));
      z1 = fmax(z1, fabs(y));
    }
  }
  return z1;
}

/* SOS_ARRET_DIFFUS_1, SOS_OS.F:3497-3547 */
static double orc_arret_diffus_1(int N, int nt, const double *i1, const double *q1, const double *u1)
{
  const int L = nt + 1;
  double z1 = 0.0;
  for (int k = -N; k <= N; ++k) {
    if (k == 0) continue;
    int ind = (k < 0) ? nt : 0;
    z1 = fmax(z1, fabs(F(i1, ind, k)));
    z1 = fmax(z1, fabs(F(q1, ind, k)));
    z1 = fmax(z1, fabs(F(u1, ind, k)));
  }
  return z1;
}

/* SOS_ARRET_DIFFUS_2, SOS_OS.F:3585-3659 */
static double orc_arret_diffus_2(int N, int nt, const double *i1, const double *q1, const double *u1,
                                 const double *i3, const double *q3, const double *u3)
{
  const int L = nt + 1;
  double z1 = 0.0;
  for (int k = -N; k <= N; ++k) {
    if (k == 0) continue;
    int ind = (k < 0) ? nt : 0;
    if (V(i3, k) != 0.0) z1 = fmax(z1, fabs(F(i1, ind, k) / V(i3, k)));
    if (V(q3, k) != 0.0) z1 = fmax(z1, fabs(F(q1, ind, k) / V(q3, k)));
    if (V(u3, k) != 0.0) z1 = fmax(z1, fabs(F(u1, ind, k) / V(u3, k)));
  }
  return z1;
}

/* SOS_ARRET_FOURIER, SOS_OS.F:3709-3796 */
static double orc_arret_fourier(int N, const double *i3, const double *q3, const double *u3,
                                const double *i4, const double *q4, const double *u4,
                                const double *i5, const double *q5, const double *u5)
{
  double z1 = 0.0;
  for (int j = -N; j <= N; ++j) {
    if (j == 0) continue;
    if (V(q4, j) != 0.0) z1 = fmax(z1, fabs(V(q3, j) / V(q4, j)));
    if (V(i4, j) != 0.0) z1 = fmax(z1, fabs(V(i3, j) / V(i4, j)));
    if (V(u4, j) != 0.0) z1 = fmax(z1, fabs(V(u3, j) / V(u4, j)));
    if (V(q5, j) != 0.0) z1 = fmax(z1, fabs(V(q3, j) / V(q5, j)));
    if (V(u5, j) != 0.0) z1 = fmax(z1, fabs(V(u3, j) / V(u5, j)));
    if (V(i5, j) != 0.0) z1 = fmax(z1, fabs(V(i3, j) / V(i5, j)));
  }
  return z1;
}

/* SOS_AJOUT_QUEUE, SOS_OS.F:3871-4018 */
static void orc_ajout_queue(int nt, int N, const double *d1, const double *e1, const double *f1,
                            const double *g1, const double *h1, const double *p1,
                            const double *d1out, const double *e1out, const double *f1out,
                            const double *g1out, const double *h1out, const double *p1out,
                            double *i3, double *q3, double *u3, double *i3out, double *q3out, double *u3out)
{
  const int L = nt + 1;
  for (int j = -N; j <= N; ++j) {
    if (j == 0) continue;
    double iq = (V(d1, j) == 0.0) ? 0.0 : V(g1, j) / (1 - V(g1, j) / V(d1, j));
    double qq = (V(e1, j) == 0.0) ? 0.0 : V(h1, j) / (1 - V(h1, j) / V(e1, j));
    double uq = (V(f1, j) == 0.0) ? 0.0 : V(p1, j) / (1 - V(p1, j) / V(f1, j));
    V(i3, j) = V(i3, j) + iq;
    V(q3, j) = V(q3, j) + qq;
    V(u3, j) = V(u3, j) + uq;
  }
  for (int j = -N; j <= N; ++j) {
    if (j == 0) continue;
    for (int i = 0; i <= nt; ++i) {
      double iq = (F(d1out, i, j) == 0.0) ? 0.0 : F(g1out, i, j) / (1 - F(g1out, i, j) / F(d1out, i, j));
      double qq = (F(e1out, i, j) == 0.0) ? 0.0 : F(h1out, i, j) / (1 - F(h1out, i, j) / F(e1out, i, j));
      double uq = (F(f1out, i, j) == 0.0) ? 0.0 : F(p1out, i, j) / (1 - F(p1out, i, j) / F(f1out, i, j));
      F(i3out, i, j) = F(i3out, i, j) + iq;
      F(q3out, i, j) = F(q3out, i, j) + qq;
      F(u3out, i, j) = F(u3out, i, j) + uq;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SOS_OS, SOS_OS.F:303-1674                                                  */
int orc_sos_os(int nbmu, double *rmu, const double *ga, int os_nb, int nt,
               int n0, double tetas, double ro, int imat_surf, int ifresnel, double ind_surf,
               const double *h, const double *xdel, const double *ydel, const double *zprof,
               double ron, double *alpha, double *beta, double *gamma, double *zeta,
               double zout, int igmax, int iborm, int ipolar, const float *surf,
               double *rec, int *n_fourier, int *n_scatter, int *stop_reason,
               double *emoins, double *eplus)
{
  const int N = nbmu, W = 2 * nbmu + 1, L = nt + 1;
  (void)beta;
  if (n_fourier) *n_fourier = 0;

  double aaa = ron / (2 - ron);                              /* :678-684 */
  aaa = (1 - aaa) / (1 + 2 * aaa);
  double beta0 = 1.0;
  double beta2 = 0.5 * aaa;
  double gamma2 = -aaa * sqrt(1.5);
  double alpha2 = 3.0 * aaa;
  if (ipolar == 0) {                                         /* :689-699 */
    gamma2 = 0.0; alpha2 = 0.0;
    for (int k = 0; k <= os_nb; ++k) { alpha[k] = 0.0; gamma[k] = 0.0; zeta[k] = 0.0; }
  }
  double tab;
  if (n0 > 0) tab = -V(rmu, n0);                             /* :706-710 */
  else tab = -cos(orc_pi() * tetas / 180.0);
  const int jk = 0;
  V(rmu, jk) = tab;                                          /* :715 */

  if (tab == 0.0) return 0;                                  /* :807 limb incidence: IER stays 0 */
  if (((zout < 0) && (zout != -1.0)) || (zout > TOA_ALT)) return -1; /* :811 */

  double *f11 = (double *)calloc(N + 1, sizeof(double));
  double *f12 = (double *)calloc(N + 1, sizeof(double));
  double *f33 = (double *)calloc(N + 1, sizeof(double));
  if (ifresnel == 1) orc_mat_fresnel_plan_refl(N, rmu, ind_surf, ipolar, f11, f12, f33);

  const size_t FS = (size_t)W * L;
#define NEWF() ((double *)calloc(FS, sizeof(double)))
#define NEWV() ((double *)calloc(W, sizeof(double)))
  double *ch = (double *)calloc(L, sizeof(double));
  for (int i = 0; i <= nt; ++i) ch[i] = exp(-h[i] / (-tab)) / 4.0;  /* :837-839 */

  double *i4 = NEWV(), *q4 = NEWV(), *u4 = NEWV(), *i5 = NEWV(), *q5 = NEWV(), *u5 = NEWV();
  double *i3 = NEWV(), *q3 = NEWV(), *u3 = NEWV();
  double *a1 = NEWV(), *b1 = NEWV(), *c1 = NEWV(), *d1 = NEWV(), *e1 = NEWV(), *f1 = NEWV();
  double *g1 = NEWV(), *h1 = NEWV(), *p1 = NEWV();
  double *i3z = NEWV(), *q3z = NEWV(), *u3z = NEWV();
  double *xpl = NEWV(), *xrl = NEWV(), *xtl = NEWV();
  double *xr = (double *)calloc(N + 1, sizeof(double));
  double *rii = (double *)calloc(N + 1, sizeof(double));
  double *rqq = (double *)calloc(N + 1, sizeof(double));
  double *ruu = (double *)calloc(N + 1, sizeof(double));
  double *riiout = (double *)calloc((size_t)(N + 1) * L, sizeof(double));
  double *rqqout = (double *)calloc((size_t)(N + 1) * L, sizeof(double));
  double *ruuout = (double *)calloc((size_t)(N + 1) * L, sizeof(double));
#define RO_(a, i, k) ((a)[(size_t)(k) * L + (i)])
  double *bp = (double *)calloc((size_t)W * W, sizeof(double));
  double *gr = (double *)calloc((size_t)W * W, sizeof(double));
  double *gt = (double *)calloc((size_t)W * W, sizeof(double));
  double *arr = (double *)calloc((size_t)W * W, sizeof(double));
  double *art = (double *)calloc((size_t)W * W, sizeof(double));
  double *att = (double *)calloc((size_t)W * W, sizeof(double));
  double *i1 = NEWF(), *q1 = NEWF(), *u1 = NEWF(), *i2 = NEWF(), *q2 = NEWF(), *u2 = NEWF();
  double *i1f = NEWF(), *q1f = NEWF(), *u1f = NEWF();
  double *i3out = NEWF(), *q3out = NEWF(), *u3out = NEWF();
  double *d1out = NEWF(), *e1out = NEWF(), *f1out = NEWF();
  double *g1out = NEWF(), *h1out = NEWF(), *p1out = NEWF();
  float *rs = (float *)calloc((size_t)9 * N * N, sizeof(float));
#define RM(m, I, J) ((double)rs[(size_t)(m) * N * N + (size_t)((J) - 1) * N + ((I) - 1)])
  /* m: 0 R11, 1 R12, 2 R13, 3 R21, 4 R22, 5 R23, 6 R31, 7 R32, 8 R33 */

  double sign = -1.0;
  int nrec = 0;
  if (emoins) *emoins = 0.0;
  if (eplus) *eplus = 0.0;

  for (int is = 0; is <= iborm; ++is) {                      /* :872 */
    sign = -sign;
    if (is > 0) beta0 = 0.0;                                 /* :890 */
    for (int j = -N; j <= N; ++j) {
      V(i3, j) = 0.0; V(q3, j) = 0.0; V(u3, j) = 0.0;
      for (int i = 0; i <= nt; ++i) { F(i3out, i, j) = 0.0; F(q3out, i, j) = 0.0; F(u3out, i, j) = 0.0; }
    }
    if (imat_surf == 1) {                                    /* :912-943 */
      memcpy(rs, surf + (size_t)is * 9 * N * N, (size_t)9 * N * N * sizeof(float));
      if (ipolar == 0) memset(rs + (size_t)N * N, 0, (size_t)8 * N * N * sizeof(float));
    }
    orc_noyaux(is, N, rmu, os_nb, alpha, beta, gamma, zeta, xpl, xrl, xtl, bp, gr, gt, arr, art, att);
    orc_fsource_ordre1(is, N, nt, jk, xdel, ydel, beta0, beta2, gamma2, xpl, xrl, xtl, bp, gr, gt, ch, i2, q2, u2);

    for (int k = 1; k <= N; ++k) {                           /* :970-992 */
      F(i1, nt, k) = 0.0; F(q1, nt, k) = 0.0; F(u1, nt, k) = 0.0;
      xr[k] = 0.0;
      if (!(ro == 0.0 || is != 0)) {
        F(i1, nt, k) = -ro * tab * exp(h[nt] / tab);
        xr[k] = F(i1, nt, k);
      }
      if (imat_surf == 1) {
        double rr = exp(h[nt] / tab) / V(rmu, k);
        F(i1, nt, k) = F(i1, nt, k) + RM(0, n0, k) * rr;
        F(q1, nt, k) = RM(3, n0, k) * rr;
        F(u1, nt, k) = RM(6, n0, k) * rr;
      }
    }
    orc_integr_epopt(N, rmu, nt, h, i2, q2, u2, i1, q1, u1); /* :998 */

    if (ifresnel == 1) {                                     /* :1010-1043 */
      orc_fsource_diff_fresnel1(f11[0], f12[0], xdel, ydel, nt, beta0, beta2, gamma2, alpha2,
                                bp, gr, gt, arr, art, xpl, xrl, xtl, is, N, tab, h, i2, q2, u2);
      for (int k = 1; k <= N; ++k) { F(i1f, nt, k) = 0.0; F(q1f, nt, k) = 0.0; F(u1f, nt, k) = 0.0; }
      orc_integr_epopt(N, rmu, nt, h, i2, q2, u2, i1f, q1f, u1f);
      for (int i = 0; i <= nt; ++i)
        for (int k = -N; k <= N; ++k) {
          if (k == 0) continue; /* column 0 is never used */
          F(i1, i, k) = F(i1, i, k) + F(i1f, i, k);
          F(q1, i, k) = F(q1, i, k) + F(q1f, i, k);
          F(u1, i, k) = F(u1, i, k) + F(u1f, i, k);
        }
    }

    for (int k = 1; k <= N; ++k) {                           /* :1051-1084 */
      rii[k] = 0.0; rqq[k] = 0.0; ruu[k] = 0.0;
      for (int i = 0; i <= nt; ++i) { RO_(riiout, i, k) = 0.0; RO_(rqqout, i, k) = 0.0; RO_(ruuout, i, k) = 0.0; }
    }
    if (imat_surf == 1) {
      for (int k = 1; k <= N; ++k) {
        for (int i = 0; i <= nt; ++i) {
          double a = -(h[nt] - h[i]) / V(rmu, k);
          a = exp(a);
          RO_(riiout, i, k) = a * (F(i1, nt, k) - xr[k]);
          RO_(rqqout, i, k) = a * F(q1, nt, k);
          RO_(ruuout, i, k) = a * F(u1, nt, k);
        }
        double a = -h[nt] / V(rmu, k);
        a = exp(a);
        rii[k] = a * (F(i1, nt, k) - xr[k]);
        rqq[k] = a * F(q1, nt, k);
        ruu[k] = a * F(u1, nt, k);
      }
    }

    for (int k = -N; k <= -1; ++k) {                         /* :1094-1113 */
      V(i3, k) = F(i1, nt, k); V(q3, k) = F(q1, nt, k); V(u3, k) = F(u1, nt, k);
      V(d1, k) = F(i1, nt, k); V(e1, k) = F(q1, nt, k); V(f1, k) = F(u1, nt, k);
      for (int i = 0; i <= nt; ++i) { F(i3out, i, k) = F(i1, i, k); F(q3out, i, k) = F(q1, i, k); F(u3out, i, k) = F(u1, i, k); }
    }
    for (int k = 1; k <= N; ++k) {                           /* :1118-1137 */
      V(i3, k) = F(i1, 0, k); V(q3, k) = F(q1, 0, k); V(u3, k) = F(u1, 0, k);
      V(d1, k) = F(i1, 0, k); V(e1, k) = F(q1, 0, k); V(f1, k) = F(u1, 0, k);
      for (int i = 0; i <= nt; ++i) { F(i3out, i, k) = F(i1, i, k); F(q3out, i, k) = F(q1, i, k); F(u3out, i, k) = F(u1, i, k); }
    }

    int ig = 1;
    int reason;
    for (;;) {                                               /* label 503 */
      ig = ig + 1;
      if (ig > igmax) { reason = SOS_STOP_IGMAX_PRE; break; }
      orc_fsource_ordreig(is, N, nt, xdel, ydel, beta0, beta2, gamma2, alpha2, xpl, xrl, xtl,
                          i1, q1, u1, bp, gr, gt, arr, art, att, ga, i2, q2, u2);
      for (int k = 1; k <= N; ++k) { F(i1, nt, k) = 0.0; F(q1, nt, k) = 0.0; F(u1, nt, k) = 0.0; xr[k] = 0.0; }
      double lsol = 0.0;                                     /* :1177-1181 */
      for (int j = 1; j <= N; ++j) lsol = lsol + V(ga, j) * F(i1, nt, -j) * V(rmu, j);
      lsol = 2 * lsol * ro;
      if (!(ro == 0.0 || is != 0)) {
        for (int j = 1; j <= N; ++j) { F(i1, nt, j) = lsol; xr[j] = lsol; }
      }
      if (imat_surf == 1) {                                  /* :1194-1220 */
        for (int k = 1; k <= N; ++k) {
          double ii2 = 0.0, qq2 = 0.0, uu2 = 0.0;
          double rrmu = 2 / V(rmu, k);
          for (int j = 1; j <= N; ++j) {
            double z = V(ga, j);
            double xi1 = F(i1, nt, -j), xq1 = F(q1, nt, -j), xu1 = F(u1, nt, -j);
            ii2 = ii2 + z * (xi1 * RM(0, j, k) + xq1 * RM(1, j, k) + xu1 * RM(2, j, k));
            qq2 = qq2 + z * (xi1 * RM(3, j, k) + xq1 * RM(4, j, k) + xu1 * RM(5, j, k));
            uu2 = uu2 + z * (xi1 * RM(6, j, k) + xq1 * RM(7, j, k) + xu1 * RM(8, j, k));
          }
          F(i1, nt, k) = ii2 * rrmu + xr[k];
          F(q1, nt, k) = qq2 * rrmu;
          F(u1, nt, k) = uu2 * rrmu;
        }
      }
      if (ifresnel == 1) {                                   /* :1225-1239 */
        for (int k = 1; k <= N; ++k) {
          F(i1, nt, k) = F(i1, nt, k) + f11[k] * F(i1, nt, -k) + f12[k] * F(q1, nt, -k);
          F(q1, nt, k) = F(q1, nt, k) + f12[k] * F(i1, nt, -k) + f11[k] * F(q1, nt, -k);
          F(u1, nt, k) = F(u1, nt, k) + f33[k] * F(u1, nt, -k);
        }
      }
      orc_integr_epopt(N, rmu, nt, h, i2, q2, u2, i1, q1, u1); /* :1244 */

      for (int k = -N; k <= -1; ++k) {                       /* :1248-1263 */
        V(g1, k) = F(i1, nt, k); V(h1, k) = F(q1, nt, k); V(p1, k) = F(u1, nt, k);
        for (int i = 0; i <= nt; ++i) { F(g1out, i, k) = F(i1, i, k); F(h1out, i, k) = F(q1, i, k); F(p1out, i, k) = F(u1, i, k); }
      }
      for (int k = 1; k <= N; ++k) {                         /* :1265-1280 */
        V(g1, k) = F(i1, 0, k); V(h1, k) = F(q1, 0, k); V(p1, k) = F(u1, 0, k);
        for (int i = 0; i <= nt; ++i) { F(g1out, i, k) = F(i1, i, k); F(h1out, i, k) = F(q1, i, k); F(p1out, i, k) = F(u1, i, k); }
      }

      if (ig != 2) {                                         /* :1285-1315 */
        double z1 = orc_param_conv(N, a1, b1, c1, d1, e1, f1, g1, h1, p1, i3, q3, u3);
        if (!(z1 > SEUIL_CV_SG)) {
          orc_ajout_queue(nt, N, d1, e1, f1, g1, h1, p1, d1out, e1out, f1out, g1out, h1out, p1out,
                          i3, q3, u3, i3out, q3out, u3out);
          reason = SOS_STOP_GEOM;
          break;
        }
      }
      for (int k = -N; k <= N; ++k) {                        /* label 506, :1323-1339 */
        V(a1, k) = V(d1, k); V(b1, k) = V(e1, k); V(c1, k) = V(f1, k);
        V(d1, k) = V(g1, k); V(e1, k) = V(h1, k); V(f1, k) = V(p1, k);
        if (k == 0) continue;
        for (int i = 0; i <= nt; ++i) { F(d1out, i, k) = F(g1out, i, k); F(e1out, i, k) = F(h1out, i, k); F(f1out, i, k) = F(p1out, i, k); }
      }
      for (int j = 1; j <= N; ++j) {                         /* :1343-1363 */
        V(i3, j) = V(i3, j) + F(i1, 0, j);
        V(q3, j) = V(q3, j) + F(q1, 0, j);
        V(u3, j) = V(u3, j) + F(u1, 0, j);
        V(i3, -j) = V(i3, -j) + F(i1, nt, -j);
        V(q3, -j) = V(q3, -j) + F(q1, nt, -j);
        V(u3, -j) = V(u3, -j) + F(u1, nt, -j);
        for (int i = 0; i <= nt; ++i) {
          F(i3out, i, j) = F(i3out, i, j) + F(i1, i, j);
          F(q3out, i, j) = F(q3out, i, j) + F(q1, i, j);
          F(u3out, i, j) = F(u3out, i, j) + F(u1, i, j);
          F(i3out, i, -j) = F(i3out, i, -j) + F(i1, i, -j);
          F(q3out, i, -j) = F(q3out, i, -j) + F(q1, i, -j);
          F(u3out, i, -j) = F(u3out, i, -j) + F(u1, i, -j);
        }
      }
      double z1 = orc_arret_diffus_1(N, nt, i1, q1, u1);     /* :1368 */
      if (!(z1 > SEUIL_VALDIF)) { reason = SOS_STOP_LOWVAL; break; }
      z1 = orc_arret_diffus_2(N, nt, i1, q1, u1, i3, q3, u3); /* :1387 */
      if (!(z1 > SEUIL_SUMDIF)) { reason = SOS_STOP_RATIO; break; }
      if (!(ig < igmax)) { reason = SOS_STOP_IGMAX; break; } /* :1406 */
    }
    if (n_scatter) n_scatter[is] = ig;
    if (stop_reason) stop_reason[is] = reason;

    if (imat_surf == 1) {                                    /* :1421-1439 */
      for (int j = 1; j <= N; ++j) {
        V(i3, j) = V(i3, j) - rii[j];
        V(q3, j) = V(q3, j) - rqq[j];
        V(u3, j) = V(u3, j) - ruu[j];
        for (int i = 0; i <= nt; ++i) {
          F(i3out, i, j) = F(i3out, i, j) - RO_(riiout, i, j);
          F(q3out, i, j) = F(q3out, i, j) - RO_(rqqout, i, j);
          F(u3out, i, j) = F(u3out, i, j) - RO_(ruuout, i, j);
        }
      }
    }
    if (is == 0) {                                           /* :1447-1456 */
      double em = 0.0, ep = 0.0;
      for (int j = 1; j <= N; ++j) {
        em = em + V(rmu, j) * V(ga, j) * V(i3, -j);
        ep = ep + V(rmu, j) * V(ga, j) * V(i3, j);
      }
      em = -em * 2 / tab;
      ep = -ep * 2 / tab;
      if (emoins) *emoins = em;
      if (eplus) *eplus = ep;
    }
    double coef = 2.0;                                       /* :1460-1473 */
    if (is == 0) coef = 1.0;
    for (int j = -N; j <= N; ++j) {
      if (j == 0) continue;
      V(i4, j) = V(i4, j) + coef * V(i3, j);
      V(q4, j) = V(q4, j) + coef * V(q3, j);
      V(u4, j) = V(u4, j) + coef * V(u3, j);
      V(i5, j) = V(i5, j) + coef * V(i3, j) * sign;
      V(q5, j) = V(q5, j) + coef * V(q3, j) * sign;
      V(u5, j) = V(u5, j) + coef * V(u3, j) * sign;
    }
    if (zout == -1) {                                        /* :1484-1534 */
      for (int k = -N; k <= -1; ++k) { V(i3z, k) = F(i3out, nt, k); V(q3z, k) = F(q3out, nt, k); V(u3z, k) = F(u3out, nt, k); }
      for (int k = 1; k <= N; ++k) { V(i3z, k) = F(i3out, 0, k); V(q3z, k) = F(q3out, 0, k); V(u3z, k) = F(u3out, 0, k); }
      V(i3z, 0) = 0.0; V(q3z, 0) = 0.0; V(u3z, 0) = 0.0; /* index 0 is unused (undefined in the reference) */
    } else {
      int j = 1;
      while (zout < zprof[j]) j = j + 1;
      double zz = (zout - zprof[j - 1]) / (zprof[j] - zprof[j - 1]);
      for (int k = -N; k <= N; ++k) {
        V(i3z, k) = (1 - zz) * F(i3out, j - 1, k) + zz * F(i3out, j, k);
        V(q3z, k) = (1 - zz) * F(q3out, j - 1, k) + zz * F(q3out, j, k);
        V(u3z, k) = (1 - zz) * F(u3out, j - 1, k) + zz * F(u3out, j, k);
      }
      V(i3z, 0) = 0.0; V(q3z, 0) = 0.0; V(u3z, 0) = 0.0;
    }
    if (rec) {                                               /* :1571-1575 */
      double *r = rec + (size_t)nrec * 3 * W;
      memcpy(r, q3z, W * sizeof(double));
      memcpy(r + W, u3z, W * sizeof(double));
      memcpy(r + 2 * W, i3z, W * sizeof(double));
    }
    nrec++;
    double z1 = orc_arret_fourier(N, i3, q3, u3, i4, q4, u4, i5, q5, u5); /* :1580 */
    if (!(z1 > SEUIL_SF)) break;                             /* :1585-1589 */
  }
  if (n_fourier) *n_fourier = nrec;

  free(f11); free(f12); free(f33); free(ch);
  free(i4); free(q4); free(u4); free(i5); free(q5); free(u5); free(i3); free(q3); free(u3);
  free(a1); free(b1); free(c1); free(d1); free(e1); free(f1); free(g1); free(h1); free(p1);
  free(i3z); free(q3z); free(u3z); free(xpl); free(xrl); free(xtl);
  free(xr); free(rii); free(rqq); free(ruu); free(riiout); free(rqqout); free(ruuout);
  free(bp); free(gr); free(gt); free(arr); free(art); free(att);
  free(i1); free(q1); free(u1); free(i2); free(q2); free(u2); free(i1f); free(q1f); free(u1f);
  free(i3out); free(q3out); free(u3out); free(d1out); free(e1out); free(f1out);
  free(g1out); free(h1out); free(p1out); free(rs);
  return 0;
#undef NEWF
#undef NEWV
#undef RO_
#undef RM
}

/* ------------------------------------------------------------------------- */
/* SOS, SOS.F:340-697 (in memory)                                             */
int orc_sos(int nt, double zout, int igmax, int ipolar, double ron, double ind_surf, double rho,
            int imat_surf, int ifresnel, const float *surf, int n0, double piz, double piztr, double a,
            double *rmu, const double *ga, double tetas, int os_nb, int nbmu,
            double *alpha, double *beta, double *gamma, double *zeta,
            const double *zprof, const double *h_in, const double *pcaer, const double *pcmol,
            int want_trans,
            double *rec, int *n_fourier, int *n_scatter, int *stop_reason,
            double *ttot_tronc, double *ttot_vrai, double *tauout, double *tdifmus, double *tdifmug,
            double *emoins, double *eplus, double *h_tr, double *xdel_tr, double *ydel_tr)
{
  const int N = nbmu;
  const int L = nt + 1;
  double *h = (double *)malloc(L * sizeof(double));
  double *htr = (double *)calloc(L, sizeof(double));
  double *xdel = (double *)malloc(L * sizeof(double));
  double *ydel = (double *)malloc(L * sizeof(double));
  memcpy(h, h_in, L * sizeof(double));
  memcpy(xdel, pcaer, L * sizeof(double));
  memcpy(ydel, pcmol, L * sizeof(double));
  int lta = 1;
  double va = 0.0, vatr = 0.0, vr = 0.0, vg;
  *ttot_vrai = h[nt];                                        /* :518 */
  htr[0] = h[0];
  if (a != 0.0) {                                            /* :525-537 */
    for (int i = 1; i <= nt; ++i) {
      va = xdel[i] * (h[i] - h[i - 1]);
      vatr = va * (1 - piz * 0.5 * a);
      vr = ydel[i] * (h[i] - h[i - 1]);
      vg = (1 - xdel[i] - ydel[i]) * (h[i] - h[i - 1]);
      htr[i] = (vatr + vr + vg) + htr[i - 1];
      xdel[i] = vatr / (vatr + vr + vg);
      ydel[i] = vr / (vatr + vr + vg);
    }
  }
  for (int i = 0; i <= nt; ++i) {                            /* :539-543 */
    if (a != 0.0) h[i] = htr[i];
    xdel[i] = xdel[i] * piztr;
    if (xdel[i] != 0.0) lta = 0;
  }
  int iborm = os_nb;                                         /* :549-550 */
  if (lta) iborm = 2;
  int ier = orc_sos_os(N, rmu, ga, os_nb, nt, n0, tetas, rho, imat_surf, ifresnel, ind_surf,
                       h, xdel, ydel, zprof, ron, alpha, beta, gamma, zeta, zout, igmax, iborm, ipolar,
                       surf, rec, n_fourier, n_scatter, stop_reason, emoins, eplus);
  if (ier == 0) {
    if (zout == -1) *tauout = h[0];                          /* :567-583 */
    else {
      int j = 1;
      while (zout < zprof[j]) j = j + 1;
      double zz = (zout - zprof[j - 1]) / (zprof[j] - zprof[j - 1]);
      *tauout = (1 - zz) * h[j - 1] + zz * h[j];
    }
    *ttot_tronc = h[nt];                                     /* :586 */
    if (want_trans) {                                        /* :605-637 */
      double eplus_solnoir;
      ier = orc_sos_os(N, rmu, ga, os_nb, nt, n0, tetas, 0.0, 0, 0, ind_surf, h, xdel, ydel, zprof, ron,
                       alpha, beta, gamma, zeta, -1.0, igmax, 0, ipolar, NULL,
                       NULL, NULL, NULL, NULL, tdifmus, &eplus_solnoir);
      for (int j = 1; j <= N && ier == 0; ++j) {
        double tmp;
        ier = orc_sos_os(N, rmu, ga, os_nb, nt, j, tetas, 0.0, 0, 0, ind_surf, h, xdel, ydel, zprof, ron,
                         alpha, beta, gamma, zeta, -1.0, igmax, 0, ipolar, NULL,
                         NULL, NULL, NULL, NULL, &tmp, &eplus_solnoir);
        V(tdifmug, j) = tmp;
      }
    }
  }
  if (h_tr) memcpy(h_tr, h, L * sizeof(double));
  if (xdel_tr) memcpy(xdel_tr, xdel, L * sizeof(double));
  if (ydel_tr) memcpy(ydel_tr, ydel, L * sizeof(double));
  free(h); free(htr); free(xdel); free(ydel);
  return ier;
}

/* ------------------------------------------------------------------------- */
/* SOS_AGGREGATE, SOS_AGGREGATE.F:289-488 (in memory)                          */
int orc_aggregate(int nbmu, double aik, const double *tmp, int ntmp,
                  double *res, int nres, int have_res,
                  double ttot_tronc_tmp, double ttot_vrai_tmp, double tauout_tmp,
                  double tdifmus_tmp, const double *tdifmug_tmp, double emoins_tmp, double eplus_tmp,
                  double *ttot_tronc, double *ttot_vrai, double *tauout,
                  double *tdifmus, double *tdifmug, double *emoins, double *eplus)
{
  const int N = nbmu, W = 2 * nbmu + 1;
  const int RW = 3 * W;
  /* record loop :355-422.  ios16/ios17 follow the reference's IOSTAT state machine */
  int ios16 = 0, ios17 = have_res ? 0 : -1;
  int r16 = 0, r17 = 0, nout = 0;
  double *cur_tmp = (double *)calloc(RW, sizeof(double));
  double *cur_res = (double *)calloc(RW, sizeof(double));
  double *out = (double *)calloc((size_t)RW * (size_t)((ntmp > nres ? ntmp : nres) + 2), sizeof(double));
  for (;;) {
    if (ios16 == 0) {
      if (r16 < ntmp) { memcpy(cur_tmp, tmp + (size_t)r16 * RW, RW * sizeof(double)); r16++; }
      else ios16 = -1;           /* EOF: list items keep their (zeroed) values */
    }
    if (ios17 == 0) {
      if (r17 < nres) { memcpy(cur_res, res + (size_t)r17 * RW, RW * sizeof(double)); r17++; }
      else ios17 = -1;
    } else {
      if (ios16 != 0) break;     /* :391 both files finished */
    }
    double *o = out + (size_t)nout * RW;
    for (int x = 0; x < RW; ++x) {                           /* :397-413 */
      o[x] = cur_res[x] + aik * cur_tmp[x];
      cur_tmp[x] = 0.0;
      cur_res[x] = 0.0;
    }
    nout++;
  }
  memcpy(res, out, (size_t)nout * RW * sizeof(double));
  free(cur_tmp); free(cur_res); free(out);

  *tdifmus = *tdifmus + aik * tdifmus_tmp;                   /* :452-459 */
  *emoins = *emoins + aik * emoins_tmp;
  *eplus = *eplus + aik * eplus_tmp;
  if (tdifmug && tdifmug_tmp)
    for (int j = -N; j <= N; ++j) V(tdifmug, j) = V(tdifmug, j) + aik * V(tdifmug_tmp, j);

  double trans;                                              /* :467-488 */
  if (*ttot_tronc != 0) trans = aik * exp(-ttot_tronc_tmp) + exp(-*ttot_tronc);
  else trans = aik * exp(-ttot_tronc_tmp);
  *ttot_tronc = -log(trans);
  if (*ttot_vrai != 0) trans = aik * exp(-ttot_vrai_tmp) + exp(-*ttot_vrai);
  else trans = aik * exp(-ttot_vrai_tmp);
  *ttot_vrai = -log(trans);
  if (*tauout != 0) trans = aik * exp(-tauout_tmp) + exp(-*tauout);
  else trans = aik * exp(-tauout_tmp);
  *tauout = -log(trans);
  return nout;
}

/* ------------------------------------------------------------------------- */
/* SOS_GLITTE, SOS_TRPHI.F:1278-1317 */
static double orc_glitte(double sig, double c0, double c1, double phi)
{
  double x1 = sqrt(1 - c1 * c1) - cos(phi) * sqrt(1 - c0 * c0);
  double x2 = sqrt(1 - c0 * c0) * sin(phi);
  double x3 = c0 + c1;
  double c0n = (x3 / (sqrt(x1 * x1 + x2 * x2 + x3 * x3)));
  double xxx = (-(1 - c0n * c0n) / (sig * (c0n * c0n)));
  if (xxx < -100) return 0.0;
  double pp = (1 / sig) * exp(xxx);
  double c2 = c0n * c0n;
  return pp / (4 * c1 * (c2 * c2));
}

/* SOS_ANGLE, SOS_TRPHI.F:1347-1375 */
static void orc_angle(double c0, double c1, double phi, double *coskip, double *cosdif)
{
  double s = 1.0;
  if (sin(phi) > 0.0) s = -1.0;
  *cosdif = -c0 * c1 + sqrt(1 - c0 * c0) * sqrt(1 - c1 * c1) * cos(phi);
  double z = s * (sqrt(1 - (*cosdif) * (*cosdif))) * (sqrt(1 - c1 * c1));
  *coskip = 0.0;
  if (fabs(z) > SEUIL_Z) *coskip = (c1 * (*cosdif) + c0) / z;
}

/* SOS_REFLEX, SOS_TRPHI.F:1433-1472 */
static void orc_reflex(double cosdif, double ind, double *r11, double *r12, double *r33)
{
  double ind2 = ind * ind;
  double cosw = sqrt(.5 * (1 - cosdif));
  double v = .5 * (1 + cosdif);
  double x = sqrt(ind2 - v);
  double rl = (ind2 * cosw - x) / (ind2 * cosw + x);
  double rr = (cosw - x) / (cosw + x);
  *r11 = (rl * rl + rr * rr) / 2.0;
  *r12 = (rl * rl - rr * rr) / 2.0;
  *r33 = rr * rl;
}

/* SOS_MATRIC, SOS_TRPHI.F:1505-1541 */
static void orc_matric(double coskip, double r11, double r12, double *m11, double *m21, double *m31)
{
  double x = 1.0 - fabs(coskip);
  double c2 = 1.0, s2 = 0.0;
  if (x >= SEUIL_X) {
    c2 = 2.0 * coskip * coskip - 1.0;
    s2 = 2.0 * coskip * sqrt(1.0 - coskip * coskip);
  }
  if (coskip == 0.0) r12 = 0.0;
  *m11 = r11;
  *m21 = c2 * r12;
  *m31 = s2 * r12;
}

/* SOS_POLAR, SOS_TRPHI.F:1843-1907 */
void orc_polar(double xi, double xq, double xu, double *xan, double *tpol, double *lpol)
{
  const double pi = orc_pi();
  if (xq != 0.0) {
    double xt = xu / xq;
    if (xq > 0.0) *xan = 90.0 * atan(xt) / pi;
    else if (xu > 0.0) *xan = 90.0 + 90.0 * atan(xt) / pi;
    else *xan = -90.0 + 90.0 * atan(xt) / pi;
  } else {
    if (xu > 0.0) *xan = 45.0;
    else if (xu < 0) *xan = -45.0;
    else *xan = VALEUR_INDEF;
  }
  *lpol = sqrt(xq * xq + xu * xu);
  if (xi != 0.0) *tpol = 100.0 * (*lpol) / xi;
  else *tpol = VALEUR_INDEF;
}

/* SOS_TRPHI, SOS_TRPHI.F:749-1243 (glitter and flat-sea direct terms) */
int orc_trphi(const double *rec, int nrec, int nbmu, const double *rmu, double tau, double tauout,
              double phi, int igli, int n0, double wind, double ind_surf, int ifresnel, int ipolar,
              double *xit, double *xqt, double *xut, double *angdiff)
{
  const int N = nbmu, W = 2 * nbmu + 1;
  const double pi = orc_pi();
  if (nrec < 1) return -1;
  double c0 = V(rmu, n0);                                    /* :882 */
  for (int j = -N; j <= N; ++j) {
    double cosdif = -c0 * V(rmu, j) + sin(acos(c0)) * sin(acos(V(rmu, j))) * cos(phi);
    V(angdiff, j) = acos(cosdif) * 180.0 / pi;
  }
  {
    const double *q3 = rec, *u3 = rec + W, *i3 = rec + 2 * W; /* :908-918 */
    for (int j = -N; j <= N; ++j) {
      if (j == 0) { V(xqt, j) = 0.0; V(xut, j) = 0.0; V(xit, j) = 0.0; continue; }
      V(xqt, j) = V(q3, j); V(xut, j) = V(u3, j); V(xit, j) = V(i3, j);
    }
  }
  for (int is = 1; is < nrec; ++is) {                        /* :921-940 */
    const double *q3 = rec + (size_t)is * 3 * W, *u3 = q3 + W, *i3 = q3 + 2 * W;
    double xphi = is * phi;
    for (int j = -N; j <= N; ++j) {
      if (j == 0) continue;
      V(xqt, j) = V(xqt, j) + 2.0 * V(q3, j) * cos(xphi);
      V(xut, j) = V(xut, j) + 2.0 * V(u3, j) * sin(xphi);
      V(xit, j) = V(xit, j) + 2.0 * V(i3, j) * cos(xphi);
    }
  }
  if (igli == 1) {                                           /* :946-1001 */
    c0 = V(rmu, n0);
    double at0 = exp(-tau / c0);
    double sigma2 = (double)0.003f + (double)0.00512f * wind; /* .003 + .00512*WIND :963 */
    for (int j = 1; j <= N; ++j) {
      double atj = at0 * exp(-(tau - tauout) / V(rmu, j));
      double c1 = V(rmu, j);
      double p = orc_glitte(sigma2, c0, c1, phi);
      double coskip, cosdif, r11, r12, r33, m11, m21, m31;
      orc_angle(c0, c1, phi, &coskip, &cosdif);
      orc_reflex(cosdif, ind_surf, &r11, &r12, &r33);
      orc_matric(coskip, r11, r12, &m11, &m21, &m31);
      V(xit, j) = V(xit, j) + m11 * atj * p;
      if (ipolar == 1) {
        V(xqt, j) = V(xqt, j) + m21 * atj * p;
        V(xut, j) = V(xut, j) + m31 * atj * p;
      }
    }
  }
  if (ifresnel == 1) {                                       /* :1008-1039 */
    if ((cos(phi) == 1.0) && (n0 > 0)) {
      c0 = V(rmu, n0);
      double at0 = exp(-tau / c0);
      double atj = at0 * exp(-(tau - tauout) / c0);
      double cosdif = 1.0 - 2.0 * c0 * c0;
      double r11, r12, r33;
      orc_reflex(cosdif, ind_surf, &r11, &r12, &r33);
      double coef_sun = pi / SOLAR_DISC;
      V(xit, n0) = V(xit, n0) + r11 * coef_sun * atj;
      if (ipolar == 1) V(xqt, n0) = V(xqt, n0) + r12 * coef_sun * atj;
    }
  }
  for (int j = -N; j <= N; ++j) {                            /* :1212-1218 */
    if (V(xit, j) <= 1.e-99) V(xit, j) = 0.0;
    if (fabs(V(xqt, j)) < THRESHOLD_QU) V(xqt, j) = 0.0;
    if (fabs(V(xut, j)) < THRESHOLD_QU) V(xut, j) = 0.0;
  }
  return 0;
}

/* SOS_TRPHI_OPTION, SOS_TRPHI.F:285-636 */
int orc_trphi_option(const double *rec, int nrec, int nbmu, const double *rmu, double tau, double tauout,
                     int igli, int n0, double wind, double ind_surf, int ifresnel,
                     int itrphi, double phios, int pas_phi, int ipolar,
                     double *phi_fin, double *theta_fin, double *up, double *down, int nphi_cap)
{
  const int N = nbmu, W = 2 * nbmu + 1;
  const double pi = orc_pi();
  double *xit = (double *)calloc(W, sizeof(double));
  double *xqt = (double *)calloc(W, sizeof(double));
  double *xut = (double *)calloc(W, sizeof(double));
  double *ang = (double *)calloc(W, sizeof(double));
  int nphi = -1;
#define T(tab, t, ip, jj) (tab)[((size_t)(t) * nphi_cap + (ip)) * N + (jj)]
  if (itrphi == 1 && nphi_cap >= 2) {
    for (int pass = 0; pass < 2; ++pass) {
      double phi = (pass == 0) ? pi + phios * pi / 180.0 : phios * pi / 180.0; /* :435, :494 */
      if (orc_trphi(rec, nrec, N, rmu, tau, tauout, phi, igli, n0, wind, ind_surf, ifresnel, ipolar,
                    xit, xqt, xut, ang) != 0) goto done;
      if (pass == 0) phi_fin[0] = phios + 180; else phi_fin[0] = phios;  /* :445, :504 (slot 0 both times) */
      for (int j = 1; j <= N; ++j) {
        double teta = acos(V(rmu, j)) * 180.0 / pi;
        int jj = j - 1;
        double xan, tpol, lpol;
        theta_fin[jj] = (pass == 0) ? -teta : teta;
        orc_polar(V(xit, j), V(xqt, j), V(xut, j), &xan, &tpol, &lpol);
        T(up, 0, pass, jj) = V(ang, j); T(up, 1, pass, jj) = V(xit, j); T(up, 2, pass, jj) = V(xqt, j);
        T(up, 3, pass, jj) = V(xut, j); T(up, 4, pass, jj) = xan; T(up, 5, pass, jj) = tpol; T(up, 6, pass, jj) = lpol;
        orc_polar(V(xit, -j), V(xqt, -j), V(xut, -j), &xan, &tpol, &lpol);
        T(down, 0, pass, jj) = V(ang, -j); T(down, 1, pass, jj) = V(xit, -j); T(down, 2, pass, jj) = V(xqt, -j);
        T(down, 3, pass, jj) = V(xut, -j); T(down, 4, pass, jj) = xan; T(down, 5, pass, jj) = tpol; T(down, 6, pass, jj) = lpol;
      }
    }
    nphi = 2;
  } else if (itrphi == 2) {
    int ip = 0;
    for (int iphi = 0; iphi <= 360; iphi += pas_phi) {       /* :558-613 */
      if (ip >= nphi_cap) break;
      double phi = pi * iphi / 180.0;
      if (orc_trphi(rec, nrec, N, rmu, tau, tauout, phi, igli, n0, wind, ind_surf, ifresnel, ipolar,
                    xit, xqt, xut, ang) != 0) goto done;
      for (int j = 1; j <= N; ++j) {
        double teta = acos(V(rmu, j)) * 180.0 / pi;
        int jj = j - 1;
        double xan, tpol, lpol;
        phi_fin[ip] = (double)iphi;
        theta_fin[jj] = teta;
        orc_polar(V(xit, j), V(xqt, j), V(xut, j), &xan, &tpol, &lpol);
        T(up, 0, ip, jj) = V(ang, j); T(up, 1, ip, jj) = V(xit, j); T(up, 2, ip, jj) = V(xqt, j);
        T(up, 3, ip, jj) = V(xut, j); T(up, 4, ip, jj) = xan; T(up, 5, ip, jj) = tpol; T(up, 6, ip, jj) = lpol;
        orc_polar(V(xit, -j), V(xqt, -j), V(xut, -j), &xan, &tpol, &lpol);
        T(down, 0, ip, jj) = V(ang, -j); T(down, 1, ip, jj) = V(xit, -j); T(down, 2, ip, jj) = V(xqt, -j);
        T(down, 3, ip, jj) = V(xut, -j); T(down, 4, ip, jj) = xan; T(down, 5, ip, jj) = tpol; T(down, 6, ip, jj) = lpol;
      }
      ip++;
    }
    nphi = ip;
  }
done:
  free(xit); free(xqt); free(xut); free(ang);
  return nphi;
#undef T
}
