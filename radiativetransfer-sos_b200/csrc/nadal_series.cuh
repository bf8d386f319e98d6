// nadal_series.cuh -- Fourier series of Nadal's BPDF divided by Fresnel's F21 (-SURF.Type 6), as the pieces
// k_glitter (glitter_kernel.cu, gmodel 4) gives to single threads:
//   F(theta1, theta2, phi)              SOS_CALC_F21_NADAL_SUR_FRESNEL   (SOS_SURFACE_BPDF.F:1129-1223)
//   E(s) = (sum_i U(i) cos(s phi_i)) Q / pi, i ascending                 SOS_F21SF_NADAL (:856-861)
//   T1_s(i) = E(0) + 2 sum_{s'<=s} E(s') cos(s' phi_i), running in s     (:868-879)
//   the cut of the series from B1(s) = max_i |(T1_s(i) - F_i) / F_i|     (:888-900)
// Every function is the reference's own sequence of operations (compiled without FMA contraction); only exp / cos differ from
// glibc's on the device.  __host__ __device__ so that tests/surface_host.cpp can step the same functions against the reference
// library on a machine without a GPU; the library only ever runs them inside k_glitter.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define NAD_HD __host__ __device__ inline
#else
#define NAD_HD static inline
#endif

#define NAD_PH_NU 1024                         // CTE_PH_NU (SOS.h:319)
#define NAD_SEUIL ((double)0.001f)             // CTE_SEUIL_SF_NADAL (SOS.h:332), REAL*4 literal compared with a double
#define NAD_PI 3.141592653589793               // 4.D+00*DATAN(1.D+00) (:1166): glibc returns this double

// SOS_CALC_F21_NADAL_SUR_FRESNEL.  (S1 and S2 are arguments of the reference routine that it does not use: the sines are
// recomputed from the cosines, :1172.)
NAD_HD double nadal_f(double ind, double alpha, double beta, double c1, double c2, double phi)
{
  const double coef = 4. * c1 * NAD_PI;
  const double cosdif = -c1 * c2 + sqrt(1. - c1 * c1) * sqrt(1. - c2 * c2) * cos(phi);
  const double cosw = sqrt(.5 * (1 - cosdif));
  const double v = .5 * (1 + cosdif);
  const double ind2 = ind * ind;
  const double x = sqrt(ind2 - v);
  const double rl = (ind2 * cosw - x) / (ind2 * cosw + x);
  const double rr = (cosw - x) / (cosw + x);
  const double f21fresnel = 0.5 * (rr * rr - rl * rl);
  double f21nadal = -beta * f21fresnel / (c1 + c2);
  f21nadal = alpha * (1. - exp(f21nadal));
  double f;
  if (f21fresnel < 1.0e-10) f = alpha * beta / (c1 + c2);           // expansion near F21 of Fresnel = 0 (:1209-1210)
  else f = f21nadal / f21fresnel;
  return f * coef * c2 * c1;                                         // reflectance -> normalised radiance, adapted to the OS code (:1218)
}

// E(IS) from the samples U(0:NU) (:856-861): one thread per order, the reference's order of additions
NAD_HD double nadal_coef(const double *U, int is, double q, double pi)
{
  double y = 0.;
  for (int i = 0; i <= NAD_PH_NU; ++i) {
    const double phi = i * q;
    y = y + U[i] * cos(is * phi);
  }
  return y * q / pi;
}

// one step of the recombination at azimuth phi: T1 after adding order IS2 (:875-877)
NAD_HD double nadal_recomb_step(double t1, double e, int is2, double phi)
{
  return t1 + 2. * e * cos(is2 * phi);
}

// the cut (:888-900): IL = first order with B1 <= threshold; else the order before the first one where B1 stops decreasing;
// else OS_NB.  B1[s] for all s in 0..nb (the reference computes them one by one and stops at the cut).
NAD_HD int nadal_cut(const double *B1, int nb)
{
  double b1_prec = 1.e300;
  for (int is = 0; is <= nb; ++is) {
    const double b1 = B1[is];
    if (!(b1 > NAD_SEUIL)) return is;
    if (b1 < b1_prec) { b1_prec = b1; continue; }
    return is - 1;
  }
  return nb;
}
