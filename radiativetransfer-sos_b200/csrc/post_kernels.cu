// post_kernels.cu -- azimuth synthesis of the Fourier series onto the view directions
// (SOS_TRPHI / SOS_TRPHI_OPTION, SOS_TRPHI.F:285-636, 749-1243, helpers :1278-1541, 1843-1907).
// Compiled with -fmad=false: same operation order as the reference.
// Bandwidth-type kernel: one thread per (group, azimuth, direction); each thread walks the Fourier
// records of its group in ascending order (the reference's summation order).
#include "sosgpu_internal.h"
#include "post_kernels.h"
#include <math.h>

#define SEUIL_Z      ((double)0.0001f)    // CTE_SEUIL_Z  SOS.h:407 (REAL*4 literal)
#define SEUIL_X      ((double)0.00001f)   // CTE_SEUIL_X  SOS.h:413
#define THRESHOLD_QU (1.0e-15)            // CTE_THRESHOLD_Q_U_NULL SOS.h:418
#define SOLAR_DISC   (6.8e-05)            // CTE_SOLAR_DISC_SOLID_ANGLE SOS.h:426
#define VALEUR_INDEF (-999.0)             // INCTE_VALEUR_INDEF SOS_TRPHI.F:134
#define SOSGPU_NB_MAX_DEV 200             // CTE_OS_NB_MAX SOS.h:480

// SOS_GLITTE, SOS_TRPHI.F:1278-1317
__device__ static double glitte(double sig, double c0, double c1, double phi)
{
  const double x1 = sqrt(1 - c1 * c1) - cos(phi) * sqrt(1 - c0 * c0);
  const double x2 = sqrt(1 - c0 * c0) * sin(phi);
  const double x3 = c0 + c1;
  const double c0n = (x3 / (sqrt(x1 * x1 + x2 * x2 + x3 * x3)));
  const double xxx = (-(1 - c0n * c0n) / (sig * (c0n * c0n)));
  if (xxx < -100) return 0.0;
  const double pp = (1 / sig) * exp(xxx);
  const double c2 = c0n * c0n;
  return pp / (4 * c1 * (c2 * c2));
}
// SOS_ANGLE, SOS_TRPHI.F:1347-1375
__device__ static void angle(double c0, double c1, double phi, double &coskip, double &cosdif)
{
  double s = 1.0;
  if (sin(phi) > 0.0) s = -1.0;
  cosdif = -c0 * c1 + sqrt(1 - c0 * c0) * sqrt(1 - c1 * c1) * cos(phi);
  const double z = s * (sqrt(1 - cosdif * cosdif)) * (sqrt(1 - c1 * c1));
  coskip = 0.0;
  if (fabs(z) > SEUIL_Z) coskip = (c1 * cosdif + c0) / z;
}
// SOS_REFLEX, SOS_TRPHI.F:1433-1472
__device__ static void reflex(double cosdif, double ind, double &r11, double &r12, double &r33)
{
  const double ind2 = ind * ind;
  const double cosw = sqrt(.5 * (1 - cosdif));
  const double v = .5 * (1 + cosdif);
  const double x = sqrt(ind2 - v);
  const double rl = (ind2 * cosw - x) / (ind2 * cosw + x);
  const double rr = (cosw - x) / (cosw + x);
  r11 = (rl * rl + rr * rr) / 2.0;
  r12 = (rl * rl - rr * rr) / 2.0;
  r33 = rr * rl;
}
// SOS_MATRIC, SOS_TRPHI.F:1505-1541
__device__ static void matric(double coskip, double r11, double r12, double &m11, double &m21, double &m31)
{
  const double x = 1.0 - fabs(coskip);
  double c2 = 1.0, s2 = 0.0;
  if (x >= SEUIL_X) {
    c2 = 2.0 * coskip * coskip - 1.0;
    s2 = 2.0 * coskip * sqrt(1.0 - coskip * coskip);
  }
  if (coskip == 0.0) r12 = 0.0;
  m11 = r11;
  m21 = c2 * r12;
  m31 = s2 * r12;
}
// SOS_CALC_F_ROUJEAN, SOS_ROUJEAN.F:891-1022 (CTE_TETAS/TETAV_LIM_ROUJEAN = 60 degrees, CTE_SEUIL_NUM = 1.D-10: SOS.h:347,355,361)
__device__ static double calc_f_roujean(double k0, double k1, double k2, double c1, double s1, double c2, double s2, double phi)
{
  const double pi = 4.0 * atan(1.0);
  double xphi = phi;
  if (xphi < 0.0) xphi = -xphi;
  if (xphi > pi) xphi = 2. * pi - xphi;
  double xc1 = c1, xs1 = s1, xc2 = c2, xs2 = s2;
  if (acos(c1) * 180. / pi > 60) { xc1 = cos(60 * pi / 180.0); xs1 = sin(60 * pi / 180.0); }
  if (acos(c2) * 180. / pi > 60) { xc2 = cos(60 * pi / 180.0); xs2 = sin(60 * pi / 180.0); }
  const double cosphi = cos(xphi);
  const double tants = xs1 / xc1, tantv = xs2 / xc2;
  double f1 = 0.5 * ((pi - xphi) * cosphi + sin(xphi)) * tants * tantv;
  f1 = f1 - tants - tantv;
  f1 = f1 - sqrt(tants * tants + tantv * tantv - 2. * tantv * tants * cosphi);
  f1 = f1 / pi;
  double coszeta = xc1 * xc2 + xs1 * xs2 * cosphi;
  if (fabs(fabs(coszeta) - 1.0) <= 1.e-10) {
    if ((coszeta >= (1. - 1.e-10)) && (coszeta <= (1. + 1.e-10))) coszeta = 1.0;
    else coszeta = -1.0;
  }
  const double zeta = acos(coszeta);
  double f2 = 4. * ((pi / 2. - zeta) * coszeta + sin(zeta)) / (3. * pi * (xc1 + xc2));
  f2 = f2 - (1.0 / 3.0);
  double f = k0 + k1 * f1 + k2 * f2;
  f = f * c2 * c1;
  return f;
}
// SOS_CALCG_MAIGNAN, SOS_SURFACE_BPDF.F:1606-1641
__device__ static double calcg_maignan(double c1, double c2, double s12, double phi, double coef_c)
{
  const double cos_2i = c1 * c2 - s12 * cos(phi);
  double tan2_i = (1 - cos_2i) / (1 + cos_2i);
  if (tan2_i < 0.0) tan2_i = 0.0;
  return coef_c * exp(-sqrt(tan2_i)) / (1. / c1 + 1. / c2);
}
// SOS_POLAR, SOS_TRPHI.F:1843-1907
__device__ static void polar(double xi, double xq, double xu, double pi, double &xan, double &tpol, double &lpol)
{
  if (xq != 0.0) {
    const double xt = xu / xq;
    if (xq > 0.0) xan = 90.0 * atan(xt) / pi;
    else if (xu > 0.0) xan = 90.0 + 90.0 * atan(xt) / pi;
    else xan = -90.0 + 90.0 * atan(xt) / pi;
  } else {
    if (xu > 0.0) xan = 45.0;
    else if (xu < 0) xan = -45.0;
    else xan = VALEUR_INDEF;
  }
  lpol = sqrt(xq * xq + xu * xu);
  if (xi != 0.0) tpol = 100.0 * lpol / xi;
  else tpol = VALEUR_INDEF;
}

// grid: (nphi, ngroup); block: threads over the 2N directions.
// out: [ngroup][2 (up, down)][7][nphi][N]   tables SCA, I, Q, U, POL_ANG, POL_RATE, L_POL
__global__ void k_trphi(const TrphiGroup *groups, const double *phis, int nphi, int nout, TrphiParams prm, double *out)
{
  const TrphiGroup g = groups[blockIdx.y];
  const int ip = blockIdx.x;
  const int N = g.nbmu, W = g.wstride;
  const double phi = phis[ip];
  const double pi = prm.pi;
  if (nout < 0) nout = N;
  if (g.nrec <= 0) return;
  // cos(s*phi), sin(s*phi) depend on (order, azimuth) only: computed once per CTA, not once per direction
  __shared__ double s_cos[SOSGPU_NB_MAX_DEV + 1], s_sin[SOSGPU_NB_MAX_DEV + 1];
  for (int is = threadIdx.x; is < g.nrec && is <= SOSGPU_NB_MAX_DEV; is += blockDim.x) {
    const double xphi = is * phi;
    s_cos[is] = cos(xphi); s_sin[is] = sin(xphi);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * N; t += blockDim.x) {
    const int j = (t < N) ? (t + 1) : -(t - N + 1);
    const double rmuj = g.rmu[j + N];
    double c0 = g.rmu[g.n0 + N];                               // :882
    const double cosang = -c0 * rmuj + sin(acos(c0)) * sin(acos(rmuj)) * cos(phi);
    const double angdiff = acos(cosang) * 180.0 / pi;
    double xq = 0.0, xu = 0.0, xi = 0.0;
    if (g.nrec > 0) {                                          // :908-940
      xq = g.rec[j + N]; xu = g.rec[W + j + N]; xi = g.rec[2 * W + j + N];
      for (int is = 1; is < g.nrec; ++is) {
        const double *r = g.rec + (size_t)is * 3 * W;
        xq = xq + 2.0 * r[j + N] * s_cos[is];
        xu = xu + 2.0 * r[W + j + N] * s_sin[is];
        xi = xi + 2.0 * r[2 * W + j + N] * s_cos[is];
      }
    }
    if (prm.igli == 1 && j > 0) {                              // :946-1001
      const double at0 = exp(-g.tau / c0);
      const double sigma2 = (double)0.003f + (double)0.00512f * prm.wind;
      const double atj = at0 * exp(-(g.tau - g.tauout) / rmuj);
      const double c1 = rmuj;
      const double p = glitte(sigma2, c0, c1, phi);
      double coskip, cosdif, r11, r12, r33, m11, m21, m31;
      angle(c0, c1, phi, coskip, cosdif);
      reflex(cosdif, prm.ind_surf, r11, r12, r33);
      matric(coskip, r11, r12, m11, m21, m31);
      xi = xi + m11 * atj * p;
      if (prm.ipolar == 1) { xq = xq + m21 * atj * p; xu = xu + m31 * atj * p; }
    }
    if (prm.ifresnel == 1 && j == g.n0) {                      // :1008-1039
      if ((cos(phi) == 1.0) && (g.n0 > 0)) {
        const double at0 = exp(-g.tau / c0);
        const double atj = at0 * exp(-(g.tau - g.tauout) / c0);
        const double cosdif = 1.0 - 2.0 * c0 * c0;
        double r11, r12, r33;
        reflex(cosdif, prm.ind_surf, r11, r12, r33);
        const double coef_sun = pi / SOLAR_DISC;
        xi = xi + r11 * coef_sun * atj;
        if (prm.ipolar == 1) xq = xq + r12 * coef_sun * atj;
      }
    }
    if (prm.iroujean == 1 && j > 0) {                          // Roujean BRDF direct term (:1047-1072)
      const double at0 = exp(-g.tau / c0);
      const double s0 = sqrt(1.0 - c0 * c0);
      const double c1 = rmuj;
      const double atj = at0 * exp(-(g.tau - g.tauout) / c1);
      const double s1 = sqrt(1.0 - c1 * c1);
      const double phirj = pi - phi;
      const double f = calc_f_roujean(prm.k0, prm.k1, prm.k2, c0, s0, c1, s1, phirj);
      xi = xi + atj * f / c1;
    }
    if ((prm.irondeaux == 1 || prm.ibreon == 1 || prm.imaignan == 1) && j > 0) {   // vegetation / soil BPDF (:1080-1122)
      const double at0 = exp(-g.tau / c0);
      const double s0 = sqrt(1.0 - c0 * c0);
      const double c1 = rmuj;
      const double atj = at0 * exp(-(g.tau - g.tauout) / c1);
      double coskip, cosdif, r11, r12, r33, m11, m21, m31;
      angle(c0, c1, phi, coskip, cosdif);
      reflex(cosdif, prm.ind_surf, r11, r12, r33);
      matric(coskip, r11, r12, m11, m21, m31);
      double p = 0.0;
      if (prm.irondeaux == 1) p = 1. / (4. * (1 + c1 / c0));
      if (prm.ibreon == 1) p = 1. / (4. * c1);
      if (prm.imaignan == 1) {
        const double s1 = sqrt(1.0 - c1 * c1);
        const double s12 = s0 * s1;
        p = calcg_maignan(c0, c1, s12, phi, prm.coef_c_maignan);
        p = p / (4. * c1);
      }
      xi = xi + m11 * atj * p;
      if (prm.ipolar == 1) { xq = xq + m21 * atj * p; xu = xu + m31 * atj * p; }
    }
    if (prm.inadal == 1 && j > 0) {                            // Nadal BPDF (:1130-1178)
      const double at0 = exp(-g.tau / c0);
      const double c1 = rmuj;
      const double atj = at0 * exp(-(g.tau - g.tauout) / c1);
      double coskip, cosdif, r11, r12, r33, m11, m21, m31;
      angle(c0, c1, phi, coskip, cosdif);
      reflex(cosdif, prm.ind_surf, r11, r12, r33);
      matric(coskip, r11, r12, m11, m21, m31);
      const double f21fresnel = -r12;
      double f21nadal = -prm.beta_nadal * f21fresnel / (c0 + c1);
      f21nadal = prm.alpha_nadal * (1.0 - exp(f21nadal));
      double p;
      if (f21fresnel < 1.0e-10) p = prm.alpha_nadal * prm.beta_nadal / (c0 + c1);
      else p = f21nadal / f21fresnel;
      xi = xi + m11 * atj * p;
      if (prm.ipolar == 1) { xq = xq + m21 * atj * p; xu = xu + m31 * atj * p; }
    }
    if (xi <= 1.e-99) xi = 0.0;                                // :1212-1218
    if (fabs(xq) < THRESHOLD_QU) xq = 0.0;
    if (fabs(xu) < THRESHOLD_QU) xu = 0.0;
    double xan, tpol, lpol;
    polar(xi, xq, xu, pi, xan, tpol, lpol);
    const int ud = (j > 0) ? 0 : 1, jj = (j > 0 ? j : -j) - 1;
    double *o = out + (((size_t)blockIdx.y * 2 + ud) * 7) * nphi * nout + (size_t)ip * nout + jj;   // nout = row pitch (max N)
    const size_t ts = (size_t)nphi * nout;
    o[0] = angdiff; o[ts] = xi; o[2 * ts] = xq; o[3 * ts] = xu; o[4 * ts] = xan; o[5 * ts] = tpol; o[6 * ts] = lpol;
  }
}

extern "C" void sos_launch_trphi(const TrphiGroup *groups, int ngroup, const double *phis, int nphi,
                                 TrphiParams prm, double *out, cudaStream_t st)
{
  if (ngroup <= 0 || nphi <= 0) return;
  // single wavelength: the output pitch is that group's N (read from the descriptor on the host side by the caller)
  dim3 grid(nphi, ngroup);
  k_trphi<<<grid, 160, 0, st>>>(groups, phis, nphi, -1, prm, out);
}
extern "C" void sos_launch_trphi_stride(const TrphiGroup *groups, int ngroup, const double *phis, int nphi, int nout,
                                        TrphiParams prm, double *out, cudaStream_t st)
{
  if (ngroup <= 0 || nphi <= 0) return;
  dim3 grid(nphi, ngroup);
  k_trphi<<<grid, 160, 0, st>>>(groups, phis, nphi, nout, prm, out);
}

// RES = RES + AIK*TMP (SOS_AGGREGATE.F:401-403) for the file-level drop-in sos_aggregate_
__global__ void k_axpy(double *res, const double *tmp, double aik, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) res[i] = res[i] + aik * tmp[i];
}
extern "C" void sos_launch_axpy(double *res, const double *tmp, double aik, size_t n, cudaStream_t st)
{
  if (n == 0) return;
  k_axpy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(res, tmp, aik, n);
}


// =================================================================================================
// SOS_ROUJEAN (SOS_ROUJEAN.F:212-416) = SOS_FSF_ROUJEAN (:417-832) + SOS_MISE_FORMAT_RJ (:1102-1224): Fourier series of
// Roujean's BRDF F(theta1, theta2, phi) for every (incidence, reflection) pair, written straight into the surface-file
// record layout (only P11 is non-zero).  One CTA per pair (I1, I2):
//   U(i) = F(phi_i), i = 0..1024                                   one thread per sample (:513-521)
//   E(s) = (sum_i U(i) cos(s phi_i)) * Q / pi, i ascending         one thread per order s: the reference's own sequence (:556-563)
//   B1(s) = max_i |(T1_s(i) - F_i)/F_i|, T1_s(i) = E(0) + 2 sum_{s'<=s} E(s') cos(s' phi_i)   one thread per sample walks s
//           upwards with a running sum (the reference's order of additions, :571-586); the maximum is order independent
//   the series stops at the first s where B1 no longer decreases (:588-593): later coefficients stay zero.
// status[pair] = 1 when a negative BRDF value was met (the reference aborts with IER = -1, label 993).
#define RJ_NU 1024
__device__ __forceinline__ void atomic_max_pos(double *addr, double v)
{
  // non-negative doubles order like their bit patterns; NaN (0/0) has a larger pattern than any finite value, which makes the
  // comparison B1 > threshold / B1 < B1_PREC false, as in the reference
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}
__global__ void k_roujean(int N, const double *__restrict__ rmu, int os_nb, double k0, double k1, double k2,
                          float *__restrict__ surf, int *__restrict__ status)
{
  __shared__ double U[RJ_NU + 1];
  __shared__ double E[SOSGPU_NB_MAX_DEV + 1], B1[SOSGPU_NB_MAX_DEV + 1];
  __shared__ int s_last, s_neg;
  const int pair = blockIdx.x;
  const int I1 = pair / N + 1, I2 = pair % N + 1;
  const double pi = acos(-1.0);
  const double c1 = rmu[I1 + N], c2 = rmu[I2 + N];
  const double s1 = sqrt(1 - c1 * c1), s2 = sqrt(1 - c2 * c2);
  const double q = pi / RJ_NU;
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (tid == 0) s_neg = 0;
  for (int s = tid; s <= os_nb; s += nthr) B1[s] = 0.0;
  __syncthreads();
  for (int i = tid; i <= RJ_NU; i += nthr) {
    const double phios = q * i;
    const double f = calc_f_roujean(k0, k1, k2, c1, s1, c2, s2, pi - phios);
    U[i] = f;
    if (f < 0.0) s_neg = 1;
  }
  __syncthreads();
  for (int s = tid; s <= os_nb; s += nthr) {
    double y = 0.0;
    for (int i = 0; i <= RJ_NU; ++i) {
      const double phios = i * q;
      y = y + U[i] * cos(s * phios);
    }
    E[s] = y * q / pi;
  }
  __syncthreads();
  for (int i = tid; i <= RJ_NU; i += nthr) {
    const double phios = q * i;
    const double f = U[i];
    double t1 = E[0];
    atomic_max_pos(&B1[0], fabs((t1 - f) / f));
    for (int s = 1; s <= os_nb; ++s) {
      t1 = t1 + 2. * E[s] * cos(s * phios);
      atomic_max_pos(&B1[s], fabs((t1 - f) / f));
    }
  }
  __syncthreads();
  if (tid == 0) {                                             // :588-596: stop when the recombination error stops decreasing
    double b1_prec = 1.e300;
    int last = os_nb;                                         // last order whose coefficient was computed by the reference
    for (int s = 0; s <= os_nb; ++s) {
      if (B1[s] < b1_prec) { b1_prec = B1[s]; continue; }
      last = s;
      break;
    }
    s_last = last;
    if (status) status[pair] = s_neg;
  }
  __syncthreads();
  const int last = s_last;
  const size_t nn = (size_t)N * N;
  for (int s = tid; s <= os_nb; s += nthr) {
    float *rec = surf + (size_t)s * 9 * nn;                   // record s, matrix P11, element (I, J) column-major
    rec[(size_t)(I2 - 1) * N + (I1 - 1)] = (s <= last) ? (float)E[s] : 0.f;
  }
}

extern "C" void sos_launch_roujean(int N, const double *rmu, int os_nb, double k0, double k1, double k2, float *surf, int *status,
                                   cudaStream_t st)
{
  k_roujean<<<N * N, 256, 0, st>>>(N, rmu, os_nb, k0, k1, k2, surf, status);
}

// SOS_BPDF_AJOUT_BRDF (SOS_SURFACE.F:2503-2669): the nine REAL*4 matrices of two surface files are added record by record
__global__ void k_ajout_brdf(float *out, const float *a, const float *b, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
extern "C" void sos_launch_ajout_brdf(float *out, const float *a, const float *b, size_t n, cudaStream_t st)
{
  if (n) k_ajout_brdf<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, a, b, n);
}
