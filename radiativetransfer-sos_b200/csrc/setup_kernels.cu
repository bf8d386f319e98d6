// setup_kernels.cu -- order-faithful small kernels of libsosgpu.so (compiled with -fmad=false so that
// every multiply/add rounds exactly like the reference's non-FMA x86-64 gfortran build):
//   k_basis / k_kernels : SOS_NOYAUX            (SOS_OS.F:1857-2158)
//   k_pack              : matrix form of SOS_FSOURCE_ORDREIG / _ORDRE1 / _DIFF_FRESNEL1 coefficients
//                         (SOS_OS.F:2520-2560, 2859-2905, 3237-3289)
//   k_att               : exp(-dtau/mu) table   (SOS_OS.F:2291,2335)
//   k_init / k_test     : accumulators and stop tests of the scattering loop (SOS_OS.F:1094-1137, 1248-1417,
//                         3377-3461, 3497-3659, 3871-4018)
//   k_fourier           : per-order bookkeeping + Fourier stop (SOS_OS.F:1421-1589, 3709-3796)
//   k_aggregate         : CKD-weighted sum (SOS_AGGREGATE.F:397-413)
#include "sosgpu_internal.h"
#include <math.h>

#define SEUIL_CV_SG  ((double)0.00001f)   // SOS.h:389 REAL*4 literal
#define SEUIL_SUMDIF ((double)0.00001f)   // SOS.h:394
#define SEUIL_VALDIF (1.0e-50)            // SOS.h:395
#define SEUIL_SF     ((double)0.00001f)   // SOS.h:400

// ------------------------------------------------------------------------------------------------
// SOS_NOYAUX part 1: generalised Legendre functions PSL, RSL, TSL(l, j), l = -1..NB, j = -N..N.
// One block per kernel set, one thread per j in 0..N (each thread writes only its +-j entries).
// basis must be zero-initialised (the reference relies on zeroed static storage).
__global__ void k_basis(const KsetDev *ksets, const OpticsDev *optics)
{
  const KsetDev ks = ksets[blockIdx.x];
  const OpticsDev &op = optics[ks.optics];
  const int N = op.nbmu, W = op.W, NB = op.os_nb, is = ks.is;
  const int LD = NB + 2;
  double *psl = ks.basis, *rsl = ks.basis + (size_t)LD * W, *tsl = ks.basis + (size_t)2 * LD * W;
#define PS(l, j) psl[(size_t)((l) + 1) * W + ((j) + N)]
#define RS(l, j) rsl[(size_t)((l) + 1) * W + ((j) + N)]
#define TS(l, j) tsl[(size_t)((l) + 1) * W + ((j) + N)]
  const int j = threadIdx.x;
  if (j > N) return;
  const double c = op.rmu[j + N];
  const double rac3 = sqrt(3.0);
  const double x26 = 2.0 * sqrt(6.0);
  if (is == 0) {                                              // :1968-1993
    PS(0, -j) = 1.0; PS(0, j) = 1.0;
    PS(1, j) = c;    PS(1, -j) = -c;
    double x = (3.0 * c * c - 1.0) * 0.5;
    PS(2, -j) = x;   PS(2, j) = x;
    RS(1, j) = 0.0;  RS(1, -j) = 0.0;
    x = 3.0 * (1.0 - c * c) / x26;
    RS(2, -j) = x;   RS(2, j) = x;
    TS(1, j) = 0.0;  TS(1, -j) = 0.0;
    TS(2, j) = 0.0;  TS(2, -j) = 0.0;
    if (j == 0) { PS(1, 0) = op.rmu[N]; RS(1, 0) = 0.0; }
  } else if (is == 1) {                                       // :1997-2023
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
    if (j == 0) { PS(2, 0) = -PS(2, 0); RS(2, 0) = -RS(2, 0); RS(1, 0) = 0.0; TS(1, 0) = 0.0; }
  } else {                                                    // :2027-2052
    double a = 1.0;
    for (int i = 1; i <= is; ++i) {
      double x = (double)i;
      a = a * sqrt((double)(i + is) / x) * 0.5;
    }
    double b = a * sqrt((double)is / ((double)is + 1.0)) * sqrt(((double)is - 1.0) / ((double)is + 2.0));
    double xx = 1.0 - c * c;
    double yy = (double)((float)is * 0.5f - 1.0f);
    PS(is - 1, j) = 0.0; RS(is - 1, j) = 0.0; TS(is - 1, j) = 0.0;
    double x = a * pow(xx, (double)((float)is * 0.5f));
    PS(is, -j) = x;  PS(is, j) = x;
    x = b * (1.0 + c * c) * pow(xx, yy);
    RS(is, -j) = x;  RS(is, j) = x;
    x = 2.0 * b * c * pow(xx, yy);
    TS(is, -j) = -x; TS(is, j) = x;
  }
  int k0 = 2;                                                 // :2058-2100
  if (is > 2) k0 = is;
  if (k0 != NB) {
    int ig = -1;
    if (is == 1) ig = 1;
    for (int l = k0; l <= NB - 1; ++l) {
      const int lp = l + 1, lm = l - 1;
      double a = (2.0 * l + 1.0) / sqrt(((double)(l + is) + 1.0) * ((double)(l - is) + 1.0));
      double b = sqrt((double)((l + is) * (l - is))) / (2.0 * l + 1.0);
      double d = ((double)l + 1.0) * (2.0 * l + 1.0) /
                 sqrt(((double)l + 3.0) * ((double)l - 1.0) * ((double)(l + is) + 1.0) * ((double)(l - is) + 1.0));
      double e = sqrt(((double)l + 2.0) * ((double)l - 2.0) * (double)(l + is) * (double)(l - is)) /
                 ((double)l * (2.0 * l + 1.0));
      float ff = __fdiv_rn(2.0f * (float)is, (float)l * ((float)l + 1.0f));   // all-REAL*4 expression (:2079)
      double f = (double)ff;
      double x = a * (c * PS(l, j) - b * PS(lm, j));
      PS(lp, j) = x;
      x = d * (c * RS(l, j) - f * TS(l, j) - e * RS(lm, j));
      RS(lp, j) = x;
      x = d * (c * TS(l, j) - f * RS(l, j) - e * TS(lm, j));
      TS(lp, j) = x;
      if (j != 0) {
        PS(lp, -j) = ig * PS(lp, j);
        RS(lp, -j) = ig * RS(lp, j);
        TS(lp, -j) = -ig * TS(lp, j);
      }
      ig = -ig;
    }
  }
  // XPL, XRL, XTL = l=2 rows (:2107-2111)
  ks.xpl[j + N] = PS(2, j);        ks.xpl[-j + N] = PS(2, -j);
  ks.xpl[W + j + N] = RS(2, j);    ks.xpl[W - j + N] = RS(2, -j);
  ks.xpl[2 * W + j + N] = TS(2, j); ks.xpl[2 * W - j + N] = TS(2, -j);
}

// SOS_NOYAUX part 2 (:2121-2155): six kernels, l ascending, one thread per (j,k)
__global__ void k_kernels(const KsetDev *ksets, const OpticsDev *optics)
{
  const KsetDev ks = ksets[blockIdx.y];
  const OpticsDev &op = optics[ks.optics];
  const int N = op.nbmu, W = op.W, NB = op.os_nb, is = ks.is;
  const int LD = NB + 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * W) return;
  const int j = idx % W - N, k = idx / W - N;                 // Fortran (j,k): first index fastest
  const double *psl = ks.basis, *rsl = ks.basis + (size_t)LD * W, *tsl = ks.basis + (size_t)2 * LD * W;
  double sbp = 0, satt = 0, sarr = 0, sgr = 0, sgt = 0, sart = 0;
  if (!(is > NB)) {
    for (int l = is; l <= NB; ++l) {
      const double pj = PS(l, j), pk = PS(l, k), rj = RS(l, j), rk = RS(l, k), tj = TS(l, j), tk = TS(l, k);
      const double al = op.alpha[l], be = op.beta[l], ga = op.gamma[l], ze = op.zeta[l];
      double r1 = tj * tk;
      double r2 = rj * rk;
      sbp = sbp + be * pj * pk;
      satt = satt + al * r1 + ze * r2;
      sarr = sarr + ze * r1 + al * r2;
      sgr = sgr + ga * pj * rk;
      sgt = sgt + ga * pj * tk;
      sart = sart + al * rk * tj + ze * rj * tk;
    }
  }
  const size_t WW = (size_t)W * W;
  ks.ker[0 * WW + idx] = sbp;
  ks.ker[1 * WW + idx] = sgr;
  ks.ker[2 * WW + idx] = sgt;
  ks.ker[3 * WW + idx] = sarr;
  ks.ker[4 * WW + idx] = sart;
  ks.ker[5 * WW + idx] = satt;
}
#undef PS
#undef RS
#undef TS

// ------------------------------------------------------------------------------------------------
// Matrix form of the order-n source (SOS_OS.F:2894-2905): J = XDEL*(A_A X) + YDEL*(A_R X),
// A[ro][co] = 0.5 * GA(j) * sign * E(.,.), rows = outputs (d_out, stokes, k), cols = inputs (d_in, stokes, j).
// Also the per-row coefficients of the first-order and flat-sea sources.
__device__ __forceinline__ double kerA(const KsetDev &ks, int which, int a, int b, int N, int W)
{
  return ks.ker[(size_t)which * W * W + (size_t)(b + N) * W + (a + N)];
}
// Rayleigh kernel elements built from the l=2 rows (SOS_OS.F:2859-2876)
__device__ __forceinline__ double kerR(const KsetDev &ks, const OpticsDev &op, int which, int a, int b, int N, int W)
{
  const double *xpl = ks.xpl, *xrl = ks.xpl + W, *xtl = ks.xpl + 2 * W;
  switch (which) {
    case 0: return ks.beta0 + op.beta2 * xpl[a + N] * xpl[b + N];       // BP
    case 1: return op.gamma2 * xpl[a + N] * xrl[b + N];                 // GR(a,b) = gamma2*XPL(a)*XRL(b)
    case 2: return op.gamma2 * xpl[a + N] * xtl[b + N];                 // GT
    case 3: return op.alpha2 * xrl[a + N] * xrl[b + N];                 // ARR
    case 4: return op.alpha2 * xtl[a + N] * xrl[b + N];                 // ART(a,b) = alpha2*XTL(a)*XRL(b)
    default: return op.alpha2 * xtl[a + N] * xtl[b + N];                // ATT
  }
}

__global__ void k_pack(const KsetDev *ksets, const OpticsDev *optics)
{
  const KsetDev ks = ksets[blockIdx.y];
  const OpticsDev &op = optics[ks.optics];
  const int N = op.nbmu, W = op.W, HB = op.HB, KP = op.KP;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)KP * KP) return;
  const int ro = (int)(idx / KP), co = (int)(idx % KP);
  const int dout = ro / HB, qo = ro % HB, din = co / HB, qi = co % HB;
  double va = 0.0, vr = 0.0;
  if (qo < 3 * N && qi < 3 * N) {
    const int so = qo / N, k = qo % N + 1, si = qi / N, j = qi % N + 1;
    const bool same = (dout == din);
    int which, a, b;
    double sign = 1.0;
    if (so == 0) {
      if (si == 0)      { which = 0; a = j; b = same ? k : -k; }
      else if (si == 1) { which = 1; a = k; b = same ? j : -j; }
      else              { which = 2; a = k; b = same ? j : -j; sign = (dout == 0) ? -1.0 : 1.0; }
    } else if (so == 1) {
      if (si == 0)      { which = 1; a = j; b = same ? k : -k; }
      else if (si == 1) { which = 3; a = j; b = same ? k : -k; }
      else              { which = 4; a = j; b = same ? k : -k; sign = (din == 0) ? -1.0 : 1.0; }
    } else {
      if (si == 0)      { which = 2; a = j; b = same ? k : -k; sign = (din == 0) ? -1.0 : 1.0; }
      else if (si == 1) { which = 4; a = k; b = same ? j : -j; sign = (dout == 0) ? -1.0 : 1.0; }
      else              { which = 5; a = j; b = same ? k : -k; }
    }
    const double z = 0.5 * op.ga[j + N] * sign;
    va = z * kerA(ks, which, a, b, N, W);
    if (ks.dual) vr = z * kerR(ks, op, which, a, b, N, W);
  }
  // slab-major, k swizzled by 4*(row&3): one contiguous R*128-byte block per (tile, k-slab) lands in shared
  // memory as a bank-conflict-free DMMA A tile (step_kernel.cu)
  ks.apackA[((size_t)(co >> 4) * KP + ro) * 16 + ((co & 15) ^ (4 * (ro & 3)))] = va;
  if (ks.dual) ks.apackR[idx] = vr;

  if (co == 0) {                                   // one thread per row: first-order coefficients
    double c1 = 0.0, c2 = 0.0, f1 = 0.0, f2 = 0.0, ur = 0.0;
    if (qo < 3 * N) {
      const int so = qo / N, k = qo % N + 1;
      const int jj = (dout == 0) ? k : -k;
      const double *xpl = ks.xpl, *xrl = ks.xpl + W, *xtl = ks.xpl + 2 * W;
      const double spl = xpl[N];                               // XPL(JK), JK = 0 (:2526)
      ur = (so == 0) ? xpl[k + N] : (so == 1 ? xrl[k + N] : xtl[k + N]);
      if (so == 0)      { c2 = kerA(ks, 0, 0, jj, N, W);  if (ks.dual) c1 = ks.beta0 + op.beta2 * xpl[jj + N] * spl; }
      else if (so == 1) { c2 = kerA(ks, 1, 0, jj, N, W);  if (ks.dual) c1 = op.gamma2 * xrl[jj + N] * spl; }
      else              { c2 = -kerA(ks, 2, 0, jj, N, W); if (ks.dual) c1 = -(op.gamma2 * xtl[jj + N] * spl); }
      if (op.ifresnel == 1) {                                  // SOS_FSOURCE_DIFF_FRESNEL1 (:3237-3289)
        const int jm = -jj;                                    // mirrored direction
        const double F1 = op.f11sun, F2 = op.f12sun;
        if (so == 0) {
          f2 = F1 * kerA(ks, 0, 0, jm, N, W) + F2 * kerA(ks, 1, jm, 0, N, W);
          if (ks.dual) f1 = F1 * (ks.beta0 + op.beta2 * xpl[jm + N] * spl) + F2 * (op.gamma2 * xrl[N] * xpl[jm + N]);
        } else if (so == 1) {
          f2 = F1 * kerA(ks, 1, 0, jm, N, W) + F2 * kerA(ks, 3, 0, jm, N, W);
          if (ks.dual) f1 = F1 * (xrl[jm + N] * spl * op.gamma2) + F2 * (op.alpha2 * xrl[N] * xrl[jm + N]);
        } else {
          f2 = F1 * kerA(ks, 2, 0, jm, N, W) + F2 * kerA(ks, 4, jm, 0, N, W);
          if (ks.dual) f1 = F1 * (op.gamma2 * spl * xtl[jm + N]) + F2 * (op.alpha2 * xtl[jm + N] * xrl[N]);
        }
      }
    }
    ks.c1[ro] = c1; ks.c2[ro] = c2; ks.fz1[ro] = f1; ks.fz2[ro] = f2; ks.urow[ro] = ur;
  }
}

// Rank-4 factorisation of the molecular part per direction block (SOS_OS.F:2859-2876: every element of the
// Rayleigh phase matrix is built from the l=2 rows only, so for a fixed output direction block
//   I rows:  A_R[I(k)] = v0 + XPL(k) v1,   Q rows: A_R[Q(k)] = XRL(k) v2,   U rows: A_R[U(k)] = XTL(k) v3 ).
// The functionals v_r are read off the dense rows with the best-conditioned k.  One thread per (dir, column).
__global__ void k_pack_v(const KsetDev *ksets, const OpticsDev *optics)
{
  const KsetDev ks = ksets[blockIdx.y];
  if (!ks.dual) return;
  const OpticsDev &op = optics[ks.optics];
  const int N = op.nbmu, W = op.W, HB = op.HB, KP = op.KP;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * KP) return;
  const int d = idx / KP, co = idx % KP;
  const double *xpl = ks.xpl, *xrl = ks.xpl + W, *xtl = ks.xpl + 2 * W;
  int k1 = 1, k2 = 1, kr = 1, kt = 1;
  for (int k = 2; k <= N; ++k) {
    if (xpl[k + N] > xpl[k1 + N]) k1 = k;
    if (xpl[k + N] < xpl[k2 + N]) k2 = k;
    if (fabs(xrl[k + N]) > fabs(xrl[kr + N])) kr = k;
    if (fabs(xtl[k + N]) > fabs(xtl[kt + N])) kt = k;
  }
  const double *R = ks.apackR;
  const size_t rowI = (size_t)(d * HB), rowQ = (size_t)(d * HB + N), rowU = (size_t)(d * HB + 2 * N);
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  const double m1 = R[(rowI + k1 - 1) * KP + co], m2 = R[(rowI + k2 - 1) * KP + co];
  const double p1 = xpl[k1 + N], p2 = xpl[k2 + N];
  if (ks.beta0 == 0.0) {                           // the isotropic term exists for the Fourier order 0 only (:890)
    const int kb = (fabs(p1) >= fabs(p2)) ? k1 : k2;
    const double pb = xpl[kb + N];
    if (pb != 0.0) v[1] = R[(rowI + kb - 1) * KP + co] / pb;
  } else if (p1 != p2) {
    v[1] = (m1 - m2) / (p1 - p2);
    v[0] = m1 - p1 * v[1];
  } else v[0] = m1;
  if (xrl[kr + N] != 0.0) v[2] = R[(rowQ + kr - 1) * KP + co] / xrl[kr + N];
  if (xtl[kt + N] != 0.0) v[3] = R[(rowU + kt - 1) * KP + co] / xtl[kt + N];
  double *vp = ks.vpack + (size_t)d * 8 * KP;
  for (int r = 0; r < 8; ++r)
    vp[((size_t)(co >> 4) * 8 + r) * 16 + ((co & 15) ^ (4 * (r & 3)))] = (r < 4) ? v[r] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// att[layer][k-1] = a = exp(-dt[layer]/mu_k)  (SOS_OS.F:2291 DEXP(-DTAU/RMUK); :2335 is the same value), and the two weights
// of the layer integration in the form the sweep kernel uses:  with S the source function at the two levels of the layer,
//   SOS_OS.F:2288-2309:  (1-a)*(A*mu + S_i) - A*a*dt,  A = (S_j - S_i)/dt   ==   (1 - a - g)*S_i + g*S_j,
//   g = (1-a)*mu/dt - a   (same g for the downward sweep, :2332-2353).
// g = x/2 - x^2/3 + ... for x = dt/mu -> 0: the closed form cancels there, so small x uses the series
// g = sum_{n>=1} (-1)^(n+1) n x^n / (n+1)!  (error below 1e-17 for x < 0.25 with 14 terms).
__global__ void k_att(const TermDev *terms, const OpticsDev *optics)
{
  const TermDev tm = terms[blockIdx.y];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= tm.nt * N) return;
  const int layer = idx / N, k = idx % N + 1;
  const double x = tm.dt[layer] / op.rmu[k + N];
  const double a = exp(-tm.dt[layer] / op.rmu[k + N]);
  const_cast<double *>(tm.att)[idx] = a;
  const double t = -expm1(-x);                                // 1 - a without cancellation
  double g;
  if (x < 0.25) {
    double term = x * 0.5, sum = term;                         // n = 1
    for (int n = 2; n <= 14; ++n) {
      term = -term * x * (double)n / ((double)(n - 1) * (double)(n + 1));   // ratio of consecutive terms: -x n / ((n-1)(n+1))
      sum = sum + term;
    }
    g = sum;
  } else g = t / x - a;
  const double b = t - g;
  const_cast<double *>(tm.gco)[idx] = g;
  const_cast<double *>(tm.bco)[idx] = b;
  // Aerosol-only Fourier orders (S = XDEL * accumulator): weights with XDEL folded in, indexed by the level they update.
  //   upward   level i = layer,     neighbour i+1:  c = pup[i]*acc(i) + qup[i]*acc(i+1)
  //   downward level i = layer + 1, neighbour i-1:  c = pdn[i]*acc(i) + qdn[i]*acc(i-1)
  const double xd_lo = tm.xdel[layer], xd_hi = tm.xdel[layer + 1];
  const size_t up_i = (size_t)layer * N + (k - 1), dn_i = (size_t)(layer + 1) * N + (k - 1);
  const_cast<double *>(tm.pup)[up_i] = b * xd_lo;
  const_cast<double *>(tm.qup)[up_i] = g * xd_hi;
  const_cast<double *>(tm.pdn)[dn_i] = b * xd_hi;
  const_cast<double *>(tm.qdn)[dn_i] = g * xd_lo;
  // First scattering order (k_order1): the source c2(row)*sx + c1(row)*sy of the two levels folded with b and g (k_beam ran before).
  const double sx_lo = tm.sxd[layer], sx_hi = tm.sxd[layer + 1], sy_lo = tm.syd[layer], sy_hi = tm.syd[layer + 1];
  const_cast<double *>(tm.o1u2)[up_i] = b * sx_lo + g * sx_hi;
  const_cast<double *>(tm.o1u1)[up_i] = b * sy_lo + g * sy_hi;
  const_cast<double *>(tm.adn)[dn_i] = a;
  const_cast<double *>(tm.o1d2)[dn_i] = b * sx_hi + g * sx_lo;
  const_cast<double *>(tm.o1d1)[dn_i] = b * sy_hi + g * sy_lo;
}

// Per-level attenuation of the direct solar beam: CH(i) = exp(-H(i)/mu_s)/4 (SOS_OS.F:837-839) and, for a flat sea, the beam
// reflected by the surface CF(i) = exp(2 H(NT)/TAB)/4 * exp(-H(i)/TAB) (SOS_OS.F:3219, 3278, 3285); TAB = -mu_s.
__global__ void k_beam(const TermDev *terms, const OpticsDev *optics)
{
  const TermDev tm = terms[blockIdx.x];
  const OpticsDev &op = optics[tm.optics];
  const double coefnt = exp(2.0 * tm.h[tm.nt] / op.tab) / 4.0;
  for (int i = threadIdx.x; i <= tm.nt; i += blockDim.x) {
    const double ch = exp(-tm.h[i] / (-op.tab)) / 4.0;
    const_cast<double *>(tm.ch)[i] = ch;
    const_cast<double *>(tm.cf)[i] = coefnt * exp(-tm.h[i] / op.tab);
    const_cast<double *>(tm.sxd)[i] = ch * tm.xdel[i];          // order-1 source = c2(row)*sxd + c1(row)*syd (SOS_OS.F:2553-2560)
    const_cast<double *>(tm.syd)[i] = ch * tm.ydel[i];
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_max(double v, double *red)
{
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmax(r, red[i]);
  return r;
}

// component c -> (row of the packed field, level of the TOA/BOA sample)
__device__ __forceinline__ int comp_row(int c, int N, int HB) { const int d = c / (3 * N); return d * HB + (c - d * 3 * N); }

// After the first-order field: accumulators, histories, direct surface term (SOS_OS.F:1051-1137)
__global__ void k_init(ItemDev *items, const TermDev *terms, const OpticsDev *optics)
{
  ItemDev &it = items[blockIdx.x];
  const TermDev &tm = terms[it.term];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu, HB = op.HB, NT = tm.nt, KP = op.KP;
  const double *x1 = it.x[1];
  for (int c = threadIdx.x; c < 6 * N; c += blockDim.x) {
    const int d = c / (3 * N);
    const int r = comp_row(c, N, HB);
    const double v = x1[SOS_XIDX(KP, r, d == 0 ? 0 : NT)];
    it.sum3[c] = v;
    it.hist_d[c] = v;
    it.hist_a[c] = 0.0;
    if (it.sumout) {
      it.sumout[c] = x1[SOS_XIDX(KP, r, tm.jout - 1)];
      it.sumout[6 * N + c] = x1[SOS_XIDX(KP, r, tm.jout)];
    }
    if (d == 0) {
      double rv = 0.0, ro0 = 0.0, ro1 = 0.0;
      if (op.imat_surf == 1) {
        const int so = c / N, k = c % N + 1;
        const double mu = op.rmu[k + N];
        double g = x1[SOS_XIDX(KP, r, NT)];                   // I1(NT,K) (boundary value is kept by the integration)
        if (so == 0) {
          double xr = 0.0;
          if (!(op.ro == 0.0 || it.is != 0)) xr = -op.ro * op.tab * tm.eground;   // XR(K) (:979-980)
          g = g - xr;
        }
        double a = -tm.h[NT] / mu;
        rv = exp(a) * g;                                      // :1076-1080
        if (it.riiout) {
          a = -(tm.h[NT] - tm.h[tm.jout - 1]) / mu; ro0 = exp(a) * g;   // :1068-1072
          a = -(tm.h[NT] - tm.h[tm.jout]) / mu;     ro1 = exp(a) * g;
        }
      }
      it.rii[c] = rv;
      if (it.riiout) { it.riiout[c] = ro0; it.riiout[3 * N + c] = ro1; }
    }
  }
  if (threadIdx.x == 0) {
    it.n = 1;
    it.active = 1;
    it.reason = -1;
    if (op.igmax < 2) { it.n = 2; it.active = 0; it.reason = 0; }   // IG=2 > IGMAX at label 503 (:1152)
  }
}

// One scattering order has been integrated (field of order n+1 sits in x[(n+1)&1]):
// SOS_OS.F:1248-1417 with SOS_PARAM_CONV, SOS_AJOUT_QUEUE, SOS_ARRET_DIFFUS_1/2.
__global__ void k_test(ItemDev *items, const TermDev *terms, const OpticsDev *optics,
                       const int *list_cur, const int *count_cur, int *list_next, int *count_next)
{
  __shared__ double red[32];
  if (count_cur && (int)blockIdx.x >= *count_cur) return;      // the grid is an upper bound: the wave loop never waits for the count
  const int item = list_cur ? list_cur[blockIdx.x] : (int)blockIdx.x;
  ItemDev &it = items[item];
  if (!it.active) return;                                     // IG=2 > IGMAX at label 503: no scattering order at all
  const TermDev &tm = terms[it.term];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu, HB = op.HB, NT = tm.nt, KP = op.KP;
  const int ig = it.n + 1;                                    // label 503: IG=IG+1
  const double *xn = it.x[ig & 1];
  const double *xp = it.x[(ig - 1) & 1];
  const int NC = 6 * N;

  double zloc = 0.0;
  if (ig != 2) {                                              // SOS_PARAM_CONV (:3430-3458)
    for (int c = threadIdx.x; c < NC; c += blockDim.x) {
      const int d = c / (3 * N);
      const double g = xn[SOS_XIDX(KP, comp_row(c, N, HB), d == 0 ? 0 : NT)];
      const double a1 = it.hist_a[c], d1 = it.hist_d[c], s3 = it.sum3[c];
      if (a1 != 0.0 && d1 != 0.0 && s3 != 0.0) {
        const double r = 1 - g / d1;
        const double y = (g / d1 - d1 / a1) / (r * r) * (g / s3);
        zloc = fmax(zloc, fabs(y));
      }
    }
  }
  const double zconv = block_max(zloc, red);
  if (ig != 2 && !(zconv > SEUIL_CV_SG)) {                    // geometric tail (:1293-1315, SOS_AJOUT_QUEUE)
    for (int c = threadIdx.x; c < NC; c += blockDim.x) {
      const int d = c / (3 * N);
      const int crow = comp_row(c, N, HB);
      const double g = xn[SOS_XIDX(KP, crow, d == 0 ? 0 : NT)];
      const double d1 = it.hist_d[c];
      const double q = (d1 == 0.0) ? 0.0 : g / (1 - g / d1);
      it.sum3[c] = it.sum3[c] + q;
      if (it.sumout) {
        for (int lv = 0; lv < 2; ++lv) {
          const int level = tm.jout - 1 + lv;
          const double go = xn[SOS_XIDX(KP, crow, level)], dprev = xp[SOS_XIDX(KP, crow, level)];
          const double qo = (dprev == 0.0) ? 0.0 : go / (1 - go / dprev);
          it.sumout[lv * NC + c] = it.sumout[lv * NC + c] + qo;
        }
      }
    }
    if (threadIdx.x == 0) { it.n = ig; it.active = 0; it.reason = 1; }
    return;
  }
  // label 506: shift histories, accumulate (:1323-1363), then tests (:1368-1406)
  double z1 = 0.0, z2 = 0.0;
  for (int c = threadIdx.x; c < NC; c += blockDim.x) {
    const int d = c / (3 * N);
    const int crow = comp_row(c, N, HB);
    const double g = xn[SOS_XIDX(KP, crow, d == 0 ? 0 : NT)];
    it.hist_a[c] = it.hist_d[c];
    it.hist_d[c] = g;
    const double s3 = it.sum3[c] + g;
    it.sum3[c] = s3;
    if (it.sumout) {
      it.sumout[c] = it.sumout[c] + xn[SOS_XIDX(KP, crow, tm.jout - 1)];
      it.sumout[NC + c] = it.sumout[NC + c] + xn[SOS_XIDX(KP, crow, tm.jout)];
    }
    z1 = fmax(z1, fabs(g));
    if (s3 != 0.0) z2 = fmax(z2, fabs(g / s3));
  }
  z1 = block_max(z1, red);
  z2 = block_max(z2, red);
  if (threadIdx.x == 0) {
    it.n = ig;
    int reason = -1;
    if (!(z1 > SEUIL_VALDIF)) reason = 2;
    else if (!(z2 > SEUIL_SUMDIF)) reason = 3;
    else if (!(ig < op.igmax)) reason = 4;
    if (reason >= 0) { it.active = 0; it.reason = reason; }
    else { const int p = atomicAdd(count_next, 1); list_next[p] = item; }
  }
}

// Per term, in Fourier-order sequence over the wave [s0,s1): SOS_OS.F:1421-1589
__global__ void k_fourier(ItemDev *items, TermDev *terms, const OpticsDev *optics,
                          const int *item_of, int s0, int s1, int rec_stride, int wdev,
                          double *rec, int *n_fourier, int *n_scatter, int *stop_reason,
                          double *emoins, double *eplus, int *done)
{
  __shared__ double red[32];
  extern __shared__ double sh[];                              // I3,Q3,U3 of the current order [6N]
  const int b = blockIdx.x;
  if (done[b]) return;
  TermDev &tm = terms[b];
  const OpticsDev &op = optics[tm.optics];
  const int N = op.nbmu, NC = 6 * N, WS = s1 - s0;
  double *i4 = tm.i4, *i5 = tm.i4 + NC;
  for (int s = s0; s < s1 && s <= tm.iborm; ++s) {
    const int item = item_of[(size_t)b * WS + (s - s0)];
    const ItemDev &it = items[item];
    const double sign = (s & 1) ? -1.0 : 1.0;
    const double coef = (s == 0) ? 1.0 : 2.0;
    for (int c = threadIdx.x; c < NC; c += blockDim.x) {
      double v = it.sum3[c];
      if (op.imat_surf == 1 && c < 3 * N) v = v - it.rii[c];  // :1421-1439
      sh[c] = v;
    }
    __syncthreads();
    if (s == 0 && threadIdx.x == 0) {                         // fluxes (:1447-1456)
      double em = 0.0, ep = 0.0;
      for (int j = 1; j <= N; ++j) {
        em = em + op.rmu[j + N] * op.ga[j + N] * sh[3 * N + (j - 1)];
        ep = ep + op.rmu[j + N] * op.ga[j + N] * sh[j - 1];
      }
      emoins[b] = -em * 2 / op.tab;
      eplus[b] = -ep * 2 / op.tab;
    }
    double z = 0.0;
    for (int c = threadIdx.x; c < NC; c += blockDim.x) {      // :1460-1473 and SOS_ARRET_FOURIER
      const double v = sh[c];
      const double a4 = i4[c] + coef * v;
      const double a5 = i5[c] + coef * v * sign;
      i4[c] = a4; i5[c] = a5;
      if (a4 != 0.0) z = fmax(z, fabs(v / a4));
      if (a5 != 0.0) z = fmax(z, fabs(v / a5));
      // record in file order Q,U,I (:1572-1574); element j at [j+N]
      const int d = c / (3 * N), cc = c - d * 3 * N, so = cc / N, k = cc % N + 1;
      const int j = (d == 0) ? k : -k;
      const int slot = (so == 0) ? 2 : (so == 1 ? 0 : 1);
      double out = v;
      if (it.sumout) {                                        // zout != -1 (:1514-1532)
        double v0 = it.sumout[c], v1 = it.sumout[NC + c];
        if (op.imat_surf == 1 && c < 3 * N) { v0 = v0 - it.riiout[c]; v1 = v1 - it.riiout[3 * N + c]; }
        out = (1 - tm.zz) * v0 + tm.zz * v1;
      }
      rec[(((size_t)b * rec_stride + s) * 3 + slot) * wdev + (j + N)] = out;
    }
    z = block_max(z, red);
    if (threadIdx.x == 0) {
      n_fourier[b] = s + 1;
      n_scatter[(size_t)b * rec_stride + s] = it.n;
      stop_reason[(size_t)b * rec_stride + s] = it.reason;
    }
    if (!(z > SEUIL_SF)) {                                    // :1585-1589
      if (threadIdx.x == 0) done[b] = 1;
      return;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && s1 > tm.iborm) done[b] = 1;
}

// RES = RES + AIK*TMP per Fourier record, in term order (SOS_AGGREGATE.F:397-413)
__global__ void k_aggregate(const TermDev *terms, const int *group_start, const int *group_terms,
                            const double *rec, const int *n_fourier, int rec_stride, int wdev,
                            double *grec, int *gnrec)
{
  const int g = blockIdx.y;
  const size_t per = (size_t)rec_stride * 3 * wdev;
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = group_start[g], t1 = group_start[g + 1];
  if (x == 0) {
    int m = 0;
    for (int t = t0; t < t1; ++t) m = max(m, n_fourier[group_terms[t]]);
    gnrec[g] = m;
  }
  if (x >= per) return;
  double r = 0.0;
  for (int t = t0; t < t1; ++t) {
    const int b = group_terms[t];
    r = r + terms[b].aik * rec[(size_t)b * per + x];
  }
  grec[(size_t)g * per + x] = r;
}

// ------------------------------------------------------------------------------------------------
extern "C" {
void sos_launch_basis(const KsetDev *ksets, const OpticsDev *optics, int nkset, cudaStream_t st)
{
  if (nkset > 0) k_basis<<<nkset, 96, 0, st>>>(ksets, optics);
}
void sos_launch_kernels(const KsetDev *ksets, const OpticsDev *optics, int nkset, int maxW, cudaStream_t st)
{
  if (nkset <= 0) return;
  dim3 grid((maxW * maxW + 127) / 128, nkset);
  k_kernels<<<grid, 128, 0, st>>>(ksets, optics);
}
void sos_launch_pack(const KsetDev *ksets, const OpticsDev *optics, int nkset, int maxKP, cudaStream_t st)
{
  if (nkset <= 0) return;
  dim3 grid((unsigned)(((size_t)maxKP * maxKP + 255) / 256), nkset);
  k_pack<<<grid, 256, 0, st>>>(ksets, optics);
  dim3 gridv((unsigned)((2 * maxKP + 127) / 128), nkset);
  k_pack_v<<<gridv, 128, 0, st>>>(ksets, optics);
}
void sos_launch_att(const TermDev *terms, const OpticsDev *optics, int nterm, int max_elems, cudaStream_t st)
{
  if (nterm <= 0) return;
  dim3 grid((max_elems + 127) / 128, nterm);
  k_beam<<<nterm, 128, 0, st>>>(terms, optics);                 // k_att reads its sxd / syd
  k_att<<<grid, 128, 0, st>>>(terms, optics);
}
void sos_launch_init(ItemDev *items, const TermDev *terms, const OpticsDev *optics, int nitem, cudaStream_t st)
{
  if (nitem > 0) k_init<<<nitem, 256, 0, st>>>(items, terms, optics);
}
void sos_launch_test(ItemDev *items, const TermDev *terms, const OpticsDev *optics,
                     const int *list_cur, const int *count_cur, int ncur, int *list_next, int *count_next, cudaStream_t st)
{
  if (ncur > 0) k_test<<<ncur, 256, 0, st>>>(items, terms, optics, list_cur, count_cur, list_next, count_next);
}
void sos_launch_fourier(ItemDev *items, TermDev *terms, const OpticsDev *optics, int nterm,
                        const int *item_of, int s0, int s1, int rec_stride_dev, int wdev,
                        double *rec, int *n_fourier, int *n_scatter, int *stop_reason,
                        double *emoins, double *eplus, int *done, cudaStream_t st)
{
  if (nterm <= 0) return;
  k_fourier<<<nterm, 256, 6 * 80 * sizeof(double), st>>>(items, terms, optics, item_of, s0, s1, rec_stride_dev, wdev,
                                                        rec, n_fourier, n_scatter, stop_reason, emoins, eplus, done);
}
void sos_launch_aggregate(const TermDev *terms, const int *group_start, const int *group_terms, int ngroup,
                          const double *rec, const int *n_fourier, int rec_stride_dev, int wdev,
                          double *grec, int *gnrec, cudaStream_t st)
{
  if (ngroup <= 0) return;
  const size_t per = (size_t)rec_stride_dev * 3 * wdev;
  dim3 grid((unsigned)((per + 255) / 256), ngroup);
  k_aggregate<<<grid, 256, 0, st>>>(terms, group_start, group_terms, rec, n_fourier, rec_stride_dev, wdev, grec, gnrec);
}
}
