// glitter_kernel.cu -- rough-sea (Cox & Munk) reflection matrices, one fused kernel per surface file:
//   SOS_GSF + SOS_CALCG           (SOS_GLITTER.F:451-711, 755-784)  Fourier series of G(theta1,theta2,phi)
//   SOS_NOYAUX_FRESNEL            (SOS_SURFACE.F:2029-2227)         Fresnel phase-matrix kernels per Fourier order
//   SOS_MAT_REFLEXION             (SOS_SURFACE.F:1708-1973)         M_ij(s) = sum_k (+-)(G(k+s) +- G(|k-s|))/4 * kernel_k / sigma^2
//   SOS_MISE_FORMAT               (SOS_SURFACE.F:2307-2443)         pair-major -> order-major records, REAL*4
// One CTA per (theta1 >= theta2) pair; the three temporary files of the reference never exist.
// Compiled with -fmad=false: same operation order as the reference (only exp/cos/pow differ from glibc by <= 2 ulp).
// Compute-type kernel (FP64 CUDA cores + SFU): ~1025 exp and up to (OS_NM+1)*1023 cos per pair.
#include "sosgpu_internal.h"
#include "post_kernels.h"
#include "nadal_series.cuh"
#include <math.h>

#define PH_NU 1024      // CTE_PH_NU   SOS.h:319
#define PH_NQ 10        // CTE_PH_NQ   SOS.h:325
#define PH_TEST 10000   // CTE_PH_TEST SOS.h:312

// SOS_CALCG, SOS_GLITTER.F:779-781
__device__ __forceinline__ double calcg(double cs12, double c12, double s12, double sig, double phi)
{
  const double costetad = -c12 + s12 * cos(phi);
  const double x = (1 - costetad) / cs12;
  return x * x * exp(-(x - 1) / sig);
}

// SOS_CALCG_MAIGNAN, SOS_SURFACE_BPDF.F:1606-1641
__device__ __forceinline__ double calcg_maignan(double c1, double c2, double s12, double phi, double coef_c)
{
  const double cos_2i = c1 * c2 - s12 * cos(phi);
  double tan2_i = (1 - cos_2i) / (1 + cos_2i);
  if (tan2_i < 0.0) tan2_i = 0.0;
  return coef_c * exp(-sqrt(tan2_i)) / (1. / c1 + 1. / c2);
}

// B1(s) = max over the azimuths of a non-negative deviation: non-negative doubles order like their bit patterns, and NaN (0/0) has
// a larger pattern than any finite value, which makes the comparisons B1 > threshold / B1 < B1_PREC false, as in the reference
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// smem: U[1025] | G[nm_alloc] | K[6][2][ns+1] | B1[nb+1] (gmodel 4 only) | scalars
__global__ void k_glitter(GlitterParams p, float *__restrict__ surf, int *__restrict__ il_out)
{
  extern __shared__ double sh[];
  double *U = sh;
  double *G = U + PH_NU + 1;
  const int ng = p.os_nm + p.os_ns + p.os_nb + 2;
  double *K = G + ng;                                       // kernel (which, is, dir) at K[(which*(ns+1) + is)*2 + dir]
  __shared__ double s_phib, s_q, s_gmax, s_gmin;
  __shared__ int s_il;

  // pair index -> (I, J<=I), pairs ordered I outer, J inner (SOS_GLITTER.F:523,532)
  const int pair = blockIdx.x;
  int I = (int)((sqrt(8.0 * pair + 1.0) - 1.0) * 0.5) + 1;
  while (I * (I - 1) / 2 > pair) --I;
  while ((I + 1) * I / 2 <= pair) ++I;
  const int J = pair - I * (I - 1) / 2 + 1;
  const int N = p.nbmu, NS = p.os_ns, NB = p.os_nb, NM = p.os_nm;
  const double pi = p.pi, sig = p.sig;
  const double c1 = p.rmu[I + N], c2 = p.rmu[J + N];
  const double s1 = sqrt(1 - c1 * c1), s2 = sqrt(1 - c2 * c2);
  const double c12 = c1 * c2, s12 = s1 * s2;
  double cs12 = (c1 + c2);
  cs12 = .5 * cs12 * cs12;
  const int tid = threadIdx.x, nthr = blockDim.x;
  // G(theta1, theta2, phi): Cox-Munk wave-slope function (SOS_CALCG) or Maignan's BPDF function (gmodel 3, SOS_CALCG_MAIGNAN)
  auto gfun = [&](double phi) -> double {
    return (p.gmodel == 3) ? calcg_maignan(c1, c2, s12, phi, p.coef_c) : calcg(cs12, c12, s12, sig, phi);
  };

  if (p.gmodel == 1 || p.gmodel == 2) {
    // ---------------- SOS_GSF_RONDEAUX_BREON (SOS_SURFACE_BPDF.F:463-591): G does not depend on the azimuth ----------------
    for (int i = tid; i < ng; i += nthr) G[i] = 0.0;
    __syncthreads();
    if (tid == 0) {
      G[0] = (p.gmodel == 1) ? 1. / (1. / c1 + 1. / c2) : 1.;    // Rondeaux : Breon
      s_il = 0;
      if (il_out) il_out[pair] = 0;
    }
    __syncthreads();
  } else if (p.gmodel == 4) {
    // ---------------- SOS_F21SF_NADAL (SOS_SURFACE_BPDF.F:686-1069): series of F(phi) up to OS_NB with a recombination test ----------------
    // The reference writes one series per (I1, I2) for all N*N pairs and SOS_MAT_REFLEXION reads the file sequentially for the
    // N(N+1)/2 pairs (I, J <= I) (SOS_SURFACE.F:1832-1842): pair number `pair` gets the series of (pair / N + 1, pair mod N + 1).
    // nadal_pairing 0 reproduces that file; 1 takes the series of the pair (I, J) itself.
    const int a1 = p.nadal_pairing ? I : pair / N + 1, a2 = p.nadal_pairing ? J : pair % N + 1;
    const double n1 = p.rmu[a1 + N], n2 = p.rmu[a2 + N];
    const double q = pi / NAD_PH_NU;
    double *B1 = K + 12 * (NS + 1);
    for (int i = tid; i <= NAD_PH_NU; i += nthr) U[i] = nadal_f(p.ind, p.alpha_nadal, p.beta_nadal, n1, n2, q * i);   // :773-782
    for (int i = tid; i < ng; i += nthr) G[i] = 0.0;
    for (int s = tid; s <= NB; s += nthr) B1[s] = 0.0;
    __syncthreads();
    for (int s = tid; s <= NB; s += nthr) G[s] = nadal_coef(U, s, q, pi);                                             // :856-861
    __syncthreads();
    for (int i = tid; i <= NAD_PH_NU; i += nthr) {           // :868-883, one azimuth per thread walking the orders upwards
      const double phi = i * q, f = U[i];
      double t1 = G[0];
      atomic_max_nonneg(&B1[0], fabs((t1 - f) / f));
      for (int s = 1; s <= NB; ++s) {
        t1 = nadal_recomb_step(t1, G[s], s, phi);
        atomic_max_nonneg(&B1[s], fabs((t1 - f) / f));
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int il = nadal_cut(B1, NB);                      // :888-900
      s_il = il;
      if (il_out) il_out[pair] = il;
    }
    __syncthreads();
  } else {
  // ---------------- SOS_GSF: support [0, PHIB] of G(phi) ----------------
  if (tid == 0) {
    double g = gfun(0.0);
    const double gmax = g;
    g = gfun(pi);
    double gmin = g, phib, q;
    double x = PH_TEST * gmin;
    if (x >= gmax) {                                         // :568-573
      phib = pi;
      q = pi / PH_NU;
    } else {                                                 // bisection :586-620
      double phi1 = 0, phi2 = pi;
      for (;;) {
        phib = .5 * (phi1 + phi2);
        g = gfun(phib);
        x = PH_TEST * g;
        if (fabs(x - gmax) < (double)0.01f * gmax) break;
        if (x <= gmax) phi2 = phib; else phi1 = phib;
      }
      q = phib / PH_NU;
      gmin = -1.0;                                           // marker: GMIN = U(NU) after the table is filled (:635)
    }
    s_phib = phib; s_q = q; s_gmax = gmax; s_gmin = gmin;
    U[0] = gmax;
  }
  __syncthreads();
  const double phib = s_phib, q = s_q, gmax = s_gmax;
  for (int i = 1 + tid; i <= PH_NU; i += nthr) U[i] = gfun(q * i);   // :574-578 / :630-634
  for (int i = tid; i < ng; i += nthr) G[i] = 0.0;
  __syncthreads();
  const double gmin = (s_gmin < 0.0) ? U[PH_NU] : s_gmin;

  // ---------------- Fourier coefficients E(IS) by nested trapezoid refinement (:644-663) ----------------
  for (int is = tid; is <= NM; is += nthr) {
    double z = .5 * (gmax + gmin * cos(is * phib));
    int ia = 1;
    for (int i = 1; i <= PH_NQ; ++i) {
      ia = 2 * ia;
      const int ip = PH_NU / ia;
      double y = 0;
      for (int j = 1; j <= ia; j += 2) {
        const int k = ip * j;
        y = y + U[k] * cos((is * k) * q);
      }
      y = 2 * y / ia;
      const double xt = fabs(z - y) / z;
      if (xt < (double)0.0001f) break;
      z = .5 * (y + z);
    }
    G[is] = phib * z / pi;
  }
  __syncthreads();
  if (tid == 0) {                                            // series cut (:664-679), sequential in IS
    int il = NM;
    double t1 = G[0];
    for (int is = 1; is <= NM; ++is) {
      t1 = t1 + 2 * G[is];
      const double b1 = fabs(t1 - gmax) / gmax;
      if (!(b1 > (double)0.001f)) { il = is; break; }           // same cut in SOS_GSF (:664-679) and SOS_GSF_MAIGNAN (SOS_SURFACE_BPDF.F:1497-1502);
                                                                // Maignan's cusped G rarely converges to 1e-3 below OS_NM, but some pairs do
    }
    s_il = il;
    if (il_out) il_out[pair] = il;
  }
  __syncthreads();
  }
  const int lim = s_il;
  for (int i = lim + 1 + tid; i <= NM; i += nthr) G[i] = 0.0;        // SOS_SURFACE.F:1846-1848

  // ---------------- SOS_NOYAUX_FRESNEL: one thread per Fourier order, sums accumulated along the recurrence ----------------
  for (int is = tid; is <= NS; is += nthr) {
    const double r[2] = {c1, c2};
    double pm[2], pc[2], rm[2], rc[2], tm_[2], tc[2];       // values at L-1 and L
    double sbp[2] = {0, 0}, sgr[2] = {0, 0}, sgt[2] = {0, 0}, sarr[2] = {0, 0}, sart[2] = {0, 0}, satt[2] = {0, 0};
    // term L of the six sums for K = 1,2 (J = 3-K): index d = K-1, other direction o = 1-d
#define ACCUM(L_)                                                                                      \
    for (int d = 0; d < 2; ++d) {                                                                      \
      const int o = 1 - d;                                                                             \
      sbp[d] = sbp[d] + p.beta[L_] * pc[o] * pc[d];                                                    \
      sgr[d] = sgr[d] + p.gamma[L_] * pc[o] * rc[d];                                                   \
      sgt[d] = sgt[d] + p.gamma[L_] * pc[o] * tc[d];                                                   \
      satt[d] = satt[d] + p.alpha[L_] * tc[o] * tc[d] + p.zeta[L_] * rc[o] * rc[d];                    \
      sarr[d] = sarr[d] + p.zeta[L_] * tc[o] * tc[d] + p.alpha[L_] * rc[o] * rc[d];                    \
      sart[d] = sart[d] + p.alpha[L_] * rc[d] * tc[o] + p.zeta[L_] * rc[o] * tc[d];                    \
    }
    int lstart;                                              // first L whose (L-1, L) pair is initialised
    const double rac3 = sqrt(3.0), x26 = 2. * sqrt(6.0);
    if (is == 0) {                                           // :2106-2119
      for (int j = 0; j < 2; ++j) { pc[j] = 1; rc[j] = 0; tc[j] = 0; }
      ACCUM(0)                                               // L=0: PSL=1, RSL/TSL(0) zero-initialised storage
      for (int j = 0; j < 2; ++j) { pm[j] = pc[j]; rm[j] = 0; tm_[j] = 0; pc[j] = r[j]; rc[j] = 0; tc[j] = 0; }
      if (NS >= 1) { ACCUM(1) }
      for (int j = 0; j < 2; ++j) {
        const double c = r[j];
        pm[j] = c; rm[j] = 0; tm_[j] = 0;
        pc[j] = (3 * c * c - 1) * 0.5; rc[j] = 3 * (1 - c * c) / x26; tc[j] = 0.;
      }
      if (NS >= 2) { ACCUM(2) }
      lstart = 2;
    } else if (is == 1) {                                    // :2123-2135
      for (int j = 0; j < 2; ++j) { const double c = r[j]; const double x = 1 - c * c; pc[j] = sqrt(x * 0.5); rc[j] = 0; tc[j] = 0.; }
      ACCUM(1)
      for (int j = 0; j < 2; ++j) {
        const double c = r[j];
        const double x = 1 - c * c;
        pm[j] = pc[j]; rm[j] = 0; tm_[j] = 0.;
        pc[j] = c * pm[j] * rac3; rc[j] = -c * sqrt(x) * 0.5; tc[j] = -sqrt(x) * 0.5;
      }
      if (NS >= 2) { ACCUM(2) }
      lstart = 2;
    } else {                                                 // :2139-2159
      double a = 1;
      for (int i = 1; i <= is; ++i) { const double x = i; a = a * sqrt((i + is) / x) * 0.5; }
      const double b = a * sqrt(is / (is + 1.0)) * sqrt((is - 1.0) / (is + 2.));
      for (int j = 0; j < 2; ++j) {
        const double c = r[j];
        const double xx = 1 - c * c;
        double yy = (double)(is * 0.5f);
        pm[j] = 0.; rm[j] = 0.; tm_[j] = 0.;
        double x = pow(xx, yy);
        pc[j] = a * x;
        yy = yy - 1;
        x = pow(xx, yy);
        rc[j] = b * (1 + c * c) * x;
        tc[j] = 2 * b * c * x;
      }
      ACCUM(is)
      lstart = is;
    }
    for (int l = lstart; l <= NS - 1; ++l) {                 // recurrence :2167-2189, then term L+1 of the sums
      const double a = (2 * l + 1.) / sqrt((l + is + 1.0) * (l - is + 1.));
      const double b = sqrt((double)((l + is) * (l - is))) / (2. * l + 1.);
      const double d = (l + 1.) * (2 * l + 1.) / sqrt((l + 3.0) * (l - 1.) * (l + is + 1.) * (l - is + 1.));
      const double e = sqrt((l + 2.0) * (l - 2.) * (l + is) * (l - is)) / (l * (2. * l + 1.));
      const double f = (double)__fdiv_rn(2.f * is, l * (l + 1.f));        // all-REAL*4 expression (:2176)
      for (int j = 0; j < 2; ++j) {
        const double c = r[j];
        const double pn = a * (c * pc[j] - b * pm[j]);
        const double rn = d * (c * rc[j] - f * tc[j] - e * rm[j]);
        const double tn = d * (c * tc[j] - f * rc[j] - e * tm_[j]);
        pm[j] = pc[j]; rm[j] = rc[j]; tm_[j] = tc[j];
        pc[j] = pn; rc[j] = rn; tc[j] = tn;
      }
      ACCUM(l + 1)
    }
#undef ACCUM
    for (int d = 0; d < 2; ++d) {
      K[(0 * (NS + 1) + is) * 2 + d] = sbp[d];
      K[(1 * (NS + 1) + is) * 2 + d] = sgr[d];
      K[(2 * (NS + 1) + is) * 2 + d] = sgt[d];
      K[(3 * (NS + 1) + is) * 2 + d] = sarr[d];
      K[(4 * (NS + 1) + is) * 2 + d] = sart[d];
      K[(5 * (NS + 1) + is) * 2 + d] = satt[d];
    }
  }
  __syncthreads();

  // ---------------- SOS_MAT_REFLEXION (:1864-1933) + SOS_MISE_FORMAT (:2378-2395) ----------------
  const double coef = p.coef;
  const size_t NN = (size_t)N * N;
#define KK(w, k, d) K[((w) * (NS + 1) + (k)) * 2 + (d)]
  for (int is = tid; is <= NB; is += nthr) {
    double x = coef * G[is] / 4.;
    double r111 = x * KK(0, 0, 0), r121 = x * KK(1, 0, 0), r122 = x * KK(1, 0, 1), r131 = 0, r132 = 0, r231 = 0, r232 = 0;
    double r211 = x * KK(1, 0, 1), r212 = x * KK(1, 0, 0), r221 = x * KK(3, 0, 1), r222 = x * KK(3, 0, 0);
    double r311 = 0, r312 = 0, r321 = 0, r322 = 0, r331 = x * KK(5, 0, 1), r332 = x * KK(5, 0, 0);
    int im = 1;
    for (int k = 1; k <= NS; ++k) {
      im = -im;
      const int i1 = k + is, i2 = abs(k - is);
      if (i1 > lim && i2 > lim) continue;
      const double xx = coef * im * (G[i1] + G[i2]) / 4.;
      const double yy = coef * im * (G[i2] - G[i1]) / 4.;
      r111 = r111 + KK(0, k, 0) * xx;
      r121 = r121 + KK(1, k, 0) * xx;  r122 = r122 + KK(1, k, 1) * xx;
      r131 = r131 + KK(2, k, 0) * yy;  r132 = r132 + KK(2, k, 1) * yy;
      r211 = r211 + KK(1, k, 1) * xx;  r212 = r212 + KK(1, k, 0) * xx;
      r221 = r221 + KK(3, k, 1) * xx;  r222 = r222 + KK(3, k, 0) * xx;
      r231 = r231 + KK(4, k, 1) * yy;  r232 = r232 + KK(4, k, 0) * yy;
      r311 = r311 + KK(2, k, 1) * yy;  r312 = r312 + KK(2, k, 0) * yy;
      r321 = r321 + KK(4, k, 0) * yy;  r322 = r322 + KK(4, k, 1) * yy;
      r331 = r331 + KK(5, k, 1) * xx;  r332 = r332 + KK(5, k, 0) * xx;
    }
    float *rec = surf + (size_t)is * 9 * NN;
    const float m1[9] = {(float)r111, (float)r121, (float)r131, (float)r211, (float)r221, (float)r231,
                         (float)-r311, (float)-r321, (float)-r331};
    const float m2[9] = {(float)r111, (float)r122, (float)r132, (float)r212, (float)r222, (float)r232,
                         (float)-r312, (float)-r322, (float)-r332};
    for (int m = 0; m < 9; ++m) {
      if (I != J) rec[m * NN + (size_t)(J - 1) * N + (I - 1)] = m1[m];   // P(I,J); for I==J the second store wins (:2378-2379)
      rec[m * NN + (size_t)(I - 1) * N + (J - 1)] = m2[m];               // P(J,I)
    }
  }
#undef KK
}

extern "C" void sos_launch_glitter(GlitterParams p, float *surf, int *il_out, cudaStream_t st)
{
  const int npair = p.nbmu * (p.nbmu + 1) / 2;
  const size_t smem = (size_t)(PH_NU + 1 + p.os_nm + p.os_ns + p.os_nb + 2 + 12 * (p.os_ns + 1) + (p.gmodel == 4 ? p.os_nb + 1 : 0)) * sizeof(double);
  k_glitter<<<npair, 256, smem, st>>>(p, surf, il_out);
}


// =================================================================================================
// SOS_MAT_FRESNEL (SOS_SURFACE.F:1235-1603): Legendre / generalised-Legendre expansion of the Fresnel reflection matrix
// by Gauss quadrature.  One block; thread k owns expansion order k: it runs the polynomial recurrences up to its own order
// for every quadrature angle and accumulates in the reference's angle order (j = -N..N), so every coefficient is the
// reference's own sequence of operations (the recurrences are recomputed per thread instead of shared: O(NS^2 N) flops,
// microseconds).  out: [4][ns+1] = ALPHA, BETA, GAMMA, ZETA before the 4(E15.8) text channel (rounded on the host).
__global__ void k_mat_fresnel(int N, const double *__restrict__ rmu, const double *__restrict__ chr, double ind, int ns,
                              double *__restrict__ out)
{
  extern __shared__ double sh[];
  double *r11 = sh, *r12 = sh + (2 * N + 1), *r33 = sh + 2 * (2 * N + 1);
  double *beta = sh + 3 * (2 * N + 1), *delta = beta + (ns + 1), *gamma = delta + (ns + 1);
  const int W = 2 * N + 1;
  for (int t = threadIdx.x; t < W; t += blockDim.x) {           // :1346-1381
    const int j = t - N;
    if (j == 0) continue;
    double c = rmu[t];
    c = sqrt(.5 * (1 + c));
    const double a = sqrt(ind * ind - 1.0 + c * c), b = ind * ind * c;
    const double rl = -(b - a) / (b + a), rr = (c - a) / (c + a);
    r11[t] = .5 * (rl * rl + rr * rr); r12[t] = .5 * (rl * rl - rr * rr); r33[t] = rl * rr;
  }
  __syncthreads();
  const double sq6 = sqrt(6.0);
  for (int k = threadIdx.x; k <= ns; k += blockDim.x) {
    double bk = 0.0, dk = 0.0, gk = 0.0;
    for (int j = -N; j <= N; ++j) {
      if (j == 0) continue;
      const double xrmu = rmu[j + N];
      // Legendre recurrence (:1387-1400, :1448-1454): PL(K+1) after K steps starting from PL(-1)=0, PL(0)=1
      double p0 = 0.0, p1 = 1.0;
      for (int kk = 0; kk < k; ++kk) {
        const double p2 = ((2 * kk + 1.) * xrmu * p1 - kk * p0) / (kk + 1.);
        p0 = p1; p1 = p2;
      }
      bk = bk + (r11[j + N] * chr[j + N]) * p1;
      dk = dk + (chr[j + N] * r33[j + N]) * p1;
      if (k >= 2) {                                             // generalised function POL(K) (:1433-1447)
        double q0 = 0.0, q1 = 3. * (1. - xrmu * xrmu) / 2. / sq6;   // POL(1), POL(2)
        for (int kk = 2; kk < k; ++kk) {
          const double d = (2. * kk + 1.) / sqrt(1.0 * (kk + 3.) * (kk - 1.));
          const double e = sqrt(1.0 * (kk + 2.) * (kk - 2.)) / (2. * kk + 1.);
          const double q2 = d * (xrmu * q1 - e * q0);
          q0 = q1; q1 = q2;
        }
        gk = gk + (chr[j + N] * r12[j + N]) * q1;
      }
    }
    beta[k] = (2 * k + 1) * bk * .5;                            // :1404
    delta[k] = dk * (2. * k + 1.) * .5;                         // :1460-1463
    gamma[k] = gk * (2. * k + 1.) * .5;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= ns; i += blockDim.x) {         // :1521-1546 (CO1, CO2 are REAL*4 expressions)
    double al = 0.0, ze = 0.0;
    if (i >= 2) {
      const float co1f = 4 * (2 * i + 1.f) / (float)i / (i - 1.f) / (i + 1.f) / (i + 2.f);
      const float co2f = i * (i - 1.f) / ((i + 1.f) * (i + 2.f));
      const double co1 = co1f;
      double co2 = co2f;
      const double co3 = co2 * delta[i];
      co2 = co2 * beta[i];
      const int nn = (int)(i * .5f), mm = (int)((i - 1) * .5f);
      double som1 = 0, som2 = 0, som3 = 0, som4 = 0;
      for (int j = 1; j <= nn; ++j) {
        const double x2 = (double)((i - 1.f) * (i - 1.f) - 3.f * (2 * j - 1.f) * (i - j));
        som1 = som1 + x2 * beta[i - 2 * j]; som2 = som2 + x2 * delta[i - 2 * j];
      }
      for (int j = 0; j <= mm; ++j) {
        const double x2 = (double)((i - 1.f) * (i - 1.f) - 3.f * j * (2 * i - 2 * j - 1.f));
        som3 = som3 + x2 * beta[i - 2 * j - 1]; som4 = som4 + x2 * delta[i - 2 * j - 1];
      }
      ze = co3 - co1 * (som2 - som3);
      al = co2 - co1 * (som1 - som4);
    }
    out[i] = al; out[(ns + 1) + i] = beta[i]; out[2 * (ns + 1) + i] = gamma[i]; out[3 * (ns + 1) + i] = ze;
  }
}

extern "C" void sos_launch_mat_fresnel(int N, const double *rmu, const double *chr, double ind, int ns, double *out, cudaStream_t st)
{
  const size_t smem = (size_t)(3 * (2 * N + 1) + 3 * (ns + 1)) * sizeof(double);
  k_mat_fresnel<<<1, 160, smem, st>>>(N, rmu, chr, ind, ns, out);
}
