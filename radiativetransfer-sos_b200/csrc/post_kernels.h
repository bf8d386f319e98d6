// post_kernels.h -- descriptors of the azimuth-synthesis kernel (SOS_TRPHI)
#pragma once
#include <cuda_runtime.h>

struct TrphiGroup {          // one aggregated wavelength
  const double *rec;         // [nrec][3][2N+1] Fourier coefficients Q,U,I
  const double *rmu;         // [2N+1]
  int nrec, nbmu, n0;
  int wstride;               // doubles between the Q, U, I rows of a record (>= 2N+1)
  double tau, tauout;        // TTOT_TRONC, TAUOUT after aggregation
};
struct TrphiParams {
  int igli, ifresnel, ipolar;
  double wind, ind_surf, pi;
};
struct GlitterParams {       // SOS_GLITTER (SOS_GLITTER.F:229): one surface file
  int nbmu, os_nb, os_ns, os_nm;
  double sig, coef, pi;      // sigma^2 = .003 + .00512*W (REAL*4 literals), COEF = 1/sigma^2
  const double *rmu;         // [2N+1] device
  const double *alpha, *beta, *gamma, *zeta;   // [os_ns+1] Fresnel expansion (after the E15.8 channel), device
};
#ifdef __cplusplus
extern "C" {
#endif
void sos_launch_glitter(GlitterParams p, float *surf, int *il_out, cudaStream_t st);
void sos_launch_trphi(const TrphiGroup *groups, int ngroup, const double *phis, int nphi,
                      TrphiParams prm, double *out, cudaStream_t st);
void sos_launch_trphi_stride(const TrphiGroup *groups, int ngroup, const double *phis, int nphi, int nout,
                             TrphiParams prm, double *out, cudaStream_t st);
void sos_launch_axpy(double *res, const double *tmp, double aik, size_t n, cudaStream_t st);
#ifdef __cplusplus
}
#endif
