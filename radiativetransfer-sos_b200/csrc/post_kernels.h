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
  // direct-beam terms of the land-surface models (SOS_TRPHI.F:1047-1200); all zero = none
  int iroujean, irondeaux, ibreon, inadal, imaignan;
  double k0, k1, k2;                 // Roujean BRDF
  double alpha_nadal, beta_nadal;    // Nadal BPDF
  double coef_c_maignan;             // Maignan BPDF
};
struct GlitterParams {       // SOS_GLITTER (SOS_GLITTER.F:229): one surface file
  int nbmu, os_nb, os_ns, os_nm;
  int gmodel;                // G function: 0 Cox-Munk (SOS_GSF), 1 Rondeaux, 2 Breon (SOS_GSF_RONDEAUX_BREON, azimuth independent),
                             // 3 Maignan (SOS_GSF_MAIGNAN), 4 Nadal (SOS_F21SF_NADAL)
  double coef_c;             // Maignan's coefficient C
  double ind, alpha_nadal, beta_nadal;   // gmodel 4: Nadal's BPDF over F21 of Fresnel (SOS_F21SF_NADAL)
  int nadal_pairing;         // 0: the series the reference's SOS_MAT_REFLEXION reads for each pair (glitter_kernel.cu); 1: the pair's own
  double sig, coef, pi;      // sigma^2 = .003 + .00512*W (REAL*4 literals), COEF = 1/sigma^2
  const double *rmu;         // [2N+1] device
  const double *alpha, *beta, *gamma, *zeta;   // [os_ns+1] Fresnel expansion (after the E15.8 channel), device
};
#ifdef __cplusplus
extern "C" {
#endif
void sos_launch_glitter(GlitterParams p, float *surf, int *il_out, cudaStream_t st);
// SOS_MAT_FRESNEL: out [4][ns+1] = ALPHA, BETA, GAMMA, ZETA (unrounded); rmu, chr: [2N+1] device
void sos_launch_mat_fresnel(int N, const double *rmu, const double *chr, double ind, int ns, double *out, cudaStream_t st);
void sos_launch_trphi(const TrphiGroup *groups, int ngroup, const double *phis, int nphi,
                      TrphiParams prm, double *out, cudaStream_t st);
void sos_launch_trphi_stride(const TrphiGroup *groups, int ngroup, const double *phis, int nphi, int nout,
                             TrphiParams prm, double *out, cudaStream_t st);
// SOS_ROUJEAN: surf [os_nb+1][9][N][N] REAL*4 (zeroed by the caller), status [N*N] (1: negative BRDF met)
void sos_launch_roujean(int N, const double *rmu, int os_nb, double k0, double k1, double k2, float *surf, int *status,
                        cudaStream_t st);
void sos_launch_ajout_brdf(float *out, const float *a, const float *b, size_t n, cudaStream_t st);
void sos_launch_axpy(double *res, const double *tmp, double aik, size_t n, cudaStream_t st);
#ifdef __cplusplus
}
#endif
