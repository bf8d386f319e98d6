// post_kernels.h -- descriptors of the azimuth-synthesis kernel (SOS_TRPHI)
#pragma once
#include <cuda_runtime.h>

struct TrphiGroup {          // one aggregated wavelength
  const double *rec;         // [nrec][3][2N+1] Fourier coefficients Q,U,I
  const double *rmu;         // [2N+1]
  int nrec, nbmu, n0;
  double tau, tauout;        // TTOT_TRONC, TAUOUT after aggregation
};
struct TrphiParams {
  int igli, ifresnel, ipolar;
  double wind, ind_surf, pi;
};
#ifdef __cplusplus
extern "C" {
#endif
void sos_launch_trphi(const TrphiGroup *groups, int ngroup, const double *phis, int nphi,
                      TrphiParams prm, double *out, cudaStream_t st);
void sos_launch_axpy(double *res, const double *tmp, double aik, size_t n, cudaStream_t st);
#ifdef __cplusplus
}
#endif
