// surface_host.cpp -- TEST INFRASTRUCTURE: a host build of the functions of csrc/nadal_series.cuh, stepped in the way k_glitter
// (gmodel 4) gives them to its threads (samples, one order per "thread", one azimuth per "thread" walking the orders upwards, the
// cut), so that the CPU test suite can compare the Nadal series with the reference library where no GPU exists.  The library
// never runs this: libsosgpu.so executes the same functions only inside its CUDA kernel.  Built by tests/test_surface_nadal.py
// with g++ -O2 -ffp-contract=off.
#include "../radiativetransfer-sos_b200/csrc/nadal_series.cuh"
#include <vector>

extern "C" {
double sfh_nadal_f(double ind, double alpha, double beta, double c1, double c2, double phi) { return nadal_f(ind, alpha, beta, c1, c2, phi); }

// series of the pair (c1, c2): returns IL, fills e[0..nb] (all orders, also those past the cut) and b1[0..nb]
int sfh_nadal_series(double ind, double alpha, double beta, double c1, double c2, int nb, double pi, double *e, double *b1)
{
  const double q = pi / NAD_PH_NU;
  std::vector<double> U(NAD_PH_NU + 1);
  for (int i = 0; i <= NAD_PH_NU; ++i) U[i] = nadal_f(ind, alpha, beta, c1, c2, q * i);
  for (int is = 0; is <= nb; ++is) { e[is] = nadal_coef(U.data(), is, q, pi); b1[is] = 0.0; }
  for (int i = 0; i <= NAD_PH_NU; ++i) {
    const double phi = i * q, f = U[i];
    double t1 = e[0];
    for (int is = 0; is <= nb; ++is) {
      if (is > 0) t1 = nadal_recomb_step(t1, e[is], is, phi);
      const double d = fabs((t1 - f) / f);
      if (d > b1[is]) b1[is] = d;
    }
  }
  return nadal_cut(b1, nb);
}
}
