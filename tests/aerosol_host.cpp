// aerosol_host.cpp -- TEST INFRASTRUCTURE: a host build of the functions of csrc/aerosol_chain.cuh (one "thread", no barrier), so
// that the CPU test suite can step the aerosol chain (Mie -> size distribution -> mixture -> Legendre expansion) against the
// reference library where no GPU exists.  The library never runs this: libsosgpu.so executes the same functions only inside its
// CUDA kernels (sosgpu_aerosols.cu).  Built by tests/test_aerosol_chain.py with g++ -O2 -ffp-contract=off.
#include "../radiativetransfer-sos_b200/csrc/aerosol_chain.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" {
// number of records SOS_MIE writes for [alpha0, alphaf], or -1 (its error 997)
int ach_mie_count(double alpha0, double alphaf)
{
  if (trunc(alphaf + alphaf + 20) > AC_MIE_DIM) return -1;
  int n = 0;
  for (double a = alpha0;; ) { ++n; a = a + ac_mie_step(a); if (!(a <= alphaf)) break; }
  return n;
}
int ach_mie(int nbmu, const double *rmu, double rn, double in, double alpha0, double alphaf, int cap, float *rec, double *g, float *imie,
            float *qmie, float *umie)
{
  const int nrec = ach_mie_count(alpha0, alphaf);
  if (nrec < 0 || nrec > cap) return -1;
  const size_t stride = (size_t)trunc(alphaf + alphaf + 20) + 4, nang = 2 * (size_t)nbmu + 1;
  std::vector<double> buf(AC_WORK_ARRAYS * stride, 0.0);
  const AcMieWork w = ac_work(buf.data(), stride);
  int sh_n[2]; double sh_q[4];
  double a = alpha0;
  for (int k = 0; k < nrec; ++k) {
    ac_mie_record(0, 1, AcNoSync(), a, rn, in, nbmu, rmu, w, sh_n, sh_q, rec + 3 * (size_t)k, g + k, imie + k * nang, qmie + k * nang,
                  umie + k * nang);
    a = a + ac_mie_step(a);
  }
  return nrec;
}
// The two-stage layout of the kernels (one size parameter per lane, interleaved work arrays, chunks that fit an arena budget),
// lane by lane on the host with the same plan, phase functions and offsets as sosgpu_aerosols.cu: tables t = 0..ntab-1 with
// (rn, in, alpha0, alphaf), records concatenated in table order.
int ach_mie_lanes(int nbmu, const double *rmu, int ntab, const double *tab, long long budget_doubles, int cap, float *rec, double *g, float *imie,
                  float *qmie, float *umie)
{
  const size_t nang = 2 * (size_t)nbmu + 1;
  std::vector<long long> rec0; std::vector<int> nrec; std::vector<double> alpha;
  for (int t = 0; t < ntab; ++t) {
    const double a0 = tab[4 * t + 2], af = tab[4 * t + 3];
    const int n = ach_mie_count(a0, af);
    if (n < 0) return -1;
    rec0.push_back((long long)alpha.size()); nrec.push_back(n);
    double a = a0;
    for (int k = 0; k < n; ++k) { alpha.push_back(a); a = a + ac_mie_step(a); }
  }
  if ((int)alpha.size() > cap) return -1;
  const AcMiePlan p = ac_mie_plan(rec0, nrec, alpha, (size_t)budget_doubles);
  std::vector<double> arena(p.arena, 0.0);
  const size_t nitem = p.item_table.size();
  std::vector<int> it_n2(nitem, 0); std::vector<double> it_q(nitem, 0.0);
  const int nw = 4;
  for (size_t c = 0; c + 1 < p.chunk_first.size(); ++c) {
    for (int grp = p.chunk_first[c]; grp < p.chunk_first[c + 1]; ++grp) {          // stage 1: one CTA per group
      AcLaneState st[32];
      for (int ph = 0; ph < AC_COEF_PHASES; ++ph)
        for (int wr = 0; wr < nw; ++wr)
          for (int lane = 0; lane < 32; ++lane) {
            const size_t it = (size_t)grp * 32 + lane;
            const int t = p.item_table[it];
            if (t < 0) continue;
            const size_t r = (size_t)rec0[t] + p.item_rec[it];
            const AcMieWork w = ac_work(arena.data() + p.group_off[grp] + lane, (size_t)p.group_stride[grp], 32);
            ac_coef_phase(ph, wr, nw, alpha[r], tab[4 * t], tab[4 * t + 1], w, st[lane], rec + 3 * r, g + r, &it_n2[it], &it_q[it]);
          }
    }
    for (int grp = p.chunk_first[c]; grp < p.chunk_first[c + 1]; ++grp)            // stage 2: one warp per (group, angle)
      for (size_t j = 0; j < nang; ++j)
        for (int lane = 0; lane < 32; ++lane) {
          const size_t it = (size_t)grp * 32 + lane;
          const int t = p.item_table[it];
          if (t < 0) continue;
          const size_t r = (size_t)rec0[t] + p.item_rec[it];
          const AcMieWork w = ac_work(arena.data() + p.group_off[grp] + lane, (size_t)p.group_stride[grp], 32);
          const AcCoef cf{w.ra, w.ia, w.rb, w.ib, w.es};
          ac_mie_phase(rmu[j], alpha[r], it_q[it], it_n2[it], cf, &imie[r * nang + j], &qmie[r * nang + j], &umie[r * nang + j]);
        }
  }
  return (int)alpha.size();
}
int ach_granu(int nrec, const float *rec, const float *imie, const float *qmie, const float *umie, int nang, double alphaf, int igranu,
              double v1, double v2, double v3, double wa, double *out, double *p11, double *p12, double *p33)
{
  std::vector<double> scratch(3 * (size_t)nrec + 3);
  int sh_k = 0, ier = -2;
  ac_granu(0, 1, AcNoSync(), nrec, rec, imie, qmie, umie, nang, alphaf, igranu, v1, v2, v3, wa, scratch.data(), &sh_k, out, p11, p12, p33, &ier);
  return ier;
}
int ach_model(int nbmu, const double *xmu, const double *xhr, const double *comp_k, const double *p11c, const double *p12c,
              const double *p33c, const double *p22c, int ncomp, const int *comp, const double *w, int itronc, int os_nb, double *scal,
              double *coef, double *phase)
{
  const size_t nang = 2 * (size_t)nbmu + 1;
  std::vector<double> pl((os_nb + 1) * nang), pol((os_nb + 1) * nang);
  for (size_t j = 0; j < nang; ++j) ac_legendre_column(xmu[j], os_nb, pl.data() + j, pol.data() + j, nang);
  AcModel m{};
  m.ncomp = ncomp; m.itronc = itronc; m.os_nb = os_nb;
  for (int i = 0; i < (ncomp ? ncomp : 1); ++i) { m.comp[i] = comp[i]; m.w[i] = w ? w[i] : 1.0; }
  AcModelShared *s = new AcModelShared();
  int ier = -2;
  ac_model(0, 1, AcNoSync(), nbmu, xmu, xhr, pl.data(), pol.data(), comp_k, p11c, p12c, p33c, p22c, m, *s, scal, coef, phase, &ier);
  delete s;
  return ier;
}
}
