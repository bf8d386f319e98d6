// profile_host.cpp -- TEST INFRASTRUCTURE: a host build of the device functions of csrc/profile_chain.cuh, so that the CPU test
// suite can step the profile chain against the reference library where no GPU exists.  The library never runs this: libsosgpu.so
// executes the same functions only inside its CUDA kernels (sosgpu_profile.cu).  Built by tests/test_profile_chain.py with
// g++ -O2 -ffp-contract=off.
#include "../radiativetransfer-sos_b200/csrc/profile_chain.cuh"
#include <string.h>

// The warp-wide search of the kernels (sosgpu_profile.cu, PcWarp), lane by lane on the host: the same pc_tree_* / pc_disc_step /
// pc_first_tau_* functions, with the ballot and the shuffle written as loops.
struct PcWarpEmu {
  double disc(double dt, const PcColumn &c, double tim1, double zmax_init, double zlim) const
  {
    const double ti = tim1 + dt;
    double zmax = zmax_init, zmin = zlim;
    for (;;) {
      unsigned stop = 0, dir = 0;
      double cand[32];
      for (int lane = 1; lane < 32; ++lane) {
        double lo, hi;
        pc_tree_interval(lane, pc_tree_depth(lane), zmin, zmax, &lo, &hi);
        cand[lane] = (hi + lo) / 2.0;
        const int r = pc_disc_step(c, ti, cand[lane]);
        if (r & 1) stop |= 1u << lane;
        if (r & 2) dir |= 1u << lane;
      }
      int node;
      if (pc_tree_walk(stop, dir, &node)) return cand[node];
      pc_tree_interval(node, 5, zmin, zmax, &zmin, &zmax);
    }
  }
  void first(bool gas, const PcColumn &c, double t_first, double *z, double *dtau) const
  {
    if (!(0.0 < t_first)) return;
    double z0 = *z;
    for (;;) {
      double zk[32], dk[32];
      for (int lane = 0; lane < 32; ++lane) {
        double zz = z0;
        for (int k = 0; k <= lane; ++k) zz = zz - PC_DELTA_Z;
        zk[lane] = zz;
        dk[lane] = gas ? pc_first_tau_gas(c, zz) : pc_first_tau_ng(c, zz);
      }
      for (int lane = 0; lane < 32; ++lane)
        if (!(dk[lane] < t_first)) { *z = zk[lane]; *dtau = dk[lane]; return; }
      z0 = zk[31];
    }
  }
};

extern "C" {
int pch_absprofile(int nb_temp, int nb_pres, int nb_conc, const double *tab_temp, const double *tab_pres, const double *tab_conc,
                   const int *nexp, const double *ki, const double *ki_h2o, const double *userprofil, const double *ro, int lamb,
                   const int *ik, double *tauabs)
{
  PcCkd c;
  c.nb_temp = nb_temp; c.nb_pres = nb_pres; c.nb_conc = nb_conc;
  c.tab_temp = tab_temp; c.tab_pres = tab_pres; c.tab_conc = tab_conc; c.nexp = nexp; c.ki = ki; c.ki_h2o = ki_h2o;
  double tau[PC_NLEV];
  for (int j = 1; j <= PC_NLEV - 1; ++j) {
    const int rc = pc_absprofile_layer(c, userprofil, ro, lamb, ik, j, &tau[j - 1]);
    if (rc) return rc;
  }
  pc_absprofile_scan(tau, tauabs);
  return 0;
}
// wide = 0: the reference's serial searches; 1: the kernels' 31-candidate tree search, emulated
int pch_profile(int iprofil, double tr, double hr, double ta, double ha, double zmin, double zmax, int absprofil, const double *altabs,
                const double *tabs, int text_hop, int wide, int *nt, double *zprof, double *h, double *pcaer, double *pcmol)
{
  double scratch[PC_LEVELS];
  memset(scratch, 0, sizeof scratch);
  const int rc = wide ? pc_profile(PcWarpEmu(), iprofil, tr, hr, ta, ha, zmin, zmax, absprofil, altabs, tabs, scratch, zprof, h, pcaer, pcmol, nt)
                      : pc_profile(PcSerial(), iprofil, tr, hr, ta, ha, zmin, zmax, absprofil, altabs, tabs, scratch, zprof, h, pcaer, pcmol, nt);
  if (rc) return rc;
  if (text_hop)
    for (int i = 0; i <= *nt; ++i) { zprof[i] = pc_round_f5(zprof[i]); h[i] = pc_round_e8(h[i]); pcaer[i] = pc_round_e8(pcaer[i]); pcmol[i] = pc_round_e8(pcmol[i]); }
  return 0;
}
double pch_round_e8(double x) { return pc_round_e8(x); }
double pch_round_f5(double x) { return pc_round_f5(x); }
}
